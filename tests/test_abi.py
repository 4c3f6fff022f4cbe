"""The C-ABI shared library loads without a GPU and exports every entry point that include/gops_b200.h
declares; the ctypes prototypes in gops_b200/_lib.py cover exactly that set; struct sizes agree with the
compiled library's view (probed through plan_create's argument validation, no compute)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "gops_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gops_b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from gops_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    handle = C.CDLL(_lib.LIB_PATH)
    names = header_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/gops_b200.h but not exported"
    assert sorted(_lib.PROTOTYPES) == names, "ctypes prototypes and header are out of sync"
    assert _lib.lib().gops_b200_version() == 4


def test_error_reporting_without_gpu_or_with_bad_args():
    from gops_b200 import _lib
    lib = _lib.lib()
    desc = _lib.PlanDesc()
    desc.alg = 99
    handle = C.c_void_p()
    rc = lib.gops_b200_plan_create(C.byref(desc), C.byref(handle))
    assert rc != 0 and b"algorithm" in lib.gops_b200_last_error()
    with pytest.raises(RuntimeError):
        _lib.check(rc)


def test_missing_library_fails_loudly(monkeypatch):
    from gops_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libgops_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.lib()


def test_unknown_ids_raise_like_the_reference():
    from gops_b200.create_pkg.create_alg import create_alg
    from gops_b200.create_pkg.create_apprfunc import create_apprfunc
    from gops_b200.create_pkg.create_env_model import create_env_model
    with pytest.raises(KeyError, match="No registered env with id"):
        create_env_model("no_such_env")
    with pytest.raises(KeyError, match="No registered algorithm with id"):
        create_alg(algorithm="NOPE")
    with pytest.raises(KeyError, match="No registered apprfunc with id"):
        create_apprfunc(apprfunc="MLP", name="Nope")


def test_wrapper_chain_order_matches_reference():
    from gops_b200.create_pkg.create_env_model import create_env_model
    from gops_b200.env.fused import collect_chain
    m = create_env_model("pyth_lq", lq_config="s4a2", reward_scale=0.5)
    names = []
    x = m
    while hasattr(x, "model"):
        names.append(type(x).__name__)
        x = x.model
    assert names == ["ScaleActionModel", "ClipActionModel", "ClipObservationModel", "ShapingRewardModel",
                     "MaskAtDoneModel"] and type(x).__name__ == "LqModel"
    cfg = collect_chain(m)
    assert cfg["reward_scale"] == 0.5 and cfg["mask_at_done"] == 1 and cfg["clip_obs"] == 1
    assert m.action_lower_bound.tolist() == [-1.0, -1.0] and m.unwrapped.action_upper_bound.tolist() == [8.0, 8.0]
    m2 = create_env_model("pyth_lq", lq_config="s4a2", repeat_num=2, obs_scale=[1.0, 2.0, 1.0, 0.5])
    cfg2 = collect_chain(m2)
    assert cfg2["repeat_num"] == 2 and cfg2["obs_scaling"] == 1 and list(cfg2["obs_scale"]) == [1.0, 2.0, 1.0, 0.5]
    assert list(cfg2["obs_shift"]) == [0.0] * 4
