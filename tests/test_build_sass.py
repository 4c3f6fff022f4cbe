"""The built objects must contain the instructions the design claims (checked on the SASS of the sm_100a cubins, no GPU
needed): tcgen05 tensor-core MMAs with TMEM loads / stores and commit barriers in the rollout and inference kernels,
TF32 mma.sync in the mma path, TMA bulk copies for the weight staging."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _count(obj, needles):
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    counts = dict.fromkeys(needles, 0)
    p = subprocess.Popen([exe, "-sass", obj], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    for line in p.stdout:
        for n in needles:
            if n in line:
                counts[n] += 1
    p.wait()
    return counts


def _objects():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.build()
    return os.path.join(ROOT, "build", "obj")


def test_rollout_objects_carry_tcgen05_tma_and_tf32_mma():
    c = _count(os.path.join(_objects(), "kernels_idp.o"),
               ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "HMMA.1688.F32.TF32", "UBLKCP", "SYNCS"])
    assert c["UTCHMMA"] > 100, c          # tcgen05.mma (rollout_tc2: BF16x3 layer / delta / weight-gradient products)
    assert c["LDTM"] > 10 and c["STTM"] >= 4, c  # tcgen05.ld / tcgen05.st (TMEM epilogues; act'(layer 1) parked in TMEM)
    assert c["UTCBAR"] > 8, c             # tcgen05.commit -> mbarrier
    assert c["HMMA.1688.F32.TF32"] > 100, c      # mma.sync 3xTF32 path
    assert c["UBLKCP"] > 4, c             # cp.async.bulk weight staging


def test_layerwise_objects_carry_tcgen05():
    """dense_tc.o: the forward / dgrad / wgrad GEMMs of the layer-wise path; kernels_vehtrack.o only hosts the per-step
    kernels of C3 (its dense products run in dense_tc.o)."""
    c = _count(os.path.join(_objects(), "dense_tc.o"), ["UTCHMMA", "LDTM", "UTCBAR", "UBLKCP"])
    assert c["UTCHMMA"] >= 90 and c["LDTM"] >= 5 and c["UTCBAR"] >= 5 and c["UBLKCP"] >= 4, c


def test_inference_object_carries_tcgen05():
    c = _count(os.path.join(_objects(), "gops_b200.o"), ["UTCHMMA", "LDTM", "UTCBAR", "UBLKCP"])
    assert c["UTCHMMA"] >= 18 and c["LDTM"] >= 2 and c["UTCBAR"] >= 2 and c["UBLKCP"] >= 2, c
