"""Algorithm factory (reference: gops/create_pkg/create_alg.py:33-119): one module per algorithm in
gops_b200/algorithm/, exporting the CamelCase class and `ApproxContainer`.

Data parallelism differs by design: the reference spawns Ray actor replicas for `off_sync` /
`off_async` trainers (create_alg.py:88-93); here every rank of a torchrun job builds ONE in-process
algorithm and the gradient mean is a single NCCL all-reduce inside `local_update`."""
import importlib
import os
from dataclasses import dataclass, field
from typing import Callable, Dict

from gops_b200.utils.gops_path import algorithm_path, underline2camel


@dataclass
class Spec:
    algorithm: str
    entry_point: Callable
    approx_container_cls: Callable
    kwargs: dict = field(default_factory=dict)


registry: Dict[str, Spec] = {}


def register(algorithm: str, entry_point: Callable, approx_container_cls: Callable, **kwargs):
    registry[algorithm] = Spec(algorithm=algorithm, entry_point=entry_point,
                               approx_container_cls=approx_container_cls, kwargs=kwargs)


for _file in sorted(os.listdir(algorithm_path)):
    if _file.endswith(".py") and _file[0] != "_" and _file != "base.py":
        _name = _file[:-3]
        _mdl = importlib.import_module("gops_b200.algorithm." + _name)
        _camel = underline2camel(_name, first_upper=True)
        register(algorithm=_camel, entry_point=getattr(_mdl, _camel),
                 approx_container_cls=getattr(_mdl, "ApproxContainer"))


def _with_defaults(spec_, kwargs):
    _kwargs = spec_.kwargs.copy()
    _kwargs.update(kwargs)
    if _kwargs.get("seed") is None:
        _kwargs["seed"] = 0
    if _kwargs.get("cnn_shared") is None:
        _kwargs["cnn_shared"] = False
    _kwargs.setdefault("trainer", "off_serial_trainer")
    return _kwargs


def create_alg(**kwargs) -> object:
    algorithm = kwargs["algorithm"]
    spec_ = registry.get(algorithm)
    if spec_ is None:
        raise KeyError(f"No registered algorithm with id: {algorithm}")
    if not callable(spec_.entry_point):
        raise RuntimeError(f"{spec_.algorithm} registered but entry_point is not specified")
    return spec_.entry_point(**_with_defaults(spec_, kwargs))


def create_approx_contrainer(algorithm: str, **kwargs) -> object:
    spec_ = registry.get(algorithm)
    if spec_ is None:
        raise KeyError(f"No registered algorithm with id: {algorithm}")
    if not callable(spec_.approx_container_cls):
        raise RuntimeError(f"{spec_.algorithm} registered but approx_container_cls is not specified")
    return spec_.approx_container_cls(**_with_defaults(spec_, kwargs))
