"""gym.envs.registration subset: register(id, entry_point) and make("module:Id", **kwargs) (gym_grasper/__init__.py:4-7)."""
import importlib

registry = {}


def register(id, entry_point=None, **kwargs):
    registry[id] = dict(entry_point=entry_point, kwargs=kwargs.get("kwargs", {}))


def make(id, **kwargs):
    if ":" in id:
        module, id = id.split(":", 1)
        importlib.import_module(module)
    if id not in registry:
        raise KeyError(f"No registered env with id: {id}")
    spec = registry[id]
    ep = spec["entry_point"]
    if isinstance(ep, str):
        mod, attr = ep.split(":")
        ep = getattr(importlib.import_module(mod), attr)
    kw = dict(spec["kwargs"])
    kw.update(kwargs)
    return ep(**kw)
