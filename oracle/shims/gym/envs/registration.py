"""gym.envs.registration stand-in (oracle only)."""
import importlib

registry = {}


def register(id, entry_point=None, kwargs=None, **_ignored):
    registry[id] = (entry_point, dict(kwargs or {}))


def make(id, **kwargs):
    # "module:EnvId-v0" imports the module first (old gym behaviour), then looks the id up.
    if ":" in id:
        mod, id = id.split(":", 1)
        importlib.import_module(mod)
    entry_point, kw = registry[id]
    kw = dict(kw)
    kw.update(kwargs)
    if isinstance(entry_point, str):
        mod_name, attr = entry_point.split(":")
        entry_point = getattr(importlib.import_module(mod_name), attr)
    return entry_point(**kw)
