"""marbler_b200: B200-native batched implementation of MARBLER's environment step.

Registers the reference's five gym ids (robotarium_gym/__init__.py:4-23) with gym or gymnasium when
one of them is installed and always in the internal registry used by `marbler_b200.make`."""
from .config import GYM_KEYS, SCENARIOS, default_config_path

__all__ = ["make", "registry", "Wrapper", "SCENARIOS"]

registry = {}
for _scn, _key in GYM_KEYS.items():
    registry[_key] = {"entry_point": "marbler_b200.wrapper:Wrapper",
                      "kwargs": {"env_name": _scn, "config_path": default_config_path(_scn)}}


def _register_with_gym():
    for modname in ("gym", "gymnasium"):
        try:                                    # pragma: no cover - optional dependency
            reg = __import__(modname + ".envs.registration", fromlist=["register"]).register
        except Exception:
            continue
        for key, spec in registry.items():
            try:
                reg(key, entry_point=spec["entry_point"], kwargs=dict(spec["kwargs"]))
            except Exception:
                pass
        return modname
    return None


gym_backend = _register_with_gym()


def make(key, **kwargs):
    """gym.make equivalent: accepts 'PredatorCapturePrey-v0', 'marbler_b200:PredatorCapturePrey-v0' or
    'robotarium_gym:PredatorCapturePrey-v0' (the id EPyMARL configs use, README.md:19-28)."""
    from .wrapper import Wrapper
    key = key.split(":")[-1]
    spec = registry[key]
    kw = dict(spec["kwargs"])
    kw.update(kwargs)
    return Wrapper(**kw)


def __getattr__(name):
    if name == "Wrapper":
        from .wrapper import Wrapper
        return Wrapper
    raise AttributeError(name)
