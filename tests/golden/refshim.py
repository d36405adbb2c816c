"""Import shim for the UNMODIFIED reference (/root/reference), used ONLY by the golden-vector
generator (tests/golden/make_golden.py) in the build container.  /root/reference does not exist
on the GPU box, so nothing in tests/, bench.py or smoke() imports this at run time.

Stubs habitat / gym / quaternion (absent, no network) and registers pointnav_vo sub-packages as
namespace modules so pointnav_vo/__init__.py (which imports the simulator-bound trainers) is
never executed.  Recipe documented in SURVEY.md section 8c / appendix B.
"""
import collections
import importlib.machinery
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("PNVO_REFERENCE_ROOT", "/root/reference")


class _Any(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        m = _Any(self.__name__ + "." + k)
        setattr(self, k, m)
        return m

    def __call__(self, *a, **k):
        return None


def _stub(name):
    parts = name.split(".")
    for i in range(1, len(parts) + 1):
        n = ".".join(parts[:i])
        if n not in sys.modules:
            m = _Any(n)
            m.__path__ = []
            m.__spec__ = importlib.machinery.ModuleSpec(n, None, is_package=True)
            sys.modules[n] = m
            if i > 1:
                setattr(sys.modules[".".join(parts[: i - 1])], parts[i - 1], m)


class _Registry:
    mapping = collections.defaultdict(dict)

    @classmethod
    def _register_impl(cls, _type, to_register, name, assert_type=None):
        def wrap(c):
            cls.mapping[_type][c.__name__ if name is None else name] = c
            return c

        return wrap if to_register is None else wrap(to_register)

    @classmethod
    def _get_impl(cls, _type, name):
        return cls.mapping[_type].get(name)


class Box:
    def __init__(self, low, high, shape, dtype=np.float32):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype


class Dict:
    def __init__(self, spaces):
        self.spaces = spaces


class Discrete:
    def __init__(self, n):
        self.n = n


_installed = False


def install():
    global _installed
    if _installed:
        return
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    for a, v in (("int", int), ("float", float), ("quaternion", object)):
        if not hasattr(np, a):
            setattr(np, a, v)
    for n in [
        "habitat.core.registry",
        "habitat.core.simulator",
        "habitat.tasks.utils",
        "habitat.tasks.nav.nav",
        "habitat.utils.geometry_utils",
        "habitat.utils.visualizations.utils",
        "habitat.config",
        "gym.spaces",
        "quaternion",
        "habitat_baselines.common.baseline_registry",
    ]:
        _stub(n)
    sys.modules["habitat.core.registry"].Registry = _Registry
    g = sys.modules["gym.spaces"]
    g.Box, g.Dict, g.Discrete = Box, Dict, Discrete
    root = os.path.join(REF_ROOT, "pointnav_vo")
    for sub in [
        "", "vo", "vo.models", "vo.common", "rl", "rl.policies", "rl.common", "rl.ppo",
        "utils", "model_utils", "model_utils.visual_encoders", "model_utils.rnns",
    ]:
        name = "pointnav_vo" + ("." + sub if sub else "")
        m = types.ModuleType(name)
        m.__path__ = [root + ("/" + sub.replace(".", "/") if sub else "")]
        sys.modules[name] = m
    _installed = True
