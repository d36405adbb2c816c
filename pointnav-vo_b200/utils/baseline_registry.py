"""String -> class registry with the reference's decorator API (pointnav_vo/utils/baseline_registry.py:26-112,
which subclasses habitat.core.registry.Registry; habitat is not a dependency here).

When the real `pointnav_vo` package is importable, `install_into_reference()` registers these classes
under the same names in ITS registry, which is how the modules drop into the reference's engines /
trainers (they look models up by name: vo_cnn_regression_geo_invariance_engine.py:52-74,
base_trainer_with_vo.py:41-81)."""
import collections


class BaselineRegistry:
    mapping = collections.defaultdict(dict)

    @classmethod
    def _register_impl(cls, _type, to_register, name, assert_type=None):
        def wrap(klass):
            if assert_type is not None:
                assert issubclass(klass, assert_type), f"{klass} must be a subclass of {assert_type}"
            cls.mapping[_type][klass.__name__ if name is None else name] = klass
            return klass

        return wrap if to_register is None else wrap(to_register)

    @classmethod
    def _get_impl(cls, _type, name):
        return cls.mapping[_type].get(name, None)

    @classmethod
    def register_vo_model(cls, to_register=None, *, name=None):
        return cls._register_impl("vo_model", to_register, name)

    @classmethod
    def get_vo_model(cls, name):
        return cls._get_impl("vo_model", name)

    @classmethod
    def register_policy(cls, to_register=None, *, name=None):
        return cls._register_impl("policy", to_register, name)

    @classmethod
    def get_policy(cls, name):
        return cls._get_impl("policy", name)

    @classmethod
    def register_vo_engine(cls, to_register=None, *, name=None):
        return cls._register_impl("vo_engine", to_register, name)

    @classmethod
    def get_vo_engine(cls, name):
        return cls._get_impl("vo_engine", name)


baseline_registry = BaselineRegistry()


def install_into_reference(overwrite=True):
    """Registers every B200 VO model / policy in the reference's own registry (if it is importable)."""
    try:
        from pointnav_vo.utils.baseline_registry import baseline_registry as ref_registry
    except Exception as e:  # reference (or habitat) not importable in this process
        raise RuntimeError(f"the reference package `pointnav_vo` is not importable: {e}")
    for _type in ("vo_model", "policy"):
        for name, klass in BaselineRegistry.mapping[_type].items():
            if overwrite or ref_registry._get_impl(_type, name) is None:
                ref_registry.mapping[_type][name] = klass
    return ref_registry
