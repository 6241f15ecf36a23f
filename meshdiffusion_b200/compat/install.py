"""Makes ``import ml_collections`` resolve to the in-tree stand-in when the real package is absent."""
import sys
import types

from .config_dict import ConfigDict


def ensure_ml_collections():
    try:
        import ml_collections  # noqa: F401
        return
    except ImportError:
        pass
    mod = types.ModuleType("ml_collections")
    mod.ConfigDict = ConfigDict
    sub = types.ModuleType("ml_collections.config_dict")
    sub.ConfigDict = ConfigDict
    mod.config_dict = sub
    sys.modules["ml_collections"] = mod
    sys.modules["ml_collections.config_dict"] = sub
