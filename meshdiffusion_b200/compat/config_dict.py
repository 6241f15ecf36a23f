"""Minimal stand-in for ``ml_collections.ConfigDict`` (the package is not available offline).

Supports what the reference's config files and entry point rely on (configs/default_configs.py, main_diffusion.py:13-16):
attribute and item access, nested dicts, creation of new keys (``lock_config=False`` semantics), dotted overrides.
"""
import ast


class ConfigDict(dict):
    def __init__(self, initial=None):
        super().__init__()
        if initial:
            for k, v in dict(initial).items():
                self[k] = v

    def __setitem__(self, key, value):
        if isinstance(value, dict) and not isinstance(value, ConfigDict):
            value = ConfigDict(value)
        super().__setitem__(key, value)

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value

    def __delattr__(self, name):
        del self[name]

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, ConfigDict) else v) for k, v in self.items()}

    def set_by_path(self, dotted, value):
        node = self
        parts = dotted.split(".")
        for p in parts[:-1]:
            if p not in node or not isinstance(node[p], ConfigDict):
                node[p] = ConfigDict()
            node = node[p]
        node[parts[-1]] = value

    def get_by_path(self, dotted, default=None):
        node = self
        for p in dotted.split("."):
            if not isinstance(node, dict) or p not in node:
                return default
            node = node[p]
        return node


def parse_override_value(text):
    """`--config.a.b=v`: python literal if it parses, else the raw string (ml_collections behaves the same way)."""
    try:
        return ast.literal_eval(text)
    except (ValueError, SyntaxError):
        low = text.lower()
        if low in ("true", "false"):
            return low == "true"
        return text
