"""Minimal behavioural stand-in for yacs.config.CfgNode (test infrastructure only).

Used solely to import the reference's lib/config/default.py when generating golden vectors
(oracle/ref_harness.py).  Dict with attribute access, recursive wrapping, yaml merge, list merge;
freeze/defrost are no-ops.
"""
import ast
import yaml


class CfgNode(dict):
    def __init__(self, init_dict=None, key_list=None, new_allowed=False):
        super().__init__()
        for k, v in (init_dict or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value

    def defrost(self):
        pass

    def freeze(self):
        pass

    def clone(self):
        import copy
        return copy.deepcopy(self)

    def _merge(self, other):
        for k, v in other.items():
            if isinstance(v, dict):
                if k in self and isinstance(self[k], CfgNode):
                    self[k]._merge(v)
                else:
                    self[k] = CfgNode(v)
            else:
                if isinstance(v, tuple):
                    v = list(v)
                self[k] = v

    def merge_from_file(self, path):
        with open(path, "r") as f:
            self._merge(yaml.safe_load(f) or {})

    def merge_from_other_cfg(self, other):
        self._merge(other)

    def merge_from_list(self, lst):
        assert len(lst) % 2 == 0
        for full_key, v in zip(lst[0::2], lst[1::2]):
            d = self
            keys = full_key.split(".")
            for sub in keys[:-1]:
                d = d[sub]
            if isinstance(v, str):
                try:
                    v = ast.literal_eval(v)
                except (ValueError, SyntaxError):
                    pass
            d[keys[-1]] = v
