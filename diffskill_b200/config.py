"""Minimal yacs-free configuration layer.

Mirrors the behaviour of the reference's config loader for the hot path's
inputs (plb/config/default_config.py:10-22,78-82 and plb/config/utils.py:4-40):
defaults + YAML merge, attribute access, string tuples such as
``"(0.65, 0.08, 0.5)"`` decoded with ``ast.literal_eval`` (yacs does the same in
``_decode_cfg_value``).  yacs itself is not available offline.
"""
import ast
import copy

import yaml


class CfgNode(dict):
    """dict with attribute access; just enough of yacs.CfgNode for the callers."""

    def __init__(self, init=None, new_allowed=True):
        super().__init__()
        if init:
            for k, v in init.items():
                self[k] = _wrap(v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = _wrap(v)

    def clone(self):
        return copy.deepcopy(self)

    def defrost(self):
        return self

    def freeze(self):
        return self

    def merge_from_other_cfg(self, other):
        _merge(self, other)

    def merge_from_file(self, path):
        with open(path) as f:
            _merge(self, yaml.safe_load(f))

    def merge_from_list(self, opts):
        assert len(opts) % 2 == 0
        for k, v in zip(opts[0::2], opts[1::2]):
            node = self
            parts = k.split('.')
            for p in parts[:-1]:
                node = node[p]
            node[parts[-1]] = _wrap(_decode(v))


def _decode(v):
    if isinstance(v, str):
        try:
            return ast.literal_eval(v)
        except (ValueError, SyntaxError):
            return v
    return v


def _wrap(v):
    if isinstance(v, CfgNode):
        return v
    if isinstance(v, dict):
        return CfgNode(v)
    if isinstance(v, list):
        return [_wrap(i) for i in v]
    return _decode(v)


def _merge(dst, src):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = _wrap(v)


def get_cfg_defaults():
    """Defaults of plb/config/default_config.py (simulator + env parts)."""
    c = CfgNode()
    c.SIMULATOR = CfgNode(dict(
        dim=3, quality=1, quality_multiplier=1., yield_stress=50., dtype="float32", max_steps=1024,
        n_particles=30000, lower_bound=0., E=5e3, nu=0.15, ground_friction=1.5, gravity=(0, -1, 0)))
    c.PRIMITIVES = []
    c.SHAPES = []
    c.RENDERER = CfgNode(dict(name='tina'))
    c.ENV = CfgNode(dict(n_observed_particles=200, cached_state_path='', env_name='',
                         loss=dict(soft_contact=False, target_path='', weight=dict(sdf=10, density=10, contact=1))))
    c.VARIANTS = []
    return c


def load(path=None, opts=None, data=None):
    """``load(path, opts)`` as plb/config/utils.py:33-40; ``data`` accepts an already parsed dict."""
    cfg = get_cfg_defaults()
    if path is not None:
        cfg.merge_from_file(path)
    if data is not None:
        _merge(cfg, copy.deepcopy(data))
    if opts is not None:
        cfg.merge_from_list(opts)
    return cfg


def make_cls_config(obj, cfg=None, **kwargs):
    """plb/config/utils.py:4-13."""
    out = obj.default_config()
    if cfg is not None:
        if isinstance(cfg, str):
            out.merge_from_file(cfg)
        else:
            out.merge_from_other_cfg(cfg)
    if kwargs:
        out.merge_from_list(sum(list(kwargs.items()), ()))
    return out
