"""Minimal stand-in for the mmcv registry / config machinery the reference relies on, so
the reference's config files (configs/**.py, `model` dict) build the B200 modules by the
same registry names without mmcv (plugin mechanism: tools/test.py:136-157,
mmdet3d_plugin/__init__.py:1-15)."""
import os

PRECISION = os.environ.get('SRFDET_B200_PRECISION', 'bf16')


def set_precision(p):
    """'bf16' (tcgen05 tensor-core path, tol 1e-2) or 'fp32' (SIMT exact path, tol 1e-4)."""
    global PRECISION
    assert p in ('bf16', 'fp32')
    PRECISION = p


def get_precision():
    return PRECISION


class Registry:
    def __init__(self, name):
        self.name = name
        self._modules = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            key = name if isinstance(name, str) else cls.__name__
            if key in self._modules and not force:
                raise KeyError(f'{key} already registered in {self.name}')
            self._modules[key] = cls
            return cls
        if module is not None:
            return deco(module)
        if isinstance(name, type):
            cls, name = name, None
            return deco(cls)
        return deco

    def get(self, key):
        return self._modules.get(key)

    def build(self, cfg, **default_args):
        cfg = dict(cfg)
        for k, v in default_args.items():
            cfg.setdefault(k, v)
        t = cfg.pop('type')
        cls = t if isinstance(t, type) else self._modules.get(t)
        if cls is None:
            raise KeyError(f"{t} is not in the {self.name} registry")
        return cls(**cfg)

    def __contains__(self, key):
        return key in self._modules


VOXEL_ENCODERS = Registry('voxel_encoder')
MIDDLE_ENCODERS = Registry('middle_encoder')
ROI_EXTRACTORS = Registry('roi_extractor')
HEADS = Registry('head')
NORM_LAYERS = Registry('norm_layer')
DETECTORS = Registry('detector')


def build_voxel_encoder(cfg):
    return VOXEL_ENCODERS.build(cfg)


def build_middle_encoder(cfg):
    return MIDDLE_ENCODERS.build(cfg)


def build_roi_extractor(cfg):
    return ROI_EXTRACTORS.build(cfg)


def build_head(cfg):
    return HEADS.build(cfg)


def build_norm_layer(cfg, num_features):
    """mmcv.cnn.build_norm_layer for the 1-D norms used on this path -> (name, layer)."""
    import torch.nn as nn
    cfg = dict(cfg)
    t = cfg.pop('type')
    cfg.pop('requires_grad', None)
    cls = NORM_LAYERS.get(t)
    if cls is None:
        if t not in ('BN1d', 'BN'):
            raise KeyError(f'unsupported norm layer {t}')
        cls = nn.BatchNorm1d
    return 'bn', cls(num_features, **cfg)


def load_config(path):
    """Execute a reference config file (plain Python, no _base_ inheritance) and return its
    namespace; `cfg['model']` is the model dict."""
    ns = {'__file__': path}
    with open(path) as f:
        exec(compile(f.read(), path, 'exec'), ns)
    return {k: v for k, v in ns.items() if not k.startswith('__')}
