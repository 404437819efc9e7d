"""Import the reference's own Python (read-only, /root/reference) in this container.

The reference imports mmcv / mmdet / mmdet3d / spconv, none of which is installed.  This
module installs *minimal stand-ins* for exactly the third-party symbols the hot-path files
touch, so that the reference-OWNED code (boxes3d_to_corners3d, the RoI-sampling methods,
DynamicConv, apply_deltas_lidar, DynamicVFECustom.forward, SparseEncoderCustom's layer
construction) executes unmodified on CPU and can emit golden vectors.  The stand-ins for
third-party *kernels* (RoIAlign, DynamicScatter) are supplied by the caller
(make_golden.py uses torchvision / torch.unique), and are recorded as such in the fixture
metadata: goldens pin the reference's code, not mmcv's kernels.

Only make_golden.py uses this; /root/reference does not exist on the GPU box.
"""
import importlib
import sys
import types

import torch
from torch import nn

REF_ROOT = '/root/reference'


class _Registry:
    def __init__(self, name):
        self.name = name
        self.modules = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            self.modules[name or cls.__name__] = cls
            return cls
        if module is not None:
            return deco(module)
        if isinstance(name, type):
            cls, name = name, None
            return deco(cls)
        return deco

    def build(self, cfg, **kw):
        cfg = dict(cfg)
        return self.modules[cfg.pop('type')](**cfg, **kw)


def _passthrough_decorator(*dargs, **dkw):
    if len(dargs) == 1 and callable(dargs[0]) and not dkw:
        return dargs[0]

    def deco(fn):
        return fn
    return deco


class _BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg

    def init_weights(self):
        pass


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []
    sys.modules[name] = m
    parent, _, child = name.rpartition('.')
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], child, m)
    return m


class Recorder:
    """Records make_sparse_convmodule / SparseBasicBlock construction calls."""
    calls = []


def install(dynamic_scatter_cls=None, bbox2roi=None):
    regs = {n: _Registry(n) for n in ['HEADS', 'DETECTORS', 'MIDDLE_ENCODERS', 'VOXEL_ENCODERS',
                                      'BACKBONES', 'NORM_LAYERS', 'ROI_EXTRACTORS']}

    def build_norm_layer(cfg, num_features, postfix=''):
        cfg = dict(cfg)
        t = cfg.pop('type')
        cfg.pop('requires_grad', None)
        if t in regs['NORM_LAYERS'].modules:
            layer = regs['NORM_LAYERS'].modules[t](num_features, **cfg)
        else:
            layer = nn.BatchNorm1d(num_features, **cfg)
        return 'bn', layer

    def build_activation_layer(cfg):
        return nn.ReLU(inplace=cfg.get('inplace', False))

    def make_sparse_convmodule(in_channels, out_channels, kernel_size, indice_key=None, stride=1,
                               padding=0, conv_type='SubMConv3d', norm_cfg=None,
                               order=('conv', 'norm', 'act')):
        Recorder.calls.append(dict(fn='make_sparse_convmodule', cin=in_channels, cout=out_channels,
                                   ksize=kernel_size, key=indice_key, stride=stride, pad=padding,
                                   conv_type=conv_type, order=tuple(order)))
        return nn.Identity()

    class SparseBasicBlock(nn.Module):
        def __init__(self, inplanes, planes, stride=1, downsample=None, conv_cfg=None, norm_cfg=None):
            super().__init__()
            Recorder.calls.append(dict(fn='SparseBasicBlock', cin=inplanes, cout=planes,
                                       conv_type=conv_cfg['type']))

    class SparseSequential(nn.Sequential):
        pass

    _mod('mmcv')
    _mod('mmcv.runner', force_fp32=_passthrough_decorator, auto_fp16=_passthrough_decorator,
         BaseModule=_BaseModule, ModuleList=nn.ModuleList)
    _mod('mmcv.cnn', build_activation_layer=build_activation_layer, ConvModule=None,
         build_conv_layer=None, build_norm_layer=build_norm_layer, NORM_LAYERS=regs['NORM_LAYERS'])
    _mod('mmcv.cnn.bricks')
    _mod('mmcv.cnn.bricks.transformer', build_transformer_layer_sequence=None)
    _mod('mmcv.ops', MultiScaleDeformableAttention=None, points_in_boxes_all=None,
         three_interpolate=None, three_nn=None, SparseConvTensor=None,
         SparseSequential=SparseSequential)
    _mod('mmdet')
    _mod('mmdet.core', build_assigner=None, bbox2roi=bbox2roi, multi_apply=None, build_sampler=None)
    _mod('mmdet.core.utils', reduce_mean=None)
    _mod('mmdet.models')
    _mod('mmdet.models.dense_heads')
    _mod('mmdet.models.dense_heads.base_dense_head', BaseDenseHead=_BaseModule)
    _mod('mmdet.models.losses', sigmoid_focal_loss=None, smooth_l1_loss=None)
    _mod('mmdet3d')
    _mod('mmdet3d.core', box3d_multiclass_nms=None, xywhr2xyxyr=None, bbox3d2result=None)
    _mod('mmdet3d.models', HEADS=regs['HEADS'], build_loss=None, build_head=None,
         build_roi_extractor=None)
    _mod('mmdet3d.models.builder', MIDDLE_ENCODERS=regs['MIDDLE_ENCODERS'],
         VOXEL_ENCODERS=regs['VOXEL_ENCODERS'], build_fusion_layer=None, DETECTORS=regs['DETECTORS'],
         BACKBONES=regs['BACKBONES'])
    _mod('mmdet3d.ops', SparseBasicBlock=SparseBasicBlock, make_sparse_convmodule=make_sparse_convmodule,
         DynamicScatter=dynamic_scatter_cls, Voxelization=None)
    _mod('mmdet3d.ops.spconv', IS_SPCONV2_AVAILABLE=False)

    # reference package skeleton: packages exist but their __init__ files are NOT executed
    # (they import datasets / visualisers that need yet more third-party code).
    for pkg in ['mmdet3d_plugin', 'mmdet3d_plugin.models', 'mmdet3d_plugin.core',
                'mmdet3d_plugin.core.bbox', 'mmdet3d_plugin.models.sparse_heads',
                'mmdet3d_plugin.models.voxel_encoders', 'mmdet3d_plugin.models.middle_encoders',
                'mmdet3d_plugin.ops']:
        m = _mod(pkg)
        m.__path__ = [REF_ROOT + '/' + pkg.replace('.', '/')]
    # the reference hard-codes .cuda() (core/bbox/util.py:134,143-145): run it on CPU
    torch.Tensor.cuda = lambda self, *a, **k: self
    return regs


def ref_import(name):
    return importlib.import_module(name)
