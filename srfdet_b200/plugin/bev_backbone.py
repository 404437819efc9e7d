"""Dense BEV backbone and neck between the sparse encoder and the RoI stage:
`SECONDCustom` (mmdet3d_plugin/models/backbones/second_custom.py:11-91) and mmdet's `FPN` as the
reference configs use it (configs/nus/srfdet_voxel_nusc_L.py:55-75: norm BN2d + ReLU ConvModules,
start_level 0, num_outs 4, add_extra_convs='on_output'), under the same registry names, constructor
arguments and state-dict keys (`blocks.{i}.{3j}.weight`, `blocks.{i}.{3j+1}.*`;
`lateral_convs.{i}.conv.weight`, `lateral_convs.{i}.bn.*`, `fpn_convs.{i}.conv.weight`, `fpn_convs.{i}.bn.*`).

HOW: activations are NHWC pixel rows in the mode's 16-bit encoding; every 3x3 convolution (any
stride) runs on the tcgen05 gather-GEMM kernel (`srf_spconv_tc`) with a dense, per-shape static
rulebook (`srf_dense_rulebook`), BatchNorm2d folded into the weights and bias + ReLU fused in its
epilogue; 1x1 lateral convolutions are plain `srf_linear_tc` GEMMs over the pixel rows.  The four FPN
outputs are written fp32 NHWC, i.e. exactly the `torch.channels_last` tensors the RoI samplers and
the DPG kernels read in place.
"""
import ctypes
import os

import torch
from torch import nn

from .. import _lib as L
from . import registry
from .head import _cached, _f, encode_rows
from .registry import BACKBONES, NECKS
from .voxel_encoder import fold_bn


HALO_CONV = os.environ.get('SRF_HALO_CONV', '1') != '0'      # A/B switch: '0' sends every dense conv through the gather-GEMM kernel


def _tc_enc(precision):
    """Backbone convolutions always run on the tensor-core kernels ('fp32_simt' uses the hi + lo split
    form: the FFMA conv kernel stops at 128 channels)."""
    enc = registry.act_enc(precision)
    return registry.act_enc('fp32') if enc is None else enc


class _Grid:
    """Static dense rulebooks, one per (n, h, w, ksize, stride, pad, device)."""
    cache = {}

    @classmethod
    def get(cls, n, h, w, ks, stride, pad, device):
        key = (n, h, w, ks, stride, pad, str(device))
        hit = cls.cache.get(key)
        if hit is None:
            ho, wo = (h + 2 * pad - ks) // stride + 1, (w + 2 * pad - ks) // stride + 1
            cap = (n * ho * wo + 127) // 128 * 128
            nbr = torch.empty((ks * ks, cap), dtype=torch.int32, device=device)
            mask = torch.empty((cap // 128,), dtype=torch.int32, device=device)
            L.check(L.load().srf_dense_rulebook(n, h, w, ks, stride, pad, cap, L.ptr(nbr), L.ptr(mask), L.stream_ptr()),
                    'srf_dense_rulebook')
            hit = cls.cache[key] = (nbr, mask, ho, wo, cap)
        return hit


def conv_bn_act_rows(x_rows, n, h, w, conv, bn, enc, cache, key, relu=True, out_enc=None):
    """x_rows (>= n*h*w, cin | 2cin) pixel rows in `enc` -> (cap, cout) rows of conv(k x k, stride, pad) + folded BN (+ReLU).
    Returns (rows, ho, wo); rows past n*ho*wo are padding."""
    lib = L.load()
    dev = x_rows.device
    ks, stride, pad = conv.kernel_size[0], conv.stride[0], conv.padding[0]
    cin, cout = conv.in_channels, conv.out_channels
    assert conv.kernel_size[0] == conv.kernel_size[1] and conv.groups == 1 and (conv.bias is None or bn is None)
    src = [conv.weight] + ([bn.weight, bn.bias, bn.running_mean, bn.running_var] if bn is not None else []) + (
        [conv.bias] if conv.bias is not None else [])

    def pack():
        if bn is not None:
            wf, bf = fold_bn(conv.weight, bn)
        else:                                   # plain Conv2d (+ its own bias): SRFDetHead.img_convs
            wf = conv.weight.detach().float()
            bf = conv.bias.detach().float() if conv.bias is not None else torch.zeros(cout)
        kio = wf.to(dev).permute(2, 3, 1, 0).reshape(ks * ks, cin, cout).contiguous()       # (cout,cin,ky,kx) -> (ky*ks+kx, cin, cout)
        wp = torch.empty(L.enc_width(enc, kio.numel()), dtype=L.enc_torch_dtype(enc), device=dev)
        L.check(lib.srf_pack_weight_tc(L.ptr(kio), ks * ks, cin, cout, enc, L.ptr(wp), L.stream_ptr()), 'srf_pack_weight_tc')
        return wp, bf.to(dev).contiguous()
    wp, bias = _cached(cache, (key, enc), src, pack)
    out_enc = enc if out_enc is None else out_enc
    if HALO_CONV and ks == 3 and stride == 1 and pad == 1 and not L.enc_is_split(enc) and cin % 128 == 0 and cin <= 512 and cout % 128 == 0:
        # stride-1 3x3 layers with >= 128 channels: halo-tile kernel, no rulebook (csrc/conv3x3_halo.cu)
        cap = (n * h * w + 127) // 128 * 128
        y = torch.empty((cap, L.enc_width(out_enc, cout)), dtype=L.enc_torch_dtype(out_enc), device=dev)
        L.check(lib.srf_conv3x3_rows(L.ptr(x_rows), enc, n, h, w, cin, L.ptr(wp), cout, L.ptr(bias), int(relu), L.ptr(y), out_enc,
                                     L.stream_ptr()), 'srf_conv3x3_rows')
        return y, h, w
    nbr, mask, ho, wo, cap = _Grid.get(n, h, w, ks, stride, pad, dev)
    y = torch.empty((cap, L.enc_width(out_enc, cout)), dtype=L.enc_torch_dtype(out_enc), device=dev)
    a = L.ConvArgs()
    a.in_, a.in_dtype, a.in_rows = L.ptr(x_rows), enc, x_rows.shape[0]
    a.cin, a.cout, a.kvol = cin, cout, ks * ks
    a.nbr, a.tile_mask, a.cap_out, a.d_n_out = L.ptr(nbr), L.ptr(mask), cap, None
    a.w, a.bias, a.residual, a.relu = L.ptr(wp), L.ptr(bias), None, int(relu)
    a.out, a.out_dtype = L.ptr(y), out_enc
    L.check(lib.srf_spconv_tc(ctypes.byref(a), L.stream_ptr()), 'srf_spconv_tc')
    return y, ho, wo


def rows_as_map(rows, n, h, w, c):
    """fp32 pixel rows (>= n*h*w, c) -> logical (n, c, h, w) tensor in torch.channels_last memory format (a view)."""
    return rows[:n * h * w].view(n, h, w, c).permute(0, 3, 1, 2)


@BACKBONES.register_module()
class SECONDCustom(nn.Module):
    def __init__(self, in_channels=128, out_channels=[128, 128, 256], layer_nums=[3, 5, 5], layer_strides=[2, 2, 2],
                 norm_cfg=dict(type='BN', eps=1e-3, momentum=0.01), conv_cfg=dict(type='Conv2d', bias=False), init_cfg=None,
                 pretrained=None):
        super().__init__()
        assert len(layer_strides) == len(layer_nums) == len(out_channels)
        assert conv_cfg.get('type', 'Conv2d') == 'Conv2d' and not conv_cfg.get('bias', False)
        bn = lambda c: nn.BatchNorm2d(c, eps=norm_cfg.get('eps', 1e-5), momentum=norm_cfg.get('momentum', 0.1))
        in_filters = [in_channels, *out_channels[:-1]]
        blocks = []
        for i, layer_num in enumerate(layer_nums):
            block = [nn.Conv2d(in_filters[i], out_channels[i], 3, stride=layer_strides[i], padding=1, bias=False), bn(out_channels[i]),
                     nn.ReLU(inplace=True)]
            for _ in range(layer_num):
                block += [nn.Conv2d(out_channels[i], out_channels[i], 3, padding=1, bias=False), bn(out_channels[i]), nn.ReLU(inplace=True)]
            blocks.append(nn.Sequential(*block))
        self.blocks = nn.ModuleList(blocks)
        self.in_channels = in_channels
        self._cache = {}

    def forward_rows(self, x_rows, n, h, w, precision=None):
        """x_rows: pixel rows of the input map in the mode's encoding -> [(rows, h_i, w_i, c_i)] per block."""
        enc = _tc_enc(precision or registry.get_precision())
        outs = []
        for i, block in enumerate(self.blocks):
            for j in range(0, len(block), 3):
                x_rows, h, w = conv_bn_act_rows(x_rows, n, h, w, block[j], block[j + 1], enc, self._cache, ('b', i, j))
            outs.append((x_rows, h, w, block[0].out_channels))
        return outs

    def forward(self, x, precision=None):
        """x (N, C, H, W) fp32 (the sparse encoder's dense map) -> tuple of (N, C_i, H_i, W_i) fp32 maps
        (torch.channels_last memory format)."""
        if self.training:
            raise NotImplementedError('srfdet_b200 implements the inference path only')
        precision = precision or registry.get_precision()
        enc = _tc_enc(precision)
        n, c, h, w = x.shape
        outs = self.forward_rows(nchw_to_rows(x, enc), n, h, w, precision)
        return tuple(rows_as_map(L.decode(r[:n * hh * ww], cc).contiguous(), n, hh, ww, cc) for r, hh, ww, cc in outs)


def nchw_to_rows(x, enc):
    x = x.contiguous().float()
    n, c, h, w = x.shape
    rows = torch.empty(((n * h * w + 127) // 128 * 128, L.enc_width(enc, c)), dtype=L.enc_torch_dtype(enc), device=x.device)
    L.check(L.load().srf_nchw_to_rows(L.ptr(x), n, c, h, w, enc, L.ptr(rows), L.stream_ptr()), 'srf_nchw_to_rows')
    return rows


class ConvModule(nn.Module):
    """[3P] mmcv ConvModule parameter holder (conv -> bn -> ReLU), keys `conv.weight`, `bn.*`."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, groups=1, norm_cfg=None, bias=False):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding, groups=groups, bias=bias)
        self.bn = nn.BatchNorm2d(out_channels, eps=norm_cfg.get('eps', 1e-5), momentum=norm_cfg.get('momentum', 0.1)) if norm_cfg else None


@NECKS.register_module()
class FPN(nn.Module):
    """[3P] mmdet 2.28.2 FPN in the form the reference configs use: every ConvModule has a norm and ReLU."""

    def __init__(self, in_channels, out_channels, num_outs, start_level=0, end_level=-1, add_extra_convs=False,
                 relu_before_extra_convs=False, no_norm_on_lateral=False, conv_cfg=None, norm_cfg=None, act_cfg=None,
                 upsample_cfg=dict(mode='nearest'), init_cfg=None):
        super().__init__()
        if start_level != 0 or end_level not in (-1, len(in_channels) - 1) or no_norm_on_lateral or norm_cfg is None or act_cfg is None \
                or upsample_cfg.get('mode', 'nearest') != 'nearest' or relu_before_extra_convs \
                or add_extra_convs not in (False, 'on_output', True):
            raise NotImplementedError('FPN: only the configuration of the reference configs is built '
                                      '(all levels, BN + ReLU ConvModules, nearest upsampling, extra convs on_output or max-pool)')
        self.in_channels, self.out_channels, self.num_outs = list(in_channels), out_channels, num_outs
        self.add_extra_convs = 'on_output' if add_extra_convs is True else add_extra_convs   # mmdet maps True -> 'on_input'; reject below
        if add_extra_convs is True:
            raise NotImplementedError("FPN: add_extra_convs=True ('on_input') is not built")
        self.lateral_convs = nn.ModuleList([ConvModule(c, out_channels, 1, norm_cfg=norm_cfg) for c in in_channels])
        self.fpn_convs = nn.ModuleList([ConvModule(out_channels, out_channels, 3, padding=1, norm_cfg=norm_cfg) for _ in in_channels])
        extra = num_outs - len(in_channels)
        if self.add_extra_convs and extra >= 1:
            for _ in range(extra):
                self.fpn_convs.append(ConvModule(out_channels, out_channels, 3, stride=2, padding=1, norm_cfg=norm_cfg))
        self._cache = {}

    def forward_rows(self, feats, n, precision=None):
        """feats: [(rows in the mode's encoding, h, w, c)] -> list of fp32 (n, C, H_l, W_l) channels_last maps."""
        from .head import _LinearView, _linear
        precision = precision or registry.get_precision()
        enc = _tc_enc(precision)
        tc_mode = precision if registry.act_enc(precision) is not None else 'fp32'
        lib = L.load()
        lat = []
        for i, (rows, h, w, c) in enumerate(feats):
            cm = self.lateral_convs[i]
            key = ('lat', i)
            wf_b = _cached(self._cache, (key, 'fold'), [cm.conv.weight, cm.bn.weight, cm.bn.bias, cm.bn.running_mean, cm.bn.running_var],
                           lambda cm=cm: tuple(t.to(rows.device).contiguous() for t in fold_bn(cm.conv.weight.flatten(1), cm.bn)))
            lin = _LinearView(wf_b[0], wf_b[1])
            lat.append((_linear(rows[:n * h * w], lin, tc_mode, self._cache, key, relu=True), h, w))
        for i in range(len(lat) - 1, 0, -1):
            (hi, h, w), (lo, hl, wl) = lat[i - 1], lat[i]
            L.check(lib.srf_upsample_add(L.ptr(hi), L.ptr(lo), n, h, w, hl, wl, self.out_channels, enc, L.stream_ptr()), 'srf_upsample_add')
        outs = []
        for i, (rows, h, w) in enumerate(lat):
            cm = self.fpn_convs[i]
            y, ho, wo = conv_bn_act_rows(rows, n, h, w, cm.conv, cm.bn, enc, self._cache, ('fpn', i), out_enc=L.F32)
            outs.append((y, ho, wo))
        if self.num_outs > len(outs):
            if not self.add_extra_convs:
                for _ in range(self.num_outs - len(outs)):        # mmdet: F.max_pool2d(outs[-1], 1, stride=2)
                    y, h, w = outs[-1]
                    m = rows_as_map(y, n, h, w, self.out_channels)[:, :, ::2, ::2]
                    ho, wo = m.shape[2], m.shape[3]
                    outs.append((m.permute(0, 2, 3, 1).reshape(n * ho * wo, self.out_channels).contiguous(), ho, wo))
            else:
                for i in range(len(lat), self.num_outs):
                    y, h, w = outs[-1]
                    cm = self.fpn_convs[i]
                    x_rows = encode_rows(y[:n * h * w], enc)
                    y2, ho, wo = conv_bn_act_rows(x_rows, n, h, w, cm.conv, cm.bn, enc, self._cache, ('fpn', i), out_enc=L.F32)
                    outs.append((y2, ho, wo))
        return [rows_as_map(y, n, h, w, self.out_channels) for y, h, w in outs]

    def forward(self, inputs, precision=None):
        """inputs: tuple of (N, C_i, H_i, W_i) fp32 maps -> list of num_outs (N, C, H_l, W_l) fp32 maps."""
        precision = precision or registry.get_precision()
        enc = _tc_enc(precision)
        n = inputs[0].shape[0]
        feats = [(nchw_to_rows(x, enc), x.shape[2], x.shape[3], x.shape[1]) for x in inputs]
        return self.forward_rows(feats, n, precision)
