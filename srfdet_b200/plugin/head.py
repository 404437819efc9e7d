"""Sparse-region-fusion head stages under the reference names
(mmdet3d_plugin/models/sparse_heads/srfdet_head.py): DynamicConv (:2633-2693),
SingleSRFDetHeadLiDAR (:1348-1623) and SingleSRFDetHead (:2103-2629).

Hot path on the CUDA kernels of this package: RoI sampling,
fusion projection, DynamicConv.  The remaining dense rows of a stage (self-attention over
the proposals, FFN, cls/reg towers, apply_deltas) run as plain torch ops
so the five stages chain with real boxes; they are not part of the measured path.
Parameter names equal the reference's, so its checkpoints load unchanged.
"""
import math

import torch
import torch.nn.functional as F
from torch import nn

from .. import _lib as L
from . import registry
from .registry import HEADS
from .roi import img_feats_sampling_bboxes_roi, maps_channels_last, points_feats_sampling_bboxes_roi

_DEFAULT_SCALE_CLAMP = math.log(100000.0 / 16)


def _linear(x, lin, precision, cache, key, relu=False, ln=None, out_bf16=None):
    """x (M,K) f32|bf16 -> (M,N).  fp32: SIMT FFMA.  bf16: tcgen05 GEMM (weights packed once)."""
    lib = L.load()
    m, k = x.shape
    n = lin.out_features
    dev = x.device
    st = L.stream_ptr()
    bias = lin.bias.detach().float().contiguous() if lin.bias is not None else None
    if precision == 'fp32':
        x = x.float().contiguous()
        w = lin.weight.detach().float().contiguous()
        out = torch.empty((m, n), dtype=torch.float32, device=dev)
        fuse_relu = relu and ln is None
        L.check(lib.srf_linear_f32(L.ptr(x), m, k, L.ptr(w), n, L.ptr(bias), int(fuse_relu), L.ptr(out), st), 'srf_linear_f32')
        if ln is not None:
            L.check(lib.srf_layernorm(L.ptr(out), L.F32, m, n, 1, None, L.ptr(ln.weight.detach().float().contiguous()),
                                      L.ptr(ln.bias.detach().float().contiguous()), ln.eps, int(relu), L.ptr(out), st),
                    'srf_layernorm')
        return out
    if x.dtype != torch.bfloat16:
        xb = torch.empty((m, k), dtype=torch.bfloat16, device=dev)
        L.check(lib.srf_f32_to_bf16(L.ptr(x.float().contiguous()), m, k, k, L.ptr(xb), st), 'srf_f32_to_bf16')
        x = xb
    if key not in cache:
        wp = torch.empty((n * k,), dtype=torch.bfloat16, device=dev)
        L.check(lib.srf_pack_linear_bf16(L.ptr(lin.weight.detach().float().contiguous()), n, k, L.ptr(wp), st),
                'srf_pack_linear_bf16')
        cache[key] = wp
    out_bf16 = True if out_bf16 is None else out_bf16
    # few output tiles but a long reduction (DynamicConv.out_layer: 900 x 6272 -> 128): split K
    # over CTAs, fp32 partials meet in a zeroed buffer, bias + LayerNorm + ReLU in one pass after
    tiles = ((m + 127) // 128) * ((n + 127) // 128)
    kvol = k // min(k, 128)
    if ln is not None and tiles * 4 <= 148 and kvol >= 8:
        splits = lib.srf_linear_splits(k, min(kvol, max(1, 148 // tiles)))
        part = torch.empty((splits, m, n), dtype=torch.float32, device=dev)
        L.check(lib.srf_linear_bf16(L.ptr(x.contiguous()), m, k, L.ptr(cache[key]), n, None, 0, None, None, L.ptr(part), L.F32,
                                    splits, st), 'srf_linear_bf16')
        acc = torch.empty((m, n), dtype=torch.float32, device=dev)
        L.check(lib.srf_layernorm(L.ptr(part), L.F32, m, n, splits, L.ptr(bias), L.ptr(ln.weight.detach().float().contiguous()),
                                  L.ptr(ln.bias.detach().float().contiguous()), ln.eps, int(relu), L.ptr(acc), st), 'srf_layernorm')
        return acc.to(torch.bfloat16) if out_bf16 else acc
    out = torch.empty((m, n), dtype=torch.bfloat16 if out_bf16 else torch.float32, device=dev)
    fuse_ln = ln is not None and n <= 128
    epi = (1 if relu and (ln is None or fuse_ln) else 0) | (2 if fuse_ln else 0)
    lnw = ln.weight.detach().float().contiguous() if fuse_ln else None
    lnb = ln.bias.detach().float().contiguous() if fuse_ln else None
    L.check(lib.srf_linear_bf16(L.ptr(x.contiguous()), m, k, L.ptr(cache[key]), n, L.ptr(bias), epi, L.ptr(lnw), L.ptr(lnb),
                                L.ptr(out), L.BF16 if out_bf16 else L.F32, 1, st), 'srf_linear_bf16')
    if ln is not None and not fuse_ln:
        L.check(lib.srf_layernorm(L.ptr(out), L.BF16 if out_bf16 else L.F32, m, n, 1, None,
                                  L.ptr(ln.weight.detach().float().contiguous()), L.ptr(ln.bias.detach().float().contiguous()),
                                  ln.eps, int(relu), L.ptr(out), st), 'srf_layernorm')
    return out


class DynamicConv(nn.Module):
    def __init__(self, feat_channels, dynamic_dim=64, dynamic_num=2, pooler_resolution=7):
        super().__init__()
        assert dynamic_num == 2 and pooler_resolution == 7
        self.feat_channels = feat_channels
        self.dynamic_dim = dynamic_dim
        self.dynamic_num = dynamic_num
        self.num_params = feat_channels * dynamic_dim
        self.dynamic_layer = nn.Linear(feat_channels, dynamic_num * self.num_params)
        self.norm1 = nn.LayerNorm(dynamic_dim)
        self.norm2 = nn.LayerNorm(feat_channels)
        self.activation = nn.ReLU(inplace=True)
        self.out_layer = nn.Linear(feat_channels * pooler_resolution ** 2, feat_channels)
        self.norm3 = nn.LayerNorm(feat_channels)
        self._cache = {}

    def load_state_dict(self, *a, **k):
        self._cache = {}
        return super().load_state_dict(*a, **k)

    def make_params(self, prop_feats, precision=None):
        """dynamic_layer(prop_feats): (K,C) -> (K, 2*C*d).  Depends only on the proposal features,
        so callers may run it concurrently with the RoI sampling of the same stage."""
        precision = precision or registry.get_precision()
        return _linear(prop_feats, self.dynamic_layer, precision, self._cache, ('dyn', str(prop_feats.device)))

    def forward_kc(self, prop_feats, roi_feats, precision=None, params=None):
        """prop_feats (K,C); roi_feats (K,49,C) f32|bf16 (channel-last RoI features) -> (K,C) f32."""
        precision = precision or registry.get_precision()
        lib = L.load()
        k, c = prop_feats.shape
        d = self.dynamic_dim
        dev = prop_feats.device
        bf = precision == 'bf16'
        if params is None:
            params = self.make_params(prop_feats, precision)
        roi_feats = roi_feats.contiguous()
        inter = torch.empty((k, 49 * c), dtype=torch.bfloat16 if bf else torch.float32, device=dev)
        f = lambda t: L.ptr(t.detach().float().contiguous())
        L.check(lib.srf_dynconv_interact(L.ptr(roi_feats), L.BF16 if roi_feats.dtype == torch.bfloat16 else L.F32,
                                         L.ptr(params), L.BF16 if bf else L.F32, k, c, d, f(self.norm1.weight),
                                         f(self.norm1.bias), f(self.norm2.weight), f(self.norm2.bias), L.ptr(inter),
                                         L.BF16 if bf else L.F32, L.stream_ptr()), 'srf_dynconv_interact')
        out = _linear(inter, self.out_layer, precision, self._cache, ('out', str(dev)), relu=True, ln=self.norm3,
                      out_bf16=False)
        return out.float()

    def forward(self, prop_feats, roi_feats):
        """Reference signature: prop_feats (1, K, C), roi_feats (49, K, C) -> (K, C)."""
        return self.forward_kc(prop_feats[0], roi_feats.permute(1, 0, 2).contiguous())


class _SingleHeadBase(nn.Module):
    def _build_common(self, num_classes, feat_channels, pooler_resolution, use_focal_loss, use_fed_loss,
                      dim_feedforward, num_cls_convs, num_reg_convs, num_heads, dropout, scale_clamp, bbox_weights,
                      dynamic_conv, pc_range, voxel_size):
        self.feat_channels_lidar = feat_channels
        self.pc_range_lidar = pc_range
        self.voxel_size_lidar = voxel_size
        self.self_attn_lidar = nn.MultiheadAttention(feat_channels, num_heads, dropout=dropout)
        self.inst_interact_lidar = DynamicConv(feat_channels=feat_channels, pooler_resolution=pooler_resolution,
                                               dynamic_dim=dynamic_conv['dynamic_dim'],
                                               dynamic_num=dynamic_conv['dynamic_num'])
        self.linear1_lidar = nn.Linear(feat_channels, dim_feedforward)
        self.dropout_lidar = nn.Dropout(dropout)
        self.linear2_lidar = nn.Linear(dim_feedforward, feat_channels)
        self.norm1_lidar = nn.LayerNorm(feat_channels)
        self.norm2_lidar = nn.LayerNorm(feat_channels)
        self.norm3_lidar = nn.LayerNorm(feat_channels)
        self.dropout1_lidar = nn.Dropout(dropout)
        self.dropout2_lidar = nn.Dropout(dropout)
        self.dropout3_lidar = nn.Dropout(dropout)
        self.activation_lidar = nn.ReLU(inplace=True)

        def tower(n):
            mods = []
            for _ in range(n):
                mods += [nn.Linear(feat_channels, feat_channels, False), nn.LayerNorm(feat_channels), nn.ReLU(inplace=True)]
            return nn.ModuleList(mods)
        self.cls_module_lidar = tower(num_cls_convs)
        self.reg_module_lidar = tower(num_reg_convs)
        self.use_focal_loss = use_focal_loss
        self.use_fed_loss = use_fed_loss
        self.class_logits_lidar = nn.Linear(feat_channels, num_classes if (use_focal_loss or use_fed_loss) else num_classes + 1)
        self.bboxes_delta_lidar = nn.Linear(feat_channels, len(bbox_weights))
        self.scale_clamp = scale_clamp
        self.bbox_weights = bbox_weights

    # dense tail of a stage ('next' rows): attention, interaction, FFN, towers, box update
    def _stage_tail(self, roi_feats_kc, bboxes, prop_feats, bs, n_p, precision=None):
        C = self.feat_channels_lidar
        prop = prop_feats.view(bs, n_p, C).permute(1, 0, 2)
        prop2 = self.self_attn_lidar(prop, prop, value=prop)[0]
        prop = self.norm1_lidar(prop + self.dropout1_lidar(prop2))
        prop = prop.permute(1, 0, 2).reshape(bs * n_p, C)
        prop2 = self.inst_interact_lidar.forward_kc(prop.contiguous(), roi_feats_kc, precision)
        obj = self.norm2_lidar(prop + self.dropout2_lidar(prop2))
        obj2 = self.linear2_lidar(self.dropout_lidar(self.activation_lidar(self.linear1_lidar(obj))))
        obj = self.norm3_lidar(obj + self.dropout3_lidar(obj2))
        cls_f, reg_f = obj.clone(), obj.clone()
        for layer in self.cls_module_lidar:
            cls_f = layer(cls_f)
        for layer in self.reg_module_lidar:
            reg_f = layer(reg_f)
        logits = self.class_logits_lidar(cls_f)
        deltas = self.bboxes_delta_lidar(reg_f)
        pred = self.apply_deltas_lidar(deltas, bboxes.view(-1, len(self.bbox_weights)))
        return logits.view(bs, n_p, -1), pred.view(bs, n_p, -1), obj

    def apply_deltas_lidar(self, deltas, boxes):
        """srfdet_head.py:2331-2420 (boxes carry ABSOLUTE centres after the in-place de-normalisation)."""
        boxes = boxes.to(deltas.dtype)
        w = self.bbox_weights
        d = deltas
        ctr = boxes[:, 0:3]
        size = torch.exp(boxes[:, 3:6])
        dxyz = torch.stack([d[:, 0] / w[0], d[:, 1] / w[1], d[:, 2] / w[2]], -1)
        dwlh = torch.stack([d[:, 3] / w[3], d[:, 4] / w[4], d[:, 5] / w[5]], -1).clamp(max=self.scale_clamp)
        pred_ctr = dxyz * size + ctr
        pred_size = torch.exp(dwlh) * size
        r = self.pc_range_lidar
        lo = pred_ctr.new_tensor(r[:3])
        span = pred_ctr.new_tensor([r[3] - r[0], r[4] - r[1], r[5] - r[2]])
        pred_ctr = ((pred_ctr - lo) / span).clamp(min=0.0, max=1.0)
        return torch.cat([pred_ctr, pred_size.log(), d[:, 6:len(w)]], dim=-1)


@HEADS.register_module()
class SingleSRFDetHeadLiDAR(_SingleHeadBase):
    def __init__(self, num_classes=80, feat_channels=256, pooler_resolution=7, use_focal_loss=True, use_fed_loss=False,
                 dim_feedforward=2048, num_cls_convs=1, num_reg_convs=3, num_heads=8, dropout=0.0,
                 scale_clamp=_DEFAULT_SCALE_CLAMP, bbox_weights=[1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 0.2, 0.2],
                 act_cfg=dict(type='ReLU', inplace=True), dynamic_conv=dict(dynamic_dim=64, dynamic_num=2),
                 pc_range=None, voxel_size=None, init_cfg=None, is_kitti=None):
        super().__init__()
        self._build_common(num_classes, feat_channels, pooler_resolution, use_focal_loss, use_fed_loss, dim_feedforward,
                           num_cls_convs, num_reg_convs, num_heads, dropout, scale_clamp, bbox_weights, dynamic_conv,
                           pc_range, voxel_size)

    def region_features(self, point_feats, bboxes, pooler):
        """RoI features, channel-last (bs*n_p, 49, C).  Mutates bboxes[..., :3] (reference :1638-1646)."""
        return points_feats_sampling_bboxes_roi(point_feats, bboxes, pooler, self.pc_range_lidar, self.voxel_size_lidar,
                                                channel_last=True)

    def forward(self, point_feats, bboxes, prop_feats, pooler, img_metas=None, precision=None):
        bs, n_p = bboxes.shape[:2]
        roi = self.region_features(point_feats, bboxes, pooler)
        if prop_feats is None:
            prop_feats = roi.mean(1).view(bs, n_p, -1)
        return self._stage_tail(roi, bboxes, prop_feats, bs, n_p, precision)


@HEADS.register_module()
class SingleSRFDetHead(_SingleHeadBase):
    def __init__(self, num_classes=80, feat_channels=256, pooler_resolution=7, use_focal_loss=True, use_fed_loss=False,
                 dim_feedforward=2048, num_cls_convs=1, num_reg_convs=3, num_heads=8, dropout=0.0,
                 scale_clamp=_DEFAULT_SCALE_CLAMP, bbox_weights=[1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 0.2, 0.2],
                 act_cfg=dict(type='ReLU', inplace=True), dynamic_conv=dict(dynamic_dim=64, dynamic_num=2),
                 pc_range=None, use_fusion=False, voxel_size=None, is_kitti=False, init_cfg=None):
        super().__init__()
        self.feat_channels = feat_channels
        self.use_fusion = use_fusion
        self.is_kitti = is_kitti
        self._build_common(num_classes, feat_channels, pooler_resolution, use_focal_loss, use_fed_loss, dim_feedforward,
                           num_cls_convs, num_reg_convs, num_heads, dropout, scale_clamp, bbox_weights, dynamic_conv,
                           pc_range, voxel_size)
        if use_fusion:
            self.output_fused_proj = nn.Linear(2 * feat_channels, feat_channels)
        self._cache = {}

    def load_state_dict(self, *a, **k):
        self._cache = {}
        return super().load_state_dict(*a, **k)

    def region_features(self, img_feats, point_feats, bboxes, pooler, pooler_img, lidar2img, precision=None):
        """Fused RoI features (bs*n_p, 49, C) channel-last: image RoIs (camera sum), BEV RoIs,
        concat + Linear(2C->C) (srfdet_head.py:2236-2264)."""
        precision = precision or registry.get_precision()
        img_roi = pts_roi = None
        if (self.use_fusion and img_feats is not None and point_feats is not None
                and maps_channels_last([f[0] for f in img_feats], pooler_img.num_inputs)
                and maps_channels_last(point_feats, pooler.num_inputs)):
            # both samplers write their half of cat(img, pts) (srfdet_head.py:2257) in the GEMM's dtype
            k, c = bboxes.shape[0] * bboxes.shape[1], point_feats[0].shape[1]
            cat = torch.empty((k, 49, 2 * c), device=bboxes.device,
                              dtype=torch.bfloat16 if precision == 'bf16' else torch.float32)
            img_feats_sampling_bboxes_roi(img_feats, bboxes, pooler_img, lidar2img, self.pc_range_lidar,
                                          channel_last=True, out=cat, ch_offset=0)
            points_feats_sampling_bboxes_roi(point_feats, bboxes, pooler, self.pc_range_lidar, self.voxel_size_lidar,
                                             channel_last=True, out=cat, ch_offset=c)
            fused = _linear(cat.view(k * 49, 2 * c), self.output_fused_proj, precision, self._cache, ('fuse', str(cat.device)))
            return fused.view(k, 49, -1)
        if img_feats is not None:
            img_roi = img_feats_sampling_bboxes_roi(img_feats, bboxes, pooler_img, lidar2img, self.pc_range_lidar,
                                                    channel_last=True)
        if point_feats is not None:
            pts_roi = points_feats_sampling_bboxes_roi(point_feats, bboxes, pooler, self.pc_range_lidar,
                                                       self.voxel_size_lidar, channel_last=True)
        if img_roi is not None and pts_roi is not None and self.use_fusion:
            k = pts_roi.shape[0]
            cat = torch.cat((img_roi, pts_roi), dim=2).view(k * 49, -1)
            fused = _linear(cat, self.output_fused_proj, precision, self._cache, ('fuse', str(cat.device)))
            return fused.view(k, 49, -1)
        if not self.use_fusion and img_roi is not None and pts_roi is None:
            return img_roi
        if not self.use_fusion and pts_roi is not None and img_roi is None:
            return pts_roi
        raise ValueError('inconsistent fusion inputs')

    def forward(self, img_feats, point_feats, bboxes, prop_feats, pooler, img_metas, pooler_img=None, precision=None):
        bs, n_p = bboxes.shape[:2]
        lidar2img = None
        if img_feats is not None:
            import numpy as np
            lidar2img = torch.as_tensor(np.asarray([m['lidar2img'] for m in img_metas]), dtype=torch.float32,
                                        device=bboxes.device)
        roi = self.region_features(img_feats, point_feats, bboxes, pooler, pooler_img, lidar2img, precision)
        if prop_feats is None:
            prop_feats = roi.float().mean(1).view(bs, n_p, -1)
        return self._stage_tail(roi, bboxes, prop_feats, bs, n_p, precision)
