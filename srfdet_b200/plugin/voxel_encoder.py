"""Voxel feature encoders of the hot path, under the reference's registry names:
HardSimpleVFE ([3P] mmdet3d; cfg configs/nus/srfdet_voxel_nusc_L.py:40),
DynamicVFECustom (mmdet3d_plugin/models/voxel_encoders/voxel_encoder.py:10-240),
DynamicVFELayer (voxel_encoders/utils.py:8-45) and naiveSyncBN1dCustom
(mmdet3d_plugin/ops/norm.py:27-85; eval path = plain BatchNorm1d, folded here).

State-dict keys match the reference modules, so its checkpoints load unchanged.
Inference only: the CUDA path implements eval-mode forward.
"""
import ctypes

import torch
from torch import nn

from .. import _lib as L
from .ops import DynamicScatter, _ws
from .registry import NORM_LAYERS, VOXEL_ENCODERS, build_norm_layer


@NORM_LAYERS.register_module('naiveSyncBN1dCustom')
class NaiveSyncBatchNorm1dCustom(nn.BatchNorm1d):
    """Eval-mode semantics of ops/norm.py:57-58 (the cross-rank statistics sync at :60-84
    is training-only and outside the hot path)."""

    def forward(self, input):
        assert input.dtype == torch.float32, f'input should be in float32 type, got {input.dtype}'
        if self.training:
            raise NotImplementedError('srfdet_b200 implements the inference path only')
        return super().forward(input)


def fold_bn(weight, bn):
    """(out,in) linear/conv weight + eval BatchNorm -> (folded weight, bias)."""
    s = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    return weight.detach().float() * s.view(-1, *([1] * (weight.dim() - 1))), \
        bn.bias.detach().float() - bn.running_mean.detach().float() * s


@VOXEL_ENCODERS.register_module()
class HardSimpleVFE(nn.Module):
    """Mean of the points of each voxel (M,T,C),(M,) -> (M,num_features)."""

    def __init__(self, num_features=4):
        super().__init__()
        self.num_features = num_features

    def forward(self, features, num_points, coors=None):
        # the fused path (Voxelization.hard_padded(want_mean=True)) never materialises
        # `features`; this standalone form keeps the mmdet3d contract.
        pts = features[:, :, :self.num_features].sum(dim=1)
        return (pts / num_points.type_as(features).view(-1, 1)).contiguous()


class DynamicVFELayer(nn.Module):
    def __init__(self, in_channels, out_channels, norm_cfg=dict(type='BN1d', eps=1e-3, momentum=0.01)):
        super().__init__()
        self.norm = build_norm_layer(norm_cfg, out_channels)[1]
        self.linear = nn.Linear(in_channels, out_channels, bias=False)


@VOXEL_ENCODERS.register_module()
class DynamicVFECustom(nn.Module):
    def __init__(self, in_channels=4, feat_channels=[], with_distance=False, with_cluster_center=False,
                 with_voxel_center=False, voxel_size=(0.2, 0.2, 4), point_cloud_range=(0, -40, -3, 70.4, 40, 1),
                 norm_cfg=dict(type='BN1d', eps=1e-3, momentum=0.01), mode='max', fusion_layer=None,
                 return_point_feats=False, with_centroid_aware_vox=True, centroid_to_point_pos_emb_dims=32):
        super().__init__()
        assert mode in ['avg', 'max'] and len(feat_channels) > 0
        # the fused kernel implements the configuration every reference config uses
        # (configs/waymo/srfdet_dvoxel_waymo_L.py:36-48, configs/kitti/srfdet_voxel_kitti_L.py:40-52)
        if not (with_cluster_center and with_voxel_center and with_centroid_aware_vox) or with_distance \
                or mode != 'max' or fusion_layer is not None or return_point_feats \
                or centroid_to_point_pos_emb_dims != 32 or len(feat_channels) > 2:
            raise NotImplementedError('DynamicVFECustom: only the configuration of the reference configs '
                                      '(cluster centre + voxel centre + centroid-aware, max, <=2 layers) is built')
        self.raw_in_channels = in_channels
        self.in_channels = in_channels + centroid_to_point_pos_emb_dims + 3
        self.vx, self.vy, self.vz = voxel_size
        self.x_offset = self.vx / 2 + point_cloud_range[0]
        self.y_offset = self.vy / 2 + point_cloud_range[1]
        self.z_offset = self.vz / 2 + point_cloud_range[2]
        self.point_cloud_range = point_cloud_range
        self.scatter = DynamicScatter(voxel_size, point_cloud_range, True)
        chans = [self.in_channels] + list(feat_channels)
        layers = []
        for i in range(len(chans) - 1):
            layers.append(DynamicVFELayer(chans[i] * (2 if i > 0 else 1), chans[i + 1], norm_cfg))
        self.vfe_layers = nn.ModuleList(layers)
        self.num_vfe = len(layers)
        self.vfe_scatter = DynamicScatter(voxel_size, point_cloud_range, False)
        self.cluster_scatter = DynamicScatter(voxel_size, point_cloud_range, average_points=True)
        d = centroid_to_point_pos_emb_dims
        self.cen2point_pos_enc = nn.Sequential(nn.Linear(3, d, bias=False), nn.BatchNorm1d(d), nn.Tanh(),
                                               nn.Linear(d, d, bias=False), nn.BatchNorm1d(d), nn.Tanh())
        self._packed = None

    def _weights_version(self):
        """(data_ptr, version) of every parameter / buffer: changes on load_state_dict at any level
        of the module tree and on in-place updates."""
        return tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))

    def _pack(self, device):
        ver = self._weights_version()
        if self._packed is not None and self._packed['dev'] == device and self._packed['ver'] == ver:
            return self._packed
        enc = self.cen2point_pos_enc
        w0, b0 = fold_bn(enc[0].weight, enc[1])
        w1, b1 = fold_bn(enc[3].weight, enc[4])
        v0w, v0b = fold_bn(self.vfe_layers[0].linear.weight, self.vfe_layers[0].norm)
        t = dict(dev=device, ver=ver, w0=w0, b0=b0, w1=w1, b1=b1, v0w=v0w, v0b=v0b, v1w=None, v1b=None)
        if self.num_vfe == 2:
            t['v1w'], t['v1b'] = fold_bn(self.vfe_layers[1].linear.weight, self.vfe_layers[1].norm)
        for k, v in list(t.items()):
            if isinstance(v, torch.Tensor):
                t[k] = v.to(device).contiguous()
        p = L.VfeParams()
        p.pos_w0, p.pos_b0, p.pos_w1, p.pos_b1 = L.ptr(t['w0']), L.ptr(t['b0']), L.ptr(t['w1']), L.ptr(t['b1'])
        p.vfe_w0, p.vfe_b0 = L.ptr(t['v0w']), L.ptr(t['v0b'])
        p.vfe_w1, p.vfe_b1 = L.ptr(t['v1w']), L.ptr(t['v1b'])
        p.cin = self.raw_in_channels
        p.c0 = self.vfe_layers[0].linear.out_features
        p.c1 = self.vfe_layers[1].linear.out_features if self.num_vfe == 2 else 0
        p.vx, p.vy, p.vz = self.vx, self.vy, self.vz
        p.x_off, p.y_off, p.z_off = self.x_offset, self.y_offset, self.z_offset
        t['params'] = p
        self._packed = t
        return t

    def forward_padded(self, features, coors, batch_size=None):
        """No host sync.  -> (voxel_feats (N,C) , voxel_coors (N,4), count (1,) int32 device)."""
        if self.training:
            raise NotImplementedError('srfdet_b200 implements the inference path only')
        features = features.contiguous().float()
        coors = coors.contiguous().int()
        n = features.shape[0]
        dev = features.device
        t = self._pack(dev)
        if batch_size is None:
            batch_size = int(coors[-1, 0].item()) + 1 if n else 1   # voxel_encoder.py:138 syncs too
        dims = [int(batch_size)] + self.scatter.grid_zyx
        ncells = dims[0] * dims[1] * dims[2] * dims[3]
        lib = L.load()
        ws = _ws(lib.srf_dynamic_vfe_ws_bytes(ncells, n), dev)
        c_last = t['params'].c1 or t['params'].c0
        vf = torch.empty((max(n, 1), c_last), dtype=torch.float32, device=dev)
        vc = torch.empty((max(n, 1), 4), dtype=torch.int32, device=dev)
        count = torch.zeros((1,), dtype=torch.int32, device=dev)
        L.check(lib.srf_dynamic_vfe(L.ptr(features), L.ptr(coors), n, L.i4(dims), ctypes.byref(t['params']),
                                    L.ptr(vf), L.ptr(vc), L.ptr(count), L.ptr(ws), ws.numel(), L.stream_ptr()),
                'srf_dynamic_vfe')
        return vf, vc, count

    def forward(self, features, coors, points=None, img_feats=None, img_metas=None):
        vf, vc, count = self.forward_padded(features, coors)
        m = int(count.item())
        return vf[:m], vc[:m]
