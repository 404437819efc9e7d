"""Point-cloud side of the SRFDet detector (mmdet3d_plugin/models/detectors/srfdet.py):
`voxelize` (:204-247) and `extract_point_features` (:249-276) up to the middle encoder,
built from the reference's unmodified `model` config dict.

The dense BEV backbone / FPN that follow in the reference (SECONDCustom + FPN,
srfdet.py:260-262) are outside the hot path and are not built here.
"""
import torch
from torch import nn

from . import registry
from .ops import Voxelization
from .registry import DETECTORS, build_middle_encoder, build_voxel_encoder
from .voxel_encoder import HardSimpleVFE


@DETECTORS.register_module(name='SRFDet')
@DETECTORS.register_module(name='SRFDetWaymo')
class SRFDetPointPath(nn.Module):
    def __init__(self, pts_voxel_layer=None, pts_voxel_encoder=None, pts_middle_encoder=None, **unused):
        super().__init__()
        self.pts_voxel_layer_cfg = dict(pts_voxel_layer)
        self.pts_voxel_layer = Voxelization(**pts_voxel_layer)
        self.pts_voxel_encoder = build_voxel_encoder(pts_voxel_encoder)
        self.pts_middle_encoder = build_middle_encoder(pts_middle_encoder)
        self.unused_cfg = unused
        self.eval()

    @classmethod
    def from_config(cls, path_or_model):
        model = registry.load_config(path_or_model)['model'] if isinstance(path_or_model, str) else dict(path_or_model)
        model = dict(model)
        model.pop('type', None)
        enc = dict(model['pts_middle_encoder'])
        enc.pop('init_cfg', None)   # checkpoints are loaded explicitly (no network / ckpts here)
        model['pts_middle_encoder'] = enc
        return cls(**model)

    @property
    def is_dynamic(self):
        return self.pts_voxel_layer_cfg['max_num_points'] == -1

    @torch.no_grad()
    def voxelize(self, points):
        """Reference contract (host-synchronising, like srfdet.py:204-247).
        hard   : (voxels (M,T,C), num_points (M,), coors (M,4))
        dynamic: (points (sum N, C), coors (sum N, 4))"""
        if not self.is_dynamic:
            vs, cs, ns = [], [], []
            for i, res in enumerate(points):
                o = self.pts_voxel_layer.hard_padded(res, batch_idx=i)
                m = int(o['count'].item())
                vs.append(o['voxels'][:m])
                cs.append(o['coors'][:m])
                ns.append(o['num_points'][:m])
            return torch.cat(vs, 0), torch.cat(ns, 0), torch.cat(cs, 0)
        coors = [self.pts_voxel_layer.dynamic(res, batch_idx=i) for i, res in enumerate(points)]
        return torch.cat(points, 0), torch.cat(coors, 0)

    @torch.no_grad()
    def extract_point_features(self, points, precision=None):
        """points: list of (N_i, C) fp32 CUDA tensors -> dense BEV (B, C*D, H, W) fp32.
        Single-frame inputs (all reference test configs use batch 1) run without any host
        synchronisation: voxel counts stay on the device end to end."""
        bs = len(points)
        if not self.is_dynamic:
            if bs == 1 and isinstance(self.pts_voxel_encoder, HardSimpleVFE):
                o = self.pts_voxel_layer.hard_padded(points[0], batch_idx=0, want_voxels=False, want_mean=True)
                c = self.pts_voxel_encoder.num_features
                mean = o['mean'] if o['mean'].shape[1] == c else o['mean'][:, :c].contiguous()
                return self.pts_middle_encoder(mean, o['coors'], 1, num_voxels=o['count'], precision=precision)
            voxels, num_points, coors = self.voxelize(points)
            feats = self.pts_voxel_encoder(voxels, num_points, coors)
            return self.pts_middle_encoder(feats, coors, bs, precision=precision)
        pts, coors = self.voxelize(points)
        vf, vc, count = self.pts_voxel_encoder.forward_padded(pts, coors, batch_size=bs)
        return self.pts_middle_encoder(vf, vc, bs, num_voxels=count, precision=precision)
