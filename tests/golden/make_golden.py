"""Generate tests/golden/*.npz|json from the reference's own Python (run in the build
container only: `python tests/golden/make_golden.py`; needs /root/reference).

What is pinned here is the reference-OWNED code on the hot path, executed unmodified
(see ref_stubs.py).  Third-party kernels it calls are replaced by stand-ins, recorded in
each fixture's `standins` field:
  RoIAlign            -> torchvision.ops.roi_align(aligned=True, sampling_ratio=2)
  SingleRoIExtractor  -> restated mmdet 2.28.2 level mapping around that RoIAlign
  bbox2roi            -> restated mmdet bbox2roi
  DynamicScatter      -> torch.unique(dim=0, sorted) + scatter_reduce
Inputs are tiny and seeded; fixtures total a few hundred KB.
"""
import json
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F
import torchvision

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, '/root/reference')
import ref_stubs  # noqa: E402


# ---------------- third-party stand-ins ----------------
def bbox2roi(bbox_list):
    out = []
    for i, b in enumerate(bbox_list):
        out.append(torch.cat([b.new_full((b.size(0), 1), i), b[:, :4]], dim=-1))
    return torch.cat(out, 0)


class Pooler:
    """mmdet SingleRoIExtractor stand-in (finest_scale=56, RoIAlign 7x7 sr=2 aligned avg)."""

    def __init__(self, strides):
        self.strides = strides
        self.num_inputs = len(strides)

    def __call__(self, feats, rois):
        scale = torch.sqrt((rois[:, 3] - rois[:, 1]) * (rois[:, 4] - rois[:, 2]))
        lv = torch.floor(torch.log2(scale / 56 + 1e-6)).clamp(min=0, max=len(feats) - 1).long()
        out = feats[0].new_zeros(rois.size(0), feats[0].size(1), 7, 7)
        for i, f in enumerate(feats):
            inds = (lv == i).nonzero().flatten()
            if inds.numel():
                out[inds] = torchvision.ops.roi_align(f, rois[inds], (7, 7), 1.0 / self.strides[i], 2, True)
        return out


class DynamicScatterStandin(torch.nn.Module):
    def __init__(self, voxel_size, point_cloud_range, average_points):
        super().__init__()
        self.average_points = average_points

    def forward_single(self, points, coors):
        bad = (coors < 0).any(-1, keepdim=True)
        clean = coors.masked_fill(bad, -1)
        oc, inv, cnt = torch.unique(clean, dim=0, sorted=True, return_inverse=True, return_counts=True)
        if oc[0, 0] < 0:
            oc, cnt, inv = oc[1:], cnt[1:], inv - 1
        keep = inv >= 0
        idx = inv[keep].view(-1, 1).expand(-1, points.size(1))
        if self.average_points:
            red = points.new_zeros(oc.size(0), points.size(1)).scatter_add_(0, idx, points[keep])
            red = red / cnt.view(-1, 1)
        else:
            red = points.new_full((oc.size(0), points.size(1)), -float('inf'))
            red.scatter_reduce_(0, idx, points[keep], reduce='amax')
        return red, oc

    def forward(self, points, coors):
        if coors.size(-1) == 3:
            return self.forward_single(points, coors)
        bs = int(coors[-1, 0]) + 1
        vs, cs = [], []
        for i in range(bs):
            inds = torch.where(coors[:, 0] == i)
            v, c = self.forward_single(points[inds], coors[inds][:, 1:])
            vs.append(v)
            cs.append(F.pad(c, (1, 0), value=i))
        return torch.cat(vs), torch.cat(cs)


def sd_np(module):
    return {k: v.detach().numpy() for k, v in module.state_dict().items()}


def randomize_bn(module, gen):
    for m in module.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=gen) * 0.3)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=gen) + 0.5)
            m.weight.data.copy_(torch.rand(m.weight.shape, generator=gen) + 0.5)
            m.bias.data.copy_(torch.randn(m.bias.shape, generator=gen) * 0.2)


def load_cfg_model(path):
    ns = {}
    with open(path) as f:
        exec(compile(f.read(), path, 'exec'), ns)
    return ns['model']


def main():
    sys.path.insert(0, os.path.join(HERE, '..', '..'))
    from srfdet_b200 import synth
    ref_stubs.install(dynamic_scatter_cls=DynamicScatterStandin, bbox2roi=bbox2roi)
    util = ref_stubs.ref_import('mmdet3d_plugin.core.bbox.util')
    head = ref_stubs.ref_import('mmdet3d_plugin.models.sparse_heads.srfdet_head')
    vfe_mod = ref_stubs.ref_import('mmdet3d_plugin.models.voxel_encoders.voxel_encoder')
    enc_mod = ref_stubs.ref_import('mmdet3d_plugin.models.middle_encoders.sparse_encoder_custom')
    standins = 'RoIAlign=torchvision.ops.roi_align; SingleRoIExtractor,bbox2roi=restated mmdet; ' \
               'DynamicScatter=torch.unique+scatter_reduce'
    g = torch.Generator().manual_seed(1234)
    pc_range = [-55.2, -55.2, -5.0, 55.2, 55.2, 3.0]
    voxel_size = [0.075, 0.075, 0.2]

    def rand_boxes(bs, n, dims):
        b = torch.zeros(bs, n, dims)
        b[..., :3] = torch.rand(bs, n, 3, generator=g)
        b[..., 3:6] = torch.log(torch.tensor([1.9, 4.6, 1.7])) + 0.3 * torch.randn(bs, n, 3, generator=g)
        yaw = (torch.rand(bs, n, generator=g) * 2 - 1) * np.pi
        b[..., 6], b[..., 7] = torch.sin(yaw), torch.cos(yaw)
        # a few big boxes so that all four pyramid levels are hit
        b[:, ::7, 3:5] += 2.0
        b[:, ::11, 3:5] += 3.2
        return b

    # ---- a6: boxes3d_to_corners3d (core/bbox/util.py:84-176)
    boxes = rand_boxes(2, 37, 8)
    boxes[..., :3] = boxes[..., :3] * 100 - 50
    corners = util.boxes3d_to_corners3d(boxes.clone(), bottom_center=False, ry=False)
    np.savez(os.path.join(HERE, 'corners.npz'), boxes=boxes.numpy(), corners=corners.numpy(),
             standins='none')

    # ---- a7: points_feats_sampling_bboxes_roi (srfdet_head.py:2568-2629)
    C = 16
    strides = [8, 16, 32, 64]
    pfeats = [torch.as_tensor(synth.hash_field((2, C, 184 // 2 ** i, 184 // 2 ** i), 100 + i)) for i in range(4)]
    boxes = rand_boxes(2, 48, 10)
    self_ns = types.SimpleNamespace(pc_range_lidar=pc_range, voxel_size_lidar=voxel_size, is_kitti=False)
    b_in = boxes.clone()
    out = head.SingleSRFDetHead.points_feats_sampling_bboxes_roi(self_ns, pfeats, b_in, Pooler(strides), None)
    np.savez(os.path.join(HERE, 'bev_roi.npz'), boxes=boxes.numpy(), boxes_after=b_in.numpy(),
             pc_range=np.array(pc_range), voxel_size=np.array(voxel_size), strides=np.array(strides),
             feat_seed=100, C=C, out=out.numpy(), standins=standins)

    # ---- a8: img_feats_sampling_bboxes_roi (srfdet_head.py:2424-2565), B=1
    l2i = synth.lidar2img(6, 1)
    istrides = [4, 8, 16, 32]
    ifeats = [torch.as_tensor(synth.hash_field((1, 6, C, 232 // 2 ** i, 400 // 2 ** i), 200 + i)) for i in range(4)]
    boxes = rand_boxes(1, 40, 10)
    metas = [dict(lidar2img=l2i[0])]
    out = head.SingleSRFDetHead.img_feats_sampling_bboxes_roi(self_ns, ifeats, boxes.clone(), Pooler(istrides), metas)
    np.savez(os.path.join(HERE, 'img_roi.npz'), boxes=boxes.numpy(), lidar2img=l2i,
             pc_range=np.array(pc_range), strides=np.array(istrides),
             feat_seed=200, C=C, out=out.numpy(), standins=standins)

    # ---- a10: DynamicConv (srfdet_head.py:2633-2693)
    torch.manual_seed(7)
    dc = head.DynamicConv(feat_channels=C, dynamic_dim=4, dynamic_num=2, pooler_resolution=7).eval()
    for p in dc.parameters():
        p.data.copy_(torch.randn(p.shape, generator=g) * (0.3 if p.dim() > 1 else 0.5) + (1.0 if p.dim() == 1 and 'norm' in '' else 0))
    K = 23
    prop = torch.randn(1, K, C, generator=g)
    roi = torch.randn(49, K, C, generator=g)
    with torch.no_grad():
        out = dc(prop, roi)
    np.savez(os.path.join(HERE, 'dynconv.npz'), prop=prop.numpy(), roi=roi.numpy(), out=out.numpy(),
             dynamic_dim=4, **{'p.' + k: v for k, v in sd_np(dc).items()}, standins='none')

    # ---- a7..a10 chained + 'next' rows: SingleSRFDetHeadLiDAR.forward (srfdet_head.py:1455-1529)
    torch.manual_seed(11)
    lid = head.SingleSRFDetHeadLiDAR(num_classes=10, feat_channels=C, dim_feedforward=32, num_cls_convs=2,
                                     num_reg_convs=3, num_heads=2, dropout=0.1,
                                     dynamic_conv=dict(dynamic_dim=4, dynamic_num=2), pc_range=pc_range,
                                     voxel_size=voxel_size).eval()
    boxes = rand_boxes(1, 24, 10)
    b_in = boxes.clone()
    with torch.no_grad():
        # NB: prop_feats=None would crash in the reference itself (srfdet_head.py:1476 reads
        # self.feat_channels, which SingleSRFDetHeadLiDAR never sets), so give proposals.
        prop_in = torch.randn(1, 24, C, generator=g)
        logits, pred, obj = lid([f[:1] for f in pfeats], b_in[:1], prop_in, Pooler(strides), None)
    np.savez(os.path.join(HERE, 'head_lidar.npz'), boxes=boxes.numpy(), boxes_after=b_in.numpy(), prop=prop_in.numpy(),
             logits=logits.numpy(), pred=pred.numpy(), obj=obj.numpy(), pc_range=np.array(pc_range),
             voxel_size=np.array(voxel_size), strides=np.array(strides),
             feat_seed=100, C=C,
             **{'p.' + k: v for k, v in sd_np(lid).items()}, standins=standins)

    # ---- a7..a10 with fusion: SingleSRFDetHead.forward (srfdet_head.py:2221-2326), use_fusion=True
    torch.manual_seed(13)
    fus = head.SingleSRFDetHead(num_classes=10, feat_channels=C, dim_feedforward=32, num_cls_convs=2,
                                num_reg_convs=3, num_heads=2, dropout=0.1,
                                dynamic_conv=dict(dynamic_dim=4, dynamic_num=2), pc_range=pc_range,
                                voxel_size=voxel_size, use_fusion=True).eval()
    boxes = rand_boxes(1, 24, 10)
    b_in = boxes.clone()
    pf1t = [f[:1] for f in pfeats]
    with torch.no_grad():
        logits, pred, obj = fus(ifeats, pf1t, b_in, None, Pooler(strides), metas, pooler_img=Pooler(istrides))
    np.savez(os.path.join(HERE, 'head_fusion.npz'), boxes=boxes.numpy(), boxes_after=b_in.numpy(),
             logits=logits.numpy(), pred=pred.numpy(), obj=obj.numpy(), lidar2img=l2i,
             pc_range=np.array(pc_range), voxel_size=np.array(voxel_size), strides=np.array(strides),
             istrides=np.array(istrides),
             feat_seed=100, ifeat_seed=200, C=C,
             **{'p.' + k: v for k, v in sd_np(fus).items()}, standins=standins)

    # ---- a4: DynamicVFECustom.forward (voxel_encoders/voxel_encoder.py:162-240), Waymo-style
    ops_norm = ref_stubs.ref_import('mmdet3d_plugin.ops.norm')  # registers naiveSyncBN1dCustom
    assert ops_norm is not None
    wcfg = load_cfg_model('/root/reference/configs/waymo/srfdet_dvoxel_waymo_L.py')['pts_voxel_encoder']
    for tag, cfg in [('waymo', wcfg),
                     ('kitti', load_cfg_model('/root/reference/configs/kitti/srfdet_voxel_kitti_L.py')['pts_voxel_encoder'])]:
        cfg = dict(cfg)
        cfg.pop('type')
        torch.manual_seed(17)
        vfe = vfe_mod.DynamicVFECustom(**cfg).eval()
        randomize_bn(vfe, g)
        rng = np.random.default_rng(5)
        n = 600
        cin = cfg['in_channels']
        lo = np.array(cfg['point_cloud_range'][:3])
        hi = np.array(cfg['point_cloud_range'][3:])
        vsz = np.array(cfg['voxel_size'])
        centre = lo + (hi - lo) * np.array([0.5, 0.5, 0.3])
        pts = np.zeros((n, cin), np.float32)
        pts[:, :3] = centre + rng.uniform(-1, 1, (n, 3)) * np.array([1.0, 1.0, 0.5])
        pts[:, 3:] = rng.uniform(0, 1, (n, cin - 3))
        pts[::50, 0] = hi[0] + 1.0  # out-of-range points
        grid = np.round((hi - lo) / vsz).astype(np.int64)
        c = np.floor((pts[:, :3] - lo.astype(np.float32)) / vsz.astype(np.float32)).astype(np.int64)
        ok = ((c >= 0) & (c < grid)).all(1)
        coors = np.full((n, 4), -1, np.int64)
        coors[:, 0] = 0
        coors[ok, 1:] = c[ok][:, ::-1]
        coors_t = torch.as_tensor(coors, dtype=torch.int32)
        # two samples in the batch: second half is sample 1
        coors_t[n // 2:, 0] = 1
        with torch.no_grad():
            vf, vc = vfe(torch.as_tensor(pts), coors_t)
        np.savez(os.path.join(HERE, f'vfe_{tag}.npz'), points=pts, coors=coors_t.numpy(), voxel_feats=vf.numpy(),
                 voxel_coors=vc.numpy(), voxel_size=vsz, pc_range=np.array(cfg['point_cloud_range']),
                 **{'p.' + k: v for k, v in sd_np(vfe).items()}, standins=standins)

    # ---- a5: SparseEncoderCustom layer construction (sparse_encoder_custom.py:73-107,142-216)
    plans = {}
    for tag, path in [('nusc', 'configs/nus/srfdet_voxel_nusc_L.py'),
                      ('waymo', 'configs/waymo/srfdet_dvoxel_waymo_L.py'),
                      ('kitti', 'configs/kitti/srfdet_voxel_kitti_L.py')]:
        cfg = dict(load_cfg_model('/root/reference/' + path)['pts_middle_encoder'])
        cfg.pop('type')
        cfg.pop('init_cfg', None)
        ref_stubs.Recorder.calls = []
        enc_mod.SparseEncoderCustom(**cfg)
        calls = ref_stubs.Recorder.calls
        # the reference builds conv_input, then the stages, then conv_out
        plans[tag] = dict(cfg={k: v for k, v in cfg.items() if k != 'norm_cfg'}, calls=calls)
    with open(os.path.join(HERE, 'encoder_plan.json'), 'w') as f:
        json.dump(plans, f, indent=1, default=list)
    print('golden fixtures written to', HERE)


if __name__ == '__main__':
    main()
