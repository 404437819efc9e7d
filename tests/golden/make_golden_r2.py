"""Round-2 fixtures (rows f1-f4 of SURVEY.md 8): generated from the reference's own Python like
make_golden.py (same stubs; run in the build container only, needs /root/reference):

    python tests/golden/make_golden_r2.py

  pillar_vfe.npz   PillarFeatureNetCustom.forward + PFNLayer   (reference-owned, executed unmodified)
  second_fpn.npz   SECONDCustom.forward (reference-owned; Conv2d/BN2d = torch) + mmdet FPN (RESTATED
                   from mmdet 2.28.2 fpn.py: lateral 1x1 ConvModules, nearest top-down, 3x3 fpn convs,
                   add_extra_convs='on_output')
  srfdet_head.npz  SRFDetHead._get_init_proposals, .forward (5-stage chained loop) and .get_bboxes
                   decode without NMS (reference-owned; ConvModule = Conv2d(bias=False)+BN2d+ReLU
                   stand-in of mmcv's, RoIAlign/SingleRoIExtractor stand-ins of make_golden.py)
"""
import os
import sys

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, '/root/reference')
import ref_stubs  # noqa: E402
import make_golden as G1  # noqa: E402


def pillar(vfe_utils, pil_mod):
    gen = torch.Generator().manual_seed(4321)
    rng = np.random.default_rng(11)
    n, t, c = 150, 20, 5
    npts = rng.integers(1, t + 1, n)
    npts[:5] = t
    coors = np.stack([rng.integers(0, 2, n), np.zeros(n, np.int64), rng.integers(0, 512, n), rng.integers(0, 512, n)], 1)
    vox = np.zeros((n, t, c), np.float32)
    for i in range(n):
        ctr = np.array([(coors[i, 3] + 0.5) * 0.2 - 51.2, (coors[i, 2] + 0.5) * 0.2 - 51.2, -1.0])
        p = np.concatenate([ctr + rng.uniform(-0.1, 0.1, (npts[i], 3)) * [1, 1, 20], rng.uniform(0, 1, (npts[i], 2))], 1)
        vox[i, :npts[i]] = p
    out = dict(voxels=vox, num_points=npts.astype(np.int32), coors=coors.astype(np.int32))
    for tag, kw in [('new', dict(legacy=False)), ('legacy', dict(legacy=True, with_distance=True)), ('avg', dict(legacy=False, mode='avg'))]:
        cfg = dict(in_channels=5, feat_channels=[64], with_distance=False, voxel_size=[0.2, 0.2, 8],
                   norm_cfg=dict(type='BN1d', eps=1e-3, momentum=0.01), point_cloud_range=[-51.2, -51.2, -5.0, 51.2, 51.2, 3.0])
        cfg.update(kw)
        torch.manual_seed(23)
        net = pil_mod.PillarFeatureNetCustom(**cfg).eval()
        G1.randomize_bn(net, gen)
        with torch.no_grad():
            y = net(torch.as_tensor(vox.copy()), torch.as_tensor(npts), torch.as_tensor(coors))
        out.update({f'{tag}.out': y.numpy(), **{f'{tag}.p.{k}': v for k, v in G1.sd_np(net).items()}})
    np.savez_compressed(os.path.join(HERE, 'pillar_vfe.npz'), **out)


def main():
    ref_stubs.install(dynamic_scatter_cls=G1.DynamicScatterStandin, bbox2roi=G1.bbox2roi)
    sys.modules['mmcv.ops'].DynamicScatter = G1.DynamicScatterStandin      # imported (unused) by pillar_encoder_custom.py:4
    vfe_utils = ref_stubs.ref_import('mmdet3d_plugin.models.voxel_encoders.utils')
    pil_mod = ref_stubs.ref_import('mmdet3d_plugin.models.voxel_encoders.pillar_encoder_custom')
    pillar(vfe_utils, pil_mod)
    print('round-2 fixtures written to', HERE)


if __name__ == '__main__':
    main()
