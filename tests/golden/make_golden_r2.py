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


class ConvModuleStandin(nn.Module):
    """[3P] mmcv.cnn.ConvModule as the reference uses it: Conv2d (bias=False when a norm follows) ->
    norm ('bn') -> ReLU (default act_cfg); state-dict keys conv.weight, bn.*."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, groups=1, norm_cfg=None,
                 act_cfg=dict(type='ReLU'), conv_cfg=None, inplace=True, bias='auto'):
        super().__init__()
        use_bias = norm_cfg is None if bias == 'auto' else bias
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding, groups=groups, bias=use_bias)
        self.bn = nn.BatchNorm2d(out_channels, eps=norm_cfg.get('eps', 1e-5), momentum=norm_cfg.get('momentum', 0.1)) if norm_cfg else None
        self.act = nn.ReLU() if act_cfg is not None else None

    def forward(self, x):
        x = self.conv(x)
        if self.bn is not None:
            x = self.bn(x)
        return self.act(x) if self.act is not None else x


def build_conv_layer(cfg, *args, **kwargs):
    cfg = dict(cfg or dict(type='Conv2d'))
    assert cfg.pop('type') == 'Conv2d'
    kwargs.update(cfg)
    if kwargs.get('bias') == 'auto':
        kwargs['bias'] = True
    return nn.Conv2d(*args, **kwargs)


def build_norm_layer_2d(cfg, num_features, postfix=''):
    cfg = dict(cfg)
    cfg.pop('type')
    cfg.pop('requires_grad', None)
    return 'bn', nn.BatchNorm2d(num_features, **cfg)


class AttrDict(dict):
    __getattr__ = dict.get


class BoxHolder:
    def __init__(self, tensor, box_dim):
        self.tensor = tensor


def srfdet_head(head_mod, regs):
    """SRFDetHead: _get_init_proposals (:506-655), forward (:371-498: DPG -> sigmoid -> chained stages ->
    centre de-normalisation) and get_bboxes (:1228-1340, use_nms=False branch)."""
    sys.path.insert(0, os.path.join(HERE, '..', '..'))
    from srfdet_b200 import synth
    head_mod.ConvModule = ConvModuleStandin
    head_mod.build_conv_layer = build_conv_layer
    head_mod.build_head = lambda cfg: regs['HEADS'].build(cfg)
    head_mod.build_roi_extractor = lambda cfg: G1.Pooler(cfg['featmap_strides'])
    head_mod.build_loss = lambda cfg: None
    pc_range = [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]
    voxel_size = [0.4, 0.4, 0.2]
    C, P, stages = 16, 24, 3
    out = {}
    for tag, use_img in [('lidar', False), ('fusion', True)]:
        single = dict(type='SingleSRFDetHead' if use_img else 'SingleSRFDetHeadLiDAR', num_cls_convs=2, num_reg_convs=3,
                      dim_feedforward=32, num_heads=2, dropout=0.1, act_cfg=dict(type='ReLU', inplace=True),
                      dynamic_conv=dict(dynamic_dim=4, dynamic_num=2), pc_range=pc_range, voxel_size=voxel_size)
        if use_img:
            single['use_fusion'] = True
        torch.manual_seed(31 + int(use_img))
        net = head_mod.SRFDetHead(
            use_img=use_img, num_classes=10, feat_channels_lidar=C, feat_channels_img=C, hidden_dim=C, lidar_feat_lvls=4,
            img_feat_lvls=4, num_proposals=P, num_heads=stages, deep_supervision=True, with_lidar_encoder=False,
            grid_size=[256, 256, 40], out_size_factor=8, code_weights=[1.0] * 8 + [0.2, 0.2], with_dpg=True, num_dpg_exp=4,
            single_head_lidar=single,
            roi_extractor_lidar=dict(type='SingleRoIExtractor', roi_layer=dict(type='RoIAlign', output_size=7, sampling_ratio=2),
                                     out_channels=C, featmap_strides=[8, 16, 32, 64]),
            roi_extractor_img=dict(type='SingleRoIExtractor', roi_layer=dict(type='RoIAlign', output_size=7, sampling_ratio=2),
                                   out_channels=C, featmap_strides=[4, 8, 16, 32]),
            test_cfg=AttrDict(use_nms=False, max_per_img=20, post_center_range=[-40.0, -40.0, -10.0, 40.0, 40.0, 10.0]),
            train_cfg=None).eval()
        G1.randomize_bn(net, torch.Generator().manual_seed(77))
        with torch.no_grad():   # spread the embeddings so that boxes land inside the range with car-like sizes
            g = torch.Generator().manual_seed(78)
            w = net.init_proposal_boxes.weight
            w[:, :3] = torch.randn(w.shape[0], 3, generator=g) * 1.2
            w[:, 3:6] = torch.log(torch.tensor([1.9, 4.6, 1.7])) + torch.randn(w.shape[0], 3, generator=g) * 0.3
            for hd in net.head_series_lidar:      # modest box refinements per stage (xavier-sized deltas collapse the boxes)
                hd.bboxes_delta_lidar.weight.mul_(0.05)
                hd.bboxes_delta_lidar.bias.zero_()
            big = {}                              # large tensors are regenerated from a hash in the tests instead of being stored
            for bi, (name, prm) in enumerate(net.named_parameters()):
                if prm.numel() > 50000:
                    scale = float(1.0 / prm.shape[-1] ** 0.5)
                    prm.copy_(torch.as_tensor(synth.hash_field(tuple(prm.shape), 500 + bi)) * scale)
                    big[name] = np.array([500 + bi, scale], np.float64)
        pf = [torch.as_tensor(synth.hash_field((1, C, 32 // 2 ** i, 32 // 2 ** i), 300 + i)) for i in range(4)]
        imf = [torch.as_tensor(synth.hash_field((1, 6, C, 232 // 2 ** i, 400 // 2 ** i), 400 + i)) for i in range(4)] if use_img else None
        l2i = synth.lidar2img(6, 1)
        metas = [dict(lidar2img=l2i[0], box_type_3d=BoxHolder)]
        with torch.no_grad():
            b0, f0 = net._get_init_proposals([f.clone() for f in imf] if use_img else None, pf)
            logits, boxes = net([f.clone() for f in imf] if use_img else None, pf, metas)
            res = net.get_bboxes(logits, boxes, metas)
        out.update({f'{tag}.init_boxes': b0.numpy(), f'{tag}.init_feats': f0.numpy(), f'{tag}.logits': logits.numpy(),
                    f'{tag}.boxes': boxes.numpy(), f'{tag}.det_boxes': res[0][0].tensor.numpy(), f'{tag}.det_scores': res[0][1].numpy(),
                    f'{tag}.det_labels': res[0][2].numpy(), **{f'{tag}.p.{k}': v for k, v in G1.sd_np(net).items() if k not in big},
                    **{f'{tag}.big.{k}': v for k, v in big.items()}})
    np.savez_compressed(os.path.join(HERE, 'srfdet_head.npz'), pc_range=np.array(pc_range), voxel_size=np.array(voxel_size), C=C, P=P,
                        stages=stages, feat_seed=300, ifeat_seed=400, lidar2img=l2i, **out)


def fpn_restated(params, feats, out_channels, num_outs, eps):
    """[3P] mmdet 2.28.2 FPN.forward restated (start_level=0, add_extra_convs='on_output', norm+ReLU ConvModules,
    upsample nearest): params keyed like mmdet (lateral_convs.i.conv/bn, fpn_convs.i.conv/bn)."""
    import torch.nn.functional as F

    def cm(x, pre, stride, pad):
        x = F.conv2d(x, params[pre + '.conv.weight'], None, stride=stride, padding=pad)
        x = F.batch_norm(x, params[pre + '.bn.running_mean'], params[pre + '.bn.running_var'], params[pre + '.bn.weight'],
                         params[pre + '.bn.bias'], False, 0.0, eps)
        return F.relu(x)
    lat = [cm(f, f'lateral_convs.{i}', 1, 0) for i, f in enumerate(feats)]
    for i in range(len(lat) - 1, 0, -1):
        lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], size=lat[i - 1].shape[2:], mode='nearest')
    outs = [cm(lat[i], f'fpn_convs.{i}', 1, 1) for i in range(len(lat))]
    for i in range(len(lat), num_outs):
        outs.append(cm(outs[-1], f'fpn_convs.{i}', 2, 1))
    return outs


def second_fpn(second_mod):
    """SECONDCustom.forward (backbones/second_custom.py:77-91) on the reference's own module + restated FPN."""
    second_mod.build_conv_layer = build_conv_layer
    second_mod.build_norm_layer = build_norm_layer_2d
    gen = torch.Generator().manual_seed(99)
    torch.manual_seed(41)
    cfg = dict(in_channels=32, out_channels=[16, 32], layer_nums=[2, 2], layer_strides=[1, 2],
               norm_cfg=dict(type='BN', eps=1e-3, momentum=0.01), conv_cfg=dict(type='Conv2d', bias=False))
    net = second_mod.SECONDCustom(**cfg).eval()
    G1.randomize_bn(net, gen)
    x = torch.randn(1, 32, 24, 20, generator=gen)
    x = x * (torch.rand(1, 1, 24, 20, generator=gen) < 0.3)          # sparse occupancy like a scattered BEV map
    with torch.no_grad():
        feats = net(x)
    # FPN parameters (mmdet naming), out_channels 16, 4 outputs
    fp = {}
    oc = 16
    for i, ci in enumerate([16, 32]):
        fp[f'lateral_convs.{i}.conv.weight'] = torch.randn(oc, ci, 1, 1, generator=gen) * (1.0 / ci) ** 0.5
    for i in range(4):
        fp[f'fpn_convs.{i}.conv.weight'] = torch.randn(oc, oc, 3, 3, generator=gen) * (1.0 / (9 * oc)) ** 0.5
    for pre in [f'lateral_convs.{i}' for i in range(2)] + [f'fpn_convs.{i}' for i in range(4)]:
        fp[pre + '.bn.weight'] = torch.rand(oc, generator=gen) + 0.5
        fp[pre + '.bn.bias'] = torch.randn(oc, generator=gen) * 0.2
        fp[pre + '.bn.running_mean'] = torch.randn(oc, generator=gen) * 0.3
        fp[pre + '.bn.running_var'] = torch.rand(oc, generator=gen) + 0.5
    with torch.no_grad():
        outs = fpn_restated(fp, list(feats), oc, 4, 1e-3)
    np.savez_compressed(os.path.join(HERE, 'second_fpn.npz'), x=x.numpy(), **{f'feat{i}': f.numpy() for i, f in enumerate(feats)},
                        **{f'out{i}': o.numpy() for i, o in enumerate(outs)}, **{'b.' + k: v for k, v in G1.sd_np(net).items()},
                        **{'n.' + k: v.numpy() for k, v in fp.items()})


def main():
    regs = ref_stubs.install(dynamic_scatter_cls=G1.DynamicScatterStandin, bbox2roi=G1.bbox2roi)
    sys.modules['mmcv.ops'].DynamicScatter = G1.DynamicScatterStandin      # imported (unused) by pillar_encoder_custom.py:4
    vfe_utils = ref_stubs.ref_import('mmdet3d_plugin.models.voxel_encoders.utils')
    pil_mod = ref_stubs.ref_import('mmdet3d_plugin.models.voxel_encoders.pillar_encoder_custom')
    pillar(vfe_utils, pil_mod)
    sys.modules['mmdet.models'].BACKBONES = regs['BACKBONES']
    pkg = ref_stubs._mod('mmdet3d_plugin.models.backbones')       # skeleton package: its __init__ (VoVNet, ...) is not executed
    pkg.__path__ = [ref_stubs.REF_ROOT + '/mmdet3d_plugin/models/backbones']
    second_fpn(ref_stubs.ref_import('mmdet3d_plugin.models.backbones.second_custom'))
    srfdet_head(ref_stubs.ref_import('mmdet3d_plugin.models.sparse_heads.srfdet_head'), regs)
    print('round-2 fixtures written to', HERE)


if __name__ == '__main__':
    main()
