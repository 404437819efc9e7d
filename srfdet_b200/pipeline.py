"""Frame pipeline of the measured hot path (BASELINE.json metric: frames/s of
voxelize + SparseEncoder + region fusion), assembled from the drop-in modules.

Two scopes:
  scope='full' (default): the reference's whole LiDAR(+camera) forward after the image backbone --
      points -> Voxelization (+VFE) -> SparseEncoderCustom -> SECONDCustom -> FPN -> SRFDetHead (Dynamic
      Proposal Generation, 5 CHAINED stages: RoI sampling on the REAL FPN maps [+ 6-camera image RoIs + fusion],
      self-attention, DynamicConv, FFN, towers, apply_deltas feeding the next stage's boxes) -> decode.
  scope='path': the round-1 definition (SURVEY.md 8a rows only), kept for continuity:

One frame =
  points (N,C) --Voxelization(+HardSimpleVFE | DynamicVFECustom)--> voxel features
         --SparseEncoderCustom--> dense BEV map (1, 256, H, W)
  for each of the 5 cascade stages (srfdet_head.py:428-458):
         proposals --BEV RoIAlign [+ multi-view image RoIAlign + fusion Linear]--> (P,49,C)
         --DynamicConv--> (P, C) region features

What is NOT in the frame (outside the hot path): the dense BEV backbone/FPN
between the encoder and the RoI stage, the image backbone, and the attention / FFN /
box-regression rows of a stage.  The RoI stage therefore samples *synthetic* FPN pyramids
(N(0,1) maps of the configs' shapes) with seeded per-stage proposal boxes; the proposal
features chain from stage to stage through DynamicConv.
"""
import numpy as np
import contextlib

import torch

from . import synth
from .plugin import (DynamicConv, SingleRoIExtractor, SRFDetPointPath, img_feats_sampling_bboxes_roi,
                     points_feats_sampling_bboxes_roi)
from .plugin import head as _head
from .plugin import registry
from . import _lib as _L

MODEL_CFG = {
    # configs/nus/srfdet_voxel_nusc_L.py:34-50 (+ LC variant :40-84 for the image branch)
    'nusc': dict(
        pts_voxel_layer=dict(max_num_points=10, voxel_size=[0.075, 0.075, 0.2], max_voxels=(120000, 160000),
                             point_cloud_range=[-55.2, -55.2, -5.0, 55.2, 55.2, 3.0]),
        pts_voxel_encoder=dict(type='HardSimpleVFE', num_features=5),
        pts_middle_encoder=dict(type='SparseEncoderCustom', in_channels=5, sparse_shape=[41, 1472, 1472], output_channels=128,
                                order=('conv', 'norm', 'act'),
                                encoder_channels=((16, 16, 32), (32, 32, 64), (64, 64, 128), (128, 128)),
                                encoder_paddings=((0, 0, 1), (0, 0, 1), (0, 0, [0, 1, 1]), (0, 0)), block_type='basicblock')),
    # configs/waymo/srfdet_dvoxel_waymo_L.py:26-59
    'waymo': dict(
        pts_voxel_layer=dict(voxel_size=[0.1, 0.1, 0.15], max_num_points=-1, point_cloud_range=[-76.8, -76.8, -2, 76.8, 76.8, 4],
                             max_voxels=(-1, -1)),
        pts_voxel_encoder=dict(type='DynamicVFECustom', in_channels=5, feat_channels=[5, 5], with_distance=False,
                               voxel_size=[0.1, 0.1, 0.15], with_cluster_center=True, with_voxel_center=True,
                               point_cloud_range=[-76.8, -76.8, -2, 76.8, 76.8, 4],
                               norm_cfg=dict(type='naiveSyncBN1dCustom', eps=1e-3, momentum=0.01)),
        pts_middle_encoder=dict(type='SparseEncoderCustom', in_channels=5, sparse_shape=[41, 1536, 1536], output_channels=128,
                                order=('conv', 'norm', 'act'),
                                encoder_channels=((16, 16, 32), (32, 32, 64), (64, 64, 128), (128, 128)),
                                encoder_paddings=((0, 0, 1), (0, 0, 1), (0, 0, [0, 1, 1]), (0, 0)), block_type='basicblock')),
    # configs/kitti/srfdet_voxel_kitti_L.py:30-56
    'kitti': dict(
        pts_voxel_layer=dict(voxel_size=[0.05, 0.05, 0.1], max_num_points=-1, point_cloud_range=[0, -40, -3, 70.4, 40, 1],
                             max_voxels=(-1, -1)),
        pts_voxel_encoder=dict(type='DynamicVFECustom', in_channels=4, feat_channels=[4], with_distance=False,
                               voxel_size=[0.05, 0.05, 0.1], with_cluster_center=True, with_voxel_center=True,
                               point_cloud_range=[0, -40, -3, 70.4, 40, 1],
                               norm_cfg=dict(type='naiveSyncBN1dCustom', eps=1e-3, momentum=0.01)),
        pts_middle_encoder=dict(type='SparseEncoderCustom', in_channels=4, sparse_shape=[41, 1600, 1408], order=('conv', 'norm', 'act'))),
}

HEAD_CFG = {  # feat channels C, dynamic dim d, BEV base size, box dims (Appendix A of SURVEY.md)
    'nusc': dict(C=128, d=32, bev_hw=(184, 184), box_dim=10, ff=512, grid=[1472, 1472, 40], classes=10),
    'waymo': dict(C=128, d=32, bev_hw=(192, 192), box_dim=8, ff=512, grid=[1536, 1536, 40], classes=3),
    'kitti': dict(C=256, d=64, bev_hw=(200, 176), box_dim=8, ff=1024, grid=[1600, 1408, 40], classes=3),
}


def backbone_cfg(kind):
    """pts_backbone / pts_neck / bbox_head of the reference configs (configs/nus/srfdet_voxel_nusc_L.py:55-141,
    configs/nus/srfdet_voxel_nusc_LC.py:84-179, configs/waymo/srfdet_dvoxel_waymo_L.py:60-152,
    configs/kitti/srfdet_voxel_kitti_L.py:65-160)."""
    h = HEAD_CFG[kind]
    neck = dict(type='FPN', norm_cfg=dict(type='BN2d', eps=1e-3, momentum=0.01), act_cfg=dict(type='ReLU'), in_channels=[128, 256],
                out_channels=h['C'], start_level=0, num_outs=4)
    if kind != 'kitti':
        neck['add_extra_convs'] = 'on_output'
    return dict(pts_backbone=dict(type='SECONDCustom', in_channels=256, out_channels=[128, 256], layer_nums=[5, 5], layer_strides=[1, 2],
                                  norm_cfg=dict(type='BN', eps=1e-3, momentum=0.01), conv_cfg=dict(type='Conv2d', bias=False)),
                pts_neck=neck)


def head_cfg(kind, fusion):
    h = HEAD_CFG[kind]
    layer = MODEL_CFG[kind]['pts_voxel_layer']
    wts = [1.0] * 8 + ([0.2, 0.2] if h['box_dim'] == 10 else [])
    single = dict(type='SingleSRFDetHead' if fusion else 'SingleSRFDetHeadLiDAR', num_cls_convs=2, num_reg_convs=3, dim_feedforward=h['ff'],
                  num_heads=8, dropout=0.1, act_cfg=dict(type='ReLU', inplace=True), dynamic_conv=dict(dynamic_dim=h['d'], dynamic_num=2),
                  pc_range=layer['point_cloud_range'], voxel_size=layer['voxel_size'], bbox_weights=wts)
    if fusion:
        single['use_fusion'] = True
    roi = lambda strides: dict(type='SingleRoIExtractor', roi_layer=dict(type='RoIAlign', output_size=7, sampling_ratio=2),
                               out_channels=h['C'], featmap_strides=strides)
    return dict(type='SRFDetHead', use_img=fusion, num_classes=h['classes'], feat_channels_lidar=h['C'], feat_channels_img=256,
                hidden_dim=128, lidar_feat_lvls=4, img_feat_lvls=4, num_proposals=N_PROP, num_heads=N_STAGES, deep_supervision=True,
                with_lidar_encoder=False, grid_size=h['grid'], out_size_factor=8, code_weights=wts, with_dpg=True, num_dpg_exp=4,
                single_head_lidar=single, roi_extractor_lidar=roi([8, 16, 32, 64]), roi_extractor_img=roi([4, 8, 16, 32]),
                test_cfg=dict(use_nms=False, max_per_img=300, post_center_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0]))
N_STAGES = 5
N_PROP = 900


def init_synthetic_head(head, seed):
    """Synthetic weights of SRFDetHead (no checkpoints offline): default torch inits, then (a) proposal embeddings spread over
    the range with car-like log sizes, (b) box-delta projections scaled down so five chained refinements stay in range."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        w = head.init_proposal_boxes.weight
        w[:, :3] = torch.randn(w.shape[0], 3, generator=g) * 1.2
        w[:, 3:6] = torch.log(torch.tensor([1.9, 4.6, 1.7])) + torch.randn(w.shape[0], 3, generator=g) * 0.3
        w[:, 6:8] = torch.nn.functional.normalize(torch.randn(w.shape[0], 2, generator=g), dim=1)
        for st in head.head_series_lidar:
            st.bboxes_delta_lidar.weight.mul_(0.05)
            st.bboxes_delta_lidar.bias.zero_()
        # DPG: the FC input is a sum over 4C ReLU channels (O(C) per pixel); scale fc1 so the expert logits are O(1)
        # and the softmax over experts MIXES the embeddings (a trained gate), instead of a brittle hard selection
        for fc in [getattr(head, n) for n in ('dpg_fc1_lidar', 'dpg_fc1_img') if hasattr(head, n)]:
            fc.weight.mul_(1.0 / (2.0 * head.feat_channels_lidar))


def init_synthetic_backbone(module, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in module.modules():
            if isinstance(m, torch.nn.Conv2d):
                fan_in = m.in_channels // m.groups * m.kernel_size[0] * m.kernel_size[1]
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) * (2.0 / fan_in) ** 0.5)


def _randomize_bn(module, seed):
    g = torch.Generator().manual_seed(seed)
    for m in module.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.2)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.weight.shape, generator=g) * 0.5 + 0.75)
            m.bias.data.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)


class RegionFeaturePipeline:
    """kind in {'nusc','waymo','kitti'}; fusion=True adds the 6-camera image branch + fusion
    projection of srfdet_voxel_nusc_LC.  Weights: torch.manual_seed(0) random init with
    non-trivial BatchNorm statistics (no checkpoints offline)."""

    def __init__(self, kind='nusc', fusion=False, device='cuda', precision=None, seed=0, channels_last=True, use_graph=False,
                 scope='path'):
        assert scope in ('path', 'full')
        self.scope = scope
        self.kind, self.fusion, self.device = kind, fusion, torch.device(device)
        self.use_graph = use_graph
        self._graphs = {}
        self.overlap_stage = True
        self.overlap_image_branch = True      # scope 'full': the image half of DPG runs on a side stream under the encoder
        self._aux = None
        self._aux2 = None
        self.precision = precision or registry.get_precision()
        h = HEAD_CFG[kind]
        self.C, self.d, self.box_dim = h['C'], h['d'], h['box_dim']
        self.pc_range = MODEL_CFG[kind]['pts_voxel_layer']['point_cloud_range']
        self.voxel_size = MODEL_CFG[kind]['pts_voxel_layer']['voxel_size']
        torch.manual_seed(seed)
        self.detector = SRFDetPointPath(**MODEL_CFG[kind])
        _randomize_bn(self.detector, seed + 1)
        self.detector = self.detector.to(self.device).eval()
        self.pooler = SingleRoIExtractor(dict(type='RoIAlign', output_size=7, sampling_ratio=2), self.C, [8, 16, 32, 64])
        self.pooler_img = SingleRoIExtractor(dict(type='RoIAlign', output_size=7, sampling_ratio=2), self.C, [4, 8, 16, 32])
        self.dynconvs = [DynamicConv(self.C, self.d).to(self.device).eval() for _ in range(N_STAGES)]
        self.fuse = [torch.nn.Linear(2 * self.C, self.C).to(self.device).eval() for _ in range(N_STAGES)] if fusion else None
        self._fuse_cache = [dict() for _ in range(N_STAGES)]
        # synthetic FPN pyramids (stand-ins for the dense backbones)
        # channels_last=True: the maps are torch.channels_last tensors (logical NCHW, NHWC in memory),
        # i.e. what a channels_last FPN emits for free; False = the reference's contiguous NCHW.
        self.channels_last = channels_last
        fmt = (lambda t: t.contiguous(memory_format=torch.channels_last)) if channels_last else (lambda t: t)
        self.bev_feats = [fmt(torch.as_tensor(f).to(self.device)) for f in synth.feature_pyramid(seed + 10, self.C, h['bev_hw'], 4)]
        self.img_feats = self.lidar2img = None
        if fusion:
            self.img_feats = [fmt(torch.as_tensor(f[0]).to(self.device)).unsqueeze(0)
                              for f in synth.feature_pyramid(seed + 11, self.C, (232, 400), 4, lead=(1, 6))]
            self.lidar2img = torch.as_tensor(synth.lidar2img(6, 1)[0]).to(self.device)
        self.stage_boxes = [torch.as_tensor(synth.proposals(seed + 20 + s, N_PROP, self.box_dim, 1)).to(self.device)
                            for s in range(N_STAGES)]
        self.prop0 = (torch.randn(N_PROP, self.C, generator=torch.Generator().manual_seed(seed + 30))).to(self.device)
        self.backbone = self.neck = self.head = None
        self.last = None
        if scope == 'full':
            from .plugin import registry as R
            torch.manual_seed(seed + 40)
            cfg = backbone_cfg(kind)
            self.backbone = R.build_backbone(cfg['pts_backbone'])
            self.neck = R.build_neck(cfg['pts_neck'])
            self.head = R.build_head(head_cfg(kind, fusion))
            init_synthetic_backbone(self.backbone, seed + 41)
            init_synthetic_backbone(self.neck, seed + 42)
            init_synthetic_head(self.head, seed + 43)
            for i, m in enumerate((self.backbone, self.neck, self.head)):
                _randomize_bn(m, seed + 44 + i)
            self.backbone, self.neck, self.head = [m.to(self.device).eval() for m in (self.backbone, self.neck, self.head)]

    def state(self):
        """Weights / synthetic inputs as numpy, for the CPU oracle."""
        sd = lambda m: {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
        full = {}
        if self.scope == 'full':
            h = HEAD_CFG[self.kind]
            layer = MODEL_CFG[self.kind]['pts_voxel_layer']
            full = dict(backbone=sd(self.backbone), neck=sd(self.neck), head=sd(self.head),
                        head_cfg=dict(pc_range=layer['point_cloud_range'], voxel_size=layer['voxel_size'], C=self.C, strides=[8, 16, 32, 64],
                                      istrides=[4, 8, 16, 32], attn_heads=8, d=self.d, n_cls=2, n_reg=3,
                                      bbox_weights=[1.0] * 8 + ([0.2, 0.2] if self.box_dim == 10 else []),
                                      scale_clamp=float(np.log(100000.0 / 16)), n_exp=4, n_p=N_PROP, stages=N_STAGES,
                                      extra_convs=self.kind != 'kitti'))
        return dict(scope=self.scope, **full, encoder=sd(self.detector.pts_middle_encoder),
                    vfe=sd(self.detector.pts_voxel_encoder),
                    dynconv=[sd(m) for m in self.dynconvs],
                    fuse=[sd(m) for m in self.fuse] if self.fusion else None,
                    bev_feats=[f.contiguous().cpu().numpy() for f in self.bev_feats],
                    img_feats=[f.contiguous().cpu().numpy() for f in self.img_feats] if self.fusion else None,
                    lidar2img=self.lidar2img.cpu().numpy() if self.fusion else None,
                    stage_boxes=[b.cpu().numpy() for b in self.stage_boxes], prop0=self.prop0.cpu().numpy())

    @torch.no_grad()
    def encode(self, points):
        return self.detector.extract_point_features([points], precision=self.precision)

    @torch.no_grad()
    def region_stages(self):
        prop = self.prop0
        main = torch.cuda.current_stream() if self.device.type == 'cuda' else None
        for s in range(N_STAGES):
            # within a stage the parameter generation (needs the proposal features) and the RoI
            # sampling (needs the boxes) are independent: they run on two streams and join
            # before the interaction.  Stages themselves stay sequential, as in the head.
            params = None
            if main is not None and self.overlap_stage:
                if self._aux is None:
                    self._aux = torch.cuda.Stream()
                self._aux.wait_event(main.record_event())
                with torch.cuda.stream(self._aux):
                    params = self.dynconvs[s].make_params(prop, self.precision)
                    ev_params = self._aux.record_event()
                if not torch.cuda.is_current_stream_capturing():
                    params.record_stream(main)
            boxes = self.stage_boxes[s]                  # static inputs: sampled with mutate=False (no in-place de-normalisation)
            enc = registry.act_enc(self.precision)
            if self.fusion and self.channels_last:
                # both samplers fill their half of the concatenated fusion input
                # (cat(img, pts), srfdet_head.py:2257) directly, in the GEMM's dtype, concurrently.
                # The image sampler reads the still-normalised stage boxes, the BEV one its clone.
                cenc = _L.F32 if enc is None else enc
                cat = torch.empty((N_PROP, 49, _L.enc_width(cenc, 2 * self.C)), device=self.device, dtype=_L.enc_torch_dtype(cenc))
                fork = main is not None and self.overlap_stage
                if fork:
                    if self._aux2 is None:
                        self._aux2 = torch.cuda.Stream()
                    self._aux2.wait_event(main.record_event())
                with (torch.cuda.stream(self._aux2) if fork else contextlib.nullcontext()):
                    img_feats_sampling_bboxes_roi(self.img_feats, self.stage_boxes[s], self.pooler_img, self.lidar2img,
                                                  self.pc_range, channel_last=True, out=cat, ch_offset=0, out_enc=cenc)
                    if fork:
                        ev_img = self._aux2.record_event()
                points_feats_sampling_bboxes_roi(self.bev_feats, boxes, self.pooler, self.pc_range, self.voxel_size,
                                                 channel_last=True, out=cat, ch_offset=self.C, mutate=False, out_enc=cenc)
                if fork:
                    main.wait_event(ev_img)
                roi = _head._linear(cat.view(N_PROP * 49, -1), self.fuse[s], self.precision, self._fuse_cache[s],
                                    'fuse').view(N_PROP, 49, -1)
            elif self.fusion:
                img_roi = img_feats_sampling_bboxes_roi(self.img_feats, boxes, self.pooler_img, self.lidar2img, self.pc_range,
                                                        channel_last=True)
                pts_roi = points_feats_sampling_bboxes_roi(self.bev_feats, boxes, self.pooler, self.pc_range, self.voxel_size,
                                                           channel_last=True, mutate=False)
                cat = torch.cat((img_roi, pts_roi), dim=2).view(N_PROP * 49, 2 * self.C)
                roi = _head._linear(cat, self.fuse[s], self.precision, self._fuse_cache[s], 'fuse').view(N_PROP, 49, -1)
            elif self.channels_last and enc is not None:
                # the interaction MMA consumes 16-bit (or hi + lo) operands: the sampler encodes once, on store
                roi = torch.empty((N_PROP, 49, _L.enc_width(enc, self.C)), dtype=_L.enc_torch_dtype(enc), device=self.device)
                points_feats_sampling_bboxes_roi(self.bev_feats, boxes, self.pooler, self.pc_range, self.voxel_size,
                                                 channel_last=True, out=roi, mutate=False, out_enc=enc)
            else:
                roi = points_feats_sampling_bboxes_roi(self.bev_feats, boxes, self.pooler, self.pc_range, self.voxel_size,
                                                       channel_last=True, mutate=False)
            if params is not None:
                main.wait_event(ev_params)
            prop = self.dynconvs[s].forward_kc(prop, roi, precision=self.precision, params=params)
        return prop

    @torch.no_grad()
    def calibrate(self, points):
        """Set-up step for the synthetic weights of scope 'full' (no checkpoints offline): give every BatchNorm2d of the
        dense backbone, the neck and the DPG staircase the running statistics of its own input on one calibration
        frame -- what training would have produced -- so activations stay O(1) through the 12 + 6 conv layers
        instead of growing heavy tails (random statistics make the chained head ill-conditioned: max / std of the
        FPN maps reaches 40).  Plain torch ops ON THE HOST (the encoder output is copied to the CPU: no third-party GPU
        kernel runs anywhere in this package), run once, not part of any measured frame."""
        import torch.nn.functional as F
        assert self.scope == 'full'
        g = torch.Generator().manual_seed(12345)

        def fit(conv, bn, x, stride, pad, groups=1):
            y = F.conv2d(x, conv.weight.detach().float().cpu(), None, stride=stride, padding=pad, groups=groups)
            bn.running_mean.copy_(y.mean((0, 2, 3)))
            bn.running_var.copy_(y.var((0, 2, 3), unbiased=False).clamp_min(1e-6))
            bn.weight.copy_(torch.rand(bn.weight.shape, generator=g) * 0.5 + 0.75)
            bn.bias.copy_(torch.randn(bn.bias.shape, generator=g) * 0.1)
            return F.relu(F.batch_norm(y, bn.running_mean.cpu(), bn.running_var.cpu(), bn.weight.cpu(), bn.bias.cpu(), False, 0.0, bn.eps))
        old_tf32 = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        try:
            x = self.detector.extract_point_features([points], precision='fp32').float().cpu().contiguous()
            feats = []
            for block in self.backbone.blocks:
                for j in range(0, len(block), 3):
                    x = fit(block[j], block[j + 1], x, block[j].stride[0], 1)
                feats.append(x)
            lat = [fit(cm.conv, cm.bn, f, 1, 0) for cm, f in zip(self.neck.lateral_convs, feats)]
            for i in range(len(lat) - 1, 0, -1):
                lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], size=lat[i - 1].shape[2:], mode='nearest')
            outs = [fit(self.neck.fpn_convs[i].conv, self.neck.fpn_convs[i].bn, lat[i], 1, 1) for i in range(len(lat))]
            for i in range(len(lat), self.neck.num_outs):
                if self.neck.add_extra_convs:
                    outs.append(fit(self.neck.fpn_convs[i].conv, self.neck.fpn_convs[i].bn, outs[-1], 2, 1))
                else:
                    outs.append(F.max_pool2d(outs[-1], 1, stride=2))

            def staircase(convs, maps):
                xx = None
                for l, cm in enumerate(convs):
                    inp = maps[l] if xx is None else torch.cat([maps[l], xx], dim=1)
                    xx = fit(cm.conv, cm.bn, inp, 2, 1, groups=inp.shape[1])
            staircase(self.head.dpg_dw_convs_lidar, outs)
            if self.fusion:
                staircase(self.head.dpg_dw_convs_img, [f[0].float().cpu().contiguous() for f in self.img_feats])
        finally:
            torch.backends.cudnn.allow_tf32 = old_tf32
        self._graphs = {}

    @torch.no_grad()
    def full_chain(self, bev, dpg_img=None):
        """dense BEV map -> SECONDCustom -> FPN -> SRFDetHead (DPG, chained stages) -> decode.
        Returns the decoded boxes + scores; logits / boxes / pyramid in self.last."""
        from .plugin.bev_backbone import nchw_to_rows, _tc_enc
        n, c, h, w = bev.shape
        feats = self.backbone.forward_rows(nchw_to_rows(bev, _tc_enc(self.precision)), n, h, w, self.precision)
        pyramid = self.neck.forward_rows(feats, n, self.precision)
        logits, boxes = self.head(self.img_feats if self.fusion else None, pyramid, None, lidar2img=self.lidar2img, precision=self.precision,
                                  dpg_img_logits=dpg_img)
        scores, dec = self.head.decode(logits, boxes)
        self.last = dict(pyramid=pyramid, logits=logits, boxes=boxes, scores=scores, det_boxes=dec)
        self.last_pyramid = pyramid
        return torch.cat([dec[0], scores[0]], dim=1)          # (P, box_dim-1 + classes): what a caller post-processes

    @torch.no_grad()
    def _run_frame_eager(self, points):
        if self.scope != 'full':
            return self.encode(points), self.region_stages()
        dpg_img = ev = None
        if self.fusion:
            # the image half of Dynamic Proposal Generation depends on the image FPN maps only: start it before the
            # encoder, on a side stream, and join right before the proposals are mixed
            main = torch.cuda.current_stream()
            if self.overlap_image_branch:
                if self._aux2 is None:
                    self._aux2 = torch.cuda.Stream()
                self._aux2.wait_event(main.record_event())
                with torch.cuda.stream(self._aux2):
                    dpg_img = self.head.dpg_image_logits(self.img_feats)
                    ev = self._aux2.record_event()
                if not torch.cuda.is_current_stream_capturing():
                    dpg_img.record_stream(main)
            else:
                dpg_img = self.head.dpg_image_logits(self.img_feats)
        bev = self.encode(points)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)
        return bev, self.full_chain(bev, dpg_img)

    @torch.no_grad()
    def run_frames(self, clouds, host=False):
        """Frame-level concurrency on one GPU (BASELINE config 5): frame i of `clouds` runs as its own CUDA graph
        (slot i: private input / intermediate / output buffers) on its own stream; the streams fork from and
        join the current stream.  host=True: `clouds` are pinned host tensors, results come back as pinned host
        tensors (H2D / D2H on the frame's stream).  Returns the per-frame outputs."""
        main = torch.cuda.current_stream()
        if not hasattr(self, '_slot_streams'):
            self._slot_streams = []
        while len(self._slot_streams) < len(clouds):
            self._slot_streams.append(torch.cuda.Stream())
        fork = main.record_event()
        outs = []
        for i, pts in enumerate(clouds):
            st = self._slot_streams[i]
            st.wait_event(fork)
            with torch.cuda.stream(st):
                outs.append(self.run_frame_host(pts, slot=i, sync=False) if host else self.run_frame(pts, slot=i))
            main.wait_event(st.record_event())
        if host:
            main.synchronize()
        return outs

    @torch.no_grad()
    def run_frame(self, points, slot=0):
        """points (N,C) fp32 CUDA tensor -> (dense BEV map, region features (P,C)).

        use_graph=True: the whole frame (every kernel reads its counts from device memory, so
        the launch sequence is independent of the data) is captured once per point-count into a
        CUDA graph and replayed; outputs then live in graph-owned buffers that the next replay
        overwrites."""
        if not self.use_graph:
            return self._run_frame_eager(points)
        key = (tuple(points.shape), self.precision, self.scope, slot)
        g = self._graphs.get(key)
        if g is None:
            static_in = torch.empty_like(points)
            static_in.copy_(points)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):          # warm-up off the capture: lazy kernel attributes, weight packing
                for _ in range(2):
                    self._run_frame_eager(static_in)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self._run_frame_eager(static_in)
            g = self._graphs[key] = (graph, static_in, out)
        graph, static_in, out = g
        if static_in.data_ptr() != points.data_ptr():
            static_in.copy_(points, non_blocking=True)
        graph.replay()
        return out

    @torch.no_grad()
    def run_frame_host(self, points_pinned, slot=0, sync=True):
        """Public end-to-end call: HOST (pinned) points in, HOST region features out."""
        key = (tuple(points_pinned.shape), self.precision, self.scope, slot)
        if self.use_graph and key in self._graphs:
            pts = self._graphs[key][1]                      # H2D straight into the graph's input buffer
            pts.copy_(points_pinned, non_blocking=True)
        else:
            pts = points_pinned.to(self.device, non_blocking=True)
        bev, obj = self.run_frame(pts, slot=slot)
        hkey = ('host_out', tuple(obj.shape), slot)
        out = self._graphs.get(hkey)
        if out is None:
            out = self._graphs[hkey] = torch.empty(obj.shape, dtype=obj.dtype, pin_memory=True)
        out.copy_(obj, non_blocking=True)
        if sync:
            torch.cuda.current_stream().synchronize()
        return bev, out
