"""Region-feature kernels (corners, BEV / image RoI sampling, fusion, DynamicConv) on CUDA
vs the reference goldens (tests/golden, produced by the reference's own Python) and vs the
oracle at production sizes."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O
from srfdet_b200 import synth
from util import cuda, rel_err

pytestmark = pytest.mark.gpu
PC = [-55.2, -55.2, -5.0, 55.2, 55.2, 3.0]
VS = [0.075, 0.075, 0.2]


def _z(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _pooler(strides, c):
    from srfdet_b200.plugin import SingleRoIExtractor
    return SingleRoIExtractor(dict(type='RoIAlign', output_size=7, sampling_ratio=2), c, strides)


def test_corners_golden(golden_dir):
    from srfdet_b200.plugin import boxes3d_to_corners3d
    z = _z(golden_dir, 'corners.npz')
    got = boxes3d_to_corners3d(cuda(z['boxes'])).cpu().numpy()
    np.testing.assert_allclose(got, z['corners'], rtol=1e-6, atol=3e-5)


def test_bev_roi_golden(golden_dir):
    from srfdet_b200.plugin import points_feats_sampling_bboxes_roi
    z = _z(golden_dir, 'bev_roi.npz')
    C = int(z['C'])
    feats = [cuda(synth.hash_field((2, C, 184 // 2 ** i, 184 // 2 ** i), int(z['feat_seed']) + i)) for i in range(4)]
    boxes = cuda(z['boxes'].copy())
    pooler = _pooler(z['strides'].tolist(), C)
    out = points_feats_sampling_bboxes_roi(feats, boxes, pooler, z['pc_range'].tolist(), z['voxel_size'].tolist())
    np.testing.assert_allclose(boxes.cpu().numpy(), z['boxes_after'], rtol=0, atol=1e-5)    # in-place centre mutation
    np.testing.assert_allclose(out.cpu().numpy(), z['out'], rtol=0, atol=2e-4)
    # channel-last layout is the same numbers transposed
    boxes = cuda(z['boxes'].copy())
    cl = points_feats_sampling_bboxes_roi(feats, boxes, pooler, z['pc_range'].tolist(), z['voxel_size'].tolist(), channel_last=True)
    np.testing.assert_array_equal(cl.permute(0, 2, 1).reshape(out.shape).cpu().numpy(), out.cpu().numpy())
    # mutate=False: same features, the caller's boxes stay normalised
    boxes = cuda(z['boxes'].copy())
    keep = points_feats_sampling_bboxes_roi(feats, boxes, pooler, z['pc_range'].tolist(), z['voxel_size'].tolist(), mutate=False)
    np.testing.assert_array_equal(boxes.cpu().numpy(), z['boxes'])
    np.testing.assert_array_equal(keep.cpu().numpy(), out.cpu().numpy())


def test_img_roi_golden(golden_dir):
    from srfdet_b200.plugin import img_feats_sampling_bboxes_roi
    z = _z(golden_dir, 'img_roi.npz')
    C = int(z['C'])
    feats = [cuda(synth.hash_field((1, 6, C, 232 // 2 ** i, 400 // 2 ** i), int(z['feat_seed']) + i)) for i in range(4)]
    boxes = cuda(z['boxes'].copy())
    out = img_feats_sampling_bboxes_roi(feats, boxes, _pooler(z['strides'].tolist(), C), cuda(z['lidar2img'][0]), z['pc_range'].tolist())
    np.testing.assert_array_equal(boxes.cpu().numpy(), z['boxes'])        # image branch works on a clone
    np.testing.assert_allclose(out.cpu().numpy(), z['out'], rtol=0, atol=5e-4)


def test_roi_extractor_generic_vs_oracle():
    rng = np.random.default_rng(0)
    feats = [rng.standard_normal((2, 12, 40 // 2 ** i, 56 // 2 ** i)).astype(np.float32) for i in range(4)]
    xy = rng.uniform(-40, 500, (300, 2, 2))
    rois = np.stack([rng.integers(0, 2, 300), xy[:, 0].min(1), xy[:, 1].min(1), xy[:, 0].max(1), xy[:, 1].max(1)], 1).astype(np.float32)
    rois[0, 1:] = [10, 10, 10, 10]            # empty RoI
    rois[1, 1:] = [-900, -900, -800, -800]    # fully outside
    strides = [8, 16, 32, 64]
    ref = O.single_roi_extractor(feats, rois, strides)
    got = _pooler(strides, 12)([cuda(f) for f in feats], cuda(rois)).cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=0, atol=2e-5)
    assert np.abs(ref[1]).max() == 0


def test_bev_roi_production_size_vs_oracle():
    from srfdet_b200.plugin import points_feats_sampling_bboxes_roi
    feats = synth.feature_pyramid(3, 128, (184, 184), 4, lead=(1,))
    boxes = synth.proposals(4, 900, 10, 1)
    ref_boxes = boxes.copy()
    ref = O.points_roi_feats(feats, ref_boxes, PC, VS, [8, 16, 32, 64])
    b = cuda(boxes.copy())
    got, rois = points_feats_sampling_bboxes_roi([cuda(f) for f in feats], b, _pooler([8, 16, 32, 64], 128), PC, VS, return_rois=True)
    np.testing.assert_allclose(b.cpu().numpy(), ref_boxes, rtol=0, atol=1e-5)
    np.testing.assert_allclose(rois.cpu().numpy(), O.bev_rois(boxes.copy(), PC, VS).numpy(), rtol=0, atol=2e-3)
    assert rel_err(got.cpu().numpy(), ref) < 1e-4


@pytest.mark.parametrize('C', [128, 256])
def test_channels_last_maps_match_nchw(C):
    """torch.channels_last feature maps take the coalesced kernels; same numbers as the NCHW kernels."""
    from srfdet_b200.plugin import img_feats_sampling_bboxes_roi, points_feats_sampling_bboxes_roi
    cl = lambda t: t.contiguous(memory_format=torch.channels_last)
    feats = [cuda(f) for f in synth.feature_pyramid(7, C, (184, 184), 4, lead=(2,))]
    boxes = synth.proposals(8, 300, 10, 2)
    pool = _pooler([8, 16, 32, 64], C)
    ref = points_feats_sampling_bboxes_roi(feats, cuda(boxes.copy()), pool, PC, VS)
    b2 = cuda(boxes.copy())
    got = points_feats_sampling_bboxes_roi([cl(f) for f in feats], b2, pool, PC, VS)
    assert rel_err(got.cpu().numpy(), ref.cpu().numpy()) < 1e-5
    gotl = points_feats_sampling_bboxes_roi([cl(f) for f in feats], cuda(boxes.copy()), pool, PC, VS, channel_last=True)
    np.testing.assert_array_equal(gotl.permute(0, 2, 1).reshape(got.shape).cpu().numpy(), got.cpu().numpy())
    rois = O.bev_rois(boxes.copy(), PC, VS)
    gen = pool([cl(f) for f in feats], cuda(rois.numpy()))
    assert rel_err(gen.cpu().numpy(), ref.cpu().numpy()) < 1e-4     # rois recomputed on the host: last-ulp rectangle differences
    if C == 128:
        ifeats = [cuda(synth.hash_field((6, C, 232 // 2 ** i, 400 // 2 ** i), 60 + i)) for i in range(4)]
        l2i = cuda(synth.lidar2img(6, 1)[0])
        pooli = _pooler([4, 8, 16, 32], C)
        b1 = cuda(boxes[:1].copy())
        r0 = img_feats_sampling_bboxes_roi([f.unsqueeze(0) for f in ifeats], b1, pooli, l2i, PC)
        r1 = img_feats_sampling_bboxes_roi([cl(f).unsqueeze(0) for f in ifeats], b1, pooli, l2i, PC)
        assert rel_err(r1.cpu().numpy(), r0.cpu().numpy()) < 1e-5


@pytest.mark.parametrize('enc_name', ['f32', 'bf16', 'f16', 'bf16x2', 'f16x2'])
def test_samplers_fill_concatenated_buffer(enc_name):
    """out= / ch_offset=: image and BEV samplers write the two halves of cat(img, pts)
    (srfdet_head.py:2257) into one channel-last buffer: fp32 exact, 16-bit rounded, or hi + lo split
    rows [hi(2C) | lo(2C)] whose sum reproduces the fp32 value to the format's 2x precision."""
    from srfdet_b200 import _lib as L
    from srfdet_b200.plugin import img_feats_sampling_bboxes_roi, points_feats_sampling_bboxes_roi
    enc = {'f32': L.F32, 'bf16': L.BF16, 'f16': L.F16, 'bf16x2': L.BF16X2, 'f16x2': L.F16X2}[enc_name]
    dtype = L.enc_torch_dtype(enc)
    C = 128
    cl = lambda t: t.contiguous(memory_format=torch.channels_last)
    feats = [cl(cuda(f)) for f in synth.feature_pyramid(7, C, (184, 184), 4, lead=(1,))]
    ifeats = [cl(cuda(synth.hash_field((6, C, 232 // 2 ** i, 400 // 2 ** i), 60 + i))).unsqueeze(0) for i in range(4)]
    l2i = cuda(synth.lidar2img(6, 1)[0])
    boxes = synth.proposals(8, 300, 10, 1)
    pool, pooli = _pooler([8, 16, 32, 64], C), _pooler([4, 8, 16, 32], C)
    img = img_feats_sampling_bboxes_roi(ifeats, cuda(boxes.copy()), pooli, l2i, PC, channel_last=True)
    pts = points_feats_sampling_bboxes_roi(feats, cuda(boxes.copy()), pool, PC, VS, channel_last=True)
    ref = torch.cat((img, pts), dim=2)
    cat = torch.full((300, 49, L.enc_width(enc, 2 * C)), float('nan'), dtype=dtype, device='cuda')
    r = img_feats_sampling_bboxes_roi(ifeats, cuda(boxes.copy()), pooli, l2i, PC, channel_last=True, out=cat, ch_offset=0, out_enc=enc)
    assert r.data_ptr() == cat.data_ptr()
    points_feats_sampling_bboxes_roi(feats, cuda(boxes.copy()), pool, PC, VS, channel_last=True, out=cat, ch_offset=C, out_enc=enc)
    if enc == L.F32:
        assert torch.equal(cat, ref)
    elif not L.enc_is_split(enc):
        assert torch.equal(cat, ref.to(dtype))
    else:
        hi, lo = cat[..., :2 * C], cat[..., 2 * C:]
        assert torch.equal(hi, ref.to(dtype))
        assert torch.equal(lo, (ref - hi.float()).to(dtype))
        assert rel_err(L.decode(cat, 2 * C).cpu().numpy(), ref.cpu().numpy()) < (2e-5 if enc == L.BF16X2 else 1e-6)
    # NCHW maps cannot take the strided form: loud error, no silent fallback
    with pytest.raises(RuntimeError):
        points_feats_sampling_bboxes_roi([f.contiguous() for f in feats], cuda(boxes.copy()), pool, PC, VS,
                                         channel_last=True, out=cat, ch_offset=C)


def test_img_roi_production_size_vs_oracle():
    """6 cameras x 900 proposals (configs/nus/srfdet_voxel_nusc_LC.py), C reduced to 32 to bound
    oracle time; includes behind-camera boxes (degenerate rectangles that must read zeros)."""
    from srfdet_b200.plugin import img_feats_sampling_bboxes_roi
    C = 32
    feats = [synth.hash_field((1, 6, C, 232 // 2 ** i, 400 // 2 ** i), 50 + i) for i in range(4)]
    boxes = synth.proposals(5, 900, 10, 1)
    l2i = synth.lidar2img(6, 1)
    ref = O.img_roi_feats(feats, boxes, l2i, PC, [4, 8, 16, 32])
    got, rois = img_feats_sampling_bboxes_roi([cuda(f) for f in feats], cuda(boxes), _pooler([4, 8, 16, 32], C), cuda(l2i[0]), PC, return_rois=True)
    got = got.cpu().numpy()
    ref_rois = O.img_rois(boxes, l2i, PC).numpy()
    # rectangles: relative agreement (behind-camera coordinates reach 1e7 pixels)
    r_got, r_ref = rois.cpu().numpy()[:, 1:], ref_rois[:, 1:]
    near = np.abs(r_ref) < 2e3
    np.testing.assert_allclose(r_got[near], r_ref[near], rtol=2e-4, atol=2e-2)
    # beyond that the corner sits on the camera plane (depth clipped at 1e-5): the quotient
    # amplifies the last-ulp difference of the depth, only the magnitude is meaningful
    np.testing.assert_allclose(r_got[~near], r_ref[~near], rtol=2e-2)
    # features: per-proposal comparison; a proposal whose rectangle edge sits within float
    # rounding of a sampling / level threshold may legitimately differ -> allow a handful
    err = np.abs(got - ref).reshape(900, -1).max(1)
    assert (err > 2e-3).sum() <= 4, (err > 2e-3).sum()
    assert np.median(err) < 1e-4
    assert (np.abs(ref).reshape(900, -1).max(1) > 0).sum() > 300


def _dc_params(z):
    return {k[2:]: z[k] for k in z.files if k.startswith('p.')}


def test_dynconv_golden_fp32(golden_dir):
    from srfdet_b200.plugin import DynamicConv
    z = _z(golden_dir, 'dynconv.npz')
    c, d = z['prop'].shape[2], int(z['dynamic_dim'])
    dc = DynamicConv(c, d).eval()
    dc.load_state_dict({k: torch.as_tensor(v) for k, v in _dc_params(z).items()})
    dc = dc.cuda()
    from srfdet_b200.plugin import registry, set_precision
    old = registry.get_precision()
    for mode in ('fp32', 'fp32_simt'):      # golden dims (c, d) are tiny: both modes take the FFMA interaction kernel
        set_precision(mode)
        try:
            out = dc(cuda(z['prop']), cuda(z['roi']))
        finally:
            set_precision(old)
        assert rel_err(out.cpu().numpy(), z['out']) < 1e-4


@pytest.mark.parametrize('c,d', [(128, 32), (256, 64)])
@pytest.mark.parametrize('precision,tol', [('fp32', 1e-4), ('fp32_simt', 1e-4), ('fp16', 1e-2), ('bf16', 2e-2)])
def test_dynconv_production_dims_vs_oracle(c, d, precision, tol):
    from srfdet_b200.plugin import DynamicConv
    torch.manual_seed(1)
    dc = DynamicConv(c, d).eval()
    g = torch.Generator().manual_seed(2)
    k = 900 if c == 128 else 300
    prop = torch.randn(k, c, generator=g)
    roi = torch.randn(k, c, 7, 7, generator=g)
    ref = O.dynamic_conv({n: p.detach().numpy() for n, p in dc.state_dict().items()}, prop.numpy(), roi.numpy(), d)
    roi_kc = roi.reshape(k, c, 49).permute(0, 2, 1).contiguous()
    out = dc.cuda().forward_kc(prop.cuda(), roi_kc.cuda(), precision=precision)
    assert rel_err(out.cpu().numpy(), ref) < tol


def _load_head(cls, z, **kw):
    sd = {k[2:]: torch.as_tensor(z[k]) for k in z.files if k.startswith('p.')}
    C = int(z['C'])
    head = cls(num_classes=10, feat_channels=C, dim_feedforward=32, num_cls_convs=2, num_reg_convs=3, num_heads=2, dropout=0.1,
               dynamic_conv=dict(dynamic_dim=4, dynamic_num=2), pc_range=z['pc_range'].tolist(), voxel_size=z['voxel_size'].tolist(), **kw).eval()
    head.load_state_dict(sd, strict=True)
    return head.cuda(), C


def test_single_head_lidar_golden(golden_dir):
    """Whole stage of SingleSRFDetHeadLiDAR.forward (srfdet_head.py:1455-1529) vs the reference's output."""
    from srfdet_b200.plugin import SingleSRFDetHeadLiDAR
    z = _z(golden_dir, 'head_lidar.npz')
    head, C = _load_head(SingleSRFDetHeadLiDAR, z)
    feats = [cuda(synth.hash_field((2, C, 184 // 2 ** i, 184 // 2 ** i), int(z['feat_seed']) + i)[:1]) for i in range(4)]
    boxes = cuda(z['boxes'].copy())
    with torch.no_grad():
        logits, pred, obj = head(feats, boxes, cuda(z['prop']), _pooler(z['strides'].tolist(), C), None, precision='fp32')
    np.testing.assert_allclose(boxes.cpu().numpy(), z['boxes_after'], rtol=0, atol=1e-5)
    assert rel_err(obj.cpu().numpy(), z['obj']) < 2e-4
    assert rel_err(logits.cpu().numpy(), z['logits']) < 2e-4
    assert rel_err(pred.cpu().numpy(), z['pred']) < 2e-4


@pytest.mark.parametrize('maps_cl', [False, True])
def test_single_head_fusion_golden(golden_dir, maps_cl):
    """SingleSRFDetHead.forward with use_fusion=True (srfdet_head.py:2221-2326): image RoIs + BEV RoIs + fusion.
    maps_cl: torch.channels_last maps -> the samplers write the concatenated fusion input directly."""
    from srfdet_b200.plugin import SingleSRFDetHead
    z = _z(golden_dir, 'head_fusion.npz')
    head, C = _load_head(SingleSRFDetHead, z, use_fusion=True)
    pf = [cuda(synth.hash_field((2, C, 184 // 2 ** i, 184 // 2 ** i), int(z['feat_seed']) + i)[:1]) for i in range(4)]
    imf = [cuda(synth.hash_field((1, 6, C, 232 // 2 ** i, 400 // 2 ** i), int(z['ifeat_seed']) + i)) for i in range(4)]
    if maps_cl:
        pf = [f.contiguous(memory_format=torch.channels_last) for f in pf]
        imf = [f[0].contiguous(memory_format=torch.channels_last).unsqueeze(0) for f in imf]
    boxes = cuda(z['boxes'].copy())
    metas = [dict(lidar2img=z['lidar2img'][0])]
    with torch.no_grad():
        logits, pred, obj = head(imf, pf, boxes, None, _pooler(z['strides'].tolist(), C), metas,
                                 pooler_img=_pooler(z['istrides'].tolist(), C), precision='fp32')
    np.testing.assert_allclose(boxes.cpu().numpy(), z['boxes_after'], rtol=0, atol=1e-5)
    assert rel_err(obj.cpu().numpy(), z['obj']) < 5e-4
    assert rel_err(pred.cpu().numpy(), z['pred']) < 5e-4
