"""Pin the CPU oracle against the golden vectors produced by the reference's own Python
(tests/golden/make_golden.py) and against independent implementations available here
(torchvision roi_align, torch.unique, dense conv3d)."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import oracle as O
from srfdet_b200 import synth


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _params(z, prefix='p.'):
    return {k[len(prefix):]: z[k] for k in z.files if k.startswith(prefix)}


def test_corners_golden(golden_dir):
    z = _load(golden_dir, 'corners.npz')
    got = O.boxes3d_to_corners3d(z['boxes']).numpy()
    np.testing.assert_allclose(got, z['corners'], rtol=0, atol=2e-5)


def test_bev_roi_golden(golden_dir):
    z = _load(golden_dir, 'bev_roi.npz')
    C = int(z['C'])
    feats = [synth.hash_field((2, C, 184 // 2 ** i, 184 // 2 ** i), int(z['feat_seed']) + i) for i in range(4)]
    boxes = z['boxes'].copy()
    out = O.points_roi_feats(feats, boxes, z['pc_range'].tolist(), z['voxel_size'].tolist(), z['strides'].tolist())
    np.testing.assert_allclose(boxes, z['boxes_after'], rtol=0, atol=1e-5)   # in-place mutation
    np.testing.assert_allclose(out, z['out'], rtol=0, atol=2e-5)
    assert len(np.unique(O.map_roi_levels(O.bev_rois(z['boxes'].copy(), z['pc_range'].tolist(), z['voxel_size'].tolist()), 4).numpy())) >= 3


def test_img_roi_golden(golden_dir):
    z = _load(golden_dir, 'img_roi.npz')
    C = int(z['C'])
    feats = [synth.hash_field((1, 6, C, 232 // 2 ** i, 400 // 2 ** i), int(z['feat_seed']) + i) for i in range(4)]
    out = O.img_roi_feats(feats, z['boxes'], z['lidar2img'], z['pc_range'].tolist(), z['strides'].tolist())
    np.testing.assert_allclose(out, z['out'], rtol=0, atol=5e-5)
    assert np.abs(z['out']).max() > 0.1


def test_dynconv_golden(golden_dir):
    z = _load(golden_dir, 'dynconv.npz')
    roi = torch.as_tensor(z['roi'])                      # (49,K,C) as the reference passes it
    k, c = roi.shape[1], roi.shape[2]
    roi_kc77 = roi.permute(1, 2, 0).reshape(k, c, 7, 7).numpy()
    out = O.dynamic_conv(_params(z), z['prop'][0], roi_kc77, int(z['dynamic_dim']))
    np.testing.assert_allclose(out, z['out'], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('tag', ['waymo', 'kitti'])
def test_dynamic_vfe_golden(golden_dir, tag):
    z = _load(golden_dir, f'vfe_{tag}.npz')
    p = _params(z)
    params = {}
    for k, v in p.items():
        if k.startswith('cen2point_pos_enc.'):
            params['pos.' + k[len('cen2point_pos_enc.'):]] = v
        elif k.startswith('vfe_layers.'):
            params['vfe.' + k[len('vfe_layers.'):]] = v
    vf, vc = O.dynamic_vfe_custom(params, z['points'], z['coors'], z['voxel_size'].tolist(), z['pc_range'].tolist())
    np.testing.assert_array_equal(vc, z['voxel_coors'])
    np.testing.assert_allclose(vf, z['voxel_feats'], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('tag', ['nusc', 'waymo', 'kitti'])
def test_encoder_plan_golden(golden_dir, tag):
    """The flat conv list derived by the oracle equals what the reference constructor
    asked mmdet3d to build (recorded make_sparse_convmodule / SparseBasicBlock calls)."""
    with open(os.path.join(golden_dir, 'encoder_plan.json')) as f:
        rec = json.load(f)[tag]
    cfg = rec['cfg']
    plan = O.encoder_layer_plan(cfg['in_channels'], cfg.get('base_channels', 16), cfg.get('output_channels', 128),
                                cfg.get('encoder_channels', ((16,), (32, 32, 32), (64, 64, 64), (64, 64, 64))),
                                cfg.get('encoder_paddings', ((1,), (1, 1, 1), (1, 1, 1), ((0, 1, 1), 1, 1))),
                                cfg.get('block_type', 'conv_module'))
    flat = []
    for c in rec['calls']:
        if c['fn'] == 'SparseBasicBlock':
            flat.append(('subm', c['cin'], c['cout'], None))
            flat.append(('subm', c['cout'], c['cout'], None))
        else:
            kind = 'subm' if c['conv_type'] == 'SubMConv3d' else 'spconv'
            flat.append((kind, c['cin'], c['cout'], c['key'], c['ksize'], c['stride'], c['pad']))
    assert len(flat) == len(plan)
    t3 = lambda v: tuple(v) if isinstance(v, (list, tuple)) else (v, v, v)
    for a, b in zip(flat, plan):
        assert a[0] == b['kind'] and a[1] == b['cin'] and a[2] == b['cout'] and a[3] == b['key']
        if len(a) > 4:
            assert t3(a[4]) == tuple(b['ksize']) and t3(a[5]) == tuple(b['stride'])
            if b['kind'] == 'spconv':
                assert t3(a[6]) == tuple(b['pad'])
    n = {'nusc': 21, 'waymo': 21, 'kitti': 12}[tag]
    assert len(plan) == n


def test_roi_align_vs_torchvision():
    import torchvision
    rng = np.random.default_rng(0)
    feat = rng.standard_normal((2, 5, 23, 31)).astype(np.float32)
    rois = np.array([[0, 2.2, 3.1, 90.7, 60.2], [1, -30, -20, 10, 15], [0, 200, 100, 300, 200],
                     [1, 0, 0, 0, 0], [0, 100.5, 50.5, 101, 51], [1, -500, -500, 900, 900]], np.float32)
    for scale in (0.25, 0.125):
        got = O.roi_align(feat, rois, scale)
        ref = torchvision.ops.roi_align(torch.as_tensor(feat), torch.as_tensor(rois), (7, 7), scale, 2, True).numpy()
        np.testing.assert_allclose(got, ref, rtol=0, atol=1e-5)


def test_dynamic_scatter_drops_invalid():
    feats = np.arange(12, dtype=np.float32).reshape(6, 2)
    coors = np.array([[0, 1, 1, 1], [0, -1, -1, -1], [0, 1, 1, 1], [0, 0, 5, 2], [1, 0, 0, 0], [1, 3, -1, 2]], np.int32)
    vf, vc, p2v = O.dynamic_scatter(feats, coors, 'max')
    np.testing.assert_array_equal(vc, [[0, 0, 5, 2], [0, 1, 1, 1], [1, 0, 0, 0]])
    np.testing.assert_array_equal(vf, [[6, 7], [4, 5], [8, 9]])
    np.testing.assert_array_equal(p2v, [1, -1, 1, 0, 2, -1])
    vf, _, _ = O.dynamic_scatter(feats, coors, 'mean')
    np.testing.assert_allclose(vf, [[6, 7], [2, 3], [8, 9]])


def test_hard_voxelize_micro():
    """3-point hand case + points exactly on the range faces + max_voxels / max_points overflow."""
    vs, rng_ = [1.0, 1.0, 1.0], [0, 0, 0, 4, 4, 2]
    pts = np.array([[0.5, 0.5, 0.5, 9], [3.5, 0.5, 1.5, 8], [0.6, 0.4, 0.2, 7], [4.0, 1, 1, 6], [0, 0, 0, 5],
                    [-0.001, 1, 1, 4], [0.7, 0.7, 0.7, 3], [2.5, 2.5, 0.5, 2]], np.float32)
    v, c, n, p2v = O.hard_voxelize(pts, vs, rng_, max_points=3, max_voxels=2)
    np.testing.assert_array_equal(c, [[0, 0, 0], [1, 0, 3]])
    np.testing.assert_array_equal(n, [3, 1])
    np.testing.assert_array_equal(p2v, [0, 1, 0, -1, 0, -1, -1, -1])   # p6 beyond max_points, p7 beyond max_voxels
    np.testing.assert_array_equal(v[0, :, 3], [9, 7, 5])


def test_hard_voxelize_matches_bruteforce():
    g = synth.GEOM['nusc']
    pts = synth.dense_cloud('nusc', 3, 9000, extent=0.3)
    v, c, n, p2v = O.hard_voxelize(pts, g['voxel_size'], g['pc_range'], 10, 300)
    dc = O.dynamic_voxelize(pts, g['voxel_size'], g['pc_range'])
    seen, cnt = {}, {}
    for i, q in enumerate(map(tuple, dc)):
        if q[0] < 0:
            assert p2v[i] == -1
            continue
        if q not in seen:
            if len(seen) >= 300:
                assert p2v[i] == -1
                continue
            seen[q] = len(seen)
            cnt[q] = 0
        if q in seen and cnt[q] < 10:
            assert p2v[i] == seen[q]
            np.testing.assert_array_equal(v[seen[q], cnt[q]], pts[i])
            cnt[q] += 1
        else:
            assert p2v[i] == -1
    assert len(seen) == 300 and n.max() == 10


def _rand_sparse(rng, n, batch, dims):
    cells = set()
    while len(cells) < n:
        cells.add((int(rng.integers(0, batch)), int(rng.integers(0, dims[0])), int(rng.integers(0, dims[1])), int(rng.integers(0, dims[2]))))
    c = np.array(sorted(cells), np.int32)
    rng.shuffle(c)
    return c


def test_sparse_conv_vs_dense_conv3d():
    """SubM and strided sparse conv == dense conv3d evaluated at the active output sites."""
    rng = np.random.default_rng(1)
    dims, batch, cin, cout = (6, 9, 8), 2, 3, 4
    coors = _rand_sparse(rng, 120, batch, dims)
    feats = rng.standard_normal((120, cin)).astype(np.float32)
    dense = torch.zeros(batch, cin, *dims)
    cc = torch.as_tensor(coors, dtype=torch.int64)
    dense[cc[:, 0], :, cc[:, 1], cc[:, 2], cc[:, 3]] = torch.as_tensor(feats)
    w = rng.standard_normal((cout, 3, 3, 3, cin)).astype(np.float32)          # spconv2 layout
    wd = torch.as_tensor(w).permute(0, 4, 1, 2, 3).contiguous()                # conv3d layout
    # SubM
    pairs = O.rulebook_subm(coors, batch, dims)
    got = O.sparse_conv_c(feats, O.spconv2_weight_to_kio(w), pairs, len(coors))
    ref = F.conv3d(dense, wd, padding=1)[cc[:, 0], :, cc[:, 1], cc[:, 2], cc[:, 3]].numpy()
    np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(O.sparse_conv_mm(feats, O.spconv2_weight_to_kio(w), pairs, len(coors)), ref, rtol=1e-4, atol=1e-4)
    # strided, two padding variants incl. the (0,1,1) of encoder stage 3
    for pad in [(1, 1, 1), (0, 1, 1)]:
        oc, od, pairs = O.rulebook_strided(coors, batch, dims, (3, 3, 3), (2, 2, 2), pad)
        got = O.sparse_conv_c(feats, O.spconv2_weight_to_kio(w), pairs, len(oc))
        full = F.conv3d(F.pad(dense, (pad[2], pad[2], pad[1], pad[1], pad[0], pad[0])), wd, stride=2)
        assert tuple(full.shape[2:]) == tuple(od)
        o = torch.as_tensor(oc, dtype=torch.int64)
        np.testing.assert_allclose(got, full[o[:, 0], :, o[:, 1], o[:, 2], o[:, 3]].numpy(), rtol=1e-4, atol=1e-4)
        # every non-zero dense output is an active output
        mask = torch.zeros(batch, *od, dtype=torch.bool)
        mask[o[:, 0], o[:, 1], o[:, 2], o[:, 3]] = True
        assert float(full.abs().sum(1)[~mask].max()) == 0.0
        lin = ((oc[:, 0].astype(np.int64) * od[0] + oc[:, 1]) * od[1] + oc[:, 2]) * od[2] + oc[:, 3]
        assert np.all(np.diff(lin) > 0)
    # conv_out: k (3,1,1) s (2,1,1) p 0
    w2 = rng.standard_normal((cout, 3, 1, 1, cin)).astype(np.float32)
    oc, od, pairs = O.rulebook_strided(coors, batch, dims, (3, 1, 1), (2, 1, 1), (0, 0, 0))
    got = O.sparse_conv_c(feats, O.spconv2_weight_to_kio(w2), pairs, len(oc))
    full = F.conv3d(dense, torch.as_tensor(w2).permute(0, 4, 1, 2, 3).contiguous(), stride=(2, 1, 1))
    o = torch.as_tensor(oc, dtype=torch.int64)
    np.testing.assert_allclose(got, full[o[:, 0], :, o[:, 1], o[:, 2], o[:, 3]].numpy(), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize('tag,kw', [('new', dict(legacy=False)), ('legacy', dict(legacy=True, with_distance=True)),
                                    ('avg', dict(legacy=False, mode='avg'))])
def test_pillar_vfe_golden(golden_dir, tag, kw):
    """Oracle restatement of PillarFeatureNetCustom vs the reference's own output."""
    z = _load(golden_dir, 'pillar_vfe.npz')
    params = {k[len(tag) + 3:]: z[k] for k in z.files if k.startswith(tag + '.p.')}
    out = O.pillar_feature_net(params, z['voxels'], z['num_points'], z['coors'], [0.2, 0.2, 8],
                               [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], **kw)
    np.testing.assert_allclose(out, z[f'{tag}.out'], rtol=1e-5, atol=1e-5)


def _head_inputs(z, tag):
    """Parameters (large tensors regenerated from their hash seeds) and inputs of the srfdet_head fixture."""
    from srfdet_b200 import synth
    C = int(z['C'])
    params = {k[len(tag) + 3:]: z[k] for k in z.files if k.startswith(tag + '.p.')}
    for k in z.files:
        if k.startswith(tag + '.big.'):
            name = k[len(tag) + 5:]
            seed, scale = z[k]
            shape = {'dpg_fc2_lidar.weight': (4 * int(z['P']), 1024), 'dpg_fc1_img.weight': (1500, 900),
                     'dpg_fc2_img.weight': (4 * int(z['P']), 1500)}[name]
            params[name] = synth.hash_field(shape, int(seed)) * np.float32(scale)
    pf = [synth.hash_field((1, C, 32 // 2 ** i, 32 // 2 ** i), int(z['feat_seed']) + i) for i in range(4)]
    imf = [synth.hash_field((1, 6, C, 232 // 2 ** i, 400 // 2 ** i), int(z['ifeat_seed']) + i) for i in range(4)] if tag == 'fusion' else None
    cfg = dict(pc_range=z['pc_range'].tolist(), voxel_size=z['voxel_size'].tolist(), C=C, strides=[8, 16, 32, 64], istrides=[4, 8, 16, 32],
               attn_heads=2, d=4, n_cls=2, n_reg=3, bbox_weights=[1.0] * 8 + [0.2, 0.2], scale_clamp=float(np.log(100000.0 / 16)),
               n_exp=4, n_p=int(z['P']), stages=int(z['stages']))
    return params, pf, imf, cfg


@pytest.mark.parametrize('tag', ['lidar', 'fusion'])
def test_srfdet_head_golden(golden_dir, tag):
    """Oracle restatement of SRFDetHead (DPG, chained stages, decode) vs the reference's own outputs."""
    z = _load(golden_dir, 'srfdet_head.npz')
    params, pf, imf, cfg = _head_inputs(z, tag)
    b0, f0 = O.dpg_init_proposals(params, imf, pf, cfg['n_exp'], cfg['n_p'], imf is not None)
    np.testing.assert_allclose(b0, z[f'{tag}.init_boxes'], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(f0, z[f'{tag}.init_feats'], rtol=1e-4, atol=1e-5)
    logits, boxes = O.srfdet_head_forward(params, imf, pf, z['lidar2img'], cfg)
    # (fusion: one proposal's image rectangle sits within float rounding of a RoIAlign sampling threshold in the
    # C loop vs torchvision -- same effect as tests/test_gpu_roi_head.py::test_img_roi_production_size_vs_oracle)
    tol = 2e-4 if tag == 'lidar' else 1e-3
    np.testing.assert_allclose(logits, z[f'{tag}.logits'], rtol=0, atol=tol)
    np.testing.assert_allclose(boxes, z[f'{tag}.boxes'], rtol=0, atol=tol)
    scores, dec = O.decode_boxes(logits[-1], boxes[-1])
    sc, idx = torch.as_tensor(scores[0]).flatten().topk(20)
    bx = torch.as_tensor(dec[0])[idx // 10]
    rng = torch.tensor([-40.0, -40.0, -10.0, 40.0, 40.0, 10.0])
    m = (bx[:, :3] >= rng[:3]).all(1) & (bx[:, :3] <= rng[3:]).all(1)
    np.testing.assert_array_equal((idx % 10)[m].numpy(), z[f'{tag}.det_labels'])
    np.testing.assert_allclose(bx[m].numpy(), z[f'{tag}.det_boxes'], rtol=0, atol=tol)
    np.testing.assert_allclose(sc[m].numpy(), z[f'{tag}.det_scores'], rtol=0, atol=1e-5)


def test_second_fpn_golden(golden_dir):
    """Oracle SECONDCustom (reference-owned forward) + FPN restatement vs the fixture."""
    z = _load(golden_dir, 'second_fpn.npz')
    feats = O.second_custom({k[2:]: z[k] for k in z.files if k.startswith('b.')}, z['x'], [2, 2], [1, 2])
    for i, f in enumerate(feats):
        np.testing.assert_allclose(f, z[f'feat{i}'], rtol=0, atol=1e-5)
    outs = O.fpn({k[2:]: z[k] for k in z.files if k.startswith('n.')}, feats, 4)
    for i, o in enumerate(outs):
        np.testing.assert_allclose(o, z[f'out{i}'], rtol=0, atol=1e-5)
