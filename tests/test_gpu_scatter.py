"""DynamicScatter / DynamicVFECustom on CUDA vs the oracle and the reference goldens."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O
from srfdet_b200 import synth
from util import cuda, randomize_bn_, rel_err

pytestmark = pytest.mark.gpu


def _coors(kind, pts, batch=0):
    g = synth.GEOM[kind]
    c = O.dynamic_voxelize(pts, g['voxel_size'], g['pc_range'])
    return np.concatenate([np.full((len(c), 1), batch, np.int32), c], 1)


@pytest.mark.parametrize('mode', ['max', 'mean'])
@pytest.mark.parametrize('cd', [3, 4])
def test_dynamic_scatter_waymo_full(mode, cd):
    from srfdet_b200.plugin import DynamicScatter
    g = synth.GEOM['waymo']
    pts = synth.cloud('waymo', 7)
    coors = _coors('waymo', pts)
    if cd == 4:
        coors[len(coors) // 2:, 0] = 1      # two samples in the batch (valid because the batch index is sorted)
    else:
        coors = coors[:, 1:].copy()
    feats = pts - np.array([0, 0, 0, 0.5, 0.5], np.float32)   # negative values exercise the signed max
    rf, rc, rp = O.dynamic_scatter(feats, coors, mode)
    sc = DynamicScatter(g['voxel_size'], g['pc_range'], mode == 'mean')
    f, c, count, p2v = sc.forward_padded(cuda(feats), cuda(coors), want_p2v=True)
    m = int(count)
    assert m == len(rc)
    np.testing.assert_array_equal(c[:m].cpu().numpy(), rc)
    np.testing.assert_array_equal(p2v.cpu().numpy(), rp)
    if mode == 'max':
        np.testing.assert_array_equal(f[:m].cpu().numpy(), rf)
    else:
        np.testing.assert_allclose(f[:m].cpu().numpy(), rf, rtol=1e-5, atol=1e-5)
    f2, c2 = sc(cuda(feats), cuda(coors))
    assert f2.shape == rf.shape and c2.shape == rc.shape


def _load_vfe(z, cfg_extra):
    from srfdet_b200.plugin import DynamicVFECustom
    sd = {k[2:]: torch.as_tensor(z[k]) for k in z.files if k.startswith('p.')}
    nl = len({k.split('.')[1] for k in sd if k.startswith('vfe_layers.')})
    c_out = [int(sd[f'vfe_layers.{i}.linear.weight'].shape[0]) for i in range(nl)]
    cin = int(z['points'].shape[1])
    vfe = DynamicVFECustom(in_channels=cin, feat_channels=c_out, with_cluster_center=True, with_voxel_center=True,
                           voxel_size=z['voxel_size'].tolist(), point_cloud_range=z['pc_range'].tolist(),
                           norm_cfg=dict(type='naiveSyncBN1dCustom', eps=1e-3, momentum=0.01)).eval()
    missing = vfe.load_state_dict(sd, strict=True)
    return vfe.cuda()


@pytest.mark.parametrize('tag', ['waymo', 'kitti'])
def test_dynamic_vfe_reference_golden(golden_dir, tag):
    """Golden produced by the reference's DynamicVFECustom.forward (tests/golden/make_golden.py)."""
    z = np.load(os.path.join(golden_dir, f'vfe_{tag}.npz'))
    vfe = _load_vfe(z, {})
    vf, vc = vfe(cuda(z['points']), cuda(z['coors']))
    np.testing.assert_array_equal(vc.cpu().numpy(), z['voxel_coors'])
    assert rel_err(vf.cpu().numpy(), z['voxel_feats']) < 1e-4


@pytest.mark.parametrize('kind', ['waymo', 'kitti'])
def test_dynamic_vfe_full_size_vs_oracle(kind):
    from srfdet_b200.plugin import SRFDetPointPath
    cfgp = {'waymo': 'configs/waymo/srfdet_dvoxel_waymo_L.py', 'kitti': 'configs/kitti/srfdet_voxel_kitti_L.py'}[kind]
    g = synth.GEOM[kind]
    cin = g['in_channels']
    feat_channels = [5, 5] if kind == 'waymo' else [4]
    from srfdet_b200.plugin import DynamicVFECustom
    torch.manual_seed(3)
    vfe = DynamicVFECustom(in_channels=cin, feat_channels=feat_channels, with_cluster_center=True, with_voxel_center=True,
                           voxel_size=g['voxel_size'], point_cloud_range=g['pc_range'],
                           norm_cfg=dict(type='naiveSyncBN1dCustom', eps=1e-3, momentum=0.01)).eval()
    randomize_bn_(vfe, 4)
    pts = synth.cloud(kind, 8)
    coors = _coors(kind, pts)
    sd = {k: v.numpy() for k, v in vfe.state_dict().items()}
    params = {}
    for k, v in sd.items():
        if k.startswith('cen2point_pos_enc.'):
            params['pos.' + k[len('cen2point_pos_enc.'):]] = v
        elif k.startswith('vfe_layers.'):
            params['vfe.' + k[len('vfe_layers.'):]] = v
    rf, rc = O.dynamic_vfe_custom(params, pts, coors, g['voxel_size'], g['pc_range'])
    vf, vc = vfe.cuda()(cuda(pts), cuda(coors))
    np.testing.assert_array_equal(vc.cpu().numpy(), rc)
    assert rel_err(vf.cpu().numpy(), rf) < 1e-4
    assert len(rc) > 40000


# ---------------------------------------------------------------- pillar path
@pytest.mark.parametrize('tag,kw', [('new', dict(legacy=False)), ('legacy', dict(legacy=True, with_distance=True)),
                                    ('avg', dict(legacy=False, mode='avg'))])
def test_pillar_vfe_golden(golden_dir, tag, kw):
    """PillarFeatureNetCustom.forward (pillar_encoder_custom.py:95-161) vs the reference's own output."""
    import os
    from srfdet_b200.plugin import PillarFeatureNetCustom
    z = np.load(os.path.join(golden_dir, 'pillar_vfe.npz'))
    net = PillarFeatureNetCustom(in_channels=5, feat_channels=[64], voxel_size=[0.2, 0.2, 8],
                                 point_cloud_range=[-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], **kw).eval()
    net.load_state_dict({k[len(tag) + 3:]: torch.as_tensor(z[k]) for k in z.files if k.startswith(tag + '.p.')}, strict=True)
    out = net.cuda()(cuda(z['voxels']), cuda(z['num_points']), cuda(z['coors']))
    np.testing.assert_allclose(out.cpu().numpy(), z[f'{tag}.out'], rtol=1e-5, atol=2e-5)


def test_pillar_path_vs_oracle():
    """hard voxelization (T=20, 40000 pillars, configs/nus/srfdet_pillar_nusc_L.py:37-54) ->
    PillarFeatureNetCustom -> PointPillarsScatter at full size vs the oracle; no host sync."""
    from oracle import oracle as O
    from srfdet_b200 import synth
    from srfdet_b200.plugin import PillarFeatureNetCustom, PointPillarsScatter, Voxelization
    vs, pc = [0.2, 0.2, 8], [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]
    pts = synth.cloud('nusc', 77)
    torch.manual_seed(5)
    net = PillarFeatureNetCustom(in_channels=5, feat_channels=[64], voxel_size=vs, point_cloud_range=pc, legacy=False).eval()
    from util import randomize_bn_
    randomize_bn_(net, 6)
    v, c, n, _ = O.hard_voxelize(pts, vs, pc, 20, 40000)
    coors = np.concatenate([np.zeros((len(c), 1), np.int32), c], 1)
    ref = O.pillar_feature_net({k: t.numpy() for k, t in net.state_dict().items()}, v, n, coors, vs, pc, legacy=False)
    ref_canvas = O.pillars_scatter(ref, coors, 1, 512, 512)
    vox = Voxelization(vs, pc, 20, (40000, 40000)).eval()
    o = vox.hard_padded(cuda(pts), batch_idx=0)
    m = int(o['count'])
    assert m == len(c)
    np.testing.assert_array_equal(o['coors'][:m].cpu().numpy(), coors)
    feats = net.cuda()(o['voxels'], o['num_points'], o['coors'], num_voxels=o['count'])
    np.testing.assert_allclose(feats[:m].cpu().numpy(), ref, rtol=1e-5, atol=2e-5)
    for cl in (False, True):
        sc = PointPillarsScatter(64, (512, 512))
        sc.channels_last = cl
        canvas = sc(feats, o['coors'], batch_size=1, num_voxels=o['count'])
        assert canvas.shape == (1, 64, 512, 512)
        np.testing.assert_allclose(canvas.cpu().numpy(), ref_canvas, rtol=1e-5, atol=2e-5)


def test_dynamic_scatter_max_of_negative_zero():
    """A voxel channel whose maximum is exactly -0.0 must come out as (-)0.0, not -inf (sign-bit branch of the float atomic max)."""
    from srfdet_b200.plugin import DynamicScatter
    sc = DynamicScatter([0.5, 0.5, 0.5], [0, 0, 0, 4, 4, 4], average_points=False)
    feats = torch.tensor([[-0.0, -1.0], [-2.0, -0.0], [-3.0, -3.0]], device='cuda')
    coors = torch.tensor([[1, 1, 1], [1, 1, 1], [1, 1, 1]], dtype=torch.int32, device='cuda')
    vf, vc = sc(feats, coors)
    assert vf.shape == (1, 2)
    assert torch.equal(vf.abs().cpu(), torch.zeros(1, 2)) and bool(torch.isfinite(vf).all())
