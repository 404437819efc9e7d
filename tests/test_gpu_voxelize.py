"""CUDA voxelization vs the CPU oracle (bit-exact integer outputs), through the C ABI."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from srfdet_b200 import synth
from util import cuda

pytestmark = pytest.mark.gpu


def _vox(kind, **kw):
    from srfdet_b200.plugin import Voxelization
    g = synth.GEOM[kind]
    args = dict(voxel_size=g['voxel_size'], point_cloud_range=g['pc_range'], max_num_points=-1, max_voxels=(-1, -1))
    args.update(kw)
    return Voxelization(**args).eval(), g


@pytest.mark.parametrize('kind', ['kitti', 'nusc', 'waymo'])
def test_dynamic_voxelize_full_size(kind):
    vox, g = _vox(kind)
    assert vox.grid_size == O.grid_size(g['voxel_size'], g['pc_range']).tolist()
    pts = synth.cloud(kind, 1)
    ref = O.dynamic_voxelize(pts, g['voxel_size'], g['pc_range'])
    got = vox(cuda(pts)).cpu().numpy()
    np.testing.assert_array_equal(got, ref)
    got4 = vox.dynamic(cuda(pts), batch_idx=3).cpu().numpy()
    np.testing.assert_array_equal(got4[:, 1:], ref)
    assert (got4[:, 0] == 3).all()
    assert (ref[:, 0] < 0).sum() > 100     # the cloud does exercise out-of-range points


def test_dynamic_voxelize_boundaries():
    vox, g = _vox('nusc')
    lo, hi = np.array(g['pc_range'][:3], np.float32), np.array(g['pc_range'][3:], np.float32)
    eps = np.float32(1e-6)
    pts = np.array([[lo[0], lo[1], lo[2], 0, 0], [hi[0], 0, 0, 0, 0], [np.nextafter(hi[0], -np.inf, dtype=np.float32), 0, 0, 0, 0],
                    [0, hi[1], 0, 0, 0], [0, 0, hi[2], 0, 0], [np.nextafter(lo[0], -np.inf, dtype=np.float32), 0, 0, 0, 0],
                    [0.075, 0.15, 0.2, 0, 0], [0.0749999, -0.0750001, -0.2000001, 0, 0], [1e9, 0, 0, 0, 0], [0, -1e9, 0, 0, 0]], np.float32)
    ref = O.dynamic_voxelize(pts, g['voxel_size'], g['pc_range'])
    np.testing.assert_array_equal(vox(cuda(pts)).cpu().numpy(), ref)


def _check_hard(pts, kind, T, mv):
    vox, g = _vox(kind, max_num_points=T, max_voxels=(mv, mv))
    rv, rc, rn, rp = O.hard_voxelize(pts, g['voxel_size'], g['pc_range'], T, mv)
    o = vox.hard_padded(cuda(pts), want_voxels=True, want_mean=True, want_p2v=True)
    m = int(o['count'].item())
    assert m == len(rc)
    np.testing.assert_array_equal(o['coors'][:m].cpu().numpy(), rc)
    np.testing.assert_array_equal(o['num_points'][:m].cpu().numpy(), rn)
    np.testing.assert_array_equal(o['point2voxel'].cpu().numpy(), rp)
    np.testing.assert_array_equal(o['voxels'][:m].cpu().numpy(), rv)      # payload incl. zero padding, bit-exact
    ref_mean = O.hard_simple_vfe(rv, rn, pts.shape[1])
    np.testing.assert_allclose(o['mean'][:m].cpu().numpy(), ref_mean, rtol=2e-6, atol=1e-6)
    # module contract (mmcv): sliced outputs
    v, c, n = vox(cuda(pts))
    assert v.shape == rv.shape and c.shape == rc.shape and n.shape == rn.shape
    return m, rn


def test_hard_voxelize_nusc_full_size():
    m, rn = _check_hard(synth.cloud('nusc', 2), 'nusc', 10, 160000)
    assert m > 100000


def test_hard_voxelize_overflow_max_voxels_and_points():
    pts = synth.dense_cloud('nusc', 3, 9000, extent=0.3)
    m, rn = _check_hard(pts, 'nusc', 10, 300)
    assert m == 300 and rn.max() == 10
    # big cloud that overflows max_voxels at production settings
    pts = synth.cloud('nusc', 4, n_points=400000)
    m, rn = _check_hard(pts, 'nusc', 10, 120000)
    assert m == 120000


def test_hard_voxelize_micro_and_empty():
    from srfdet_b200.plugin import Voxelization
    vs, rng_ = [1.0, 1.0, 1.0], [0, 0, 0, 4, 4, 2]
    pts = np.array([[0.5, 0.5, 0.5, 9], [3.5, 0.5, 1.5, 8], [0.6, 0.4, 0.2, 7], [4.0, 1, 1, 6], [0, 0, 0, 5],
                    [-0.001, 1, 1, 4], [0.7, 0.7, 0.7, 3], [2.5, 2.5, 0.5, 2]], np.float32)
    vox = Voxelization(vs, rng_, 3, (2, 2)).eval()
    o = vox.hard_padded(cuda(pts), want_p2v=True)
    assert int(o['count']) == 2
    np.testing.assert_array_equal(o['coors'][:2].cpu().numpy(), [[0, 0, 0], [1, 0, 3]])
    np.testing.assert_array_equal(o['num_points'][:2].cpu().numpy(), [3, 1])
    np.testing.assert_array_equal(o['point2voxel'].cpu().numpy(), [0, 1, 0, -1, 0, -1, -1, -1])
    np.testing.assert_array_equal(o['voxels'][0, :, 3].cpu().numpy(), [9, 7, 5])
    v, c, n = vox(torch.zeros((0, 4), device='cuda'))
    assert v.shape[0] == 0 and c.shape[0] == 0 and n.shape[0] == 0


def test_detector_voxelize_batched():
    """SRFDet.voxelize batch padding (detectors/srfdet.py:218-247) for a 2-sample batch."""
    from srfdet_b200.plugin import SRFDetPointPath
    g = synth.GEOM['nusc']
    cfg = dict(pts_voxel_layer=dict(max_num_points=10, voxel_size=g['voxel_size'], max_voxels=(120000, 160000), point_cloud_range=g['pc_range']),
               pts_voxel_encoder=dict(type='HardSimpleVFE', num_features=5),
               pts_middle_encoder=dict(type='SparseEncoderCustom', in_channels=5, sparse_shape=g['sparse_shape']))
    det = SRFDetPointPath(**cfg)
    clouds = [synth.cloud('nusc', 5, n_points=30000), synth.cloud('nusc', 6, n_points=20000)]
    rv, rn, rc = O.detector_voxelize_hard(clouds, g['voxel_size'], g['pc_range'], 10, 160000)
    v, n, c = det.voxelize([cuda(p) for p in clouds])
    np.testing.assert_array_equal(c.cpu().numpy(), rc)
    np.testing.assert_array_equal(n.cpu().numpy(), rn)
    np.testing.assert_array_equal(v.cpu().numpy(), rv)
