"""CPU oracle for the SRFDet3D point-cloud -> region-feature hot path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under srfdet_b200/ imports this module.

Each function restates one piece of the reference path and cites it (paths relative to
/root/reference).  "[3P]" marks third-party semantics (mmcv-full 1.7.0, mmdet 2.28.2,
mmdet3d 1.0.0rc6, spconv 2.x -- none installed / vendored) restated from their published
algorithms; the reference-owned functions are pinned by tests/golden (see make_golden.py).
Loop-heavy integer work lives in srf_oracle.c (ctypes); dense math uses torch CPU fp32.
"""
import ctypes
import math
import os
import subprocess

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_c(force=False):
    """Compile srf_oracle.c -> libsrf_oracle.so (gcc).  Returns the path."""
    so = os.path.join(_HERE, 'libsrf_oracle.so')
    src = os.path.join(_HERE, 'srf_oracle.c')
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(['make', '-s', '-C', _HERE, '-B', 'libsrf_oracle.so'])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_c())
        _LIB.orc_hard_voxelize.restype = ctypes.c_int
        _LIB.orc_strided_out_coors.restype = ctypes.c_int
        _LIB.orc_rulebook_subm.restype = ctypes.c_int
        _LIB.orc_rulebook_strided.restype = ctypes.c_int
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f3(v):
    return np.ascontiguousarray(np.asarray(v, dtype=np.float32))


def _i3(v):
    return np.ascontiguousarray(np.asarray(v, dtype=np.int32))


# --------------------------------------------------------------------------------------
# a1 / a2: voxelization  ([3P] mmcv Voxelization; call sites detectors/srfdet.py:58,221,238)
# --------------------------------------------------------------------------------------
def grid_size(voxel_size, pc_range):
    g = np.zeros(3, np.int32)
    _lib().orc_grid_size(_p(_f3(voxel_size)), _p(_f3(pc_range)), _p(g))
    return g  # (gx, gy, gz)


def dynamic_voxelize(points, voxel_size, pc_range):
    """points (N,C) f32 -> coors (N,3) int32 (z,y,x), (-1,-1,-1) when out of range."""
    points = np.ascontiguousarray(points, np.float32)
    n, c = points.shape
    g = grid_size(voxel_size, pc_range)
    coors = np.empty((n, 3), np.int32)
    _lib().orc_dynamic_voxelize(_p(points), n, c, _p(_f3(voxel_size)), _p(_f3(pc_range)), _p(g),
                                _p(coors))
    return coors


def hard_voxelize(points, voxel_size, pc_range, max_points, max_voxels):
    """-> voxels (M,T,C), coors (M,3) zyx, num_points (M,), point2voxel (N,)."""
    points = np.ascontiguousarray(points, np.float32)
    n, c = points.shape
    g = grid_size(voxel_size, pc_range)
    voxels = np.zeros((max_voxels, max_points, c), np.float32)
    coors = np.zeros((max_voxels, 3), np.int32)
    num = np.zeros((max_voxels,), np.int32)
    p2v = np.empty((n,), np.int32)
    m = _lib().orc_hard_voxelize(_p(points), n, c, _p(_f3(voxel_size)), _p(_f3(pc_range)), _p(g),
                                 int(max_points), int(max_voxels), _p(voxels), _p(coors), _p(num),
                                 _p(p2v))
    assert m >= 0
    return voxels[:m], coors[:m], num[:m], p2v


def detector_voxelize_hard(points_list, voxel_size, pc_range, max_points, max_voxels):
    """SRFDet.voxelize, hard branch (detectors/srfdet.py:218-232): per-sample voxelize,
    concatenate, pad the batch index in front of the coords."""
    vs, cs, ns = [], [], []
    for b, pts in enumerate(points_list):
        v, c, n, _ = hard_voxelize(pts, voxel_size, pc_range, max_points, max_voxels)
        vs.append(v)
        ns.append(n)
        cs.append(np.concatenate([np.full((len(c), 1), b, np.int32), c], 1))
    return np.concatenate(vs), np.concatenate(ns), np.concatenate(cs)


def detector_voxelize_dynamic(points_list, voxel_size, pc_range):
    """SRFDet.voxelize, dynamic branch (detectors/srfdet.py:233-247)."""
    cs = []
    for b, pts in enumerate(points_list):
        c = dynamic_voxelize(pts, voxel_size, pc_range)
        cs.append(np.concatenate([np.full((len(c), 1), b, np.int32), c], 1))
    return np.concatenate(points_list).astype(np.float32), np.concatenate(cs)


# --------------------------------------------------------------------------------------
# a3: HardSimpleVFE ([3P] mmdet3d; cfg configs/nus/srfdet_voxel_nusc_L.py:40)
# --------------------------------------------------------------------------------------
def hard_simple_vfe(voxels, num_points, num_features=5):
    v = torch.as_tensor(voxels)
    n = torch.as_tensor(num_points)
    return (v[:, :, :num_features].sum(dim=1) / n.type_as(v).view(-1, 1)).numpy()


# --------------------------------------------------------------------------------------
# DynamicScatter ([3P] mmcv; call sites voxel_encoders/voxel_encoder.py:82,99-102,189,232)
# --------------------------------------------------------------------------------------
def _dynamic_scatter_single(feats, coors, mode):
    feats = torch.as_tensor(feats, dtype=torch.float32)
    coors = torch.as_tensor(coors).to(torch.int64)
    bad = (coors < 0).any(-1, keepdim=True)
    clean = coors.masked_fill(bad, -1)
    out_coors, inv, counts = torch.unique(clean, dim=0, sorted=True, return_inverse=True,
                                          return_counts=True)
    if out_coors.numel() and bool(out_coors[0, 0] < 0):
        out_coors = out_coors[1:]
        counts = counts[1:]
        inv = inv - 1
    m, c = out_coors.shape[0], feats.shape[1]
    keep = inv >= 0
    idx = inv[keep].view(-1, 1).expand(-1, c)
    if mode == 'max':
        red = torch.full((m, c), -float('inf'))
        red.scatter_reduce_(0, idx, feats[keep], reduce='amax', include_self=True)
    else:
        red = torch.zeros((m, c), dtype=torch.float64)
        red.scatter_add_(0, idx, feats[keep].double())
        red = (red / counts.view(-1, 1).double()).float()
    return red, out_coors.to(torch.int32), inv


def dynamic_scatter(feats, coors, mode):
    """mode in {'max','mean'}.  coors (N,3) or (N,4 with leading batch index).
    -> (voxel_feats (M,C), voxel_coors (M,3|4) int32 sorted, point2voxel (N,) or -1)."""
    coors_t = torch.as_tensor(coors)
    feats_t = torch.as_tensor(feats, dtype=torch.float32)
    if coors_t.shape[-1] == 3:
        r, c, inv = _dynamic_scatter_single(feats_t, coors_t, mode)
        return r.numpy(), c.numpy(), inv.numpy()
    bs = int(coors_t[-1, 0]) + 1
    outs, ocs = [], []
    p2v = torch.full((coors_t.shape[0],), -1, dtype=torch.int64)
    base = 0
    for b in range(bs):
        inds = torch.where(coors_t[:, 0] == b)[0]
        r, c, inv = _dynamic_scatter_single(feats_t[inds], coors_t[inds][:, 1:], mode)
        outs.append(r)
        ocs.append(F.pad(c, (1, 0), value=b))
        p2v[inds] = torch.where(inv >= 0, inv + base, inv)
        base += r.shape[0]
    return torch.cat(outs).numpy(), torch.cat(ocs).numpy(), p2v.numpy()


# --------------------------------------------------------------------------------------
# a4: DynamicVFECustom.forward (voxel_encoders/voxel_encoder.py:162-240), eval mode.
# params: dict with
#   'pos.0.weight' (32,3), 'pos.1.{weight,bias,running_mean,running_var}', 'pos.3.weight'
#   (32,32), 'pos.4.*'  -> cen2point_pos_enc (:107-116), BatchNorm1d default eps 1e-5
#   'vfe.{i}.linear.weight', 'vfe.{i}.norm.*' -> DynamicVFELayer (utils.py:8-45), eps 1e-3
# --------------------------------------------------------------------------------------
def _bn_eval(x, p, prefix, eps):
    return F.batch_norm(x, p[prefix + 'running_mean'], p[prefix + 'running_var'],
                        p[prefix + 'weight'], p[prefix + 'bias'], False, 0.0, eps)


def dynamic_vfe_custom(params, features, coors, voxel_size, pc_range, vfe_eps=1e-3):
    p = {k: torch.as_tensor(v, dtype=torch.float32) for k, v in params.items()}
    feats = torch.as_tensor(features, dtype=torch.float32)
    coors_t = torch.as_tensor(coors)
    vx, vy, vz = [float(v) for v in voxel_size]
    x_off, y_off, z_off = vx / 2 + pc_range[0], vy / 2 + pc_range[1], vz / 2 + pc_range[2]
    # cluster centre (:188-193); the canvas lookup of map_voxel_center_to_point (:118-158)
    # is the inverse map of the scatter for every valid point (invalid points are dropped
    # by the later scatter, so what they read is irrelevant).
    vmean, _, p2v = dynamic_scatter(feats, coors_t, 'mean')
    p2v_t = torch.as_tensor(p2v)
    valid = p2v_t >= 0
    safe = p2v_t.clamp(min=0)
    pts_mean = torch.as_tensor(vmean)[safe]
    f_cluster = feats[:, :3] - pts_mean[:, :3]
    h = F.linear(f_cluster, p['pos.0.weight'])
    h = torch.tanh(_bn_eval(h, p, 'pos.1.', 1e-5))
    h = F.linear(h, p['pos.3.weight'])
    h = torch.tanh(_bn_eval(h, p, 'pos.4.', 1e-5))
    cf = coors_t.to(torch.float32)
    f_center = torch.stack([feats[:, 0] - (cf[:, 3] * vx + x_off),
                            feats[:, 1] - (cf[:, 2] * vy + y_off),
                            feats[:, 2] - (cf[:, 1] * vz + z_off)], 1)
    x = torch.cat([feats, h, f_center], -1)
    n_vfe = len([k for k in p if k.startswith('vfe.') and k.endswith('linear.weight')])
    voxel_feats = voxel_coors = None
    for i in range(n_vfe):
        pf = F.relu(_bn_eval(F.linear(x, p[f'vfe.{i}.linear.weight']), p, f'vfe.{i}.norm.', vfe_eps))
        pf_valid = torch.where(valid.view(-1, 1), pf, torch.zeros_like(pf))
        voxel_feats, voxel_coors, _ = dynamic_scatter(pf_valid, coors_t, 'max')
        if i != n_vfe - 1:
            per_point = torch.as_tensor(voxel_feats)[safe]
            x = torch.cat([pf, per_point], 1)
    return voxel_feats, voxel_coors


# --------------------------------------------------------------------------------------
# a5: SparseEncoderCustom (middle_encoders/sparse_encoder_custom.py)
# --------------------------------------------------------------------------------------
def _t3(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v, v)


def encoder_layer_plan(in_channels, base_channels=16, output_channels=128,
                       encoder_channels=((16,), (32, 32, 32), (64, 64, 64), (64, 64, 64)),
                       encoder_paddings=((1,), (1, 1, 1), (1, 1, 1), ((0, 1, 1), 1, 1)),
                       block_type='conv_module'):
    """Flat list of sparse convs in execution order.  Restates the constructor
    (sparse_encoder_custom.py:73-107) and make_encoder_layers (:142-216); each entry:
    name (state-dict prefix), kind 'subm'|'spconv', cin, cout, ksize, stride, pad, key,
    relu, save_identity / add_identity ([3P] SparseBasicBlock: conv-bn-relu-conv-bn-(+x)-relu).
    """
    plan = [dict(name='conv_input', kind='subm', cin=in_channels, cout=base_channels,
                 ksize=(3, 3, 3), stride=(1, 1, 1), pad=(1, 1, 1), key='subm1', relu=True,
                 conv='0', bn='1')]
    cin = base_channels
    nst = len(encoder_channels)
    for i, blocks in enumerate(encoder_channels):
        blocks = tuple(blocks)
        for j, cout in enumerate(blocks):
            pad = tuple(encoder_paddings[i])[j]
            pre = f'encoder_layers.encoder_layer{i + 1}.{j}'
            if i != 0 and j == 0 and block_type == 'conv_module':
                plan.append(dict(name=pre, kind='spconv', cin=cin, cout=cout, ksize=(3, 3, 3),
                                 stride=(2, 2, 2), pad=_t3(pad), key=f'spconv{i + 1}', relu=True,
                                 conv='0', bn='1'))
            elif block_type == 'basicblock':
                if j == len(blocks) - 1 and i != nst - 1:
                    plan.append(dict(name=pre, kind='spconv', cin=cin, cout=cout, ksize=(3, 3, 3),
                                     stride=(2, 2, 2), pad=_t3(pad), key=f'spconv{i + 1}',
                                     relu=True, conv='0', bn='1'))
                else:
                    plan.append(dict(name=pre, kind='subm', cin=cout, cout=cout, ksize=(3, 3, 3),
                                     stride=(1, 1, 1), pad=(1, 1, 1), key=None, relu=True,
                                     conv='conv1', bn='bn1', save_identity=True))
                    plan.append(dict(name=pre, kind='subm', cin=cout, cout=cout, ksize=(3, 3, 3),
                                     stride=(1, 1, 1), pad=(1, 1, 1), key=None, relu=True,
                                     conv='conv2', bn='bn2', add_identity=True))
            else:
                plan.append(dict(name=pre, kind='subm', cin=cin, cout=cout, ksize=(3, 3, 3),
                                 stride=(1, 1, 1), pad=_t3(pad), key=f'subm{i + 1}', relu=True,
                                 conv='0', bn='1'))
            cin = cout
    plan.append(dict(name='conv_out', kind='spconv', cin=cin, cout=output_channels,
                     ksize=(3, 1, 1), stride=(2, 1, 1), pad=(0, 0, 0), key='spconv_down2',
                     relu=True, conv='0', bn='1'))
    return plan


def rulebook_subm(coors, batch, dims, ksize=(3, 3, 3)):
    """-> list over kernel offsets of (in_rows, out_rows) int32 arrays (sorted by out row)."""
    coors = np.ascontiguousarray(coors, np.int32)
    n = coors.shape[0]
    kv = int(np.prod(ksize))
    pin = np.empty((kv, max(n, 1)), np.int32)
    pout = np.empty((kv, max(n, 1)), np.int32)
    cnt = np.zeros(kv, np.int32)
    rc = _lib().orc_rulebook_subm(_p(coors), n, int(batch), _p(_i3(dims)), _p(_i3(ksize)), _p(pin),
                                  _p(pout), _p(cnt))
    assert rc == 0
    return [(pin[k, :cnt[k]].copy(), pout[k, :cnt[k]].copy()) for k in range(kv)]


def out_shape(dims, ksize, stride, pad):
    return tuple((d + 2 * p - k) // s + 1 for d, k, s, p in zip(dims, ksize, stride, pad))


def rulebook_strided(coors, batch, dims, ksize, stride, pad):
    """-> (out_coors (M,4) sorted by linear index, out_dims, pairs list)."""
    coors = np.ascontiguousarray(coors, np.int32)
    n = coors.shape[0]
    kv = int(np.prod(ksize))
    oc = np.empty((max(n * kv, 1), 4), np.int32)
    m = _lib().orc_strided_out_coors(_p(coors), n, _p(_i3(dims)), _p(_i3(ksize)), _p(_i3(stride)),
                                     _p(_i3(pad)), _p(oc))
    oc = np.ascontiguousarray(oc[:m])
    pin = np.empty((kv, max(n, 1)), np.int32)
    pout = np.empty((kv, max(n, 1)), np.int32)
    cnt = np.zeros(kv, np.int32)
    rc = _lib().orc_rulebook_strided(_p(coors), n, _p(oc), m, int(batch), _p(_i3(dims)),
                                     _p(_i3(ksize)), _p(_i3(stride)), _p(_i3(pad)), _p(pin),
                                     _p(pout), _p(cnt))
    assert rc == 0
    pairs = [(pin[k, :cnt[k]].copy(), pout[k, :cnt[k]].copy()) for k in range(kv)]
    return oc, out_shape(dims, ksize, stride, pad), pairs


def sparse_conv_c(feats, w_kio, pairs, n_out):
    """Naive C gather-GEMM-scatter (small cases).  w_kio: (kvol, cin, cout)."""
    feats = np.ascontiguousarray(feats, np.float32)
    w = np.ascontiguousarray(w_kio, np.float32)
    kv, cin, cout = w.shape
    stride = max(max((len(a) for a, _ in pairs), default=1), 1)
    pin = np.zeros((kv, stride), np.int32)
    pout = np.zeros((kv, stride), np.int32)
    cnt = np.zeros(kv, np.int32)
    for k, (a, b) in enumerate(pairs):
        pin[k, :len(a)] = a
        pout[k, :len(b)] = b
        cnt[k] = len(a)
    out = np.zeros((n_out, cout), np.float32)
    _lib().orc_sparse_conv(_p(feats), cin, _p(w), cout, kv, stride, _p(pin), _p(pout), _p(cnt),
                           _p(out))
    return out


def sparse_conv_mm(feats, w_kio, pairs, n_out):
    """spconv 'native' algorithm with torch.mm per offset (the CPU-baseline form)."""
    x = torch.as_tensor(feats, dtype=torch.float32)
    w = torch.as_tensor(w_kio, dtype=torch.float32)
    out = torch.zeros((n_out, w.shape[2]), dtype=torch.float32)
    for k, (a, b) in enumerate(pairs):
        if len(a) == 0:
            continue
        out.index_add_(0, torch.as_tensor(b, dtype=torch.int64),
                       x[torch.as_tensor(a, dtype=torch.int64)] @ w[k])
    return out.numpy()


def spconv2_weight_to_kio(w):
    """spconv2 layout [cout, kz, ky, kx, cin] -> (kvol, cin, cout)."""
    w = torch.as_tensor(w, dtype=torch.float32)
    co, kz, ky, kx, ci = w.shape
    return w.permute(1, 2, 3, 4, 0).reshape(kz * ky * kx, ci, co).contiguous().numpy()


def sparse_encoder(params, plan, voxel_features, coors, batch_size, sparse_shape, bn_eps=1e-3,
                   conv_fn=sparse_conv_mm, return_intermediates=False):
    """SparseEncoderCustom.forward (sparse_encoder_custom.py:109-140), eval mode.
    params: state-dict style {f'{name}.{conv}.weight' (spconv2 layout), f'{name}.{bn}.*'}.
    Rulebooks are cached per indice_key like spconv; SubM convs with key None (inside
    SparseBasicBlock) rebuild -- identical result, so the cache keys on resolution here.
    -> dense (B, C*D, H, W) float32."""
    feats = np.ascontiguousarray(voxel_features, np.float32)
    cur_coors = np.ascontiguousarray(coors, np.int32)
    dims = tuple(int(d) for d in sparse_shape)
    subm_cache = {}
    identity = None
    inter = []
    for L in plan:
        w = spconv2_weight_to_kio(params[f"{L['name']}.{L['conv']}.weight"])
        if L['kind'] == 'subm':
            ck = (dims, cur_coors.shape[0])
            if ck not in subm_cache:
                subm_cache[ck] = rulebook_subm(cur_coors, batch_size, dims, L['ksize'])
            pairs = subm_cache[ck]
            n_out = cur_coors.shape[0]
            out_coors, out_dims = cur_coors, dims
        else:
            out_coors, out_dims, pairs = rulebook_strided(cur_coors, batch_size, dims, L['ksize'],
                                                          L['stride'], L['pad'])
            n_out = out_coors.shape[0]
        if L.get('save_identity'):
            identity = feats
        y = torch.as_tensor(conv_fn(feats, w, pairs, n_out))
        pre = f"{L['name']}.{L['bn']}."
        y = F.batch_norm(y, torch.as_tensor(params[pre + 'running_mean']),
                         torch.as_tensor(params[pre + 'running_var']),
                         torch.as_tensor(params[pre + 'weight']),
                         torch.as_tensor(params[pre + 'bias']), False, 0.0, bn_eps)
        if L.get('add_identity'):
            y = y + torch.as_tensor(identity)
        if L['relu']:
            y = F.relu(y)
        feats = y.numpy()
        cur_coors, dims = out_coors, out_dims
        if return_intermediates:
            inter.append((feats, cur_coors, dims))
    # SparseConvTensor.dense() ([3P] scatter_nd) + view (:135-138)
    c = feats.shape[1]
    dense = torch.zeros((batch_size, c, dims[0], dims[1], dims[2]), dtype=torch.float32)
    cc = torch.as_tensor(cur_coors, dtype=torch.int64)
    dense[cc[:, 0], :, cc[:, 1], cc[:, 2], cc[:, 3]] = torch.as_tensor(feats)
    dense = dense.view(batch_size, c * dims[0], dims[1], dims[2]).numpy()
    if return_intermediates:
        return dense, inter
    return dense


# --------------------------------------------------------------------------------------
# a6: boxes3d_to_corners3d (core/bbox/util.py:84-176), bottom_center=False, ry=False
# --------------------------------------------------------------------------------------
def boxes3d_to_corners3d(boxes):
    """boxes (B,P,>=8): cx,cy,cz,log w,log l,log h,sin,cos -> corners (B,P,8,3)."""
    b = torch.as_tensor(boxes, dtype=torch.float32)
    cx, cy, cz, w, l, h, s, c = [b[..., i] for i in range(8)]
    ry = torch.atan2(s, c)
    w, l, h = w.exp(), l.exp(), h.exp()
    sx = torch.tensor([1, -1, -1, 1, 1, -1, -1, 1], dtype=torch.float32)
    sy = torch.tensor([-1, -1, 1, 1, -1, -1, 1, 1], dtype=torch.float32)
    sz = torch.tensor([-1, -1, -1, -1, 1, 1, 1, 1], dtype=torch.float32)
    xc = (w / 2.).unsqueeze(-1) * sx
    yc = (l / 2.).unsqueeze(-1) * sy
    zc = (h / 2.).unsqueeze(-1) * sz
    cs, sn = torch.cos(ry).unsqueeze(-1), torch.sin(ry).unsqueeze(-1)
    # row-vector times R with R = [[c,-s,0],[s,c,0],[0,0,1]] (util.py:146-159)
    xr = xc * cs + yc * sn
    yr = -xc * sn + yc * cs
    return torch.stack([cx.unsqueeze(-1) + xr, cy.unsqueeze(-1) + yr, cz.unsqueeze(-1) + zc], -1)


def denorm_centres_(boxes, pc_range):
    """In-place centre de-normalisation (srfdet_head.py:2579-2587)."""
    b = torch.as_tensor(boxes)
    span = b.new_tensor([pc_range[3] - pc_range[0], pc_range[4] - pc_range[1],
                         pc_range[5] - pc_range[2]])
    lo = b.new_tensor(list(pc_range[:3]))
    b[..., :3] = b[..., :3] * span + lo
    return b


# --------------------------------------------------------------------------------------
# [3P] mmdet bbox2roi + SingleRoIExtractor (+ mmcv RoIAlign); cfg
# configs/nus/srfdet_voxel_nusc_LC.py:169-178
# --------------------------------------------------------------------------------------
def bbox2roi(bbox_list):
    rois = []
    for i, bb in enumerate(bbox_list):
        bb = torch.as_tensor(bb, dtype=torch.float32)
        rois.append(torch.cat([bb.new_full((bb.shape[0], 1), float(i)), bb[:, :4]], -1))
    return torch.cat(rois, 0)


def map_roi_levels(rois, num_levels, finest_scale=56):
    rois = torch.as_tensor(rois, dtype=torch.float32)
    scale = torch.sqrt((rois[:, 3] - rois[:, 1]) * (rois[:, 4] - rois[:, 2]))
    lvl = torch.floor(torch.log2(scale / finest_scale + 1e-6))
    return lvl.clamp(min=0, max=num_levels - 1).long()


def roi_align(feat, rois, spatial_scale, out_size=7, sampling_ratio=2):
    feat = np.ascontiguousarray(feat, np.float32)
    rois = np.ascontiguousarray(rois, np.float32)
    n, c, h, w = feat.shape
    k = rois.shape[0]
    out = np.zeros((k, c, out_size, out_size), np.float32)
    if not k:
        return out
    lib = _lib()

    def run(lo, hi):
        lib.orc_roi_align(_p(feat), n, c, h, w, _p(rois[lo:hi]), hi - lo, ctypes.c_float(spatial_scale), out_size, out_size,
                          sampling_ratio, _p(out[lo:hi]))
    nt = min(os.cpu_count() or 1, max(1, k // 32))
    if nt <= 1:
        run(0, k)
    else:       # RoIs are independent: one slice per host thread (ctypes releases the GIL; results do not depend on nt)
        from concurrent.futures import ThreadPoolExecutor
        edges = np.linspace(0, k, nt + 1).astype(int)
        with ThreadPoolExecutor(nt) as ex:
            list(ex.map(lambda i: run(int(edges[i]), int(edges[i + 1])), range(nt)))
    return out


def single_roi_extractor(feats, rois, strides, out_size=7, sampling_ratio=2, finest_scale=56,
                         roi_align_fn=roi_align):
    rois_t = torch.as_tensor(rois, dtype=torch.float32)
    nl = len(strides)
    c = feats[0].shape[1]
    out = np.zeros((rois_t.shape[0], c, out_size, out_size), np.float32)
    lv = map_roi_levels(rois_t, nl, finest_scale)
    for i in range(nl):
        inds = torch.nonzero(lv == i).flatten()
        if inds.numel():
            out[inds.numpy()] = roi_align_fn(np.asarray(feats[i]), rois_t[inds].numpy(),
                                             1.0 / strides[i], out_size, sampling_ratio)
    return out


# --------------------------------------------------------------------------------------
# a7: points_feats_sampling_bboxes_roi (srfdet_head.py:2568-2629 / :1627-1688)
# --------------------------------------------------------------------------------------
def bev_rois(boxes, pc_range, voxel_size, mutate=True):
    """Mutates `boxes` centres in place like the reference (:2587) when mutate=True."""
    b = torch.as_tensor(boxes) if mutate else torch.as_tensor(boxes).clone()
    denorm_centres_(b, pc_range)
    cor = boxes3d_to_corners3d(b[..., :8])
    cor = cor - cor.new_tensor(list(pc_range[:3]))
    x = cor[..., 0] / voxel_size[0]
    y = cor[..., 1] / voxel_size[1]
    rect = torch.stack([x.min(-1).values, y.min(-1).values, x.max(-1).values, y.max(-1).values], -1)
    return bbox2roi([rect[i] for i in range(rect.shape[0])])


def points_roi_feats(point_feats, boxes, pc_range, voxel_size, strides, mutate=True):
    rois = bev_rois(boxes, pc_range, voxel_size, mutate)
    return single_roi_extractor([np.asarray(f) for f in point_feats[:len(strides)]], rois.numpy(),
                                strides)


# --------------------------------------------------------------------------------------
# a8: img_feats_sampling_bboxes_roi (srfdet_head.py:2424-2565 / :1963-2099)
# --------------------------------------------------------------------------------------
def img_rois(boxes, lidar2img, pc_range):
    """boxes (B,P,>=8) normalised centres (not mutated: the reference clones, :2435);
    lidar2img (B,Ncam,4,4).  -> rois (Ncam*B*P, 5), cam-major, batch idx = b + cam*B."""
    b = torch.as_tensor(boxes, dtype=torch.float32).clone()
    denorm_centres_(b, pc_range)
    cor = boxes3d_to_corners3d(b[..., :8])
    hom = torch.cat([cor, torch.ones_like(cor[..., :1])], -1)          # (B,P,8,4)
    l2i = torch.as_tensor(np.asarray(lidar2img), dtype=torch.float32)   # (B,Nc,4,4)
    cam = torch.matmul(l2i[:, :, None, None], hom[:, None, :, :, :, None]).squeeze(-1)
    uv = cam[..., 0:2] / torch.maximum(cam[..., 2:3], torch.full_like(cam[..., 2:3], 1e-5))
    rect = torch.cat([uv.min(3).values, uv.max(3).values], -1)         # (B,Nc,P,4)
    bs, nc = rect.shape[:2]
    rois = []
    for c in range(nc):
        r = bbox2roi([rect[i, c] for i in range(bs)])
        r[:, 0] += c * bs
        rois.append(r)
    return torch.cat(rois, 0)


def img_roi_feats(img_feats, boxes, lidar2img, pc_range, strides):
    """img_feats: list of (B,Ncam,C,H,W).  -> (B*P, C, 7, 7), summed over cameras."""
    rois = img_rois(boxes, lidar2img, pc_range)
    bs, nc, c = img_feats[0].shape[:3]
    p = np.asarray(boxes).shape[1]
    flat = [np.asarray(f).reshape(bs * nc, c, f.shape[3], f.shape[4]) for f in img_feats]
    s = single_roi_extractor(flat[:len(strides)], rois.numpy(), strides)
    s = torch.as_tensor(s).view(nc, bs, p, c, 7, 7)
    return s.permute(1, 2, 3, 4, 5, 0).sum(-1).reshape(bs * p, c, 7, 7).numpy()


# --------------------------------------------------------------------------------------
# a9 / a10: fusion projection (srfdet_head.py:2254-2264) and DynamicConv (:2633-2693)
# --------------------------------------------------------------------------------------
def fusion_proj(img_roi, pts_roi, weight, bias):
    x = torch.cat([torch.as_tensor(img_roi), torch.as_tensor(pts_roi)], 1).permute(0, 2, 3, 1)
    y = F.linear(x, torch.as_tensor(weight), torch.as_tensor(bias))
    return y.permute(0, 3, 1, 2).contiguous().numpy()


def dynamic_conv(params, prop_feats, roi_feats_kc77, dynamic_dim):
    """prop_feats (K,C); roi_feats (K,C,7,7) -> (K,C).  params: dynamic_layer.{weight,bias},
    norm1/2/3.{weight,bias}, out_layer.{weight,bias}."""
    p = {k: torch.as_tensor(v, dtype=torch.float32) for k, v in params.items()}
    roi = torch.as_tensor(roi_feats_kc77, dtype=torch.float32)
    k, c = roi.shape[:2]
    feats = roi.reshape(k, c, -1).permute(0, 2, 1)                  # (K,49,C)  (:2277-2278, :2667)
    prm = F.linear(torch.as_tensor(prop_feats, dtype=torch.float32), p['dynamic_layer.weight'],
                   p['dynamic_layer.bias'])
    npar = c * dynamic_dim
    p1 = prm[:, :npar].reshape(k, c, dynamic_dim)
    p2 = prm[:, npar:].reshape(k, dynamic_dim, c)
    f = torch.bmm(feats, p1)
    f = F.relu(F.layer_norm(f, (dynamic_dim,), p['norm1.weight'], p['norm1.bias']))
    f = torch.bmm(f, p2)
    f = F.relu(F.layer_norm(f, (c,), p['norm2.weight'], p['norm2.bias']))
    f = F.linear(f.flatten(1), p['out_layer.weight'], p['out_layer.bias'])
    f = F.relu(F.layer_norm(f, (c,), p['norm3.weight'], p['norm3.bias']))
    return f.numpy()


def region_features_lidar(params, point_feats, boxes, prop_feats, pc_range, voxel_size, strides,
                          dynamic_dim):
    """RoI sampling + interaction of SingleSRFDetHeadLiDAR.forward (srfdet_head.py:1455-1529)
    without the attention / FFN rows (SURVEY 8f 'next').  prop_feats None -> RoI mean (:1487-1490)."""
    roi = points_roi_feats(point_feats, boxes, pc_range, voxel_size, strides, mutate=True)
    if prop_feats is None:
        prop_feats = roi.reshape(roi.shape[0], roi.shape[1], -1).mean(-1)
    return dynamic_conv(params, prop_feats, roi, dynamic_dim), roi


# ----------------------------------------------------------------------------------------
# Pillar path
# ----------------------------------------------------------------------------------------
def pillar_feature_net(params, voxels, num_points, coors, voxel_size, pc_range, legacy=True, with_cluster_center=True,
                       with_voxel_center=True, with_distance=False, mode='max', eps=1e-3):
    """PillarFeatureNetCustom.forward (models/voxel_encoders/pillar_encoder_custom.py:95-161) with one
    PFNLayer (models/voxel_encoders/utils.py:109-147).  voxels (N,T,C) zero padded, coors (N,4) b,z,y,x."""
    f = torch.as_tensor(np.asarray(voxels, np.float32)).clone()
    npts = torch.as_tensor(np.asarray(num_points)).to(torch.float32)
    co = torch.as_tensor(np.asarray(coors)).to(torch.float32)
    vx, vy, vz = [float(v) for v in voxel_size]
    xo, yo, zo = vx / 2 + pc_range[0], vy / 2 + pc_range[1], vz / 2 + pc_range[2]
    ls = [f]
    if with_cluster_center:                                        # :114-119
        mean = f[:, :, :3].sum(dim=1, keepdim=True) / npts.view(-1, 1, 1)
        ls.append(f[:, :, :3] - mean)
    if with_voxel_center:                                          # :122-144
        ctr = torch.stack([co[:, 3] * vx + xo, co[:, 2] * vy + yo, co[:, 1] * vz + zo], -1).unsqueeze(1)
        if legacy:
            f[:, :, :3] = f[:, :, :3] - ctr                        # in place on the view: raw xyz channels change too
            ls.append(f[:, :, :3])
        else:
            ls.append(f[:, :, :3] - ctr)
    if with_distance:                                              # :146-148
        ls.append(torch.norm(f[:, :, :3], 2, 2, keepdim=True))
    x = torch.cat(ls, dim=-1)
    t = x.shape[1]
    mask = (npts.view(-1, 1).int() > torch.arange(t, dtype=torch.int32).view(1, -1)).unsqueeze(-1).to(x.dtype)   # utils.py:46-66
    x = x * mask
    w = torch.as_tensor(params['pfn_layers.0.linear.weight'])
    x = x @ w.t()                                                  # utils.py:125
    g, b = torch.as_tensor(params['pfn_layers.0.norm.weight']), torch.as_tensor(params['pfn_layers.0.norm.bias'])
    mu, var = torch.as_tensor(params['pfn_layers.0.norm.running_mean']), torch.as_tensor(params['pfn_layers.0.norm.running_var'])
    x = F.relu((x - mu) / torch.sqrt(var + eps) * g + b)           # BN1d eval over the channel dim, :126-128
    if mode == 'max':
        y = x.max(dim=1)[0]
    else:
        y = x.sum(dim=1) / npts.view(-1, 1)
    return y.numpy()


def pillars_scatter(feats, coors, batch_size, ny, nx):
    """[3P] mmdet3d PointPillarsScatter.forward_batch: canvas[b, :, y*nx + x] = feats."""
    feats = np.asarray(feats, np.float32)
    c = feats.shape[1]
    canvas = np.zeros((batch_size, c, ny * nx), np.float32)
    coors = np.asarray(coors)
    canvas[coors[:, 0], :, coors[:, 2] * nx + coors[:, 3]] = feats
    return canvas.reshape(batch_size, c, ny, nx)


# ----------------------------------------------------------------------------------------
# Dense BEV backbone + neck
# ----------------------------------------------------------------------------------------
def _t(v):
    return torch.as_tensor(np.asarray(v), dtype=torch.float32)


def _conv_bn_relu(x, p, conv, bn, stride, pad, eps, groups=1):
    x = F.conv2d(x, _t(p[conv + '.weight']), None, stride=stride, padding=pad, groups=groups)
    x = F.batch_norm(x, _t(p[bn + '.running_mean']), _t(p[bn + '.running_var']), _t(p[bn + '.weight']), _t(p[bn + '.bias']),
                     False, 0.0, eps)
    return F.relu(x)


def second_custom(params, x, layer_nums, layer_strides, eps=1e-3):
    """SECONDCustom.forward (models/backbones/second_custom.py:77-91): per block a strided 3x3 conv-BN-ReLU followed
    by layer_num 3x3 conv-BN-ReLU; parameters keyed blocks.{i}.{3j} (conv) / blocks.{i}.{3j+1} (BN)."""
    x = _t(x)
    outs = []
    for i, (ln, st) in enumerate(zip(layer_nums, layer_strides)):
        for j in range(ln + 1):
            x = _conv_bn_relu(x, params, f'blocks.{i}.{3 * j}', f'blocks.{i}.{3 * j + 1}', st if j == 0 else 1, 1, eps)
        outs.append(x)
    return [o.numpy() for o in outs]


def fpn(params, feats, num_outs, eps=1e-3, extra_convs=True):
    """[3P] mmdet 2.28.2 FPN.forward as configured by the reference (configs/nus/srfdet_voxel_nusc_L.py:66-75)."""
    feats = [_t(f) for f in feats]
    lat = [_conv_bn_relu(f, params, f'lateral_convs.{i}.conv', f'lateral_convs.{i}.bn', 1, 0, eps) for i, f in enumerate(feats)]
    for i in range(len(lat) - 1, 0, -1):
        lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], size=lat[i - 1].shape[2:], mode='nearest')
    outs = [_conv_bn_relu(lat[i], params, f'fpn_convs.{i}.conv', f'fpn_convs.{i}.bn', 1, 1, eps) for i in range(len(lat))]
    for i in range(len(lat), num_outs):
        if extra_convs:
            outs.append(_conv_bn_relu(outs[-1], params, f'fpn_convs.{i}.conv', f'fpn_convs.{i}.bn', 2, 1, eps))
        else:
            outs.append(F.max_pool2d(outs[-1], 1, stride=2))
    return [o.numpy() for o in outs]


# ----------------------------------------------------------------------------------------
# SRFDetHead: proposal generation, chained stages, decoding
# ----------------------------------------------------------------------------------------
def _sub(params, prefix):
    return {k[len(prefix):]: v for k, v in params.items() if k.startswith(prefix)}


def _dpg_logits(params, feats, conv_prefix, fc1, fc2, group=1, resize=None, eps=1e-3):
    """Staircase of depthwise stride-2 ConvModules + channel sum + two FCs (srfdet_head.py:521-543 / :560-595)."""
    feats = [_t(f) for f in feats]
    x = None
    for l in range(len(feats) - 1):
        inp = feats[l] if x is None else torch.cat([feats[l], x], dim=1)
        x = _conv_bn_relu(inp, params, f'{conv_prefix}.{l}.conv', f'{conv_prefix}.{l}.bn', 2, 1, eps, groups=inp.shape[1])
    last = torch.cat([feats[-1], x], dim=1)
    if resize is not None:
        last = F.interpolate(last, list(resize))
    if group > 1:
        last = last.view(last.shape[0] // group, group, *last.shape[1:]).sum(dim=1)
    s = last.sum(dim=1).flatten(1, 2)
    h = F.relu(F.linear(s, _t(params[fc1 + '.weight']), _t(params[fc1 + '.bias'])))
    return F.linear(h, _t(params[fc2 + '.weight']), _t(params[fc2 + '.bias']))


def dpg_init_proposals(params, img_feats, point_feats, n_exp, n_p, use_img, is_kitti=False):
    """SRFDetHead._get_init_proposals with DPG (srfdet_head.py:506-640) -> boxes (bs,n_p,dim), feats (bs,n_p,C)."""
    bs = point_feats[0].shape[0]
    w = _dpg_logits(params, point_feats, 'dpg_dw_convs_lidar', 'dpg_fc1_lidar', 'dpg_fc2_lidar').reshape(bs, n_exp, n_p)
    if use_img:
        flat = [np.asarray(f).reshape(-1, *f.shape[2:]) for f in img_feats]
        wi = _dpg_logits(params, flat, 'dpg_dw_convs_img', 'dpg_fc1_img', 'dpg_fc2_img', group=flat[0].shape[0] // bs,
                         resize=(30, 15) if is_kitti else (30, 30)).reshape(bs, n_exp, n_p)
        w = (w + wi) / 2
    w = w.softmax(1)
    eb = _t(params['init_proposal_boxes.weight']).view(n_exp, n_p, -1)
    ef = _t(params['init_proposal_feats.weight']).view(n_exp, n_p, -1)
    boxes = (w.unsqueeze(-1) * eb.unsqueeze(0)).sum(1)
    feats = (w.unsqueeze(-1) * ef.unsqueeze(0)).sum(1)
    return boxes.numpy(), feats.numpy()


def apply_deltas(deltas, boxes, weights, scale_clamp, pc_range):
    """SingleSRFDetHead.apply_deltas_lidar (srfdet_head.py:2331-2420)."""
    d, b = _t(deltas), _t(boxes)
    w = [float(v) for v in weights]
    size = torch.exp(b[:, 3:6])
    dxyz = torch.stack([d[:, j] / w[j] for j in range(3)], -1)
    dwlh = torch.stack([d[:, 3 + j] / w[3 + j] for j in range(3)], -1).clamp(max=scale_clamp)
    ctr = dxyz * size + b[:, 0:3]
    psize = torch.exp(dwlh) * size
    lo = torch.tensor(pc_range[:3], dtype=torch.float32)
    span = torch.tensor([pc_range[3] - pc_range[0], pc_range[4] - pc_range[1], pc_range[5] - pc_range[2]], dtype=torch.float32)
    ctr = ((ctr - lo) / span).clamp(min=0.0, max=1.0)
    return torch.cat([ctr, psize.log(), d[:, 6:len(w)]], dim=-1).numpy()


def single_head_stage(p, img_feats, point_feats, boxes, prop, lidar2img, cfg):
    """SingleSRFDetHead(.LiDAR).forward (srfdet_head.py:2221-2326 / 1455-1529), batch 1.
    boxes (1,n_p,dim) normalised centres -- MUTATED in place like the reference; prop (n_p, C)."""
    pc, vs = cfg['pc_range'], cfg['voxel_size']
    n_p, c = boxes.shape[1], cfg['C']
    if img_feats is not None:
        img = img_roi_feats(img_feats, boxes.copy(), lidar2img, pc, list(cfg['istrides']))
        pts = points_roi_feats(point_feats, boxes, pc, vs, list(cfg['strides']))
        roi = fusion_proj(img, pts, p['output_fused_proj.weight'], p['output_fused_proj.bias'])
    else:
        roi = points_roi_feats(point_feats, boxes, pc, vs, list(cfg['strides']))
    x = _t(prop)
    heads = cfg['attn_heads']
    qkv = F.linear(x, _t(p['self_attn_lidar.in_proj_weight']), _t(p['self_attn_lidar.in_proj_bias']))
    q, k, v = [t.view(n_p, heads, c // heads).transpose(0, 1) for t in qkv.chunk(3, dim=-1)]
    att = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(c // heads), dim=-1) @ v
    att = att.transpose(0, 1).reshape(n_p, c)
    x = F.layer_norm(x + F.linear(att, _t(p['self_attn_lidar.out_proj.weight']), _t(p['self_attn_lidar.out_proj.bias'])), (c,),
                     _t(p['norm1_lidar.weight']), _t(p['norm1_lidar.bias']))
    x2 = _t(dynamic_conv(_sub(p, 'inst_interact_lidar.'), x.numpy(), roi, cfg['d']))
    obj = F.layer_norm(x + x2, (c,), _t(p['norm2_lidar.weight']), _t(p['norm2_lidar.bias']))
    ff = F.linear(F.relu(F.linear(obj, _t(p['linear1_lidar.weight']), _t(p['linear1_lidar.bias']))), _t(p['linear2_lidar.weight']),
                  _t(p['linear2_lidar.bias']))
    obj = F.layer_norm(obj + ff, (c,), _t(p['norm3_lidar.weight']), _t(p['norm3_lidar.bias']))

    def tower(prefix, n):
        f = obj
        for t in range(n):
            f = F.relu(F.layer_norm(F.linear(f, _t(p[f'{prefix}.{3 * t}.weight'])), (c,), _t(p[f'{prefix}.{3 * t + 1}.weight']),
                                    _t(p[f'{prefix}.{3 * t + 1}.bias'])))
        return f
    logits = F.linear(tower('cls_module_lidar', cfg['n_cls']), _t(p['class_logits_lidar.weight']), _t(p['class_logits_lidar.bias']))
    deltas = F.linear(tower('reg_module_lidar', cfg['n_reg']), _t(p['bboxes_delta_lidar.weight']), _t(p['bboxes_delta_lidar.bias']))
    pred = apply_deltas(deltas.numpy(), boxes.reshape(n_p, -1), cfg['bbox_weights'], cfg['scale_clamp'], pc)
    return logits.numpy(), pred, obj.numpy()


def srfdet_head_forward(params, img_feats, point_feats, lidar2img, cfg, trace=None):
    """SRFDetHead.forward (srfdet_head.py:371-498), batch 1: DPG -> sigmoid -> chained stages -> centre
    de-normalisation.  -> logits (S,1,n_p,cls), boxes (S,1,n_p,dim)."""
    use_img = img_feats is not None
    boxes, prop = dpg_init_proposals(params, img_feats, point_feats, cfg['n_exp'], cfg['n_p'], use_img)
    boxes = boxes.copy()
    boxes[..., :3] = 1.0 / (1.0 + np.exp(-boxes[..., :3].astype(np.float64))).astype(np.float32)
    prop = prop[0]
    lg, bx = [], []
    if trace is not None:
        trace.update(init_boxes=boxes.copy(), init_prop=prop.copy(), stage_in=[], stage_out=[])
    for s in range(cfg['stages']):
        if trace is not None:
            trace['stage_in'].append((boxes.copy(), prop.copy()))
        logits, pred, prop = single_head_stage(_sub(params, f'head_series_lidar.{s}.'), img_feats, point_feats, boxes, prop, lidar2img, cfg)
        lg.append(logits[None])
        bx.append(pred[None].copy())
        boxes = pred[None].copy()
        if trace is not None:
            trace['stage_out'].append((logits.copy(), pred.copy(), prop.copy()))
    lg, bx = np.stack(lg), np.stack(bx)
    pc = cfg['pc_range']
    bx[..., :3] = bx[..., :3] * np.array([pc[3] - pc[0], pc[4] - pc[1], pc[5] - pc[2]], np.float32) + np.array(pc[:3], np.float32)
    return lg, bx


def decode_boxes(logits, boxes):
    """get_bboxes decode (srfdet_head.py:1245-1268, core/bbox/util.py:41-81): sigmoid, exp sizes, atan2, bottom centre."""
    lg, b = _t(logits), _t(boxes)
    rot = torch.atan2(b[..., 6:7], b[..., 7:8])
    out = torch.cat([b[..., 0:3], b[..., 3:6].exp(), rot, b[..., 8:]], dim=-1)
    out[..., 2] = out[..., 2] - out[..., 5] * 0.5
    return torch.sigmoid(lg).numpy(), out.numpy()
