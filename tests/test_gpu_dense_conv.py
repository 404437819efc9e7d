"""Halo-tile dense 3x3 convolution (csrc/conv3x3_halo.cu, `srf_conv3x3_rows`) vs torch conv2d on the rounded
operands: shapes that are not multiples of the 16 x 8 tile, several images, both K halves (cin 256), both column
tiles (cout 256), fp32 / 16-bit outputs, and equality with the gather-GEMM path it replaces."""
import pytest
import torch

from util import rel_err

pytestmark = pytest.mark.gpu


def _run(n, h, w, cin, cout, enc_name, out_f32, relu, seed=0):
    from srfdet_b200 import _lib as L
    from srfdet_b200.plugin.head import encode_rows
    lib = L.load()
    enc = {'f16': L.F16, 'bf16': L.BF16}[enc_name]
    tdt = torch.float16 if enc_name == 'f16' else torch.bfloat16
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(n, cin, h, w, generator=g) * 0.5).cuda()
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / (9 * cin) ** 0.5).cuda()
    bias = torch.randn(cout, generator=g).cuda()
    rows = x.permute(0, 2, 3, 1).reshape(n * h * w, cin).contiguous()
    xe = encode_rows(rows, enc)
    kio = wt.permute(2, 3, 1, 0).reshape(9, cin, cout).contiguous()
    wp = torch.empty(kio.numel(), dtype=tdt, device='cuda')
    st = L.stream_ptr()
    L.check(lib.srf_pack_weight_tc(L.ptr(kio), 9, cin, cout, enc, L.ptr(wp), st), 'pack')
    out_enc = L.F32 if out_f32 else enc
    cap = (n * h * w + 127) // 128 * 128
    y = torch.full((cap, cout), 3.0, dtype=L.enc_torch_dtype(out_enc), device='cuda')
    L.check(lib.srf_conv3x3_rows(L.ptr(xe), enc, n, h, w, cin, L.ptr(wp), cout, L.ptr(bias), int(relu), L.ptr(y), out_enc, st), 'conv3x3')
    torch.cuda.synchronize()
    ref = torch.nn.functional.conv2d(x.to(tdt).double(), wt.to(tdt).double(), bias.double(), padding=1)
    if relu:
        ref = torch.relu(ref)
    ref = ref.permute(0, 2, 3, 1).reshape(n * h * w, cout)
    return y.float()[:n * h * w].cpu().numpy(), ref.cpu().numpy(), y.float()[n * h * w:].cpu().numpy()


@pytest.mark.parametrize('enc', ['f16', 'bf16'])
@pytest.mark.parametrize('n,h,w,cin,cout,out_f32,relu', [
    (1, 16, 8, 128, 128, True, False),        # exactly one tile
    (1, 23, 23, 256, 256, False, True),       # ragged both ways, two K halves, two column tiles
    (2, 46, 46, 128, 128, True, True),        # two images
    (1, 5, 3, 128, 256, True, True),          # smaller than a tile
    (1, 184, 184, 128, 128, False, True),     # production size (block 0 of SECONDCustom)
    (1, 92, 92, 256, 128, True, False),
])
def test_conv3x3_halo_vs_torch(n, h, w, cin, cout, out_f32, relu, enc):
    got, ref, tail = _run(n, h, w, cin, cout, enc, out_f32, relu)
    tol = 2e-3 if out_f32 else (3e-3 if enc == 'f16' else 1.5e-2)       # (16-bit output rounding)
    assert rel_err(got, ref) < tol
    assert (tail == 3.0).all()                # padding rows of the output buffer are not written


def test_halo_conv_tma_and_cp_async_producers_agree():
    """The halo is staged by TMA tensor copies (cp.async.bulk.tensor.5d, zero padding = the copy engine's out-of-bounds fill) on
    a B200 box; the cp.async gather producers remain as the fall-back.  Both feed the same MMAs: identical results."""
    import os
    import subprocess
    import sys
    from srfdet_b200 import _lib as L
    got, ref, _ = _run(2, 37, 29, 256, 128, 'f16', True, True, seed=3)
    assert L.load().srf_conv3x3_last_used_tma() == 1, 'tensor-map encoder unavailable: the TMA producer did not run'
    code = ('import sys, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r)\n'
            'import test_gpu_dense_conv as t; from srfdet_b200 import _lib as L\n'
            'got, ref, _ = t._run(2, 37, 29, 256, 128, "f16", True, True, seed=3)\n'
            'assert L.load().srf_conv3x3_last_used_tma() == 0\n'
            'np.save(sys.argv[1], got)\n') % (os.path.dirname(os.path.abspath(__file__)), os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import tempfile
    import numpy as np
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, 'cp.npy')
        subprocess.run([sys.executable, '-c', code, out], check=True, env=dict(os.environ, SRF_HALO_TMA='0'), timeout=300)
        other = np.load(out)
    assert np.array_equal(got, other)
    assert rel_err(got, ref) < 2e-3


def test_halo_conv_equals_gather_path():
    """SECONDCustom through the halo kernel == through the gather-GEMM kernel over the dense rulebook (fp16 mode):
    same operands, same fp32 accumulation, only the summation order inside the tensor core differs."""
    from srfdet_b200.plugin import bev_backbone as bb
    torch.manual_seed(4)
    net = bb.SECONDCustom(in_channels=256, out_channels=[128, 128, 256], layer_nums=[1, 1, 1], layer_strides=[1, 2, 2]).cuda().eval()
    x = torch.randn(1, 256, 40, 56, device='cuda')
    old = bb.HALO_CONV
    try:
        bb.HALO_CONV = True
        with torch.no_grad():
            a = [t.clone() for t in net(x, precision='fp16')]
        net._cache.clear()
        bb.HALO_CONV = False
        with torch.no_grad():
            b = net(x, precision='fp16')
    finally:
        bb.HALO_CONV = old
    for u, v in zip(a, b):
        assert rel_err(u.cpu().numpy(), v.cpu().numpy()) < 2e-3


@pytest.mark.parametrize('precision,tol', [('fp16', 2e-3), ('fp32', 2e-4)])
def test_img_convs_vs_torch(precision, tol):
    """SRFDetHead.img_convs (srfdet_head.py:146-157, :404-416: Conv2d(256 -> 128, 3x3, pad 1, bias) per image FPN level over all
    cameras) on the library's dense conv kernels vs torch conv2d in float64; ragged map sizes; the per-frame memo is
    shared by dpg_image_logits() / forward() and follows in-place weight updates."""
    from srfdet_b200.pipeline import head_cfg
    from srfdet_b200.plugin import registry
    torch.manual_seed(7)
    head = registry.build_head(head_cfg('nusc', True)).cuda().eval()
    assert head.feat_channels_img == 256 and head.hidden_dim == 128 and len(head.img_convs) == 4
    sizes = [(29, 50), (15, 25), (8, 13), (4, 7)]
    feats = [torch.randn(1, 3, 256, h, w, device='cuda') for h, w in sizes]
    with torch.no_grad():
        maps = head._image_maps(feats, precision)
        again = head._image_maps(feats, precision)
    assert all(a is b for a, b in zip(maps, again))                    # one evaluation per set of input maps
    for i, (m, f) in enumerate(zip(maps, feats)):
        assert m.shape == (1, 3, 128, *sizes[i])
        assert m[0].is_contiguous(memory_format=torch.channels_last)   # what the samplers / DPG kernels read in place
        cv = head.img_convs[i]
        ref = torch.nn.functional.conv2d(f[0].double(), cv.weight.detach().double(), cv.bias.detach().double(), padding=1)
        assert rel_err(m[0].cpu().numpy(), ref.cpu().numpy()) < tol
    with torch.no_grad():
        head.img_convs[0].bias.add_(1.0)
        new = head._image_maps(feats, precision)
    assert rel_err((new[0] - maps[0]).cpu().numpy(), torch.ones_like(maps[0]).cpu().numpy()) < 1e-3
