"""tcgen05 implicit-GEMM kernel in its dense form vs torch matmul (isolates descriptor / pipeline
bugs), in every operand encoding: bf16, f16 (one MMA per product) and the hi + lo split forms
(three MMAs per product, fp32-equivalent)."""
import numpy as np
import pytest
import torch

from util import rel_err

pytestmark = pytest.mark.gpu


def _encs():
    from srfdet_b200 import _lib as L
    return {'bf16': L.BF16, 'f16': L.F16, 'bf16x2': L.BF16X2, 'f16x2': L.F16X2}


# tolerance vs an fp64 matmul of the ROUNDED operands (plain forms: only summation order differs) or of
# the fp32 operands (split forms: what is left is the 3-term product error, 2^-16 / 2^-21 relative)
TOL = {'bf16': 2e-3, 'f16': 2e-3, 'bf16x2': 3e-5, 'f16x2': 2e-5}      # (f16x2: fp32 accumulation over K = 6272 dominates)


def _round(t, name):
    if name.startswith('bf16'):
        return t.to(torch.bfloat16).float()
    return t.to(torch.float16).float()


def _run(m, k, n, enc_name='bf16', relu=False, ln=False, out='f32', seed=0, eps=1e-5):
    from srfdet_b200 import _lib as L
    from srfdet_b200.plugin.head import encode_rows
    lib = L.load()
    enc = _encs()[enc_name]
    g = torch.Generator().manual_seed(seed)
    a = (torch.randn(m, k, generator=g) * 0.5).cuda()
    w = (torch.randn(n, k, generator=g) / k ** 0.5).cuda()
    bias = torch.randn(n, generator=g).cuda()
    lnw = (torch.rand(n, generator=g) + 0.5).cuda()
    lnb = torch.randn(n, generator=g).cuda()
    ae = encode_rows(a, enc)
    wp = torch.empty(L.enc_width(enc, n * k), dtype=L.enc_torch_dtype(enc), device='cuda')
    st = L.stream_ptr()
    L.check(lib.srf_pack_linear_tc(L.ptr(w), n, k, enc, L.ptr(wp), st), 'pack')
    out_enc = L.F32 if out == 'f32' else (enc if out == 'same' else {True: L.F16, False: L.BF16}[enc in (L.F16, L.F16X2)])
    o = torch.empty((m, L.enc_width(out_enc, n)), dtype=L.enc_torch_dtype(out_enc), device='cuda')
    epi = (1 if relu else 0) | (2 if ln else 0)
    L.check(lib.srf_linear_tc(L.ptr(ae), enc, m, k, L.ptr(wp), n, L.ptr(bias), None, epi, L.ptr(lnw) if ln else None,
                              L.ptr(lnb) if ln else None, eps, L.ptr(o), out_enc, 1, st), 'linear')
    torch.cuda.synchronize()
    if L.enc_is_split(enc):
        ref = a.double() @ w.double().t() + bias.double()
    else:
        ref = _round(a, enc_name).double() @ _round(w, enc_name).double().t() + bias.double()
    if ln:
        ref = torch.nn.functional.layer_norm(ref, (n,), lnw.double(), lnb.double(), eps=eps)
    if relu:
        ref = torch.relu(ref)
    return L.decode(o, n).cpu().numpy(), ref.cpu().numpy()


@pytest.mark.parametrize('enc', ['bf16', 'f16', 'bf16x2', 'f16x2'])
@pytest.mark.parametrize('m,k,n', [(128, 16, 16), (130, 32, 32), (1, 64, 64), (900, 128, 128), (257, 16, 128),
                                   (900, 128, 8192), (900, 6272, 128), (4410, 256, 128), (300, 64, 32)])
def test_linear_matches_matmul(m, k, n, enc):
    got, ref = _run(m, k, n, enc)
    assert rel_err(got, ref) < TOL[enc]


@pytest.mark.parametrize('enc', ['bf16', 'f16', 'bf16x2', 'f16x2'])
def test_linear_epilogues(enc):
    got, ref = _run(900, 6272, 128, enc, relu=True, ln=True)
    assert rel_err(got, ref) < TOL[enc] * 2
    got, ref = _run(333, 128, 128, enc, relu=False, ln=True, eps=1e-3)      # explicit LayerNorm eps
    assert rel_err(got, ref) < TOL[enc] * 2
    got, ref = _run(333, 128, 64, enc, relu=True, out='same')                # output in the operand encoding
    assert rel_err(got, ref) < (1e-2 if enc == 'bf16' else 2e-3 if enc == 'f16' else 3e-5)
    if enc.endswith('x2'):
        got, ref = _run(333, 128, 64, enc, out='plain')                      # split operands, plain 16-bit output
        assert rel_err(got, ref) < (1e-2 if enc == 'bf16x2' else 2e-3)


def test_legacy_bf16_entry_points():
    """srf_linear_bf16 / srf_pack_linear_bf16 (round-1 names) are the SRF_BF16 forms."""
    from srfdet_b200 import _lib as L
    lib = L.load()
    g = torch.Generator().manual_seed(9)
    a = torch.randn(300, 256, generator=g).cuda().to(torch.bfloat16).contiguous()
    w = (torch.randn(128, 256, generator=g) / 16).cuda()
    wp = torch.empty(128 * 256, dtype=torch.bfloat16, device='cuda')
    st = L.stream_ptr()
    L.check(lib.srf_pack_linear_bf16(L.ptr(w), 128, 256, L.ptr(wp), st), 'pack')
    out = torch.empty((300, 128), dtype=torch.float32, device='cuda')
    L.check(lib.srf_linear_bf16(L.ptr(a), 300, 256, L.ptr(wp), 128, None, 0, None, None, L.ptr(out), L.F32, 1, st), 'linear')
    ref = a.float() @ w.to(torch.bfloat16).float().t()
    assert rel_err(out.cpu().numpy(), ref.cpu().numpy()) < 2e-3


@pytest.mark.parametrize('enc', ['bf16', 'f16x2'])
@pytest.mark.parametrize('m,k,n,splits', [(900, 6272, 128, 18), (300, 12544, 256, 9), (100, 1024, 64, 8)])
def test_linear_split_k(m, k, n, splits, enc):
    from srfdet_b200 import _lib as L
    from srfdet_b200.plugin.head import encode_rows
    lib = L.load()
    e = _encs()[enc]
    g = torch.Generator().manual_seed(3)
    a = (torch.randn(m, k, generator=g) * 0.5).cuda()
    w = (torch.randn(n, k, generator=g) / k ** 0.5).cuda()
    ae = encode_rows(a, e)
    wp = torch.empty(L.enc_width(e, n * k), dtype=L.enc_torch_dtype(e), device='cuda')
    st = L.stream_ptr()
    L.check(lib.srf_pack_linear_tc(L.ptr(w), n, k, e, L.ptr(wp), st), 'pack')
    eff = lib.srf_linear_splits_enc(k, e, splits)
    part = torch.empty((eff, m, n), dtype=torch.float32, device='cuda')
    L.check(lib.srf_linear_tc(L.ptr(ae), e, m, k, L.ptr(wp), n, None, None, 0, None, None, 1e-5, L.ptr(part), L.F32, splits, st), 'linear')
    if L.enc_is_split(e):
        ref = (a.double() @ w.double().t()).float()
    else:
        ref = _round(a, enc) @ _round(w, enc).t()
    assert rel_err(part.sum(0).cpu().numpy(), ref.cpu().numpy()) < TOL[enc] * 2
    # slabs are summed in order with bias + LayerNorm + ReLU by srf_layernorm
    bias = torch.randn(n, generator=g).cuda(); lw = (torch.rand(n, generator=g) + 0.5).cuda(); lb = torch.randn(n, generator=g).cuda()
    out = torch.empty((m, n), dtype=torch.float32, device='cuda')
    L.check(lib.srf_layernorm(L.ptr(part), L.F32, m, n, eff, L.ptr(bias), L.ptr(lw), L.ptr(lb), 1e-5, 1, L.ptr(out), st), 'ln')
    ref2 = torch.relu(torch.nn.functional.layer_norm(ref + bias, (n,), lw, lb))
    assert rel_err(out.cpu().numpy(), ref2.cpu().numpy()) < max(TOL[enc] * 2, 1e-5)


# ------------------------------------------------------------------------------ sparse conv on arbitrary neighbour tables
def _conv_case(c, enc_name, rows, n_in, nbr, residual, relu, n_out=None):
    """srf_spconv_tc (cin = cout = c, kvol 27, 16-bit in / out) on an arbitrary neighbour table vs fp64 torch."""
    from srfdet_b200 import _lib as L
    from srfdet_b200.plugin.head import encode_rows
    lib = L.load()
    enc = _encs()[enc_name]
    g = torch.Generator().manual_seed(11)
    x = (torch.randn(n_in, c, generator=g) * 0.7).cuda()
    w = (torch.randn(27, c, c, generator=g) / (9 * c) ** 0.5).cuda()
    bias = torch.randn(c, generator=g).cuda()
    res = (torch.randn(rows, c, generator=g)).cuda() if residual else None
    xe = encode_rows(x, enc)
    wp = torch.empty(27 * c * c, dtype=L.enc_torch_dtype(enc), device='cuda')
    st = L.stream_ptr()
    L.check(lib.srf_pack_weight_tc(L.ptr(w), 27, c, c, enc, L.ptr(wp), st), 'pack')
    cnt = torch.tensor([rows if n_out is None else n_out], dtype=torch.int32, device='cuda')
    out = torch.full((rows, c), 7.0, dtype=L.enc_torch_dtype(enc), device='cuda')
    re = encode_rows(res, enc) if residual else None
    a = L.ConvArgs()
    a.in_, a.in_dtype, a.in_rows = L.ptr(xe), enc, n_in
    a.cin, a.cout, a.kvol = c, c, 27
    a.nbr, a.tile_mask, a.cap_out, a.d_n_out = L.ptr(nbr), None, rows, L.ptr(cnt)
    a.w, a.bias, a.residual, a.relu = L.ptr(wp), L.ptr(bias), (L.ptr(re) if residual else None), int(relu)
    a.out, a.out_dtype = L.ptr(out), enc
    import ctypes
    L.check(lib.srf_spconv_tc(ctypes.byref(a), st), 'conv')
    torch.cuda.synchronize()
    xr, wr = _round(x, enc_name).double(), _round(w, enc_name).double()
    ref = bias.double().repeat(rows, 1)
    for k in range(27):
        idx = nbr[k].long()
        ok = idx >= 0
        ref[ok] += xr[idx[ok]] @ wr[k]
    if residual:
        ref += _round(res, enc_name).double()
    if relu:
        ref = torch.relu(ref)
    live = rows if n_out is None else n_out
    return out.float()[:live].cpu().numpy(), ref[:live].cpu().numpy(), out.float()[live:].cpu().numpy()


@pytest.mark.parametrize('enc', ['f16', 'bf16'])
@pytest.mark.parametrize('c', [16, 32])
@pytest.mark.parametrize('pattern', ['random', 'banded', 'two_clusters', 'empty'])
def test_sparse_conv_arbitrary_rulebooks(c, enc, pattern):
    """The 16-bit sparse-conv kernels (tcgen05 gather-GEMM, warp-MMA 16 -> 16) on neighbour tables that are NOT rulebooks of
    a real cloud: random tables, a banded SubM-like table, neighbours in two far-apart clusters, all-missing offset groups
    and a whole table of -1, with a ragged device-side row count; rows beyond the count must stay untouched."""
    rows, n_in = 1024, 6000
    g = torch.Generator().manual_seed(5)
    if pattern == 'random':
        nbr = torch.randint(0, n_in, (27, rows), generator=g, dtype=torch.int32)
        nbr[torch.rand(27, rows, generator=g) < 0.5] = -1
    elif pattern == 'banded':
        base = torch.arange(rows, dtype=torch.int32)[None, :] * 5
        nbr = base + torch.randint(-150, 150, (27, rows), generator=g, dtype=torch.int32)
        nbr = nbr.clamp_(0, n_in - 1)
        nbr[torch.rand(27, rows, generator=g) < 0.4] = -1
    elif pattern == 'two_clusters':
        lowc = torch.randint(0, 200, (27, rows), generator=g, dtype=torch.int32)
        highc = torch.randint(n_in - 200, n_in, (27, rows), generator=g, dtype=torch.int32)
        nbr = torch.where(torch.rand(27, rows, generator=g) < 0.5, lowc, highc)
        nbr[torch.rand(27, rows, generator=g) < 0.3] = -1
        nbr[9:18, 128:256] = -1                                   # a whole (tile, dz) group without neighbours
    else:
        nbr = torch.full((27, rows), -1, dtype=torch.int32)
    nbr = nbr.cuda().contiguous()
    got, ref, tail = _conv_case(c, enc, rows, n_in, nbr, residual=(pattern != 'random'), relu=(pattern != 'banded'), n_out=1000 - 37)
    tol = 3e-3 if enc == 'f16' else 1.5e-2        # 16-bit OUTPUT rounding of values up to ~4
    assert rel_err(got, ref) < tol
    assert (tail == 7.0).all()                    # rows beyond the device-side count are not written
