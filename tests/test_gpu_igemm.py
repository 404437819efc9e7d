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
