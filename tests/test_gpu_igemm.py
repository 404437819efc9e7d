"""tcgen05 implicit-GEMM kernel in its dense form vs torch matmul (isolates descriptor / pipeline bugs)."""
import numpy as np
import pytest
import torch

from util import rel_err

pytestmark = pytest.mark.gpu


def _run(m, k, n, relu=False, ln=False, out_bf16=False, seed=0):
    from srfdet_b200 import _lib as L
    lib = L.load()
    g = torch.Generator().manual_seed(seed)
    a = (torch.randn(m, k, generator=g) * 0.5).cuda()
    w = (torch.randn(n, k, generator=g) / k ** 0.5).cuda()
    bias = torch.randn(n, generator=g).cuda()
    lnw = (torch.rand(n, generator=g) + 0.5).cuda()
    lnb = torch.randn(n, generator=g).cuda()
    ab = a.to(torch.bfloat16).contiguous()
    wp = torch.empty(n * k, dtype=torch.bfloat16, device='cuda')
    st = L.stream_ptr()
    L.check(lib.srf_pack_linear_bf16(L.ptr(w), n, k, L.ptr(wp), st), 'pack')
    out = torch.empty((m, n), dtype=torch.bfloat16 if out_bf16 else torch.float32, device='cuda')
    epi = (1 if relu else 0) | (2 if ln else 0)
    L.check(lib.srf_linear_bf16(L.ptr(ab), m, k, L.ptr(wp), n, L.ptr(bias), epi, L.ptr(lnw) if ln else None,
                                L.ptr(lnb) if ln else None, L.ptr(out), L.BF16 if out_bf16 else L.F32, 1, st), 'linear')
    torch.cuda.synchronize()
    ref = ab.float() @ w.to(torch.bfloat16).float().t() + bias
    if ln:
        ref = torch.nn.functional.layer_norm(ref, (n,), lnw, lnb)
    if relu:
        ref = torch.relu(ref)
    return out.float().cpu().numpy(), ref.cpu().numpy()


@pytest.mark.parametrize('m,k,n', [(128, 16, 16), (130, 32, 32), (1, 64, 64), (900, 128, 128), (257, 16, 128),
                                   (900, 128, 8192), (900, 6272, 128), (4410, 256, 128), (300, 64, 32)])
def test_linear_bf16_matches_matmul(m, k, n):
    got, ref = _run(m, k, n)
    assert rel_err(got, ref) < 2e-3      # same bf16-rounded operands, fp32 accumulate: only summation order differs


def test_linear_bf16_epilogues():
    got, ref = _run(900, 6272, 128, relu=True, ln=True)
    assert rel_err(got, ref) < 2e-3
    got, ref = _run(333, 128, 64, relu=True, out_bf16=True)
    assert rel_err(got, ref) < 1e-2
    got, ref = _run(333, 128, 128, relu=False, ln=True, out_bf16=True)
    assert rel_err(got, ref) < 1e-2


@pytest.mark.parametrize('m,k,n,splits', [(900, 6272, 128, 18), (300, 12544, 256, 9), (100, 1024, 64, 8)])
def test_linear_bf16_split_k(m, k, n, splits):
    from srfdet_b200 import _lib as L
    lib = L.load()
    g = torch.Generator().manual_seed(3)
    a = (torch.randn(m, k, generator=g) * 0.5).cuda().to(torch.bfloat16).contiguous()
    w = (torch.randn(n, k, generator=g) / k ** 0.5).cuda()
    wp = torch.empty(n * k, dtype=torch.bfloat16, device='cuda')
    st = L.stream_ptr()
    L.check(lib.srf_pack_linear_bf16(L.ptr(w), n, k, L.ptr(wp), st), 'pack')
    eff = lib.srf_linear_splits(k, splits)
    part = torch.empty((eff, m, n), dtype=torch.float32, device='cuda')
    L.check(lib.srf_linear_bf16(L.ptr(a), m, k, L.ptr(wp), n, None, 0, None, None, L.ptr(part), L.F32, splits, st), 'linear')
    ref = a.float() @ w.to(torch.bfloat16).float().t()
    assert rel_err(part.sum(0).cpu().numpy(), ref.cpu().numpy()) < 2e-3
    # slabs are summed in order with bias + LayerNorm + ReLU by srf_layernorm
    bias = torch.randn(n, generator=g).cuda(); lw = (torch.rand(n, generator=g) + 0.5).cuda(); lb = torch.randn(n, generator=g).cuda()
    out = torch.empty((m, n), dtype=torch.float32, device='cuda')
    L.check(lib.srf_layernorm(L.ptr(part), L.F32, m, n, eff, L.ptr(bias), L.ptr(lw), L.ptr(lb), 1e-5, 1, L.ptr(out), st), 'ln')
    ref2 = torch.relu(torch.nn.functional.layer_norm(ref + bias, (n,), lw, lb))
    assert rel_err(out.cpu().numpy(), ref2.cpu().numpy()) < 2e-3
