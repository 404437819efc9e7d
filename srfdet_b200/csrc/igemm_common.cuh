// Shared pieces of the tcgen05 implicit-GEMM kernels: PTX wrappers, argument block, fused epilogue.
#pragma once
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace srf {

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// kind::f16 covers bf16 and f16 operands (selected by the instruction descriptor)
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread i <- lane (base_lane + i)
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=0
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}

// one elected lane of a converged warp (ptxas recognises elect.sync and emits the single-thread
// region without per-instruction broadcast loops around UTCHMMA / UTCBAR)
__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred = 0, laneid = 0;
  asm volatile(
      "{\n\t.reg .b32 %%rx;\n\t.reg .pred %%px;\n\t"
      "elect.sync %%rx|%%px, %2;\n\t"
      "@%%px mov.s32 %1, 1;\n\t"
      "mov.s32 %0, %%rx;\n\t}"
      : "+r"(laneid), "+r"(pred)
      : "r"(0xFFFFFFFFu));
  return pred;
}

// descriptor from its two 32-bit halves; the high half of every no-swizzle K-major descriptor used
// here is constant: SBO = 128 B (>>4 at bit 32) | version 1 (bit 46)
constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);
__device__ __forceinline__ uint64_t desc_pack(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}

__device__ __forceinline__ int k_first(int tile, int kvol) { return (int)(((unsigned)tile * 11u) % (unsigned)kvol); }

struct IgemmArgs {
  const uint16_t* in;   // 16-bit elements (bf16 or f16, see fmt); split rows are [hi | lo]
  long long in_stride;  // elements between A rows
  long long k_stride;   // element offset of the k-th slice inside a row (0 for sparse conv)
  long long in_lo_off;  // split operands: elements from a row's hi part to its lo part
  const int32_t* nbr;   // (kvol, cap_out) or null (dense: identity rows)
  const uint32_t* tile_mask;
  const int32_t* d_n_out;
  int cap_out;  // rows bound (multiple of 128 for the sparse path)
  int m_rows;   // dense: number of rows
  int kvol;
  int n_tiles;
  const uint16_t* w;  // packed [n_tile][k][q][hi|lo][KC/8][COUT][8]
  const float* bias;
  const void* residual;   // same encoding / stride as out
  int relu, ln;
  const float *ln_w, *ln_b;
  float ln_eps;
  void* out;
  int fmt;       // 16-bit operand format: 0 = bf16, 1 = f16
  int out_enc;   // SRF_F32 | SRF_BF16 | SRF_F16 | SRF_BF16X2 | SRF_F16X2 (16-bit forms use fmt)
  long long out_stride;   // elements of the out array between rows
  long long out_lo_off;   // split output: elements from hi to lo part
  void* out2;             // dense linear: optional second copy of the result (e.g. fp32 trunk + 16-bit GEMM operand)
  int out2_enc;
  int ln_per_tile;        // LayerNorm parameters indexed by global column (one norm per column tile: merged towers)
  float* dense;
  const int4* out_coors;
  int D, H, W;
  int k_splits;  // dense linear only: K slices are dealt to k_splits CTAs per output tile, fp32 partials in k_splits slabs
  int dbg;       // -DSRF_IGEMM_PROF builds only: ablation switches / launch id
};

// residual add of NC consecutive channels starting at column c (same encoding as the output)
template <int NC>
__device__ __forceinline__ void add_residual(const IgemmArgs& a, size_t row, int c, float* v) {
  if (a.out_enc == SRF_F32) {
    const float4* rp = (const float4*)((const float*)a.residual + row * a.out_stride + c);
#pragma unroll
    for (int i = 0; i < NC; i += 4) {
      const float4 u = __ldg(rp + i / 4);
      v[i] += u.x; v[i + 1] += u.y; v[i + 2] += u.z; v[i + 3] += u.w;
    }
    return;
  }
  const bool f16 = a.fmt != 0;
  const uint16_t* base = (const uint16_t*)a.residual + row * a.out_stride + c;
#pragma unroll
  for (int part = 0; part < 2; ++part) {
    if (part == 1 && !enc_is_split(a.out_enc)) break;
    const uint4* rp = (const uint4*)(base + (part ? a.out_lo_off : 0));
#pragma unroll
    for (int i = 0; i < NC; i += 8) {
      const uint4 u = __ldg(rp + i / 8);
      const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = unpack16x2(f16, w4[q]);
        v[i + 2 * q] += f.x;
        v[i + 2 * q + 1] += f.y;
      }
    }
  }
}

// store of NC consecutive channels of one row at column c in encoding `enc` (row stride / lo offset in elements)
template <int NC>
__device__ __forceinline__ void store_cols_to(void* out, int enc, long long stride, long long lo_off, bool f16, size_t row, int c,
                                              const float* v) {
  if (enc == SRF_F32) {
    float4* op = (float4*)((float*)out + row * stride + c);
#pragma unroll
    for (int i = 0; i < NC; i += 4) op[i / 4] = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    return;
  }
  uint16_t* base = (uint16_t*)out + row * stride + c;
  if (!enc_is_split(enc)) {
    uint4* op = (uint4*)base;
#pragma unroll
    for (int i = 0; i < NC; i += 8)
      op[i / 8] = make_uint4(pack16x2(f16, v[i], v[i + 1]), pack16x2(f16, v[i + 2], v[i + 3]), pack16x2(f16, v[i + 4], v[i + 5]),
                             pack16x2(f16, v[i + 6], v[i + 7]));
    return;
  }
  uint4* oh = (uint4*)base;
  uint4* ol = (uint4*)(base + lo_off);
#pragma unroll
  for (int i = 0; i < NC; i += 8) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) split16x2(f16, v[i + 2 * q], v[i + 2 * q + 1], h[q], l[q]);
    oh[i / 8] = make_uint4(h[0], h[1], h[2], h[3]);
    ol[i / 8] = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

template <int NC>
__device__ __forceinline__ void store_cols(const IgemmArgs& a, size_t row, int c, const float* v) {
  store_cols_to<NC>(a.out, a.out_enc, a.out_stride, a.out_lo_off, a.fmt != 0, row, c, v);
}

// Sparse-conv epilogue of NC consecutive output channels [c0, c0+NC) of one row (the row's
// accumulator is drained from TMEM in chunks to keep the epilogue's register footprint small,
// which is what lets the kernel afford more gather warps): bias(BN) + residual + ReLU + store.
template <int COUT, int NC>
__device__ __forceinline__ void epilogue_chunk(const IgemmArgs& a, int row, int c0, float* v) {
  if (a.bias) {
#pragma unroll
    for (int c = 0; c < NC; ++c) v[c] += __ldg(a.bias + c0 + c);
  }
  if (a.residual) add_residual<NC>(a, (size_t)row, c0, v);
  if (a.relu) {
#pragma unroll
    for (int c = 0; c < NC; ++c) v[c] = fmaxf(v[c], 0.f);
  }
  if (a.dense) {
    const int4 q = __ldg(a.out_coors + row);
    const size_t hw = (size_t)a.H * a.W;
    float* dp = a.dense + ((size_t)q.x * COUT * a.D + q.y) * hw + (size_t)q.z * a.W + q.w + (size_t)c0 * a.D * hw;
#pragma unroll
    for (int c = 0; c < NC; ++c) dp[(size_t)c * a.D * hw] = v[c];
  } else {
    store_cols<NC>(a, (size_t)row, c0, v);
  }
}

// dense-linear epilogue of a whole COUT-wide row: bias / residual / LayerNorm / ReLU / store.
// sp: the tile's bias | ln_w | ln_b (3 x 128 floats) staged in shared memory by the epilogue warps -- with one thread per
// row every lane needs the same 3 x COUT parameters, and fetching them with 3 x COUT broadcast global loads per thread
// made the epilogue (not the main loop) the per-tile cost of the short-K GEMMs of the head (ncu: 16 k cycles per tile).
template <int COUT>
__device__ __forceinline__ void epilogue_row(const IgemmArgs& a, int row, int nt, float* v, const float* sp, int ks = 0) {
  if (a.k_splits > 1) {   // partial sum of one K range -> its own (m, n) slab; summed in fixed order by srf_layernorm
    float4* op = (float4*)((float*)a.out + ((size_t)ks * a.m_rows + row) * a.out_stride + nt * COUT);
#pragma unroll
    for (int c = 0; c < COUT; c += 4) op[c / 4] = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
    return;
  }
  const int col0 = nt * COUT;
  if (a.bias) {
#pragma unroll
    for (int c = 0; c < COUT; c += 4) {
      const float4 b4 = *reinterpret_cast<const float4*>(sp + c);
      v[c] += b4.x; v[c + 1] += b4.y; v[c + 2] += b4.z; v[c + 3] += b4.w;
    }
  }
  if (a.residual) add_residual<COUT>(a, (size_t)row, col0, v);
  if (a.ln) {
    // four partial sums: the reductions are latency chains in a single thread
    float m4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < COUT; c += 4) { m4[0] += v[c]; m4[1] += v[c + 1]; m4[2] += v[c + 2]; m4[3] += v[c + 3]; }
    const float mean = ((m4[0] + m4[1]) + (m4[2] + m4[3])) * (1.f / COUT);
    float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < COUT; c += 4) {
#pragma unroll
      for (int q = 0; q < 4; ++q) { const float d = v[c + q] - mean; q4[q] = fmaf(d, d, q4[q]); }
    }
    const float rstd = rsqrtf(((q4[0] + q4[1]) + (q4[2] + q4[3])) * (1.f / COUT) + a.ln_eps);
#pragma unroll
    for (int c = 0; c < COUT; c += 4) {
      const float4 w4 = *reinterpret_cast<const float4*>(sp + 128 + c), b4 = *reinterpret_cast<const float4*>(sp + 256 + c);
      v[c] = (v[c] - mean) * rstd * w4.x + b4.x;
      v[c + 1] = (v[c + 1] - mean) * rstd * w4.y + b4.y;
      v[c + 2] = (v[c + 2] - mean) * rstd * w4.z + b4.z;
      v[c + 3] = (v[c + 3] - mean) * rstd * w4.w + b4.w;
    }
  }
  if (a.relu) {
#pragma unroll
    for (int c = 0; c < COUT; ++c) v[c] = fmaxf(v[c], 0.f);
  }
  store_cols<COUT>(a, (size_t)row, col0, v);
  if (a.out2) {
    const long long n_total = enc_is_split(a.out_enc) ? a.out_stride / 2 : a.out_stride;   // logical columns of a row
    store_cols_to<COUT>(a.out2, a.out2_enc, enc_is_split(a.out2_enc) ? 2 * n_total : n_total, n_total, a.fmt != 0, (size_t)row, col0, v);
  }
}

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ void cp_async16_ca(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}


}  // namespace srf
