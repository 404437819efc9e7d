// Tensor-core self-attention core over the proposals (nn.MultiheadAttention between in_proj and
// out_proj, /root/reference/mmdet3d_plugin/models/dense_heads/sparse_heads/srfdet_head.py:2281-2285)
// for the 16-bit precision modes: softmax(q k^T / sqrt(hd)) v per (sample, head).
//
// One CTA = 64 queries of one (sample, head): 4 warps x 16 queries.  Keys / values of the head are
// streamed in 64-key chunks: fp32 qkv rows -> registers (one chunk ahead of the math) -> 16-bit
// shared-memory tiles (K row-major, V transposed, both padded so the fragment loads below are
// bank-conflict free).  Per chunk a warp issues  S = Q K^T  (mma.sync m16n8k16, fp32 accumulate),
// runs the online softmax on the accumulator fragments (exp2, quad shuffles for the row maxima)
// and feeds P straight back as the A operand of  O += P V : the accumulator layout of two adjacent
// 8-key tiles IS the A-fragment layout of one 16-key step, so the probabilities never leave registers.
// SPLIT (the tensor-core form of the fp32 mode, output rows [hi | lo]): q, k, v and the probabilities are held as hi + lo
// 16-bit pairs and every product is three MMAs  lo.hi + hi.lo + hi.hi  into the same fp32 accumulator (exact to ~2^-21);
// the fp32 FFMA kernel (head_tail.cu) remains the path of fp32 OUTPUT rows and of other head widths.
#include "common.cuh"

namespace srf {

constexpr int ATM_Q = 64, ATM_K = 64, ATM_T = 128;

template <bool F16>
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if (F16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int HD, bool F16, bool SPLIT>
__global__ void __launch_bounds__(ATM_T) mha_attention_mma_kernel(const float* __restrict__ qkv, int n_p, int n_heads, float scale_log2e,
                                                                  uint16_t* __restrict__ out) {
  constexpr int NPART = SPLIT ? 2 : 1;   // hi (+ lo) copies of the K / V tiles
  constexpr int KS = HD + 8;          // K tile row stride (16-bit elements): 12 / 20 words -> the 8 rows of a fragment hit distinct banks
  constexpr int VS = ATM_K + 8;       // V^T tile row stride: 36 words
  constexpr int F4 = HD / 4;          // float4 per key row
  constexpr int LD = ATM_K * F4 / ATM_T;   // float4 per thread per chunk, for K and for V
  __shared__ __align__(16) uint16_t sK[2][NPART][ATM_K * KS];
  __shared__ __align__(16) uint16_t sV[2][NPART][HD * VS];
  const int C = n_heads * HD;
  const int b = blockIdx.z, h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const size_t row0 = (size_t)b * n_p;
  const int q0 = blockIdx.x * ATM_Q + warp * 16;
  const float* base = qkv + row0 * 3 * C + h * HD;

  // Q fragments (pre-scaled by log2(e) / sqrt(hd): the softmax below runs on exp2)
  uint32_t qa[HD / 16][4], ql[SPLIT ? HD / 16 : 1][4];
#pragma unroll
  for (int ks = 0; ks < HD / 16; ++ks)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int q = q0 + g + (r & 1) * 8, d = ks * 16 + 2 * t + (r >> 1) * 8;
      float2 v = make_float2(0.f, 0.f);
      if (q < n_p) v = __ldg(reinterpret_cast<const float2*>(base + (size_t)q * 3 * C + d));
      if (SPLIT) {
        uint32_t hi, lo;
        split16x2(F16, v.x * scale_log2e, v.y * scale_log2e, hi, lo);
        qa[ks][r] = hi;
        ql[ks][r] = lo;
      } else {
        qa[ks][r] = pack16x2(F16, v.x * scale_log2e, v.y * scale_log2e);
      }
    }

  float4 rk[LD], rv[LD];
  auto prefetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < LD; ++i) {
      const int e = threadIdx.x + i * ATM_T, key = e / F4, d4 = e % F4;
      rk[i] = rv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + key < n_p) {
        const float* p = base + (size_t)(k0 + key) * 3 * C + C + d4 * 4;
        rk[i] = __ldg(reinterpret_cast<const float4*>(p));
        rv[i] = __ldg(reinterpret_cast<const float4*>(p + C));
      }
    }
  };
  auto stage = [&](int buf) {
#pragma unroll
    for (int i = 0; i < LD; ++i) {
      const int e = threadIdx.x + i * ATM_T, key = e / F4, d4 = e % F4;
      uint32_t kh0, kl0, kh1, kl1, vh0, vl0, vh1, vl1;
      split16x2(F16, rk[i].x, rk[i].y, kh0, kl0);
      split16x2(F16, rk[i].z, rk[i].w, kh1, kl1);
      split16x2(F16, rv[i].x, rv[i].y, vh0, vl0);
      split16x2(F16, rv[i].z, rv[i].w, vh1, vl1);
      *reinterpret_cast<uint2*>(&sK[buf][0][key * KS + d4 * 4]) = make_uint2(kh0, kh1);
      sV[buf][0][(d4 * 4 + 0) * VS + key] = (uint16_t)(vh0 & 0xffffu);
      sV[buf][0][(d4 * 4 + 1) * VS + key] = (uint16_t)(vh0 >> 16);
      sV[buf][0][(d4 * 4 + 2) * VS + key] = (uint16_t)(vh1 & 0xffffu);
      sV[buf][0][(d4 * 4 + 3) * VS + key] = (uint16_t)(vh1 >> 16);
      if (SPLIT) {
        *reinterpret_cast<uint2*>(&sK[buf][NPART - 1][key * KS + d4 * 4]) = make_uint2(kl0, kl1);
        sV[buf][NPART - 1][(d4 * 4 + 0) * VS + key] = (uint16_t)(vl0 & 0xffffu);
        sV[buf][NPART - 1][(d4 * 4 + 1) * VS + key] = (uint16_t)(vl0 >> 16);
        sV[buf][NPART - 1][(d4 * 4 + 2) * VS + key] = (uint16_t)(vl1 & 0xffffu);
        sV[buf][NPART - 1][(d4 * 4 + 3) * VS + key] = (uint16_t)(vl1 >> 16);
      }
    }
  };

  float o[HD / 8][4];
#pragma unroll
  for (int nd = 0; nd < HD / 8; ++nd) o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.f;
  float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};

  const int n_chunks = (n_p + ATM_K - 1) / ATM_K;
  prefetch(0);
  for (int c = 0; c < n_chunks; ++c) {
    const int buf = c & 1, k0 = c * ATM_K;
    stage(buf);
    __syncthreads();          // (the buffer written two chunks later is only reached after the next chunk's barrier)
    if (c + 1 < n_chunks) prefetch(k0 + ATM_K);
    const uint16_t* K = sK[buf][0];
    const uint16_t* V = sV[buf][0];
    const uint16_t* Kl = sK[buf][NPART - 1];      // lo planes (SPLIT only)
    const uint16_t* Vl = sV[buf][NPART - 1];
    // S = Q K^T : 16 queries x 64 keys
    float s[ATM_K / 8][4];
#pragma unroll
    for (int nt = 0; nt < ATM_K / 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks) {
        const int ko = (nt * 8 + g) * KS + ks * 16 + 2 * t;
        const uint16_t* kp = K + ko;
        if (SPLIT) {   // small cross terms first, then the main product
          const uint16_t* kq = Kl + ko;
          mma16816<F16>(s[nt], ql[ks], *reinterpret_cast<const uint32_t*>(kp), *reinterpret_cast<const uint32_t*>(kp + 8));
          mma16816<F16>(s[nt], qa[ks], *reinterpret_cast<const uint32_t*>(kq), *reinterpret_cast<const uint32_t*>(kq + 8));
        }
        mma16816<F16>(s[nt], qa[ks], *reinterpret_cast<const uint32_t*>(kp), *reinterpret_cast<const uint32_t*>(kp + 8));
      }
    }
    // online softmax on the fragments: thread holds rows g (values 0,1) and g + 8 (values 2,3), keys nt*8 + 2t, +1
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < ATM_K / 8; ++nt) {
      const int key = k0 + nt * 8 + 2 * t;
      if (key >= n_p) s[nt][0] = s[nt][2] = -INFINITY;
      if (key + 1 >= n_p) s[nt][1] = s[nt][3] = -INFINITY;
      mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
    }
    float corr[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float mn = fmaxf(m[r], mx[r]);      // finite: every chunk holds at least one live key
      corr[r] = exp2f(m[r] - mn);
      m[r] = mn;
      l[r] *= corr[r];
    }
#pragma unroll
    for (int nd = 0; nd < HD / 8; ++nd) {
      o[nd][0] *= corr[0]; o[nd][1] *= corr[0];
      o[nd][2] *= corr[1]; o[nd][3] *= corr[1];
    }
#pragma unroll
    for (int nt = 0; nt < ATM_K / 8; ++nt) {
      s[nt][0] = exp2f(s[nt][0] - m[0]); s[nt][1] = exp2f(s[nt][1] - m[0]);
      s[nt][2] = exp2f(s[nt][2] - m[1]); s[nt][3] = exp2f(s[nt][3] - m[1]);
      l[0] += s[nt][0] + s[nt][1];
      l[1] += s[nt][2] + s[nt][3];
    }
    // O += P V : the accumulator fragments of key tiles 2kk, 2kk+1 are the A fragment of 16-key step kk
#pragma unroll
    for (int kk = 0; kk < ATM_K / 16; ++kk) {
      uint32_t pa[4], pl[4];
      split16x2(F16, s[2 * kk][0], s[2 * kk][1], pa[0], pl[0]);
      split16x2(F16, s[2 * kk][2], s[2 * kk][3], pa[1], pl[1]);
      split16x2(F16, s[2 * kk + 1][0], s[2 * kk + 1][1], pa[2], pl[2]);
      split16x2(F16, s[2 * kk + 1][2], s[2 * kk + 1][3], pa[3], pl[3]);
#pragma unroll
      for (int nd = 0; nd < HD / 8; ++nd) {
        const int vo = (nd * 8 + g) * VS + kk * 16 + 2 * t;
        const uint16_t* vp = V + vo;
        if (SPLIT) {
          const uint16_t* vq = Vl + vo;
          mma16816<F16>(o[nd], pl, *reinterpret_cast<const uint32_t*>(vp), *reinterpret_cast<const uint32_t*>(vp + 8));
          mma16816<F16>(o[nd], pa, *reinterpret_cast<const uint32_t*>(vq), *reinterpret_cast<const uint32_t*>(vq + 8));
        }
        mma16816<F16>(o[nd], pa, *reinterpret_cast<const uint32_t*>(vp), *reinterpret_cast<const uint32_t*>(vp + 8));
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
    l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
    const int q = q0 + g + r * 8;
    if (q < n_p) {
      const float inv = 1.f / l[r];
      uint16_t* op = out + (row0 + q) * (size_t)(NPART * C) + h * HD + 2 * t;     // split rows are [hi(C) | lo(C)]
#pragma unroll
      for (int nd = 0; nd < HD / 8; ++nd) {
        uint32_t hi, lo;
        split16x2(F16, o[nd][2 * r] * inv, o[nd][2 * r + 1] * inv, hi, lo);
        *reinterpret_cast<uint32_t*>(op + nd * 8) = hi;
        if (SPLIT) *reinterpret_cast<uint32_t*>(op + C + nd * 8) = lo;
      }
    }
  }
}

// true when the tensor-core kernel took the call (16-bit outputs, or their split forms: the fp32 mode)
bool mha_attention_mma_launch(const float* qkv, int n_batch, int n_p, int n_heads, int head_dim, void* out, int out_enc, cudaStream_t st) {
  if (!enc_is_16(out_enc) || !(head_dim == 16 || head_dim == 32) || n_p < 1) return false;
  if ((n_heads * head_dim) % 4 != 0 || ((uintptr_t)qkv & 15) != 0 || ((uintptr_t)out & 3) != 0) return false;
  const float scale_log2e = 1.4426950408889634f / sqrtf((float)head_dim);
  const dim3 grid(cdiv(n_p, ATM_Q), n_heads, n_batch);
  const bool f16 = enc_is_f16(out_enc), split = enc_is_split(out_enc);
  uint16_t* o = (uint16_t*)out;
#define SRF_ATT(hd_, f16_, split_) mha_attention_mma_kernel<hd_, f16_, split_><<<grid, ATM_T, 0, st>>>(qkv, n_p, n_heads, scale_log2e, o)
  if (head_dim == 16) {
    if (split) { if (f16) SRF_ATT(16, true, true); else SRF_ATT(16, false, true); }
    else { if (f16) SRF_ATT(16, true, false); else SRF_ATT(16, false, false); }
  } else {
    if (split) { if (f16) SRF_ATT(32, true, true); else SRF_ATT(32, false, true); }
    else { if (f16) SRF_ATT(32, true, false); else SRF_ATT(32, false, false); }
  }
#undef SRF_ATT
  return true;
}

}  // namespace srf
