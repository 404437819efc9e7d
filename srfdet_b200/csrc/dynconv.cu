// DynamicConv interaction core (srfdet_head.py:2679-2686):
//   f = relu(LN_d( F(49,C) . P1(C,d) ));  g = relu(LN_C( f . P2(d,C) ))
// one CTA per proposal; F, P1, P2 and the 49 x d intermediate stay in shared memory, so
// the only HBM traffic is the RoI features, the generated parameters and the output.
#include "common.cuh"

namespace srf {

constexpr int DC_ROWS = 49;

template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ void st_f(T* p, float v);
template <>
__device__ __forceinline__ void st_f<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void st_f<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }

// LayerNorm + ReLU over rows of length n in shared memory, one warp per row
__device__ __forceinline__ void ln_relu_rows(float* rows, int nrows, int n, const float* __restrict__ g,
                                             const float* __restrict__ b, float eps) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int r = warp; r < nrows; r += nwarps) {
    float* x = rows + (size_t)r * n;
    float s = 0.f;
    for (int j = lane; j < n; j += 32) s += x[j];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / n;
    float v = 0.f;
    for (int j = lane; j < n; j += 32) { float d = x[j] - mean; v += d * d; }
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rstd = rsqrtf(v / n + eps);
    for (int j = lane; j < n; j += 32) x[j] = fmaxf((x[j] - mean) * rstd * __ldg(g + j) + __ldg(b + j), 0.f);
  }
}

template <typename TR, typename TP, typename TO>
__global__ void __launch_bounds__(256) dynconv_interact_kernel(const TR* __restrict__ roi, const TP* __restrict__ params,
                                                              int c, int d, const float* __restrict__ ln1_w,
                                                              const float* __restrict__ ln1_b, const float* __restrict__ ln2_w,
                                                              const float* __restrict__ ln2_b, float eps1, float eps2, TO* __restrict__ out) {
  extern __shared__ float sm[];
  float* sF = sm;                   // 49 x c   (later reused for the 49 x c output)
  float* sP1 = sF + DC_ROWS * c;    // c x d
  float* sP2 = sP1 + c * d;         // d x c
  float* sT = sP2 + c * d;          // 49 x d
  const int k = blockIdx.x;
  const TR* r = roi + (size_t)k * DC_ROWS * c;
  const TP* p = params + (size_t)k * 2 * c * d;
  for (int e = threadIdx.x; e < DC_ROWS * c; e += blockDim.x) sF[e] = to_f<TR>(r[e]);
  for (int e = threadIdx.x; e < 2 * c * d; e += blockDim.x) sP1[e] = to_f<TP>(p[e]);  // P1 then P2 contiguous
  __syncthreads();
  for (int e = threadIdx.x; e < DC_ROWS * d; e += blockDim.x) {
    const int s = e / d, j = e - s * d;
    const float* f = sF + s * c;
    float acc = 0.f;
    for (int i = 0; i < c; ++i) acc = fmaf(f[i], sP1[i * d + j], acc);
    sT[e] = acc;
  }
  __syncthreads();
  ln_relu_rows(sT, DC_ROWS, d, ln1_w, ln1_b, eps1);
  __syncthreads();
  for (int e = threadIdx.x; e < DC_ROWS * c; e += blockDim.x) {
    const int s = e / c, j = e - s * c;
    const float* t = sT + s * d;
    float acc = 0.f;
    for (int i = 0; i < d; ++i) acc = fmaf(t[i], sP2[i * c + j], acc);
    sF[e] = acc;
  }
  __syncthreads();
  ln_relu_rows(sF, DC_ROWS, c, ln2_w, ln2_b, eps2);
  __syncthreads();
  TO* o = out + (size_t)k * DC_ROWS * c;
  for (int e = threadIdx.x; e < DC_ROWS * c; e += blockDim.x) st_f<TO>(o + e, sF[e]);
}

// Register-tiled variant for the production shapes (C,D) = (128,32), (256,64): each thread
// owns RPT rows x 1 column of F.P1 and RPT rows x 4 columns of f.P2, operands are read from
// shared memory as float4 (row operands broadcast), ~4 FMA per shared-memory load.
template <int C, int D, typename TR, typename TP, typename TO>
__global__ void __launch_bounds__(256) dynconv_interact_tiled_kernel(const TR* __restrict__ roi, const TP* __restrict__ params,
                                                                    const float* __restrict__ ln1_w, const float* __restrict__ ln1_b,
                                                                    const float* __restrict__ ln2_w, const float* __restrict__ ln2_b,
                                                                    float eps1, float eps2, TO* __restrict__ out) {
  constexpr int G1 = 256 / D;                         // row groups in phase 1
  constexpr int RPT1 = (DC_ROWS + G1 - 1) / G1;
  constexpr int G2 = 256 / (C / 4);                   // row groups in phase 3
  constexpr int RPT2 = (DC_ROWS + G2 - 1) / G2;
  constexpr int FP = C + 4;                           // padded row pitch of F (keeps float4 alignment)
  constexpr int TP_ = D + 4;
  extern __shared__ __align__(16) float sm[];
  float* sF = sm;                      // 52 x FP (rows >= 49 zero)
  float* sP1 = sF + 52 * FP;           // C x D
  float* sP2 = sP1 + C * D;            // D x C
  float* sT = sP2 + C * D;             // 52 x TP_
  const int k = blockIdx.x;
  const TR* r = roi + (size_t)k * DC_ROWS * C;
  const TP* p = params + (size_t)k * 2 * C * D;
  for (int e = threadIdx.x; e < 52 * C; e += 256) {
    int s = e / C, i = e - s * C;
    sF[s * FP + i] = s < DC_ROWS ? to_f<TR>(r[e]) : 0.f;
  }
  for (int e = threadIdx.x; e < 2 * C * D; e += 256) sP1[e] = to_f<TP>(p[e]);
  __syncthreads();
  {  // phase 1: T = F . P1
    const int j = threadIdx.x % D, g = threadIdx.x / D;
    float acc[RPT1];
#pragma unroll
    for (int q = 0; q < RPT1; ++q) acc[q] = 0.f;
    const int r0 = g * RPT1;
    for (int i = 0; i < C; i += 4) {
      const float b0 = sP1[(i + 0) * D + j], b1 = sP1[(i + 1) * D + j], b2 = sP1[(i + 2) * D + j], b3 = sP1[(i + 3) * D + j];
#pragma unroll
      for (int q = 0; q < RPT1; ++q) {
        const int row = min(r0 + q, 51);
        const float4 f = *reinterpret_cast<const float4*>(sF + row * FP + i);
        acc[q] = fmaf(f.x, b0, fmaf(f.y, b1, fmaf(f.z, b2, fmaf(f.w, b3, acc[q]))));
      }
    }
#pragma unroll
    for (int q = 0; q < RPT1; ++q)
      if (r0 + q < 52) sT[(r0 + q) * TP_ + j] = acc[q];
  }
  __syncthreads();
  {  // LayerNorm(D) + ReLU per row, one warp per row
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int row = warp; row < DC_ROWS; row += 8) {
      float* x = sT + row * TP_;
      float s = 0.f;
      for (int j = lane; j < D; j += 32) s += x[j];
      for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s / D;
      float v = 0.f;
      for (int j = lane; j < D; j += 32) { float d = x[j] - mean; v += d * d; }
      for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      const float rstd = rsqrtf(v / D + eps1);
      for (int j = lane; j < D; j += 32) x[j] = fmaxf((x[j] - mean) * rstd * __ldg(ln1_w + j) + __ldg(ln1_b + j), 0.f);
    }
  }
  __syncthreads();
  {  // phase 3: G = T . P2  -> sF (F is dead)
    const int jc = (threadIdx.x % (C / 4)) * 4, g = threadIdx.x / (C / 4);
    float acc[RPT2][4];
#pragma unroll
    for (int q = 0; q < RPT2; ++q) { acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.f; }
    const int r0 = g * RPT2;
    for (int i = 0; i < D; i += 4) {
      const float4 w0 = *reinterpret_cast<const float4*>(sP2 + (i + 0) * C + jc);
      const float4 w1 = *reinterpret_cast<const float4*>(sP2 + (i + 1) * C + jc);
      const float4 w2 = *reinterpret_cast<const float4*>(sP2 + (i + 2) * C + jc);
      const float4 w3 = *reinterpret_cast<const float4*>(sP2 + (i + 3) * C + jc);
#pragma unroll
      for (int q = 0; q < RPT2; ++q) {
        const int row = min(r0 + q, 51);
        const float4 t = *reinterpret_cast<const float4*>(sT + row * TP_ + i);
        acc[q][0] = fmaf(t.x, w0.x, fmaf(t.y, w1.x, fmaf(t.z, w2.x, fmaf(t.w, w3.x, acc[q][0]))));
        acc[q][1] = fmaf(t.x, w0.y, fmaf(t.y, w1.y, fmaf(t.z, w2.y, fmaf(t.w, w3.y, acc[q][1]))));
        acc[q][2] = fmaf(t.x, w0.z, fmaf(t.y, w1.z, fmaf(t.z, w2.z, fmaf(t.w, w3.z, acc[q][2]))));
        acc[q][3] = fmaf(t.x, w0.w, fmaf(t.y, w1.w, fmaf(t.z, w2.w, fmaf(t.w, w3.w, acc[q][3]))));
      }
    }
    __syncthreads();   // every thread is done reading sF-independent data; now overwrite F rows
#pragma unroll
    for (int q = 0; q < RPT2; ++q)
      if (r0 + q < DC_ROWS) *reinterpret_cast<float4*>(sF + (r0 + q) * FP + jc) = make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]);
  }
  __syncthreads();
  {  // LayerNorm(C) + ReLU per row, write out
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    TO* o = out + (size_t)k * DC_ROWS * C;
    for (int row = warp; row < DC_ROWS; row += 8) {
      const float* x = sF + row * FP;
      float s = 0.f;
      for (int j = lane; j < C; j += 32) s += x[j];
      for (int o2 = 16; o2; o2 >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o2);
      const float mean = s / C;
      float v = 0.f;
      for (int j = lane; j < C; j += 32) { float d = x[j] - mean; v += d * d; }
      for (int o2 = 16; o2; o2 >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o2);
      const float rstd = rsqrtf(v / C + eps2);
      for (int j = lane; j < C; j += 32)
        st_f<TO>(o + (size_t)row * C + j, fmaxf((x[j] - mean) * rstd * __ldg(ln2_w + j) + __ldg(ln2_b + j), 0.f));
    }
  }
}

template <int C, int D, typename TR, typename TP, typename TO>
static int launch_dc_tiled(const void* roi, const void* params, int k, const float* a, const float* b, const float* e,
                           const float* f, float eps1, float eps2, void* out, cudaStream_t st) {
  size_t smem = (size_t)(52 * (C + 4) + 2 * C * D + 52 * (D + 4)) * sizeof(float);
  auto kern = dynconv_interact_tiled_kernel<C, D, TR, TP, TO>;
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) { set_error("dynconv: cannot get %zu B shared memory: %s", smem, cudaGetErrorString(err)); return SRF_ERR_CUDA; }
  SRF_COUNT(1);
  kern<<<k, 256, smem, st>>>((const TR*)roi, (const TP*)params, a, b, e, f, eps1, eps2, (TO*)out);
  return SRF_OK;
}

// ------------------------------------------------------------------------------------
// Tensor-core variant (16-bit and split modes): the two per-proposal GEMMs are 64x128x32 /
// 64x32x128-sized, far below a tcgen05 tile, so they run on warp-level mma.sync.m16n8k16
// (bf16 or f16 in, fp32 accumulate) with ldmatrix-fed fragments; 8 warps = 4 row tiles x 2
// column halves.  LayerNorms run on the fp32 accumulator fragments.  SPLIT: every operand is
// held as hi + lo in shared memory (RoI features, generated parameters and the intermediate),
// each GEMM is three passes  Al.Bh + Ah.Bl + Ah.Bh  into the same accumulators.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
template <bool F16>
__device__ __forceinline__ void mma_16816(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  if (F16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// warp-tile GEMM: acc[NT][4] += A(16 x K, smem pitch lda) . B(K x (NT*8), smem pitch ldb, row-major [k][n])
template <int K, int NT, bool F16>
__device__ __forceinline__ void warp_gemm(const uint16_t* sA, int lda, const uint16_t* sB, int ldb, float (*acc)[4]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int k0 = 0; k0 < K; k0 += 16) {
    uint32_t a0, a1, a2, a3;
    ldsm_x4((uint32_t)__cvta_generic_to_shared(sA + (size_t)((lane & 7) + ((lane >> 3) & 1) * 8) * lda + k0 + (lane >> 4) * 8), a0, a1, a2, a3);
#pragma unroll
    for (int nt = 0; nt < NT; nt += 2) {
      uint32_t b0, b1, b2, b3;   // (k 0-7, n-tile nt), (k 8-15, nt), (k 0-7, nt+1), (k 8-15, nt+1)
      ldsm_x4_t((uint32_t)__cvta_generic_to_shared(sB + (size_t)(k0 + (lane & 7) + ((lane >> 3) & 1) * 8) * ldb + (nt + (lane >> 4)) * 8), b0, b1, b2, b3);
      mma_16816<F16>(acc[nt], a0, a1, a2, a3, b0, b1);
      mma_16816<F16>(acc[nt + 1], a0, a1, a2, a3, b2, b3);
    }
  }
}

// LayerNorm + ReLU applied directly on mma accumulator fragments.  A row of the 64 x N result
// is spread over the 4 lanes of a quad (columns) and over the 2 warps that own the two column
// halves; only the per-row partial sums cross warps (through `st`, 64 rows x 2 halves).
// acc[i][0..1] belong to row r0, acc[i][2..3] to row r0+8; column of acc[i][e] = colbase + i*8 + 2t + (e&1).
template <int NT, int N, typename Emit>
__device__ __forceinline__ void fragment_ln_relu(float (*acc)[4], float* st, float* st2, int r0, int nh, int colbase, int t,
                                                 const float* __restrict__ gw, const float* __restrict__ gb, float eps, Emit emit) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int i = 0; i < NT; ++i) { s0 += acc[i][0] + acc[i][1]; s1 += acc[i][2] + acc[i][3]; }
  s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
  s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
  if (t == 0) { st[r0 * 2 + nh] = s0; st[(r0 + 8) * 2 + nh] = s1; }
  __syncthreads();
  const float m0 = (st[r0 * 2] + st[r0 * 2 + 1]) * (1.f / N), m1 = (st[(r0 + 8) * 2] + st[(r0 + 8) * 2 + 1]) * (1.f / N);
  float q0 = 0.f, q1 = 0.f;
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    float d;
    d = acc[i][0] - m0; q0 += d * d; d = acc[i][1] - m0; q0 += d * d;
    d = acc[i][2] - m1; q1 += d * d; d = acc[i][3] - m1; q1 += d * d;
  }
  q0 += __shfl_xor_sync(0xffffffffu, q0, 1); q0 += __shfl_xor_sync(0xffffffffu, q0, 2);
  q1 += __shfl_xor_sync(0xffffffffu, q1, 1); q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
  if (t == 0) { st2[r0 * 2 + nh] = q0; st2[(r0 + 8) * 2 + nh] = q1; }
  __syncthreads();
  const float rs0 = rsqrtf((st2[r0 * 2] + st2[r0 * 2 + 1]) * (1.f / N) + eps);
  const float rs1 = rsqrtf((st2[(r0 + 8) * 2] + st2[(r0 + 8) * 2 + 1]) * (1.f / N) + eps);
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    const int col = colbase + i * 8 + 2 * t;
    const float w0 = __ldg(gw + col), w1 = __ldg(gw + col + 1), b0 = __ldg(gb + col), b1 = __ldg(gb + col + 1);
    emit(r0, col, fmaxf((acc[i][0] - m0) * rs0 * w0 + b0, 0.f), fmaxf((acc[i][1] - m0) * rs0 * w1 + b1, 0.f));
    emit(r0 + 8, col, fmaxf((acc[i][2] - m1) * rs1 * w0 + b0, 0.f), fmaxf((acc[i][3] - m1) * rs1 * w1 + b1, 0.f));
  }
}

struct DcMmaArgs {
  const void* roi;      // (K, 49, C)
  int roi_enc;          // SRF_F32, the kernel's 16-bit format, or its split form
  const void* params;   // (K, 2*C*D)
  int param_enc;        // SRF_F32 or the kernel's 16-bit format
  const float *ln1_w, *ln1_b, *ln2_w, *ln2_b;
  float eps1, eps2;
  uint16_t* out;        // (K, 49*C) 16-bit, or split rows [hi(49C) | lo(49C)]
};

// 8 consecutive logical values (row, col..col+7) of a (rows, c) buffer -> hi (and lo) words
template <bool F16, bool SPLIT>
__device__ __forceinline__ void load8_split(const void* base, int enc, size_t row, int c, int col, uint4& hi, uint4& lo) {
  lo = make_uint4(0u, 0u, 0u, 0u);
  if (enc == SRF_F32) {
    const float4* p = reinterpret_cast<const float4*>((const float*)base + row * c + col);
    const float4 f0 = __ldg(p), f1 = __ldg(p + 1);
    uint32_t h[4], l[4];
    split16x2(F16, f0.x, f0.y, h[0], l[0]);
    split16x2(F16, f0.z, f0.w, h[1], l[1]);
    split16x2(F16, f1.x, f1.y, h[2], l[2]);
    split16x2(F16, f1.z, f1.w, h[3], l[3]);
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    if (SPLIT) lo = make_uint4(l[0], l[1], l[2], l[3]);
  } else if (enc_is_split(enc)) {
    const uint16_t* p = (const uint16_t*)base + row * 2 * c + col;
    hi = __ldg(reinterpret_cast<const uint4*>(p));
    if (SPLIT) lo = __ldg(reinterpret_cast<const uint4*>(p + c));
  } else {
    hi = __ldg(reinterpret_cast<const uint4*>((const uint16_t*)base + row * c + col));
  }
}

// resident CTAs per SM aimed at by the 16-bit form for C <= 128 (A/B knob; 900 proposals over 148 SMs: 3 -> three rounds, 4 -> two)
#ifndef SRF_DC_MINB
#define SRF_DC_MINB 4
#endif
template <int C, int D, bool F16, bool SPLIT>
__global__ void __launch_bounds__(256, C <= 128 ? (SPLIT ? 2 : SRF_DC_MINB) : 1) dynconv_interact_mma_kernel(const DcMmaArgs a) {
  constexpr int LDF = C + 8, LDP1 = D + 8, LDP2 = C + 8, LDT = D + 8;   // 16-bit pitches (odd multiples of 16 B)
  constexpr int NP = SPLIT ? 2 : 1;                                      // hi (+ lo) copies of every operand
  constexpr int SZF = 64 * LDF, SZP1 = C * LDP1, SZP2 = D * LDP2, SZT = 64 * LDT;
  extern __shared__ __align__(16) uint8_t smraw[];
  uint16_t* sF = reinterpret_cast<uint16_t*>(smraw);      // [NP] 64 x LDF (rows >= 49 zero)
  uint16_t* sP1 = sF + NP * SZF;                          // [NP] C x LDP1
  uint16_t* sP2 = sP1 + NP * SZP1;                        // [NP] D x LDP2
  uint16_t* sTb = sP2 + NP * SZP2;                        // [NP] 64 x LDT  relu(LN(F.P1))
  float* st = reinterpret_cast<float*>(sTb + NP * SZT);   // 4 x (64 x 2) row partial sums
  const int k = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int e = threadIdx.x; e < C * D / 8; e += 256) {          // 8-element pieces of P1 (C x D) and P2 (D x C)
    uint4 hi, lo;
    load8_split<F16, SPLIT>(a.params, a.param_enc, (size_t)k, 2 * C * D, e * 8, hi, lo);
    const int row1 = (e * 8) / D, col1 = (e * 8) % D;
    *reinterpret_cast<uint4*>(sP1 + row1 * LDP1 + col1) = hi;
    if (SPLIT) *reinterpret_cast<uint4*>(sP1 + SZP1 + row1 * LDP1 + col1) = lo;
    load8_split<F16, SPLIT>(a.params, a.param_enc, (size_t)k, 2 * C * D, C * D + e * 8, hi, lo);
    const int row2 = (e * 8) / C, col2 = (e * 8) % C;
    *reinterpret_cast<uint4*>(sP2 + row2 * LDP2 + col2) = hi;
    if (SPLIT) *reinterpret_cast<uint4*>(sP2 + SZP2 + row2 * LDP2 + col2) = lo;
  }
  for (int e = threadIdx.x; e < 64 * C / 8; e += 256) {          // 8 channels per thread-step
    const int s = (e * 8) / C, i = (e * 8) % C;
    uint4 hi = make_uint4(0u, 0u, 0u, 0u), lo = hi;
    if (s < DC_ROWS) load8_split<F16, SPLIT>(a.roi, a.roi_enc, (size_t)k * DC_ROWS + s, C, i, hi, lo);
    *reinterpret_cast<uint4*>(sF + s * LDF + i) = hi;
    if (SPLIT) *reinterpret_cast<uint4*>(sF + SZF + s * LDF + i) = lo;
  }
  __syncthreads();
  const int mt = warp >> 1, nh = warp & 1;
  const int g = lane >> 2, t = lane & 3;
  {  // T1 = F . P1 (64 x D), warp tile 16 x D/2; LayerNorm(D) + ReLU on the fragments -> 16-bit A operand
    constexpr int NT = D / 16;
    float acc[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    if (SPLIT) {
      warp_gemm<C, NT, F16>(sF + SZF + mt * 16 * LDF, LDF, sP1 + nh * (D / 2), LDP1, acc);
      warp_gemm<C, NT, F16>(sF + mt * 16 * LDF, LDF, sP1 + SZP1 + nh * (D / 2), LDP1, acc);
    }
    warp_gemm<C, NT, F16>(sF + mt * 16 * LDF, LDF, sP1 + nh * (D / 2), LDP1, acc);
    fragment_ln_relu<NT, D>(acc, st, st + 128, mt * 16 + g, nh, nh * (D / 2), t, a.ln1_w, a.ln1_b, a.eps1,
                            [&](int row, int col, float y0, float y1) {
                              if (row >= DC_ROWS) { y0 = 0.f; y1 = 0.f; }
                              uint32_t h, l;
                              split16x2(F16, y0, y1, h, l);
                              *reinterpret_cast<uint32_t*>(sTb + row * LDT + col) = h;
                              if (SPLIT) *reinterpret_cast<uint32_t*>(sTb + SZT + row * LDT + col) = l;
                            });
  }
  __syncthreads();
  {  // G = T . P2 (64 x C), warp tile 16 x C/2; LayerNorm(C) + ReLU on the fragments -> global
    constexpr int NT = C / 16;
    float acc[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    if (SPLIT) {
      warp_gemm<D, NT, F16>(sTb + SZT + mt * 16 * LDT, LDT, sP2 + nh * (C / 2), LDP2, acc);
      warp_gemm<D, NT, F16>(sTb + mt * 16 * LDT, LDT, sP2 + SZP2 + nh * (C / 2), LDP2, acc);
    }
    warp_gemm<D, NT, F16>(sTb + mt * 16 * LDT, LDT, sP2 + nh * (C / 2), LDP2, acc);
    uint16_t* o = a.out + (size_t)k * DC_ROWS * C * NP;
    fragment_ln_relu<NT, C>(acc, st + 256, st + 384, mt * 16 + g, nh, nh * (C / 2), t, a.ln2_w, a.ln2_b, a.eps2,
                            [&](int row, int col, float y0, float y1) {
                              if (row >= DC_ROWS) return;
                              uint32_t h, l;
                              split16x2(F16, y0, y1, h, l);
                              *reinterpret_cast<uint32_t*>(o + (size_t)row * C + col) = h;
                              if (SPLIT) *reinterpret_cast<uint32_t*>(o + (size_t)DC_ROWS * C + (size_t)row * C + col) = l;
                            });
  }
}

template <int C, int D, bool F16, bool SPLIT>
static int launch_dc_mma(const DcMmaArgs& a, int k, cudaStream_t st) {
  constexpr int NP = SPLIT ? 2 : 1;
  size_t smem = (size_t)NP * (64 * (C + 8) + C * (D + 8) + D * (C + 8) + 64 * (D + 8)) * 2 + 4 * 128 * sizeof(float);
  auto kern = dynconv_interact_mma_kernel<C, D, F16, SPLIT>;
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) { set_error("dynconv mma: cannot get %zu B shared memory: %s", smem, cudaGetErrorString(err)); return SRF_ERR_CUDA; }
  SRF_COUNT(1);
  kern<<<k, 256, smem, st>>>(a);
  return SRF_OK;
}

template <int C, int D>
static int dispatch_dc_mma(const DcMmaArgs& a, int k, bool f16, bool split, cudaStream_t st) {
  if (f16) return split ? launch_dc_mma<C, D, true, true>(a, k, st) : launch_dc_mma<C, D, true, false>(a, k, st);
  return split ? launch_dc_mma<C, D, false, true>(a, k, st) : launch_dc_mma<C, D, false, false>(a, k, st);
}

template <typename TR, typename TP, typename TO>
static int launch_dc(const void* roi, const void* params, int k, int c, int d, const float* a, const float* b,
                     const float* e, const float* f, float eps1, float eps2, void* out, cudaStream_t st) {
  if (c == 128 && d == 32) return launch_dc_tiled<128, 32, TR, TP, TO>(roi, params, k, a, b, e, f, eps1, eps2, out, st);
  if (c == 256 && d == 64) return launch_dc_tiled<256, 64, TR, TP, TO>(roi, params, k, a, b, e, f, eps1, eps2, out, st);
  size_t smem = (size_t)(DC_ROWS * c + 2 * c * d + DC_ROWS * d) * sizeof(float);
  auto kern = dynconv_interact_kernel<TR, TP, TO>;
  if (smem > 48 * 1024) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) { set_error("dynconv: cannot get %zu B shared memory: %s", smem, cudaGetErrorString(err)); return SRF_ERR_CUDA; }
  }
  SRF_COUNT(1);
  kern<<<k, 256, smem, st>>>((const TR*)roi, (const TP*)params, c, d, a, b, e, f, eps1, eps2, (TO*)out);
  return SRF_OK;
}

}  // namespace srf

using namespace srf;

extern "C" {

int srf_dynconv_interact_tc(const void* roi, int32_t roi_enc, const void* params, int32_t param_enc, int32_t k,
                            int32_t c, int32_t d, const float* ln1_w, const float* ln1_b, float ln1_eps, const float* ln2_w,
                            const float* ln2_b, float ln2_eps, void* out, int32_t out_enc, void* stream) {
  SRF_CHECK_ARG(roi && params && out && ln1_w && ln1_b && ln2_w && ln2_b, "srf_dynconv_interact: null arg");
  SRF_CHECK_ARG(k >= 0 && c >= 1 && c <= 256 && d >= 1 && d <= 64, "srf_dynconv_interact: need c<=256, d<=64");
  if (k == 0) return SRF_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if (enc_is_16(out_enc) && ((c == 128 && d == 32) || (c == 256 && d == 64))) {
    // tensor-core kernel: the output encoding fixes the operand format (bf16 / f16) and the split form
    const bool f16 = enc_is_f16(out_enc), split = enc_is_split(out_enc);
    const int plain = f16 ? SRF_F16 : SRF_BF16;
    SRF_CHECK_ARG(roi_enc == SRF_F32 || roi_enc == plain || (split && roi_enc == out_enc),
                  "srf_dynconv_interact: RoI features must be f32 or match the output's 16-bit format");
    SRF_CHECK_ARG(param_enc == SRF_F32 || (!split && param_enc == plain), "srf_dynconv_interact: parameters must be f32%s",
                  split ? "" : " or the output's 16-bit format");
    DcMmaArgs a{roi, roi_enc, params, param_enc, ln1_w, ln1_b, ln2_w, ln2_b, ln1_eps, ln2_eps, (uint16_t*)out};
    rc = c == 128 ? dispatch_dc_mma<128, 32>(a, k, f16, split, st) : dispatch_dc_mma<256, 64>(a, k, f16, split, st);
  } else {
    const bool rb = roi_enc == SRF_BF16, pb = param_enc == SRF_BF16, ob = out_enc == SRF_BF16;
    SRF_CHECK_ARG((rb || roi_enc == SRF_F32) && (pb || param_enc == SRF_F32) && (ob || out_enc == SRF_F32),
                  "srf_dynconv_interact: the SIMT kernel takes f32 / bf16 buffers (other encodings need (c,d) = (128,32) or (256,64))");
    if (!rb && !pb && !ob) rc = launch_dc<float, float, float>(roi, params, k, c, d, ln1_w, ln1_b, ln2_w, ln2_b, ln1_eps, ln2_eps, out, st);
    else if (rb && pb && ob) rc = launch_dc<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16>(roi, params, k, c, d, ln1_w, ln1_b, ln2_w, ln2_b, ln1_eps, ln2_eps, out, st);
    else if (!rb && pb && ob) rc = launch_dc<float, __nv_bfloat16, __nv_bfloat16>(roi, params, k, c, d, ln1_w, ln1_b, ln2_w, ln2_b, ln1_eps, ln2_eps, out, st);
    else if (rb && !pb && ob) rc = launch_dc<__nv_bfloat16, float, __nv_bfloat16>(roi, params, k, c, d, ln1_w, ln1_b, ln2_w, ln2_b, ln1_eps, ln2_eps, out, st);
    else { set_error("srf_dynconv_interact: unsupported dtype combination"); return SRF_ERR_UNSUPPORTED; }
  }
  if (rc) return rc;
  SRF_LAUNCH_CHECK();
  return SRF_OK;
}

int srf_dynconv_interact(const void* roi, int32_t roi_dtype, const void* params, int32_t param_dtype, int32_t k,
                         int32_t c, int32_t d, const float* ln1_w, const float* ln1_b, const float* ln2_w,
                         const float* ln2_b, void* out, int32_t out_dtype, void* stream) {
  return srf_dynconv_interact_tc(roi, roi_dtype, params, param_dtype, k, c, d, ln1_w, ln1_b, 1e-5f, ln2_w, ln2_b, 1e-5f, out,
                                 out_dtype, stream);
}

}  // extern "C"
