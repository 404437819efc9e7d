// Implicit gather-GEMM on tcgen05 / TMEM (sm_100a): the BF16 mode of the sparse
// convolution and of the dense projections of the region-fusion
// head (rows a9, a10).
//
//   out[m, :] = epi( sum_k  A[idx_k(m), k-slice] . W_k  + bias )
//
// sparse conv : idx_k(m) = nbr[k][m] (rulebook, -1 -> zero row), k-slice = whole row
// dense linear: idx_k(m) = m, k-slice = columns [k*CIN, (k+1)*CIN)
//
// Output-stationary: one CTA owns a 128-row output tile; the 128xCOUT fp32 accumulator
// lives in TMEM (double buffered so the epilogue of tile i overlaps the main loop of tile
// i+1) and is written exactly once with bias / residual / ReLU / LayerNorm fused.
//
// Warp roles (288 threads):
//   warps 0-3  epilogue  : tcgen05.ld (warp w owns TMEM lanes 32w..32w+31 = rows), fused
//                          epilogue, global stores
//   warps 4-7  producers : one thread per tile row; cp.async 16-byte row chunks (zero-fill
//                          for missing neighbours) + the W_k tile into a STAGES-deep ring
//   warp  8    MMA       : one lane issues tcgen05.mma (M=128, N=COUT, K=16) per 16 input
//                          channels, tcgen05.commit releases ring slots / publishes the
//                          accumulator
// TMA cannot express this gather (row indices are data dependent and -1 rows must read
// zeros), hence cp.async into the canonical no-swizzle K-major core-matrix layout:
//   operand byte offset(row r, 16B chunk c) = c * LBO + r * 16     (SBO = 128)
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
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
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

struct IgemmArgs {
  const __nv_bfloat16* in;
  long long in_stride;  // elements between A rows
  long long k_stride;   // element offset of the k-th slice inside a row (0 for sparse conv)
  const int32_t* nbr;   // (kvol, cap_out) or null (dense: identity rows)
  const uint32_t* tile_mask;
  const int32_t* d_n_out;
  int cap_out;  // rows bound (multiple of 128 for the sparse path)
  int m_rows;   // dense: number of rows
  int kvol;
  int n_tiles;
  const __nv_bfloat16* w;  // packed [n_tile][k][CIN/8][COUT][8]
  const float* bias;
  const void* residual;
  int relu, ln;
  const float *ln_w, *ln_b;
  void* out;
  int out_bf16;
  long long out_stride;
  float* dense;
  const int4* out_coors;
  int D, H, W;
};

template <int CIN, int COUT, bool SPARSE>
struct Cfg {
  static constexpr int CH = CIN / 8;  // 16-byte chunks per A row
  static constexpr int A_PAD = CH == 2 ? 64 : (CH == 4 ? 32 : 16);
  static constexpr int A_LBO = 128 * 16 + A_PAD;
  static constexpr int A_BYTES = CH * A_LBO;
  static constexpr int B_LBO = COUT * 16;
  static constexpr int B_BYTES = CH * B_LBO;
  // sparse conv with small channel counts: all 27 weight tiles stay resident in shared
  // memory for the CTA's lifetime instead of being re-streamed with every ring slot
  static constexpr bool WRES = SPARSE && (27 * B_BYTES <= 56 * 1024);
  static constexpr int W_BYTES = WRES ? 27 * B_BYTES : 0;
  static constexpr int STAGE_BYTES = (A_BYTES + (WRES ? 0 : B_BYTES) + 127) / 128 * 128;
  static constexpr int BUDGET = (STAGE_BYTES * 3 + W_BYTES > 100 * 1024) ? 200 * 1024 : 104 * 1024;
  static constexpr int STAGES_RAW = (BUDGET - W_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 12 ? 12 : (STAGES_RAW < 2 ? 2 : STAGES_RAW);
  static constexpr int TMEM_COLS = 2 * COUT < 32 ? 32 : 2 * COUT;
  static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
  static constexpr int SMEM_BYTES = W_BYTES + STAGES * STAGE_BYTES + BAR_BYTES + 128;
  // kind::f16: D fp32 (bit 4), A bf16 (bit 7), B bf16 (bit 10), K-major both, N>>3 @17, M>>4 @24
  static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(COUT >> 3) << 17) | ((128u >> 4) << 24);
};

__device__ __forceinline__ void cp_async16_ca(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

template <int CIN, int COUT, bool SPARSE>
__global__ void __launch_bounds__(288, (Cfg<CIN, COUT, SPARSE>::SMEM_BYTES <= 110 * 1024 && COUT <= 64) ? 2 : 1)
igemm_umma_kernel(const IgemmArgs a) {
  using C = Cfg<CIN, COUT, SPARSE>;
  constexpr int S = C::STAGES;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
  uint8_t* w_res = smem;
  uint8_t* stage_base = smem + C::W_BYTES;
  uint64_t* bars = (uint64_t*)(stage_base + S * C::STAGE_BYTES);
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * S + 4);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (S + s); };
  auto tfull_bar = [&](int b) { return bar0 + 8u * (2 * S + b); };
  auto tempty_bar = [&](int b) { return bar0 + 8u * (2 * S + 2 + b); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_rows = SPARSE ? (a.d_n_out ? min(*a.d_n_out, a.cap_out) : a.cap_out) : a.m_rows;
  const int m_tiles = (m_rows + 127) >> 7;
  const int total_tiles = m_tiles * a.n_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 128); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)C::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (C::WRES) {
    // resident weights: packed global image == shared image ([k][CIN/8][COUT][8] bf16)
    const uint4* wsrc = reinterpret_cast<const uint4*>(a.w);
    uint4* wdst = reinterpret_cast<uint4*>(w_res);
    const int n16 = a.kvol * (C::B_BYTES / 16);
    for (int j = threadIdx.x; j < n16; j += blockDim.x) wdst[j] = __ldg(wsrc + j);
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 4 && warp < 8) {
    // ------------------------------------------------------------------ producers
    const int pt = threadIdx.x - 128;
    int it = 0;
    // neighbour indices of this thread's row for the current tile (idx) and the next one
    // (nxt): 27 independent loads issued a whole tile ahead, so the ring never stalls on an
    // index -> address dependency.
    int idx[27], nxt[27];
    uint32_t mask = 0xffffffffu, mask_nxt = 0xffffffffu;
    auto load_idx = [&](int tile, int* dst, uint32_t& m) {
      m = 0xffffffffu;
      if (!SPARSE) return;
      const bool live = tile < total_tiles;
      const int mt = live ? tile % m_tiles : 0;
      const int row = mt * 128 + pt;
      if (a.tile_mask) m = live ? __ldg(a.tile_mask + mt) : 0u;
#pragma unroll
      for (int k = 0; k < 27; ++k) {
        dst[k] = -1;
        if (live && k < a.kvol && ((m >> k) & 1u) && row < m_rows) dst[k] = __ldg(a.nbr + (size_t)k * a.cap_out + row);
      }
    };
    load_idx(blockIdx.x, idx, mask);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int mt = tile % m_tiles, nt = tile / m_tiles;
      const int row = mt * 128 + pt;
      load_idx(tile + gridDim.x, nxt, mask_nxt);
      auto fill = [&](int k, int src_row) {
        const int s = it % S;
        const uint32_t ph = (uint32_t)(it / S) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        const __nv_bfloat16* src = src_row >= 0 ? a.in + (size_t)src_row * a.in_stride + (size_t)k * a.k_stride : a.in;
        const uint32_t nbytes = src_row >= 0 ? 16u : 0u;
        const uint32_t sa = smem_u32(stage_base + s * C::STAGE_BYTES);
#pragma unroll
        for (int c = 0; c < C::CH; ++c) cp_async16_ca(sa + c * C::A_LBO + pt * 16, src + c * 8, nbytes);
        if (!C::WRES) {
          const uint32_t sb = sa + C::A_BYTES;
          const __nv_bfloat16* wsrc = a.w + ((size_t)nt * a.kvol + k) * (size_t)(CIN * COUT);
#pragma unroll
          for (int j = pt; j < C::CH * COUT; j += 128) cp_async16(sb + j * 16, wsrc + (size_t)j * 8, 16u);
        }
        // the hardware arrives on full[s] for this thread when its copies have landed
        // (cutlass::arch::cpasync_barrier_arrive_noinc pattern): producers never wait on data
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(full_bar(s)) : "memory");
        ++it;
      };
      if (SPARSE) {
#pragma unroll
        for (int k = 0; k < 27; ++k) {
          if (k < a.kvol && ((mask >> k) & 1u)) fill(k, idx[k]);
        }
#pragma unroll
        for (int k = 0; k < 27; ++k) idx[k] = nxt[k];
        mask = mask_nxt;
      } else {
        const int src_row = row < m_rows ? row : -1;
        for (int k = 0; k < a.kvol; ++k) fill(k, src_row);
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (warp == 8) {
    // ------------------------------------------------------------------ MMA issuer
    int it = 0, tcount = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      const int mt = tile % m_tiles;
      const uint32_t mask = (SPARSE && a.tile_mask) ? __ldg(a.tile_mask + mt) : 0xffffffffu;
      const int buf = tcount & 1;
      const uint32_t tph = (uint32_t)(tcount >> 1) & 1u;
      mbar_wait(tempty_bar(buf), tph ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(buf * COUT);
      uint32_t accumulate = 0;
      for (int k = 0; k < a.kvol; ++k) {
        if (SPARSE && !((mask >> k) & 1u)) continue;
        const int s = it % S;
        const uint32_t ph = (uint32_t)(it / S) & 1u;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(stage_base + s * C::STAGE_BYTES);
          const uint32_t sb = C::WRES ? smem_u32(w_res) + (uint32_t)k * C::B_BYTES : sa + C::A_BYTES;
#pragma unroll
          for (int j = 0; j < CIN / 16; ++j) {
            uint64_t ad = make_desc(sa + j * 2 * C::A_LBO, C::A_LBO, 128);
            uint64_t bd = make_desc(sb + j * 2 * C::B_LBO, C::B_LBO, 128);
            tc_mma_bf16(tmem_d, ad, bd, C::IDESC, accumulate);
            accumulate = 1;
          }
          tc_commit(empty_bar(s));
        }
        __syncwarp();
        accumulate = 1;
        ++it;
      }
      if (lane == 0) tc_commit(tfull_bar(buf));
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    int tcount = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      const int mt = tile % m_tiles, nt = tile / m_tiles;
      const int buf = tcount & 1;
      const uint32_t tph = (uint32_t)(tcount >> 1) & 1u;
      mbar_wait(tfull_bar(buf), tph);
      tc_fence_after();
      const int row = mt * 128 + warp * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * COUT);
      float v[COUT];
#pragma unroll
      for (int c0 = 0; c0 < COUT; c0 += 16) tc_ld16(taddr + c0, v + c0);
      // accumulator is in registers: hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      mbar_arrive(tempty_bar(buf));
      if (row < m_rows) {
        const int col0 = nt * COUT;
        if (a.bias) {
#pragma unroll
          for (int c = 0; c < COUT; ++c) v[c] += __ldg(a.bias + col0 + c);
        }
        if (a.residual) {
          if (a.out_bf16) {
            const uint4* rp = (const uint4*)((const __nv_bfloat16*)a.residual + (size_t)row * a.out_stride + col0);
#pragma unroll
            for (int c = 0; c < COUT; c += 8) {
              uint4 u = __ldg(rp + c / 8);
              uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                v[c + 2 * q] += __uint_as_float(w4[q] << 16);
                v[c + 2 * q + 1] += __uint_as_float(w4[q] & 0xffff0000u);
              }
            }
          } else {
            const float4* rp = (const float4*)((const float*)a.residual + (size_t)row * a.out_stride + col0);
#pragma unroll
            for (int c = 0; c < COUT; c += 4) {
              float4 u = __ldg(rp + c / 4);
              v[c] += u.x; v[c + 1] += u.y; v[c + 2] += u.z; v[c + 3] += u.w;
            }
          }
        }
        if (a.ln) {
          float mean = 0.f;
#pragma unroll
          for (int c = 0; c < COUT; ++c) mean += v[c];
          mean *= (1.f / COUT);
          float var = 0.f;
#pragma unroll
          for (int c = 0; c < COUT; ++c) { float d = v[c] - mean; var += d * d; }
          const float rstd = rsqrtf(var * (1.f / COUT) + 1e-5f);
#pragma unroll
          for (int c = 0; c < COUT; ++c) v[c] = (v[c] - mean) * rstd * __ldg(a.ln_w + c) + __ldg(a.ln_b + c);
        }
        if (a.relu) {
#pragma unroll
          for (int c = 0; c < COUT; ++c) v[c] = fmaxf(v[c], 0.f);
        }
        if (a.dense) {
          const int4 q = __ldg(a.out_coors + row);
          const size_t hw = (size_t)a.H * a.W;
          float* dp = a.dense + ((size_t)q.x * COUT * a.D + q.y) * hw + (size_t)q.z * a.W + q.w;
#pragma unroll
          for (int c = 0; c < COUT; ++c) dp[(size_t)c * a.D * hw] = v[c];
        } else if (a.out_bf16) {
          uint4* op = (uint4*)((__nv_bfloat16*)a.out + (size_t)row * a.out_stride + col0);
#pragma unroll
          for (int c = 0; c < COUT; c += 8) {
            uint32_t w4[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              __nv_bfloat162 h = __floats2bfloat162_rn(v[c + 2 * q], v[c + 2 * q + 1]);
              w4[q] = *reinterpret_cast<uint32_t*>(&h);
            }
            op[c / 8] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
          }
        } else {
          float4* op = (float4*)((float*)a.out + (size_t)row * a.out_stride + col0);
#pragma unroll
          for (int c = 0; c < COUT; c += 4) op[c / 4] = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
  }
}

template <int CIN, int COUT, bool SPARSE>
static int launch_igemm(const IgemmArgs& a, int host_tiles, cudaStream_t st) {
  using C = Cfg<CIN, COUT, SPARSE>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(igemm_umma_kernel<CIN, COUT, SPARSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) { set_error("igemm<%d,%d>: cannot set %d B dynamic smem: %s", CIN, COUT, C::SMEM_BYTES, cudaGetErrorString(e)); return SRF_ERR_CUDA; }
    configured = true;
  }
  int per_sm = (227 * 1024) / (C::SMEM_BYTES + 1024);
  int by_tmem = 512 / C::TMEM_COLS;
  if (per_sm > by_tmem) per_sm = by_tmem;
  if (per_sm > 2) per_sm = 2;
  if (per_sm < 1) per_sm = 1;
  int grid = sm_count() * per_sm;
  if (grid > host_tiles) grid = host_tiles;
  if (grid < 1) grid = 1;
  SRF_COUNT(1);
  igemm_umma_kernel<CIN, COUT, SPARSE><<<grid, 288, C::SMEM_BYTES, st>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("igemm<%d,%d> launch failed: %s", CIN, COUT, cudaGetErrorString(e)); return SRF_ERR_CUDA; }
  return SRF_OK;
}

template <bool SPARSE>
static int dispatch_igemm(int cin, int cout, const IgemmArgs& a, int host_tiles, cudaStream_t st) {
#define SRF_CASE(ci, co) if (cin == ci && cout == co) return launch_igemm<ci, co, SPARSE>(a, host_tiles, st);
  SRF_CASE(16, 16) SRF_CASE(16, 32) SRF_CASE(32, 32) SRF_CASE(32, 64) SRF_CASE(64, 64) SRF_CASE(64, 128)
  SRF_CASE(128, 128) SRF_CASE(16, 128) SRF_CASE(32, 128) SRF_CASE(64, 32) SRF_CASE(128, 64) SRF_CASE(128, 32)
  SRF_CASE(64, 16) SRF_CASE(32, 16) SRF_CASE(16, 64) SRF_CASE(128, 16)
#undef SRF_CASE
  set_error("igemm: unsupported channel pair cin=%d cout=%d (each must be 16/32/64/128)", cin, cout);
  return SRF_ERR_UNSUPPORTED;
}

}  // namespace srf

using namespace srf;

extern "C" {

int srf_linear_tile_k(int32_t k);
int srf_linear_tile_n(int32_t n);

int srf_spconv_bf16(const srf_conv_args* c, void* stream) {
  SRF_CHECK_ARG(c && c->in && c->nbr && c->w && (c->out || c->dense), "srf_spconv_bf16: null arg");
  SRF_CHECK_ARG(c->in_dtype == SRF_BF16, "srf_spconv_bf16: input features must be bf16");
  SRF_CHECK_ARG(c->kvol >= 1 && c->kvol <= 27, "srf_spconv_bf16: kvol must be in [1,27]");
  SRF_CHECK_ARG(c->cap_out > 0 && c->cap_out % 128 == 0, "srf_spconv_bf16: cap_out must be a multiple of 128");
  SRF_CHECK_ARG(!c->dense || c->out_coors, "srf_spconv_bf16: dense output needs out_coors");
  IgemmArgs a = {};
  a.in = (const __nv_bfloat16*)c->in;
  a.in_stride = c->cin;
  a.k_stride = 0;
  a.nbr = c->nbr;
  a.tile_mask = c->tile_mask;
  a.d_n_out = c->d_n_out;
  a.cap_out = c->cap_out;
  a.kvol = c->kvol;
  a.n_tiles = 1;
  a.w = (const __nv_bfloat16*)c->w;
  a.bias = c->bias;
  a.residual = c->residual;
  a.relu = c->relu;
  a.out = c->out;
  a.out_bf16 = c->out_dtype == SRF_BF16;
  a.out_stride = c->cout;
  a.dense = c->dense;
  a.out_coors = (const int4*)c->out_coors;
  a.D = c->out_dims[1];
  a.H = c->out_dims[2];
  a.W = c->out_dims[3];
  return dispatch_igemm<true>(c->cin, c->cout, a, c->cap_out / 128, (cudaStream_t)stream);
}

int srf_linear_bf16(const void* a_bf16, int32_t m, int32_t k, const void* w_packed, int32_t n, const float* bias,
                    int32_t epi, const float* ln_w, const float* ln_b, void* out, int32_t out_dtype, void* stream) {
  SRF_CHECK_ARG(a_bf16 && w_packed && out && m >= 0 && k > 0 && n > 0, "srf_linear_bf16: bad args");
  if (m == 0) return SRF_OK;
  int tk = srf_linear_tile_k(k), tn = srf_linear_tile_n(n);
  SRF_CHECK_ARG(k % tk == 0 && n % tn == 0, "srf_linear_bf16: n=%d k=%d not tileable", n, k);
  SRF_CHECK_ARG(!(epi & 2) || (n == tn && ln_w && ln_b), "srf_linear_bf16: fused LayerNorm needs n <= 128 and ln weights");
  IgemmArgs a = {};
  a.in = (const __nv_bfloat16*)a_bf16;
  a.in_stride = k;
  a.k_stride = tk;
  a.m_rows = m;
  a.cap_out = m;
  a.kvol = k / tk;
  a.n_tiles = n / tn;
  a.w = (const __nv_bfloat16*)w_packed;
  a.bias = bias;
  a.relu = epi & 1;
  a.ln = (epi & 2) ? 1 : 0;
  a.ln_w = ln_w;
  a.ln_b = ln_b;
  a.out = out;
  a.out_bf16 = out_dtype == SRF_BF16;
  a.out_stride = n;
  return dispatch_igemm<false>(tk, tn, a, cdiv(m, 128) * (n / tn), (cudaStream_t)stream);
}

}  // extern "C"
