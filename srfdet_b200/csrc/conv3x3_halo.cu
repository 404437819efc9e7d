// Dense 3x3 / stride 1 / pad 1 convolution + folded BatchNorm2d + ReLU over NHWC pixel rows on tcgen05, with the
// im2col done by the shared-memory DESCRIPTOR instead of by gathers: the convolutions of SECONDCustom
// (/root/reference/mmdet3d_plugin/models/backbones/second_custom.py:23-91) and of the FPN's output convs
// (configs/nus/srfdet_voxel_nusc_L.py:55-75), SURVEY.md 8f rank 1.
//
// The gather-GEMM kernel (igemm_umma.cu) treats these layers as sparse convolutions over a static dense
// rulebook: every input pixel row is copied into shared memory 9 times (once per kernel offset, 16-byte
// cp.async pieces) and the 128 x 128 weight tile of every offset is streamed for every 128-row tile, which
// makes the 128 / 256-channel layers L2 -> SM bandwidth bound (profiles/r02_notes.md).  Here an output tile is
// 16 rows x 8 pixels of the image and its 18 x 10 input halo is staged ONCE, in K-chunk planes
//        plane c (8 channels = 16 bytes) : [18 halo rows][10 halo pixels][16 B]
// In the no-swizzle K-major UMMA layout a core matrix is 8 rows x 16 B = 128 contiguous bytes and the stride
// between 8-row groups (SBO) is free: with "row" = pixel x inside a tile row and "8-row group" = tile row y, the A
// operand of kernel offset (dy, dx) is the SAME halo read through a descriptor whose start address is moved by
// (dy * 10 + dx) * 16 bytes and whose SBO is one halo row (160 B).  Nine offsets, zero copies: A traffic drops
// from 9 to 1.4 rows per output row, the producers issue 6.4x fewer cp.async, zero padding comes from
// cp.async zero-fill of out-of-image halo pixels.
//
// Warp roles: 0-3 epilogue (tcgen05.ld, bias + ReLU, store in any encoding; tile i's epilogue overlaps tile i+1's
// MMAs through two TMEM buffers), 4-7 halo producers (cp.async, hardware arrive on the slot's mbarrier), 8 weight
// producer (one lane, cp.async.bulk of the packed 128 x 128 tile of (offset, K half)), 9 MMA issuer.
// Operands f16 / bf16 (run-time descriptor field), fp32 accumulate; weights in the packing of srf_pack_weight_tc.
#include <cuda.h>   // CUtensorMap + cuTensorMapEncodeTiled prototype (types only: the entry point is fetched through cudart, no libcuda link)
#include "igemm_common.cuh"

namespace srf {

#ifdef SRF_HALO_TRACE
// development build: globaltimer event trace of CTA 0 (tools/halo_trace.py): (code << 56) | ns
__device__ unsigned long long g_halo_trace[4096];
__device__ int g_halo_trace_n;
#define HTRACE(code_) do { if (blockIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_)); const int i_ = atomicAdd(&g_halo_trace_n, 1); if (i_ < 4096) g_halo_trace[i_] = ((unsigned long long)(code_) << 56) | (t_ & 0x00ffffffffffffffull); } } while (0)
#else
#define HTRACE(code_)
#endif

constexpr int HC_TY = 16, HC_TX = 8;                 // output tile: 16 image rows x 8 pixels = 128 GEMM rows
constexpr int HC_HY = HC_TY + 2, HC_HX = HC_TX + 2;  // halo
constexpr int HC_KC = 128;                           // input channels per A stage / weight slot
constexpr int HC_PLANE = HC_HY * HC_HX * 16 + 16;    // bytes per K-chunk plane (+16: the 16 planes land on distinct 16-byte bank groups)
constexpr int HC_A_BYTES = (HC_KC / 8) * HC_PLANE;   // 46336
constexpr int HC_W_BYTES = HC_KC * 128 * 2;          // 32768: [KC/8][128][8] 16-bit
constexpr int HC_NA = 2, HC_NW = 4;
constexpr int HC_THREADS = 448;                      // 4 + 4 epilogue warps (0-3, 10-13), 4 halo producers (4-7), weight producer (8), MMA (9)
constexpr int HC_SMEM = HC_NA * HC_A_BYTES + HC_NW * HC_W_BYTES + (2 * HC_NA + 2 * HC_NW + 4) * 8 + 16 + 2 * 128 * 4 + 128;   // + bias of two tiles in flight

struct HaloArgs {
  alignas(64) CUtensorMap tmap;   // 5-D view of the input rows: (c % 8, x, y, c / 8, image); box = one K-chunk plane of a halo
  IgemmArgs e;          // epilogue fields (bias, relu, out, out_enc, out_stride, out_lo_off, fmt)
  int n, h, w, cin;     // image batch / size, input channels (multiple of 128)
  int n_tiles;          // cout / 128
  int use_tma;          // halo planes by TMA tensor copies (OOB zero-fill = the conv's zero padding); 0: cp.async gathers
};

constexpr int HC_PLANE_TMA = HC_HY * HC_HX * 16;   // the tensor copy writes the 16 planes of a stage densely (2880-byte planes)

// the 16 K-chunk planes (128 channels) of an 18 x 10 halo in one tensor copy: [16][18][10][16 B], zero-filled outside the image
__device__ __forceinline__ void tma_halo(uint32_t dst, const CUtensorMap* tm, int x, int y, int plane, int img, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(0), "r"(x), "r"(y), "r"(plane), "r"(img), "r"(bar) : "memory");
}

__global__ void __launch_bounds__(HC_THREADS, 1) conv3x3_halo_kernel(const __grid_constant__ HaloArgs p) {
  extern __shared__ __align__(128) uint8_t hc_smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)hc_smem_raw + 127) & ~(uintptr_t)127);
  const uint32_t a0 = smem_u32(smem), w0 = a0 + HC_NA * HC_A_BYTES;
  uint64_t* bars = (uint64_t*)(smem + HC_NA * HC_A_BYTES + HC_NW * HC_W_BYTES);
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * HC_NA + 2 * HC_NW + 4);
  float* bias_s = reinterpret_cast<float*>(tmem_slot + 4);          // [2][128]: bias of the column tile, per TMEM buffer
  const uint32_t bar0 = smem_u32(bars);
  auto afull = [&](int s) { return bar0 + 8u * s; };
  auto aempty = [&](int s) { return bar0 + 8u * (HC_NA + s); };
  auto wfull = [&](int s) { return bar0 + 8u * (2 * HC_NA + s); };
  auto wempty = [&](int s) { return bar0 + 8u * (2 * HC_NA + HC_NW + s); };
  auto tfull = [&](int b) { return bar0 + 8u * (2 * HC_NA + 2 * HC_NW + b); };
  auto tempty = [&](int b) { return bar0 + 8u * (2 * HC_NA + 2 * HC_NW + 2 + b); };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const IgemmArgs& a = p.e;

  if (threadIdx.x == 0) HTRACE(1);
  pdl_trigger();
  if (threadIdx.x == 0) {
    for (int s = 0; s < HC_NA; ++s) { mbar_init(afull(s), p.use_tma ? 1 : 128); mbar_init(aempty(s), 1); }
    for (int s = 0; s < HC_NW; ++s) { mbar_init(wfull(s), 1); mbar_init(wempty(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull(b), 1); mbar_init(tempty(b), 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __reduce_or_sync(0xffffffffu, *tmem_slot);
  pdl_wait();
  if (threadIdx.x == 0) HTRACE(2);

  const int ty_n = (p.h + HC_TY - 1) / HC_TY, tx_n = (p.w + HC_TX - 1) / HC_TX;
  const int kq = p.cin / HC_KC;
  const int total = p.n * ty_n * tx_n * p.n_tiles;
  // tile -> (column tile, image, tile row, tile column); the column tiles of one spatial tile are adjacent (same halo: L2 hits)
  auto decode = [&](int tile, int& nt, int& img, int& y0, int& x0) {
    nt = tile % p.n_tiles;
    int t = tile / p.n_tiles;
    x0 = (t % tx_n) * HC_TX;
    t /= tx_n;
    y0 = (t % ty_n) * HC_TY;
    img = t / ty_n;
  };

  if (warp < 4 || warp >= 10) {
    // ------------------------------------------------------------------ epilogue (8 warps)
    // A warp may only read the TMEM lanes of its quarter (warp % 4): warps 0-3 drain columns 0-63 of their 32 rows, warps
    // 10-13 columns 64-127.  The tile's 128 bias values are staged in shared memory once (a per-thread broadcast load per
    // column made the epilogue, not the MMAs, the per-tile cost); 32 columns per tcgen05.ld, two loads in flight per wait.
    int tcount = 0;
    const int quarter = warp & 3, half = warp >= 10 ? 1 : 0;
    const int r = quarter * 32 + lane, py = r >> 3, px = r & 7;
    const int et = (warp < 4 ? warp : warp - 6) * 32 + lane;       // 0..255 among the epilogue threads
    const bool f16 = a.fmt != 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++tcount) {
      int nt, img, y0, x0;
      decode(tile, nt, img, y0, x0);
      const int buf = tcount & 1;
      // bias of this column tile -> smem (buffer `buf` was last read two tiles ago: every reader has since passed a bar.sync)
      if (et < 128) bias_s[buf * 128 + et] = a.bias ? __ldg(a.bias + nt * 128 + et) : 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(tfull(buf), (uint32_t)(tcount >> 1) & 1u);
      if (et == 0) HTRACE(10);
      tc_fence_after();
      const int y = y0 + py, x = x0 + px;
      const bool ok = y < p.h && x < p.w;
      const size_t row = (size_t)(img * p.h + y) * p.w + x;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * 128 + half * 64);
      uint32_t v[64];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                     "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                     "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                   : "r"(taddr) : "memory");
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]),
                     "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]),
                     "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
                   : "r"(taddr + 32u) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      // the accumulator half is in registers: hand the TMEM buffer back before the stores
      tc_fence_before();
      mbar_arrive(tempty(buf));
      if (ok) {
        const float* bs = bias_s + buf * 128 + half * 64;
        const int col0 = nt * 128 + half * 64;
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 8) {
          const float4 b0 = *reinterpret_cast<const float4*>(bs + c0), b1 = *reinterpret_cast<const float4*>(bs + c0 + 4);
          float o[8] = {__uint_as_float(v[c0]) + b0.x, __uint_as_float(v[c0 + 1]) + b0.y, __uint_as_float(v[c0 + 2]) + b0.z, __uint_as_float(v[c0 + 3]) + b0.w,
                        __uint_as_float(v[c0 + 4]) + b1.x, __uint_as_float(v[c0 + 5]) + b1.y, __uint_as_float(v[c0 + 6]) + b1.z, __uint_as_float(v[c0 + 7]) + b1.w};
          if (a.relu) {
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = fmaxf(o[i], 0.f);
          }
          store_cols_to<8>(a.out, a.out_enc, a.out_stride, a.out_lo_off, f16, row, col0 + c0, o);
        }
      }
      if (et == 0) HTRACE(11);
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------------ halo producers
    const int pt = threadIdx.x - 128;
    if (p.use_tma) {
      // TMA form: one thread issues ONE tensor copy per stage (box 8 ch x 10 px x 18 rows x 16 planes at (x0 - 1, y0 - 1));
      // the copy engine zero-fills what lies outside the image, which is the convolution's padding.
      if (pt == 0) {
        int it = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
          int nt, img, y0, x0;
          decode(tile, nt, img, y0, x0);
          for (int q = 0; q < kq; ++q, ++it) {
            const int s = it % HC_NA;
            mbar_wait(aempty(s), (((uint32_t)(it / HC_NA)) & 1u) ^ 1u);
            HTRACE(20);
            const uint32_t fb = afull(s);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"((uint32_t)((HC_KC / 8) * HC_HY * HC_HX * 16)) : "memory");
            tma_halo(a0 + (uint32_t)s * HC_A_BYTES, &p.tmap, x0 - 1, y0 - 1, q * (HC_KC / 8), img, fb);
            HTRACE(21);
          }
        }
      }
    } else {
    const int c = pt & 15;                          // K-chunk plane of this thread (consecutive lanes: one pixel's 256 contiguous bytes)
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      int nt, img, y0, x0;
      decode(tile, nt, img, y0, x0);
      for (int q = 0; q < kq; ++q, ++it) {
        const int s = it % HC_NA;
        mbar_wait(aempty(s), (((uint32_t)(it / HC_NA)) & 1u) ^ 1u);
        if (pt == 0) HTRACE(20);
        const uint32_t dst = a0 + (uint32_t)s * HC_A_BYTES + (uint32_t)c * HC_PLANE;
        const uint16_t* src_c = a.in + (size_t)q * HC_KC + c * 8;
#pragma unroll 4
        for (int pix = pt >> 4; pix < HC_HY * HC_HX; pix += 8) {
          const int hy = pix / HC_HX, hx = pix - hy * HC_HX;
          const int y = y0 - 1 + hy, x = x0 - 1 + hx;
          const bool ok = y >= 0 && y < p.h && x >= 0 && x < p.w;
          const uint16_t* src = ok ? src_c + ((size_t)(img * p.h + y) * p.w + x) * a.in_stride : a.in;
          cp_async16(dst + (uint32_t)pix * 16u, src, ok ? 16u : 0u);
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(afull(s)) : "memory");
        if (pt == 0) HTRACE(21);
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------------ weight producer
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int nt = tile % p.n_tiles;
        for (int q = 0; q < kq; ++q)
          for (int i = 0; i < 9; ++i, ++it) {
            const int k = i;
            const int s = it % HC_NW;
            mbar_wait(wempty(s), (((uint32_t)(it / HC_NW)) & 1u) ^ 1u);
            HTRACE(30);
            const uint32_t fb = wfull(s);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"((uint32_t)HC_W_BYTES) : "memory");
            bulk_g2s(w0 + (uint32_t)s * HC_W_BYTES, a.w + ((size_t)(nt * 9 + k) * kq + q) * (size_t)(HC_W_BYTES / 2), HC_W_BYTES, fb);
          }
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer
    // kind::f16, D fp32, A/B format from a.fmt, K-major both, N = 128, M = 128
    const uint32_t idesc = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24) | (a.fmt ? 0u : ((1u << 7) | (1u << 10)));
    constexpr uint32_t A_HI = ((HC_HX * 16u) >> 4) | (1u << 14);      // SBO = one halo row
    const uint32_t plane = p.use_tma ? (uint32_t)HC_PLANE_TMA : (uint32_t)HC_PLANE;   // LBO: bytes between K-chunk planes
    int ia = 0, iw = 0, tcount = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++tcount) {
      const int buf = tcount & 1;
      mbar_wait(tempty(buf), ((uint32_t)(tcount >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(buf * 128);
      uint32_t accumulate = 0;
      for (int q = 0; q < kq; ++q, ++ia) {
        const int sa = ia % HC_NA;
        mbar_wait(afull(sa), ((uint32_t)(ia / HC_NA)) & 1u);
        if (lane == 0) HTRACE(40);
        const uint32_t abase = a0 + (uint32_t)sa * HC_A_BYTES;
#pragma unroll 1
        for (int i = 0; i < 9; ++i, ++iw) {
          const int k = i;
          const int sw = iw % HC_NW;
          mbar_wait(wfull(sw), ((uint32_t)(iw / HC_NW)) & 1u);
          if (lane == 0) HTRACE(41);
          tc_fence_after();
          const int dy = k / 3, dx = k - dy * 3;
          const uint32_t a_lo0 = (((abase + (uint32_t)(dy * HC_HX + dx) * 16u) >> 4) & 0x3fffu) | ((plane >> 4) << 16);
          const uint32_t b_lo0 = (((w0 + (uint32_t)sw * HC_W_BYTES) >> 4) & 0x3fffu) | ((uint32_t)((128 * 16) >> 4) << 16);
          if (elect_one_sync()) {
#pragma unroll
            for (int j = 0; j < HC_KC / 16; ++j) {
              const uint64_t ad = desc_pack(a_lo0 + (uint32_t)j * ((2u * plane) >> 4), A_HI);
              const uint64_t bd = desc_pack(b_lo0 + (uint32_t)(j * ((2 * 128 * 16) >> 4)), DESC_HI);
              tc_mma_f16(tmem_d, ad, bd, idesc, accumulate);
              accumulate = 1;
            }
            tc_commit(wempty(sw));
            if (i == 8) tc_commit(aempty(sa));
          }
          __syncwarp();
          accumulate = 1;
        }
      }
      if (elect_one_sync()) tc_commit(tfull(buf));
      __syncwarp();
      if (lane == 0) HTRACE(42);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) HTRACE(3);
  if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
}

}  // namespace srf

using namespace srf;

// cuTensorMapEncodeTiled through the runtime's driver entry point table (the library links only cudart)
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_tiled_fn encode_tiled() {
  static encode_tiled_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* e = getenv("SRF_HALO_TMA");
    if (!(e && e[0] == '0')) {
      void* ptr = nullptr;
      cudaDriverEntryPointQueryResult q;
      if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
        fn = (encode_tiled_fn)ptr;
      (void)cudaGetLastError();
    }
  }
  return fn;
}

// (n, h, w, cin) 16-bit NHWC rows as the 5-D tensor (c % 8, x, y, c / 8, image); a box is one K-chunk plane of an 18 x 10 halo
static bool make_halo_tmap(CUtensorMap* tm, const void* in, int n, int h, int w, int cin) {
  encode_tiled_fn enc = encode_tiled();
  if (!enc) return false;
  const cuuint64_t dims[5] = {8, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)(cin / 8), (cuuint64_t)n};
  const cuuint64_t strides[4] = {(cuuint64_t)cin * 2, (cuuint64_t)w * cin * 2, 16, (cuuint64_t)h * w * cin * 2};   // bytes, dims 1..4
  const cuuint32_t box[5] = {8, (cuuint32_t)HC_HX, (cuuint32_t)HC_HY, (cuuint32_t)(HC_KC / 8), 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 5, const_cast<void*>(in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int g_last_tma = -1;
extern "C" int srf_conv3x3_last_used_tma(void) { return g_last_tma; }

extern "C" int srf_conv3x3_rows(const void* in, int32_t enc, int32_t n, int32_t h, int32_t w, int32_t cin, const void* w_packed,
                                int32_t cout, const float* bias, int32_t relu, void* out, int32_t out_enc, void* stream) {
  SRF_CHECK_ARG(in && w_packed && out && n >= 1 && h >= 1 && w >= 1, "srf_conv3x3_rows: bad args");
  if (!(enc == SRF_F16 || enc == SRF_BF16) || cin % HC_KC != 0 || cin > 512 || cout % 128 != 0 ||
      !(out_enc == SRF_F32 || (enc_is_16(out_enc) && enc_is_f16(out_enc) == enc_is_f16(enc))) ||
      (long long)n * h * w >= (1ll << 31) || (((uintptr_t)in | (uintptr_t)w_packed | (uintptr_t)out) & 15) != 0) {
    set_error("srf_conv3x3_rows: plain f16 / bf16 rows, cin a multiple of 128 (<= 512), cout a multiple of 128");
    return SRF_ERR_UNSUPPORTED;
  }
  static bool configured = false;
  if (!configured) {
    SRF_CUDA(cudaFuncSetAttribute(conv3x3_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HC_SMEM));
    configured = true;
  }
  HaloArgs p = {};
  p.e.in = (const uint16_t*)in;
  p.e.in_stride = cin;
  p.e.w = (const uint16_t*)w_packed;
  p.e.bias = bias;
  p.e.relu = relu;
  p.e.out = out;
  p.e.fmt = enc_is_f16(enc) ? 1 : 0;
  p.e.out_enc = out_enc;
  p.e.out_stride = enc_is_split(out_enc) ? 2 * cout : cout;
  p.e.out_lo_off = cout;
  p.n = n; p.h = h; p.w = w; p.cin = cin;
  p.n_tiles = cout / 128;
  p.use_tma = make_halo_tmap(&p.tmap, in, n, h, w, cin) ? 1 : 0;
  g_last_tma = p.use_tma;
  const int tiles = n * cdiv(h, HC_TY) * cdiv(w, HC_TX) * p.n_tiles;
  int grid = sm_count();
  if (grid > tiles) grid = tiles;
  SRF_COUNT(1);
  cudaError_t e = launch_pdl(conv3x3_halo_kernel, dim3(grid), dim3(HC_THREADS), (size_t)HC_SMEM, (cudaStream_t)stream, p);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("conv3x3_halo launch failed: %s", cudaGetErrorString(e)); return SRF_ERR_CUDA; }
  return SRF_OK;
}

#ifdef SRF_HALO_TRACE
extern "C" int srf_halo_trace_read(unsigned long long* host, int reset) {     // host[4096]; returns the number of events
  cudaDeviceSynchronize();
  int n = 0;
  cudaMemcpyFromSymbol(&n, srf::g_halo_trace_n, 4);
  cudaMemcpyFromSymbol(host, srf::g_halo_trace, sizeof(srf::g_halo_trace));
  if (reset) { int z = 0; cudaMemcpyToSymbol(srf::g_halo_trace_n, &z, 4); }
  return n < 4096 ? n : 4096;
}
#endif
