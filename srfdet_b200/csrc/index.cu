// Cell index (bitmap + rank), device-wide scan, rulebook construction.
//
// Why a bitmap and not a hash + sort: every consumer on this path wants the occupied
// cells enumerated in ascending linear (b,z,y,x) order -- mmcv's DynamicScatter returns
// at::unique_dim order, and the canonical rulebook order is ascending linear index.  A
// bitmap over the dense grid (<= 92.4 M cells -> 11.6 MB, L2 resident on B200) plus a
// popcount scan gives that order AND a perfect hash (rank) with two loads per probe,
// no collisions, no sort passes.
#include <stdarg.h>

#include "common.cuh"

namespace srf {

static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached = 0;
  if (!cached) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) cached = 148;
  }
  return cached;
}

// --------------------------------------------------------------------------------------
// scan
// --------------------------------------------------------------------------------------
template <bool POPC>
__device__ __forceinline__ uint32_t scan_val(uint32_t v) {
  return POPC ? (uint32_t)__popc(v) : v;
}

__device__ __forceinline__ uint32_t block_reduce_sum(uint32_t v, uint32_t* sh) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  uint32_t t = 0;
  if (w == 0) {
    t = l < (blockDim.x >> 5) ? sh[l] : 0;
    for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (l == 0) sh[0] = t;
  }
  __syncthreads();
  t = sh[0];
  __syncthreads();
  return t;
}

// exclusive scan of one value per thread over the block; returns prefix, *total = block sum
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* sh, uint32_t* total) {
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  uint32_t inc = v;
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (l >= o) inc += t;
  }
  if (l == 31) sh[w] = inc;
  __syncthreads();
  if (w == 0) {
    int nw = blockDim.x >> 5;
    uint32_t s = l < nw ? sh[l] : 0;
    uint32_t si = s;
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, si, o);
      if (l >= o) si += t;
    }
    if (l < nw) sh[l] = si - s;  // exclusive warp offsets
    if (l == 31) sh[32] = si;    // block total
  }
  __syncthreads();
  uint32_t res = inc - v + sh[w];
  *total = sh[32];
  __syncthreads();
  return res;
}

template <bool POPC>
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const uint32_t* __restrict__ in,
                                                                  uint32_t* __restrict__ blocksum,
                                                                  int64_t n, int64_t chunk) {
  __shared__ uint32_t sh[33];
  int64_t beg = (int64_t)blockIdx.x * chunk;
  int64_t end = beg + chunk < n ? beg + chunk : n;
  uint32_t acc = 0;
  for (int64_t i = beg + threadIdx.x; i < end; i += SCAN_THREADS) acc += scan_val<POPC>(in[i]);
  uint32_t t = block_reduce_sum(acc, sh);
  if (threadIdx.x == 0) blocksum[blockIdx.x] = t;
}

__global__ void __launch_bounds__(SCAN_BLOCKS) scan_blocksums_kernel(uint32_t* blocksum, int nb,
                                                                    int32_t* d_total) {
  __shared__ uint32_t sh[33];
  uint32_t v = (int)threadIdx.x < nb ? blocksum[threadIdx.x] : 0;
  uint32_t total;
  uint32_t p = block_excl_scan(v, sh, &total);
  if ((int)threadIdx.x < nb) blocksum[threadIdx.x] = p;
  if (threadIdx.x == 0 && d_total) *d_total = (int32_t)total;
}

template <bool POPC>
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const uint32_t* __restrict__ in,
                                                                 uint32_t* __restrict__ out,
                                                                 const uint32_t* __restrict__ blocksum,
                                                                 int64_t n, int64_t chunk) {
  __shared__ uint32_t sh[33];
  int64_t beg = (int64_t)blockIdx.x * chunk;
  int64_t end = beg + chunk < n ? beg + chunk : n;
  uint32_t carry = blocksum[blockIdx.x];
  for (int64_t base = beg; base < end; base += SCAN_THREADS) {
    int64_t i = base + threadIdx.x;
    uint32_t v = i < end ? scan_val<POPC>(in[i]) : 0;
    uint32_t total;
    uint32_t p = block_excl_scan(v, sh, &total);
    if (i < end) out[i] = carry + p;
    carry += total;
  }
}

int scan_flags_launch(const uint32_t* in, uint32_t* out_excl, uint32_t* blocksum, int64_t n,
                      int32_t* d_total, int popcount_mode, cudaStream_t st) {
  if (n <= 0) {
    if (d_total) SRF_CUDA(cudaMemsetAsync(d_total, 0, sizeof(int32_t), st));
    return SRF_OK;
  }
  int nb = (int)((n + 4 * SCAN_THREADS - 1) / (4 * SCAN_THREADS));
  if (nb > SCAN_BLOCKS) nb = SCAN_BLOCKS;
  if (nb < 1) nb = 1;
  int64_t chunk = (n + nb - 1) / nb;
  chunk = (chunk + SCAN_THREADS - 1) / SCAN_THREADS * SCAN_THREADS;
  nb = (int)((n + chunk - 1) / chunk);
  if (popcount_mode) {
    SRF_COUNT(3);
    scan_reduce_kernel<true><<<nb, SCAN_THREADS, 0, st>>>(in, blocksum, n, chunk);
    scan_blocksums_kernel<<<1, SCAN_BLOCKS, 0, st>>>(blocksum, nb, d_total);
    scan_apply_kernel<true><<<nb, SCAN_THREADS, 0, st>>>(in, out_excl, blocksum, n, chunk);
  } else {
    SRF_COUNT(3);
    scan_reduce_kernel<false><<<nb, SCAN_THREADS, 0, st>>>(in, blocksum, n, chunk);
    scan_blocksums_kernel<<<1, SCAN_BLOCKS, 0, st>>>(blocksum, nb, d_total);
    scan_apply_kernel<false><<<nb, SCAN_THREADS, 0, st>>>(in, out_excl, blocksum, n, chunk);
  }
  SRF_LAUNCH_CHECK();
  return SRF_OK;
}

// --------------------------------------------------------------------------------------
// index kernels
// --------------------------------------------------------------------------------------
__global__ void index_mark_kernel(uint32_t* __restrict__ bits, Dims4 d, const int4* __restrict__ coors,
                                  int n, const int32_t* __restrict__ d_n) {
  int nn = d_n ? min(*d_n, n) : n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += gridDim.x * blockDim.x) {
    int4 q = __ldg(coors + i);
    if (q.x < 0 || q.y < 0 || q.z < 0 || q.w < 0) continue;
    if (q.x >= d.b || q.y >= d.z || q.z >= d.y || q.w >= d.x) continue;
    int64_t cell = cell_of(d, q.x, q.y, q.z, q.w);
    atomicOr(bits + (cell >> 5), 1u << (cell & 31));
  }
}

struct Conv3 {
  int32_t k[3], s[3], p[3];
};

// POW2: every stride is a power of two (every reference config: 1 or 2) -> the divisibility test and the division of the
// 27 candidate offsets are a mask and a shift (runtime divisors made this kernel instruction bound: 18-24 us per level),
// and the grid has < 2^31 cells -> 32-bit cell arithmetic.
template <bool POW2>
__global__ void index_mark_strided_kernel(uint32_t* __restrict__ bits, Dims4 od,
                                          const int4* __restrict__ in_coors, int cap_in,
                                          const int32_t* __restrict__ d_n_in, Conv3 cv, int sh0, int sh1, int sh2) {
  int nn = d_n_in ? min(*d_n_in, cap_in) : cap_in;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += gridDim.x * blockDim.x) {
    int4 q = __ldg(in_coors + i);
    if (q.x < 0 || q.y < 0) continue;
    for (int kz = 0; kz < cv.k[0]; ++kz) {
      int nz = q.y + cv.p[0] - kz;
      if (nz < 0 || (POW2 ? (nz & (cv.s[0] - 1)) : (nz % cv.s[0]))) continue;
      int z = POW2 ? (nz >> sh0) : nz / cv.s[0];
      if (z >= od.z) continue;
      for (int ky = 0; ky < cv.k[1]; ++ky) {
        int ny = q.z + cv.p[1] - ky;
        if (ny < 0 || (POW2 ? (ny & (cv.s[1] - 1)) : (ny % cv.s[1]))) continue;
        int y = POW2 ? (ny >> sh1) : ny / cv.s[1];
        if (y >= od.y) continue;
        for (int kx = 0; kx < cv.k[2]; ++kx) {
          int nx = q.w + cv.p[2] - kx;
          if (nx < 0 || (POW2 ? (nx & (cv.s[2] - 1)) : (nx % cv.s[2]))) continue;
          int x = POW2 ? (nx >> sh2) : nx / cv.s[2];
          if (x >= od.x) continue;
          if (POW2) {
            const uint32_t cell = (((uint32_t)q.x * (uint32_t)od.z + (uint32_t)z) * (uint32_t)od.y + (uint32_t)y) * (uint32_t)od.x + (uint32_t)x;
            atomicOr(bits + (cell >> 5), 1u << (cell & 31));
          } else {
            int64_t cell = cell_of(od, q.x, z, y, x);
            atomicOr(bits + (cell >> 5), 1u << (cell & 31));
          }
        }
      }
    }
  }
}

// coordinates of the occupied cells in rank order.  A word's 32 cells are consecutive along x: the word's first cell is
// decoded once (32-bit divisions when the grid has < 2^31 cells, which every config's has), each set bit then only steps
// x and carries into y / z / b when it runs off the line (64-bit divisions per voxel made this kernel 15-30 us on the
// coarse levels, whose words hold many voxels each).
template <typename I>
__global__ void index_emit_kernel(const uint32_t* __restrict__ bits, const uint32_t* __restrict__ rank,
                                  int64_t nwords, Dims4 d, int4* __restrict__ out, int cap) {
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords;
       w += (int64_t)gridDim.x * blockDim.x) {
    uint32_t word = __ldg(bits + w);
    if (!word) continue;
    int r = (int)__ldg(rank + w);
    I cell = (I)(w << 5);
    const int x0 = (int)(cell % (I)d.x); cell /= (I)d.x;
    const int y0 = (int)(cell % (I)d.y); cell /= (I)d.y;
    const int z0 = (int)(cell % (I)d.z);
    const int b0 = (int)(cell / (I)d.z);
    while (word) {
      const int b = __ffs(word) - 1;
      word &= word - 1;
      int4 q = make_int4(b0, z0, y0, x0 + b);
      while (q.w >= d.x) {
        q.w -= d.x;
        if (++q.z == d.y) { q.z = 0; if (++q.y == d.z) { q.y = 0; ++q.x; } }
      }
      if (r < cap) out[r] = q;
      ++r;
    }
  }
}

template <bool PERM>
__global__ void index_lookup_kernel(const uint32_t* __restrict__ bits, const uint32_t* __restrict__ rank,
                                    Dims4 d, const int4* __restrict__ coors, int n,
                                    const int32_t* __restrict__ d_n, int32_t* __restrict__ out) {
  int nn = d_n ? min(*d_n, n) : n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += gridDim.x * blockDim.x) {
    int4 q = __ldg(coors + i);
    int r = -1;
    if (q.x >= 0 && q.y >= 0 && q.z >= 0 && q.w >= 0 && q.x < d.b && q.y < d.z && q.z < d.y && q.w < d.x)
      r = index_rank(bits, rank, cell_of(d, q.x, q.y, q.z, q.w));
    if (PERM) {
      if (r >= 0) out[r] = i;
    } else {
      out[i] = r;
    }
  }
}

// --------------------------------------------------------------------------------------
// rulebook: one thread per output row, lanes = consecutive (sorted) rows so that the
// bitmap / rank words a warp probes for a given offset are shared; the per-tile offset
// mask is aggregated with a warp ballot (one atomicOr per warp and offset).
// --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rulebook_kernel(const uint32_t* __restrict__ bits,
                                                      const uint32_t* __restrict__ rank, Dims4 id,
                                                      const int32_t* __restrict__ in_perm,
                                                      const int4* __restrict__ out_coors, int cap_out,
                                                      const int32_t* __restrict__ d_n_out, Conv3 cv,
                                                      int32_t* __restrict__ nbr,
                                                      uint32_t* __restrict__ tile_mask) {
  int n_out = d_n_out ? min(*d_n_out, cap_out) : cap_out;
  int n_pad = min((n_out + 127) / 128 * 128, cap_out);
  for (int base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; base < n_pad;
       base += gridDim.x * blockDim.x) {
    int o = base + (threadIdx.x & 31);
    bool live = o < n_out;
    int4 q = live ? __ldg(out_coors + o) : make_int4(0, 0, 0, 0);
    int kk = 0;
    for (int kz = 0; kz < cv.k[0]; ++kz) {
      int z = q.y * cv.s[0] - cv.p[0] + kz;
      for (int ky = 0; ky < cv.k[1]; ++ky) {
        int y = q.z * cv.s[1] - cv.p[1] + ky;
        for (int kx = 0; kx < cv.k[2]; ++kx, ++kk) {
          int x = q.w * cv.s[2] - cv.p[2] + kx;
          int r = -1;
          if (live && z >= 0 && z < id.z && y >= 0 && y < id.y && x >= 0 && x < id.x) {
            r = index_rank(bits, rank, cell_of(id, q.x, z, y, x));
            if (r >= 0 && in_perm) r = __ldg(in_perm + r);
          }
          if (o < n_pad) nbr[(size_t)kk * cap_out + o] = r;
          unsigned any = __ballot_sync(0xffffffffu, r >= 0);
          if (any && (threadIdx.x & 31) == 0) atomicOr(tile_mask + (base >> 7), 1u << kk);
        }
      }
    }
  }
}

// Unrolled form for the kernel shapes the encoders use ((3,3,3) and the (3,1,1) conv_out): the K2 cells of a (kz, ky) line
// are consecutive along x, so they share one bitmap / rank word (two when the run crosses a word boundary): 9-12 word
// loads per row instead of 27 probes; every probe of a row is resolved into registers before the first store (all loads
// independent), 32-bit cell arithmetic (I) when the grid has < 2^31 cells, one tile-mask atomic per warp.
template <int K0, int K1, int K2, typename I>
__global__ void __launch_bounds__(256) rulebook_fast_kernel(const uint32_t* __restrict__ bits, const uint32_t* __restrict__ rank, Dims4 id,
                                                           const int32_t* __restrict__ in_perm, const int4* __restrict__ out_coors,
                                                           int cap_out, const int32_t* __restrict__ d_n_out, Conv3 cv,
                                                           int32_t* __restrict__ nbr, uint32_t* __restrict__ tile_mask) {
  const int n_out = d_n_out ? min(*d_n_out, cap_out) : cap_out;
  const int n_pad = min((n_out + 127) / 128 * 128, cap_out);
  for (int base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; base < n_pad; base += gridDim.x * blockDim.x) {
    const int o = base + (threadIdx.x & 31);
    const bool live = o < n_out;
    const int4 q = live ? __ldg(out_coors + o) : make_int4(0, 0, 0, 0);
    int rr[K0 * K1 * K2];
    const int xs = q.w * cv.s[2] - cv.p[2];
#pragma unroll
    for (int kz = 0; kz < K0; ++kz) {
      const int z = q.y * cv.s[0] - cv.p[0] + kz;
#pragma unroll
      for (int ky = 0; ky < K1; ++ky) {
        const int y = q.z * cv.s[1] - cv.p[1] + ky;
        const bool ok = live && z >= 0 && z < id.z && y >= 0 && y < id.y;
        const I line = (((I)q.x * (I)id.z + (I)(ok ? z : 0)) * (I)id.y + (I)(ok ? y : 0)) * (I)id.x;
        // words of the first and of the last in-range cell of the run
        const int xa = max(xs, 0), xb = min(xs + K2 - 1, id.x - 1);
        const bool any = ok && xa <= xb;
        const I wa = (line + (I)(any ? xa : 0)) >> 5, wb = (line + (I)(any ? xb : 0)) >> 5;
        const uint32_t ba = any ? __ldg(bits + wa) : 0u;
        const uint32_t bb = (any && wb != wa) ? __ldg(bits + wb) : ba;
        const uint32_t ra = ba ? __ldg(rank + wa) : 0u;
        const uint32_t rb = (wb != wa && bb) ? __ldg(rank + wb) : ra;
#pragma unroll
        for (int kx = 0; kx < K2; ++kx) {
          const int x = xs + kx;
          int r = -1;
          if (any && x >= 0 && x < id.x) {
            const I cell = line + (I)x;
            const bool first = (cell >> 5) == wa;
            const uint32_t word = first ? ba : bb, bit = 1u << (uint32_t)(cell & 31);
            if (word & bit) r = (int)((first ? ra : rb) + __popc(word & (bit - 1u)));
          }
          rr[(kz * K1 + ky) * K2 + kx] = r;
        }
      }
    }
    if (in_perm) {
#pragma unroll
      for (int kk = 0; kk < K0 * K1 * K2; ++kk)
        if (rr[kk] >= 0) rr[kk] = __ldg(in_perm + rr[kk]);
    }
    uint32_t mask = 0u;
#pragma unroll
    for (int kk = 0; kk < K0 * K1 * K2; ++kk) {
      nbr[(size_t)kk * cap_out + o] = rr[kk];              // o < n_pad: base is a multiple of 32 below n_pad (a multiple of 128)
      if (__ballot_sync(0xffffffffu, rr[kk] >= 0)) mask |= 1u << kk;
    }
    if (mask && (threadIdx.x & 31) == 0) atomicOr(tile_mask + (base >> 7), mask);
  }
}

__global__ void dense_rulebook_kernel(int n, int h, int w, int ho, int wo, int ks, int stride, int pad, int cap_out,
                                      int32_t* __restrict__ nbr, uint32_t* __restrict__ tile_mask) {
  const int64_t total = (int64_t)cap_out * ks * ks;
  const int n_out = n * ho * wo;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(e % cap_out), k = (int)(e / cap_out);
    int src = -1;
    if (row < n_out) {
      const int ox = row % wo, oy = (row / wo) % ho, img = row / (wo * ho);
      const int iy = oy * stride - pad + k / ks, ix = ox * stride - pad + k % ks;
      if (iy >= 0 && iy < h && ix >= 0 && ix < w) src = (img * h + iy) * w + ix;
    }
    nbr[e] = src;
    if (k == 0 && (row & 127) == 0) tile_mask[row >> 7] = row < n_out ? ((1u << (ks * ks)) - 1u) : 0u;
  }
}

}  // namespace srf

using namespace srf;

extern "C" {

int srf_version(void) { return 100; }
const char* srf_last_error(void) { return g_err; }
int srf_sm_count(void) { return sm_count(); }
unsigned long long srf_launch_count(void) { return g_launches; }

size_t srf_index_bytes(int64_t ncells) {
  int64_t nw = (index_nwords(ncells) + 3) / 4 * 4;
  return (size_t)(2 * nw + SCAN_BLOCKS + 4) * sizeof(uint32_t);
}

int srf_index_clear(void* index, int64_t ncells, void* stream) {
  SRF_CHECK_ARG(index && ncells > 0, "srf_index_clear: bad args");
  IndexView v = index_view(index, ncells);
  SRF_CUDA(cudaMemsetAsync(v.bits, 0, (size_t)v.nwords * sizeof(uint32_t), (cudaStream_t)stream));
  return SRF_OK;
}

static Dims4 dims4(const int32_t d[4]) { return Dims4{d[0], d[1], d[2], d[3]}; }
static int64_t ncells4(const int32_t d[4]) { return (int64_t)d[0] * d[1] * d[2] * d[3]; }
static int grid_for(int64_t n, int threads) {
  int64_t g = (n + threads - 1) / threads;
  int64_t cap = (int64_t)sm_count() * 16;
  if (g > cap) g = cap;
  return g < 1 ? 1 : (int)g;
}

int srf_index_mark(void* index, const int32_t dims[4], const int32_t* coors, int32_t n,
                   const int32_t* d_n, void* stream) {
  SRF_CHECK_ARG(index && dims && coors && n >= 0, "srf_index_mark: bad args");
  if (n == 0) return SRF_OK;
  IndexView v = index_view(index, ncells4(dims));
  SRF_COUNT(1);
  index_mark_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(v.bits, dims4(dims),
                                                                        (const int4*)coors, n, d_n);
  SRF_LAUNCH_CHECK();
  return SRF_OK;
}

static int make_conv3(Conv3* cv, const int32_t k[3], const int32_t s[3], const int32_t p[3]) {
  for (int j = 0; j < 3; ++j) {
    cv->k[j] = k[j];
    cv->s[j] = s[j];
    cv->p[j] = p[j];
    if (k[j] < 1 || s[j] < 1 || p[j] < 0) return -1;
  }
  if (k[0] * k[1] * k[2] > 27) return -1;
  return 0;
}

int srf_index_mark_strided(void* out_index, const int32_t out_dims[4], const int32_t* in_coors,
                           int32_t cap_in, const int32_t* d_n_in, const int32_t ksize[3],
                           const int32_t stride[3], const int32_t pad[3], void* stream) {
  SRF_CHECK_ARG(out_index && out_dims && in_coors && cap_in >= 0, "srf_index_mark_strided: bad args");
  Conv3 cv;
  SRF_CHECK_ARG(make_conv3(&cv, ksize, stride, pad) == 0, "srf_index_mark_strided: bad conv geometry");
  if (cap_in == 0) return SRF_OK;
  IndexView v = index_view(out_index, ncells4(out_dims));
  SRF_COUNT(1);
  int sh[3];
  bool pow2 = ncells4(out_dims) < (1ll << 31);
  for (int j = 0; j < 3; ++j) {
    sh[j] = 0;
    while ((1 << sh[j]) < cv.s[j]) ++sh[j];
    pow2 = pow2 && (1 << sh[j]) == cv.s[j];
  }
  if (pow2)
    index_mark_strided_kernel<true><<<grid_for(cap_in, 256), 256, 0, (cudaStream_t)stream>>>(
        v.bits, dims4(out_dims), (const int4*)in_coors, cap_in, d_n_in, cv, sh[0], sh[1], sh[2]);
  else
    index_mark_strided_kernel<false><<<grid_for(cap_in, 256), 256, 0, (cudaStream_t)stream>>>(
        v.bits, dims4(out_dims), (const int4*)in_coors, cap_in, d_n_in, cv, sh[0], sh[1], sh[2]);
  SRF_LAUNCH_CHECK();
  return SRF_OK;
}

int srf_index_finalize(void* index, int64_t ncells, int32_t* d_num, void* stream) {
  SRF_CHECK_ARG(index && ncells > 0, "srf_index_finalize: bad args");
  IndexView v = index_view(index, ncells);
  return scan_flags_launch(v.bits, v.rank, v.blocksum, v.nwords, d_num, 1, (cudaStream_t)stream);
}

int srf_index_emit_coors(const void* index, const int32_t dims[4], int32_t* coors_out, int32_t cap,
                         void* stream) {
  SRF_CHECK_ARG(index && dims && coors_out && cap >= 0, "srf_index_emit_coors: bad args");
  IndexView v = index_view(index, ncells4(dims));
  SRF_COUNT(1);
  if (ncells4(dims) < (1ll << 31))
    index_emit_kernel<uint32_t><<<grid_for(v.nwords, 256), 256, 0, (cudaStream_t)stream>>>(v.bits, v.rank, v.nwords, dims4(dims), (int4*)coors_out, cap);
  else
    index_emit_kernel<int64_t><<<grid_for(v.nwords, 256), 256, 0, (cudaStream_t)stream>>>(v.bits, v.rank, v.nwords, dims4(dims), (int4*)coors_out, cap);
  SRF_LAUNCH_CHECK();
  return SRF_OK;
}

int srf_index_lookup(const void* index, const int32_t dims[4], const int32_t* coors, int32_t n,
                     const int32_t* d_n, int32_t* rows, void* stream) {
  SRF_CHECK_ARG(index && dims && coors && rows && n >= 0, "srf_index_lookup: bad args");
  if (n == 0) return SRF_OK;
  IndexView v = index_view(index, ncells4(dims));
  SRF_COUNT(1);
  index_lookup_kernel<false><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
      v.bits, v.rank, dims4(dims), (const int4*)coors, n, d_n, rows);
  SRF_LAUNCH_CHECK();
  return SRF_OK;
}

int srf_index_perm(const void* index, const int32_t dims[4], const int32_t* coors, int32_t n,
                   const int32_t* d_n, int32_t* perm, void* stream) {
  SRF_CHECK_ARG(index && dims && coors && perm && n >= 0, "srf_index_perm: bad args");
  if (n == 0) return SRF_OK;
  IndexView v = index_view(index, ncells4(dims));
  SRF_COUNT(1);
  index_lookup_kernel<true><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
      v.bits, v.rank, dims4(dims), (const int4*)coors, n, d_n, perm);
  SRF_LAUNCH_CHECK();
  return SRF_OK;
}

// Rulebook of a DENSE 2-D convolution over (n, h, w) row-major pixel rows (the BEV backbone / neck
// convs of SURVEY.md 8f rank 1 run on the same gather-GEMM kernel as the sparse layers): every
// output pixel exists, a neighbour is missing only in the zero padding.  Static per shape.
int srf_dense_rulebook(int32_t n, int32_t h, int32_t w, int32_t ksize, int32_t stride, int32_t pad, int32_t cap_out, int32_t* nbr,
                       uint32_t* tile_mask, void* stream) {
  SRF_CHECK_ARG(nbr && tile_mask && n >= 1 && h >= 1 && w >= 1 && ksize >= 1 && ksize * ksize <= 27 && stride >= 1 && pad >= 0,
                "srf_dense_rulebook: bad args");
  const int ho = (h + 2 * pad - ksize) / stride + 1, wo = (w + 2 * pad - ksize) / stride + 1;
  SRF_CHECK_ARG(cap_out % 128 == 0 && cap_out >= n * ho * wo, "srf_dense_rulebook: cap_out must be a multiple of 128 covering n*ho*wo");
  SRF_COUNT(1);
  dense_rulebook_kernel<<<grid_for((int64_t)cap_out * ksize * ksize, 256), 256, 0, (cudaStream_t)stream>>>(n, h, w, ho, wo, ksize, stride, pad,
                                                                                                       cap_out, nbr, tile_mask);
  SRF_LAUNCH_CHECK();
  return SRF_OK;
}

int srf_rulebook_build(const void* in_index, const int32_t in_dims[4], const int32_t* in_perm,
                       const int32_t* out_coors, int32_t cap_out, const int32_t* d_n_out,
                       const int32_t ksize[3], const int32_t stride[3], const int32_t pad[3],
                       int32_t* nbr, uint32_t* tile_mask, void* stream) {
  SRF_CHECK_ARG(in_index && in_dims && out_coors && nbr && tile_mask, "srf_rulebook_build: null arg");
  SRF_CHECK_ARG(cap_out > 0 && cap_out % 128 == 0, "srf_rulebook_build: cap_out must be a positive multiple of 128");
  Conv3 cv;
  SRF_CHECK_ARG(make_conv3(&cv, ksize, stride, pad) == 0, "srf_rulebook_build: bad conv geometry");
  cudaStream_t st = (cudaStream_t)stream;
  SRF_CUDA(cudaMemsetAsync(tile_mask, 0, (size_t)(cap_out / 128) * sizeof(uint32_t), st));
  IndexView v = index_view(in_index, ncells4(in_dims));
  SRF_COUNT(1);
  const bool small = ncells4(in_dims) < (1ll << 31);
  const int grid = grid_for(cap_out, 256);
#define SRF_RB(k0, k1, k2, I) rulebook_fast_kernel<k0, k1, k2, I><<<grid, 256, 0, st>>>(v.bits, v.rank, dims4(in_dims), in_perm, (const int4*)out_coors, cap_out, d_n_out, cv, nbr, tile_mask)
  if (cv.k[0] == 3 && cv.k[1] == 3 && cv.k[2] == 3) {
    if (small) SRF_RB(3, 3, 3, uint32_t); else SRF_RB(3, 3, 3, int64_t);
  } else if (cv.k[0] == 3 && cv.k[1] == 1 && cv.k[2] == 1) {
    if (small) SRF_RB(3, 1, 1, uint32_t); else SRF_RB(3, 1, 1, int64_t);
  } else {
    rulebook_kernel<<<grid, 256, 0, st>>>(v.bits, v.rank, dims4(in_dims), in_perm, (const int4*)out_coors, cap_out, d_n_out, cv, nbr, tile_mask);
  }
#undef SRF_RB
  SRF_LAUNCH_CHECK();
  return SRF_OK;
}

}  // extern "C"
