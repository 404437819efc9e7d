// Pillar path of the srfdet_pillar_* configs:
//   PillarFeatureNetCustom.forward (mmdet3d_plugin/models/voxel_encoders/pillar_encoder_custom.py:95-161)
//   with its single PFNLayer (voxel_encoders/utils.py:109-147; Linear(no bias) -> BN1d(eval) -> ReLU ->
//   max | avg over the T point slots) fused into one pass over the hard-voxelized pillars, and
//   mmdet3d PointPillarsScatter (cfg configs/nus/srfdet_pillar_nusc_L.py:53-54).
// The decorated point features (raw C, offset to the pillar's point mean, offset to the pillar
// centre, optional range) never touch HBM; padded slots are zeroed AFTER decoration exactly like the
// reference (features *= mask), so a padded slot contributes relu(folded bias) to the max.
#include "common.cuh"

namespace srf {

constexpr int PILLAR_MAXT = 64;     // point slots per pillar
constexpr int PILLAR_MAXF = 16;     // decorated features per point

struct PillarArgs {
  const float* voxels;      // (cap, T, C)
  const int32_t* num_points;
  const int32_t* coors;     // (cap, 4) b,z,y,x
  const int32_t* d_n;
  int cap, T, C;
  const float* w;           // (cout, F) Linear weight with BatchNorm folded
  const float* b;           // (cout) folded BatchNorm shift
  int cout, F;
  float vx, vy, vz, xo, yo, zo;
  int with_cluster, with_center, with_distance, legacy, avg;
  float* out;               // (cap, cout)
};

// one block per pillar, one thread per output channel (cout <= 128)
__global__ void __launch_bounds__(128) pillar_vfe_kernel(const PillarArgs a) {
  __shared__ float sf[PILLAR_MAXT][PILLAR_MAXF + 1];
  __shared__ float smean[3];
  const int n = a.d_n ? min(*a.d_n, a.cap) : a.cap;
  for (int v = blockIdx.x; v < n; v += gridDim.x) {
    const float* pv = a.voxels + (size_t)v * a.T * a.C;
    const int np = __ldg(a.num_points + v);
    __syncthreads();
    if (threadIdx.x < 3) {
      // points_mean = features[:, :, :3].sum(dim=1) / num_points (padded slots are zeros, in slot order)
      float s = 0.f;
      for (int t = 0; t < a.T; ++t) s += __ldg(pv + t * a.C + threadIdx.x);
      smean[threadIdx.x] = s / (float)np;
    }
    __syncthreads();
    if (threadIdx.x < a.T) {
      const int t = threadIdx.x;
      const int4 q = __ldg(reinterpret_cast<const int4*>(a.coors) + v);
      float raw[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) raw[c] = c < a.C ? __ldg(pv + t * a.C + c) : 0.f;     // fully unrolled: stays in registers
      const float cx = raw[0] - ((float)q.w * a.vx + a.xo), cy = raw[1] - ((float)q.z * a.vy + a.yo),
                  cz = raw[2] - ((float)q.y * a.vz + a.zo);
      // features are written straight to shared memory one scalar at a time (a register array filled through a
      // running index made ptxas pair the stores into 8-byte local stores at 4-byte-aligned offsets)
      const float m = t < np ? 1.f : 0.f;     // get_paddings_indicator
      float* f = sf[t];
      int k = 0;
      // legacy=True: f_center aliases features[:, :, :3] and is modified in place, so the raw xyz
      // channels carry the centre offsets too (pillar_encoder_custom.py:133-143)
      const bool alias = a.with_center && a.legacy;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (c < a.C) f[k++] = m * ((alias && c < 3) ? (c == 0 ? cx : (c == 1 ? cy : cz)) : raw[c]);
      if (a.with_cluster) {
        f[k++] = m * (raw[0] - smean[0]);
        f[k++] = m * (raw[1] - smean[1]);
        f[k++] = m * (raw[2] - smean[2]);
      }
      if (a.with_center) {
        f[k++] = m * cx;
        f[k++] = m * cy;
        f[k++] = m * cz;
      }
      if (a.with_distance) {
        const float dx = alias ? cx : raw[0], dy = alias ? cy : raw[1], dz = alias ? cz : raw[2];
        f[k++] = m * sqrtf(dx * dx + dy * dy + dz * dz);
      }
    }
    __syncthreads();
    if (threadIdx.x < a.cout) {
      const float* wr = a.w + (size_t)threadIdx.x * a.F;
      float wreg[PILLAR_MAXF];
#pragma unroll
      for (int j = 0; j < PILLAR_MAXF; ++j) wreg[j] = j < a.F ? __ldg(wr + j) : 0.f;
      const float bias = __ldg(a.b + threadIdx.x);
      float red = a.avg ? 0.f : -INFINITY;
      for (int t = 0; t < a.T; ++t) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < PILLAR_MAXF; ++j)
          if (j < a.F) acc = fmaf(sf[t][j], wreg[j], acc);
        const float y = fmaxf(acc + bias, 0.f);
        red = a.avg ? red + y : fmaxf(red, y);
      }
      a.out[(size_t)v * a.cout + threadIdx.x] = a.avg ? red / (float)np : red;
    }
  }
}

// PointPillarsScatter: canvas[b, c, y, x] = feats[row, c] (canvas zeroed by the caller).
// channels_last: the canvas is (B, ny, nx, C) in memory (a torch.channels_last tensor).
__global__ void pillars_scatter_kernel(const float* __restrict__ feats, const int32_t* __restrict__ coors, const int32_t* __restrict__ d_n,
                                       int cap, int c, int ny, int nx, int channels_last, float* __restrict__ canvas) {
  const int n = d_n ? min(*d_n, cap) : cap;
  const int64_t total = (int64_t)n * c;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(e / c), ch = (int)(e % c);
    const int4 q = __ldg(reinterpret_cast<const int4*>(coors) + row);
    if (q.x < 0 || q.z < 0 || q.z >= ny || q.w < 0 || q.w >= nx) continue;
    const size_t cell = (size_t)q.z * nx + q.w;
    const size_t dst = channels_last ? ((size_t)q.x * ny * nx + cell) * c + ch : ((size_t)q.x * c + ch) * ny * nx + cell;
    canvas[dst] = __ldg(feats + e);
  }
}

}  // namespace srf

using namespace srf;

extern "C" {

int srf_pillar_vfe(const float* voxels, const int32_t* num_points, const int32_t* coors, int32_t cap, const int32_t* d_n,
                   int32_t t, int32_t c, const float* w_folded, const float* b_folded, int32_t cout, const float voxel_size[3],
                   const float offsets[3], int32_t flags, float* out, void* stream) {
  SRF_CHECK_ARG(voxels && num_points && coors && w_folded && b_folded && out && voxel_size && offsets, "srf_pillar_vfe: null arg");
  SRF_CHECK_ARG(cap >= 0 && t >= 1 && t <= PILLAR_MAXT && c >= 3 && c <= 8 && cout >= 1 && cout <= 128,
                "srf_pillar_vfe: need 1 <= T <= %d, 3 <= C <= 8, cout <= 128", PILLAR_MAXT);
  if (cap == 0) return SRF_OK;
  PillarArgs a;
  a.voxels = voxels; a.num_points = num_points; a.coors = coors; a.d_n = d_n; a.cap = cap; a.T = t; a.C = c;
  a.w = w_folded; a.b = b_folded; a.cout = cout;
  a.vx = voxel_size[0]; a.vy = voxel_size[1]; a.vz = voxel_size[2];
  a.xo = offsets[0]; a.yo = offsets[1]; a.zo = offsets[2];
  a.with_cluster = flags & 1; a.with_center = (flags >> 1) & 1; a.with_distance = (flags >> 2) & 1;
  a.legacy = (flags >> 3) & 1; a.avg = (flags >> 4) & 1;
  a.F = c + (a.with_cluster ? 3 : 0) + (a.with_center ? 3 : 0) + (a.with_distance ? 1 : 0);
  SRF_CHECK_ARG(a.F <= PILLAR_MAXF, "srf_pillar_vfe: too many decorated features");
  int grid = cap < sm_count() * 16 ? cap : sm_count() * 16;
  SRF_COUNT(1);
  pillar_vfe_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(a);
  SRF_LAUNCH_CHECK();
  return SRF_OK;
}

int srf_pillars_scatter(const float* feats, const int32_t* coors, int32_t cap, const int32_t* d_n, int32_t c, int32_t ny,
                        int32_t nx, int32_t channels_last, float* canvas, void* stream) {
  SRF_CHECK_ARG(feats && coors && canvas && cap >= 0 && c >= 1 && ny >= 1 && nx >= 1, "srf_pillars_scatter: bad args");
  if (cap == 0) return SRF_OK;
  int64_t total = (int64_t)cap * c;
  int64_t g = (total + 255) / 256;
  if (g > sm_count() * 16) g = sm_count() * 16;
  SRF_COUNT(1);
  pillars_scatter_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(feats, coors, d_n, cap, c, ny, nx, channels_last, canvas);
  SRF_LAUNCH_CHECK();
  return SRF_OK;
}

}  // extern "C"
