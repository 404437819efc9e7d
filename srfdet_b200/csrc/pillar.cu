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

constexpr int PILLAR_MAXT = 1024;   // point slots per pillar
constexpr int PILLAR_MAXF = 64;     // decorated features per point

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

// one thread per (pillar, output channel): the T x F decorated features are recomputed by every channel's thread
// (a few hundred flops; the pillar's 100-odd input floats are L1 hits after the first thread) -- no staging, no
// dynamically indexed arrays.  Weights are read through the read-only cache (one row of F floats per thread).
__global__ void __launch_bounds__(256) pillar_vfe_kernel(const PillarArgs a) {
  const int n = a.d_n ? min(*a.d_n, a.cap) : a.cap;
  const long long total = (long long)n * a.cout;
  const bool alias = a.with_center && a.legacy;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(e / a.cout), co = (int)(e % a.cout);
    const float* pv = a.voxels + (size_t)v * a.T * a.C;
    const int np = __ldg(a.num_points + v);
    const int32_t* cq = a.coors + (size_t)v * 4;
    const float ctr_x = (float)__ldg(cq + 3) * a.vx + a.xo, ctr_y = (float)__ldg(cq + 2) * a.vy + a.yo,
                ctr_z = (float)__ldg(cq + 1) * a.vz + a.zo;
    // points_mean = features[:, :, :3].sum(dim=1) / num_points (padded slots are zeros, summed in slot order)
    float mx = 0.f, my = 0.f, mz = 0.f;
    for (int t = 0; t < a.T; ++t) {
      mx += __ldg(pv + t * a.C);
      my += __ldg(pv + t * a.C + 1);
      mz += __ldg(pv + t * a.C + 2);
    }
    mx /= (float)np; my /= (float)np; mz /= (float)np;
    const float* wr = a.w + (size_t)co * a.F;
    const float bias = __ldg(a.b + co);
    float red = a.avg ? 0.f : -INFINITY;
    for (int t = 0; t < a.T; ++t) {
      const float m = t < np ? 1.f : 0.f;     // get_paddings_indicator: the decorated features of a padded slot are zeroed
      const float* pt = pv + t * a.C;
      const float x = __ldg(pt), y = __ldg(pt + 1), z = __ldg(pt + 2);
      const float cx = x - ctr_x, cy = y - ctr_y, cz = z - ctr_z;
      int j = 0;
      float acc = 0.f;
      // legacy=True: f_center aliases features[:, :, :3] and is modified in place, so the raw xyz channels carry the
      // centre offsets too (pillar_encoder_custom.py:133-143)
      acc = fmaf(m * (alias ? cx : x), __ldg(wr + j++), acc);
      acc = fmaf(m * (alias ? cy : y), __ldg(wr + j++), acc);
      acc = fmaf(m * (alias ? cz : z), __ldg(wr + j++), acc);
      for (int c = 3; c < a.C; ++c) acc = fmaf(m * __ldg(pt + c), __ldg(wr + j++), acc);
      if (a.with_cluster) {
        acc = fmaf(m * (x - mx), __ldg(wr + j++), acc);
        acc = fmaf(m * (y - my), __ldg(wr + j++), acc);
        acc = fmaf(m * (z - mz), __ldg(wr + j++), acc);
      }
      if (a.with_center) {
        acc = fmaf(m * cx, __ldg(wr + j++), acc);
        acc = fmaf(m * cy, __ldg(wr + j++), acc);
        acc = fmaf(m * cz, __ldg(wr + j++), acc);
      }
      if (a.with_distance) {
        const float dx = alias ? cx : x, dy = alias ? cy : y, dz = alias ? cz : z;
        acc = fmaf(m * sqrtf(dx * dx + dy * dy + dz * dz), __ldg(wr + j++), acc);
      }
      const float yv = fmaxf(acc + bias, 0.f);
      red = a.avg ? red + yv : fmaxf(red, yv);
    }
    a.out[e] = a.avg ? red / (float)np : red;
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
  SRF_CHECK_ARG(cap >= 0 && t >= 1 && t <= PILLAR_MAXT && c >= 3 && c <= 8 && cout >= 1,
                "srf_pillar_vfe: need 1 <= T <= %d, 3 <= C <= 8", PILLAR_MAXT);
  if (cap == 0) return SRF_OK;
  PillarArgs a = {};
  a.out = out;
  a.voxels = voxels; a.num_points = num_points; a.coors = coors; a.d_n = d_n; a.cap = cap; a.T = t; a.C = c;
  a.w = w_folded; a.b = b_folded; a.cout = cout;
  a.vx = voxel_size[0]; a.vy = voxel_size[1]; a.vz = voxel_size[2];
  a.xo = offsets[0]; a.yo = offsets[1]; a.zo = offsets[2];
  a.with_cluster = flags & 1; a.with_center = (flags >> 1) & 1; a.with_distance = (flags >> 2) & 1;
  a.legacy = (flags >> 3) & 1; a.avg = (flags >> 4) & 1;
  a.F = c + (a.with_cluster ? 3 : 0) + (a.with_center ? 3 : 0) + (a.with_distance ? 1 : 0);
  SRF_CHECK_ARG(a.F <= PILLAR_MAXF, "srf_pillar_vfe: too many decorated features");
  long long g = ((long long)cap * cout + 255) / 256;
  if (g > sm_count() * 16) g = sm_count() * 16;
  SRF_COUNT(1);
  pillar_vfe_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(a);
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
