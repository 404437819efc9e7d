// DynamicScatter and the fused DynamicVFECustom forward.
//
// mmcv's DynamicPointToVoxelForward sorts the point coordinates (at::unique_dim) and then
// reduces with atomics; DynamicVFECustom additionally builds a dense int64 canvas of the
// whole grid (0.7 GB) just to map points back to their voxel
// (voxel_encoders/voxel_encoder.py:118-158).  Here the bitmap index gives the sorted voxel
// order and the point->voxel map directly, so one pass marks, one scan ranks, and the
// per-point MLPs + reductions run fused over the points.
#include "common.cuh"

namespace srf {

// (branch on the SIGN BIT, not on v >= 0: -0.0f must take the unsigned path, otherwise its pattern 0x80000000 =
// INT_MIN never beats the -inf initial value and a channel whose maximum is -0.0 would come out as -inf)
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (__float_as_int(v) >= 0)
    atomicMax((int*)addr, __float_as_int(v));
  else
    atomicMin((unsigned int*)addr, __float_as_uint(v));
}

template <int CD>
__device__ __forceinline__ bool load_cell(const int32_t* __restrict__ coors, int i, const Dims4& d,
                                          int64_t& cell) {
  int b, z, y, x;
  if (CD == 4) {
    int4 q = __ldg(reinterpret_cast<const int4*>(coors) + i);
    b = q.x; z = q.y; y = q.z; x = q.w;
  } else {
    const int32_t* q = coors + (size_t)i * 3;
    b = 0; z = __ldg(q); y = __ldg(q + 1); x = __ldg(q + 2);
  }
  if (b < 0 || z < 0 || y < 0 || x < 0 || b >= d.b || z >= d.z || y >= d.y || x >= d.x) return false;
  cell = cell_of(d, b, z, y, x);
  return true;
}

template <int CD>
__global__ void sc_mark_kernel(uint32_t* __restrict__ bits, Dims4 d, const int32_t* __restrict__ coors, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int64_t cell;
    if (load_cell<CD>(coors, i, d, cell)) atomicOr(bits + (cell >> 5), 1u << (cell & 31));
  }
}

template <int CD>
__global__ void sc_emit_kernel(const uint32_t* __restrict__ bits, const uint32_t* __restrict__ rank,
                               int64_t nwords, Dims4 d, int32_t* __restrict__ out, int cap) {
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords;
       w += (int64_t)gridDim.x * blockDim.x) {
    uint32_t word = __ldg(bits + w);
    if (!word) continue;
    int r = (int)__ldg(rank + w);
    while (word) {
      int bpos = __ffs(word) - 1;
      word &= word - 1;
      int64_t cell = (w << 5) + bpos;
      int x = (int)(cell % d.x); cell /= d.x;
      int y = (int)(cell % d.y); cell /= d.y;
      int z = (int)(cell % d.z); cell /= d.z;
      if (r < cap) {
        if (CD == 4) reinterpret_cast<int4*>(out)[r] = make_int4((int)cell, z, y, x);
        else { out[(size_t)r * 3] = z; out[(size_t)r * 3 + 1] = y; out[(size_t)r * 3 + 2] = x; }
      }
      ++r;
    }
  }
}

__global__ void fill_kernel(float* __restrict__ p, int64_t n, float v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = v;
}

template <int CD, bool MEAN>
__global__ void sc_reduce_kernel(const uint32_t* __restrict__ bits, const uint32_t* __restrict__ rank,
                                 Dims4 d, const float* __restrict__ feats, const int32_t* __restrict__ coors,
                                 int n, int c, float* __restrict__ out, int32_t* __restrict__ count,
                                 int32_t* __restrict__ p2v) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int64_t cell;
    int r = -1;
    if (load_cell<CD>(coors, i, d, cell)) r = index_rank(bits, rank, cell);
    if (p2v) p2v[i] = r;
    if (r < 0) continue;
    const float* f = feats + (size_t)i * c;
    float* o = out + (size_t)r * c;
    for (int j = 0; j < c; ++j) {
      float v = __ldg(f + j);
      if (MEAN) atomicAdd(o + j, v);
      else atomic_max_float(o + j, v);
    }
    if (MEAN) atomicAdd(count + r, 1);
  }
}

__global__ void sc_divide_kernel(float* __restrict__ out, const int32_t* __restrict__ count,
                                 const int32_t* __restrict__ d_num, int c) {
  int64_t total = (int64_t)(*d_num) * c;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
    out[e] = __fdiv_rn(out[e], (float)count[e / c]);
}

// ---- fused DynamicVFECustom -----------------------------------------------------------
struct VfeDev {
  const float *pos_w0, *pos_b0, *pos_w1, *pos_b1, *vfe_w0, *vfe_b0, *vfe_w1, *vfe_b1;
  int cin, c0, c1;
  float vx, vy, vz, x_off, y_off, z_off;
};

// cluster sums: per point rank + atomicAdd xyz, count  (voxel_encoder.py:189 cluster_scatter)
__global__ void vfe_cluster_kernel(const uint32_t* __restrict__ bits, const uint32_t* __restrict__ rank, Dims4 d,
                                   const float* __restrict__ pts, const int32_t* __restrict__ coors, int n, int cin,
                                   float* __restrict__ sums /* (cap,4): x,y,z,count */, int32_t* __restrict__ p2v) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int64_t cell;
    int r = -1;
    if (load_cell<4>(coors, i, d, cell)) r = index_rank(bits, rank, cell);
    p2v[i] = r;
    if (r < 0) continue;
    const float* p = pts + (size_t)i * cin;
    float* s = sums + (size_t)r * 4;
    atomicAdd(s + 0, __ldg(p));
    atomicAdd(s + 1, __ldg(p + 1));
    atomicAdd(s + 2, __ldg(p + 2));
    atomicAdd(s + 3, 1.0f);
  }
}

constexpr int VFE_MAXC = 8;
constexpr int VFE_POS = 32;

// per point: centroid-aware position encoding MLP, voxel-centre offsets, first VFE layer
// (voxel_encoder.py:193-230), max-reduced into vmax0.
__global__ void __launch_bounds__(128) vfe_layer0_kernel(const float* __restrict__ pts, const int32_t* __restrict__ coors,
                                                        const int32_t* __restrict__ p2v, const float* __restrict__ sums,
                                                        int n, VfeDev P, float* __restrict__ pf0, float* __restrict__ vmax0) {
  __shared__ float s_w0[VFE_POS * 3], s_b0[VFE_POS], s_w1[VFE_POS * VFE_POS], s_b1[VFE_POS];
  __shared__ float s_v0[VFE_MAXC * (VFE_MAXC + VFE_POS + 3)], s_vb0[VFE_MAXC];
  const int fin = P.cin + VFE_POS + 3;
  for (int t = threadIdx.x; t < VFE_POS * 3; t += blockDim.x) s_w0[t] = P.pos_w0[t];
  for (int t = threadIdx.x; t < VFE_POS; t += blockDim.x) { s_b0[t] = P.pos_b0[t]; s_b1[t] = P.pos_b1[t]; }
  for (int t = threadIdx.x; t < VFE_POS * VFE_POS; t += blockDim.x) s_w1[t] = P.pos_w1[t];
  for (int t = threadIdx.x; t < P.c0 * fin; t += blockDim.x) s_v0[t] = P.vfe_w0[t];
  for (int t = threadIdx.x; t < P.c0; t += blockDim.x) s_vb0[t] = P.vfe_b0[t];
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int r = p2v[i];
    if (r < 0) continue;
    float feat[VFE_MAXC];
    const float* p = pts + (size_t)i * P.cin;
#pragma unroll
    for (int j = 0; j < VFE_MAXC; ++j) feat[j] = j < P.cin ? __ldg(p + j) : 0.f;
    float4 s = __ldg(reinterpret_cast<const float4*>(sums) + r);
    float fc[3] = {feat[0] - __fdiv_rn(s.x, s.w), feat[1] - __fdiv_rn(s.y, s.w), feat[2] - __fdiv_rn(s.z, s.w)};
    float h1[VFE_POS];
#pragma unroll
    for (int o = 0; o < VFE_POS; ++o)
      h1[o] = tanhf(s_w0[o * 3] * fc[0] + s_w0[o * 3 + 1] * fc[1] + s_w0[o * 3 + 2] * fc[2] + s_b0[o]);
    int4 q = __ldg(reinterpret_cast<const int4*>(coors) + i);
    float ctr[3] = {feat[0] - ((float)q.w * P.vx + P.x_off), feat[1] - ((float)q.z * P.vy + P.y_off),
                    feat[2] - ((float)q.y * P.vz + P.z_off)};
    float acc[VFE_MAXC];
#pragma unroll
    for (int o = 0; o < VFE_MAXC; ++o) {
      acc[o] = 0.f;
      if (o < P.c0) {
        float a = s_vb0[o];
        const float* w = s_v0 + o * fin;
        for (int j = 0; j < P.cin; ++j) a += w[j] * feat[j];
        a += w[P.cin + VFE_POS] * ctr[0] + w[P.cin + VFE_POS + 1] * ctr[1] + w[P.cin + VFE_POS + 2] * ctr[2];
        acc[o] = a;
      }
    }
    // second pos-enc layer streamed: h2[o] is consumed as soon as it is produced
    for (int o = 0; o < VFE_POS; ++o) {
      float a = s_b1[o];
      const float* w = s_w1 + o * VFE_POS;
#pragma unroll
      for (int j = 0; j < VFE_POS; ++j) a += w[j] * h1[j];
      float h2 = tanhf(a);
#pragma unroll
      for (int oo = 0; oo < VFE_MAXC; ++oo)
        if (oo < P.c0) acc[oo] += s_v0[oo * fin + P.cin + o] * h2;
    }
#pragma unroll
    for (int o = 0; o < VFE_MAXC; ++o)
      if (o < P.c0) {
        float y = fmaxf(acc[o], 0.f);
        if (pf0) pf0[(size_t)i * P.c0 + o] = y;
        atomicMax((int*)(vmax0 + (size_t)r * P.c0 + o), __float_as_int(y));  // y >= 0
      }
  }
}

// second VFE layer: cat(point_feats, voxel max of layer 0) (voxel_encoder.py:233-237)
__global__ void __launch_bounds__(128) vfe_layer1_kernel(const int32_t* __restrict__ p2v, const float* __restrict__ pf0,
                                                        const float* __restrict__ vmax0, int n, VfeDev P,
                                                        float* __restrict__ vmax1) {
  __shared__ float s_w[VFE_MAXC * 2 * VFE_MAXC], s_b[VFE_MAXC];
  for (int t = threadIdx.x; t < P.c1 * 2 * P.c0; t += blockDim.x) s_w[t] = P.vfe_w1[t];
  for (int t = threadIdx.x; t < P.c1; t += blockDim.x) s_b[t] = P.vfe_b1[t];
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int r = p2v[i];
    if (r < 0) continue;
    float in[2 * VFE_MAXC];
#pragma unroll
    for (int j = 0; j < VFE_MAXC; ++j) {
      in[j] = j < P.c0 ? pf0[(size_t)i * P.c0 + j] : 0.f;
      in[VFE_MAXC + j] = j < P.c0 ? vmax0[(size_t)r * P.c0 + j] : 0.f;
    }
#pragma unroll
    for (int o = 0; o < VFE_MAXC; ++o)
      if (o < P.c1) {
        float a = s_b[o];
        const float* w = s_w + o * 2 * P.c0;
        for (int j = 0; j < P.c0; ++j) a += w[j] * in[j] + w[P.c0 + j] * in[VFE_MAXC + j];
        atomicMax((int*)(vmax1 + (size_t)r * P.c1 + o), __float_as_int(fmaxf(a, 0.f)));
      }
  }
}

static int lgrid(int64_t n, int threads) {
  int64_t g = (n + threads - 1) / threads;
  int64_t cap = (int64_t)sm_count() * 8;
  if (g > cap) g = cap;
  return g < 1 ? 1 : (int)g;
}

static size_t al256(size_t b) { return (b + 255) / 256 * 256; }

}  // namespace srf

using namespace srf;

extern "C" {

size_t srf_scatter_ws_bytes(int64_t ncells, int32_t n, int32_t c) {
  (void)c;
  return al256(srf_index_bytes(ncells)) + al256((size_t)n * 4) + 256;
}

int srf_dynamic_scatter(const float* feats, const int32_t* coors, int32_t n, int32_t c, int32_t coor_dim,
                        const int32_t dims[4], int32_t mode, float* out_feats, int32_t* out_coors,
                        int32_t* d_num_voxels, int32_t* point2voxel, void* ws, size_t ws_bytes, void* stream) {
  SRF_CHECK_ARG(feats && coors && dims && out_feats && out_coors && d_num_voxels && ws, "srf_dynamic_scatter: null arg");
  SRF_CHECK_ARG(coor_dim == 3 || coor_dim == 4, "srf_dynamic_scatter: coor_dim must be 3 or 4");
  SRF_CHECK_ARG(mode == 0 || mode == 1, "srf_dynamic_scatter: mode must be 0 (max) or 1 (mean)");
  SRF_CHECK_ARG(n >= 0 && c > 0, "srf_dynamic_scatter: bad sizes");
  int64_t ncells = (int64_t)dims[0] * dims[1] * dims[2] * dims[3];
  SRF_CHECK_ARG(ncells > 0, "srf_dynamic_scatter: empty grid");
  SRF_CHECK_ARG(ws_bytes >= srf_scatter_ws_bytes(ncells, n, c), "srf_dynamic_scatter: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    SRF_CUDA(cudaMemsetAsync(d_num_voxels, 0, 4, st));
    return SRF_OK;
  }
  IndexView v = index_view(ws, ncells);
  int32_t* count = (int32_t*)((char*)ws + al256(srf_index_bytes(ncells)));
  Dims4 d{dims[0], dims[1], dims[2], dims[3]};
  SRF_CUDA(cudaMemsetAsync(v.bits, 0, (size_t)v.nwords * 4, st));
  int g = lgrid(n, 256);
  SRF_COUNT(4);
  if (coor_dim == 4) sc_mark_kernel<4><<<g, 256, 0, st>>>(v.bits, d, coors, n);
  else sc_mark_kernel<3><<<g, 256, 0, st>>>(v.bits, d, coors, n);
  int rc = scan_flags_launch(v.bits, v.rank, v.blocksum, v.nwords, d_num_voxels, 1, st);
  if (rc) return rc;
  if (coor_dim == 4) sc_emit_kernel<4><<<lgrid(v.nwords, 256), 256, 0, st>>>(v.bits, v.rank, v.nwords, d, out_coors, n);
  else sc_emit_kernel<3><<<lgrid(v.nwords, 256), 256, 0, st>>>(v.bits, v.rank, v.nwords, d, out_coors, n);
  if (mode == 1) {
    SRF_CUDA(cudaMemsetAsync(out_feats, 0, (size_t)n * c * 4, st));
    SRF_CUDA(cudaMemsetAsync(count, 0, (size_t)n * 4, st));
    if (coor_dim == 4) sc_reduce_kernel<4, true><<<g, 256, 0, st>>>(v.bits, v.rank, d, feats, coors, n, c, out_feats, count, point2voxel);
    else sc_reduce_kernel<3, true><<<g, 256, 0, st>>>(v.bits, v.rank, d, feats, coors, n, c, out_feats, count, point2voxel);
    sc_divide_kernel<<<lgrid((int64_t)n * c, 256), 256, 0, st>>>(out_feats, count, d_num_voxels, c);
  } else {
    fill_kernel<<<lgrid((int64_t)n * c, 256), 256, 0, st>>>(out_feats, (int64_t)n * c, -INFINITY);
    if (coor_dim == 4) sc_reduce_kernel<4, false><<<g, 256, 0, st>>>(v.bits, v.rank, d, feats, coors, n, c, out_feats, count, point2voxel);
    else sc_reduce_kernel<3, false><<<g, 256, 0, st>>>(v.bits, v.rank, d, feats, coors, n, c, out_feats, count, point2voxel);
  }
  SRF_LAUNCH_CHECK();
  return SRF_OK;
}

size_t srf_dynamic_vfe_ws_bytes(int64_t ncells, int32_t n) {
  // index | p2v (n) | sums (n,4) | pf0 (n,8) | vmax0 (n,8)
  return al256(srf_index_bytes(ncells)) + al256((size_t)n * 4) + al256((size_t)n * 16) +
         2 * al256((size_t)n * VFE_MAXC * 4) + 256;
}

int srf_dynamic_vfe(const float* points, const int32_t* coors, int32_t n, const int32_t dims[4],
                    const srf_vfe_params* p, float* out_feats, int32_t* out_coors, int32_t* d_num_voxels,
                    void* ws, size_t ws_bytes, void* stream) {
  SRF_CHECK_ARG(points && coors && dims && p && out_feats && out_coors && d_num_voxels && ws, "srf_dynamic_vfe: null arg");
  SRF_CHECK_ARG(p->cin >= 3 && p->cin <= VFE_MAXC && p->c0 >= 1 && p->c0 <= VFE_MAXC && p->c1 >= 0 && p->c1 <= VFE_MAXC,
                "srf_dynamic_vfe: channel counts must be <= %d", VFE_MAXC);
  SRF_CHECK_ARG(p->pos_w0 && p->pos_b0 && p->pos_w1 && p->pos_b1 && p->vfe_w0 && p->vfe_b0, "srf_dynamic_vfe: null weights");
  SRF_CHECK_ARG(p->c1 == 0 || (p->vfe_w1 && p->vfe_b1), "srf_dynamic_vfe: second layer weights missing");
  int64_t ncells = (int64_t)dims[0] * dims[1] * dims[2] * dims[3];
  SRF_CHECK_ARG(ncells > 0 && n >= 0, "srf_dynamic_vfe: bad sizes");
  SRF_CHECK_ARG(ws_bytes >= srf_dynamic_vfe_ws_bytes(ncells, n), "srf_dynamic_vfe: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    SRF_CUDA(cudaMemsetAsync(d_num_voxels, 0, 4, st));
    return SRF_OK;
  }
  char* base = (char*)ws;
  IndexView v = index_view(base, ncells);
  base += al256(srf_index_bytes(ncells));
  int32_t* p2v = (int32_t*)base; base += al256((size_t)n * 4);
  float* sums = (float*)base; base += al256((size_t)n * 16);
  float* pf0 = (float*)base; base += al256((size_t)n * VFE_MAXC * 4);
  float* vmax0 = (float*)base;
  Dims4 d{dims[0], dims[1], dims[2], dims[3]};
  VfeDev P{p->pos_w0, p->pos_b0, p->pos_w1, p->pos_b1, p->vfe_w0, p->vfe_b0, p->vfe_w1, p->vfe_b1,
           p->cin, p->c0, p->c1, p->vx, p->vy, p->vz, p->x_off, p->y_off, p->z_off};
  const int c_last = p->c1 ? p->c1 : p->c0;
  SRF_CUDA(cudaMemsetAsync(v.bits, 0, (size_t)v.nwords * 4, st));
  SRF_CUDA(cudaMemsetAsync(sums, 0, (size_t)n * 16, st));
  // post-ReLU features are >= 0, so 0 is the identity of the max reduction
  SRF_CUDA(cudaMemsetAsync(out_feats, 0, (size_t)n * c_last * 4, st));
  if (p->c1) SRF_CUDA(cudaMemsetAsync(vmax0, 0, (size_t)n * p->c0 * 4, st));
  int g = lgrid(n, 256);
  SRF_COUNT(p->c1 ? 5 : 4);
  sc_mark_kernel<4><<<g, 256, 0, st>>>(v.bits, d, coors, n);
  int rc = scan_flags_launch(v.bits, v.rank, v.blocksum, v.nwords, d_num_voxels, 1, st);
  if (rc) return rc;
  sc_emit_kernel<4><<<lgrid(v.nwords, 256), 256, 0, st>>>(v.bits, v.rank, v.nwords, d, out_coors, n);
  vfe_cluster_kernel<<<g, 256, 0, st>>>(v.bits, v.rank, d, points, coors, n, p->cin, sums, p2v);
  int g128 = lgrid(n, 128);
  if (p->c1) {
    vfe_layer0_kernel<<<g128, 128, 0, st>>>(points, coors, p2v, sums, n, P, pf0, vmax0);
    vfe_layer1_kernel<<<g128, 128, 0, st>>>(p2v, pf0, vmax0, n, P, out_feats);
  } else {
    vfe_layer0_kernel<<<g128, 128, 0, st>>>(points, coors, p2v, sums, n, P, nullptr, out_feats);
  }
  SRF_LAUNCH_CHECK();
  return SRF_OK;
}

}  // extern "C"
