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
                                             const float* __restrict__ b) {
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
    const float rstd = rsqrtf(v / n + 1e-5f);
    for (int j = lane; j < n; j += 32) x[j] = fmaxf((x[j] - mean) * rstd * __ldg(g + j) + __ldg(b + j), 0.f);
  }
}

template <typename TR, typename TP, typename TO>
__global__ void __launch_bounds__(256) dynconv_interact_kernel(const TR* __restrict__ roi, const TP* __restrict__ params,
                                                              int c, int d, const float* __restrict__ ln1_w,
                                                              const float* __restrict__ ln1_b, const float* __restrict__ ln2_w,
                                                              const float* __restrict__ ln2_b, TO* __restrict__ out) {
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
  ln_relu_rows(sT, DC_ROWS, d, ln1_w, ln1_b);
  __syncthreads();
  for (int e = threadIdx.x; e < DC_ROWS * c; e += blockDim.x) {
    const int s = e / c, j = e - s * c;
    const float* t = sT + s * d;
    float acc = 0.f;
    for (int i = 0; i < d; ++i) acc = fmaf(t[i], sP2[i * c + j], acc);
    sF[e] = acc;
  }
  __syncthreads();
  ln_relu_rows(sF, DC_ROWS, c, ln2_w, ln2_b);
  __syncthreads();
  TO* o = out + (size_t)k * DC_ROWS * c;
  for (int e = threadIdx.x; e < DC_ROWS * c; e += blockDim.x) st_f<TO>(o + e, sF[e]);
}

template <typename TR, typename TP, typename TO>
static int launch_dc(const void* roi, const void* params, int k, int c, int d, const float* a, const float* b,
                     const float* e, const float* f, void* out, cudaStream_t st) {
  size_t smem = (size_t)(DC_ROWS * c + 2 * c * d + DC_ROWS * d) * sizeof(float);
  auto kern = dynconv_interact_kernel<TR, TP, TO>;
  if (smem > 48 * 1024) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) { set_error("dynconv: cannot get %zu B shared memory: %s", smem, cudaGetErrorString(err)); return SRF_ERR_CUDA; }
  }
  SRF_COUNT(1);
  kern<<<k, 256, smem, st>>>((const TR*)roi, (const TP*)params, c, d, a, b, e, f, (TO*)out);
  return SRF_OK;
}

}  // namespace srf

using namespace srf;

extern "C" {

int srf_dynconv_interact(const void* roi, int32_t roi_dtype, const void* params, int32_t param_dtype, int32_t k,
                         int32_t c, int32_t d, const float* ln1_w, const float* ln1_b, const float* ln2_w,
                         const float* ln2_b, void* out, int32_t out_dtype, void* stream) {
  SRF_CHECK_ARG(roi && params && out && ln1_w && ln1_b && ln2_w && ln2_b, "srf_dynconv_interact: null arg");
  SRF_CHECK_ARG(k >= 0 && c >= 1 && c <= 256 && d >= 1 && d <= 64, "srf_dynconv_interact: need c<=256, d<=64");
  if (k == 0) return SRF_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  const bool rb = roi_dtype == SRF_BF16, pb = param_dtype == SRF_BF16, ob = out_dtype == SRF_BF16;
  if (!rb && !pb && !ob) rc = launch_dc<float, float, float>(roi, params, k, c, d, ln1_w, ln1_b, ln2_w, ln2_b, out, st);
  else if (rb && pb && ob) rc = launch_dc<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16>(roi, params, k, c, d, ln1_w, ln1_b, ln2_w, ln2_b, out, st);
  else if (!rb && pb && ob) rc = launch_dc<float, __nv_bfloat16, __nv_bfloat16>(roi, params, k, c, d, ln1_w, ln1_b, ln2_w, ln2_b, out, st);
  else if (rb && !pb && ob) rc = launch_dc<__nv_bfloat16, float, __nv_bfloat16>(roi, params, k, c, d, ln1_w, ln1_b, ln2_w, ln2_b, out, st);
  else { set_error("srf_dynconv_interact: unsupported dtype combination"); return SRF_ERR_UNSUPPORTED; }
  if (rc) return rc;
  SRF_LAUNCH_CHECK();
  return SRF_OK;
}

}  // extern "C"
