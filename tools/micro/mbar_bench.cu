// Microbenchmark: cost of one producer->consumer->producer ring hand-off per stage as a function
// of how many threads arrive on the "full" mbarrier and with which instruction (development tool).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mb mbar_bench.cu && /tmp/mb
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}

// V: 0 = every producer thread cp.async.mbarrier.arrive.noinc; 1 = every producer thread mbarrier.arrive;
//    2 = one lane per producer warp arrives; 3 = 0 + one dummy 16-byte cp.async per thread before the arrive
template <int V, int S, int IDLE = 0>
__global__ void ring(int iters, int npw, const uint8_t* src, long long* clk) {
  __shared__ __align__(8) uint64_t bars[2 * S + 1];
  extern __shared__ __align__(128) uint8_t buf[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t b0 = smem_u32(bars);
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(b0 + 8 * s, V == 2 ? npw : npw * 32); mbar_init(b0 + 8 * (S + s), 1); }
    mbar_init(b0 + 8 * 2 * S, 1);   // parked warps wait here until the ring is done
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  if (warp < npw) {
    for (int it = 0; it < iters; ++it) {
      const int s = it % S;
      const uint32_t ph = (uint32_t)(it / S) & 1u;
      mbar_wait(b0 + 8 * (S + s), ph ^ 1u);
      if (V == 3) asm volatile("cp.async.ca.shared.global [%0], [%1], 16, 16;" ::"r"(smem_u32(buf) + threadIdx.x * 16), "l"(src + threadIdx.x * 16) : "memory");
      if (V == 0 || V == 3) asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(b0 + 8 * s) : "memory");
      else if (V == 1) mbar_arrive(b0 + 8 * s);
      else { __syncwarp(); if (lane == 0) mbar_arrive(b0 + 8 * s); }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (warp == npw) {
    for (int it = 0; it < iters; ++it) {
      const int s = it % S;
      const uint32_t ph = (uint32_t)(it / S) & 1u;
      mbar_wait(b0 + 8 * s, ph);
      if (lane == 0) mbar_arrive(b0 + 8 * (S + s));
      __syncwarp();
    }
    if (IDLE && lane == 0) mbar_arrive(b0 + 8 * 2 * S);
  } else if (IDLE && warp > npw) {
    mbar_wait(b0 + 8 * 2 * S, 0);      // like the epilogue warps: blocked on one barrier for a whole tile
  }
  __syncthreads();
  if (threadIdx.x == 0) clk[blockIdx.x] = clock64() - t0;
}

template <int V, int S, int IDLE = 0>
static void run(int npw, int ctas_per_sm, const uint8_t* src, long long* clk) {
  const int iters = 20000;
  const int smem = ctas_per_sm == 1 ? 120 * 1024 : 60 * 1024;
  cudaFuncSetAttribute(ring<V, S, IDLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int grid = 148 * ctas_per_sm;
  for (int rep = 0; rep < 2; ++rep) ring<V, S, IDLE><<<grid, (npw + 1 + IDLE) * 32, smem>>>(iters, npw, src, clk);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  static long long h[296];
  cudaMemcpy(h, clk, grid * 8, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < grid; ++i) avg += h[i];
  avg /= grid;
  const char* names[] = {"all threads cp.async.mbarrier.arrive.noinc", "all threads mbarrier.arrive", "one lane per warp mbarrier.arrive", "1 cp.async + noinc arrive per thread"};
  printf("stages %d  producer warps %2d  parked warps %d  CTAs/SM %d  %-44s : %7.1f clk per ring slot\n", S, npw, IDLE, ctas_per_sm, names[V], avg / iters);
}

int main() {
  uint8_t* src;
  cudaMalloc(&src, 1 << 20);
  long long* clk;
  cudaMalloc(&clk, 296 * 8);
  for (int c = 1; c <= 2; ++c)
    for (int npw : {4, 8, 16}) {
      run<0, 4>(npw, c, src, clk);
      run<1, 4>(npw, c, src, clk);
      run<2, 4>(npw, c, src, clk);
      run<3, 4>(npw, c, src, clk);
    }
  run<0, 4, 4>(4, 2, src, clk);
  run<1, 4, 4>(4, 2, src, clk);
  run<0, 4, 4>(4, 1, src, clk);
  run<0, 4, 8>(4, 2, src, clk);
  run<0, 3>(8, 2, src, clk);
  run<0, 8>(8, 2, src, clk);
  run<2, 3>(8, 2, src, clk);
  run<2, 8>(8, 2, src, clk);
  return 0;
}
