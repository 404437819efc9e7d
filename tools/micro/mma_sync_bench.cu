// Throughput of legacy mma.sync.m16n8k16 (f16, fp32 accumulate) on sm_100a as a function of warps per SM and of the
// number of independent accumulators per warp.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_sync_bench mma_sync_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int NACC>
__global__ void k(float* out, int iters) {
  float d[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 9, b0 = 11, b1 = 13;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  if (s == 123.456f) out[0] = s;
}

template <int NACC>
void run(int warps_per_sm, int sms) {
  const int iters = 4096;
  float* out;
  cudaMalloc(&out, 4);
  dim3 grid(sms), block(32 * warps_per_sm);
  k<NACC><<<grid, block>>>(out, 16);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<NACC><<<grid, block>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double flops = 2.0 * 16 * 8 * 16 * (double)NACC * iters * warps_per_sm * sms;
  const double clk_per_mma_per_smsp = ms * 1e-3 * 1.965e9 / ((double)NACC * iters * warps_per_sm / 4.0);
  printf("warps/SM %2d  acc/warp %2d : %7.1f TFLOP/s   %.2f clk per MMA per SM sub-partition\n", warps_per_sm, NACC, flops / ms * 1e-9, clk_per_mma_per_smsp);
  cudaFree(out);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  for (int w : {4, 8, 16, 32}) { run<1>(w, sms); run<4>(w, sms); run<8>(w, sms); run<16>(w, sms); }
  return 0;
}
