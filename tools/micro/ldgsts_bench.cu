// Microbenchmark: throughput of 16-byte cp.async (LDGSTS) row gathers into shared memory as a
// function of row size, lane mapping, cache operator and DESTINATION layout (development tool;
// informs the A-operand layout of the sparse-conv kernel).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/lb ldgsts_bench.cu && /tmp/lb
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <bool CA>
__device__ __forceinline__ void cp16(uint32_t dst, const void* src) {
  if (CA) asm volatile("cp.async.ca.shared.global [%0], [%1], 16, 16;" ::"r"(dst), "l"(src) : "memory");
  else asm volatile("cp.async.cg.shared.global [%0], [%1], 16, 16;" ::"r"(dst), "l"(src) : "memory");
}
template <bool CA>
__device__ __forceinline__ void cp16z(uint32_t dst, const void* src, uint32_t nbytes) {
  if (CA) asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
  else asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
}
__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

constexpr int U = 8;            // instructions per warp per commit group
constexpr int WARP_SMEM = 12 * 1024;

// RB: row bytes.  TPR: thread-per-row mapping (each lane copies a whole row with RB/16 instructions) vs
// coalesced (RB/16 consecutive lanes copy one row).  SWZ: destination = row-contiguous with the UMMA
// 32/64/128-byte XOR swizzle (256 B rows = two 128 B-swizzled K blocks) vs no-swizzle core-matrix planes
// (chunk c of row r at c*(R*16+16) + r*16).
// MISS: 0 = every row present; 1 = ~30 % of the rows missing, zero-filled through the src-size operand (src = table base);
// 2 = missing rows read a (cache-resident) all-zero row with a plain 16-byte copy; 3 = src-size form, nothing missing
template <bool CA, int RB, bool TPR, bool SWZ, int MISS>
__global__ void __launch_bounds__(512, 1) bench(const uint8_t* table, uint32_t row_mask, int iters, long long* clk) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = smem_u32(smem) + warp * WARP_SMEM;
  constexpr int CH = RB / 16;
  constexpr int RPI = TPR ? 32 : 32 / CH;                  // rows touched per instruction
  constexpr int R = TPR ? 32 * (U / CH > 0 ? U / CH : 1) : RPI * U;   // rows per group
  constexpr int PLANE = R * 16 + 16;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < U; ++j) {
      int r, c;
      if (TPR) { r = (j / CH) * 32 + lane; c = j % CH; }
      else { r = j * RPI + lane / CH; c = lane % CH; }
      uint32_t row = hash32(it * 131u + r * 7919u + (blockIdx.x * 16 + warp) * 104729u) & row_mask;
      if (MISS == 4) {   // neighbour walk: consecutive rows, every row requested by three consecutive instructions (dx = -1, 0, +1)
        const int grp = j / 3, dx = j % 3;
        row = ((uint32_t)((blockIdx.x * 16 + warp) * 4099 + it * (U / 3) * RPI + grp * RPI + (TPR ? lane : lane / CH) + dx)) & row_mask;
      }
      const uint8_t* src = table + (size_t)row * RB + c * 16;
      const bool missing = MISS && MISS != 3 && (hash32(row * 31u + it) % 10u) < 3u;
      uint32_t dst;
      if (!SWZ) dst = base + c * PLANE + r * 16;
      else if (RB == 256) dst = base + (c >> 3) * (R * 128) + r * 128 + (((c & 7) ^ (r & 7)) * 16);
      else if (RB == 128) dst = base + r * 128 + ((c ^ (r & 7)) * 16);
      else if (RB == 64) dst = base + r * 64 + ((c ^ ((r >> 1) & 3)) * 16);
      else dst = base + r * 32 + ((c ^ ((r >> 2) & 1)) * 16);
      if (MISS == 0 || MISS == 4) cp16<CA>(dst, src);
      else if (MISS == 1) cp16z<CA>(dst, missing ? table : src, missing ? 0u : 16u);
      else if (MISS == 2) cp16<CA>(dst, missing ? table + c * 16 : src);
      else cp16z<CA>(dst, src, 16u + (row >> 31));
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 3;" ::: "memory");
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

static uint8_t* table;
static long long* clk;

template <bool CA, int RB, bool TPR, bool SWZ, int MISS = 0>
static void run(size_t bytes) {
  const int iters = 1000;
  auto k = bench<CA, RB, TPR, SWZ, MISS>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * WARP_SMEM);
  const uint32_t mask = (uint32_t)(bytes / RB) - 1;
  for (int rep = 0; rep < 2; ++rep) k<<<148, 512, 16 * WARP_SMEM>>>(table, mask, iters, clk);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); exit(1); }
  long long h[148];
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  avg /= 148;
  const double instr = 16.0 * U * iters;
  printf("miss-mode %d %s rows %3d B  %-10s -> %-8s  table %7zu KB : clk/instr %6.2f   B/clk/SM %6.1f\n", MISS, CA ? ".ca" : ".cg", RB,
         TPR ? "thread/row" : "coalesced", SWZ ? "swizzled" : "planes", bytes >> 10, avg / instr, instr * 512 / avg);
}

template <bool CA, int RB>
static void sweep(size_t bytes) {
  run<CA, RB, false, false>(bytes);
  run<CA, RB, false, true>(bytes);
  run<CA, RB, true, false>(bytes);
  run<CA, RB, true, true>(bytes);
}

int main(int argc, char** argv) {
  cudaMalloc(&table, 512ull << 20);
  cudaMemset(table, 1, 512ull << 20);
  cudaMalloc(&clk, 148 * 8);
  if (argc > 2) {   // neighbour-walk pattern vs random rows (L2-resident 16 MB table)
    const size_t bytes = (size_t)16 << 20;
    run<true, 64, false, false, 0>(bytes); run<true, 64, false, false, 4>(bytes);
    run<true, 128, false, false, 0>(bytes); run<true, 128, false, false, 4>(bytes);
    run<true, 256, false, false, 0>(bytes); run<true, 256, false, false, 4>(bytes);
    run<false, 128, false, true, 0>(bytes); run<false, 128, false, true, 4>(bytes);
    return 0;
  }
  if (argc > 1) {   // missing-row handling
    for (size_t bytes : {(size_t)32 << 10, (size_t)16 << 20}) {
      run<true, 128, false, false, 0>(bytes); run<true, 128, false, false, 3>(bytes); run<true, 128, false, false, 1>(bytes); run<true, 128, false, false, 2>(bytes);
      run<false, 128, false, true, 0>(bytes); run<false, 128, false, true, 3>(bytes); run<false, 128, false, true, 1>(bytes); run<false, 128, false, true, 2>(bytes);
      run<true, 64, false, false, 0>(bytes); run<true, 64, false, false, 1>(bytes); run<true, 64, false, false, 2>(bytes);
      run<true, 64, true, false, 0>(bytes); run<true, 64, true, false, 1>(bytes); run<true, 64, true, false, 2>(bytes);
      run<false, 64, false, true, 0>(bytes); run<false, 64, false, true, 1>(bytes); run<false, 64, false, true, 2>(bytes);
    }
    return 0;
  }
  for (size_t bytes : {(size_t)32 << 10, (size_t)16 << 20, (size_t)512 << 20}) {
    sweep<true, 32>(bytes); sweep<true, 64>(bytes); sweep<true, 128>(bytes); sweep<true, 256>(bytes);
    sweep<false, 32>(bytes); sweep<false, 64>(bytes); sweep<false, 128>(bytes); sweep<false, 256>(bytes);
  }
  return 0;
}
