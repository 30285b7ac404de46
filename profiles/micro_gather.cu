// profiles/micro_gather.cu -- microbenchmark behind the LCA record layout: how many bytes does B200 move
// per RANDOM 32-byte record read, depending on the load instruction and the L2 fetch granularity limit?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o micro_gather micro_gather.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 mix(u64 z) { z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }

template <int MODE>
__device__ __forceinline__ u64 load_rec(const char* p) {
  u64 a, b, c, d;
  if (MODE == 0) {  // two plain 128-bit loads
    const uint4 x = *reinterpret_cast<const uint4*>(p), y = *reinterpret_cast<const uint4*>(p + 16);
    return (u64)x.x + x.y + x.z + x.w + y.x + y.y + y.z + y.w;
  } else if (MODE == 1) {  // two 128-bit nc / no_allocate loads
    uint4 x, y;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "l"(p));
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4+16];" : "=r"(y.x), "=r"(y.y), "=r"(y.z), "=r"(y.w) : "l"(p));
    return (u64)x.x + x.y + x.z + x.w + y.x + y.y + y.z + y.w;
  } else if (MODE == 2) {  // one 256-bit load, evict_first
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    return a + b + c + d;
  } else if (MODE == 3) {  // one 256-bit load, plain
    asm volatile("ld.global.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    return a + b + c + d;
  } else if (MODE == 4) {  // one 128-bit load only (16-byte record)
    const uint4 x = *reinterpret_cast<const uint4*>(p);
    return (u64)x.x + x.y + x.z + x.w;
  } else if (MODE == 5) {  // one 64-bit load
    return *reinterpret_cast<const u64*>(p);
  } else {  // 6: two 128-bit loads, volatile-free __ldg
    const uint4 x = __ldg(reinterpret_cast<const uint4*>(p)), y = __ldg(reinterpret_cast<const uint4*>(p + 16));
    return (u64)x.x + x.y + x.z + x.w + y.x + y.y + y.z + y.w;
  }
}

template <int MODE, int K>
__global__ void __launch_bounds__(256) gather(const char* __restrict__ base, u64 nrec, u64 q, u64* __restrict__ out, u64 salt) {
  const u64 tid = blockIdx.x * (u64)blockDim.x + threadIdx.x, T = (u64)gridDim.x * blockDim.x;
  u64 acc = 0;
  for (u64 i = tid; i < q; i += T * K) {
    u64 idx[K];
#pragma unroll
    for (int j = 0; j < K; ++j) idx[j] = mix(i + j * T + salt) % nrec;
#pragma unroll
    for (int j = 0; j < K; ++j) acc += load_rec<MODE>(base + idx[j] * 32);
  }
  if (acc == 0x1234567) out[0] = acc;
}

template <int MODE>
float run(const char* base, u64 nrec, u64 q, u64* out, int grid) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    gather<MODE, 4><<<grid, 256>>>(base, nrec, q, out, 77 + rep);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
  }
  return best;
}

int main(int argc, char** argv) {
  const u64 nrec = 20000000ull, q = 200000000ull;
  char* base; u64* out;
  cudaMalloc(&base, nrec * 32); cudaMemset(base, 1, nrec * 32); cudaMalloc(&out, 8);
  const char* names[] = {"2x128 plain", "2x128 nc.no_allocate", "1x256 nc.no_alloc.evict_first", "1x256 plain", "1x128 (16B of the record)", "1x64", "2x128 __ldg"};
  for (int gran = 0; gran < 3; ++gran) {
    size_t g = 0;
    if (gran == 1) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
    if (gran == 2) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 128);
    cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
    for (int grid : {148 * 8, 148 * 16}) {
      float t[7] = {run<0>(base, nrec, q, out, grid), run<1>(base, nrec, q, out, grid), run<2>(base, nrec, q, out, grid), run<3>(base, nrec, q, out, grid),
                    run<4>(base, nrec, q, out, grid), run<5>(base, nrec, q, out, grid), run<6>(base, nrec, q, out, grid)};
      for (int m = 0; m < 7; ++m)
        printf("L2fetch=%zu grid=%d %-32s %.3f ms  %.1f G rec/s  (%.0f GB/s at 32 B/rec)\n", g, grid, names[m], t[m], q / t[m] / 1e6, q * 32.0 / t[m] / 1e6);
    }
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
