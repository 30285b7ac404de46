// csrc/cuda_common.cuh -- shared CUDA helpers for libptgpu.so (sm_100a only)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/pt_gpu.h"
#include "pt_error.h"

#define PT_CUDA(call)                                                                                              \
  do {                                                                                                             \
    cudaError_t err__ = (call);                                                                                    \
    if (err__ != cudaSuccess)                                                                                      \
      return ptgpu::set_error(err__ == cudaErrorMemoryAllocation ? PT_ERR_NOMEM : PT_ERR_CUDA, "%s:%d: %s: %s", __FILE__, __LINE__, #call, \
                              cudaGetErrorString(err__));                                                          \
  } while (0)

namespace ptgpu {

// B200: 148 SMs, 227 KB opt-in shared memory per CTA (queried at run time, these are only the design numbers)
constexpr int kMaxSmemPerBlock = 227 * 1024;

inline int require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return set_error(PT_ERR_NO_DEVICE, "no CUDA device available (%s); libptgpu has no CPU path", e == cudaSuccess ? "count is 0" : cudaGetErrorString(e));
  }
  return PT_OK;
}

struct DeviceProps { int sms; int smem_optin; int device; };
inline int device_props(DeviceProps* p) {
  PT_CUDA(cudaGetDevice(&p->device));
  PT_CUDA(cudaDeviceGetAttribute(&p->sms, cudaDevAttrMultiProcessorCount, p->device));
  PT_CUDA(cudaDeviceGetAttribute(&p->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, p->device));
  return PT_OK;
}

// Caching device allocator: pt_batch_create / pt_batch_free sit on the end-to-end path (one handle per batch), and
// cudaMalloc / cudaFree of the ~1 GB a 10^5-instance batch needs cost more than the kernels.  Freed blocks are kept
// (up to kCacheLimit bytes) and handed out again for requests of the same size class or up to a quarter smaller.  Every entry
// point allocates through it: in a process that already holds a few GB one raw cudaMalloc / cudaFree pair costs 10-20 ms
// (pt_tree_who_displays: 151 ms with raw calls, 5 ms with cached blocks, same kernels).
int dev_alloc(void** p, size_t bytes);
void dev_free(void* p);

// streaming loads/stores: data touched once should not displace the L2-resident tables
__device__ __forceinline__ int4 ld_stream_v4(const int4* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ int ld_stream(const int* p) {
  int r;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(int* p, int v) { asm volatile("st.global.cs.s32 [%0], %1;" ::"l"(p), "r"(v)); }
__device__ __forceinline__ void st_stream_v4(int4* p, int4 v) { asm volatile("st.global.cs.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)); }

// L2 eviction policies (sm_100a accepts the .L2::evict_* qualifier directly only on 256-bit loads; narrower loads
// take a policy register through .L2::cache_hint)
__device__ __forceinline__ unsigned long long l2_policy_evict_last() { unsigned long long p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ unsigned long long ld_u64_hint(const unsigned long long* p, unsigned long long pol) {
  unsigned long long r;
  asm volatile("ld.global.nc.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(r) : "l"(p), "l"(pol));
  return r;
}
// one 32-byte sector in one instruction (LDG.E.256), not kept in L1, first to leave L2
__device__ __forceinline__ void ld_sector_evict_first(const void* p, unsigned long long& a, unsigned long long& b, unsigned long long& c, unsigned long long& d) {
  asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
}

}  // namespace ptgpu
