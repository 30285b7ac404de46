// csrc/tc_kernels.cu -- batched tree containment on sm_100a: one warp per (sub)instance, instance state in shared
// memory, persistent warps pulling work items from an atomic ticket; branching is level-synchronous over a pool of
// compacted child instances (one kernel launch per branching depth).
//
// Entry points: pt_batch_create / pt_batch_contain / pt_batch_free (include/pt_gpu.h), standing in for
// TreeInNetContainment(...).displayed() (utils/containment.hpp:1179-1202) and TreeInTreeContainment(...).displayed() (:50-83).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <vector>
#include "cuda_common.cuh"
#include "tc_core.cuh"

#include <cooperative_groups.h>
#include <chrono>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>

namespace ptgpu {

// ---- caching device allocator (see cuda_common.cuh) -----------------------------------------------------------------
namespace {
std::mutex g_alloc_mu;
std::multimap<size_t, void*> g_free_blocks;          // (device << 48 | size) -> block
std::unordered_map<void*, size_t> g_block_size;      // every live or cached block we own -> its key
size_t g_cached_bytes = 0;
constexpr size_t kCacheLimit = (size_t)64 << 30;  // of 180 GB: beyond it blocks go back to the driver, and cudaFree costs ~20 ms in a process holding this much
size_t size_class(size_t b) { const size_t g = b < ((size_t)1 << 20) ? 4096 : ((size_t)1 << 20); return (b + g - 1) / g * g; }
}  // namespace
int dev_alloc(void** p, size_t bytes) {
  const size_t sz = size_class(bytes ? bytes : 1);
  int dev = 0;
  cudaGetDevice(&dev);
  const size_t key = ((size_t)dev << 48) | sz;
  {
    std::lock_guard<std::mutex> lk(g_alloc_mu);
    // smallest cached block of this device that fits and is at most a quarter larger (the block keeps its own size class)
    auto it = g_free_blocks.lower_bound(key);
    if (it != g_free_blocks.end() && (it->first >> 48) == (size_t)dev) {
      const size_t have = it->first & (((size_t)1 << 48) - 1);
      if (have <= sz + sz / 4) { *p = it->second; g_free_blocks.erase(it); g_cached_bytes -= have; return PT_OK; }
    }
  }
  cudaError_t e = cudaMalloc(p, sz);
  if (e != cudaSuccess) {  // release the cache and retry once
    cudaGetLastError();
    {
      std::lock_guard<std::mutex> lk(g_alloc_mu);
      for (auto& kv : g_free_blocks) { cudaFree(kv.second); g_block_size.erase(kv.second); }
      g_free_blocks.clear(); g_cached_bytes = 0;
    }
    e = cudaMalloc(p, sz);
    if (e != cudaSuccess) { cudaGetLastError(); *p = nullptr; return set_error(PT_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", sz, cudaGetErrorString(e)); }
  }
  std::lock_guard<std::mutex> lk(g_alloc_mu);
  g_block_size[*p] = key;
  return PT_OK;
}
void dev_free(void* p) {
  if (!p) return;
  std::lock_guard<std::mutex> lk(g_alloc_mu);
  auto it = g_block_size.find(p);
  if (it == g_block_size.end()) { cudaFree(p); return; }
  const size_t sz = it->second & (((size_t)1 << 48) - 1);
  if (g_cached_bytes + sz > kCacheLimit) { cudaFree(p); g_block_size.erase(it); return; }
  g_free_blocks.emplace(it->second, p); g_cached_bytes += sz;
}

// ---- the warp executor: PT_FOR = lane-strided loop, barrier = __syncwarp, atomics on shared memory ---------------
#ifdef PT_PROFILE
// profiling build (make prof -> libptgpu_prof.so, never shipped): cycles per solver pass, summed over warps
__device__ unsigned long long g_prof[2][64];  // [0] cycles  [1] entries
struct PhaseProbe {
  int cur; long long t0;
  __device__ __forceinline__ void operator=(int p) {
    const long long t = clock64();
    if ((threadIdx.x & 31) == 0) { atomicAdd(&g_prof[0][cur & 63], (unsigned long long)(t - t0)); atomicAdd(&g_prof[1][p & 63], 1ull); }
    cur = p; t0 = clock64();
  }
};
#define PT_PHASE_INIT {0, clock64()}
#else
typedef int PhaseProbe;
#define PT_PHASE_INIT 0
#endif
struct WarpEx {
  int lane;
  PhaseProbe phase;  // last pass entered (debug aid, see CtaEx)
  __device__ __forceinline__ int tid() const { return lane; }
  __device__ __forceinline__ int nth() const { return 32; }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
  __device__ __forceinline__ int map(int i, int) const { return i; }
  // shared memory: between the two barriers nobody writes, so every lane reads the same value (one broadcast load)
  __device__ __forceinline__ int uniform(const int32_t* p) const { __syncwarp(); const int v = *(const volatile int32_t*)p; __syncwarp(); return v; }
  __device__ __forceinline__ int reduce_or(int v) const { return (int)__reduce_or_sync(0xffffffffu, (unsigned)v); }
  __device__ __forceinline__ int reduce_add(int v) const { return __reduce_add_sync(0xffffffffu, v); }
  __device__ __forceinline__ int atomic_add(int32_t* p, int v) const { return atomicAdd(p, v); }
  __device__ __forceinline__ int atomic_min(int32_t* p, int v) const { return atomicMin(p, v); }
  __device__ __forceinline__ int atomic_max(int32_t* p, int v) const { return atomicMax(p, v); }
  __device__ __forceinline__ uint32_t atomic_or(uint32_t* p, uint32_t v) const { return atomicOr(p, v); }
  __device__ __forceinline__ int atomic_xor(int32_t* p, int v) const { return atomicXor(p, v); }
  // L2-coherent (ld.global.cg), not the read-only path: branch children are written by other SMs during the same launch, and
  // the chunks of a HOST batch land (copy engine) while the kernel is already running
  template <class S>
  __device__ __forceinline__ int ldg(const S* p) const { return (int)__ldcg(p); }
  __device__ __forceinline__ int atomic_cas(int32_t* p, int c, int v) const { return atomicCAS(p, c, v); }
  __device__ __forceinline__ int atomic_cas_idx(int16_t* p, int c, int v) const {
    return (int)(int16_t)atomicCAS((unsigned short*)p, (unsigned short)(int16_t)c, (unsigned short)(int16_t)v);
  }
  // 16-bit counters, native 32-bit shared atomics on the word that holds them.  The counters are non-negative and never
  // leave [0, 32767], so an add / sub never carries into the neighbouring half.
  __device__ __forceinline__ int atomic_add_idx(int16_t* p, int v) const {
    const unsigned odd = (unsigned)__cvta_generic_to_shared(p) & 2u, sh = odd << 3;
    unsigned* wp = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(p) - odd);
    const unsigned old = v >= 0 ? atomicAdd(wp, (unsigned)v << sh) : atomicSub(wp, (unsigned)(-v) << sh);
    return (int)(int16_t)(old >> sh);
  }
  __device__ __forceinline__ void atomic_xor_idx(int16_t* p, int v) const {
    const unsigned odd = (unsigned)__cvta_generic_to_shared(p) & 2u, sh = odd << 3;
    atomicXor(reinterpret_cast<unsigned*>(reinterpret_cast<char*>(p) - odd), (unsigned)v << sh);
  }
  // stable in-place compaction of the arc arrays (keeps at >= 0); returns the new count.
  // A tile's loads complete before its ballot does and its stores go to positions <= the ones just read.
  template <class I>
  __device__ __forceinline__ int compact_arcs(I* at, I* ah, int m, I*, I*) const {
    int carry = 0;
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll 1
    for (int base = 0; base < m; base += 32) {
      const int i = base + lane;
      int t = -1, h = 0;
      if (i < m) { t = at[i]; h = ah[i]; }
      const unsigned mask = __ballot_sync(0xffffffffu, t >= 0);
      if (t >= 0) { const int pos = carry + __popc(mask & lt); at[pos] = (I)t; ah[pos] = (I)h; }
      carry += __popc(mask);
    }
    __syncwarp();
    return carry;
  }
  // out[0..n] = exclusive prefix sums of in[0..n): see warp_exclusive_scan below (one shared, out-of-line copy: inlined at its
  // five call sites the scan was 46 KB of SASS, and instruction fetch was the kernel's first stall reason)
  template <class C, class T>
  __device__ __forceinline__ void exclusive_scan(const C* in, T* out, int n) const;
};

// tiles of 32 consecutive elements (bank-conflict free), one shuffle scan per tile with a running carry.  (The first version
// gave every lane a contiguous chunk: 8-way bank conflicts, 7 % of all stall samples of the kernel.)  Takes shared-window
// addresses so that the out-of-line copy still uses LDS / STS.
template <int IN_BYTES, int OUT_BYTES>
__device__ __noinline__ void warp_exclusive_scan(unsigned in_s, unsigned out_s, int n) {
  const int lane = threadIdx.x & 31;
  int carry = 0;
#pragma unroll 1
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    int v = 0;
    if (i < n) {
      if (IN_BYTES == 2) { short h; asm volatile("ld.shared.s16 %0, [%1];" : "=h"(h) : "r"(in_s + 2u * (unsigned)i) : "memory"); v = h; }
      else asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(in_s + 4u * (unsigned)i) : "memory");
    }
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    const int o = carry + incl - v;
    if (i < n) {
      if (OUT_BYTES == 2) asm volatile("st.shared.b16 [%0], %1;" ::"r"(out_s + 2u * (unsigned)i), "h"((short)o) : "memory");
      else asm volatile("st.shared.b32 [%0], %1;" ::"r"(out_s + 4u * (unsigned)i), "r"(o) : "memory");
    }
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) {
    if (OUT_BYTES == 2) asm volatile("st.shared.b16 [%0], %1;" ::"r"(out_s + 2u * (unsigned)n), "h"((short)carry) : "memory");
    else asm volatile("st.shared.b32 [%0], %1;" ::"r"(out_s + 4u * (unsigned)n), "r"(carry) : "memory");
  }
  __syncwarp();
}
template <class C, class T>
__device__ __forceinline__ void WarpEx::exclusive_scan(const C* in, T* out, int n) const {
  static_assert((sizeof(T) == 2 || sizeof(T) == 4) && (sizeof(C) == 2 || sizeof(C) == 4), "index type");
  __syncwarp();
  warp_exclusive_scan<(int)sizeof(C), (int)sizeof(T)>((unsigned)__cvta_generic_to_shared(in), (unsigned)__cvta_generic_to_shared(out), n);
}

// ---- the CTA executor: one thread block per instance, workspace in GLOBAL memory (int32 ids), barrier = __syncthreads ----
// For instances that do not fit one CTA's shared memory (BASELINE config 4: n = 10^6).  Same solver code, other executor.
// debug trace (host-mapped memory, see pt_debug_trace): first and last thread of block 0 publish how many barriers they have
// passed and which pass they are in, so that a hang can be located from the host while the kernel is still stuck
__device__ volatile int* g_trace = nullptr;
struct CtaEx {
  int* red;  // 40 ints of shared memory
  int phase;
  int nsync;
  __device__ __forceinline__ int tid() const { return threadIdx.x; }
  __device__ __forceinline__ int nth() const { return blockDim.x; }
  __device__ __forceinline__ void sync() {
    ++nsync;
    volatile int* t = g_trace;
    if (t && (threadIdx.x == 0 || threadIdx.x == blockDim.x - 1)) { const int o = blockIdx.x * 16 + (threadIdx.x ? 4 : 0); t[o] = nsync; t[o + 1] = phase; }
    // The workspace is in GLOBAL memory and the passes mix plain stores with atomics (which execute in L2) on the same
    // words.  A CTA barrier alone orders them only as far as the SM's L1; the fence pushes this thread's stores to L2 first.
    __threadfence();
    __syncthreads();
  }
  __device__ __forceinline__ int map(int i, int) const { return i; }
  __device__ __forceinline__ int uniform(const int32_t* p) {
    sync();
    if (threadIdx.x == 0) red[34] = *(const volatile int32_t*)p;
    sync();
    const int v = red[34];
    sync();
    return v;
  }
  __device__ __forceinline__ int reduce_or(int v) {
    v = (int)__reduce_or_sync(0xffffffffu, (unsigned)v);
    sync();
    if (threadIdx.x == 0) red[32] = 0;
    sync();
    if ((threadIdx.x & 31) == 0 && v) atomicOr(&red[32], v);
    sync();
    return red[32];
  }
  __device__ __forceinline__ int reduce_add(int v) {
    v = __reduce_add_sync(0xffffffffu, v);
    sync();
    if (threadIdx.x == 0) red[33] = 0;
    sync();
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&red[33], v);
    sync();
    return red[33];
  }
  __device__ __forceinline__ int atomic_add(int32_t* p, int v) const { return atomicAdd(p, v); }
  __device__ __forceinline__ int atomic_min(int32_t* p, int v) const { return atomicMin(p, v); }
  __device__ __forceinline__ int atomic_max(int32_t* p, int v) const { return atomicMax(p, v); }
  __device__ __forceinline__ uint32_t atomic_or(uint32_t* p, uint32_t v) const { return atomicOr(p, v); }
  __device__ __forceinline__ int atomic_xor(int32_t* p, int v) const { return atomicXor(p, v); }
  template <class S>
  __device__ __forceinline__ int ldg(const S* p) const { return (int)__ldg(p); }  // caller's arrays: read-only during the launch
  __device__ __forceinline__ int atomic_cas(int32_t* p, int c, int v) const { return atomicCAS(p, c, v); }
  __device__ __forceinline__ int atomic_cas_idx(int32_t* p, int c, int v) const { return atomicCAS(p, c, v); }
  __device__ __forceinline__ int atomic_add_idx(int32_t* p, int v) const { return atomicAdd(p, v); }
  __device__ __forceinline__ void atomic_xor_idx(int32_t* p, int v) const { atomicXor(p, v); }
  template <class I>
  __device__ __forceinline__ int compact_arcs(I* at, I* ah, int m, I*, I*) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    sync();
    int carry = 0;
    for (int base = 0; base < m; base += blockDim.x) {
      const int i = base + threadIdx.x;
      int t = -1, h = 0;
      if (i < m) { t = at[i]; h = ah[i]; }
      const unsigned mask = __ballot_sync(0xffffffffu, t >= 0);
      if (lane == 0) red[wid] = __popc(mask);
      __syncthreads();  // every thread has read its element of this tile
      int woff = 0, total = 0;
      for (int k = 0; k < nw; ++k) { const int c = red[k]; if (k < wid) woff += c; total += c; }
      if (t >= 0) { const int pos = carry + woff + __popc(mask & lt); at[pos] = (I)t; ah[pos] = (I)h; }
      carry += total;
      __syncthreads();
    }
    sync();
    return carry;
  }
  // tile-by-tile block scan with a running carry: coalesced, two barriers per tile of blockDim.x elements
  template <class C, class T>
  __device__ __forceinline__ void exclusive_scan(const C* in, T* out, int n) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    sync();
    int carry = 0;
    for (int base = 0; base < n; base += blockDim.x) {
      const int i = base + threadIdx.x;
      const int v = i < n ? in[i] : 0;
      int incl = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
      if (lane == 31) red[wid] = incl;
      __syncthreads();
      int woff = 0, total = 0;
      for (int k = 0; k < nw; ++k) { const int t = red[k]; if (k < wid) woff += t; total += t; }
      if (i < n) out[i] = (T)(carry + woff + incl - v);
      carry += total;
      __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = (T)carry;
    __syncthreads();
  }
};

#ifndef PT_CLUSTER_THREADS
#define PT_CLUSTER_THREADS 1024  // config 4, clusters of 8: 512 threads (115 registers) 28.9 ms, 1024 threads (64 registers) 24.4 ms
#endif
// ---- the cluster executor: one thread-block CLUSTER per instance, workspace in global memory -------------------------------
// For a handful of huge instances (BASELINE config 4: 8 instances of n = 10^6 per GPU) one CTA per instance leaves most SMs
// idle.  A cluster of up to 8 CTAs (hardware co-scheduled, hardware barrier) runs the same solver: a pass is spread over
// all threads of the cluster, the barrier is barrier.cluster, reductions and scans combine per-CTA partials through a few
// words of the workspace.  gsc = 64 ints at the front of the cluster's workspace: [0] reduce_or [1] reduce_add [2] ticket /
// broadcast [16..31] per-CTA partial sums.
struct ClusterEx {
  int* red;           // 40 ints of shared memory (this CTA)
  int32_t* gsc;       // 64 ints of global memory (this cluster)
  int phase;
  unsigned rank, csize;
  __device__ __forceinline__ int tid() const { return (int)(rank * blockDim.x + threadIdx.x); }
  __device__ __forceinline__ int nth() const { return (int)(csize * blockDim.x); }
  __device__ __forceinline__ void sync() {
#ifndef PT_CLUSTER_NO_FENCE
    __threadfence();  // as in CtaEx: plain stores must reach L2 before another CTA's atomics / loads of the same words
#endif
    cooperative_groups::this_cluster().sync();  // barrier.cluster.arrive.release + wait.acquire
  }
  __device__ __forceinline__ int map(int i, int) const { return i; }
  __device__ __forceinline__ int uniform(const int32_t* p) { sync(); const int v = *(const volatile int32_t*)p; sync(); return v; }
  __device__ __forceinline__ int reduce_or(int v) {
    v = (int)__reduce_or_sync(0xffffffffu, (unsigned)v);
    sync();
    if (tid() == 0) gsc[0] = 0;
    sync();
    if ((threadIdx.x & 31) == 0 && v) atomicOr(&gsc[0], v);
    sync();
    return *(volatile int32_t*)&gsc[0];
  }
  __device__ __forceinline__ int reduce_add(int v) {
    v = __reduce_add_sync(0xffffffffu, v);
    sync();
    if (tid() == 0) gsc[1] = 0;
    sync();
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&gsc[1], v);
    sync();
    return *(volatile int32_t*)&gsc[1];
  }
  __device__ __forceinline__ int atomic_add(int32_t* p, int v) const { return atomicAdd(p, v); }
  __device__ __forceinline__ int atomic_min(int32_t* p, int v) const { return atomicMin(p, v); }
  __device__ __forceinline__ int atomic_max(int32_t* p, int v) const { return atomicMax(p, v); }
  __device__ __forceinline__ uint32_t atomic_or(uint32_t* p, uint32_t v) const { return atomicOr(p, v); }
  __device__ __forceinline__ int atomic_xor(int32_t* p, int v) const { return atomicXor(p, v); }
  template <class S>
  __device__ __forceinline__ int ldg(const S* p) const { return (int)__ldg(p); }
  __device__ __forceinline__ int atomic_cas(int32_t* p, int c, int v) const { return atomicCAS(p, c, v); }
  __device__ __forceinline__ int atomic_cas_idx(int32_t* p, int c, int v) const { return atomicCAS(p, c, v); }
  __device__ __forceinline__ int atomic_add_idx(int32_t* p, int v) const { return atomicAdd(p, v); }
  __device__ __forceinline__ void atomic_xor_idx(int32_t* p, int v) const { atomicXor(p, v); }
  // sum of `mine` over the threads of this CTA (all threads get it)
  __device__ __forceinline__ int cta_sum(int mine) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    mine = __reduce_add_sync(0xffffffffu, mine);
    __syncthreads();
    if (lane == 0) red[wid] = mine;
    __syncthreads();
    int total = 0;
    for (int k = 0; k < nw; ++k) total += red[k];
    return total;
  }
  // every CTA publishes its partial; returns the sum over the lower ranks and, in *all, over the whole cluster
  __device__ __forceinline__ int cluster_offsets(int cta_total, int* all) {
    sync();
    if (threadIdx.x == 0) gsc[16 + rank] = cta_total;
    sync();
    int base = 0, sum = 0;
    for (unsigned k = 0; k < csize; ++k) { const int t = *(volatile int32_t*)&gsc[16 + k]; if (k < rank) base += t; sum += t; }
    *all = sum;
    return base;
  }
  // each CTA owns a contiguous slice of [0,n): a counting pass, the cluster-wide offsets, then the tile scan of CtaEx
  template <class C, class T>
  __device__ __forceinline__ void exclusive_scan(const C* in, T* out, int n) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int lo = (int)((long long)n * rank / csize), hi = (int)((long long)n * (rank + 1) / csize);
    int part = 0;
    sync();
    for (int i = lo + (int)threadIdx.x; i < hi; i += blockDim.x) part += in[i];
    int all = 0;
    int carry = cluster_offsets(cta_sum(part), &all);
    for (int base = lo; base < hi; base += blockDim.x) {
      const int i = base + threadIdx.x;
      const int v = i < hi ? (int)in[i] : 0;
      int incl = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
      __syncthreads();
      if (lane == 31) red[wid] = incl;
      __syncthreads();
      int woff = 0, total = 0;
      for (int k = 0; k < nw; ++k) { const int t = red[k]; if (k < wid) woff += t; total += t; }
      if (i < hi) out[i] = (T)(carry + woff + incl - v);
      carry += total;
    }
    if (tid() == 0) out[n] = (T)all;
    sync();
  }
  // stable compaction of the arc arrays (keeps at >= 0) into the two spare arrays, then the pointers are swapped: CTAs cannot
  // compact in place (a CTA would overwrite elements a lower-ranked one has not read yet)
  template <class I>
  __device__ __forceinline__ int compact_arcs(I*& at, I*& ah, int m, I*& spare_a, I*& spare_b) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const int lo = (int)((long long)m * rank / csize), hi = (int)((long long)m * (rank + 1) / csize);
    int part = 0;
    sync();
    for (int i = lo + (int)threadIdx.x; i < hi; i += blockDim.x) part += at[i] >= 0;
    int all = 0;
    int carry = cluster_offsets(cta_sum(part), &all);
    for (int base = lo; base < hi; base += blockDim.x) {
      const int i = base + threadIdx.x;
      int t = -1, h = 0;
      if (i < hi) { t = at[i]; h = ah[i]; }
      const unsigned mask = __ballot_sync(0xffffffffu, t >= 0);
      __syncthreads();
      if (lane == 0) red[wid] = __popc(mask);
      __syncthreads();
      int woff = 0, total = 0;
      for (int k = 0; k < nw; ++k) { const int c = red[k]; if (k < wid) woff += c; total += c; }
      if (t >= 0) { const int pos = carry + woff + __popc(mask & lt); spare_a[pos] = (I)t; spare_b[pos] = (I)h; }
      carry += total;
    }
    sync();
    I* x = at; at = spare_a; spare_a = x;
    x = ah; ah = spare_b; spare_b = x;
    return all;
  }
};

// where work items come from: the caller's batch (offset tables) or a pool of fixed-stride slots
struct TcSource {
  int64_t first, count;     // batch mode: this launch covers instances [first, count)
  const int32_t* count_ptr; // pool mode: device counter (clamped to capacity)
  int is_pool, capacity;
  const int64_t *hnode_off, *harc_off, *gnode_off;
  const void *succ_off, *succ_adj, *hlabel, *tparent, *tlabel;
  const int32_t *dims, *origin;  // pool mode
  int sN1, sM, sN, sNT;          // pool strides (elements)
  int solver_flags;              // bit 0: never finish by enumerating the small core (pt_options.no_small_core)
};
struct TcSink {
  int32_t *succ_off, *succ_adj, *hlabel, *tparent, *tlabel, *dims, *origin;
  int32_t* count;  // atomic
  int capacity, sN1, sM, sN, sNT;
};
enum { CNT_DISPLAYED = 0, CNT_ITEMS, CNT_BRANCH_ITEMS, CNT_CHERRIES, CNT_RET, CNT_ARCS, CNT_PASSES, CNT_ERRORS, CNT_N };

// instance `item` of the caller's batch (element type S) as a TcInst
template <class S>
__device__ __forceinline__ void batch_inst(const TcSource& src, int item, TcInst& in) {
  const int sh = sizeof(S) == 1 ? 0 : (sizeof(S) == 2 ? 1 : 2);  // log2 of the element size
  const int64_t hb = src.hnode_off[item], ab = src.harc_off[item], gb = src.gnode_off[item];
  const int64_t nn = src.hnode_off[item + 1] - hb, mm = src.harc_off[item + 1] - ab, tt = src.gnode_off[item + 1] - gb;
  in.n = (int)min(nn, (int64_t)1 << 30); in.m = (int)min(mm, (int64_t)1 << 30); in.nt = (int)min(tt, (int64_t)1 << 30);
  // the offset tables are validated HERE, per instance (reading 2.4 MB of tables on one host core before the launch cost 0.14 ms per
  // 10^5 instances): every range must lie inside the arrays, whose lengths are the last entries of the tables
  const int64_t B = src.count;
  if (nn < 0 || mm < 0 || tt < 0 || hb < 0 || ab < 0 || gb < 0 || hb + nn > src.hnode_off[B] || ab + mm > src.harc_off[B] || gb + tt > src.gnode_off[B]) { in.n = in.m = in.nt = -1; }
  // the degree layout (uint8_t) has n entries per instance in succ_off, the offset layouts n + 1
  in.succ_off = static_cast<const char*>(src.succ_off) + ((hb + (TcLayout<S>::kDegrees ? 0 : item)) << sh); in.succ_adj = static_cast<const char*>(src.succ_adj) + (ab << sh);
  in.hlabel = static_cast<const char*>(src.hlabel) + (hb << sh); in.tparent = static_cast<const char*>(src.tparent) + (gb << sh);
  in.tlabel = static_cast<const char*>(src.tlabel) + (gb << sh);
}

// ---- the persistent warp kernel: ONE launch per pt_batch_contain ------------------------------------------------------------
// Work items are the batch's instances (ticket) and the branch children other warps have pushed into the pool during the same
// launch.  A branching warp pushes all children but the first as compact instances (same element type S as the batch) and goes
// on with the first one IN PLACE (keep_only: no emit / reload); idle warps take pool items before fresh instances, so deep
// branching chains start early instead of forming the kernel's tail.  Termination: `finished` counts completed items; a warp
// with nothing to take leaves when finished == B + pool_tail (read in that order: both only grow, and every push happens
// before its parent finishes).  HOST batches are uploaded in chunks on a copy stream that writes chunk_ready[c] after chunk c;
// a warp whose ticket lies in a chunk that has not landed yet waits for the flag, so the solve overlaps the upload inside one
// launch.  Every wait is bounded by a watchdog (ctl->abort), so a lost flag becomes an error instead of a hang.
struct TcCtl {
  int32_t ticket, pool_tail, pool_head, finished, abort, overflow, done, pad;
  unsigned long long t_first, t_batch_done_inv, t_last;  // %globaltimer (ns): first warp in, ~(first warp that found the ticket exhausted) -- kept inverted so
                                                          // that the ONE memset that zeroes the control block initialises it (atomicMax of ~t = min of t) --, last warp out
};
__device__ __forceinline__ unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
struct TcPool {   // fixed-stride slots of element type S (see TcLayout); ready[slot] == epoch once the slot is complete
  void *succ_off, *succ_adj, *hlabel, *tparent, *tlabel;
  int32_t *dims, *origin, *ready;
  int capacity, sN1, sM, sN, sNT, epoch;
};
struct TcChunks { int n; int32_t end[32]; const int32_t* ready; int flags; };  // n = 0: everything is resident; flags: tuning aids (PT_TC_FLAGS)

__device__ __forceinline__ int ld_volatile(const int32_t* p) { return *(const volatile int32_t*)p; }
constexpr long long kWatchdogCycles = 8000000000ll;  // ~4 s at 1.9 GHz: far beyond any legitimate wait

// S = element type of the caller's arrays (int32_t, int16_t for index_bytes = 2, uint8_t for index_bytes = 1).  MAXT = launch
// bound = register budget: 768 threads -> 80 registers (the default: 24 warps per SM), 1024 -> 64.
// CORE = the variant that finishes a stuck instance by enumerating its small core (TcSolver::finish_small_core) instead of branching.
// It is a separate instantiation because its code is not free: 6 KB more SASS made EVERY instance 4.8 % slower (7.87 -> 8.25 ms per
// 10^5; 92 % of the instances never reach that code -- this kernel lives on its instruction cache), while what it buys is per LAUNCH
// (no chains of children in the tail: -0.08 ms) plus 1.4 % of work.  Measured break-even: ~28 000 instances per launch; the host picks
// the variant by the size of the batch it launches (one rank's shard of BASELINE config 2 at 8 GPUs: 12 500).
template <class I, class S, int MAXT, bool CORE>
__global__ void __launch_bounds__(MAXT, 1) tc_persistent_kernel(TcSource src, TcPool pool, TcChunks chunks, TcCtl* __restrict__ ctl, int capN, int capM, int capNT, int capL,
                                                                int warp_bytes, uint32_t* __restrict__ result_bits, uint8_t* __restrict__ status,
                                                                uint8_t* __restrict__ branched, unsigned long long* __restrict__ counters) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  WarpEx ex{lane, PT_PHASE_INIT};
  TcWork<I> w;
  w.carve(reinterpret_cast<char*>(smem) + (size_t)warp * warp_bytes, capN, capM, capNT, capL);
  const int B = (int)src.count;
  unsigned long long c_disp = 0, c_items = 0, c_br = 0, c_ch = 0, c_ret = 0, c_arc = 0, c_pass = 0, c_err = 0;
  bool batch_done = false;        // (lane 0) the ticket has run past B
  int avail = chunks.n ? 0 : B;   // (lane 0) instances [0, avail) are known to be resident
  int nready = 0;
  int mine = -1;                  // (lane 0) claimed pool index not consumed yet
  __shared__ int s_done;
  if (threadIdx.x == 0) s_done = 0;
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) ctl->t_first = global_ns();
  for (;;) {
    // ---- take a work item --------------------------------------------------------------------------------------------------
    // `mine` is a pool index this warp has claimed with ONE atomicAdd (no retry loop: a compare-and-swap on the head serialises
    // thousands of idle warps -- measured 3.4 us per item).  A claim is only made when the pool looks non-empty or when there is
    // nothing else to do; if the claimed slot has not been pushed yet the warp goes on with the batch and looks again later, and
    // at the end it waits on ITS OWN ready word (every idle warp polls a different address).  The end of the job is published
    // by the last finisher in ctl->done; warp 0 of a CTA relays it to the CTA through shared memory.
    int item = -1, from_pool = 0;
    if (lane == 0) {
      long long t0 = 0;
      unsigned spins = 0;
      for (;;) {
        if (mine < 0 && (batch_done || !(chunks.flags & 2))) {
          const int h = ld_volatile(&ctl->pool_head), t = ld_volatile(&ctl->pool_tail);
          if (h < t || batch_done) mine = atomicAdd(&ctl->pool_head, 1);
        }
        if (mine >= 0 && mine < pool.capacity && ld_volatile(&pool.ready[mine]) == pool.epoch) { __threadfence(); item = mine; from_pool = 1; mine = -1; break; }
        if (!batch_done) {
          const int i = atomicAdd(&ctl->ticket, 1);
          if (i < B) { item = i; break; }
          batch_done = true;
          atomicMax(&ctl->t_batch_done_inv, ~global_ns());
          continue;
        }
        // nothing to do: wait for the claimed slot or for the end of the job
        if (*(volatile int*)&s_done) { item = -2; break; }
        if (warp == 0 || (spins & 31) == 31) { if (ld_volatile(&ctl->done) || ld_volatile(&ctl->abort)) { *(volatile int*)&s_done = 1; item = -2; break; } }
        if (!t0) t0 = clock64(); else if ((spins & 255) == 255 && clock64() - t0 > kWatchdogCycles) { atomicExch(&ctl->abort, 1); item = -2; break; }
        ++spins;
        __nanosleep(spins < 64 ? 100 : 400);
      }
      if (item >= 0 && !from_pool && item >= avail) {    // the instance's chunk may still be on its way (HOST batch)
        long long t1 = 0;
        while (item >= avail) {
          if (nready < chunks.n && ld_volatile(&chunks.ready[nready])) { avail = chunks.end[nready]; ++nready; continue; }
          if (!t1) t1 = clock64(); else if (clock64() - t1 > kWatchdogCycles || nready >= chunks.n) { atomicExch(&ctl->abort, 1); item = -2; break; }
          __nanosleep(500);
        }
        __threadfence();
      }
    }
    item = __shfl_sync(0xffffffffu, item, 0);
    from_pool = __shfl_sync(0xffffffffu, from_pool, 0);
    if (item < 0) break;
    TcInst in;
    int origin;
    bool skip = false;
    if (from_pool) {
      origin = pool.origin[item];
      int done = 0;
      if (lane == 0) done = (int)((*(volatile uint32_t*)&result_bits[origin >> 5] >> (origin & 31)) & 1u);
      skip = __shfl_sync(0xffffffffu, done, 0) != 0;  // a sibling branch already displayed it
      in.n = pool.dims[4 * item]; in.m = pool.dims[4 * item + 1]; in.nt = pool.dims[4 * item + 2];
      in.succ_off = static_cast<const S*>(pool.succ_off) + (size_t)item * pool.sN1; in.succ_adj = static_cast<const S*>(pool.succ_adj) + (size_t)item * pool.sM;
      in.hlabel = static_cast<const S*>(pool.hlabel) + (size_t)item * pool.sN; in.tparent = static_cast<const S*>(pool.tparent) + (size_t)item * pool.sNT;
      in.tlabel = static_cast<const S*>(pool.tlabel) + (size_t)item * pool.sNT;
    } else {
      origin = item;
      batch_inst<S>(src, item, in);
    }
    if (!skip) {
      TcSolver<WarpEx, I> sv(ex, w);
      sv.allow_core_ = !(src.solver_flags & 1);
      int st = 0, x = -1;
      int code = sv.template begin<S>(in, from_pool != 0, &st);
      ++c_items;
      while (code == TC_CONTINUE) {
        code = sv.rule_loop(&st, &x);
        ex.phase = 13;
        if (code != TC_BRANCH) break;
#ifdef PT_AONLY
        code = TC_NOT; ++c_err; break;
#endif
        if (CORE) { const int core = sv.finish_small_core(); if (core >= 0) { code = core ? TC_DISPLAYED : TC_NOT; break; } }   // what is left is tiny: decide it here
        if (chunks.flags & 4) { code = TC_NOT; break; }   // measuring aid: no branching at all (wrong verdicts; shows the tail without chains)
        // ---- branch on x: children 1 .. d-1 go to the pool, child 0 continues in place -------------------------------------
        const int d = sv.indeg(x);
        int base = -1;
        if (lane == 0) {
          int old = ld_volatile(&ctl->pool_tail);
          for (;;) {  // reserve without ever counting a slot that will not be written
            if (old + (d - ((chunks.flags & 1) ? 0 : 1)) > pool.capacity) { base = -1; break; }
            const int seen = atomicCAS(&ctl->pool_tail, old, old + (d - ((chunks.flags & 1) ? 0 : 1)));
            if (seen == old) { base = old; break; }
            old = seen;
          }
          if (base < 0) { status[origin] = (uint8_t)TCS_POOL_OVERFLOW; atomicExch(&ctl->overflow, 1); }
          else if (!from_pool) branched[origin] = 1;
        }
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base < 0) { code = TC_ERROR; st = TCS_POOL_OVERFLOW; ++c_err; break; }
        c_br += (unsigned long long)d;
        int32_t* arcs = reinterpret_cast<int32_t*>(w.lset);  // the label sets are dead by now; d <= n <= capN * W words
        for (int k = lane; k < d; k += 32) arcs[k] = sv.in_arc(x, k);
        __syncwarp();
        const int k0 = (chunks.flags & 1) ? 0 : 1;  // tuning aid: bit 0 = no in-place continuation (the parent only emits)
        for (int k = k0; k < d; ++k) {
          const size_t slot = (size_t)(base + k - k0);
          TcEmitT<S> out{static_cast<S*>(pool.succ_off) + slot * pool.sN1, static_cast<S*>(pool.succ_adj) + slot * pool.sM, static_cast<S*>(pool.hlabel) + slot * pool.sN,
                         static_cast<S*>(pool.tparent) + slot * pool.sNT, static_cast<S*>(pool.tlabel) + slot * pool.sNT, pool.dims + slot * 4};
          sv.emit_child(x, arcs[k], out);
          __threadfence();
          __syncwarp();
          if (lane == 0) { pool.origin[slot] = origin; __threadfence(); atomicExch(&pool.ready[slot], pool.epoch); }
        }
        int done = 0;
        if (lane == 0) done = (int)((*(volatile uint32_t*)&result_bits[origin >> 5] >> (origin & 31)) & 1u);
        if (__shfl_sync(0xffffffffu, done, 0) || (chunks.flags & 1)) { code = TC_NOT; break; }  // a sibling branch displayed it meanwhile
        sv.keep_only(x, arcs[0]);
        ++c_items;
        code = TC_CONTINUE;
      }
      c_ch += sv.st.cherries; c_ret += sv.st.retcherries; c_arc += sv.st.arcs_deleted; c_pass += sv.st.passes;
      if (sv.core_switchings_) {   // decided by enumerating the switchings of the small core: not "by the rules alone"
        if (lane == 0 && !from_pool) branched[origin] = 1;
        c_br += (unsigned long long)sv.core_switchings_;
      }
      if (code == TC_DISPLAYED) {
        if (lane == 0) { atomicOr(&result_bits[origin >> 5], 1u << (origin & 31)); }
        ++c_disp;
      } else if (code == TC_ERROR && st != TCS_POOL_OVERFLOW) {
        if (lane == 0) status[origin] = (uint8_t)st;
        ++c_err;
      }
    }
    __syncwarp();
    if (lane == 0) {
      // the LAST finisher sees finished == B + pool_tail (finished is read by the add, the tail after it: both only grow and
      // every push precedes its parent's finish) and publishes the end of the job
      __threadfence();
      const int f = atomicAdd(&ctl->finished, 1) + 1;
      __threadfence();
      if (f == B + ld_volatile(&ctl->pool_tail)) atomicExch(&ctl->done, 1);
    }
  }
  if (lane == 0) {
    atomicMax(&ctl->t_last, global_ns());
    if (c_disp) atomicAdd(&counters[CNT_DISPLAYED], c_disp);
    if (c_items) atomicAdd(&counters[CNT_ITEMS], c_items);
    if (c_br) atomicAdd(&counters[CNT_BRANCH_ITEMS], c_br);
    if (c_ch) atomicAdd(&counters[CNT_CHERRIES], c_ch);
    if (c_ret) atomicAdd(&counters[CNT_RET], c_ret);
    if (c_arc) atomicAdd(&counters[CNT_ARCS], c_arc);
    if (c_pass) atomicAdd(&counters[CNT_PASSES], c_pass);
    if (c_err) atomicAdd(&counters[CNT_ERRORS], c_err);
  }
}

// reserve d consecutive pool slots; -1 if they do not fit.  The counter is only ever advanced over slots that WILL be written
// (an add-then-check would leave counted but unwritten slots behind for the next round to read)
__device__ __forceinline__ int pool_reserve(int32_t* count, int d, int capacity) {
  int old = *(volatile int32_t*)count;
  for (;;) {
    if (old + d > capacity) return -1;
    const int seen = atomicCAS(count, old, old + d);
    if (seen == old) return old;
    old = seen;
  }
}

// one CTA per work item, workspace in global memory
template <class S>
__global__ void __launch_bounds__(512) tc_reduce_cta_kernel(TcSource src, TcSink sink, int capN, int capM, int capNT, int capL, size_t ws_bytes,
                                                            unsigned char* __restrict__ ws, uint32_t* __restrict__ result_bits, uint8_t* __restrict__ status,
                                                            uint8_t* __restrict__ branched, int32_t* __restrict__ ticket, unsigned long long* __restrict__ counters) {
  using I = int32_t;
  __shared__ int s_red[40];
  CtaEx ex{s_red, 0, 0};
  TcWork<I> w;
  w.carve(reinterpret_cast<char*>(ws) + (size_t)blockIdx.x * ws_bytes, capN, capM, capNT, capL);
  const int64_t total = src.is_pool ? (int64_t)min(*src.count_ptr, src.capacity) : src.count;
  unsigned long long c_disp = 0, c_items = 0, c_br = 0, c_ch = 0, c_ret = 0, c_arc = 0, c_pass = 0, c_err = 0;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_red[36] = (int)src.first + atomicAdd(ticket, 1);
    __syncthreads();
    const int item = s_red[36];
    if (g_trace && threadIdx.x == 0) g_trace[blockIdx.x * 16 + 8] = item;
    if (item >= total) break;
    TcInst in;
    int origin;
    if (src.is_pool) {
      origin = src.origin[item];
      if (threadIdx.x == 0) s_red[37] = (int)((*(volatile uint32_t*)&result_bits[origin >> 5] >> (origin & 31)) & 1u);
      __syncthreads();
      if (s_red[37]) continue;
      in.n = src.dims[4 * item]; in.m = src.dims[4 * item + 1]; in.nt = src.dims[4 * item + 2];
      in.succ_off = static_cast<const int32_t*>(src.succ_off) + (size_t)item * src.sN1; in.succ_adj = static_cast<const int32_t*>(src.succ_adj) + (size_t)item * src.sM;
      in.hlabel = static_cast<const int32_t*>(src.hlabel) + (size_t)item * src.sN; in.tparent = static_cast<const int32_t*>(src.tparent) + (size_t)item * src.sNT;
      in.tlabel = static_cast<const int32_t*>(src.tlabel) + (size_t)item * src.sNT;
    } else {
      origin = item;
      batch_inst<S>(src, item, in);
    }
    TcSolver<CtaEx, I> sv(ex, w);
    sv.allow_core_ = !(src.solver_flags & 1);
    int st = 0, x = -1;
    const int code = sv.template solve<S>(in, src.is_pool != 0, &st, &x);
    ++c_items; c_ch += sv.st.cherries; c_ret += sv.st.retcherries; c_arc += sv.st.arcs_deleted; c_pass += sv.st.passes;
    if (sv.core_switchings_) {   // decided by enumerating the switchings of the small core: not "by the rules alone"
      if (threadIdx.x == 0) branched[origin] = 1;
      c_br += (unsigned long long)sv.core_switchings_;
    }
    if (code == TC_DISPLAYED) {
      if (threadIdx.x == 0) atomicOr(&result_bits[origin >> 5], 1u << (origin & 31));
      ++c_disp;
    } else if (code == TC_ERROR) {
      if (threadIdx.x == 0) status[origin] = (uint8_t)st;
      ++c_err;
    } else if (code == TC_BRANCH) {
      const int d = sv.indeg(x);
      if (threadIdx.x == 0) s_red[38] = pool_reserve(sink.count, d, sink.capacity);
      __syncthreads();
      const int base = s_red[38];
      if (base < 0) {
        if (threadIdx.x == 0) status[origin] = (uint8_t)TCS_POOL_OVERFLOW;
      } else {
        if (threadIdx.x == 0) branched[origin] = 1;
        c_br += (unsigned long long)d;
        int32_t* arcs = reinterpret_cast<int32_t*>(w.lset);  // signatures are dead by now; d <= n <= capN * W words
        for (int k = threadIdx.x; k < d; k += blockDim.x) arcs[k] = sv.in_arc(x, k);
        __syncthreads();
        for (int k = 0; k < d; ++k) {
          const int keep = arcs[k];
          const size_t slot = (size_t)(base + k);
          TcEmit out{sink.succ_off + slot * sink.sN1, sink.succ_adj + slot * sink.sM, sink.hlabel + slot * sink.sN,
                     sink.tparent + slot * sink.sNT, sink.tlabel + slot * sink.sNT, sink.dims + slot * 4};
          sv.emit_child(x, keep, out);
          if (threadIdx.x == 0) sink.origin[slot] = origin;
        }
      }
    }
  }
  if (threadIdx.x == 0) {
    if (c_disp) atomicAdd(&counters[CNT_DISPLAYED], c_disp);
    if (c_items) atomicAdd(&counters[CNT_ITEMS], c_items);
    if (c_br) atomicAdd(&counters[CNT_BRANCH_ITEMS], c_br);
    if (c_ch) atomicAdd(&counters[CNT_CHERRIES], c_ch);
    if (c_ret) atomicAdd(&counters[CNT_RET], c_ret);
    if (c_arc) atomicAdd(&counters[CNT_ARCS], c_arc);
    if (c_pass) atomicAdd(&counters[CNT_PASSES], c_pass);
    if (c_err) atomicAdd(&counters[CNT_ERRORS], c_err);
  }
}

// one thread-block cluster per work item (launched with cudaLaunchAttributeClusterDimension), workspace in global memory:
// 256 bytes of executor scalars, then the TcWork arrays with the spare arc arrays
template <class S>
__global__ void __launch_bounds__(PT_CLUSTER_THREADS, 1) tc_reduce_cluster_kernel(TcSource src, TcSink sink, int capN, int capM, int capNT, int capL, size_t ws_bytes,
                                                                unsigned char* __restrict__ ws, uint32_t* __restrict__ result_bits, uint8_t* __restrict__ status,
                                                                uint8_t* __restrict__ branched, int32_t* __restrict__ ticket, unsigned long long* __restrict__ counters) {
  using I = int32_t;
  __shared__ int s_red[40];
  cooperative_groups::cluster_group cl = cooperative_groups::this_cluster();
  const unsigned csize = cl.num_blocks(), rank = cl.block_rank(), cluster_id = blockIdx.x / csize;
  char* const base_ws = reinterpret_cast<char*>(ws) + (size_t)cluster_id * ws_bytes;
  ClusterEx ex{s_red, reinterpret_cast<int32_t*>(base_ws), 0, rank, csize};
  TcWork<I> w;
  w.carve(base_ws + 256, capN, capM, capNT, capL, true);
  const int64_t total = src.is_pool ? (int64_t)min(*src.count_ptr, src.capacity) : src.count;
  const bool leader = ex.tid() == 0;
  unsigned long long c_disp = 0, c_items = 0, c_br = 0, c_ch = 0, c_ret = 0, c_arc = 0, c_pass = 0, c_err = 0;
  for (;;) {
    ex.sync();
    if (leader) ex.gsc[2] = (int)src.first + atomicAdd(ticket, 1);
    const int item = ex.uniform(&ex.gsc[2]);
    if (item >= total) break;
    TcInst in;
    int origin;
    if (src.is_pool) {
      origin = src.origin[item];
      ex.sync();
      if (leader) ex.gsc[3] = (int)((*(volatile uint32_t*)&result_bits[origin >> 5] >> (origin & 31)) & 1u);
      if (ex.uniform(&ex.gsc[3])) continue;
      in.n = src.dims[4 * item]; in.m = src.dims[4 * item + 1]; in.nt = src.dims[4 * item + 2];
      in.succ_off = static_cast<const int32_t*>(src.succ_off) + (size_t)item * src.sN1; in.succ_adj = static_cast<const int32_t*>(src.succ_adj) + (size_t)item * src.sM;
      in.hlabel = static_cast<const int32_t*>(src.hlabel) + (size_t)item * src.sN; in.tparent = static_cast<const int32_t*>(src.tparent) + (size_t)item * src.sNT;
      in.tlabel = static_cast<const int32_t*>(src.tlabel) + (size_t)item * src.sNT;
    } else {
      origin = item;
      batch_inst<S>(src, item, in);
    }
    TcSolver<ClusterEx, I> sv(ex, w);
    sv.allow_core_ = !(src.solver_flags & 1);
    int st = 0, x = -1;
    const int code = sv.template solve<S>(in, src.is_pool != 0, &st, &x);
    ++c_items; c_ch += sv.st.cherries; c_ret += sv.st.retcherries; c_arc += sv.st.arcs_deleted; c_pass += sv.st.passes;
    if (sv.core_switchings_) {   // decided by enumerating the switchings of the small core: not "by the rules alone"
      if (leader) branched[origin] = 1;
      c_br += (unsigned long long)sv.core_switchings_;
    }
    if (code == TC_DISPLAYED) {
      if (leader) atomicOr(&result_bits[origin >> 5], 1u << (origin & 31));
      ++c_disp;
    } else if (code == TC_ERROR) {
      if (leader) status[origin] = (uint8_t)st;
      ++c_err;
    } else if (code == TC_BRANCH) {
      const int d = sv.indeg(x);
      ex.sync();
      if (leader) ex.gsc[4] = pool_reserve(sink.count, d, sink.capacity);
      const int base = ex.uniform(&ex.gsc[4]);
      if (base < 0) {
        if (leader) status[origin] = (uint8_t)TCS_POOL_OVERFLOW;
      } else {
        if (leader) branched[origin] = 1;
        c_br += (unsigned long long)d;
        int32_t* arcs = reinterpret_cast<int32_t*>(w.lset);  // signatures are dead by now; d <= n <= capN * W words
        for (int k = ex.tid(); k < d; k += ex.nth()) arcs[k] = sv.in_arc(x, k);
        ex.sync();
        for (int k = 0; k < d; ++k) {
          const int keep = *(volatile int32_t*)&arcs[k];
          const size_t slot = (size_t)(base + k);
          TcEmit out{sink.succ_off + slot * sink.sN1, sink.succ_adj + slot * sink.sM, sink.hlabel + slot * sink.sN,
                     sink.tparent + slot * sink.sNT, sink.tlabel + slot * sink.sNT, sink.dims + slot * 4};
          sv.emit_child(x, keep, out);
          if (leader) sink.origin[slot] = origin;
        }
      }
    }
  }
  if (leader) {
    if (c_disp) atomicAdd(&counters[CNT_DISPLAYED], c_disp);
    if (c_items) atomicAdd(&counters[CNT_ITEMS], c_items);
    if (c_br) atomicAdd(&counters[CNT_BRANCH_ITEMS], c_br);
    if (c_ch) atomicAdd(&counters[CNT_CHERRIES], c_ch);
    if (c_ret) atomicAdd(&counters[CNT_RET], c_ret);
    if (c_arc) atomicAdd(&counters[CNT_ARCS], c_arc);
    if (c_pass) atomicAdd(&counters[CNT_PASSES], c_pass);
    if (c_err) atomicAdd(&counters[CNT_ERRORS], c_err);
  }
}

// pt_options.max_rounds cut the branching short: the instances that still have children pending and no "displayed" child yet
// get a status instead of a silent "not displayed"
__global__ void tc_round_limit_kernel(const int32_t* __restrict__ origin, int items, const uint32_t* __restrict__ bits, uint8_t* __restrict__ status) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < items; i += gridDim.x * blockDim.x) {
    const int o = origin[i];
    if (!((bits[o >> 5] >> (o & 31)) & 1u)) status[o] = (uint8_t)TCS_ROUND_LIMIT;
  }
}

// per-batch maxima of n, m, nt and label id (for DEVICE input without hints)
__global__ void tc_maxima_kernel(int64_t B, const int64_t* hnode_off, const int64_t* harc_off, const int64_t* gnode_off,
                                 const void* hlabel, const void* tlabel, int idx_bytes, int32_t* out4) {
  auto lab = [idx_bytes](const void* a, int64_t i) -> int {
    if (idx_bytes == 1) { const int x = static_cast<const uint8_t*>(a)[i]; return x == 255 ? -1 : x; }
    return idx_bytes == 2 ? (int)static_cast<const short*>(a)[i] : static_cast<const int*>(a)[i];
  };
  int mn = 0, mm = 0, mt = 0, ml = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
    mn = max(mn, (int)(hnode_off[i + 1] - hnode_off[i])); mm = max(mm, (int)(harc_off[i + 1] - harc_off[i])); mt = max(mt, (int)(gnode_off[i + 1] - gnode_off[i]));
  }
  const int64_t NH = hnode_off[B], NG = gnode_off[B];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < NH; i += (int64_t)gridDim.x * blockDim.x) ml = max(ml, lab(hlabel, i) + 1);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < NG; i += (int64_t)gridDim.x * blockDim.x) ml = max(ml, lab(tlabel, i) + 1);
  for (int d = 16; d; d >>= 1) {
    mn = max(mn, __shfl_xor_sync(0xffffffffu, mn, d)); mm = max(mm, __shfl_xor_sync(0xffffffffu, mm, d));
    mt = max(mt, __shfl_xor_sync(0xffffffffu, mt, d)); ml = max(ml, __shfl_xor_sync(0xffffffffu, ml, d));
  }
  if ((threadIdx.x & 31) == 0) { atomicMax(&out4[0], mn); atomicMax(&out4[1], mm); atomicMax(&out4[2], mt); atomicMax(&out4[3], ml); }
}

}  // namespace ptgpu

using namespace ptgpu;

#ifdef PT_PROFILE
extern "C" int pt_debug_profile(unsigned long long* out128) {
  unsigned long long z[128] = {};
  if (cudaMemcpyFromSymbol(out128, g_prof, sizeof(z)) != cudaSuccess) return -1;
  return cudaMemcpyToSymbol(g_prof, z, sizeof(z)) == cudaSuccess ? 0 : -1;
}
#endif

static constexpr int kDefaultMaxWarps = 24;  // measured on config 2: 16 warps 8.5 ms, 20: 8.0, 24: 7.6, 28: 8.2 (instruction fetch, not latency, is the limit)

struct pt_batch {
  bool outputs_prezeroed = false;   // set by pt_batch_contain_sharded for the next call
  // pt_batch_free waits for THIS handle's work only (an event recorded behind the last thing a call enqueued), not for everything the
  // caller has queued on the stream since: a caller that keeps the next batch's solve queued behind this one must not be stalled by the free
  cudaEvent_t last_work = nullptr;
  bool work_recorded = false;
  int64_t B = 0;
  bool owns_input = false;
  int64_t *d_hnode_off = nullptr, *d_harc_off = nullptr, *d_gnode_off = nullptr;
  void *d_succ_off = nullptr, *d_succ_adj = nullptr, *d_hlabel = nullptr, *d_tparent = nullptr, *d_tlabel = nullptr;
  int idx_bytes = 4;  // element size of the five arrays above: 4, 2 (int16_t) or 1 (uint8_t, out-degrees instead of offsets)
  int32_t maxN = 0, maxM = 0, maxNT = 0, maxL = 0;
  // results + scratch
  uint32_t* d_bits = nullptr; uint8_t* d_status = nullptr; uint8_t* d_branched = nullptr;
  size_t scratch_bytes = 0;
  int32_t* d_ticket = nullptr;  // 256 bytes: the warp path's TcCtl; the global-memory executors: [0] ticket, [1] pool count A, [2] pool count B, [4..] chunk tickets
  unsigned long long* d_counters = nullptr;
  // pools of the global-memory executors (ping-pong, int32) and of the persistent warp kernel (one, element type of the batch)
  int pool_cap = 0;
  int32_t* d_pool[2] = {nullptr, nullptr};
  unsigned char* d_ppool = nullptr; size_t ppool_bytes = 0; int ppool_cap = 0; int32_t* d_pready = nullptr; int pool_epoch = 0;
  unsigned char* d_ws = nullptr; size_t ws_cap = 0;  // global-memory workspaces of the CTA executor
  // HOST input is uploaded in chunks of instances on a private copy stream.  The persistent warp kernel polls d_chunk_ready[c]
  // (written by the copy stream after chunk c); the global-memory executors launch one kernel per chunk behind chunk_ev[c].
  static constexpr int kMaxChunks = 32;
  cudaStream_t copy_stream = nullptr;
  int nchunks = 1; int64_t chunk_end[kMaxChunks] = {}; cudaEvent_t chunk_ev[kMaxChunks] = {};
  int32_t* d_chunk_ready = nullptr; cudaEvent_t flags_zeroed = nullptr, upload_done = nullptr;
  int32_t* h_ctl = nullptr;  // pinned copy of the TcCtl of the last warp-path launch (watchdog / overflow), read at the next call
  bool ctl_pending = false;
  int64_t last_launches = 0;
  int64_t h2d_bytes = 0;
  static constexpr int kMaxTimedRounds = 16;
  cudaEvent_t ev[2 * kMaxTimedRounds] = {};
  // the persistent launches of the last kRing calls keep their own event pairs, so that a caller who queues many steps without waiting
  // in between can read every step's kernel time afterwards (pt_batch_launch_ms)
  static constexpr int kRing = 64;
  cudaEvent_t ring[2 * kRing] = {};
  int64_t launch_seq = 0;
  int timed_rounds = 0;
  cudaStream_t used_stream = nullptr; bool used_stream_valid = false;
};

// per-handle host resources that are expensive to create (cudaHostAlloc, cudaStreamCreate: tens of microseconds each, on the
// end-to-end path once per batch): kept in process-wide free lists
namespace {
std::mutex g_res_mu;
std::vector<int32_t*> g_free_ctl;                     // pinned 64-byte blocks
std::vector<std::pair<int, cudaStream_t>> g_free_streams;  // (device, non-blocking stream)
}
static int32_t* take_pinned_ctl() {
  {
    std::lock_guard<std::mutex> lk(g_res_mu);
    if (!g_free_ctl.empty()) { int32_t* p = g_free_ctl.back(); g_free_ctl.pop_back(); return p; }
  }
  int32_t* p = nullptr;
  if (cudaHostAlloc((void**)&p, 64, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
static void give_pinned_ctl(int32_t* p) { if (p) { std::lock_guard<std::mutex> lk(g_res_mu); g_free_ctl.push_back(p); } }
static cudaStream_t take_copy_stream() {
  int dev = 0;
  cudaGetDevice(&dev);
  {
    std::lock_guard<std::mutex> lk(g_res_mu);
    for (size_t i = 0; i < g_free_streams.size(); ++i)
      if (g_free_streams[i].first == dev) { cudaStream_t s = g_free_streams[i].second; g_free_streams.erase(g_free_streams.begin() + (long)i); return s; }
  }
  cudaStream_t s = nullptr;
  if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return s;
}
static void give_copy_stream(cudaStream_t s) {
  if (!s) return;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(g_res_mu);
  g_free_streams.emplace_back(dev, s);
}

static void batch_release(pt_batch* b) {
  if (!b) return;
  // cached blocks may be handed to another handle at once: everything this handle has enqueued must be over
  if (b->work_recorded && b->last_work) cudaEventSynchronize(b->last_work);
  else if (b->used_stream_valid) cudaStreamSynchronize(b->used_stream);
  if (b->copy_stream) { cudaStreamSynchronize(b->copy_stream); give_copy_stream(b->copy_stream); }
  if (b->owns_input) {
    dev_free(b->d_hnode_off); dev_free(b->d_harc_off); dev_free(b->d_gnode_off);
    dev_free(b->d_succ_off); dev_free(b->d_succ_adj); dev_free(b->d_hlabel); dev_free(b->d_tparent); dev_free(b->d_tlabel);
  }
  dev_free(b->d_bits); dev_free(b->d_status); dev_free(b->d_ticket);   // (counters and branched flags live inside d_ticket's block)
  dev_free(b->d_pool[0]); dev_free(b->d_pool[1]); dev_free(b->d_ws); dev_free(b->d_ppool); dev_free(b->d_pready); dev_free(b->d_chunk_ready);
  for (cudaEvent_t e : b->ev) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : b->ring) if (e) cudaEventDestroy(e);
  if (b->last_work) cudaEventDestroy(b->last_work);
  for (cudaEvent_t e : b->chunk_ev) if (e) cudaEventDestroy(e);
  if (b->flags_zeroed) cudaEventDestroy(b->flags_zeroed);
  if (b->upload_done) cudaEventDestroy(b->upload_done);
  give_pinned_ctl(b->h_ctl);
  delete b;
}

// the word "1" in pinned host memory: what the copy stream writes into chunk_ready[c] (a 4-byte copy-engine transfer that is
// ordered behind the chunk's own copies; a memset could be a kernel and would not find room next to the persistent kernel)
static const int32_t* pinned_one() {
  static int32_t* one = nullptr;
  static std::once_flag once;
  std::call_once(once, []() { if (cudaHostAlloc((void**)&one, 64, cudaHostAllocPortable) == cudaSuccess) *one = 1; else { cudaGetLastError(); one = nullptr; } });
  return one;
}

template <class T>
static int upload(T** dst, const T* src, int64_t count, cudaStream_t s, int64_t* bytes) {
  const size_t nb = (size_t)std::max<int64_t>(count, 1) * sizeof(T);
  if (int rc = dev_alloc((void**)dst, nb)) return rc;
  if (count > 0) { PT_CUDA(cudaMemcpyAsync(*dst, src, (size_t)count * sizeof(T), cudaMemcpyHostToDevice, s)); *bytes += (int64_t)count * (int64_t)sizeof(T); }
  return PT_OK;
}

// marks the end of what a call has enqueued for this handle on its stream (see pt_batch::last_work)
static void batch_mark_work(pt_batch* b, cudaStream_t s) {
  b->work_recorded = false;
  if (!b->last_work && cudaEventCreateWithFlags(&b->last_work, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); b->last_work = nullptr; return; }
  if (cudaEventRecord(b->last_work, s) == cudaSuccess) b->work_recorded = true; else cudaGetLastError();
}

// common scratch of a handle (result words, status, flags, counters)
static int batch_scratch(pt_batch* b) {
  const int64_t B = b->B;
  int rc = PT_OK;
  // control block, counters and the "branched" flags share one allocation: pt_batch_contain clears them with one memset
  b->scratch_bytes = 256 + 128 + (((size_t)B + 1 + 15) & ~(size_t)15);
  if ((rc = dev_alloc((void**)&b->d_bits, ((size_t)((B + 31) / 32) + 1) * 4)) || (rc = dev_alloc((void**)&b->d_status, (size_t)B + 1)) ||
      (rc = dev_alloc((void**)&b->d_ticket, b->scratch_bytes)))
    return rc;
  b->d_counters = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(b->d_ticket) + 256);
  b->d_branched = reinterpret_cast<uint8_t*>(b->d_ticket) + 256 + 128;
  static_assert(CNT_N * sizeof(unsigned long long) <= 128, "counter block");
  if (!(b->h_ctl = take_pinned_ctl())) return set_error(PT_ERR_NOMEM, "cudaHostAlloc failed");
  memset(b->h_ctl, 0, 64);
  return PT_OK;
}

// the TcCtl of the previous warp-path launch on this handle: a fired watchdog is an error of THAT call, reported here
static int batch_check_ctl(pt_batch* b) {
  if (!b->ctl_pending) return PT_OK;
  b->ctl_pending = false;
  if (b->work_recorded && b->last_work) PT_CUDA(cudaEventSynchronize(b->last_work));
  else if (b->used_stream_valid) PT_CUDA(cudaStreamSynchronize(b->used_stream));
  const TcCtl* c = reinterpret_cast<const TcCtl*>(b->h_ctl);
  if (c->abort) return set_error(PT_ERR_CUDA, "pt_batch_contain: the device watchdog fired (a wait for an upload chunk or a pool slot did not end); results are incomplete");
  return PT_OK;
}

// launch of the persistent warp kernel for element type S
// csrc/split_kernels.cuh (compiled into lca_kernels.o): the bridge-splitting path for instances beyond shared memory
extern "C" int ptgpu_split_contain(int64_t B, const int64_t* d_hnode_off, const int64_t* d_harc_off, const int64_t* d_gnode_off, const void* d_succ_off, const void* d_succ_adj,
                                   const void* d_hlabel, const void* d_tparent, const void* d_tlabel, int idx_bytes, uint32_t* d_bits, uint8_t* d_status, uint64_t* blobs_solved,
                                   int* applied, void* stream);

constexpr int64_t kCoreMaxBatch = 28000;
template <class S>
static cudaError_t launch_persistent(int wpc, int grid, size_t smem, cudaStream_t s, const TcSource& src, const TcPool& pool, const TcChunks& ch, TcCtl* ctl, const pt_batch* b,
                                     int warp_bytes, uint32_t* bits, uint8_t* stat) {
  using I = int16_t;
  // the small-core variant for launches below the measured break-even (see the kernel), PT_TC_CORE=0 / 1 forces one (A/B aid)
  static const int env_core = [] { const char* e = getenv("PT_TC_CORE"); return e ? atoi(e) : -1; }();
  const bool core = !(src.solver_flags & 1) && (env_core >= 0 ? env_core != 0 : src.count <= kCoreMaxBatch);
  auto kern = wpc > 24 ? (core ? tc_persistent_kernel<I, S, 1024, true> : tc_persistent_kernel<I, S, 1024, false>)
                       : (core ? tc_persistent_kernel<I, S, 768, true> : tc_persistent_kernel<I, S, 768, false>);
  // the attribute is per function and device: set it again only when it has to grow (one process may drive several devices)
  static std::mutex mu; static size_t have[4][64] = {};
  int dev = 0; cudaGetDevice(&dev);
  cudaError_t e = cudaSuccess;
  { std::lock_guard<std::mutex> g(mu);
    size_t& h = have[(wpc > 24 ? 1 : 0) + (core ? 2 : 0)][dev & 63];
    if (h < smem) { e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e == cudaSuccess) h = smem; } }
  if (e != cudaSuccess) return e;
  kern<<<grid, wpc * 32, smem, s>>>(src, pool, ch, ctl, (int)b->maxN, (int)b->maxM, (int)b->maxNT, (int)b->maxL, warp_bytes, bits, stat, b->d_branched, b->d_counters);
  return cudaGetLastError();
}

extern "C" {

// debug only (not part of include/pt_gpu.h): returns a host pointer to 16 ints per block (1024 blocks) that the CTA
// executor updates while running
int* pt_debug_trace(void) {
  static int* host = nullptr;
  if (!host) {
    if (cudaHostAlloc((void**)&host, 1024 * 64, cudaHostAllocMapped) != cudaSuccess) return nullptr;
    memset(host, 0, 1024 * 64);
    int* dev = nullptr;
    if (cudaHostGetDevicePointer((void**)&dev, host, 0) != cudaSuccess) return nullptr;
    volatile int* d = dev;
    cudaMemcpyToSymbol(ptgpu::g_trace, &d, sizeof(d));
  }
  return host;
}

int pt_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
int pt_set_device(int device) {
  if (int rc = require_device()) return rc;
  PT_CUDA(cudaSetDevice(device));
  return PT_OK;
}

int pt_get_limits(pt_limits* out) {
  if (!out) return set_error(PT_ERR_INVALID, "pt_get_limits: null argument");
  // limits of the shared-memory (warp) executor: ids in int16, whole instance in the 227 KB of one CTA's shared memory.
  // Larger instances run on the global-memory (CTA) executor with int32 ids, bounded by device memory only.
  out->max_host_nodes = 30000; out->max_host_arcs = 30000; out->max_guest_nodes = 30000; out->max_labels = 30000;
  return PT_OK;
}

int pt_batch_create(pt_batch** out, const pt_batch_desc* d, void* stream) {
  if (!out || !d) return set_error(PT_ERR_INVALID, "pt_batch_create: null argument");
  *out = nullptr;
  if (d->num_instances < 0 || d->num_instances > (int64_t)INT32_MAX - 64) return set_error(PT_ERR_INVALID, "pt_batch_create: bad num_instances");
  if (!d->host_node_off || !d->host_arc_off || !d->guest_node_off || !d->host_succ_off || !d->host_label || !d->guest_parent || !d->guest_label)
    return set_error(PT_ERR_INVALID, "pt_batch_create: a required array is null");
  // host_succ_adj may only be null when the batch has no arc at all (std::vector::data() of an empty vector)
  if (!d->host_succ_adj && d->num_instances > 0 && (d->mem != PT_MEM_HOST || d->host_arc_off[d->num_instances] > 0))
    return set_error(PT_ERR_INVALID, "pt_batch_create: host_succ_adj is null");
  if (d->index_bytes != 0 && d->index_bytes != 1 && d->index_bytes != 2 && d->index_bytes != 4) return set_error(PT_ERR_INVALID, "pt_batch_create: index_bytes must be 0, 1, 2 or 4");
  if (int rc = require_device()) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  static const bool trace = getenv("PT_CREATE_TRACE") != nullptr;   // measuring aid: host time of every phase of this call
  auto t_prev = std::chrono::steady_clock::now();
  auto mark = [&](const char* what) {
    if (!trace) return;
    const auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "[create] %-28s %8.1f us\n", what, std::chrono::duration<double, std::micro>(t - t_prev).count());
    t_prev = t;
  };
  pt_batch* b = new (std::nothrow) pt_batch();
  if (!b) return set_error(PT_ERR_NOMEM, "out of host memory");
  const int64_t B = b->B = d->num_instances;
  b->idx_bytes = d->index_bytes ? d->index_bytes : 4;
  const size_t esz = (size_t)b->idx_bytes;
  const bool deg = b->idx_bytes == 1;  // succ_off holds n out-degrees per instance instead of n + 1 offsets
  auto elem = [&](const int32_t* base, int64_t i) {
    return b->idx_bytes == 1 ? (reinterpret_cast<const uint8_t*>(base)[i] == 255 ? -1 : (int32_t) reinterpret_cast<const uint8_t*>(base)[i])
                             : (b->idx_bytes == 2 ? (int32_t) reinterpret_cast<const int16_t*>(base)[i] : base[i]);
  };
  b->used_stream = s; b->used_stream_valid = true;
  int rc = PT_OK;
  auto fail = [&](int code) { batch_release(b); return code; };
  b->maxN = d->max_host_nodes; b->maxM = d->max_host_arcs; b->maxNT = d->max_guest_nodes; b->maxL = d->max_labels;
  const bool need_max = !(b->maxN > 0 && b->maxNT > 0 && b->maxL > 0) && B > 0;
  if (d->mem == PT_MEM_HOST) {
    // structural sanity of the offset tables before trusting them for the copy sizes
    // only the entries that size the copies are looked at here (first, last and the chunk boundaries below); every instance's own range is
    // checked against the array lengths by the kernel that loads it (batch_inst -> PT_INST_BAD_GRAPH)
    if (d->host_node_off[0] != 0 || d->host_arc_off[0] != 0 || d->guest_node_off[0] != 0) return fail(set_error(PT_ERR_INVALID, "pt_batch_create: offset tables must start at 0"));
    const int64_t NH = d->host_node_off[B], MH = d->host_arc_off[B], NG = d->guest_node_off[B];
    if (NH < 0 || MH < 0 || NG < 0) return fail(set_error(PT_ERR_INVALID, "pt_batch_create: offset tables must be non-decreasing"));
    if (need_max) {
      for (int64_t i = 0; i < B; ++i) {
        b->maxN = std::max<int32_t>(b->maxN, (int32_t)std::min<int64_t>(d->host_node_off[i + 1] - d->host_node_off[i], INT32_MAX));
        b->maxM = std::max<int32_t>(b->maxM, (int32_t)std::min<int64_t>(d->host_arc_off[i + 1] - d->host_arc_off[i], INT32_MAX));
        b->maxNT = std::max<int32_t>(b->maxNT, (int32_t)std::min<int64_t>(d->guest_node_off[i + 1] - d->guest_node_off[i], INT32_MAX));
      }
      int32_t ml = 0;
      for (int64_t i = 0; i < NH; ++i) ml = std::max(ml, elem(d->host_label, i) + 1);
      for (int64_t i = 0; i < NG; ++i) ml = std::max(ml, elem(d->guest_label, i) + 1);
      b->maxL = std::max(b->maxL, ml);
    }
    b->owns_input = true;
    mark("checks");
    if ((rc = dev_alloc(&b->d_succ_off, (size_t)std::max<int64_t>(NH + (deg ? 0 : B), 1) * esz)) || (rc = dev_alloc(&b->d_succ_adj, (size_t)std::max<int64_t>(MH, 1) * esz)) ||
        (rc = dev_alloc(&b->d_hlabel, (size_t)std::max<int64_t>(NH, 1) * esz)) || (rc = dev_alloc(&b->d_tparent, (size_t)std::max<int64_t>(NG, 1) * esz)) ||
        (rc = dev_alloc(&b->d_tlabel, (size_t)std::max<int64_t>(NG, 1) * esz)))
      return fail(rc);
    // Chunks that double in size (1/32, 1/32, 1/16, ... 1/2 of the batch): the persistent kernel starts on a small first chunk
    // and the copy engine stays ahead of the solver from then on (the upload is 3-4 times faster than the solve), with few
    // API calls: their host time comes BEFORE the kernel launch and delays it.
    // measured end to end at 12 500 instances (one rank's shard of config 2 at N = 8; profiles/chunks_sweep.sh): 1 chunk 1.49 ms, 2: 1.36, 3: 1.31, 4: 1.32, 6: 1.37
    b->nchunks = B >= 65536 ? 6 : (B >= 4096 ? 3 : 1);
    static const int env_chunks = [] { const char* e = getenv("PT_TC_CHUNKS"); return e ? std::max(1, std::min((int)pt_batch::kMaxChunks, atoi(e))) : 0; }();  // tuning aid
    if (env_chunks) b->nchunks = env_chunks;
    const int32_t* one = b->nchunks > 1 ? pinned_one() : nullptr;
    if (b->nchunks > 1 && (!one || !(b->copy_stream = take_copy_stream()))) { b->nchunks = 1; b->copy_stream = nullptr; }
    cudaStream_t cs = b->nchunks > 1 ? b->copy_stream : s;
    // which executor will run?  (the same test as pt_batch_contain; only the global-memory executors need one event per chunk)
    const bool warp_path = b->maxN > 0 && b->maxN <= 30000 && b->maxM <= 30000 && b->maxNT <= 30000 && b->maxL <= 30000 &&
                           TcWork<int16_t>::bytes(std::max(b->maxN, 1), std::max(b->maxM > 0 ? b->maxM : 2 * b->maxN, 1), std::max(b->maxNT, 1), std::max(b->maxL, 1)) <= (size_t)kMaxSmemPerBlock;
    if (b->nchunks > 1) {
      // the flags are zeroed on the copy stream before anything else happens there; `s` (and any other stream a later
      // pt_batch_contain uses) waits for that, so no kernel can see a stale "ready" left in the cached block
      if ((rc = dev_alloc((void**)&b->d_chunk_ready, pt_batch::kMaxChunks * 4))) return fail(rc);
      cudaError_t e = cudaMemsetAsync(b->d_chunk_ready, 0, pt_batch::kMaxChunks * 4, cs);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b->flags_zeroed, cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventRecord(b->flags_zeroed, cs);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(s, b->flags_zeroed, 0);
      if (e != cudaSuccess) return fail(set_error(PT_ERR_CUDA, "pt_batch_create: %s", cudaGetErrorString(e)));
    }
    mark("allocations + flags");
    // offset tables first (small), then the five data arrays chunk by chunk, all on the copy stream when there is one: the kernel looks at
    // nothing before it has seen chunk_ready[0], which is written behind them, and the caller's stream carries no copy of ours between
    // the previous batch's kernel and this one's
    if ((rc = upload(&b->d_hnode_off, d->host_node_off, B + 1, cs, &b->h2d_bytes)) || (rc = upload(&b->d_harc_off, d->host_arc_off, B + 1, cs, &b->h2d_bytes)) ||
        (rc = upload(&b->d_gnode_off, d->guest_node_off, B + 1, cs, &b->h2d_bytes)))
      return fail(rc);
    mark("offset tables");
    auto chunk_bound = [&](int k) -> int64_t {
      if (k <= 0) return 0;
      if (k >= b->nchunks) return B;
      return (B >> (b->nchunks - k)) / 32 * 32;   // B/32, B/16, ..., B/2
    };
    for (int c = 0; c < b->nchunks; ++c) {
      const int64_t i0 = chunk_bound(c), i1 = chunk_bound(c + 1);
      b->chunk_end[c] = i1;
      const int64_t h0 = d->host_node_off[i0], h1 = d->host_node_off[i1], a0 = d->host_arc_off[i0], a1 = d->host_arc_off[i1], g0 = d->guest_node_off[i0], g1 = d->guest_node_off[i1];
      if (h0 < 0 || h1 < h0 || h1 > NH || a0 < 0 || a1 < a0 || a1 > MH || g0 < 0 || g1 < g0 || g1 > NG)
        return fail(set_error(PT_ERR_INVALID, "pt_batch_create: offset tables must be non-decreasing (around instance %lld)", (long long)i0));
      auto cp = [&](void* dst, const void* src, int64_t from, int64_t to) -> cudaError_t {
        if (to <= from) return cudaSuccess;
        b->h2d_bytes += (to - from) * (int64_t)esz;
        return cudaMemcpyAsync(static_cast<char*>(dst) + from * esz, static_cast<const char*>(src) + from * esz, (size_t)(to - from) * esz, cudaMemcpyHostToDevice, cs);
      };
      cudaError_t e = cp(b->d_succ_off, d->host_succ_off, h0 + (deg ? 0 : i0), h1 + (deg ? 0 : i1));
      if (e == cudaSuccess) e = cp(b->d_succ_adj, d->host_succ_adj, a0, a1);
      if (e == cudaSuccess) e = cp(b->d_hlabel, d->host_label, h0, h1);
      if (e == cudaSuccess) e = cp(b->d_tparent, d->guest_parent, g0, g1);
      if (e == cudaSuccess) e = cp(b->d_tlabel, d->guest_label, g0, g1);
      if (e == cudaSuccess && b->nchunks > 1) {
        e = cudaMemcpyAsync(b->d_chunk_ready + c, one, 4, cudaMemcpyHostToDevice, cs);
        if (e == cudaSuccess && !warp_path) { e = cudaEventCreateWithFlags(&b->chunk_ev[c], cudaEventDisableTiming); if (e == cudaSuccess) e = cudaEventRecord(b->chunk_ev[c], cs); }
      }
      if (e != cudaSuccess) return fail(set_error(PT_ERR_CUDA, "pt_batch_create: upload failed: %s", cudaGetErrorString(e)));
    }
    mark("chunk copies");
    if (b->nchunks > 1) {  // for an executor that was not expected here (pt_options.path): everything has landed
      cudaError_t e = cudaEventCreateWithFlags(&b->upload_done, cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventRecord(b->upload_done, cs);
      if (e != cudaSuccess) return fail(set_error(PT_ERR_CUDA, "pt_batch_create: %s", cudaGetErrorString(e)));
    }
  } else {
    b->d_hnode_off = (int64_t*)d->host_node_off; b->d_harc_off = (int64_t*)d->host_arc_off; b->d_gnode_off = (int64_t*)d->guest_node_off;
    b->d_succ_off = (void*)d->host_succ_off; b->d_succ_adj = (void*)d->host_succ_adj; b->d_hlabel = (void*)d->host_label;
    b->d_tparent = (void*)d->guest_parent; b->d_tlabel = (void*)d->guest_label;
    if (need_max) {
      int32_t* d4 = nullptr; int32_t h4[4] = {0, 0, 0, 0};
      if (int rc = dev_alloc((void**)&d4, 16)) return fail(rc);
      cudaMemsetAsync(d4, 0, 16, s);
      tc_maxima_kernel<<<296, 256, 0, s>>>(B, b->d_hnode_off, b->d_harc_off, b->d_gnode_off, b->d_hlabel, b->d_tlabel, b->idx_bytes, d4);
      cudaError_t e = cudaMemcpyAsync(h4, d4, 16, cudaMemcpyDeviceToHost, s);
      if (e == cudaSuccess) e = cudaStreamSynchronize(s); else cudaStreamSynchronize(s);
      dev_free(d4);
      if (e != cudaSuccess) return fail(set_error(PT_ERR_CUDA, "maxima pass failed: %s", cudaGetErrorString(e)));
      b->maxN = std::max(b->maxN, h4[0]); b->maxM = std::max(b->maxM, h4[1]); b->maxNT = std::max(b->maxNT, h4[2]); b->maxL = std::max(b->maxL, h4[3]);
    }
  }
  if (b->maxM <= 0) b->maxM = std::max(1, 2 * b->maxN);
  b->maxN = std::max(b->maxN, 1); b->maxNT = std::max(b->maxNT, 1); b->maxL = std::max(b->maxL, 1);
  if (b->idx_bytes == 1 && (b->maxN > 255 || b->maxNT > 255 || b->maxL > 255))
    return fail(set_error(PT_ERR_INVALID, "pt_batch_create: index_bytes = 1 needs at most 255 host nodes, guest nodes and labels per instance"));
  if ((rc = batch_scratch(b))) return fail(rc);
  batch_mark_work(b, s);
  mark("scratch");
  *out = b;
  return PT_OK;
}

// array `which` of the batch as it lives on the device (0 host_node_off, 1 host_arc_off, 2 guest_node_off: int64; 3 succ_off,
// 4 succ_adj, 5 host_label, 6 guest_parent, 7 guest_label: int32, int16 or uint8 per *index_bytes) copied to `dst`
int pt_batch_download(pt_batch* b, int which, void* dst, int64_t capacity_bytes, int64_t* bytes, int32_t* index_bytes, int32_t maxima[4]) {
  if (!b || which < 0 || which > 7 || !bytes) return set_error(PT_ERR_INVALID, "pt_batch_download: bad argument");
  const int64_t B = b->B;
  int64_t tail[3] = {0, 0, 0};
  if (B > 0) {
    if (b->used_stream_valid) PT_CUDA(cudaStreamSynchronize(b->used_stream));
    if (b->copy_stream) PT_CUDA(cudaStreamSynchronize(b->copy_stream));
    PT_CUDA(cudaMemcpy(&tail[0], b->d_hnode_off + B, 8, cudaMemcpyDeviceToHost)); PT_CUDA(cudaMemcpy(&tail[1], b->d_harc_off + B, 8, cudaMemcpyDeviceToHost));
    PT_CUDA(cudaMemcpy(&tail[2], b->d_gnode_off + B, 8, cudaMemcpyDeviceToHost));
  }
  const int64_t esz = b->idx_bytes;
  const void* src[8] = {b->d_hnode_off, b->d_harc_off, b->d_gnode_off, b->d_succ_off, b->d_succ_adj, b->d_hlabel, b->d_tparent, b->d_tlabel};
  const int64_t n[8] = {(B + 1) * 8, (B + 1) * 8, (B + 1) * 8, (tail[0] + (esz == 1 ? 0 : B)) * esz, tail[1] * esz, tail[0] * esz, tail[2] * esz, tail[2] * esz};
  *bytes = n[which];
  if (index_bytes) *index_bytes = (int32_t)esz;
  if (maxima) { maxima[0] = b->maxN; maxima[1] = b->maxM; maxima[2] = b->maxNT; maxima[3] = b->maxL; }
  if (dst) {
    if (capacity_bytes < n[which]) return set_error(PT_ERR_INVALID, "pt_batch_download: buffer too small");
    if (n[which] > 0) PT_CUDA(cudaMemcpy(dst, src[which], (size_t)n[which], cudaMemcpyDeviceToHost));
  }
  return PT_OK;
}

int64_t pt_batch_last_launches(const pt_batch* b) { return b ? b->last_launches : 0; }

// A batch whose device arrays were produced on the device (csrc/newick_kernels.cu): the handle takes ownership of the eight
// dev_alloc'ed arrays (order: host_node_off, host_arc_off, guest_node_off, succ_off, succ_adj, host_label, guest_parent, guest_label).
int ptgpu_batch_adopt(pt_batch** out, int64_t B, void* const arrays[8], int index_bytes, const int32_t maxima[4], void* stream) {
  pt_batch* b = new (std::nothrow) pt_batch();
  if (!b) return set_error(PT_ERR_NOMEM, "out of host memory");
  b->B = B; b->owns_input = true; b->idx_bytes = index_bytes;
  b->used_stream = (cudaStream_t)stream; b->used_stream_valid = true;
  b->d_hnode_off = (int64_t*)arrays[0]; b->d_harc_off = (int64_t*)arrays[1]; b->d_gnode_off = (int64_t*)arrays[2];
  b->d_succ_off = arrays[3]; b->d_succ_adj = arrays[4]; b->d_hlabel = arrays[5]; b->d_tparent = arrays[6]; b->d_tlabel = arrays[7];
  b->maxN = std::max(maxima[0], 1); b->maxM = std::max(maxima[1], 1); b->maxNT = std::max(maxima[2], 1); b->maxL = std::max(maxima[3], 1);
  if (int rc = batch_scratch(b)) { batch_release(b); return rc; }
  *out = b;
  return PT_OK;
}
int64_t pt_batch_h2d_bytes(const pt_batch* b) { return b ? b->h2d_bytes : 0; }
// internal (comm.cu): the caller's DEVICE result words / status bytes of the next pt_batch_contain are already zero
void ptgpu_batch_outputs_prezeroed(pt_batch* b) { if (b) b->outputs_prezeroed = true; }
// debug / profiling aid (not part of include/pt_gpu.h): phases of the last persistent launch in microseconds --
// out[0] = first warp in -> first warp that found the batch ticket exhausted, out[1] = from there to the last warp out (the tail),
// out[2] = pool slots used, out[3] = items finished
int pt_debug_last_phases(pt_batch* b, double out[4]) {
  if (!b || !b->h_ctl) return PT_ERR_INVALID;
  if (b->used_stream_valid) cudaStreamSynchronize(b->used_stream);
  const TcCtl* c = reinterpret_cast<const TcCtl*>(b->h_ctl);
  const unsigned long long t_batch_done = ~c->t_batch_done_inv;
  out[0] = (double)(t_batch_done - c->t_first) / 1e3; out[1] = (double)(c->t_last - t_batch_done) / 1e3; out[2] = c->pool_tail; out[3] = c->finished;
  return PT_OK;
}
int pt_batch_round_ms(const pt_batch* bc, int round, float* ms) {
  pt_batch* b = const_cast<pt_batch*>(bc);
  if (b && ms && round == 0 && b->timed_rounds == -1) return pt_batch_launch_ms(b, b->launch_seq - 1, ms);
  if (!b || !ms || round < 0 || round >= b->timed_rounds) return set_error(PT_ERR_INVALID, "pt_batch_round_ms: round %d did not run (or was not timed)", round);
  PT_CUDA(cudaEventSynchronize(b->ev[2 * round + 1]));  // pt_batch_contain with DEVICE outputs does not wait for its kernel
  PT_CUDA(cudaEventElapsedTime(ms, b->ev[2 * round], b->ev[2 * round + 1]));
  return batch_check_ctl(b);
}
int64_t pt_batch_launch_count(const pt_batch* b) { return b ? b->launch_seq : 0; }
int pt_batch_launch_ms(const pt_batch* bc, int64_t launch, float* ms) {
  pt_batch* b = const_cast<pt_batch*>(bc);
  if (!b || !ms || launch < 0 || launch >= b->launch_seq || launch < b->launch_seq - pt_batch::kRing)
    return set_error(PT_ERR_INVALID, "pt_batch_launch_ms: launch %lld is not among the last %d persistent launches of this handle", (long long)launch, pt_batch::kRing);
  const int slot = (int)(launch % pt_batch::kRing);
  PT_CUDA(cudaEventSynchronize(b->ring[2 * slot + 1]));
  PT_CUDA(cudaEventElapsedTime(ms, b->ring[2 * slot], b->ring[2 * slot + 1]));
  return launch == b->launch_seq - 1 ? batch_check_ctl(b) : PT_OK;
}

int pt_batch_contain(pt_batch* b, const pt_options* opt, uint32_t* result_bits, uint8_t* status, pt_mem out_mem, pt_counters* counters, void* stream) {
  if (!b || !result_bits) return set_error(PT_ERR_INVALID, "pt_batch_contain: null argument");
  if (int rc = require_device()) return rc;
  if (int rc = batch_check_ctl(b)) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  DeviceProps dp;
  if (int rc = device_props(&dp)) return rc;
  using I = int16_t;
  const int64_t B = b->B;
  const size_t words = (size_t)((B + 31) / 32);
  b->last_launches = 0; b->timed_rounds = 0;
  b->work_recorded = false;   // until the end of this call is marked, pt_batch_free falls back to a stream synchronisation
  if (b->flags_zeroed && (!b->used_stream_valid || b->used_stream != s)) PT_CUDA(cudaStreamWaitEvent(s, b->flags_zeroed, 0));
  b->used_stream = s; b->used_stream_valid = true;
  uint32_t* bits = out_mem == PT_MEM_DEVICE ? result_bits : b->d_bits;
  uint8_t* stat = (out_mem == PT_MEM_DEVICE && status) ? status : b->d_status;
  const bool prezeroed = b->outputs_prezeroed && out_mem == PT_MEM_DEVICE;   // pt_batch_contain_sharded has just cleared the whole vectors
  b->outputs_prezeroed = false;
  if (!prezeroed) PT_CUDA(cudaMemsetAsync(bits, 0, words * 4, s));
  if (!prezeroed || stat == b->d_status) PT_CUDA(cudaMemsetAsync(stat, 0, (size_t)B, s));
  PT_CUDA(cudaMemsetAsync(b->d_ticket, 0, b->scratch_bytes, s));   // control block + counters + branched flags
  uint64_t rounds = 0;
  // executors over the same solver: warps with the instance in shared memory (int16 ids, ONE persistent launch), or one CTA /
  // one cluster per instance with the workspace in global memory (int32 ids, one launch per branching depth) when an instance
  // does not fit one CTA's shared memory
  // pt_options.path: 0 auto, 1 / 2 / 3 force an executor, 4 = automatic choice but never split (what the split path uses for its
  // blobs), 5 = split along the bridges whatever the size (falls back to the automatic choice when the batch is not clean)
  const bool want_split = opt && opt->path == 5, no_split = opt && opt->path == 4;
  const int force = (opt && opt->path >= 1 && opt->path <= 3) ? opt->path : 0;
  const bool fits16 = b->maxN <= 30000 && b->maxM <= 30000 && b->maxNT <= 30000 && b->maxL <= 30000;
  const size_t warp_bytes = TcWork<I>::bytes(b->maxN, b->maxM, b->maxNT, b->maxL);
  const bool use_cta = force == 2 || force == 3 || !fits16 || warp_bytes > (size_t)dp.smem_optin;
  const int sN1 = b->maxN + 1, sM = std::max(1, (int)b->maxM), sN = b->maxN, sNT = b->maxNT;
  const size_t slot_elems = (size_t)sN1 + sM + sN + 2 * (size_t)sNT;
  int want = opt && opt->pool_slots > 0 ? opt->pool_slots : (int)std::min<int64_t>(std::max<int64_t>(8192, B), 1 << 22);
  bool split_done = false;
  if (B > 0 && force == 0 && !no_split && want_split) {
    // ---- split along the bridges first (north_star subsystem 1), on request (pt_options.path = 5).  The split path decides CLEAN
    // batches completely (tree parts by cluster comparisons, bridge-free blobs as small instances); anything else goes to the
    // executors below.  It is NOT what the automatic choice takes for instances beyond shared memory: measured on BASELINE config 4
    // (8 x n = 10^6, r = 10^3, generator networks) the reticulations of an instance fall into ONE bridge-free component of ~16 500
    // nodes (+ 15 000 pendants), so the split costs two LCA builds over the whole forest (7.4 ms of sweeps) and still leaves a blob
    // for the global-memory executor (6.8 ms): 37-43 ms against 24.5 ms unsplit (21.4 against 22.9 ms for a single instance).
    // Networks with small, local blobs are where it pays (profiles/README.md).
    if (b->copy_stream) PT_CUDA(cudaStreamSynchronize(b->copy_stream));   // a HOST batch: its chunks must have landed
    int applied = 0; uint64_t blobs = 0;
    if (int rc = ptgpu_split_contain(B, b->d_hnode_off, b->d_harc_off, b->d_gnode_off, b->d_succ_off, b->d_succ_adj, b->d_hlabel, b->d_tparent, b->d_tlabel, b->idx_bytes,
                                     bits, stat, &blobs, &applied, stream)) return rc;
    if (applied) {
      split_done = true; rounds = 1; b->last_launches = 40;
      const unsigned long long items = blobs;
      PT_CUDA(cudaMemcpyAsync(b->d_counters + CNT_ITEMS, &items, 8, cudaMemcpyHostToDevice, s));
      PT_CUDA(cudaStreamSynchronize(s));
    } else {
      PT_CUDA(cudaMemsetAsync(bits, 0, words * 4, s));
      PT_CUDA(cudaMemsetAsync(stat, 0, (size_t)B, s));
    }
  }
  if (split_done) {
  } else if (force == 1 && use_cta && B > 0) {
    PT_CUDA(cudaMemsetAsync(stat, PT_INST_TOO_LARGE, (size_t)B, s));  // caller insisted on the shared-memory path
  } else if (B > 0 && !use_cta) {
    // ---- the warp executor: one persistent launch --------------------------------------------------------------------------
    // occupancy: as many warps per SM as shared memory allows (warp_bytes each), in one CTA per SM; the kernel variant
    // (register budget) follows from the warp count
    int max_warps = kDefaultMaxWarps;
    static const int env_warps = [] { const char* e = getenv("PT_TC_MAX_WARPS"); return e ? std::max(1, std::min(32, atoi(e))) : 0; }();  // tuning aid
    if (env_warps) max_warps = env_warps;
    const int wpc = std::max(1, (int)std::min<size_t>((size_t)max_warps, ((size_t)dp.smem_optin) / warp_bytes));
    const size_t smem = (size_t)wpc * warp_bytes;
    const size_t esz = (size_t)b->idx_bytes;
    if (!(opt && opt->pool_slots > 0)) want = (int)std::max<size_t>(16, std::min<size_t>((size_t)want, ((size_t)4 << 30) / (slot_elems * esz + 32)));  // <= 4 GB
    const size_t stride_bytes = (slot_elems * esz + 15) & ~(size_t)15;
    if (b->ppool_cap < want) {
      dev_free(b->d_ppool); dev_free(b->d_pready); b->d_ppool = nullptr; b->d_pready = nullptr; b->ppool_cap = 0;
      const size_t arrays = ((size_t)(sN1 + sM + sN + 2 * (size_t)sNT) * esz * (size_t)want + 63) & ~(size_t)63;
      if (int rc2 = dev_alloc((void**)&b->d_ppool, arrays + (size_t)want * 5 * 4 + 64)) return rc2;
      if (int rc2 = dev_alloc((void**)&b->d_pready, (size_t)want * 4)) return rc2;
      PT_CUDA(cudaMemsetAsync(b->d_pready, 0, (size_t)want * 4, s));
      b->ppool_cap = want; b->pool_epoch = 0; b->ppool_bytes = arrays;
    }
    (void)stride_bytes;
    if (++b->pool_epoch == 0x7fffffff) { PT_CUDA(cudaMemsetAsync(b->d_pready, 0, (size_t)b->ppool_cap * 4, s)); b->pool_epoch = 1; }
    const int cap = std::min(b->ppool_cap, want);  // an explicit pt_options.pool_slots also caps an already larger pool
    TcPool pool{};
    {
      // arrays are laid out for the handle's full capacity so that the int32 tail stays aligned whatever `cap` is
      unsigned char* p = b->d_ppool; const size_t C = (size_t)b->ppool_cap;
      pool.succ_off = p; p += (size_t)sN1 * esz * C; pool.succ_adj = p; p += (size_t)sM * esz * C; pool.hlabel = p; p += (size_t)sN * esz * C;
      pool.tparent = p; p += (size_t)sNT * esz * C; pool.tlabel = p;
      int32_t* q = reinterpret_cast<int32_t*>(b->d_ppool + b->ppool_bytes);
      pool.dims = q; q += 4 * C; pool.origin = q;
      pool.ready = b->d_pready; pool.capacity = cap; pool.sN1 = sN1; pool.sM = sM; pool.sN = sN; pool.sNT = sNT; pool.epoch = b->pool_epoch;
    }
    TcSource src{};
    src.solver_flags = (opt && opt->no_small_core) ? 1 : 0;
    src.first = 0; src.count = B; src.is_pool = 0; src.hnode_off = b->d_hnode_off; src.harc_off = b->d_harc_off; src.gnode_off = b->d_gnode_off;
    src.succ_off = b->d_succ_off; src.succ_adj = b->d_succ_adj; src.hlabel = b->d_hlabel; src.tparent = b->d_tparent; src.tlabel = b->d_tlabel;
    TcChunks ch{};
    static const int env_flags = [] { const char* e = getenv("PT_TC_FLAGS"); return e ? atoi(e) : 0; }();   // tuning / measuring aid
    ch.flags = env_flags;
    if (env_flags & 8) src.solver_flags |= 1;   // (bit 3: never finish by the small core)
    if (b->nchunks > 1 && b->d_chunk_ready) { ch.n = b->nchunks; for (int c = 0; c < b->nchunks; ++c) ch.end[c] = (int32_t)b->chunk_end[c]; ch.ready = b->d_chunk_ready; }
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)dp.sms, (B + wpc - 1) / wpc));
    const int slot = (int)(b->launch_seq % pt_batch::kRing);
    for (int j = 0; j < 2; ++j) if (!b->ring[2 * slot + j]) PT_CUDA(cudaEventCreate(&b->ring[2 * slot + j]));
    PT_CUDA(cudaEventRecord(b->ring[2 * slot], s));
    TcCtl* ctl = reinterpret_cast<TcCtl*>(b->d_ticket);
    cudaError_t e = b->idx_bytes == 1 ? launch_persistent<uint8_t>(wpc, grid, smem, s, src, pool, ch, ctl, b, (int)warp_bytes, bits, stat)
                  : b->idx_bytes == 2 ? launch_persistent<int16_t>(wpc, grid, smem, s, src, pool, ch, ctl, b, (int)warp_bytes, bits, stat)
                                      : launch_persistent<int32_t>(wpc, grid, smem, s, src, pool, ch, ctl, b, (int)warp_bytes, bits, stat);
    if (e != cudaSuccess) return set_error(PT_ERR_CUDA, "pt_batch_contain: launch failed: %s", cudaGetErrorString(e));
    PT_CUDA(cudaEventRecord(b->ring[2 * slot + 1], s));
    ++b->launch_seq;
    b->timed_rounds = -1; b->last_launches = 1; rounds = 1;   // (-1: round 0 is the newest ring slot)
    PT_CUDA(cudaMemcpyAsync(b->h_ctl, b->d_ticket, sizeof(TcCtl), cudaMemcpyDeviceToHost, s));
    b->ctl_pending = true;
  } else if (B > 0) {
    // ---- the global-memory executors: one launch per branching depth over ping-pong pools ----------------------------------
    int ctas_per_sm = 1;
    size_t ws_bytes = 0;
    int cta_grid_cap = 0, max_cluster = 8;
    {
      // one stride for both global-memory executors: 256 bytes of cluster scalars + the arrays with the spare arc arrays
      ws_bytes = (256 + TcWork<int32_t>::bytes(b->maxN, b->maxM, b->maxNT, b->maxL, true) + 255) & ~(size_t)255;
      size_t free_b = 0, total_b = 0;
      PT_CUDA(cudaMemGetInfo(&free_b, &total_b));
      const size_t budget = free_b / 3;
      cta_grid_cap = (int)std::min<size_t>((size_t)dp.sms * 2, std::max<size_t>(1, budget / ws_bytes));
      cta_grid_cap = (int)std::min<int64_t>(cta_grid_cap, B);
      // measured on config 4 (8 instances of n = 10^6): clusters of 4: 43.6 ms, 8: 29.8 ms, 16 (non-portable size): 35.0 ms
      if (const char* e = getenv("PT_TC_MAX_CLUSTER")) {  // tuning aid; 16 needs the non-portable opt-in
        max_cluster = std::max(1, std::min(16, atoi(e)));
        if (max_cluster > 8 && (cudaFuncSetAttribute(tc_reduce_cluster_kernel<int16_t>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
                                cudaFuncSetAttribute(tc_reduce_cluster_kernel<int32_t>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
                                cudaFuncSetAttribute(tc_reduce_cluster_kernel<uint8_t>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)) {
          cudaGetLastError(); max_cluster = 8;
        }
      }
      if (b->ws_cap < (size_t)cta_grid_cap * ws_bytes) {
        dev_free(b->d_ws); b->d_ws = nullptr; b->ws_cap = 0;
        if (int rc2 = dev_alloc((void**)&b->d_ws, (size_t)cta_grid_cap * ws_bytes)) return rc2;
        b->ws_cap = (size_t)cta_grid_cap * ws_bytes;
      }
    }
    // pool
    const size_t slot_ints = slot_elems + 4 + 1;
    if (!(opt && opt->pool_slots > 0)) want = (int)std::max<size_t>(16, std::min<size_t>((size_t)want, ((size_t)4 << 30) / (slot_ints * 4)));  // <= 4 GB per pool
    if (b->pool_cap < want) {
      dev_free(b->d_pool[0]); dev_free(b->d_pool[1]); b->d_pool[0] = b->d_pool[1] = nullptr; b->pool_cap = 0;
      for (int k = 0; k < 2; ++k) if (int rc2 = dev_alloc((void**)&b->d_pool[k], slot_ints * (size_t)want * 4)) return rc2;
      b->pool_cap = want;
    }
    const int cap = std::min(b->pool_cap, want);  // an explicit pt_options.pool_slots also caps an already larger pool
    auto sink_of = [&](int k) {
      TcSink t; int32_t* p = b->d_pool[k];
      t.succ_off = p; p += (size_t)sN1 * cap; t.succ_adj = p; p += (size_t)sM * cap; t.hlabel = p; p += (size_t)sN * cap;
      t.tparent = p; p += (size_t)sNT * cap; t.tlabel = p; p += (size_t)sNT * cap; t.dims = p; p += (size_t)4 * cap; t.origin = p;
      t.count = b->d_ticket + 1 + k; t.capacity = cap; t.sN1 = sN1; t.sM = sM; t.sN = sN; t.sNT = sNT;
      return t;
    };
    // every round fixes one more multi-parent node, so the depth is bounded by the number of nodes; pt_options.max_rounds can
    // only lower that bound (whatever is still pending then is REPORTED, see tc_round_limit_kernel)
    const int64_t max_rounds = opt && opt->max_rounds > 0 ? opt->max_rounds : (int64_t)b->maxN + 2;
    TcSource src{};
    const int solver_flags = (opt && opt->no_small_core) ? 1 : 0;
    src.solver_flags = solver_flags;
    src.count = B; src.is_pool = 0; src.hnode_off = b->d_hnode_off; src.harc_off = b->d_harc_off; src.gnode_off = b->d_gnode_off;
    src.succ_off = b->d_succ_off; src.succ_adj = b->d_succ_adj; src.hlabel = b->d_hlabel; src.tparent = b->d_tparent; src.tlabel = b->d_tlabel;
    int64_t items = B;
    TcSink pending{};
    for (int64_t round = 0; items > 0; ++round) {
      if (round >= max_rounds) {  // not a verdict: whoever is still in the pool is told so
        tc_round_limit_kernel<<<(unsigned)std::min<int64_t>((items + 255) / 256, 1024), 256, 0, s>>>(pending.origin, (int)items, bits, stat);
        PT_CUDA(cudaGetLastError());
        break;
      }
      const int k = (int)(round & 1);
      TcSink sink = sink_of(k);
      int grid = (int)std::max<int64_t>(1, std::min<int64_t>(cta_grid_cap, items));
      const bool timed = round < pt_batch::kMaxTimedRounds;
      if (timed) {
        for (int j = 0; j < 2; ++j) if (!b->ev[2 * round + j]) PT_CUDA(cudaEventCreate(&b->ev[2 * round + j]));
        PT_CUDA(cudaEventRecord(b->ev[2 * round], s));
      }
      // round 0 of a HOST batch: one launch per upload chunk, each waiting only for its own bytes (if the handle was created
      // expecting the warp executor there are no per-chunk events: wait for the whole upload, one launch)
      const bool per_chunk = b->nchunks > 1 && b->chunk_ev[0] != nullptr;
      if (round == 0 && b->nchunks > 1 && !per_chunk && b->upload_done) PT_CUDA(cudaStreamWaitEvent(s, b->upload_done, 0));
      const int launches = (round == 0 && per_chunk) ? b->nchunks : 1;
      for (int c = 0; c < launches; ++c) {
        int32_t* ticket = b->d_ticket;
        if (round == 0) {
          src.first = c == 0 ? 0 : b->chunk_end[c - 1];
          src.count = launches > 1 ? b->chunk_end[c] : B;
          ticket = b->d_ticket + 4 + c;
          if (b->chunk_ev[c]) PT_CUDA(cudaStreamWaitEvent(s, b->chunk_ev[c], 0));
          const int64_t cnt = src.count - src.first;
          if (cnt <= 0) continue;
          grid = (int)std::max<int64_t>(1, std::min<int64_t>(cta_grid_cap, cnt));
        }
        const int in_bytes = src.is_pool ? 4 : b->idx_bytes;  // element type of what this launch reads
        // few huge instances: a cluster of CTAs per instance, as large as keeps every instance of this launch on its own SMs
        auto cluster_kernel = in_bytes == 1 ? tc_reduce_cluster_kernel<uint8_t> : (in_bytes == 2 ? tc_reduce_cluster_kernel<int16_t> : tc_reduce_cluster_kernel<int32_t>);
        auto cta_kernel = in_bytes == 1 ? tc_reduce_cta_kernel<uint8_t> : (in_bytes == 2 ? tc_reduce_cta_kernel<int16_t> : tc_reduce_cta_kernel<int32_t>);
        int csize = 1;
        if (force != 2) { while (csize < max_cluster && (int64_t)grid * csize * 2 <= dp.sms) csize *= 2; }
        if (force == 3 && csize == 1) csize = 2;
        if (csize > 1) {
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3((unsigned)(grid * csize)); cfg.blockDim = dim3(PT_CLUSTER_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = s;
          cudaLaunchAttribute attr[1];
          attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = (unsigned)csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
          cfg.attrs = attr; cfg.numAttrs = 1;
          PT_CUDA(cudaLaunchKernelEx(&cfg, cluster_kernel, src, sink, (int)b->maxN, (int)b->maxM,
                                     (int)b->maxNT, (int)b->maxL, ws_bytes, b->d_ws, bits, stat, b->d_branched, ticket, b->d_counters));
        } else
          cta_kernel<<<grid * ctas_per_sm, 512, 0, s>>>(src, sink, b->maxN, b->maxM, b->maxNT, b->maxL, ws_bytes, b->d_ws, bits, stat, b->d_branched, ticket, b->d_counters);
        ++b->last_launches;
      }
      PT_CUDA(cudaGetLastError());
      if (timed) { PT_CUDA(cudaEventRecord(b->ev[2 * round + 1], s)); b->timed_rounds = (int)round + 1; }
      ++rounds;
      int32_t h[4];
      PT_CUDA(cudaMemcpyAsync(h, b->d_ticket, 16, cudaMemcpyDeviceToHost, s));
      PT_CUDA(cudaStreamSynchronize(s));
      items = std::min<int64_t>(h[1 + k], cap);
      if (items <= 0) break;
      // next round reads the pool just written; reset the ticket and the other pool's counter
      pending = sink;
      src = TcSource{};
      src.solver_flags = solver_flags;
      src.is_pool = 1; src.count_ptr = pending.count; src.capacity = cap; src.succ_off = pending.succ_off; src.succ_adj = pending.succ_adj; src.hlabel = pending.hlabel;
      src.tparent = pending.tparent; src.tlabel = pending.tlabel; src.dims = pending.dims; src.origin = pending.origin; src.sN1 = sN1; src.sM = sM; src.sN = sN; src.sNT = sNT;
      PT_CUDA(cudaMemsetAsync(b->d_ticket, 0, 4, s));
      PT_CUDA(cudaMemsetAsync(b->d_ticket + 1 + (k ^ 1), 0, 4, s));
    }
  }
  if (out_mem == PT_MEM_HOST) {
    PT_CUDA(cudaMemcpyAsync(result_bits, bits, words * 4, cudaMemcpyDeviceToHost, s));
    if (status) PT_CUDA(cudaMemcpyAsync(status, stat, (size_t)B, cudaMemcpyDeviceToHost, s));
  }
  if (counters) {
    unsigned long long h[CNT_N];
    std::vector<uint8_t> hb((size_t)B + 1), hs((size_t)B + 1);
    std::vector<uint32_t> hw(words + 1);
    PT_CUDA(cudaMemcpyAsync(hw.data(), bits, words * 4, cudaMemcpyDeviceToHost, s));
    PT_CUDA(cudaMemcpyAsync(h, b->d_counters, sizeof(h), cudaMemcpyDeviceToHost, s));
    PT_CUDA(cudaMemcpyAsync(hb.data(), b->d_branched, (size_t)B, cudaMemcpyDeviceToHost, s));
    PT_CUDA(cudaMemcpyAsync(hs.data(), stat, (size_t)B, cudaMemcpyDeviceToHost, s));
    PT_CUDA(cudaStreamSynchronize(s));
    memset(counters, 0, sizeof(*counters));
    uint64_t nbr = 0, nerr = 0;
    uint64_t ndisp = 0;  // per ORIGINAL instance (several branch children of one instance may each report "displayed")
    for (int64_t i = 0; i < B; ++i) { nbr += hb[(size_t)i] != 0; nerr += hs[(size_t)i] != 0; ndisp += (hw[(size_t)(i >> 5)] >> (i & 31)) & 1u; }
    counters->instances = (uint64_t)B; counters->displayed = ndisp; counters->errors = nerr;
    counters->not_displayed = (uint64_t)B - ndisp - nerr; counters->branched = nbr; counters->resolved_by_rules = (uint64_t)B - nbr - nerr;
    counters->work_items = h[CNT_ITEMS]; counters->rounds = rounds; counters->cherry_reductions = h[CNT_CHERRIES];
    counters->reticulated_cherries = h[CNT_RET]; counters->arcs_deleted = h[CNT_ARCS]; counters->rule_passes = h[CNT_PASSES];
  } else if (out_mem == PT_MEM_HOST) {
    PT_CUDA(cudaStreamSynchronize(s));
  }
  batch_mark_work(b, s);
  if (out_mem == PT_MEM_HOST || counters) return batch_check_ctl(b);  // the stream has been synchronised: report a fired watchdog now
  return PT_OK;
}

int pt_batch_free(pt_batch* b) { batch_release(b); return PT_OK; }

// One call, all GPUs of the node, no torch: contiguous instance ranges (boundaries multiples of 32, so every device owns whole
// result words) on one host thread per device; no collective is needed -- the caller's buffers are filled shard by shard.
int pt_batch_contain_multi(const pt_batch_desc* d, int num_devices, const pt_options* opt, uint32_t* result_bits, uint8_t* status, pt_counters* counters) {
  if (!d || !result_bits) return set_error(PT_ERR_INVALID, "pt_batch_contain_multi: null argument");
  if (d->mem != PT_MEM_HOST) return set_error(PT_ERR_INVALID, "pt_batch_contain_multi: HOST arrays only (a DEVICE batch lives on one GPU: use pt_batch_create there)");
  if (int rc = require_device()) return rc;
  int have = 0;
  PT_CUDA(cudaGetDeviceCount(&have));
  const int G = num_devices <= 0 ? have : std::min(num_devices, 64);  // more shards than devices: they share the devices round-robin
  const int64_t B = d->num_instances;
  if (B < 0 || !d->host_node_off || !d->host_arc_off || !d->guest_node_off) return set_error(PT_ERR_INVALID, "pt_batch_contain_multi: bad batch");
  for (int64_t i = 0; i < B; ++i)
    if (d->host_node_off[i + 1] < d->host_node_off[i] || d->host_arc_off[i + 1] < d->host_arc_off[i] || d->guest_node_off[i + 1] < d->guest_node_off[i])
      return set_error(PT_ERR_INVALID, "pt_batch_contain_multi: offset tables must be non-decreasing (instance %lld)", (long long)i);
  int prev_dev = 0;
  cudaGetDevice(&prev_dev);
  const size_t esz = d->index_bytes == 2 ? 2 : (d->index_bytes == 1 ? 1 : 4);
  const bool deg = d->index_bytes == 1;  // succ_off holds n out-degrees per instance, not n + 1 offsets
  std::vector<int> rcs((size_t)G, PT_OK);
  std::vector<std::string> msgs((size_t)G);
  std::vector<pt_counters> cnt((size_t)G);
  std::vector<std::thread> workers;
  for (int g = 0; g < G; ++g) {
    const int64_t i0 = B * g / G / 32 * 32, i1 = (g + 1 == G) ? B : B * (g + 1) / G / 32 * 32;
    workers.emplace_back([&, g, i0, i1]() {
      memset(&cnt[(size_t)g], 0, sizeof(pt_counters));
      const int64_t nb = i1 - i0;
      if (nb <= 0) return;
      auto fail = [&](int rc) { rcs[(size_t)g] = rc; msgs[(size_t)g] = pt_last_error(); };
      if (cudaSetDevice(g % have) != cudaSuccess) { cudaGetLastError(); rcs[(size_t)g] = PT_ERR_CUDA; msgs[(size_t)g] = "cudaSetDevice failed"; return; }
      // the shard as its own batch: offset tables rebased to 0, data pointers moved to the shard's first element
      std::vector<int64_t> hn((size_t)nb + 1), ha((size_t)nb + 1), gn((size_t)nb + 1);
      for (int64_t k = 0; k <= nb; ++k) { hn[(size_t)k] = d->host_node_off[i0 + k] - d->host_node_off[i0]; ha[(size_t)k] = d->host_arc_off[i0 + k] - d->host_arc_off[i0]; gn[(size_t)k] = d->guest_node_off[i0 + k] - d->guest_node_off[i0]; }
      auto at = [&](const int32_t* base, int64_t elems) { return base ? reinterpret_cast<const int32_t*>(reinterpret_cast<const char*>(base) + (size_t)elems * esz) : nullptr; };
      pt_batch_desc sd = *d;
      sd.num_instances = nb; sd.host_node_off = hn.data(); sd.host_arc_off = ha.data(); sd.guest_node_off = gn.data();
      sd.host_succ_off = at(d->host_succ_off, d->host_node_off[i0] + (deg ? 0 : i0)); sd.host_succ_adj = at(d->host_succ_adj, d->host_arc_off[i0]);
      sd.host_label = at(d->host_label, d->host_node_off[i0]); sd.guest_parent = at(d->guest_parent, d->guest_node_off[i0]); sd.guest_label = at(d->guest_label, d->guest_node_off[i0]);
      sd.host_pred_off = sd.host_pred_adj = sd.guest_child_off = sd.guest_child_adj = nullptr;
      pt_batch* h = nullptr;
      int rc = pt_batch_create(&h, &sd, nullptr);
      if (rc != PT_OK) { fail(rc); return; }
      rc = pt_batch_contain(h, opt, result_bits + i0 / 32, status ? status + i0 : nullptr, PT_MEM_HOST, &cnt[(size_t)g], nullptr);
      if (rc != PT_OK) fail(rc);
      pt_batch_free(h);
    });
  }
  for (auto& w : workers) w.join();
  cudaSetDevice(prev_dev);
  for (int g = 0; g < G; ++g) if (rcs[(size_t)g] != PT_OK) return set_error(rcs[(size_t)g], "device %d: %s", g, msgs[(size_t)g].c_str());
  if (counters) {
    memset(counters, 0, sizeof(*counters));
    for (int g = 0; g < G; ++g) {
      const pt_counters& c = cnt[(size_t)g];
      counters->instances += c.instances; counters->displayed += c.displayed; counters->not_displayed += c.not_displayed; counters->errors += c.errors;
      counters->resolved_by_rules += c.resolved_by_rules; counters->branched += c.branched; counters->work_items += c.work_items;
      counters->rounds = std::max(counters->rounds, c.rounds); counters->cherry_reductions += c.cherry_reductions;
      counters->reticulated_cherries += c.reticulated_cherries; counters->arcs_deleted += c.arcs_deleted; counters->rule_passes += c.rule_passes;
    }
  }
  return PT_OK;
}

}  // extern "C"
