// csrc/comm.cu -- the path's only collective, inside the library: one process per GPU, the instances of a
// batch cut into contiguous ranges (boundaries multiples of 32, so that every rank owns whole result words); after the local
// solve ONE NCCL all-reduce(SUM) over the packed result words of the WHOLE batch (disjoint words: SUM == OR) and one over the
// counters.  The enumerator's cross-GPU early exit uses the same communicator (all-reduce(MAX) of one found word per chunk).
//
// The result exchange itself does not go through NCCL when the GPUs can map each other's memory (NVLink / NVSwitch peer access): every rank
// owns a small SYMMETRIC buffer, mapped into all peers through CUDA IPC handles at pt_comm_init; after its solve a rank's exchange kernel
// (xchg_kernel) stores its own result words and status bytes straight into every peer's buffer, raises a flag there, waits for the flags of
// the others and copies the assembled vector out -- one launch of one CTA, no staging, no reduction (the ranges are disjoint).  NCCL
// remains for bootstrap (the handle all-gather), for the counters, for the enumerator's found word, and as the fallback (PT_COMM_P2P=0).
//
// NCCL is reached through dlopen("libnccl.so.2"): in a torch process that is the NCCL torch has already loaded (one NCCL per
// process), in a plain C++ host program the system one; the library itself links against nothing but cudart.  The unique id
// travels through whatever the host program uses for bootstrap (bench.py: a torch.distributed broadcast; tests: a file).
#include <dlfcn.h>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>
#include <cstdlib>
#include "cuda_common.cuh"

namespace {

struct NcclId { char internal[128]; };  // ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES = 128)
typedef void* NcclComm;
enum { kNcclSum = 0, kNcclMax = 2, kNcclMin = 3, kNcclUint8 = 1, kNcclUint32 = 3, kNcclUint64 = 5 };  // ncclRedOp_t / ncclDataType_t values of nccl.h

// ---- the peer exchange -------------------------------------------------------------------------------------------------------------
constexpr int kMaxRanks = 64;
constexpr size_t kSymWords = (size_t)1 << 18;          // result words per half: batches up to 8.3 M instances
constexpr size_t kSymStatus = (size_t)1 << 22;         // status bytes per half: batches up to 4.2 M instances (larger ones: NCCL path)
constexpr size_t kSymHalf = kSymWords * 4 + kSymStatus;
struct SymHeader {                                      // at the start of every rank's symmetric buffer; [parity][source rank]
  uint32_t flag[2][kMaxRanks];
  unsigned arrive[2];                                   // CTAs of my own exchange kernel that have finished pushing
  long long range[2][kMaxRanks][4];                     // w0, wl (words), first, local (instances) of the source rank's shard
  uint32_t error;                                       // a wait ran into the watchdog
};
constexpr size_t kSymHeader = (sizeof(SymHeader) + 255) & ~(size_t)255;
constexpr size_t kSymBytes = kSymHeader + 2 * kSymHalf;
constexpr long long kXchgWatchdogCycles = 8000000000ll;  // ~4 s: a peer that never arrives becomes an error, not a hang

struct XchgArgs {
  unsigned char* peer[kMaxRanks];
  int nranks, rank, parity;
  uint32_t epoch;
  long long w0, wl, first, local, words, total;
  uint32_t* full_bits;
  uint8_t* full_status;   // may be null
};

__device__ __forceinline__ void st_sys_u32(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_sys_u32(const uint32_t* p) { uint32_t v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

// kXchgCtas CTAs per rank and step: push my range to every rank (myself included), flag, wait for everybody's flag, assemble.
// (one CTA took ~20 us longer per step than the NCCL all-reduce it replaces: 25 dependent L2 reads per thread in the assembly)
constexpr int kXchgCtas = 16;
__device__ __forceinline__ int owner_of(const long long (*rng)[4], int nranks, int which, long long lo, long long hi) {  // rank whose range holds [lo, hi), or -1
  for (int r = 0; r < nranks; ++r) if (lo >= rng[r][which] && hi <= rng[r][which] + rng[r][which + 1]) return r;
  return -1;
}
__global__ void __launch_bounds__(512) xchg_kernel(XchgArgs a) {
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x, gn = (long long)gridDim.x * blockDim.x;
  const int tid = threadIdx.x;
  const uint32_t* my_bits = a.full_bits + a.w0;
  const uint8_t* my_stat = a.full_status ? a.full_status + a.first : nullptr;
  const bool vec = my_stat && ((reinterpret_cast<uintptr_t>(my_stat) & 15) == 0);
  const long long n16 = vec ? a.local / 16 : 0;
  for (int p = 0; p < a.nranks; ++p) {
    unsigned char* half = a.peer[p] + kSymHeader + (size_t)a.parity * kSymHalf;
    uint32_t* dbits = reinterpret_cast<uint32_t*>(half) + a.w0;
    for (long long i = gtid; i < a.wl; i += gn) dbits[i] = my_bits[i];
    if (my_stat) {
      uint8_t* dst = half + kSymWords * 4 + a.first;   // first is a multiple of 32 and the halves are 256-byte aligned
      for (long long i = gtid; i < n16; i += gn) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(my_stat)[i];
      for (long long i = n16 * 16 + gtid; i < a.local; i += gn) dst[i] = my_stat[i];
    }
    if (gtid == 0) {
      SymHeader* h = reinterpret_cast<SymHeader*>(a.peer[p]);
      h->range[a.parity][a.rank][0] = a.w0; h->range[a.parity][a.rank][1] = a.wl; h->range[a.parity][a.rank][2] = a.first; h->range[a.parity][a.rank][3] = a.local;
    }
  }
  SymHeader* me = reinterpret_cast<SymHeader*>(a.peer[a.rank]);
  __threadfence_system();
  __syncthreads();
  // the CTA that arrives last (its atomic sees every other CTA's pushes behind their fences) raises my flag at every rank
  __shared__ int s_last, s_bad;
  if (tid == 0) {
    s_bad = 0;
    const unsigned prev = atomicAdd(&me->arrive[a.parity], 1u);
    s_last = prev == gridDim.x - 1;
    if (s_last) me->arrive[a.parity] = 0;   // this half is used again two steps from now
  }
  __syncthreads();
  if (s_last) {
    __threadfence_system();
    if (tid < a.nranks) st_sys_u32(&reinterpret_cast<SymHeader*>(a.peer[tid])->flag[a.parity][a.rank], a.epoch);
  }
  if (tid < a.nranks) {
    const long long t0 = clock64();
    while (ld_sys_u32(&me->flag[a.parity][tid]) != a.epoch) {
      if (clock64() - t0 > kXchgWatchdogCycles) { s_bad = 1; me->error = 1; break; }
      __nanosleep(100);
    }
  }
  __syncthreads();
  if (s_bad) return;
  __threadfence_system();
  // assemble: every word / byte comes from the rank whose announced range holds it; what nobody announced is zero (a sum of nothing)
  const unsigned char* half = a.peer[a.rank] + kSymHeader + (size_t)a.parity * kSymHalf;
  __shared__ long long s_rng[kMaxRanks][4];
  for (int i = tid; i < a.nranks * 4; i += blockDim.x) s_rng[i / 4][i % 4] = *reinterpret_cast<volatile long long*>(&me->range[a.parity][i / 4][i % 4]);
  __syncthreads();
  const uint32_t* sbits = reinterpret_cast<const uint32_t*>(half);
  for (long long i = gtid; i < a.words; i += gn) {
    if (i >= a.w0 && i < a.w0 + a.wl) continue;   // my own words are in place
    a.full_bits[i] = owner_of(s_rng, a.nranks, 0, i, i + 1) >= 0 ? __ldcg(&sbits[i]) : 0u;
  }
  if (a.full_status) {
    const uint8_t* sstat = half + kSymWords * 4;
    const long long padded = (a.total + 3) / 4 * 4;   // the caller pads the vector to a multiple of 4 (pt_gpu.h)
    const bool ovec = (reinterpret_cast<uintptr_t>(a.full_status) & 15) == 0;
    const long long c16 = ovec ? padded / 16 : 0;
    for (long long i = gtid; i < c16; i += gn) {
      const long long b0 = i * 16;
      if (b0 >= a.first && b0 + 16 <= a.first + a.local) continue;
      if (owner_of(s_rng, a.nranks, 2, b0, b0 + 16) >= 0) { reinterpret_cast<uint4*>(a.full_status)[i] = __ldcg(reinterpret_cast<const uint4*>(sstat) + i); continue; }
      for (int k = 0; k < 16; ++k) {   // a chunk that straddles ranges (ragged ends only)
        const long long b = b0 + k;
        if (b >= a.first && b < a.first + a.local) continue;
        a.full_status[b] = owner_of(s_rng, a.nranks, 2, b, b + 1) >= 0 ? __ldcg(&sstat[b]) : (uint8_t)0;
      }
    }
    for (long long b = c16 * 16 + gtid; b < padded; b += gn) {
      if (b >= a.first && b < a.first + a.local) continue;
      a.full_status[b] = owner_of(s_rng, a.nranks, 2, b, b + 1) >= 0 ? __ldcg(&sstat[b]) : (uint8_t)0;
    }
  }
}
struct NcclApi {
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  bool ok = false;
  char why[200] = {0};
};
NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, []() {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { snprintf(api.why, sizeof(api.why), "libnccl.so.2 not found: %s", dlerror()); return; }
    api.GetUniqueId = (int (*)(NcclId*))dlsym(h, "ncclGetUniqueId");
    api.CommInitRank = (int (*)(NcclComm*, int, NcclId, int))dlsym(h, "ncclCommInitRank");
    api.CommDestroy = (int (*)(NcclComm))dlsym(h, "ncclCommDestroy");
    api.AllReduce = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))dlsym(h, "ncclAllReduce");
    api.AllGather = (int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t))dlsym(h, "ncclAllGather");
    api.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    api.GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
    api.GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.AllGather && api.GetErrorString && api.GroupStart && api.GroupEnd;
    if (!api.ok) snprintf(api.why, sizeof(api.why), "libnccl.so.2 lacks an expected symbol");
  });
  return api;
}
#define PT_NCCL(call)                                                                                                       \
  do {                                                                                                                      \
    const int r__ = (call);                                                                                                 \
    if (r__ != 0) return ptgpu::set_error(PT_ERR_CUDA, "%s:%d: %s: NCCL error %d (%s)", __FILE__, __LINE__, #call, r__, nccl().GetErrorString(r__)); \
  } while (0)

}  // namespace

struct pt_comm {
  NcclComm comm = nullptr;
  int nranks = 1, rank = 0;
  unsigned long long* d_cnt = nullptr;  // staging for the counters
  // peer exchange: my symmetric buffer and every rank's, as mapped into this process (own included); p2p = all ranks have all mappings
  unsigned char* sym = nullptr;
  unsigned char* peer[kMaxRanks] = {};
  bool p2p = false;
  uint32_t epoch = 0;
};

// maps every rank's symmetric buffer into every other rank (CUDA IPC over NVLink peer access).  Collective: every rank calls it; the
// outcome is agreed on (all-reduce MIN), so either all ranks exchange through peer memory or all through NCCL.
static void comm_setup_p2p(pt_comm* c) {
  static const bool disabled = [] { const char* e = getenv("PT_COMM_P2P"); return e && atoi(e) == 0; }();
  int ok = (!disabled && c->nranks <= kMaxRanks) ? 1 : 0;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (ok && (cudaMalloc((void**)&c->sym, kSymBytes) != cudaSuccess || cudaMemset(c->sym, 0, kSymBytes) != cudaSuccess || cudaIpcGetMemHandle(&mine, c->sym) != cudaSuccess)) { cudaGetLastError(); ok = 0; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  unsigned char* d_h = nullptr;
  std::vector<cudaIpcMemHandle_t> all((size_t)c->nranks);
  if (cudaMalloc((void**)&d_h, 64 * (size_t)(c->nranks + 1)) != cudaSuccess) { cudaGetLastError(); d_h = nullptr; }
  bool gathered = false;
  if (d_h) {
    cudaMemcpy(d_h + 64 * (size_t)c->nranks, &mine, 64, cudaMemcpyHostToDevice);
    if (nccl().AllGather(d_h + 64 * (size_t)c->nranks, d_h, 64, kNcclUint8, c->comm, nullptr) == 0 && cudaStreamSynchronize(nullptr) == cudaSuccess &&
        cudaMemcpy(all.data(), d_h, 64 * (size_t)c->nranks, cudaMemcpyDeviceToHost) == cudaSuccess)
      gathered = true;
    else cudaGetLastError();
  }
  if (!gathered) ok = 0;
  if (ok) {
    int dev = 0; cudaGetDevice(&dev);
    for (int p = 0; p < c->nranks && ok; ++p) {
      if (p == c->rank) { c->peer[p] = c->sym; continue; }
      void* ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, all[(size_t)p], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
      c->peer[p] = (unsigned char*)ptr;
    }
  }
  // agree: everybody or nobody
  uint32_t flag = (uint32_t)ok;
  if (d_h && cudaMemcpy(d_h, &flag, 4, cudaMemcpyHostToDevice) == cudaSuccess && nccl().AllReduce(d_h, d_h, 1, kNcclUint32, kNcclMin, c->comm, nullptr) == 0 &&
      cudaStreamSynchronize(nullptr) == cudaSuccess && cudaMemcpy(&flag, d_h, 4, cudaMemcpyDeviceToHost) == cudaSuccess)
    c->p2p = flag != 0;
  else { cudaGetLastError(); c->p2p = false; }
  if (d_h) cudaFree(d_h);
  if (!c->p2p) {
    for (int p = 0; p < c->nranks; ++p) if (p != c->rank && c->peer[p]) { cudaIpcCloseMemHandle(c->peer[p]); c->peer[p] = nullptr; }
    // (the buffer itself stays until pt_comm_free: a peer that mapped it before the vote may still be closing its mapping)
  }
}

extern "C" void ptgpu_batch_outputs_prezeroed(pt_batch* b);   // tc_kernels.cu
using ptgpu::dev_alloc;
using ptgpu::dev_free;
using ptgpu::set_error;

extern "C" int pt_comm_allreduce(pt_comm* c, void* device_words, int64_t count, int elem_bytes, int op, void* stream);
// counters: summed over the ranks (rounds: the deepest shard); synchronises the stream
static int comm_reduce_counters(pt_comm* c, const pt_counters& local, pt_counters* total, cudaStream_t s) {
  void* stream = (void*)s;
    static_assert(sizeof(pt_counters) == 12 * sizeof(uint64_t), "pt_counters layout");
    uint64_t h[12];
    memcpy(h, &local, sizeof(h));
    const uint64_t my_rounds = local.rounds;
    PT_CUDA(cudaMemcpyAsync(c->d_cnt, h, sizeof(h), cudaMemcpyHostToDevice, s));
    if (int rc = pt_comm_allreduce(c, c->d_cnt, 12, 8, 0, stream)) return rc;
    PT_CUDA(cudaMemcpyAsync(h, c->d_cnt, sizeof(h), cudaMemcpyDeviceToHost, s));
    // rounds = the deepest shard, not a sum
    unsigned long long r = my_rounds;
    PT_CUDA(cudaMemcpyAsync(c->d_cnt + 16, &r, 8, cudaMemcpyHostToDevice, s));
    if (int rc = pt_comm_allreduce(c, c->d_cnt + 16, 1, 8, 1, stream)) return rc;
    PT_CUDA(cudaMemcpyAsync(&r, c->d_cnt + 16, 8, cudaMemcpyDeviceToHost, s));
    PT_CUDA(cudaStreamSynchronize(s));
    memcpy(total, h, sizeof(h));
    total->rounds = r;
  return PT_OK;
}

extern "C" {

int pt_comm_unique_id(void* id128) {
  if (!id128) return set_error(PT_ERR_INVALID, "pt_comm_unique_id: null argument");
  if (!nccl().ok) return set_error(PT_ERR_UNSUPPORTED, "pt_comm: %s", nccl().why);
  NcclId id;
  PT_NCCL(nccl().GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return PT_OK;
}

int pt_comm_init(pt_comm** out, int nranks, int rank, const void* id128) {
  if (!out || nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !id128)) return set_error(PT_ERR_INVALID, "pt_comm_init: bad argument");
  *out = nullptr;
  if (int rc = ptgpu::require_device()) return rc;
  pt_comm* c = new (std::nothrow) pt_comm();
  if (!c) return set_error(PT_ERR_NOMEM, "out of host memory");
  c->nranks = nranks; c->rank = rank;
  if (int rc = dev_alloc((void**)&c->d_cnt, 64 * sizeof(unsigned long long))) { delete c; return rc; }
  if (nranks > 1) {
    if (!nccl().ok) { dev_free(c->d_cnt); delete c; return set_error(PT_ERR_UNSUPPORTED, "pt_comm: %s", nccl().why); }
    NcclId id;
    memcpy(&id, id128, sizeof(id));
    const int r = nccl().CommInitRank(&c->comm, nranks, id, rank);
    if (r != 0) { dev_free(c->d_cnt); delete c; return set_error(PT_ERR_CUDA, "ncclCommInitRank failed: %d (%s)", r, nccl().GetErrorString(r)); }
    comm_setup_p2p(c);
  }
  *out = c;
  return PT_OK;
}

int pt_comm_peer_exchange(const pt_comm* c) { return c && c->p2p ? 1 : 0; }
int pt_comm_rank(const pt_comm* c) { return c ? c->rank : 0; }
int pt_comm_size(const pt_comm* c) { return c ? c->nranks : 1; }

// in-place all-reduce of DEVICE words: op 0 = sum, 1 = max; elem_bytes 4 or 8
int pt_comm_allreduce(pt_comm* c, void* device_words, int64_t count, int elem_bytes, int op, void* stream) {
  if (!c || !device_words || count < 0 || (elem_bytes != 4 && elem_bytes != 8) || (op != 0 && op != 1)) return set_error(PT_ERR_INVALID, "pt_comm_allreduce: bad argument");
  if (c->nranks == 1 || count == 0) return PT_OK;
  PT_NCCL(nccl().AllReduce(device_words, device_words, (size_t)count, elem_bytes == 4 ? kNcclUint32 : kNcclUint64, op == 0 ? kNcclSum : kNcclMax, c->comm, (cudaStream_t)stream));
  return PT_OK;
}

// The sharded containment step of one rank.  `b` holds this rank's contiguous range of a batch of `total_instances`; its first
// instance has global index `first_instance` (a multiple of 32).  full_bits: DEVICE vector of ceil(total/32) words for the
// WHOLE batch; full_status (optional): total bytes, DEVICE.  The local kernel writes straight into this rank's words (no
// staging copy); then one all-reduce(SUM) over all words (every other rank's words are zeroed here first) -- and one over the
// counters when `total` is given (HOST; this call then synchronises the stream).
int pt_batch_contain_sharded(pt_batch* b, pt_comm* c, const pt_options* opt, int64_t first_instance, int64_t total_instances, int64_t local_instances,
                             uint32_t* full_bits, uint8_t* full_status, pt_counters* total, void* stream) {
  if (!b || !c || !full_bits || first_instance < 0 || (first_instance & 31) || local_instances < 0 || first_instance + local_instances > total_instances)
    return set_error(PT_ERR_INVALID, "pt_batch_contain_sharded: bad argument (the first instance of a shard must be a multiple of 32)");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t words = (total_instances + 31) / 32, w0 = first_instance / 32, wl = (local_instances + 31) / 32;
  const bool use_p2p = c->p2p && c->nranks > 1 && (size_t)words <= kSymWords && (!full_status || (size_t)((total_instances + 3) / 4 * 4) <= kSymStatus);
  if (use_p2p) {
    // ---- exchange through peer memory: solve into my own words of the caller's vector, then one CTA pushes them to every rank ----
    pt_counters local;
    memset(&local, 0, sizeof(local));
    if (local_instances > 0) {
      if (int rc = pt_batch_contain(b, opt, full_bits + w0, full_status ? full_status + first_instance : nullptr, PT_MEM_DEVICE, total ? &local : nullptr, stream)) return rc;
    }
    XchgArgs a;
    memset(&a, 0, sizeof(a));
    for (int p = 0; p < c->nranks; ++p) a.peer[p] = c->peer[p];
    a.nranks = c->nranks; a.rank = c->rank; a.epoch = ++c->epoch; a.parity = (int)(a.epoch & 1u);
    a.w0 = w0; a.wl = local_instances > 0 ? wl : 0; a.first = first_instance; a.local = local_instances; a.words = words; a.total = total_instances;
    a.full_bits = full_bits; a.full_status = full_status;
    xchg_kernel<<<kXchgCtas, 512, 0, s>>>(a);
    PT_CUDA(cudaGetLastError());
    if (total) {
      if (int rc = comm_reduce_counters(c, local, total, s)) return rc;
      uint32_t err = 0;
      PT_CUDA(cudaMemcpy(&err, &reinterpret_cast<SymHeader*>(c->sym)->error, 4, cudaMemcpyDeviceToHost));
      if (err) return set_error(PT_ERR_CUDA, "pt_batch_contain_sharded: a rank did not arrive at the exchange (watchdog); results are incomplete");
    }
    return PT_OK;
  }
  // ---- NCCL path: every other rank's words / bytes are zero here (the local kernel clears and fills its own range) ----
  PT_CUDA(cudaMemsetAsync(full_bits, 0, (size_t)words * 4, s));
  if (full_status) PT_CUDA(cudaMemsetAsync(full_status, 0, (size_t)((total_instances + 3) / 4 * 4), s));
  pt_counters local;
  memset(&local, 0, sizeof(local));
  if (local_instances > 0) {
    ptgpu_batch_outputs_prezeroed(b);
    if (int rc = pt_batch_contain(b, opt, full_bits + w0, full_status ? full_status + first_instance : nullptr, PT_MEM_DEVICE, total ? &local : nullptr, stream)) return rc;
  }
  // status bytes are disjoint per rank as well and are summed as 32-bit words (the caller pads the vector to a multiple of 4,
  // pt_gpu.h); both all-reduces go out as ONE NCCL group (one launch)
  if (c->nranks > 1) {
    PT_NCCL(nccl().GroupStart());
    int rc = pt_comm_allreduce(c, full_bits, words, 4, 0, stream);
    if (rc == PT_OK && full_status) rc = pt_comm_allreduce(c, full_status, (total_instances + 3) / 4, 4, 0, stream);
    const int ge = nccl().GroupEnd();
    if (rc != PT_OK) return rc;
    if (ge != 0) return set_error(PT_ERR_CUDA, "ncclGroupEnd failed: %d (%s)", ge, nccl().GetErrorString(ge));
  }
  if (total) { if (int rc = comm_reduce_counters(c, local, total, s)) return rc; }
  return PT_OK;
}

int pt_comm_free(pt_comm* c) {
  if (!c) return PT_OK;
  cudaDeviceSynchronize();
  for (int p = 0; p < c->nranks; ++p) if (p != c->rank && c->peer[p]) cudaIpcCloseMemHandle(c->peer[p]);
  if (c->sym) cudaFree(c->sym);
  if (c->comm) nccl().CommDestroy(c->comm);
  dev_free(c->d_cnt);
  delete c;
  return PT_OK;
}

}  // extern "C"
