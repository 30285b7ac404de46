// csrc/comm.cu -- the path's only collective, inside the library: one process per GPU, the instances of a
// batch cut into contiguous ranges (boundaries multiples of 32, so that every rank owns whole result words); after the local
// solve ONE NCCL all-reduce(SUM) over the packed result words of the WHOLE batch (disjoint words: SUM == OR) and one over the
// counters.  The enumerator's cross-GPU early exit uses the same communicator (all-reduce(MAX) of one found word per chunk).
//
// NCCL is reached through dlopen("libnccl.so.2"): in a torch process that is the NCCL torch has already loaded (one NCCL per
// process), in a plain C++ host program the system one; the library itself links against nothing but cudart.  The unique id
// travels through whatever the host program uses for bootstrap (bench.py: a torch.distributed broadcast; tests: a file).
#include <dlfcn.h>
#include <cstring>
#include <mutex>
#include <new>
#include "cuda_common.cuh"

namespace {

struct NcclId { char internal[128]; };  // ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES = 128)
typedef void* NcclComm;
enum { kNcclSum = 0, kNcclMax = 2, kNcclUint32 = 3, kNcclUint64 = 5 };  // ncclRedOp_t / ncclDataType_t values of nccl.h
struct NcclApi {
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
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
    api.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    api.GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
    api.GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.GetErrorString && api.GroupStart && api.GroupEnd;
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
};

extern "C" void ptgpu_batch_outputs_prezeroed(pt_batch* b);   // tc_kernels.cu
using ptgpu::dev_alloc;
using ptgpu::dev_free;
using ptgpu::set_error;

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
  }
  *out = c;
  return PT_OK;
}

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
  // every other rank's words / bytes are zero here (the local kernel clears and fills its own range)
  (void)w0; (void)wl;
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
  if (total) {
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
  }
  return PT_OK;
}

int pt_comm_free(pt_comm* c) {
  if (!c) return PT_OK;
  cudaDeviceSynchronize();
  if (c->comm) nccl().CommDestroy(c->comm);
  dev_free(c->d_cnt);
  delete c;
  return PT_OK;
}

}  // extern "C"
