// csrc/iso_kernels.cu -- batched network isomorphism: pt_iso_batch (include/pt_gpu.h), standing in for
// make_iso_mapper(N0, N1, flags).check_isomorph() (utils/isomorphism.hpp:33-324, examples/iso.cpp:96-112).
//
// The reference keeps, per node of N0, the set of nodes of N1 it may map to, propagates along arcs and branches on a node
// with several possibilities (deep copy of the mapper per branch).  Data-parallel version, one thread block per pair:
//   1. colour refinement: every node carries a 64-bit colour, initialised from (in-degree, out-degree, counted label) and
//      refined with a commutative hash of its children's and its parents' colours -- the same function on both networks,
//      so an isomorphism maps every node to a node of the same colour in every round;
//   2. both colour arrays are sorted (bitonic, shared memory); different sequences => NOT isomorphic;
//   3. all colours different => the mapping is forced (k-th colour to k-th colour) and is VERIFIED arc by arc and label by
//      label => isomorphic iff the check passes.  If the pair is isomorphic the forced mapping is an isomorphism, so a
//      failed check is a proof of the contrary: hash collisions can only cost extra work, never a wrong answer;
//   4. a stable colouring with ties: individualise one node u of the smallest tied class of N0 against each node v of that
//      class in N1 (both get the same fresh colour) and append the |class| sub-problems to a device pool that the next
//      launch consumes -- the branching of check_isomorph() :222-232, level-synchronous like pt_batch_contain.
// Exact by construction (positives are verified, negatives are invariants); no CPU path.
#include <algorithm>
#include <vector>
#include "cuda_common.cuh"

namespace ptgpu {

typedef unsigned long long u64;

__device__ __forceinline__ u64 iso_mix(u64 z) { z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }

struct IsoGraphs {  // flat arrays of the whole batch (device pointers); graph g = 2 * pair + side
  const int64_t *node_off, *arc_off;
  const int32_t *succ_off, *succ_adj, *pred_off, *pred_adj, *label;
};
struct IsoPool {  // sub-problems: pair id + the two colour arrays (stride 2 * maxN)
  int32_t* pair; u64* colours; int32_t* count; int capacity;
};
enum { ISO_ST_OK = 0, ISO_ST_BAD_GRAPH = 2, ISO_ST_TOO_LARGE = 3, ISO_ST_POOL_OVERFLOW = 4 };

// one view of a graph of the batch
struct IsoG {
  int n, m; const int32_t *so, *sa, *po, *pa, *lab;
  __device__ __forceinline__ int outdeg(int v) const { return so[v + 1] - so[v]; }
  __device__ __forceinline__ int indeg(int v) const { return po[v + 1] - po[v]; }
  // the label that counts under `flags` (FLAG_MAP_* utils/isomorphism.hpp:10-13: 1 leaves, 2 internal tree nodes, 4 internal reticulations)
  __device__ __forceinline__ int eff_label(int v, int flags) const {
    const int od = outdeg(v), id = indeg(v);
    const int counts = od == 0 ? (flags & 1) : (id >= 2 ? (flags & 4) : (flags & 2));
    return counts ? lab[v] : -1;
  }
};
__device__ __forceinline__ IsoG iso_graph(const IsoGraphs& G, int64_t g) {
  IsoG r;
  const int64_t nb = G.node_off[g], ab = G.arc_off[g];
  r.n = (int)(G.node_off[g + 1] - nb); r.m = (int)(G.arc_off[g + 1] - ab);
  r.so = G.succ_off + nb + g; r.sa = G.succ_adj + ab; r.po = G.pred_off + nb + g; r.pa = G.pred_adj + ab; r.lab = G.label + nb;
  return r;
}

// bitonic sort of (key, idx) pairs, NP a power of two, ascending by key then idx (ties in key are what the caller looks for)
__device__ void iso_sort(u64* key, uint32_t* idx, int NP) {
  for (int k = 2; k <= NP; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < NP; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const u64 a = key[i], b = key[l]; const uint32_t ia = idx[i], ib = idx[l];
          const bool up = (i & k) == 0;
          const bool gt = a > b || (a == b && ia > ib);
          if (gt == up) { key[i] = b; key[l] = a; idx[i] = ib; idx[l] = ia; }
        }
      }
      __syncthreads();
    }
}

// pool mode: the num_items sub-problems of `src`; batch mode (src.pair == nullptr): item = pair
__global__ void __launch_bounds__(256) iso_kernel(IsoGraphs G, int64_t num_items, IsoPool src, IsoPool dst, IsoPool cold, int maxN, int NP, int flags, int depth,
                                                  uint32_t* __restrict__ result_bits, uint8_t* __restrict__ status, int32_t* __restrict__ ticket) {
  extern __shared__ __align__(16) unsigned char iso_smem[];
  // colours c[side][v] (current), k[side][*] (next colours / sort keys, NP each), idx[side][*]
  u64* c0 = reinterpret_cast<u64*>(iso_smem); u64* c1 = c0 + NP; u64* k0 = c1 + NP; u64* k1 = k0 + NP;
  uint32_t* i0 = reinterpret_cast<uint32_t*>(k1 + NP); uint32_t* i1 = i0 + NP;
  __shared__ int s_item, s_flag, s_cnt;
  __shared__ unsigned long long s_best;
  const int tid = threadIdx.x, nth = blockDim.x;
  const int64_t total = num_items;
  for (;;) {
    __syncthreads();
    if (tid == 0) s_item = atomicAdd(ticket, 1);
    __syncthreads();
    const int item = s_item;
    if (item >= total) break;
    const int pair = src.pair ? src.pair[item] : item;
    if (pair < 0) continue;  // a cold slot reserved by a branching whose hot reservation failed (marked empty there)
    if (src.pair) {  // another branch may already have proved this pair (one thread looks, so that the block cannot diverge)
      if (tid == 0) s_flag = (int)((*(volatile uint32_t*)&result_bits[pair >> 5] >> (pair & 31)) & 1u);
      __syncthreads();
      if (s_flag) continue;
    }
    const IsoG A = iso_graph(G, 2 * (int64_t)pair), B = iso_graph(G, 2 * (int64_t)pair + 1);
    if (A.n != B.n || A.m != B.m) continue;  // isomorphism.hpp:201: different sizes => not isomorphic
    const int n = A.n;
    if (n > maxN) { if (tid == 0) status[pair] = ISO_ST_TOO_LARGE; continue; }
    if (n == 0) { if (tid == 0) atomicOr(&result_bits[pair >> 5], 1u << (pair & 31)); continue; }
    // ---- colours: fresh, or those of the sub-problem ----
    if (src.pair) {
      const u64* pc = src.colours + (size_t)item * 2 * maxN;
      for (int v = tid; v < n; v += nth) { c0[v] = pc[v]; c1[v] = pc[maxN + v]; }
    } else {
      if (tid == 0) s_flag = 0;
      __syncthreads();
      for (int v = tid; v < 2 * n; v += nth) {
        const IsoG& X = v < n ? A : B; const int x = v < n ? v : v - n;
        const int od = X.so[x + 1] - X.so[x], id = X.po[x + 1] - X.po[x];
        if (od < 0 || id < 0 || X.so[x] < 0 || X.so[x + 1] > X.m || X.po[x] < 0 || X.po[x + 1] > X.m) s_flag = 1;
        const u64 h = iso_mix(iso_mix(((u64)(unsigned)id << 32) | (unsigned)od) + 0x9E3779B97F4A7C15ull * (u64)(unsigned)(X.eff_label(x, flags) + 1));
        (v < n ? c0 : c1)[x] = h;
      }
      for (int e = tid; e < 2 * A.m; e += nth) {
        const IsoG& X = e < A.m ? A : B; const int x = e < A.m ? e : e - A.m;
        if (X.sa[x] < 0 || X.sa[x] >= n || X.pa[x] < 0 || X.pa[x] >= n) s_flag = 1;
      }
      __syncthreads();
      if (s_flag) { if (tid == 0) status[pair] = ISO_ST_BAD_GRAPH; continue; }
    }
    __syncthreads();
    // ---- refine in blocks of rounds until the number of colours stops growing ----
    int prev_distinct = -1, verdict = -1;  // verdict: 0 not isomorphic, 1 isomorphic, 2 branch
    for (int block = 0; verdict < 0; ++block) {
      for (int round = 0; round < 4; ++round) {
        for (int v = tid; v < 2 * n; v += nth) {
          const IsoG& X = v < n ? A : B; const int x = v < n ? v : v - n;
          const u64* c = v < n ? c0 : c1;
          u64 down = 0, up = 0;
          for (int j = X.so[x]; j < X.so[x + 1]; ++j) down += iso_mix(c[X.sa[j]] + 0xD6E8FEB86659FD93ull);
          for (int j = X.po[x]; j < X.po[x + 1]; ++j) up += iso_mix(c[X.pa[j]] + 0xA0761D6478BD642Full);
          (v < n ? k0 : k1)[x] = iso_mix(c[x] ^ iso_mix(down + 0x1B03738712FAD5C9ull) ^ iso_mix(up + 0x2545F4914F6CDD1Dull) * 3);
        }
        __syncthreads();
        for (int v = tid; v < n; v += nth) { c0[v] = k0[v]; c1[v] = k1[v]; }
        __syncthreads();
      }
      // sort both colour arrays
      for (int v = tid; v < NP; v += nth) {
        k0[v] = v < n ? c0[v] : ~0ull; k1[v] = v < n ? c1[v] : ~0ull; i0[v] = (uint32_t)v; i1[v] = (uint32_t)v;
      }
      if (tid == 0) { s_flag = 0; s_cnt = 0; }
      __syncthreads();
      iso_sort(k0, i0, NP);
      iso_sort(k1, i1, NP);
      int mism = 0, distinct = 0;
      for (int k = tid; k < n; k += nth) { mism |= k0[k] != k1[k]; distinct += (k == 0 || k0[k] != k0[k - 1]); }
      if (mism) s_flag = 1;
      if (distinct) atomicAdd(&s_cnt, distinct);
      __syncthreads();
      const int d = s_cnt;
      if (s_flag) verdict = 0;
      else if (d == n) {
        // ---- forced mapping: verify it ----
        uint32_t* map = reinterpret_cast<uint32_t*>(c0);  // the colours are no longer needed
        __syncthreads();
        for (int k = tid; k < n; k += nth) map[i0[k]] = i1[k];
        if (tid == 0) s_flag = 0;
        __syncthreads();
        int bad = 0;
        for (int u = tid; u < n; u += nth) {
          const int w = (int)map[u];
          if (A.outdeg(u) != B.outdeg(w) || A.indeg(u) != B.indeg(w) || A.eff_label(u, flags) != B.eff_label(w, flags)) { bad = 1; continue; }
          for (int j = A.so[u]; j < A.so[u + 1]; ++j) {  // every arc (u, c) needs the arc (w, map c), with the same multiplicity
            const int cw = (int)map[A.sa[j]];
            int have = 0, need = 0;
            for (int t = B.so[w]; t < B.so[w + 1]; ++t) have += B.sa[t] == cw;
            for (int t = A.so[u]; t < A.so[u + 1]; ++t) need += A.sa[t] == A.sa[j];
            if (have != need) bad = 1;
          }
        }
        if (bad) s_flag = 1;
        __syncthreads();
        verdict = s_flag ? 0 : 1;
      } else if (d == prev_distinct || block >= n) verdict = 2;
      prev_distinct = d;
      __syncthreads();
    }
    if (verdict == 1) { if (tid == 0) atomicOr(&result_bits[pair >> 5], 1u << (pair & 31)); continue; }
    if (verdict == 0) continue;
    // ---- ties: branch on the smallest tied class (sorted keys k0 / k1, nodes i0 / i1) ----
    if (tid == 0) s_best = ~0ull;
    __syncthreads();
    for (int k = tid; k < n; k += nth) {
      if ((k == 0 || k0[k] != k0[k - 1]) && k + 1 < n && k0[k + 1] == k0[k]) {
        int len = 2;
        while (k + len < n && k0[k + len] == k0[k]) ++len;
        atomicMin(&s_best, ((u64)(unsigned)len << 32) | (unsigned)k);
      }
    }
    __syncthreads();
    const int len = (int)(s_best >> 32), kb = (int)(s_best & 0xffffffffu);
    // the first candidate goes to the "hot" pool (searched next: depth first), its siblings to the "cold" one (searched only
    // if the hot branch fails): for an isomorphic pair whose tied class is an orbit the first descent already succeeds
    __shared__ int s_cold;
    // both reservations or neither, and a counter only ever advances over slots that WILL be written: the host pushes
    // min(count, capacity) slots of each pool onto the stack, so a counted but unwritten slot would be read as a pair index
    if (tid == 0) {
      int hot = -1, cb = -1;
      for (int old = *(volatile int*)cold.count;;) {
        if (old + len - 1 > cold.capacity) break;
        const int seen = atomicCAS(cold.count, old, old + len - 1);
        if (seen == old) { cb = old; break; }
        old = seen;
      }
      if (cb >= 0) {
        for (int old = *(volatile int*)dst.count;;) {
          if (old + 1 > dst.capacity) break;
          const int seen = atomicCAS(dst.count, old, old + 1);
          if (seen == old) { hot = old; break; }
          old = seen;
        }
        // the cold slots cannot be handed back (others may have reserved behind them): mark them empty instead
        if (hot < 0) for (int j = 0; j < len - 1; ++j) cold.pair[cb + j] = -1;
      }
      s_item = hot; s_cold = cb;
    }
    __syncthreads();
    const int hot_slot = s_item, cold_base = s_cold;
    if (hot_slot < 0 || cold_base < 0) { if (tid == 0) status[pair] = ISO_ST_POOL_OVERFLOW; continue; }
    const int u = (int)i0[kb];
    const u64 fresh = iso_mix(k0[kb] + 0x632BE59BD9B4E019ull * (u64)(depth + 1));
    for (int j = 0; j < len; ++j) {
      const int v = (int)i1[kb + j];
      u64* pc = j == 0 ? dst.colours + (size_t)hot_slot * 2 * maxN : cold.colours + (size_t)(cold_base + j - 1) * 2 * maxN;
      for (int x = tid; x < n; x += nth) { pc[x] = x == u ? fresh : c0[x]; pc[maxN + x] = x == v ? fresh : c1[x]; }
      if (tid == 0) { if (j == 0) dst.pair[hot_slot] = pair; else cold.pair[cold_base + j - 1] = pair; }
    }
  }
}

// a pair that some branch proved isomorphic is fine even if another of its branches ran out of pool space
__global__ void iso_status_kernel(int64_t P, const uint32_t* __restrict__ bits, uint8_t* __restrict__ status, int unresolved_status) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= P) return;
  if ((bits[i >> 5] >> (i & 31)) & 1u) status[i] = ISO_ST_OK;
  else if (unresolved_status && status[i] == ISO_ST_OK) status[i] = (uint8_t)unresolved_status;
}

static int iso_upload(void** d, const void* h, size_t bytes, cudaStream_t s) {
  if (int rc = dev_alloc(d, std::max<size_t>(bytes, 16))) return rc;
  if (bytes) PT_CUDA(cudaMemcpyAsync(*d, h, bytes, cudaMemcpyHostToDevice, s));
  return PT_OK;
}

}  // namespace ptgpu

using namespace ptgpu;

extern "C" int pt_iso_batch(const pt_iso_desc* d, uint32_t* result_bits, uint8_t* status, pt_mem out_mem, void* stream) {
  if (!d || !result_bits) return set_error(PT_ERR_INVALID, "pt_iso_batch: null argument");
  const int64_t P = d->num_pairs;
  if (P < 0 || P > (int64_t)INT32_MAX / 4) return set_error(PT_ERR_INVALID, "pt_iso_batch: bad num_pairs");
  if (!d->node_off || !d->arc_off || !d->succ_off || !d->pred_off || !d->label) return set_error(PT_ERR_INVALID, "pt_iso_batch: a required array is null");
  // the adjacency arrays may only be null when there is not a single arc (HOST input can be checked here)
  if ((!d->succ_adj || !d->pred_adj) && P > 0 && (d->mem != PT_MEM_HOST || d->arc_off[2 * P] > 0))
    return set_error(PT_ERR_INVALID, "pt_iso_batch: succ_adj / pred_adj is null");
  if (int rc = require_device()) return rc;
  DeviceProps dp;
  if (int rc = device_props(&dp)) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t words = (size_t)((P + 31) / 32) + 1;
  IsoGraphs G{};
  std::vector<void*> owned;
  auto cleanup = [&]() { cudaStreamSynchronize(s); for (void* p : owned) dev_free(p); };
  int rc = PT_OK;
  int32_t maxN = d->max_nodes;
  if (d->mem == PT_MEM_HOST) {
    for (int64_t g = 0; g < 2 * P; ++g)
      if (d->node_off[g + 1] < d->node_off[g] || d->arc_off[g + 1] < d->arc_off[g]) return set_error(PT_ERR_INVALID, "pt_iso_batch: offset tables must be non-decreasing");
    if (P > 0 && (d->node_off[0] != 0 || d->arc_off[0] != 0)) return set_error(PT_ERR_INVALID, "pt_iso_batch: offset tables must start at 0");
    const int64_t NN = P ? d->node_off[2 * P] : 0, MM = P ? d->arc_off[2 * P] : 0;
    if (maxN <= 0) for (int64_t g = 0; g < 2 * P; ++g) maxN = std::max<int32_t>(maxN, (int32_t)std::min<int64_t>(d->node_off[g + 1] - d->node_off[g], INT32_MAX));
    void* p[7] = {};
    const void* h[7] = {d->node_off, d->arc_off, d->succ_off, d->succ_adj, d->pred_off, d->pred_adj, d->label};
    const size_t bytes[7] = {(size_t)(2 * P + 1) * 8, (size_t)(2 * P + 1) * 8, (size_t)(NN + 2 * P) * 4, (size_t)MM * 4, (size_t)(NN + 2 * P) * 4, (size_t)MM * 4, (size_t)NN * 4};
    for (int k = 0; k < 7 && rc == PT_OK; ++k) { rc = iso_upload(&p[k], h[k], bytes[k], s); if (p[k]) owned.push_back(p[k]); }
    if (rc) { cleanup(); return rc; }
    G.node_off = (const int64_t*)p[0]; G.arc_off = (const int64_t*)p[1]; G.succ_off = (const int32_t*)p[2]; G.succ_adj = (const int32_t*)p[3];
    G.pred_off = (const int32_t*)p[4]; G.pred_adj = (const int32_t*)p[5]; G.label = (const int32_t*)p[6];
  } else {
    if (maxN <= 0) return set_error(PT_ERR_INVALID, "pt_iso_batch: max_nodes is required for DEVICE input");
    G.node_off = d->node_off; G.arc_off = d->arc_off; G.succ_off = d->succ_off; G.succ_adj = d->succ_adj; G.pred_off = d->pred_off; G.pred_adj = d->pred_adj; G.label = d->label;
  }
  maxN = std::max(maxN, 1);
  int NP = 1; while (NP < maxN) NP <<= 1;
  const size_t smem = (size_t)NP * (4 * 8 + 2 * 4);
  uint32_t* bits = nullptr; uint8_t* stat = nullptr; int32_t* ctl = nullptr;  // ctl: [0] ticket [1] pool A count [2] pool B count
  const bool dev_out = out_mem == PT_MEM_DEVICE;
  if (dev_out) { bits = result_bits; stat = status; }
  else {
    if ((rc = dev_alloc((void**)&bits, words * 4))) { cleanup(); return rc; } owned.push_back(bits);
  }
  uint8_t* stat_own = nullptr;
  if (!dev_out || !stat) { if ((rc = dev_alloc((void**)&stat_own, (size_t)P + 1))) { cleanup(); return rc; } owned.push_back(stat_own); stat = stat_own; }
  if ((rc = dev_alloc((void**)&ctl, 64))) { cleanup(); return rc; } owned.push_back(ctl);
  auto fail = [&](cudaError_t e, const char* what) { cleanup(); return set_error(PT_ERR_CUDA, "pt_iso_batch: %s: %s", what, cudaGetErrorString(e)); };
  cudaError_t e;
  if ((e = cudaMemsetAsync(bits, 0, words * 4, s)) != cudaSuccess || (e = cudaMemsetAsync(stat, 0, (size_t)P + 1, s)) != cudaSuccess ||
      (e = cudaMemsetAsync(ctl, 0, 64, s)) != cudaSuccess) return fail(e, "memset");
  if (P > 0 && smem > (size_t)dp.smem_optin) {
    // networks beyond the shared-memory sort: reported per pair, not attempted
    if ((e = cudaMemsetAsync(stat, ISO_ST_TOO_LARGE, (size_t)P, s)) != cudaSuccess) return fail(e, "memset");
  } else if (P > 0) {
    if ((e = cudaFuncSetAttribute(iso_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return fail(e, "cudaFuncSetAttribute");
    // Sub-problems of tied colourings live on a device STACK: every round pops the newest `width` of them, and what they emit
    // (into a scratch pool) is pushed back on top.  Depth-first in effect: an isomorphic pair with many symmetries (every
    // candidate of a tied class works) is settled after one descent instead of a breadth that doubles per level, and the
    // memory is bounded by depth x width.
    const size_t slot_bytes = (size_t)2 * maxN * 8;
    const int ctas_per_sm = std::max(1, std::min(8, (int)((size_t)dp.smem_optin / std::max<size_t>(smem, 1024))));
    const int width = dp.sms * ctas_per_sm * 2;
    const int cap_t = (int)std::max<size_t>(64, std::min<size_t>((size_t)std::max<int64_t>(8 * (int64_t)width, 2 * P), ((size_t)256 << 20) / slot_bytes));
    const int cap_s = (int)std::max<size_t>(64, std::min<size_t>((size_t)std::max<int64_t>(1 << 16, 4 * P), ((size_t)1 << 30) / slot_bytes));
    IsoPool stack{}, scratch{}, coldp{};
    {
      void *sp = nullptr, *sc = nullptr, *tp = nullptr, *tc = nullptr, *cp = nullptr, *cc = nullptr;
      if ((rc = dev_alloc(&sp, (size_t)cap_s * 4)) == PT_OK) owned.push_back(sp);
      if (rc == PT_OK && (rc = dev_alloc(&sc, slot_bytes * cap_s)) == PT_OK) owned.push_back(sc);
      if (rc == PT_OK && (rc = dev_alloc(&tp, (size_t)cap_t * 4)) == PT_OK) owned.push_back(tp);
      if (rc == PT_OK && (rc = dev_alloc(&tc, slot_bytes * cap_t)) == PT_OK) owned.push_back(tc);
      if (rc == PT_OK && (rc = dev_alloc(&cp, (size_t)cap_t * 4)) == PT_OK) owned.push_back(cp);
      if (rc == PT_OK && (rc = dev_alloc(&cc, slot_bytes * cap_t)) == PT_OK) owned.push_back(cc);
      if (rc) { cleanup(); return rc; }
      coldp.pair = (int32_t*)cp; coldp.colours = (u64*)cc; coldp.count = ctl + 2; coldp.capacity = cap_t;
      stack.pair = (int32_t*)sp; stack.colours = (u64*)sc; stack.count = nullptr; stack.capacity = cap_s;
      scratch.pair = (int32_t*)tp; scratch.colours = (u64*)tc; scratch.count = ctl + 1; scratch.capacity = cap_t;
    }
    int64_t top = 0, hot_on_top = 0;
    int unresolved = 0;  // status for pairs still undecided when the search is cut off
    const int64_t max_rounds = 200000;
    for (int64_t round = 0;; ++round) {
      IsoPool src{}; int64_t items; 
      if (round == 0) { src.pair = nullptr; items = P; }
      else {
        if (top == 0) break;
        if (round > max_rounds) { unresolved = ISO_ST_POOL_OVERFLOW; break; }
        items = std::min<int64_t>(hot_on_top > 0 ? hot_on_top : top, width);  // the newest first candidates, else whatever is left
        src = stack; src.pair = stack.pair + (top - items); src.colours = stack.colours + (size_t)(top - items) * 2 * maxN;
      }
      const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(items, (int64_t)dp.sms * ctas_per_sm));
      iso_kernel<<<grid, 256, smem, s>>>(G, items, src, scratch, coldp, maxN, NP, d->flags, (int)(round & 0x3fffffff), bits, stat, ctl);
      if ((e = cudaGetLastError()) != cudaSuccess) return fail(e, "iso_kernel launch");
      int32_t h[4];
      if ((e = cudaMemcpyAsync(h, ctl, 16, cudaMemcpyDeviceToHost, s)) != cudaSuccess || (e = cudaStreamSynchronize(s)) != cudaSuccess) return fail(e, "iso_kernel");
      const int64_t hot = std::min<int64_t>(h[1], cap_t), cold_n = std::min<int64_t>(h[2], cap_t);
      if (round > 0) top -= items;
      if (top + hot + cold_n > cap_s) { unresolved = ISO_ST_POOL_OVERFLOW; break; }
      auto push = [&](const IsoPool& from, int64_t cnt) -> cudaError_t {
        if (cnt <= 0) return cudaSuccess;
        cudaError_t pe = cudaMemcpyAsync(stack.pair + top, from.pair, (size_t)cnt * 4, cudaMemcpyDeviceToDevice, s);
        if (pe == cudaSuccess) pe = cudaMemcpyAsync(stack.colours + (size_t)top * 2 * maxN, from.colours, (size_t)cnt * slot_bytes, cudaMemcpyDeviceToDevice, s);
        top += cnt;
        return pe;
      };
      if ((e = push(coldp, cold_n)) != cudaSuccess || (e = push(scratch, hot)) != cudaSuccess) return fail(e, "push");  // hot ones on top
      hot_on_top = hot;
      if ((e = cudaMemsetAsync(ctl, 0, 12, s)) != cudaSuccess) return fail(e, "memset");
    }
    iso_status_kernel<<<(unsigned)((P + 255) / 256), 256, 0, s>>>(P, bits, stat, unresolved);
    if ((e = cudaGetLastError()) != cudaSuccess) return fail(e, "iso_status_kernel launch");
  }
  if (!dev_out) {
    if ((e = cudaMemcpyAsync(result_bits, bits, (size_t)((P + 31) / 32) * 4, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return fail(e, "copy");
    if (status && (e = cudaMemcpyAsync(status, stat, (size_t)P, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return fail(e, "copy");
  }
  if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return fail(e, "sync");
  cleanup();
  return PT_OK;
}
