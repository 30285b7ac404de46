// csrc/enum_kernels.cu -- exhaustive switching enumeration for ONE instance over a range of the mixed-radix switching
// index: one switching per THREAD (32 display trees in flight per warp, all lanes walking the same node order so the
// structure reads are warp-uniform broadcasts and the per-thread accumulators are coalesced), compared with the guest
// by a canonical 64-bit Merkle hash of unordered leaf-labelled trees.
//
// No counterpart in the reference: it branches on the parents of one reticulation and recurses on deep copies
// (utils/containment.hpp:1225-1266); the verdict semantics are the same (exists a switching).  This kernel is the
// engine for general networks that do not reduce (BASELINE config 5) and shards by index range across GPUs.
// pt_batch_contain does NOT use it: the rule+branch path there is exact by construction; here equality is decided by
// a 64-bit hash (collision probability ~2^-64 per comparison).
#include <algorithm>
#include <new>
#include <vector>
#include "cuda_common.cuh"

namespace ptgpu {

typedef unsigned long long u64;

__host__ __device__ __forceinline__ u64 mix64(u64 z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ u64 leaf_hash(int label) { u64 h = mix64(0x9E3779B97F4A7C15ull * (u64)(label + 1)); return h ? h : 1; }
__host__ __device__ __forceinline__ u64 child_term(u64 v) { return mix64(v + 0xD6E8FEB86659FD93ull); }
__host__ __device__ __forceinline__ u64 node_hash(u64 sum_terms) { u64 h = mix64(sum_terms ^ 0xA0761D6478BD642Full); return h ? h : 1; }

// node i in children-first order
struct EnumNode { int32_t parent;      // ACCUMULATOR index of the fixed parent, -1 root, <= -2: multi-parent, digit = -2 - parent
                  int32_t label;       // leaf label or -1
                  int32_t acc;         // accumulator index of this node (nodes with out-arcs only: a leaf never receives), else -1
                  int32_t pad; };

__global__ void __launch_bounds__(256) enum_kernel(int n, int nacc, const EnumNode* __restrict__ nodes, int r, const int32_t* __restrict__ radix,
                                                   const int32_t* __restrict__ par_off, const int32_t* __restrict__ par_pos,
                                                   u64 begin, u64 end, u64 guest_hash, int stop_at_first,
                                                   u64* __restrict__ S1, u64* __restrict__ S2, uint32_t* __restrict__ cnt,
                                                   u64* __restrict__ result /* [0]=count [1]=first [2]=stop flag */) {
  extern __shared__ int32_t sm[];
  EnumNode* snode = reinterpret_cast<EnumNode*>(sm);
  int32_t* sradix = sm + 4 * n; int32_t* soff = sradix + r; int32_t* spar = soff + r + 1;
  for (int i = threadIdx.x; i < n; i += blockDim.x) snode[i] = nodes[i];
  for (int i = threadIdx.x; i < r; i += blockDim.x) sradix[i] = radix[i];
  for (int i = threadIdx.x; i <= r; i += blockDim.x) soff[i] = par_off[i];
  for (int i = threadIdx.x; i < par_off[r]; i += blockDim.x) spar[i] = par_pos[i];
  __syncthreads();
  const u64 T = (u64)gridDim.x * blockDim.x, tid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = 0; i < nacc; ++i) { S1[(u64)i * T + tid] = 0; S2[(u64)i * T + tid] = 0; cnt[(u64)i * T + tid] = 0; }
  u64 found = 0, first = ~0ull;
  for (u64 s = begin + tid; s < end; s += T) {
    if (stop_at_first && *(volatile u64*)&result[2]) break;
    u64 val = 0;
    for (int i = 0; i < n; ++i) {
      const EnumNode nd = snode[i];
      if (nd.acc < 0) val = nd.label >= 0 ? leaf_hash(nd.label) : 0;   // a node without out-arcs: its leaf hash, or nothing
      else {
        const u64 k = (u64)nd.acc * T + tid;
        const uint32_t c = cnt[k];
        if (c == 0) val = 0;
        else if (c == 1) val = S1[k];
        else val = node_hash(S2[k]);
        if (c) { S1[k] = 0; S2[k] = 0; cnt[k] = 0; }
      }
      if (val && nd.parent != -1) {
        int p = nd.parent;
        if (p <= -2) {
          // digit of this multi-parent node: peel the mixed-radix index (digits are few; recomputing beats storing them)
          const int d = -2 - p; u64 x = s;
          for (int j = 0; j < d; ++j) x /= (u64)sradix[j];
          p = spar[soff[d] + (int)(x % (u64)sradix[d])];
        }
        const u64 kp = (u64)p * T + tid;
        S1[kp] += val; S2[kp] += child_term(val); cnt[kp] += 1;
      }
    }
    if (val == guest_hash) { ++found; if (s < first) first = s; if (stop_at_first) result[2] = 1; }
  }
  for (int d = 16; d; d >>= 1) { found += __shfl_xor_sync(0xffffffffu, found, d); const u64 o = __shfl_xor_sync(0xffffffffu, first, d); first = o < first ? o : first; }
  if ((threadIdx.x & 31) == 0) { if (found) atomicAdd(&result[0], found); if (first != ~0ull) atomicMin(&result[1], first); }
}

}  // namespace ptgpu

using namespace ptgpu;

// ---- persistent handle: everything that does not depend on the index range is done once (host preprocessing, upload,
// accumulators); a run is then a kernel launch per chunk of the range, with the found word looked at between chunks ----------
struct pt_enum {
  int n = 0, nacc = 1, r = 0, grid = 0;
  uint64_t total = 1;
  int trivial = 0;            // 0 enumerate; 1 never displayed (a guest label the host lacks); 2 always displayed (<= 2 common labels)
  u64 guest_hash = 0;
  size_t smem = 0;
  EnumNode* d_nodes = nullptr; int32_t *d_radix = nullptr, *d_off = nullptr, *d_par = nullptr; u64 *d_S1 = nullptr, *d_S2 = nullptr, *d_res = nullptr; uint32_t* d_cnt = nullptr;
  // for the exact check of a witness: the host network and the guest as given
  std::vector<int32_t> succ_off, succ_adj, host_label, guest_parent, guest_label, retis, indeg;
};

static void enum_release(pt_enum* h) {
  if (!h) return;
  cudaDeviceSynchronize();
  dev_free(h->d_nodes); dev_free(h->d_radix); dev_free(h->d_off); dev_free(h->d_par); dev_free(h->d_S1); dev_free(h->d_S2); dev_free(h->d_cnt); dev_free(h->d_res);
  delete h;
}

// Is the guest displayed under switching `s` (mixed-radix index, digit i = which in-arc of the i-th multi-parent node is kept)?
// Decided EXACTLY: the switching turns the host into a tree, and pt_batch_contain on (that tree, guest) is exact by
// construction (no hashing on that path).  Used to confirm the witness the hash-based kernel reports.
static int enum_verify(const pt_enum* h, uint64_t s, int* displayed, void* stream) {
  const int n = h->n;
  std::vector<int32_t> keep_arc(n, -1);   // for a multi-parent node: the position (among its in-arcs, arc order) that is kept
  uint64_t x = s;
  for (size_t i = 0; i < h->retis.size(); ++i) { const uint64_t rad = (uint64_t)h->indeg[h->retis[i]]; keep_arc[h->retis[i]] = (int32_t)(x % rad); x /= rad; }
  std::vector<int32_t> seen(n, 0), off(n + 1, 0), adj;
  std::vector<std::vector<int32_t>> kids(n);
  for (int u = 0; u < n; ++u) for (int e = h->succ_off[u]; e < h->succ_off[u + 1]; ++e) {
    const int v = h->succ_adj[e];
    if (keep_arc[v] < 0 || seen[v]++ == keep_arc[v]) kids[u].push_back(v);
  }
  for (int u = 0; u < n; ++u) { off[u + 1] = off[u] + (int32_t)kids[u].size(); adj.insert(adj.end(), kids[u].begin(), kids[u].end()); }
  const int64_t hn[2] = {0, n}, ha[2] = {0, (int64_t)adj.size()}, gn[2] = {0, (int64_t)h->guest_parent.size()};
  pt_batch_desc d{};
  d.num_instances = 1; d.mem = PT_MEM_HOST; d.host_node_off = hn; d.host_arc_off = ha; d.guest_node_off = gn;
  d.host_succ_off = off.data(); d.host_succ_adj = adj.data(); d.host_label = h->host_label.data(); d.guest_parent = h->guest_parent.data(); d.guest_label = h->guest_label.data();
  pt_batch* b = nullptr;
  if (int rc = pt_batch_create(&b, &d, stream)) return rc;
  uint32_t bits[2] = {0, 0}; uint8_t st[4] = {0, 0, 0, 0};
  const int rc = pt_batch_contain(b, nullptr, bits, st, PT_MEM_HOST, nullptr, stream);
  pt_batch_free(b);
  if (rc) return rc;
  if (st[0]) return set_error(PT_ERR_INVALID, "pt_enum: the exact check of switching %llu failed with instance status %d", (unsigned long long)s, (int)st[0]);
  *displayed = (int)(bits[0] & 1u);
  return PT_OK;
}

extern "C" {

int pt_enum_free(pt_enum* h) { enum_release(h); return PT_OK; }

int pt_enum_create(pt_enum** out, int32_t n, int32_t m, const int32_t* succ_off, const int32_t* succ_adj, const int32_t* host_label, int32_t nt,
                   const int32_t* guest_parent, const int32_t* guest_label, uint64_t* n_switchings, void* stream) {
  if (!out || n <= 0 || m < 0 || nt <= 0 || !succ_off || (!succ_adj && m > 0) || !host_label || !guest_parent || !guest_label)
    return set_error(PT_ERR_INVALID, "pt_enum_create: bad argument");
  *out = nullptr;
  // ---- index structures (host side; O(n)) ----
  if (succ_off[0] != 0 || succ_off[n] != m) return set_error(PT_ERR_INVALID, "pt_enum_switchings: bad CSR offsets");
  std::vector<int32_t> indeg(n, 0), tdeg(nt, 0);
  for (int v = 0; v < n; ++v) { if (succ_off[v + 1] < succ_off[v]) return set_error(PT_ERR_INVALID, "pt_enum_switchings: bad CSR offsets"); }
  for (int e = 0; e < m; ++e) { if (succ_adj[e] < 0 || succ_adj[e] >= n) return set_error(PT_ERR_INVALID, "pt_enum_switchings: arc head out of range"); indeg[succ_adj[e]]++; }
  int troot = -1;
  for (int t = 0; t < nt; ++t) { const int p = guest_parent[t]; if (p >= nt || p == t) return set_error(PT_ERR_INVALID, "pt_enum_switchings: bad guest parent"); if (p < 0) { if (troot >= 0) return set_error(PT_ERR_INVALID, "guest has several roots"); troot = t; } else tdeg[p]++; }
  if (troot < 0) return set_error(PT_ERR_INVALID, "guest has no root");
  int maxlab = -1;
  for (int v = 0; v < n; ++v) maxlab = std::max(maxlab, host_label[v]);
  for (int t = 0; t < nt; ++t) maxlab = std::max(maxlab, guest_label[t]);
  std::vector<int32_t> l2h(maxlab + 2, -1), l2t(maxlab + 2, -1);
  for (int v = 0; v < n; ++v) if (host_label[v] >= 0 && succ_off[v + 1] == succ_off[v]) { if (l2h[host_label[v]] >= 0) return set_error(PT_ERR_INVALID, "single-label map for multi-labeled tree/network"); l2h[host_label[v]] = v; }
  for (int t = 0; t < nt; ++t) if (guest_label[t] >= 0 && tdeg[t] == 0) { if (l2t[guest_label[t]] >= 0) return set_error(PT_ERR_INVALID, "single-label map for multi-labeled network/tree"); l2t[guest_label[t]] = t; }
  int common = 0; bool t_only = false;
  for (int l = 0; l <= maxlab; ++l) { if (l2t[l] >= 0 && l2h[l] < 0) t_only = true; if (l2t[l] >= 0 && l2h[l] >= 0) ++common; }
  // topological order of the host (Kahn), root first
  std::vector<int32_t> order; order.reserve(n);
  { std::vector<int32_t> deg = indeg; for (int v = 0; v < n; ++v) if (!deg[v]) order.push_back(v);
    if (order.size() != 1) return set_error(PT_ERR_INVALID, "cannot create tree/network with %zu roots", order.size());
    for (size_t i = 0; i < order.size(); ++i) { const int v = order[i]; for (int e = succ_off[v]; e < succ_off[v + 1]; ++e) if (--deg[succ_adj[e]] == 0) order.push_back(succ_adj[e]); }
    if ((int)order.size() != n) return set_error(PT_ERR_INVALID, "host network has a cycle"); }
  std::vector<int32_t> posof(n);
  for (int i = 0; i < n; ++i) posof[order[n - 1 - i]] = i;  // children first
  std::vector<int32_t> retis; std::vector<int32_t> ridx(n, -1);
  for (int v = 0; v < n; ++v) if (indeg[v] >= 2) { ridx[v] = (int)retis.size(); retis.push_back(v); }
  const int r = (int)retis.size();
  uint64_t total = 1;
  for (int v : retis) { if (total > (UINT64_MAX >> 1) / (uint64_t)indeg[v]) return set_error(PT_ERR_UNSUPPORTED, "switching index space exceeds 2^63"); total *= (uint64_t)indeg[v]; }
  if (n_switchings) *n_switchings = total;
  pt_enum* h = new (std::nothrow) pt_enum();
  if (!h) return set_error(PT_ERR_NOMEM, "out of host memory");
  h->n = n; h->r = r; h->total = total;
  h->trivial = t_only ? 1 : (common <= 2 ? 2 : 0);       // containment.hpp:1116 never displayed / :1061,1208 always displayed
  h->succ_off.assign(succ_off, succ_off + n + 1); h->succ_adj.assign(succ_adj, succ_adj + m); h->host_label.assign(host_label, host_label + n);
  h->guest_parent.assign(guest_parent, guest_parent + nt); h->guest_label.assign(guest_label, guest_label + nt); h->retis = retis; h->indeg = indeg;
  if (h->trivial) { *out = h; return PT_OK; }
  if (int rc = require_device()) { delete h; return rc; }
  std::vector<EnumNode> nodes(n); std::vector<int32_t> radix(std::max(r, 1)), par_off(r + 1, 0), par_pos;
  for (int i = 0; i < r; ++i) { radix[i] = indeg[retis[i]]; par_off[i + 1] = par_off[i] + indeg[retis[i]]; }
  par_pos.assign(std::max(par_off[r], 1), 0);
  std::vector<int32_t> fill(std::max(r, 1), 0);
  // accumulators only for nodes with out-arcs (a leaf never receives a child value): half the per-thread state of a binary network
  std::vector<int32_t> accof(n, -1);
  int nacc = 0;
  for (int i = 0; i < n; ++i) { const int v = order[n - 1 - i]; if (succ_off[v + 1] > succ_off[v]) accof[v] = nacc++; }
  for (int v = 0; v < n; ++v) {
    EnumNode& nd = nodes[posof[v]];
    nd.parent = -1; nd.pad = 0;
    const int l = host_label[v];
    nd.label = (l >= 0 && succ_off[v + 1] == succ_off[v] && l2t[l] >= 0) ? l : -1;   // host-only labels are dropped (:1117-1121)
    nd.acc = accof[v];
    if (ridx[v] >= 0) nd.parent = -2 - ridx[v];
  }
  for (int u = 0; u < n; ++u) for (int e = succ_off[u]; e < succ_off[u + 1]; ++e) {
    const int v = succ_adj[e];
    if (ridx[v] >= 0) par_pos[par_off[ridx[v]] + fill[ridx[v]]++] = accof[u];   // digit order = arc order, like oracle/pt_oracle.c
    else nodes[posof[v]].parent = accof[u];
  }
  h->nacc = std::max(nacc, 1);
  // guest canonical hash (same functions as the device)
  std::vector<std::vector<int>> tk(nt);
  for (int t = 0; t < nt; ++t) if (guest_parent[t] >= 0) tk[guest_parent[t]].push_back(t);
  std::vector<u64> th(nt, 0);
  { std::vector<int> st{troot}, post; while (!st.empty()) { int t = st.back(); st.pop_back(); post.push_back(t); for (int c : tk[t]) st.push_back(c); }
    for (size_t i = post.size(); i-- > 0;) { const int t = post[i];
      if (tk[t].empty()) { th[t] = guest_label[t] >= 0 ? leaf_hash(guest_label[t]) : 0; continue; }
      int c = 0; u64 s1 = 0, s2 = 0; for (int k : tk[t]) if (th[k]) { ++c; s1 += th[k]; s2 += child_term(th[k]); }
      th[t] = c == 0 ? 0 : (c == 1 ? s1 : node_hash(s2)); } }
  h->guest_hash = th[troot];
  cudaStream_t s = (cudaStream_t)stream;
  DeviceProps dp;
  if (int rc = device_props(&dp)) { delete h; return rc; }
  h->smem = (size_t)(4 * n + r + (r + 1) + par_off[r] + 4) * 4;
  if (h->smem > (size_t)dp.smem_optin) { const size_t need = h->smem; delete h; return set_error(PT_ERR_UNSUPPORTED, "instance too large for the enumerator (%zu bytes of shared memory)", need); }
  // threads: the per-thread accumulators (20 bytes per node with out-arcs) are touched for every switching.  Measured on BASELINE
  // config 5 (2^24 switchings, 53 such nodes): 1 CTA of 256 per SM 38.2 ms, 2: 33.0, 3: 28.1, 4: 26.3, 8: 28.8 -- keeping all
  // accumulators in L2 (1-2 CTAs: <= 80 MB) does not pay, the kernel needs the warps to hide its dependent loads
  // (profiles/ncu_enum_r02.txt: long_scoreboard 24.5 stall cycles per issue); 4 CTAs per SM unless that takes more than 1 GB.
  int per_sm = 4;
  while (per_sm > 1 && (uint64_t)per_sm * (uint64_t)dp.sms * 256ull * (uint64_t)h->nacc * 20ull > (1ull << 30)) --per_sm;
  if (const char* e = getenv("PT_ENUM_CTAS_PER_SM")) per_sm = std::max(1, std::min(8, atoi(e)));   // tuning aid
  h->grid = std::max(dp.sms * per_sm, 1);
  const size_t T = (size_t)h->grid * 256;
  auto en_malloc = [](void** p, size_t bytes) -> cudaError_t { return dev_alloc(p, bytes) == PT_OK ? cudaSuccess : cudaErrorMemoryAllocation; };  // cached: raw cudaMalloc/cudaFree cost milliseconds
#define EN_TRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cudaStreamSynchronize(s); enum_release(h); return set_error(e__ == cudaErrorMemoryAllocation ? PT_ERR_NOMEM : PT_ERR_CUDA, "%s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); } } while (0)
  EN_TRY(cudaFuncSetAttribute(enum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem));
  EN_TRY(en_malloc((void**)&h->d_nodes, nodes.size() * sizeof(EnumNode))); EN_TRY(en_malloc((void**)&h->d_radix, radix.size() * 4));
  EN_TRY(en_malloc((void**)&h->d_off, par_off.size() * 4)); EN_TRY(en_malloc((void**)&h->d_par, par_pos.size() * 4));
  EN_TRY(en_malloc((void**)&h->d_S1, T * (size_t)h->nacc * 8)); EN_TRY(en_malloc((void**)&h->d_S2, T * (size_t)h->nacc * 8)); EN_TRY(en_malloc((void**)&h->d_cnt, T * (size_t)h->nacc * 4));
  EN_TRY(en_malloc((void**)&h->d_res, 64));
  EN_TRY(cudaMemcpyAsync(h->d_nodes, nodes.data(), nodes.size() * sizeof(EnumNode), cudaMemcpyHostToDevice, s));
  EN_TRY(cudaMemcpyAsync(h->d_radix, radix.data(), radix.size() * 4, cudaMemcpyHostToDevice, s));
  EN_TRY(cudaMemcpyAsync(h->d_off, par_off.data(), par_off.size() * 4, cudaMemcpyHostToDevice, s));
  EN_TRY(cudaMemcpyAsync(h->d_par, par_pos.data(), par_pos.size() * 4, cudaMemcpyHostToDevice, s));
  EN_TRY(cudaStreamSynchronize(s));   // the host vectors go out of scope
#undef EN_TRY
  *out = h;
  return PT_OK;
}

// [begin, end) of the switching index space (end == 0: to the end).  comm (optional, pt_comm of include/pt_gpu.h): the ranks of a
// multi-GPU run each pass their own sub-range; with stop_at_first the found word is reduced across the ranks (all-reduce MAX of
// one word) after every chunk of `chunk` switchings (0 = default 2^20), so ALL ranks stop as soon as ANY rank has a witness
// (SURVEY 8(e) row 2).  n_displaying: switchings of this rank's range whose display tree hashes like the guest (64-bit Merkle
// hash, collision probability ~2^-64 per comparison); first_index: the smallest such index of this rank's range, CONFIRMED
// EXACTLY (the switching is applied and the exact solver decides; a collision is skipped and counted out), or UINT64_MAX;
// *stopped_early = 1 when a witness anywhere ended the run before the range was exhausted.
int pt_enum_run(pt_enum* h, uint64_t begin, uint64_t end, int stop_at_first, pt_comm* comm, uint64_t chunk, uint64_t* n_displaying, uint64_t* first_index,
                int* stopped_early, void* stream) {
  if (!h) return set_error(PT_ERR_INVALID, "pt_enum_run: null handle");
  if (n_displaying) *n_displaying = 0;
  if (first_index) *first_index = UINT64_MAX;
  if (stopped_early) *stopped_early = 0;
  if (end == 0 || end > h->total) end = h->total;
  if (begin > end) begin = end;
  cudaStream_t s = (cudaStream_t)stream;
  const bool collective = comm && pt_comm_size(comm) > 1 && stop_at_first;
  if (h->trivial == 2 && begin < end) { if (n_displaying) *n_displaying = end - begin; if (first_index) *first_index = begin; }
  if (h->trivial && !collective) return PT_OK;
  if (chunk == 0) chunk = 1ull << 20;
  uint64_t count = 0, first = UINT64_MAX;
  u64* found_word = h->d_res ? h->d_res + 4 : nullptr;   // [4]: the word that is reduced across ranks
  u64* tmp_word = nullptr;
  if (h->trivial) { if (dev_alloc((void**)&tmp_word, 64) != PT_OK) return PT_ERR_NOMEM; found_word = tmp_word; }
  int rc = PT_OK;
  uint64_t pos = begin;
  // every rank runs the same number of poll rounds only as long as nobody has found anything: the loop ends for all ranks
  // together, because the reduced word is the same everywhere
  for (;;) {
    const uint64_t stop = std::min(end, pos + chunk);
    u64 res[4] = {0, ~0ull, 0, 0};
    if (!h->trivial && pos < stop) {
      const u64 init[4] = {0, ~0ull, 0, 0};
      if (cudaMemcpyAsync(h->d_res, init, 32, cudaMemcpyHostToDevice, s) != cudaSuccess) { rc = set_error(PT_ERR_CUDA, "pt_enum_run: copy failed"); break; }
      const uint64_t want = (stop - pos + 255) / 256;
      const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(want, (uint64_t)h->grid));
      enum_kernel<<<grid, 256, h->smem, s>>>(h->n, h->nacc, h->d_nodes, h->r, h->d_radix, h->d_off, h->d_par, pos, stop, h->guest_hash, stop_at_first, h->d_S1, h->d_S2, h->d_cnt, h->d_res);
      if (cudaGetLastError() != cudaSuccess || cudaMemcpyAsync(res, h->d_res, 32, cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) { rc = set_error(PT_ERR_CUDA, "pt_enum_run: launch failed: %s", cudaGetErrorString(cudaGetLastError())); break; }
      // the witness of this chunk, confirmed exactly; a hash collision is counted out and the search goes on behind it
      while (res[1] != ~0ull) {
        int ok = 0;
        if ((rc = enum_verify(h, res[1], &ok, stream)) != PT_OK) break;
        if (ok) break;
        const uint64_t bad = res[1];
        res[0] -= 1; res[1] = ~0ull;
        if (bad + 1 < stop) {
          const u64 init2[4] = {0, ~0ull, 0, 0};
          cudaMemcpyAsync(h->d_res, init2, 32, cudaMemcpyHostToDevice, s);
          enum_kernel<<<1, 256, h->smem, s>>>(h->n, h->nacc, h->d_nodes, h->r, h->d_radix, h->d_off, h->d_par, bad + 1, stop, h->guest_hash, 1, h->d_S1, h->d_S2, h->d_cnt, h->d_res);
          u64 r2[4];
          if (cudaMemcpyAsync(r2, h->d_res, 32, cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) { rc = set_error(PT_ERR_CUDA, "pt_enum_run: launch failed"); break; }
          res[1] = r2[1];
        }
      }
      if (rc != PT_OK) break;
      count += res[0];
      if (res[1] < first) first = res[1];
    } else if (h->trivial == 2 && pos < stop) first = std::min(first, pos);
    pos = stop;
    u64 found = first != UINT64_MAX ? 1 : 0;
    if (collective) {
      if (cudaMemcpyAsync(found_word, &found, 8, cudaMemcpyHostToDevice, s) != cudaSuccess) { rc = set_error(PT_ERR_CUDA, "pt_enum_run: copy failed"); break; }
      // second word: "this rank still has work": the loop may only end when nobody has work or somebody has a witness
      u64 more = pos < end ? 1 : 0;
      cudaMemcpyAsync(found_word + 1, &more, 8, cudaMemcpyHostToDevice, s);
      if ((rc = pt_comm_allreduce(comm, found_word, 2, 8, 1, stream)) != PT_OK) break;
      u64 back[2] = {0, 0};
      if (cudaMemcpyAsync(back, found_word, 16, cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) { rc = set_error(PT_ERR_CUDA, "pt_enum_run: copy failed"); break; }
      if (back[0]) { if (stopped_early && (pos < end || back[1])) *stopped_early = 1; break; }
      if (!back[1]) break;
    } else {
      if (stop_at_first && found) { if (stopped_early && pos < end) *stopped_early = 1; break; }
      if (pos >= end) break;
    }
  }
  if (tmp_word) { cudaStreamSynchronize(s); dev_free(tmp_word); }
  if (rc != PT_OK) return rc;
  if (n_displaying && h->trivial != 2) *n_displaying = count;
  if (first_index && h->trivial != 2) *first_index = first;
  return PT_OK;
}

int pt_enum_switchings(int32_t n, int32_t m, const int32_t* succ_off, const int32_t* succ_adj, const int32_t* host_label, int32_t nt,
                       const int32_t* guest_parent, const int32_t* guest_label, uint64_t begin, uint64_t end, int stop_at_first,
                       uint64_t* n_switchings, uint64_t* n_displaying, uint64_t* first_index, void* stream) {
  if (n_displaying) *n_displaying = 0;
  if (first_index) *first_index = UINT64_MAX;
  pt_enum* h = nullptr;
  if (int rc = pt_enum_create(&h, n, m, succ_off, succ_adj, host_label, nt, guest_parent, guest_label, n_switchings, stream)) return rc;
  // one chunk per call (the in-kernel flag ends a stop_at_first run early inside the launch)
  const int rc = pt_enum_run(h, begin, end, stop_at_first, nullptr, ~0ull >> 1, n_displaying, first_index, nullptr, stream);
  pt_enum_free(h);
  return rc;
}

}  // extern "C"
