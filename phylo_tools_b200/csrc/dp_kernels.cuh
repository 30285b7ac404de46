// csrc/dp_kernels.cuh -- the tree-in-tree DP in its general form, on top of the LCA kernels (included by lca_kernels.cu).
//
//   pt_who_*            TreeInTreeContainment<Host, Guest>::who_displays for MULTI-LABELLED tree hosts
//                       utils/containment.hpp:27-233, utils/induced_tree.hpp:65-168, utils/matching.hpp:8-93
//   pt_net_unzip        TreeInComponent::unzip_retis                          utils/containment.hpp:262-312
//   pt_who_highest_displayed   TreeInComponent::highest_displayed_ancestor    utils/containment.hpp:328-344
//   pt_net_tree_components     TreeComponentInfos (comp_root, visible_leaf, component DAG)  utils/tree_components.hpp:41-234
//
// The reference computes, per guest node u, the preorder-sorted list M(u) of host nodes that minimally display the guest
// subtree below u: it merges the children's lists by preorder, builds the host subtree they induce (LCAs of consecutive
// candidates + nearest-smaller-depth stacks), lets every candidate tell its induced ancestors which child of u it can display
// (hash map of bitsets per induced node) and, in post-order, accepts every induced node at which u's children can be matched
// to DISTINCT induced children (bipartite matching), discarding the ancestors of an accepted node.
//
// Here: DATAFLOW over the guest tree (dp_who_flow_kernel: a thread starts at every guest leaf, whoever completes the last child of a
// node goes on with that node; no levels, no grid barrier, any guest depth).  Lists are kept as host PREORDER POSITIONS in a device
// arena.  One thread per guest node: the children's lists are merged (heap sort of
// (position, child) keys), the induced subtree is built by the classic stack sweep over the sorted candidates -- one LCA query
// per consecutive pair, through the same record / table layout as pt_lca_query -- and every induced node is finished the moment
// it is popped from the stack (its subtree is complete then): the CHILD-SET UNION it hands to its parent is one 64-bit mask
// (bit i = "a candidate of child i lies below"), the matching runs over its popped children's masks (Kuhn's augmenting paths,
// at most 64 guest children), accepted nodes are appended in pop order, which for an antichain IS ascending preorder.  Guest
// nodes whose children all carry one candidate take the closed form of the single-labelled case (k-1 LCAs, no scratch).
#pragma once

namespace ptgpu {

constexpr int DP_CHILD_BITS = 20;             // low bits of a merge key: which child of u the candidate belongs to
constexpr int DP_MAX_CHILDREN = 1024;         // general case (some child has several candidates): child sets are W = ceil(d / 64) words per induced node and the
                                              // matching is serial in one thread; guest nodes whose children all have ONE candidate (every node over a
                                              // single-labelled host) take the closed form below for ANY number of children

struct DpArena {
  int32_t* mpos;       // list arena: host preorder positions
  int64_t mcap;
  int32_t* scratch;    // scratch of the nodes that need the general case (or a sort of many children), bump-allocated
  int64_t scap;
  unsigned long long* tops;  // [0] list arena top  [1] scratch top  [2] nodes processed
  int32_t* flags;      // [0] too many children  [1] list arena overflow  [2] scratch overflow
};

// (dp_heapsort: lca_kernels.cu, next to who_displays_kernel)

// maximum matching saturating all d left nodes (the guest children)?  Right node y (an induced child) is adjacent to left i iff bit i of the
// W-word child set cov[ridx[y] * W ...].  Kuhn's augmenting paths, iterative; L / it / chosen: d ints each
__device__ inline bool dp_perfect_matching(int d, int r, int W, const u64* cov, const int32_t* ridx, int32_t* matchR, int32_t* seen, int32_t* L, int32_t* it, int32_t* chosen) {
  if (r < d) return false;
  for (int w = 0; w < W; ++w) {
    u64 all = 0;
    for (int y = 0; y < r; ++y) all |= cov[(int64_t)ridx[y] * W + w];
    const int bits = d - 64 * w;
    const u64 full = bits >= 64 ? ~0ull : ((1ull << bits) - 1ull);
    if ((all & full) != full) return false;
  }
  for (int y = 0; y < r; ++y) matchR[y] = -1;
  for (int i = 0; i < d; ++i) {
    for (int y = 0; y < r; ++y) seen[y] = 0;
    int depth = 0; L[0] = i; it[0] = 0;
    bool ok = false;
    while (depth >= 0) {
      const int l = L[depth];
      int found = -1;
      while (it[depth] < r) { const int y = it[depth]++; if (((cov[(int64_t)ridx[y] * W + (l >> 6)] >> (l & 63)) & 1ull) && !seen[y]) { seen[y] = 1; found = y; break; } }
      if (found < 0) { --depth; continue; }
      chosen[depth] = found;
      if (matchR[found] < 0) { for (int k = 0; k <= depth; ++k) matchR[chosen[k]] = L[k]; ok = true; break; }
      if (depth + 1 >= d) break;   // (cannot happen: the left nodes of an alternating path are distinct)
      L[depth + 1] = matchR[found]; it[depth + 1] = 0; ++depth;
    }
    if (!ok) return false;
  }
  return true;
}

// mlen[u]: -2 dead (cleaned away: no labelled leaf below), -1 not computed yet, >= 0 length of M(u); moff[u] its start in the arena.
// One guest node: everything its children have left (mlen / moff / the list arena) is read through L2 -- other SMs wrote it during
// the same launch.
__device__ __forceinline__ void dp_node(const int32_t u, const int32_t* __restrict__ child_off, const int32_t* __restrict__ child_adj, const int32_t* __restrict__ tlabel,
                                        const int32_t* __restrict__ lab_off, const int32_t* __restrict__ lab_pos, int64_t nlab, const LcaRecord* __restrict__ rec,
                                        const u64* __restrict__ keypos, const int32_t* __restrict__ nodeat, const u64* __restrict__ table, const int64_t* __restrict__ level_base,
                                        int64_t* moff, int32_t* mlen, const DpArena& A) {
  const int32_t cb = child_off[u], ce = child_off[u + 1];
  if (cb == ce) {  // guest leaf: the host leaves with its label (construct_base_cases :115-126)
    const int32_t lab = tlabel[u];
    if (lab < 0) { mlen[u] = -2; return; }
    const int32_t b = (int64_t)lab < nlab ? lab_off[lab] : 0, e = (int64_t)lab < nlab ? lab_off[lab + 1] : 0;
    const int len = e - b;
    const unsigned long long at = atomicAdd(&A.tops[0], (unsigned long long)len);
    if ((int64_t)(at + len) > A.mcap) { A.flags[1] = 1; mlen[u] = 0; moff[u] = 0; return; }
    for (int k = 0; k < len; ++k) A.mpos[at + k] = lab_pos[b + k];
    moff[u] = (int64_t)at; mlen[u] = len;
    return;
  }
  // live children, K = total number of candidates
  int d = 0; int64_t K = 0; bool missing = false, single = true; int32_t only = -1;
  for (int32_t j = cb; j < ce; ++j) {
    const int32_t c = child_adj[j], len = __ldcg(&mlen[c]);
    if (len == -2) continue;
    if (len == 0) missing = true;   // merge_child_poss :180-230: an empty child list empties the parent
    ++d; K += len; single &= len == 1; only = c;
  }
  if (d == 0) { mlen[u] = -2; return; }
  if (missing) { mlen[u] = 0; moff[u] = 0; return; }
  if (d == 1) { mlen[u] = __ldcg(&mlen[only]); moff[u] = __ldcg(&moff[only]); return; }   // :167 -- the same list (shared, never modified)
  if (single) {
    // every child has exactly one candidate: u is displayed iff the candidates sit in d different branches of their
    // common ancestor v, i.e. all consecutive LCAs (in preorder) are v and no candidate is v itself
    uint32_t few[8]; uint32_t* ps = few;
    if (d > 8) {   // many children (a star, a wide multifurcation): sorted in a scratch span instead of registers
      const int64_t need = d + 2;
      const unsigned long long sat = atomicAdd(&A.tops[1], (unsigned long long)need);
      if ((int64_t)(sat + need) > A.scap) { A.flags[2] = 1; mlen[u] = 0; moff[u] = 0; return; }
      ps = reinterpret_cast<uint32_t*>(A.scratch + sat);
    }
    int k = 0;
    for (int32_t j = cb; j < ce; ++j) {
      const int32_t c = child_adj[j];
      if (__ldcg(&mlen[c]) != 1) continue;
      const uint32_t p = (uint32_t)__ldcg(&A.mpos[__ldcg(&moff[c])]);
      if (d > 8) { ps[k++] = p; continue; }
      int a = k++;
      while (a > 0 && ps[a - 1] > p) { ps[a] = ps[a - 1]; --a; }
      ps[a] = p;
    }
    if (d > 8) dp_heapsort(ps, (int64_t)d);
    const int32_t v = lca_one(rec, keypos, table, level_base, nodeat[ps[0]], nodeat[ps[d - 1]]);
    const uint32_t vp = rec[v].pos;
    bool ok = true;
    for (int a = 0; a < d && ok; ++a) ok = ps[a] != vp;
    for (int a = 0; a + 1 < d && ok; ++a) ok = ps[a] != ps[a + 1] && lca_one(rec, keypos, table, level_base, nodeat[ps[a]], nodeat[ps[a + 1]]) == v;
    if (!ok) { mlen[u] = 0; moff[u] = 0; return; }
    const unsigned long long at = atomicAdd(&A.tops[0], 1ull);
    if ((int64_t)(at + 1) > A.mcap) { A.flags[1] = 1; mlen[u] = 0; moff[u] = 0; return; }
    A.mpos[at] = (int32_t)vp; moff[u] = (int64_t)at; mlen[u] = 1;
    return;
  }
  if (d > DP_MAX_CHILDREN) { A.flags[0] = 1; mlen[u] = 0; moff[u] = 0; return; }
  // ---- general case: W = ceil(d / 64) words per child set.  scratch = keys (2K words) | cov (2VW words) | per induced node (V <= 2K): pos, depth,
  //      first, next, nchild, sb | stack V | ridx V | matchR V | seen V | result K | L, it, chosen (d each)
  const int W = (d + 63) >> 6;
  const int64_t V = 2 * K;
  const int64_t need = 2 * K + 2 * V * W + 6 * V + 4 * V + K + 3 * (int64_t)d + 16;
  const unsigned long long sat = atomicAdd(&A.tops[1], (unsigned long long)need);
  if ((int64_t)(sat + need) > A.scap) { A.flags[2] = 1; mlen[u] = 0; moff[u] = 0; return; }
  int32_t* sp = A.scratch + sat + (sat & 1);   // 8-byte alignment for the 64-bit arrays
  u64* keys = reinterpret_cast<u64*>(sp); sp += 2 * K;
  u64* cov = reinterpret_cast<u64*>(sp); sp += 2 * V * W;
  int32_t* vpos = sp; sp += V; int32_t* vdepth = sp; sp += V; int32_t* vfirst = sp; sp += V; int32_t* vnext = sp; sp += V;
  int32_t* vnch = sp; sp += V; int32_t* vsb = sp; sp += V; int32_t* stack = sp; sp += V; int32_t* ridx = sp; sp += V; int32_t* matchR = sp; sp += V; int32_t* seen = sp; sp += V;
  int32_t* result = sp; sp += K;
  int32_t* mL = sp; sp += d; int32_t* mit = sp; sp += d; int32_t* mchosen = sp;
  {
    int64_t k = 0; int ci = 0;
    for (int32_t j = cb; j < ce; ++j) {
      const int32_t c = child_adj[j], len = __ldcg(&mlen[c]);
      if (len == -2) continue;
      const int64_t o = __ldcg(&moff[c]);
      for (int t = 0; t < len; ++t) keys[k++] = ((u64)(uint32_t)__ldcg(&A.mpos[o + t]) << DP_CHILD_BITS) | (u64)ci;
      ++ci;
    }
  }
  dp_heapsort(keys, K);
  int nv = 0, top = -1, nres = 0, cur = -1;
  // a node is FINISHED when it leaves the stack: its subtree is complete, so it can be matched and handed to its parent
  auto finish = [&](int x, int q) {
    int succ = 0;
    if (!vsb[x] && vnch[x] >= d) {
      int r = 0;
      for (int y = vfirst[x]; y >= 0; y = vnext[y]) ridx[r++] = y;
      if (dp_perfect_matching(d, r, W, cov, ridx, matchR, seen, mL, mit, mchosen)) { result[nres++] = vpos[x]; succ = 1; }   // perfect_child_matching :174-177
    }
    if (q >= 0) {   // :140-149, :160
      for (int w = 0; w < W; ++w) cov[(int64_t)q * W + w] |= cov[(int64_t)x * W + w];
      vsb[q] |= succ | vsb[x]; vnext[x] = vfirst[q]; vfirst[q] = x; vnch[q]++;
    }
  };
  auto newnode = [&](uint32_t p) {
    const int x = nv++; vpos[x] = (int32_t)p; vdepth[x] = (int32_t)(keypos[p] >> 32); vfirst[x] = -1; vnext[x] = -1; vnch[x] = 0; vsb[x] = 0;
    for (int w = 0; w < W; ++w) cov[(int64_t)x * W + w] = 0ull;
    return x;
  };
  for (int64_t j = 0; j < K; ++j) {
    const uint32_t p = (uint32_t)(keys[j] >> DP_CHILD_BITS); const int ci = (int)(keys[j] & ((1u << DP_CHILD_BITS) - 1u));
    if (cur >= 0 && (uint32_t)vpos[cur] == p) { cov[(int64_t)cur * W + (ci >> 6)] |= 1ull << (ci & 63); continue; }   // the same host node is a candidate of several children
    const int w = newnode(p);
    cov[(int64_t)w * W + (ci >> 6)] |= 1ull << (ci & 63);
    cur = w;
    if (top < 0) { stack[++top] = w; continue; }
    const int t = stack[top];
    const int32_t lnode = lca_one(rec, keypos, table, level_base, nodeat[vpos[t]], nodeat[p]);   // induced_tree.hpp:128-138
    const uint32_t lp = rec[lnode].pos; const int32_t ld = (int32_t)(keypos[lp] >> 32);
    if (lp != (uint32_t)vpos[t]) {
      while (top >= 1 && vdepth[stack[top - 1]] >= ld) { finish(stack[top], stack[top - 1]); --top; }
      if ((uint32_t)vpos[stack[top]] != lp) { const int Lx = newnode(lp); finish(stack[top], Lx); stack[top] = Lx; }
    }
    stack[++top] = w;
  }
  while (top >= 1) { finish(stack[top], stack[top - 1]); --top; }
  if (top == 0) finish(stack[0], -1);
  if (nres == 0) { mlen[u] = 0; moff[u] = 0; return; }
  const unsigned long long at = atomicAdd(&A.tops[0], (unsigned long long)nres);
  if ((int64_t)(at + nres) > A.mcap) { A.flags[1] = 1; mlen[u] = 0; moff[u] = 0; return; }
  for (int k = 0; k < nres; ++k) A.mpos[at + k] = result[k];
  moff[u] = (int64_t)at; mlen[u] = nres;
}

// DATAFLOW over the guest tree, no levels and no grid barrier: one thread starts at every guest leaf; whoever completes the LAST
// child of a node (a counter per node, decremented behind a fence) goes on with that node at once, everybody else is done.  The
// level-synchronous version this replaces spent 204 stall cycles per issued instruction at its grid barrier (62 levels for a
// random guest of 10^6 leaves, most of them nearly empty) and paid one barrier per level of a deep guest.
__global__ void __launch_bounds__(256) dp_who_flow_kernel(int64_t nt, const int32_t* __restrict__ gparent, int32_t* __restrict__ pending, const int32_t* __restrict__ child_off,
                                                          const int32_t* __restrict__ child_adj, const int32_t* __restrict__ tlabel, const int32_t* __restrict__ lab_off,
                                                          const int32_t* __restrict__ lab_pos, int64_t nlab, const LcaRecord* __restrict__ rec, const u64* __restrict__ keypos,
                                                          const int32_t* __restrict__ nodeat, const u64* __restrict__ table, const int64_t* __restrict__ level_base,
                                                          int64_t* moff, int32_t* mlen, DpArena A) {
  unsigned long long done = 0;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < nt; idx += (int64_t)gridDim.x * blockDim.x) {
    int32_t u = (int32_t)idx;
    if (child_off[u] != child_off[u + 1]) continue;   // start at the leaves
    for (;;) {
      dp_node(u, child_off, child_adj, tlabel, lab_off, lab_pos, nlab, rec, keypos, nodeat, table, level_base, moff, mlen, A);
      ++done;
      const int32_t p = gparent[u];
      if (p < 0 || p >= nt) break;
      __threadfence();                                 // my list is out before my parent can count me
      if (atomicSub(&pending[p], 1) != 1) break;       // somebody else will complete p
      __threadfence();
      u = p;
    }
  }
  if (done) atomicAdd(&A.tops[2], done);   // nodes processed: all of them unless the parent array has a cycle
}

// host leaves grouped by label: count, then fill with POSITIONS; every group is sorted afterwards (small groups: insertion sort)
__global__ void dp_label_count_kernel(int64_t n, const int32_t* __restrict__ deg, const int32_t* __restrict__ hlabel, int64_t nlab, int32_t* __restrict__ cnt, int32_t* __restrict__ flags) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
    const int32_t lab = hlabel[v];
    if (lab >= 0 && deg[v] == 0) { if (lab >= nlab) flags[3] = 1; else atomicAdd(&cnt[lab], 1); }
  }
}
__global__ void dp_label_fill_kernel(int64_t n, const int32_t* __restrict__ deg, const int32_t* __restrict__ hlabel, int64_t nlab, const int32_t* __restrict__ lab_off, int32_t* __restrict__ cursor,
                                     const LcaRecord* __restrict__ rec, int32_t* __restrict__ lab_pos) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
    const int32_t lab = hlabel[v];
    if (lab >= 0 && lab < nlab && deg[v] == 0) lab_pos[lab_off[lab] + atomicAdd(&cursor[lab], 1)] = (int32_t)rec[v].pos;
  }
}
__global__ void dp_label_sort_kernel(int64_t nlab, const int32_t* __restrict__ lab_off, int32_t* __restrict__ lab_pos) {
  for (int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; l < nlab; l += (int64_t)gridDim.x * blockDim.x) {
    const int32_t b = lab_off[l], e = lab_off[l + 1];
    for (int32_t i = b + 1; i < e; ++i) { const int32_t x = lab_pos[i]; int32_t j = i; while (j > b && lab_pos[j - 1] > x) { lab_pos[j] = lab_pos[j - 1]; --j; } lab_pos[j] = x; }
  }
}
__global__ void dp_lengths_kernel(int64_t nt, const int32_t* __restrict__ mlen, int32_t* __restrict__ len) {
  for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < nt; u += (int64_t)gridDim.x * blockDim.x) len[u] = mlen[u] > 0 ? mlen[u] : 0;
}
__global__ void dp_gather_kernel(int64_t nt, const int32_t* __restrict__ mlen, const int64_t* __restrict__ moff, const int32_t* __restrict__ mpos, const int32_t* __restrict__ nodeat,
                                 const int32_t* __restrict__ out_off, int32_t* __restrict__ out_adj) {
  for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < nt; u += (int64_t)gridDim.x * blockDim.x) {
    const int len = mlen[u] > 0 ? mlen[u] : 0; const int64_t o = moff[u]; const int32_t b = out_off[u];
    for (int k = 0; k < len; ++k) out_adj[b + k] = nodeat[mpos[o + k]];
  }
}
// the one-entry view that highest_displayed_ancestor needs: -1 empty; host_root iff the list is exactly { host_root }; else a node != host_root
__global__ void dp_who1_kernel(int64_t nt, const int32_t* __restrict__ off, const int32_t* __restrict__ adj, int32_t host_root, int32_t* __restrict__ who1) {
  for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < nt; u += (int64_t)gridDim.x * blockDim.x) {
    const int32_t b = off[u], len = off[u + 1] - b;
    who1[u] = len == 0 ? -1 : (len == 1 ? adj[b] : (adj[b] != host_root ? adj[b] : adj[b + 1]));
  }
}

// ---- unzip_retis ------------------------------------------------------------------------------------------------------------
// usize[x] = number of nodes of the unfolding below (and including) x, for nodes of out-degree != 1; -1 = not known yet.
// One round resolves every node all of whose (chain-skipped) children are resolved: rounds = height of the DAG.
__device__ __forceinline__ int32_t dp_skip(const int32_t* __restrict__ off, const int32_t* __restrict__ adj, int32_t y, int64_t n) {
  int64_t guard = 0;
  while (off[y + 1] - off[y] == 1 && guard++ <= n) y = adj[off[y]];
  return y;
}
__global__ void __launch_bounds__(256) dp_unzip_sizes_kernel(int64_t n, const int32_t* __restrict__ off, const int32_t* __restrict__ adj, long long* __restrict__ usize, int32_t* __restrict__ flags, long long cap) {
  cg::grid_group grid = cg::this_grid();
  const int64_t gtid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, gsize = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = gtid; v < n; v += gsize) usize[v] = (off[v + 1] == off[v]) ? 1 : -1;
  grid.sync();
  for (int64_t round = 0; round <= n; ++round) {
    if (gtid == 0) flags[0] = 0;
    grid.sync();
    for (int64_t v = gtid; v < n; v += gsize) {
      if (usize[v] >= 0 || off[v + 1] - off[v] == 1) continue;
      long long s = 1; bool ready = true;
      for (int32_t k = off[v]; k < off[v + 1] && ready; ++k) { const long long c = *(volatile long long*)&usize[dp_skip(off, adj, adj[k], n)]; if (c < 0) ready = false; else { s += c; if (s > cap) s = cap + 1; } }
      if (ready) { usize[v] = s; flags[0] = 1; }
    }
    grid.sync();
    if (!*(volatile int32_t*)&flags[0]) break;
    grid.sync();
  }
}
// level-synchronous expansion: a frontier entry is (copy id, host node); the ids of its children follow from the sizes
__global__ void __launch_bounds__(256) dp_unzip_expand_kernel(int64_t n, const int32_t* __restrict__ off, const int32_t* __restrict__ adj, const int32_t* __restrict__ label,
                                                              const long long* __restrict__ usize, int32_t root, int32_t* __restrict__ fr_a, int32_t* __restrict__ fr_b,
                                                              int32_t* __restrict__ ctr, int32_t* __restrict__ out_parent, int32_t* __restrict__ out_label, int32_t* __restrict__ out_origin) {
  cg::grid_group grid = cg::this_grid();
  const int64_t gtid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, gsize = (int64_t)gridDim.x * blockDim.x;
  if (gtid == 0) {
    out_parent[0] = -1; out_origin[0] = root; out_label[0] = (off[root + 1] == off[root]) ? label[root] : -1;
    fr_a[0] = 0; fr_a[1] = root; ctr[0] = (off[root + 1] - off[root] >= 2) ? 1 : 0; ctr[1] = 0;
  }
  grid.sync();
  int32_t* cur = fr_a; int32_t* nxt = fr_b; int which = 0;
  for (;;) {
    const int32_t cnt = *(volatile int32_t*)&ctr[which];
    if (cnt == 0) break;
    for (int64_t i = gtid; i < cnt; i += gsize) {
      const int32_t id = cur[2 * i], x = cur[2 * i + 1];
      long long cid = (long long)id + 1;
      for (int32_t k = off[x]; k < off[x + 1]; ++k) {
        const int32_t y = dp_skip(off, adj, adj[k], n);
        out_parent[cid] = id; out_origin[cid] = y; out_label[cid] = (off[y + 1] == off[y]) ? label[y] : -1;
        if (off[y + 1] - off[y] >= 2) { const int32_t slot = atomicAdd(&ctr[which ^ 1], 1); nxt[2 * slot] = (int32_t)cid; nxt[2 * slot + 1] = y; }
        cid += usize[y];
      }
    }
    grid.sync();
    if (gtid == 0) ctr[which] = 0;
    which ^= 1; int32_t* t = cur; cur = nxt; nxt = t;
    grid.sync();
  }
}

// ---- TreeComponentInfos ---------------------------------------------------------------------------------------------------------
// comp_root: -1 = NoNode, -2 = not known yet (internal).  A node is a reticulation iff its in-degree is > 1.
__global__ void tc_indeg_kernel(int64_t n, const int32_t* __restrict__ off, const int32_t* __restrict__ adj, int32_t* __restrict__ indeg, int32_t* __restrict__ one_parent) {
  for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < n; u += (int64_t)gridDim.x * blockDim.x)
    for (int32_t k = off[u]; k < off[u + 1]; ++k) { atomicAdd(&indeg[adj[k]], 1); one_parent[adj[k]] = (int32_t)u; }
}
__global__ void tc_pred_fill_kernel(int64_t n, const int32_t* __restrict__ off, const int32_t* __restrict__ adj, const int32_t* __restrict__ poff, int32_t* __restrict__ cursor, int32_t* __restrict__ padj) {
  for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < n; u += (int64_t)gridDim.x * blockDim.x)
    for (int32_t k = off[u]; k < off[u + 1]; ++k) { const int32_t h = adj[k]; padj[poff[h] + atomicAdd(&cursor[h], 1)] = (int32_t)u; }
}
// what the caller sees: a reticulation with several children keeps NoNode (update_reticulations only stores into nodes of
// out-degree <= 1, tree_components.hpp:112), although its value is what the nodes below it inherit
__global__ void tc_output_kernel(int64_t n, const int32_t* __restrict__ off, const int32_t* __restrict__ indeg, const int32_t* __restrict__ comp, int32_t* __restrict__ out) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x)
    out[v] = (indeg[v] > 1 && off[v + 1] - off[v] > 1) ? -1 : (comp[v] < 0 ? -1 : comp[v]);
}
// tree nodes (in-degree <= 1, out-degree >= 1): up[u] = u for the root and for nodes whose parent is a reticulation (component roots),
// else the parent; pointer jumping then gives the component root (tree_components.hpp:69-91)
__global__ void tc_init_kernel(int64_t n, const int32_t* __restrict__ off, const int32_t* __restrict__ indeg, const int32_t* __restrict__ one_parent, int32_t* __restrict__ up) {
  for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < n; u += (int64_t)gridDim.x * blockDim.x) {
    const bool tree_inner = indeg[u] <= 1 && off[u + 1] > off[u];
    if (!tree_inner) { up[u] = -2; continue; }
    up[u] = (indeg[u] == 0 || indeg[one_parent[u]] > 1) ? (int32_t)u : one_parent[u];
  }
}
__global__ void tc_jump_kernel(int64_t n, int32_t* __restrict__ up, int32_t* __restrict__ changed) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
    const int32_t a = up[v];
    if (a < 0) continue;
    const int32_t b = up[a];
    if (b >= 0 && a != b) { up[v] = b; *changed = 1; }
  }
}
// reticulations and leaves: the common component root of all parents (a reticulation parent contributes its own value), or NoNode
// (update_reticulations :96-114).  One round resolves every node whose reticulation parents are resolved.
__global__ void tc_consensus_kernel(int64_t n, const int32_t* __restrict__ off, const int32_t* __restrict__ indeg, const int32_t* __restrict__ poff, const int32_t* __restrict__ padj,
                                    int32_t* __restrict__ comp, int32_t* __restrict__ changed) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    if (comp[r] != -2) continue;
    int32_t rt = -3; bool ready = true;   // -3 = uninitialised ("r" in the reference), -1 = more than one component
    for (int32_t k = poff[r]; k < poff[r + 1]; ++k) {
      const int32_t c = *(volatile int32_t*)&comp[padj[k]];
      if (c == -2) { ready = false; break; }
      if (rt == -3) rt = c;
      if (rt != c) rt = -1;
    }
    if (!ready) continue;
    comp[r] = rt == -3 ? -1 : rt;
    *changed = 1;
  }
}
__global__ void tc_visible_kernel(int64_t n, const int32_t* __restrict__ off, const int32_t* __restrict__ comp, int32_t* __restrict__ visible) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x)
    if (off[v + 1] == off[v] && comp[v] >= 0) atomicMax(&visible[comp[v]], (int32_t)v);
}
// component DAG: for every non-trivial component root u (tree node below a reticulation) the component roots of all tree nodes
// reached by climbing through the reticulations above u (compute_edges :117-131).  One thread per root, explicit stack in a
// bounded local array; pass 0 counts, pass 1 writes (duplicates are removed on the host side of the call).
__global__ void tc_dag_kernel(int64_t n, const int32_t* __restrict__ off, const int32_t* __restrict__ indeg, const int32_t* __restrict__ one_parent, const int32_t* __restrict__ poff,
                              const int32_t* __restrict__ padj, const int32_t* __restrict__ comp, int32_t* __restrict__ count, long long* __restrict__ edges, int64_t cap, int32_t* __restrict__ flags) {
  for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < n; u += (int64_t)gridDim.x * blockDim.x) {
    const bool tree_inner = indeg[u] <= 1 && off[u + 1] > off[u];
    if (!tree_inner || indeg[u] == 0 || indeg[one_parent[u]] <= 1) continue;
    int32_t st[64]; int sp = 0;
    st[sp++] = one_parent[u];
    while (sp) {
      const int32_t r = st[--sp];
      for (int32_t k = poff[r]; k < poff[r + 1]; ++k) {
        const int32_t p = padj[k];
        if (indeg[p] > 1) { if (sp < 64) st[sp++] = p; else flags[0] = 1; }
        else if (comp[p] >= 0) { const int32_t at = atomicAdd(count, 1); if (at < cap) edges[at] = ((long long)comp[p] << 32) | (long long)(uint32_t)u; }
      }
    }
  }
}

}  // namespace ptgpu

// =====================================================================================================================
struct pt_who {
  int64_t n = 0, nt = 0, total = 0;
  int32_t* d_off = nullptr;   // nt + 1
  int32_t* d_adj = nullptr;   // total host node ids, every list ascending in host preorder
  int32_t* d_gparent = nullptr;
};

extern "C" {

int pt_who_free(pt_who* h) {
  if (!h) return PT_OK;
  cudaDeviceSynchronize();
  dev_free(h->d_off); dev_free(h->d_adj); dev_free(h->d_gparent);
  delete h;
  return PT_OK;
}

int pt_who_build(pt_who** out, const int32_t* host_parent, const int32_t* host_label, int64_t n, const int32_t* guest_parent, const int32_t* guest_label, int64_t nt,
                 pt_mem mem, void* stream) {
  if (!out || !host_parent || !host_label || !guest_parent || !guest_label || n <= 0 || nt <= 0 || n > INT32_MAX / 4 || nt > INT32_MAX / 4)
    return set_error(PT_ERR_INVALID, "pt_who_build: bad argument");
  *out = nullptr;
  if (int rc = require_device()) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  DeviceProps dp;
  if (int rc = device_props(&dp)) return rc;
  std::vector<void*> owned;
  pt_lca* H = nullptr;
  pt_who* W = new (std::nothrow) pt_who();
  if (!W) return set_error(PT_ERR_NOMEM, "out of host memory");
  W->n = n; W->nt = nt;
  auto cleanup = [&]() { cudaStreamSynchronize(s); for (void* p : owned) dev_free(p); if (H) pt_lca_free(H); };
  auto fail = [&](int code) { cleanup(); pt_who_free(W); return code; };
  auto dalloc = [&](size_t bytes) -> void* { void* p = nullptr; if (lca_malloc(&p, bytes ? bytes : 4) != cudaSuccess) { cudaGetLastError(); return nullptr; } owned.push_back(p); return p; };
#define DP_TRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return fail(set_error(PT_ERR_CUDA, "%s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e__))); } while (0)
#define DP_PTR(p) do { if (!(p)) return fail(set_error(PT_ERR_NOMEM, "pt_who_build: out of device memory")); } while (0)
  const int32_t *hp = host_parent, *hl = host_label, *gl = guest_label;
  DP_TRY(lca_malloc((void**)&W->d_gparent, (size_t)nt * 4));
  if (mem == PT_MEM_HOST) {
    int32_t* a = (int32_t*)dalloc((size_t)n * 4); int32_t* b = (int32_t*)dalloc((size_t)n * 4); int32_t* d = (int32_t*)dalloc((size_t)nt * 4);
    DP_PTR(a); DP_PTR(b); DP_PTR(d);
    DP_TRY(cudaMemcpyAsync(a, host_parent, (size_t)n * 4, cudaMemcpyHostToDevice, s)); DP_TRY(cudaMemcpyAsync(b, host_label, (size_t)n * 4, cudaMemcpyHostToDevice, s));
    DP_TRY(cudaMemcpyAsync(W->d_gparent, guest_parent, (size_t)nt * 4, cudaMemcpyHostToDevice, s)); DP_TRY(cudaMemcpyAsync(d, guest_label, (size_t)nt * 4, cudaMemcpyHostToDevice, s));
    hp = a; hl = b; gl = d;
  } else DP_TRY(cudaMemcpyAsync(W->d_gparent, guest_parent, (size_t)nt * 4, cudaMemcpyDeviceToDevice, s));
  const int32_t* gp = W->d_gparent;
  if (int rc = lca_build_impl(&H, hp, n, PT_MEM_DEVICE, stream, true)) return fail(rc);
  const int grid = dp.sms * 8;
  const int64_t nlab = n + nt;
  int32_t* hdeg = (int32_t*)dalloc((size_t)n * 4 + 4); int32_t* small = (int32_t*)dalloc(256); int32_t* lcnt = (int32_t*)dalloc((size_t)nlab * 4 + 4); int32_t* loff = (int32_t*)dalloc((size_t)nlab * 4 + 4);
  int32_t* tdeg = (int32_t*)dalloc((size_t)nt * 4 + 4); int32_t* toff = (int32_t*)dalloc((size_t)nt * 4 + 4); int32_t* tadj = (int32_t*)dalloc((size_t)nt * 4);
  int32_t* mlen = (int32_t*)dalloc((size_t)nt * 4);
  int64_t* moff = (int64_t*)dalloc((size_t)nt * 8); int32_t* lpos = (int32_t*)dalloc((size_t)n * 4);
  int64_t* scan_tmp = (int64_t*)dalloc(((size_t)((std::max(nlab, nt) + SCAN_TILE - 1) / SCAN_TILE) + 1) * 8);
  DP_PTR(hdeg); DP_PTR(small); DP_PTR(lcnt); DP_PTR(loff); DP_PTR(tdeg); DP_PTR(toff); DP_PTR(tadj);
  DP_PTR(mlen); DP_PTR(moff); DP_PTR(lpos); DP_PTR(scan_tmp);
  DP_TRY(cudaMemsetAsync(hdeg, 0, (size_t)n * 4 + 4, s)); DP_TRY(cudaMemsetAsync(tdeg, 0, (size_t)nt * 4 + 4, s)); DP_TRY(cudaMemsetAsync(small, 0, 256, s));
  DP_TRY(cudaMemsetAsync(lcnt, 0, (size_t)nlab * 4 + 4, s));
  lca_count_kernel<<<grid, 256, 0, s>>>(hp, n, hdeg, small);
  lca_count_kernel<<<grid, 256, 0, s>>>(gp, nt, tdeg, small + 8);
  dp_label_count_kernel<<<grid, 256, 0, s>>>(n, hdeg, hl, nlab, lcnt, small + 16);
  if (int rc = device_exclusive_scan(lcnt, nlab, loff, scan_tmp, s)) return fail(rc);
  DP_TRY(cudaMemsetAsync(lcnt, 0, (size_t)nlab * 4, s));
  dp_label_fill_kernel<<<grid, 256, 0, s>>>(n, hdeg, hl, nlab, loff, lcnt, H->rec, lpos);
  dp_label_sort_kernel<<<grid, 256, 0, s>>>(nlab, loff, lpos);
  if (int rc = device_exclusive_scan(tdeg, nt, toff, scan_tmp, s)) return fail(rc);
  DP_TRY(cudaMemsetAsync(tdeg, 0, (size_t)nt * 4, s));
  lca_fill_children_kernel<<<grid, 256, 0, s>>>(gp, nt, toff, tdeg, tadj);
  int32_t info[64];
  DP_TRY(cudaMemcpyAsync(info, small, 256, cudaMemcpyDeviceToHost, s));
  DP_TRY(cudaStreamSynchronize(s));
  if (info[8] != 1 || info[10]) return fail(set_error(PT_ERR_INVALID, "pt_who_build: guest is not a rooted tree (%d roots)", info[8]));
  if (info[19]) return fail(set_error(PT_ERR_INVALID, "pt_who_build: label id out of range (must be < n + nt)"));
  // arenas: the lists of the nodes of one guest level are disjoint sets of candidates over disjoint label sets, so a level holds at most
  // one entry per labelled host leaf; the total over all levels (and the scratch of the nodes that need the general case: ~17 words per
  // merged candidate, never reused -- there are no levels to reuse it between) is unknown in advance -> grow and repeat on overflow
  int64_t mcap = 2 * (n + nt) + 1024, scap = 48 * (n + 64) + 64 * 1024;
  int32_t* pending = (int32_t*)dalloc((size_t)nt * 4 + 4); DP_PTR(pending);
  for (int attempt = 0;; ++attempt) {
    int32_t* mpos = nullptr; int32_t* scratch = nullptr; unsigned long long* tops = nullptr;
    DP_TRY(lca_malloc((void**)&mpos, (size_t)mcap * 4)); owned.push_back(mpos);
    DP_TRY(lca_malloc((void**)&scratch, (size_t)scap * 4)); owned.push_back(scratch);
    DP_TRY(lca_malloc((void**)&tops, 64)); owned.push_back(tops);
    DP_TRY(cudaMemsetAsync(tops, 0, 64, s)); DP_TRY(cudaMemsetAsync(small + 32, 0, 64, s));
    DP_TRY(cudaMemsetAsync(mlen, 0xff, (size_t)nt * 4, s));
    DP_TRY(cudaMemcpyAsync(pending, tdeg, (size_t)nt * 4, cudaMemcpyDeviceToDevice, s));   // children still to come, per guest node
    DpArena A{mpos, mcap, scratch, scap, tops, small + 32};
    dp_who_flow_kernel<<<(unsigned)std::max<int64_t>(1, std::min<int64_t>((int64_t)dp.sms * 16, (nt + 255) / 256)), 256, 0, s>>>(
        nt, gp, pending, toff, tadj, gl, loff, lpos, nlab, H->rec, H->keypos, H->nodeat, H->table, H->d_level_base, moff, mlen, A);
    DP_TRY(cudaGetLastError());
    unsigned long long processed = 0;
    DP_TRY(cudaMemcpyAsync(info, small, 256, cudaMemcpyDeviceToHost, s));
    DP_TRY(cudaMemcpyAsync(&processed, tops + 2, 8, cudaMemcpyDeviceToHost, s));
    DP_TRY(cudaStreamSynchronize(s));
    if (processed != (unsigned long long)nt) return fail(set_error(PT_ERR_INVALID, "pt_who_build: guest parent array has a cycle"));
    if (info[32]) return fail(set_error(PT_ERR_UNSUPPORTED, "pt_who_build: a guest node whose children have several candidates each has more than %d children", DP_MAX_CHILDREN));
    if (!info[33] && !info[34]) {
      // compact into CSR order by guest node, positions -> host node ids
      int32_t* len = (int32_t*)dalloc((size_t)nt * 4 + 4); DP_PTR(len);
      DP_TRY(lca_malloc((void**)&W->d_off, (size_t)nt * 4 + 4));
      dp_lengths_kernel<<<grid, 256, 0, s>>>(nt, mlen, len);
      if (int rc = device_exclusive_scan(len, nt, W->d_off, scan_tmp, s)) return fail(rc);
      int32_t tot = 0;
      DP_TRY(cudaMemcpyAsync(&tot, W->d_off + nt, 4, cudaMemcpyDeviceToHost, s));
      DP_TRY(cudaStreamSynchronize(s));
      W->total = tot;
      DP_TRY(lca_malloc((void**)&W->d_adj, (size_t)std::max<int64_t>(tot, 1) * 4));
      dp_gather_kernel<<<grid, 256, 0, s>>>(nt, mlen, moff, mpos, H->nodeat, W->d_off, W->d_adj);
      DP_TRY(cudaGetLastError());
      break;
    }
    if (attempt >= 6) return fail(set_error(PT_ERR_NOMEM, "pt_who_build: the candidate lists do not fit (%lld list words, %lld scratch words)", (long long)mcap, (long long)scap));
    if (info[33]) mcap *= 4;
    if (info[34]) scap *= 4;
  }
#undef DP_TRY
#undef DP_PTR
  cleanup();
  *out = W;
  return PT_OK;
}

int pt_who_total(const pt_who* h, int64_t* total_entries) {
  if (!h || !total_entries) return set_error(PT_ERR_INVALID, "pt_who_total: null argument");
  *total_entries = h->total;
  return PT_OK;
}

int pt_who_lists(const pt_who* h, int32_t* out_off, int32_t* out_adj, pt_mem mem, void* stream) {
  if (!h || !out_off) return set_error(PT_ERR_INVALID, "pt_who_lists: null argument");
  cudaStream_t s = (cudaStream_t)stream;
  const cudaMemcpyKind kind = mem == PT_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  PT_CUDA(cudaMemcpyAsync(out_off, h->d_off, (size_t)(h->nt + 1) * 4, kind, s));
  if (out_adj && h->total > 0) PT_CUDA(cudaMemcpyAsync(out_adj, h->d_adj, (size_t)h->total * 4, kind, s));
  PT_CUDA(cudaStreamSynchronize(s));
  return PT_OK;
}

int pt_who_highest_displayed(const pt_who* h, int32_t host_root, int32_t* out, pt_mem mem, void* stream) {
  if (!h || !out || host_root < 0) return set_error(PT_ERR_INVALID, "pt_who_highest_displayed: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  int32_t* who1 = nullptr; int32_t* res = out;
  PT_CUDA(lca_malloc((void**)&who1, (size_t)h->nt * 4));
  if (mem == PT_MEM_HOST) { if (lca_malloc((void**)&res, (size_t)h->nt * 4) != cudaSuccess) { dev_free(who1); return set_error(PT_ERR_NOMEM, "out of device memory"); } }
  dp_who1_kernel<<<592, 256, 0, s>>>(h->nt, h->d_off, h->d_adj, host_root, who1);
  int rc = pt_tree_highest_displayed(h->d_gparent, who1, h->nt, host_root, res, PT_MEM_DEVICE, stream);
  if (rc == PT_OK && mem == PT_MEM_HOST) { if (cudaMemcpyAsync(out, res, (size_t)h->nt * 4, cudaMemcpyDeviceToHost, s) != cudaSuccess) rc = set_error(PT_ERR_CUDA, "copy failed"); }
  cudaStreamSynchronize(s);
  dev_free(who1);
  if (mem == PT_MEM_HOST) dev_free(res);
  return rc;
}

int pt_net_unzip(int64_t n, int64_t m, const int32_t* succ_off, const int32_t* succ_adj, const int32_t* label, int32_t root, int32_t* out_parent, int32_t* out_label,
                 int32_t* out_origin, int64_t capacity, int64_t* count, pt_mem mem, void* stream) {
  if (n <= 0 || m < 0 || !succ_off || (!succ_adj && m > 0) || !label || root < 0 || root >= n || !count || capacity < 0 || n > INT32_MAX / 2 || capacity > INT32_MAX / 2)
    return set_error(PT_ERR_INVALID, "pt_net_unzip: bad argument");
  if (int rc = require_device()) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  DeviceProps dp;
  if (int rc = device_props(&dp)) return rc;
  std::vector<void*> owned;
  auto cleanup = [&]() { cudaStreamSynchronize(s); for (void* p : owned) dev_free(p); };
  auto dalloc = [&](size_t bytes) -> void* { void* p = nullptr; if (lca_malloc(&p, bytes ? bytes : 4) != cudaSuccess) { cudaGetLastError(); return nullptr; } owned.push_back(p); return p; };
#define UZ_TRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup(); return set_error(PT_ERR_CUDA, "%s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); } } while (0)
#define UZ_PTR(p) do { if (!(p)) { cleanup(); return set_error(PT_ERR_NOMEM, "pt_net_unzip: out of device memory"); } } while (0)
  const int32_t *so = succ_off, *sa = succ_adj, *lb = label;
  if (mem == PT_MEM_HOST) {
    int32_t* a = (int32_t*)dalloc((size_t)(n + 1) * 4); int32_t* b = (int32_t*)dalloc((size_t)m * 4); int32_t* c = (int32_t*)dalloc((size_t)n * 4); UZ_PTR(a); UZ_PTR(b); UZ_PTR(c);
    UZ_TRY(cudaMemcpyAsync(a, succ_off, (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, s)); if (m) UZ_TRY(cudaMemcpyAsync(b, succ_adj, (size_t)m * 4, cudaMemcpyHostToDevice, s));
    UZ_TRY(cudaMemcpyAsync(c, label, (size_t)n * 4, cudaMemcpyHostToDevice, s));
    so = a; sa = b; lb = c;
  }
  long long* usize = (long long*)dalloc((size_t)n * 8); int32_t* flags = (int32_t*)dalloc(64); UZ_PTR(usize); UZ_PTR(flags);
  UZ_TRY(cudaMemsetAsync(flags, 0, 64, s));
  int occ = 0;
  UZ_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dp_unzip_sizes_kernel, 256, 0));
  {
    const int cgrid = (int)std::min<int64_t>((int64_t)dp.sms * std::max(1, std::min(occ, 4)), (n + 255) / 256);
    int64_t nn = n; long long cap = (long long)INT32_MAX / 2;
    void* args[] = {&nn, &so, &sa, &usize, &flags, &cap};
    UZ_TRY(cudaLaunchCooperativeKernel((void*)dp_unzip_sizes_kernel, dim3(std::max(cgrid, 1)), dim3(256), args, 0, s));
  }
  // the root of the unfolding is `root` itself whatever its out-degree (containment.hpp:274); a unary root has no arcs that count
  int32_t root_deg[2]; long long total = 1;
  UZ_TRY(cudaMemcpyAsync(root_deg, so + root, 8, cudaMemcpyDeviceToHost, s));
  UZ_TRY(cudaStreamSynchronize(s));
  if (root_deg[1] - root_deg[0] != 1) { UZ_TRY(cudaMemcpyAsync(&total, usize + root, 8, cudaMemcpyDeviceToHost, s)); UZ_TRY(cudaStreamSynchronize(s)); }
  if (total < 0) { cleanup(); return set_error(PT_ERR_INVALID, "pt_net_unzip: the network below the root has a cycle"); }
  *count = (int64_t)total;
  if (!out_parent || !out_label || !out_origin) { cleanup(); return PT_OK; }   // size query
  if (total > capacity) { cleanup(); return set_error(PT_ERR_UNSUPPORTED, "pt_net_unzip: the MUL-tree has %lld nodes, capacity %lld", total, (long long)capacity); }
  int32_t *dpar = out_parent, *dlab = out_label, *dorg = out_origin;
  if (mem == PT_MEM_HOST) { dpar = (int32_t*)dalloc((size_t)total * 4); dlab = (int32_t*)dalloc((size_t)total * 4); dorg = (int32_t*)dalloc((size_t)total * 4); UZ_PTR(dpar); UZ_PTR(dlab); UZ_PTR(dorg); }
  int32_t* fa = (int32_t*)dalloc((size_t)total * 8 + 8); int32_t* fb = (int32_t*)dalloc((size_t)total * 8 + 8); int32_t* ctr = (int32_t*)dalloc(64); UZ_PTR(fa); UZ_PTR(fb); UZ_PTR(ctr);
  UZ_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dp_unzip_expand_kernel, 256, 0));
  {
    const int cgrid = (int)std::min<int64_t>((int64_t)dp.sms * std::max(1, std::min(occ, 4)), (total + 255) / 256);
    int64_t nn = n; const long long* us = usize;
    void* args[] = {&nn, &so, &sa, &lb, &us, &root, &fa, &fb, &ctr, &dpar, &dlab, &dorg};
    UZ_TRY(cudaLaunchCooperativeKernel((void*)dp_unzip_expand_kernel, dim3(std::max(cgrid, 1)), dim3(256), args, 0, s));
  }
  if (mem == PT_MEM_HOST) {
    UZ_TRY(cudaMemcpyAsync(out_parent, dpar, (size_t)total * 4, cudaMemcpyDeviceToHost, s)); UZ_TRY(cudaMemcpyAsync(out_label, dlab, (size_t)total * 4, cudaMemcpyDeviceToHost, s));
    UZ_TRY(cudaMemcpyAsync(out_origin, dorg, (size_t)total * 4, cudaMemcpyDeviceToHost, s));
  }
  UZ_TRY(cudaStreamSynchronize(s));
#undef UZ_TRY
#undef UZ_PTR
  cleanup();
  return PT_OK;
}

int pt_net_tree_components(int64_t n, int64_t m, const int32_t* succ_off, const int32_t* succ_adj, int32_t* comp_root, int32_t* visible_leaf, int64_t* dag_edges, int64_t dag_capacity,
                           int64_t* dag_count, pt_mem mem, void* stream) {
  if (n <= 0 || m < 0 || !succ_off || (!succ_adj && m > 0) || !comp_root || n > INT32_MAX / 2 || m > INT32_MAX / 2) return set_error(PT_ERR_INVALID, "pt_net_tree_components: bad argument");
  if (int rc = require_device()) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  DeviceProps dp;
  if (int rc = device_props(&dp)) return rc;
  std::vector<void*> owned;
  auto cleanup = [&]() { cudaStreamSynchronize(s); for (void* p : owned) dev_free(p); };
  auto dalloc = [&](size_t bytes) -> void* { void* p = nullptr; if (lca_malloc(&p, bytes ? bytes : 4) != cudaSuccess) { cudaGetLastError(); return nullptr; } owned.push_back(p); return p; };
#define TCO_TRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup(); return set_error(PT_ERR_CUDA, "%s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); } } while (0)
#define TCO_PTR(p) do { if (!(p)) { cleanup(); return set_error(PT_ERR_NOMEM, "pt_net_tree_components: out of device memory"); } } while (0)
  const int32_t *so = succ_off, *sa = succ_adj;
  if (mem == PT_MEM_HOST) {
    int32_t* a = (int32_t*)dalloc((size_t)(n + 1) * 4); int32_t* b = (int32_t*)dalloc((size_t)m * 4); TCO_PTR(a); TCO_PTR(b);
    TCO_TRY(cudaMemcpyAsync(a, succ_off, (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, s)); if (m) TCO_TRY(cudaMemcpyAsync(b, succ_adj, (size_t)m * 4, cudaMemcpyHostToDevice, s));
    so = a; sa = b;
  }
  const int grid = dp.sms * 8;
  int32_t* indeg = (int32_t*)dalloc((size_t)n * 4 + 4); int32_t* onep = (int32_t*)dalloc((size_t)n * 4); int32_t* comp = (int32_t*)dalloc((size_t)n * 4); int32_t* vis = (int32_t*)dalloc((size_t)n * 4);
  int32_t* poff = (int32_t*)dalloc((size_t)n * 4 + 4); int32_t* padj = (int32_t*)dalloc((size_t)m * 4); int32_t* cursor = (int32_t*)dalloc((size_t)n * 4 + 4); int32_t* flag = (int32_t*)dalloc(64);
  int64_t* scan_tmp = (int64_t*)dalloc(((size_t)((n + SCAN_TILE - 1) / SCAN_TILE) + 1) * 8);
  TCO_PTR(indeg); TCO_PTR(onep); TCO_PTR(comp); TCO_PTR(vis); TCO_PTR(poff); TCO_PTR(padj); TCO_PTR(cursor); TCO_PTR(flag); TCO_PTR(scan_tmp);
  TCO_TRY(cudaMemsetAsync(indeg, 0, (size_t)n * 4 + 4, s)); TCO_TRY(cudaMemsetAsync(onep, 0xff, (size_t)n * 4, s)); TCO_TRY(cudaMemsetAsync(vis, 0xff, (size_t)n * 4, s));
  tc_indeg_kernel<<<grid, 256, 0, s>>>(n, so, sa, indeg, onep);
  // predecessor CSR (br_tree_parent-style fill through atomics; the order inside a list does not matter here)
  if (int rc = device_exclusive_scan(indeg, n, poff, scan_tmp, s)) { cleanup(); return rc; }
  TCO_TRY(cudaMemsetAsync(cursor, 0, (size_t)n * 4 + 4, s));
  tc_pred_fill_kernel<<<grid, 256, 0, s>>>(n, so, sa, poff, cursor, padj);
  tc_init_kernel<<<grid, 256, 0, s>>>(n, so, indeg, onep, comp);
  for (int it = 0; it < 64; ++it) {
    int32_t h = 0;
    TCO_TRY(cudaMemsetAsync(flag, 0, 4, s));
    tc_jump_kernel<<<grid, 256, 0, s>>>(n, comp, flag);
    TCO_TRY(cudaMemcpyAsync(&h, flag, 4, cudaMemcpyDeviceToHost, s));
    TCO_TRY(cudaStreamSynchronize(s));
    if (!h) break;
  }
  for (int64_t it = 0; it <= n; ++it) {   // one round per level of stacked reticulations
    int32_t h = 0;
    TCO_TRY(cudaMemsetAsync(flag, 0, 4, s));
    tc_consensus_kernel<<<grid, 256, 0, s>>>(n, so, indeg, poff, padj, comp, flag);
    TCO_TRY(cudaMemcpyAsync(&h, flag, 4, cudaMemcpyDeviceToHost, s));
    TCO_TRY(cudaStreamSynchronize(s));
    if (!h) break;
  }
  tc_visible_kernel<<<grid, 256, 0, s>>>(n, so, comp, vis);
  const cudaMemcpyKind kind = mem == PT_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  tc_output_kernel<<<grid, 256, 0, s>>>(n, so, indeg, comp, cursor);   // cursor is free again
  TCO_TRY(cudaMemcpyAsync(comp_root, cursor, (size_t)n * 4, kind, s));
  if (visible_leaf) TCO_TRY(cudaMemcpyAsync(visible_leaf, vis, (size_t)n * 4, kind, s));
  if (dag_count) {
    int32_t* cnt = (int32_t*)dalloc(64); TCO_PTR(cnt);
    const int64_t cap = dag_edges ? dag_capacity : 0;
    long long* de = (long long*)dalloc((size_t)std::max<int64_t>(cap, 1) * 8); TCO_PTR(de);
    TCO_TRY(cudaMemsetAsync(cnt, 0, 64, s));
    tc_dag_kernel<<<grid, 256, 0, s>>>(n, so, indeg, onep, poff, padj, comp, cnt, de, cap, cnt + 1);
    int32_t hc[2] = {0, 0};
    TCO_TRY(cudaMemcpyAsync(hc, cnt, 8, cudaMemcpyDeviceToHost, s));
    TCO_TRY(cudaStreamSynchronize(s));
    if (hc[1]) { cleanup(); return set_error(PT_ERR_UNSUPPORTED, "pt_net_tree_components: more than 64 stacked reticulations above a component root"); }
    *dag_count = hc[0];   // with duplicates; the facade removes them (the reference collects them in a set)
    if (dag_edges && hc[0] <= cap && hc[0] > 0) TCO_TRY(cudaMemcpyAsync(dag_edges, de, (size_t)hc[0] * 8, kind, s));
  }
  TCO_TRY(cudaStreamSynchronize(s));
#undef TCO_TRY
#undef TCO_PTR
  cleanup();
  return PT_OK;
}

// ---- VisibleComponentRule (utils/containment.hpp:603-725) as a query: which tree component the rule would treat, and what it would match it with.
// The device does the work that grows with the network -- component roots, visible leaves and the reticulation-path walks behind the component DAG
// (pt_net_tree_components: one arc per PATH, so the multiplicity of an arc IS the path count is_half_eligible :664-676 accumulates), the unfolding
// below the chosen root (pt_net_unzip), the tree-in-tree DP on it and the climb (pt_who_*); the eligibility test itself runs over the component
// DAG (at most r + 1 nodes) on the host.  Path counts are taken ONCE per arc: the reference re-walks the reticulations above a root v for every
// component above v and adds to the same counters each time (:670), so its answer depends on the order in which its hash containers yield the
// roots; what it calls half-eligible is always half-eligible here (counts only grow there), see tests/test_oracle.py.
int pt_net_visible_component(int64_t n, int64_t m, const int32_t* succ_off, const int32_t* succ_adj, const int32_t* host_label, int64_t nt, const int32_t* guest_parent,
                             const int32_t* guest_label, int32_t want_root, int32_t want_leaf, int64_t unzip_cap, uint8_t* eligibility, int32_t* root_out, int32_t* leaf_out,
                             int32_t* guest_node_out, void* stream) {
  if (n <= 0 || m < 0 || nt <= 0 || !succ_off || (!succ_adj && m > 0) || !host_label || !guest_parent || !guest_label || !root_out || !leaf_out || !guest_node_out ||
      want_root >= n || want_leaf >= n)
    return set_error(PT_ERR_INVALID, "pt_net_visible_component: bad argument");
  *root_out = *leaf_out = *guest_node_out = -1;
  std::vector<int32_t> comp((size_t)n), vis((size_t)n);
  std::vector<int64_t> arcs((size_t)std::max<int64_t>(4 * m, 16));
  int64_t cnt = 0;
  if (int rc = pt_net_tree_components(n, m, succ_off, succ_adj, comp.data(), vis.data(), arcs.data(), (int64_t)arcs.size(), &cnt, PT_MEM_HOST, stream)) return rc;
  if (cnt > (int64_t)arcs.size()) {
    arcs.assign((size_t)cnt, 0);
    if (int rc = pt_net_tree_components(n, m, succ_off, succ_adj, comp.data(), vis.data(), arcs.data(), (int64_t)arcs.size(), &cnt, PT_MEM_HOST, stream)) return rc;
  }
  arcs.resize((size_t)cnt);
  std::sort(arcs.begin(), arcs.end());
  // component DAG with multiplicities: kids[c] = (root below, number of reticulation paths from tree nodes of c to it)
  std::vector<int32_t> indeg((size_t)n, 0);
  for (int64_t e = 0; e < m; ++e) if (succ_adj[e] >= 0 && succ_adj[e] < n) ++indeg[(size_t)succ_adj[e]];
  int32_t net_root = -1;
  for (int64_t v = 0; v < n; ++v) if (indeg[(size_t)v] == 0) { net_root = (int32_t)v; break; }
  if (net_root < 0) return set_error(PT_ERR_INVALID, "pt_net_visible_component: the network has no root");
  std::vector<int32_t> node_of;                 // component-DAG nodes
  std::vector<int32_t> idx((size_t)n, -1);
  auto id_of = [&](int32_t v) { if (idx[(size_t)v] < 0) { idx[(size_t)v] = (int32_t)node_of.size(); node_of.push_back(v); } return idx[(size_t)v]; };
  id_of(net_root);
  std::vector<std::vector<std::pair<int32_t, int32_t>>> kids;
  for (size_t i = 0; i < arcs.size();) {
    size_t j = i;
    while (j < arcs.size() && arcs[j] == arcs[i]) ++j;
    const int32_t a = id_of((int32_t)(arcs[i] >> 32)), b = id_of((int32_t)(arcs[i] & 0xffffffffll));
    if (kids.size() < node_of.size()) kids.resize(node_of.size());
    kids[(size_t)a].emplace_back(b, (int32_t)(j - i));
    i = j;
  }
  kids.resize(node_of.size());
  const int K = (int)node_of.size();
  // children first (iterative post-order over the DAG from every node), then the test of :652-687 with every count taken once
  std::vector<int8_t> half((size_t)K, -1);
  std::vector<uint8_t> elig_below((size_t)K, 0);
  std::vector<std::vector<std::pair<int32_t, int32_t>>> paths((size_t)K);   // (root below, paths) for half-eligible nodes, sorted by root
  std::vector<std::pair<int, size_t>> stack;
  for (int start = 0; start < K; ++start) {
    if (half[(size_t)start] >= 0) continue;
    stack.emplace_back(start, 0);
    half[(size_t)start] = -2;   // on the stack
    while (!stack.empty()) {
      const int u = stack.back().first;
      if (stack.back().second < kids[(size_t)u].size()) {
        const int v = kids[(size_t)u][stack.back().second++].first;
        if (half[(size_t)v] == -1) { half[(size_t)v] = -2; stack.emplace_back(v, 0); }
        else if (half[(size_t)v] == -2) return set_error(PT_ERR_INVALID, "pt_net_visible_component: the component DAG has a cycle");
        continue;
      }
      stack.pop_back();
      bool ok = true;
      std::vector<std::pair<int32_t, int32_t>> acc;
      for (const auto& vc : kids[(size_t)u]) {
        const int v = vc.first;
        elig_below[(size_t)u] |= elig_below[(size_t)v] | (uint8_t)(half[(size_t)v] == 1 && vis[(size_t)node_of[(size_t)v]] >= 0);
        if (half[(size_t)v] != 1 || vc.second > 1) { ok = false; continue; }
        acc.emplace_back(v, 1);
        for (const auto& xp : paths[(size_t)v]) acc.push_back(xp);
      }
      if (ok) {
        std::sort(acc.begin(), acc.end());
        for (size_t i = 0; i < acc.size() && ok;) {
          size_t j = i; int32_t c = 0;
          while (j < acc.size() && acc[j].first == acc[i].first) c += acc[j++].second;
          if (c > 1) ok = false;
          paths[(size_t)u].emplace_back(acc[i].first, c);
          i = j;
        }
        if (!ok) paths[(size_t)u].clear();
      }
      half[(size_t)u] = ok ? 1 : 0;
    }
  }
  if (eligibility) {
    memset(eligibility, 0, (size_t)n);
    for (int k = 0; k < K; ++k) {
      const int32_t v = node_of[(size_t)k];
      const bool e = half[(size_t)k] == 1 && vis[(size_t)v] >= 0;
      eligibility[(size_t)v] = (uint8_t)((half[(size_t)k] == 1 ? 1 : 0) | (e ? 2 : 0) | (e && !elig_below[(size_t)k] ? 4 : 0));
    }
  }
  // the root the rule treats: the first eligible one in a post-order of the component DAG (:689-704), i.e. an eligible root without an eligible
  // root below it -- the one with the smallest id here -- or the caller's choice
  int32_t rt = want_root;
  if (rt >= 0) { if (idx[(size_t)rt] < 0) return set_error(PT_ERR_INVALID, "pt_net_visible_component: node %d is not a component root", rt); }
  else for (int k = 0; k < K; ++k) {
    const int32_t v = node_of[(size_t)k];
    if (half[(size_t)k] == 1 && vis[(size_t)v] >= 0 && !elig_below[(size_t)k] && (rt < 0 || v < rt)) rt = v;
  }
  if (rt < 0) return PT_OK;
  *root_out = rt;
  const int32_t leaf = want_leaf >= 0 ? want_leaf : vis[(size_t)rt];
  *leaf_out = leaf;
  if (leaf < 0 || host_label[(size_t)leaf] < 0) return PT_OK;
  // treat_comp_root :623-644: the guest leaf with the visible leaf's label, then the highest ancestor still displayed below rt
  std::vector<uint8_t> has_child((size_t)nt, 0);
  for (int64_t t = 0; t < nt; ++t) if (guest_parent[(size_t)t] >= 0 && guest_parent[(size_t)t] < nt) has_child[(size_t)guest_parent[(size_t)t]] = 1;
  int32_t gleaf = -1;
  for (int64_t t = 0; t < nt; ++t) if (!has_child[(size_t)t] && guest_label[(size_t)t] == host_label[(size_t)leaf]) { gleaf = (int32_t)t; break; }
  if (gleaf < 0) return PT_OK;
  int64_t count = 0;
  if (int rc = pt_net_unzip(n, m, succ_off, succ_adj, host_label, rt, nullptr, nullptr, nullptr, 0, &count, PT_MEM_HOST, stream)) return rc;
  if (count > (unzip_cap > 0 ? unzip_cap : ((int64_t)1 << 26)))
    return set_error(PT_ERR_UNSUPPORTED, "pt_net_visible_component: the unfolding below node %d has %lld nodes (cap %lld)", rt, (long long)count, (long long)unzip_cap);
  std::vector<int32_t> par((size_t)count), lab((size_t)count), org((size_t)count), hda((size_t)nt);
  if (int rc = pt_net_unzip(n, m, succ_off, succ_adj, host_label, rt, par.data(), lab.data(), org.data(), count, &count, PT_MEM_HOST, stream)) return rc;
  pt_who* W = nullptr;
  if (int rc = pt_who_build(&W, par.data(), lab.data(), count, guest_parent, guest_label, nt, PT_MEM_HOST, stream)) return rc;
  const int rc = pt_who_highest_displayed(W, 0, hda.data(), PT_MEM_HOST, stream);
  pt_who_free(W);
  if (rc) return rc;
  *guest_node_out = hda[(size_t)gleaf];
  return PT_OK;
}

}  // extern "C"
