// csrc/tc_core.cuh -- the tree-containment reducer, written ONCE as barrier-separated data-parallel passes.
//
// Replaces the reference's sequential graph rewriting (utils/containment.hpp:380-1272: NodeSuppresser,
// ReticulationMerger, CherryPicker, HostGuestMatch, the apply_rules loop :1054-1083 and the branching of
// displayed() :1202-1270) with level-synchronous passes over flat arrays.  Every loop below is a
// `PT_FOR(i, n)` = "for all i in [0,n) in parallel, in any order" followed by a barrier; the executor `Ex`
// decides what runs it: one warp per instance with the instance in shared memory (tc_kernels.cu), or the
// single-threaded simulator used by the CPU tests (tests/host_sim/, which also permutes the iteration order
// to expose order dependence).  Nothing here is CPU product code: the library only instantiates the CUDA
// executor.
//
// Rules (all verdict-preserving for arbitrary rooted DAG hosts; soundness arguments in DESIGN.md 4):
//   clean      dead-end removal, in=out=1 suppression, out-degree-1 root removal, reticulation-chain merge,
//              parallel-arc merge                                       (cf. containment.hpp:392-409,756-824)
//   TC         true (multi-)cherry: a host node whose children are all single-parent leaves must equal a guest
//              node whose children are exactly those leaves -> collapse both, else NOT displayed
//                                                                       (cf. CherryPicker :538-590, match_nodes :915-950)
//   RC         reticulated cherry: guest cherry {x,y}, host leaf x hangs on q, y (or the out-degree-1
//              reticulation above y) has q among its parents -> keep that in-arc, delete the others
//   SC         sibling constraint: guest cherry {x,y}, host leaf x hangs on q: everything else below q may only
//              lead to y.  Arcs into multi-parent nodes that cannot reach y are deleted; single-parent subtrees
//              that cannot reach y must die completely (a labelled leaf there -> NOT displayed)
//                                                                       (generalises :564-580)
//   branch     nothing applies: pick the lowest multi-parent node and emit one child instance per in-arc
//              (cf. displayed() :1225-1266; the reference deep-copies the checker per branch, we append compacted
//              instances to a pool that the next kernel launch consumes)
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define PT_HD __host__ __device__ __forceinline__
// Code size matters here: the first version was 156 KB of SASS and 40 % of its stall samples were instruction-cache misses
// (profiles/).  Real (noinline) calls were worse: the workspace pointers then live in local memory.  So: everything inline,
// every pass has ONE call site, and no loop is unrolled (PT_NOUNROLL): the kernel is latency bound, not issue bound.
#define PT_NI __host__ __device__ __forceinline__
#if defined(__CUDA_ARCH__)
#define PT_NOUNROLL _Pragma("unroll 1")
#else
#define PT_NOUNROLL
#endif
#else
#define PT_HD inline
#define PT_NI
#define PT_NOUNROLL
#endif

namespace ptgpu {
#ifdef PT_SIM_STATS
static long g_sim_csr_calls, g_sim_live_nodes, g_sim_nodes, g_sim_live_arcs;
#endif

// parallel-for over [0,n): i_ is the executor's slot, i the (possibly permuted) element
// (a tiled form with a group-uniform trip count and a predicated tail was tried: 2 % MORE instructions, no fewer BSSY / BSYNC)
#define PT_FOR(i, n) \
  PT_NOUNROLL for (int i##_ = ex.tid(), i = 0; i##_ < (n) && ((i = ex.map(i##_, (n))), true); i##_ += ex.nth())
// the passes that read the caller's arrays from global memory (ex.ldg) keep four iterations in flight, so that four load
// latencies overlap instead of one load -> store round trip per iteration.  (An out-of-line helper shared by the five
// arrays was tried: 72 instead of 75 KB of SASS, but 6 % slower end to end.)
#if defined(__CUDA_ARCH__)
#define PT_FOR_LD(i, n) \
  _Pragma("unroll 4") for (int i##_ = ex.tid(), i = 0; i##_ < (n) && ((i = ex.map(i##_, (n))), true); i##_ += ex.nth())
#else
#define PT_FOR_LD(i, n) PT_FOR(i, n)
#endif

enum TcCode : int { TC_NOT = 0, TC_DISPLAYED = 1, TC_BRANCH = 2, TC_ERROR = 3, TC_CONTINUE = 4 };
enum TcStatus : int { TCS_OK = 0, TCS_MULTILABEL = 1, TCS_BAD_GRAPH = 2, TCS_TOO_LARGE = 3, TCS_POOL_OVERFLOW = 4, TCS_ROUND_LIMIT = 5 };

// one instance in the flat layout of include/pt_gpu.h (pointers already offset to the instance)
struct TcInst {
  int n, m, nt;
  const void* succ_off;  // n+1, local
  const void* succ_adj;  // m
  const void* hlabel;    // n
  const void* tparent;   // nt
  const void* tlabel;    // nt
                         // element type S of the five arrays: int32_t, or int16_t for pt_batch_desc.index_bytes = 2
};
// Element type S of the caller's five arrays and what it implies for the layout (include/pt_gpu.h, pt_batch_desc.index_bytes):
//   int32_t / int16_t : succ_off holds n+1 arc positions, -1 = "none" for labels and the guest root
//   uint8_t           : succ_off holds the n OUT-DEGREES (arc positions may exceed 255 while every id still fits a byte),
//                       255 = "none"; the offsets are rebuilt by one prefix sum when the instance is loaded
template <class S> struct TcLayout {
  static constexpr bool kDegrees = false;
  PT_HD static int dec(int x) { return x; }
  PT_HD static S enc(int x) { return (S)x; }
};
template <> struct TcLayout<uint8_t> {
  static constexpr bool kDegrees = true;
  PT_HD static int dec(int x) { return x == 255 ? -1 : x; }
  PT_HD static uint8_t enc(int x) { return (uint8_t)(x < 0 ? 255 : x); }
};
// where a branch child is written (same layout as the input of element type S; the caller provides capacity n+1 / m / n / nt / nt)
template <class S>
struct TcEmitT {
  S* succ_off; S* succ_adj; S* hlabel; S* tparent; S* tlabel;
  int32_t* dims;  // n, m, nt, num_labels
};
typedef TcEmitT<int32_t> TcEmit;

struct TcStats { unsigned cherries, retcherries, arcs_deleted, passes; };

template <class I>
struct TcWork {
  int capN, capM, capNT, capL, W;  // W = words of the reachable-label signature per node (exact when 32 W >= labels)
  // host network
  I *at, *ah;            // arc tail / head; at < 0 = deleted.  Arcs stay SORTED BY TAIL: rewriting only deletes arcs or moves heads
  I *ooff;               // out CSR: the out-arcs of v are the arc ids [ooff[v], ooff[v+1]) (csr() compacts the arc arrays)
  I *ioff, *iarc;        // in CSR (arc ids)
  I *atmp;               // only for executors that cannot compact the arc arrays in place (cluster executor): capM spare entries
  I *hlab;               // label of a live labelled leaf, else -1 (capN+1: csr() swaps it with lvl when it renumbers the nodes)
  I *lvl;                // height (longest path to a leaf) from the wavefront; scratch otherwise
  I *hcnt, *haux;        // atomic scratch, max(capN, capNT)+1 each (the guest passes of clean_guest / emit_child borrow them)
  I* hscr;               // FIFO of the wavefront / scan target of emit_child, capN+1
  uint32_t* lset;        // capN * W : labels reachable below each node
  // guest tree
  I *tpar, *tlab;        // tpar: -1 root, -2 deleted
  I *tdeg, *tmax;        // tdeg: live children; tmax: XOR of the children's ids
  // labels
  I *l2h, *l2t;
  // scalars (shared)
  int32_t* sc;           // [0] flag  [1] fail  [2] root  [3] troot  [4] nlabels  [5] best  [6..] scratch
  static constexpr int kScalars = 16;

  // signature words: exact label sets up to kMaxExactWords * 32 labels, hashed (label mod bits) beyond: a hashed set is a
  // SUPERSET, so "cannot reach" stays certain and the sibling rule stays sound, it only fires less often.  Measured in the
  // simulator (gen instances l=100..1000): 32-, 64- and 512-bit signatures give the same number of branchings to within
  // 3 %, because the rule only ever needs to separate a handful of nearby leaves; 32 bits keep the working set small
  // (shared-memory occupancy) and make the rule independent of the instance size (n = 10^6 needs 8 MB, not 62 GB).
  static constexpr int kMaxExactWords = 1;
  PT_HD static int sig_words(int L) { const int w = (L + 31) / 32; return w < 1 ? 1 : (w > kMaxExactWords ? kMaxExactWords : w); }
  PT_HD static size_t bytes(int N, int M, int NT, int L, bool spare_arcs = false) {
    const int W = sig_words(L), S = (N > NT ? N : NT) + 1;
    size_t b = 0;
    b += (size_t)(3 * M + 5 * (N + 1) + 2 * S) * sizeof(I);   // at ah iarc | ooff ioff hlab lvl hscr | hcnt haux
    if (spare_arcs) b += (size_t)M * sizeof(I);               // atmp
    b += (size_t)(2 * NT + 2 * (NT + 1) + 2 * L) * sizeof(I); // tpar tlab | tdeg tmax | l2h l2t
    b = (b + 3) & ~(size_t)3;
    b += (size_t)N * W * 4;                                   // lset
    b += kScalars * 4;
    return (b + 15) & ~(size_t)15;
  }
  // All pointers are formed as base + offset (no pointer -> integer -> pointer round trip): the CUDA compiler then keeps the
  // address space of `base`, and the warp executor gets LDS / STS / ATOMS instead of generic LD / ST / ATOM.
  PT_HD void carve(char* base, int N, int M, int NT, int L, bool spare_arcs = false) {
    capN = N; capM = M; capNT = NT; capL = L; W = sig_words(L);
    const int S = (N > NT ? N : NT) + 1;
    size_t o = 0;
    auto take = [&](size_t nbytes) { const size_t r = o; o += nbytes; return r; };
    at = (I*)(base + take((size_t)M * sizeof(I))); ah = (I*)(base + take((size_t)M * sizeof(I))); iarc = (I*)(base + take((size_t)M * sizeof(I)));
    atmp = spare_arcs ? (I*)(base + take((size_t)M * sizeof(I))) : nullptr;
    ooff = (I*)(base + take((size_t)(N + 1) * sizeof(I))); ioff = (I*)(base + take((size_t)(N + 1) * sizeof(I)));
    hlab = (I*)(base + take((size_t)(N + 1) * sizeof(I))); lvl = (I*)(base + take((size_t)(N + 1) * sizeof(I)));
    hscr = (I*)(base + take((size_t)(N + 1) * sizeof(I)));
    hcnt = (I*)(base + take((size_t)S * sizeof(I))); haux = (I*)(base + take((size_t)S * sizeof(I)));
    tpar = (I*)(base + take((size_t)NT * sizeof(I))); tlab = (I*)(base + take((size_t)NT * sizeof(I)));
    tdeg = (I*)(base + take((size_t)(NT + 1) * sizeof(I))); tmax = (I*)(base + take((size_t)(NT + 1) * sizeof(I)));
    l2h = (I*)(base + take((size_t)L * sizeof(I))); l2t = (I*)(base + take((size_t)L * sizeof(I)));
    o = (o + 3) & ~(size_t)3;
    lset = (uint32_t*)(base + take((size_t)N * W * 4));
    sc = (int32_t*)(base + take(kScalars * 4));
  }
};

// ---------------------------------------------------------------------------------------------------------
template <class Ex, class I>
struct TcSolver {
  Ex& ex;
  TcWork<I>& w;
  int n, m, nt, L;  // uniform across the group
  TcStats st;

  PT_HD TcSolver(Ex& e, TcWork<I>& work) : ex(e), w(work), n(0), m(0), nt(0), L(0), core_switchings_(0), allow_core_(true), missing_(0), dag_state_(0) { st.cherries = st.retcherries = st.arcs_deleted = st.passes = 0; }

  enum { F_FLAG = 0, F_FAIL = 1, F_ROOT = 2, F_TROOT = 3, F_NLAB = 4, F_BEST = 5, F_STATUS = 6, F_TMP = 7, F_TMP2 = 8, F_QTAIL = 9 };
  static constexpr int TDEAD = -2;
  static constexpr int LVL_INF = sizeof(I) == 2 ? 32000 : 0x3fffffff;

  // ---- small helpers -----------------------------------------------------------------------------------
  PT_HD int outdeg(int v) const { return w.ooff[v + 1] - w.ooff[v]; }
  PT_HD int indeg(int v) const { return w.ioff[v + 1] - w.ioff[v]; }
  PT_HD int out_arc(int v, int k) const { return w.ooff[v] + k; }
  PT_HD int in_arc(int v, int k) const { return w.iarc[w.ioff[v] + k]; }
  PT_HD bool lbit(int v, int lam) const { const int b = lam % (w.W * 32); return (w.lset[v * w.W + (b >> 5)] >> (b & 31)) & 1u; }
  // read a shared flag uniformly and reset it
  // A loop / branch decision must be the SAME value in every thread of the group.  ex.uniform() has one thread read the
  // scalar and broadcasts it (shuffle / shared memory), so the decision cannot diverge even where the workspace lives in
  // global memory and threads could otherwise observe it through differently aged cache lines.
  PT_HD int take_flag(int which) {
    const int v = ex.uniform(&w.sc[which]);
    if (ex.tid() == 0) w.sc[which] = 0;
    ex.sync();
    return v;
  }
  PT_HD void set_scalar(int which, int v) { if (ex.tid() == 0) w.sc[which] = v; ex.sync(); }

  // ---- load + validate ---------------------------------------------------------------------------------
  // returns TCS_* ; sets up arcs, labels, guest arrays.  Leaves w.sc[F_FAIL]=1 if a guest label is missing in the host.
  template <class S>
  PT_NI int load(const TcInst& in) {
    ex.phase = 1;
    n = in.n; m = in.m; nt = in.nt;
    if (n < 0 || m < 0 || nt < 0) return TCS_BAD_GRAPH;   // malformed offset tables (batch_inst)
    if (n > w.capN || m > w.capM || nt > w.capNT) return TCS_TOO_LARGE;
    if (ex.tid() == 0) { PT_NOUNROLL for (int k = 0; k < TcWork<I>::kScalars; ++k) w.sc[k] = 0; }
    ex.sync();
    // structure checks on the CSR offsets (branch-free bodies: the loads of consecutive iterations overlap)
    const S* const succ_off = static_cast<const S*>(in.succ_off); const S* const succ_adj = static_cast<const S*>(in.succ_adj);
    const S* const hlabel = static_cast<const S*>(in.hlabel); const S* const tparent = static_cast<const S*>(in.tparent);
    const S* const tlabel = static_cast<const S*>(in.tlabel);
    typedef TcLayout<S> Lay;
    int bad = 0, big = 0, maxlab = -1;
    if (Lay::kDegrees) {
      // out-degrees: their sum must be m (checked before anything is indexed by them), the offsets are their prefix sums
      int sum = 0;
      PT_FOR_LD(v, n) { const int d = ex.ldg(&succ_off[v]); sum += d; w.hcnt[v] = (I)d; }
      if (ex.reduce_add(sum) != m) return TCS_BAD_GRAPH;
      ex.exclusive_scan(w.hcnt, w.ooff, n);
    } else {
      PT_FOR_LD(v, n + 1) {
        const int o = ex.ldg(&succ_off[v]);
        bad |= (o < 0) | (o > m) | (v == 0 && o != 0) | (v == n && o != m);
        w.ooff[v] = (I)(o < 0 ? 0 : (o > m ? m : o));
      }
    }
    PT_FOR_LD(v, n) {
      const int l = Lay::dec(ex.ldg(&hlabel[v]));
      w.hlab[v] = (I)l;
      big |= (l >= w.capL) | (l < -1);
      maxlab = l > maxlab ? l : maxlab;
    }
    PT_FOR_LD(e, m) {
      const int h = ex.ldg(&succ_adj[e]);
      bad |= (h < 0) | (h >= n);
      w.ah[e] = (I)h;
    }
    PT_FOR_LD(t, nt) {
      const int p = Lay::dec(ex.ldg(&tparent[t])), l = Lay::dec(ex.ldg(&tlabel[t]));
      bad |= (p < -1) | (p >= nt) | (p == t);
      big |= (l >= w.capL) | (l < -1);
      maxlab = l > maxlab ? l : maxlab;
      w.tpar[t] = (I)p; w.tlab[t] = (I)l; w.tdeg[t] = (I)0;
    }
    if (bad) w.sc[F_STATUS] = TCS_BAD_GRAPH; else if (big) w.sc[F_TMP2] = 1;
    ex.sync();
    PT_FOR(v, n) {
      if (w.ooff[v + 1] < w.ooff[v]) w.sc[F_STATUS] = TCS_BAD_GRAPH;
      PT_NOUNROLL for (int e = w.ooff[v]; e < w.ooff[v + 1]; ++e) w.at[e] = (I)v;
    }
    ex.atomic_max(&w.sc[F_NLAB], maxlab + 1);
    if (take_flag(F_STATUS)) return TCS_BAD_GRAPH;
    if (take_flag(F_TMP2)) return TCS_TOO_LARGE;
    L = ex.uniform(&w.sc[F_NLAB]);
    PT_FOR(l, L) { w.l2h[l] = (I)-1; w.l2t[l] = (I)-1; }
    ex.phase = 15;
    // guest: child counts, exactly one root, no cycles
    PT_FOR(t, nt) { if (w.tpar[t] >= 0) ex.atomic_add_idx(&w.tdeg[w.tpar[t]], 1); else ex.atomic_add(&w.sc[F_TMP], 1); }
    if (nt > 0 && ex.uniform(&w.sc[F_TMP]) != 1) return TCS_BAD_GRAPH;
    PT_FOR(t, nt) {
      int x = t, steps = 0;
      while (w.tpar[x] >= 0 && steps <= nt) { x = w.tpar[x]; ++steps; }
      if (steps > nt) w.sc[F_STATUS] = TCS_BAD_GRAPH;
      if (w.tpar[t] < 0) w.sc[F_TROOT] = t;
    }
    set_scalar(F_TMP, 0);
    if (take_flag(F_STATUS)) return TCS_BAD_GRAPH;
    ex.phase = 16;
    // label tables: only out-degree-0 nodes carry labels
    PT_FOR(v, n) {
      const int l = w.hlab[v];
      if (l >= 0) {
        if (w.ooff[v + 1] == w.ooff[v]) { if (ex.atomic_cas_idx(&w.l2h[l], -1, v) != -1) w.sc[F_STATUS] = TCS_MULTILABEL; }
        else w.hlab[v] = (I)-1;
      }
    }
    PT_FOR(t, nt) {
      const int l = w.tlab[t];
      if (l >= 0) {
        if (w.tdeg[t] == 0) { if (ex.atomic_cas_idx(&w.l2t[l], -1, t) != -1) w.sc[F_STATUS] = TCS_MULTILABEL; }
        else w.tlab[t] = (I)-1;
      }
    }
    if (take_flag(F_STATUS)) return TCS_MULTILABEL;
    // containment.hpp:1110-1125: label only in the guest -> not displayed; only in the host -> leaf loses its label
    PT_FOR(l, L) {
      const int h = w.l2h[l], t = w.l2t[l];
      if (t >= 0 && h < 0) w.sc[F_FAIL] = 1;
      if (h >= 0 && t < 0) { w.hlab[h] = (I)-1; w.l2h[l] = (I)-1; }
    }
    ex.sync();
    return TCS_OK;
  }

  // ---- guest clean-up: drop unlabelled leaves, suppress unary nodes --------------------------------------
  PT_NI void clean_guest() {
    ex.phase = 2;
    for (;;) {
      PT_FOR(t, nt) {
        if (w.tpar[t] != TDEAD && w.tdeg[t] == 0 && w.tlab[t] < 0) {
          const int p = w.tpar[t];
          w.tpar[t] = (I)TDEAD;
          if (p >= 0) ex.atomic_add_idx(&w.tdeg[p], -1);
          w.sc[F_FLAG] = 1;
        }
      }
      if (!take_flag(F_FLAG)) break;
    }
    // new parent = nearest proper ancestor with >= 2 live children (haux is the double buffer)
    PT_FOR(t, nt) {
      int np = TDEAD;
      if (w.tpar[t] != TDEAD && w.tdeg[t] != 1) {
        int p = w.tpar[t];
        while (p >= 0 && w.tdeg[p] == 1) p = w.tpar[p];
        np = p;
      }
      w.haux[t] = (I)np;
    }
    ex.sync();
    PT_FOR(t, nt) { w.tpar[t] = w.haux[t]; if (w.haux[t] == -1) w.sc[F_TROOT] = t; w.tmax[t] = (I)0; }
    ex.sync();
    // tmax[p] = XOR of the ids of p's children: the sibling of t below a binary p is tmax[p] ^ t.  Child SETS never change
    // afterwards (a collapse turns p into a leaf as a whole), so this is computed once per work item.
    PT_FOR(t, nt) { if (w.tpar[t] >= 0) ex.atomic_xor_idx(&w.tmax[w.tpar[t]], t); }
    ex.sync();
  }
  // label of the other leaf of a binary guest cherry containing the leaf labelled l, else -1
  PT_HD int cherry_mate(int l) const {
    const int t = w.l2t[l];
    if (t < 0) return -1;
    const int p = w.tpar[t];
    if (p < 0 || w.tdeg[p] != 2) return -1;
    return w.tlab[w.tmax[p] ^ t];
  }

  // ---- host CSR over live arcs ---------------------------------------------------------------------------
  // Arcs are sorted by tail from the input on and no rule ever changes a tail, so the out-CSR is a stable compaction of the
  // arc arrays (ballot / popc, no atomics) plus the positions where the tail changes; only the in-CSR needs counting.
  PT_NI void csr() {
    ex.phase = 3;
    if (ex.tid() == 0) w.sc[F_QTAIL] = 0;
    m = ex.compact_arcs(w.at, w.ah, m, w.iarc, w.atmp);  // drops dead arcs (ends with a barrier); iarc is rebuilt below
    ex.phase = 31;
    // The instance shrinks quickly (on BASELINE config 2 the average snapshot has 67 of 228 nodes alive).  Once a quarter of
    // the node ids is dead, renumber the live ones so that every later pass is proportional to what is left.
    if (4 * (m + 1) < 3 * n) {
      const int root = ex.uniform(&w.sc[F_ROOT]);
      PT_FOR(v, n + 1) { w.hcnt[v] = (I)((v < n && (w.hlab[v] >= 0 || v == root)) ? 1 : 0); }
      ex.sync();
      PT_FOR(e, m) { w.hcnt[w.at[e]] = (I)1; w.hcnt[w.ah[e]] = (I)1; }
      ex.sync();
      ex.exclusive_scan(w.hcnt, w.hscr, n);  // hscr[v] = new id of v
      PT_FOR(e, m) { w.at[e] = w.hscr[w.at[e]]; w.ah[e] = w.hscr[w.ah[e]]; }
      PT_FOR(v, n) { if (w.hcnt[v]) w.lvl[w.hscr[v]] = w.hlab[v]; }  // lvl is the second buffer of hlab
      PT_FOR(l, L) { const int h = w.l2h[l]; if (h >= 0) w.l2h[l] = w.hscr[h]; }
      if (ex.tid() == 0) w.sc[F_ROOT] = w.hscr[root];
      const int n2 = w.hscr[n];
      ex.sync();
      I* const t = w.hlab; w.hlab = w.lvl; w.lvl = t;
      n = n2;
    }
    ex.phase = 32;
    PT_FOR(v, n + 1) { w.haux[v] = (I)0; }
    ex.sync();
    PT_FOR(e, m) { ex.atomic_add_idx(&w.haux[w.ah[e]], 1); }
    PT_FOR(e, m + 1) {
      const int lo = e > 0 ? w.at[e - 1] + 1 : 0, hi = e < m ? (int)w.at[e] : n;
      PT_NOUNROLL for (int v = lo; v <= hi; ++v) w.ooff[v] = (I)e;
    }
    ex.phase = 33;
    ex.exclusive_scan(w.haux, w.ioff, n);
    ex.phase = 34;
    PT_FOR(e, m) { const int h = w.ah[e]; w.iarc[w.ioff[h] + ex.atomic_add_idx(&w.haux[h], -1) - 1] = (I)e; }
    ex.sync();
#ifdef PT_SIM_STATS
    { int live = 0; for (int v = 0; v < n; ++v) live += (indeg(v) > 0 || outdeg(v) > 0); g_sim_csr_calls++; g_sim_live_nodes += live; g_sim_nodes += n; g_sim_live_arcs += m; }
#endif
  }

  PT_HD bool suppressible(int v) const { return indeg(v) == 1 && outdeg(v) == 1; }
  // multi-parent node with a single child that is itself multi-parent: its parents can be handed down
  PT_HD bool mergeable(int v) const { return indeg(v) >= 2 && outdeg(v) == 1 && indeg(w.ah[out_arc(v, 0)]) >= 2; }

  // ---- host clean-up to a fix-point (sets F_FAIL when a label became unreachable) ------------------------------
  // One detection pass per CSR snapshot decides which rewriting pass is needed (usually none); only one kind of
  // rewriting runs per snapshot, then the CSR is rebuilt.
  enum { NEED_DEAD = 1, NEED_SUPPRESS = 2, NEED_MERGE = 4, NEED_PARALLEL = 8, NEED_FAIL = 16 };
  // dag_state: 0 = input not validated yet (untrusted), 1 = validated, 2 = trusted (only the root is looked up)
  PT_NI void clean_host(int& dag_state, int* status) {
    for (;;) {
      csr();
      ex.phase = 40;
      if (dag_state != 1) {
        const bool ok = check_dag(dag_state == 2);
        dag_state = 1;
        if (!ok) { *status = TCS_BAD_GRAPH; return; }
        ex.phase = 40;
      }
      const int root = ex.uniform(&w.sc[F_ROOT]);
      int need = 0;
      PT_FOR(v, n) {
        const int od = outdeg(v), id = indeg(v);
        int cnt = 0;
        if (od == 0) {
          if (w.hlab[v] < 0) { if (id > 0) need |= NEED_DEAD; }          // (a) unlabelled dead end
          else if (id == 0 && v != root) need |= NEED_FAIL;              // a labelled leaf lost every in-arc
        } else {
          if (id == 0 && v != root) need |= NEED_DEAD;                   // (a') orphan: unreachable, drop its out-arcs
          if (od == 1) {
            if (id == 1) need |= NEED_SUPPRESS;                          // (c)
            else if (id >= 2 && indeg(w.ah[out_arc(v, 0)]) >= 2) need |= NEED_MERGE;  // (d)
          } else {
            if (od <= 8) {                                               // (e) parallel arcs
              PT_NOUNROLL for (int a = 1; a < od; ++a)
                PT_NOUNROLL for (int b = 0; b < a; ++b)
                  if (w.ah[out_arc(v, a)] == w.ah[out_arc(v, b)]) need |= NEED_PARALLEL;
            }
            // hcnt[v] = children that are labelled leaves with a single parent; all of them -> true-cherry candidate of this
            // snapshot (only used when the snapshot turns out to be clean)
            PT_NOUNROLL for (int k = 0; k < od; ++k) { const int c = w.ah[out_arc(v, k)]; cnt += (w.hlab[c] >= 0 && indeg(c) == 1); }
            if (cnt == od) w.hscr[ex.atomic_add(&w.sc[F_QTAIL], 1)] = (I)v;
          }
        }
        w.hcnt[v] = (I)cnt;
      }
      need = ex.reduce_or(need);
      if (need & NEED_FAIL) { set_scalar(F_FAIL, 1); return; }
      if (need & NEED_DEAD) {
        ex.phase = 41;
        PT_FOR(v, n) {
          if (outdeg(v) == 0) { if (w.hlab[v] < 0) PT_NOUNROLL for (int k = 0; k < indeg(v); ++k) w.at[in_arc(v, k)] = (I)-1; }
          else if (indeg(v) == 0 && v != root) PT_NOUNROLL for (int k = 0; k < outdeg(v); ++k) w.at[out_arc(v, k)] = (I)-1;
        }
        ex.sync();
        continue;
      }
      // (b) out-degree-1 root: its child becomes the root
      if (outdeg(root) == 1) {
        ex.sync();
        if (ex.tid() == 0) { const int e = out_arc(root, 0); w.at[e] = (I)-1; w.sc[F_ROOT] = w.ah[e]; }
        ex.sync();
        continue;
      }
      if (need & NEED_SUPPRESS) {
        ex.phase = 43;
        // every arc decides for itself, so the pass is race-free: an arc whose tail is suppressible disappears; an arc
        // entering a chain from a non-suppressible tail jumps to the end of the chain
        PT_FOR(e, m) {
          const int t = w.at[e];
          if (t >= 0) {
            if (suppressible(t)) w.at[e] = (I)-1;
            else {
              int h = w.ah[e];
              if (suppressible(h)) { int guard = 0; while (suppressible(h) && ++guard <= n) h = w.ah[out_arc(h, 0)]; w.ah[e] = (I)h; if (guard > n) w.sc[F_STATUS] = TCS_BAD_GRAPH; }
            }
          }
        }
        if (take_flag(F_STATUS)) { *status = TCS_BAD_GRAPH; return; }  // a chain longer than n: cyclic structure
        continue;
      }
      if (need & NEED_MERGE) {
        ex.phase = 44;
        // reticulation chains: hand the parents of x down to the bottom of the chain (cf. ReticulationMerger :392-409)
        PT_FOR(v, n) {
          int y = -1;
          if (mergeable(v)) { int guard = 0; y = w.ah[out_arc(v, 0)]; while (mergeable(y) && ++guard <= n) y = w.ah[out_arc(y, 0)]; if (guard > n) w.sc[F_STATUS] = TCS_BAD_GRAPH; }
          w.haux[v] = (I)y;
        }
        if (take_flag(F_STATUS)) { *status = TCS_BAD_GRAPH; return; }
        PT_FOR(v, n) {
          if (w.haux[v] >= 0) { PT_NOUNROLL for (int k = 0; k < indeg(v); ++k) w.ah[in_arc(v, k)] = (I)w.haux[v]; w.at[out_arc(v, 0)] = (I)-1; }
        }
        ex.sync();
        continue;
      }
      if (need & NEED_PARALLEL) {
        ex.phase = 46;
        PT_FOR(v, n) {
          const int d = outdeg(v);
          if (d >= 2 && d <= 8)
            PT_NOUNROLL for (int a = 0; a < d; ++a)
              PT_NOUNROLL for (int b = 0; b < d; ++b) { const int ea = out_arc(v, a), eb = out_arc(v, b); if (ea < eb && w.ah[ea] == w.ah[eb]) w.at[eb] = (I)-1; }
        }
        ex.sync();
        continue;
      }
      return;
    }
  }

  PT_HD int count_labels() {
    int c = 0;
    PT_FOR(l, L) { c += (w.l2h[l] >= 0); }
    return ex.reduce_add(c);
  }

  // ---- wavefront over topological levels: heights and reachable-label bitsets -----------------------------
  // returns false if some live node was never reached (cycle)
  PT_NI bool wavefront() {
    ex.phase = 6;
    const int W = w.W;
    // Kahn order from the leaves with a FIFO in hscr: total work O(n + m) instead of one scan of all nodes per level.
    // Nodes enter the queue in non-decreasing height, so the child that releases a parent is one of its tallest children.
    set_scalar(F_TMP, 0);  // queue tail
    PT_FOR(v, n) {
      PT_NOUNROLL for (int k = 0; k < W; ++k) w.lset[v * W + k] = 0;
      w.hcnt[v] = (I)outdeg(v);
      w.lvl[v] = (I)LVL_INF;
      if (w.hlab[v] >= 0) { const int b = w.hlab[v] % (W * 32); w.lset[v * W + (b >> 5)] = 1u << (b & 31); }
      if (outdeg(v) == 0 && (indeg(v) > 0 || w.hlab[v] >= 0)) { w.lvl[v] = (I)0; w.hscr[ex.atomic_add(&w.sc[F_TMP], 1)] = (I)v; }
    }
    ex.sync();
    int head = 0;
    for (;;) {
      const int tail = ex.uniform(&w.sc[F_TMP]);  // (ends with a barrier: everybody has the tail before anybody appends)
      if (head >= tail) break;
      const int cnt = tail - head;
      PT_FOR(i, cnt) {
        const int v = w.hscr[head + i], lv = w.lvl[v];
        PT_NOUNROLL for (int k = 0; k < indeg(v); ++k) {
          const int p = w.at[in_arc(v, k)];
          PT_NOUNROLL for (int j = 0; j < W; ++j) { const uint32_t b = w.lset[v * W + j]; if (b) ex.atomic_or(&w.lset[p * W + j], b); }
          if (ex.atomic_add_idx(&w.hcnt[p], -1) == 1) { w.lvl[p] = (I)(lv + 1); w.hscr[ex.atomic_add(&w.sc[F_TMP], 1)] = (I)p; }
        }
      }
      head = tail;
      ex.sync();
    }
    set_scalar(F_TMP, 0);
    int bad = 0;
    PT_FOR(v, n) { if (outdeg(v) > 0 && w.lvl[v] == LVL_INF) bad = 1; }
    return !ex.reduce_or(bad);
  }

  // ---- TC: true (multi-)cherries, as a worklist --------------------------------------------------------------
  // A host node q all of whose (>= 2) children are labelled leaves with a single parent must equal the guest node whose
  // children are exactly those leaves: collapse both into one leaf (it keeps the smallest label), else NOT displayed
  // (cf. CherryPicker :538-590, match_nodes :915-950).  hcnt[v] counts the children of v that are such leaves (detection
  // pass of clean_host).  The queue hscr[head, F_QTAIL) holds the nodes whose count is complete; collapsing q tells q's
  // parent, and the thread that completes the parent's count queues it for the next round, so a pendant subtree that
  // matches collapses level by level and the total work is proportional to the cherries found.
  // Tolerates a CSR that is stale by arc DELETIONS: dead arcs have at = -1, degrees only over-count, so a count is then
  // never completed and the cherry is found after the next clean-up instead.
  PT_HD void notify_parent(int q) {
    if (indeg(q) != 1) return;
    const int g = w.at[in_arc(q, 0)];
    if (g >= 0 && ex.atomic_add_idx(&w.hcnt[g], 1) + 1 == outdeg(g)) w.hscr[ex.atomic_add(&w.sc[F_QTAIL], 1)] = (I)g;
  }
  // returns 1 something collapsed, 0 nothing, -1 NOT displayed
  PT_NI int rule_true_cherry(int& head) {
    ex.phase = 7;
    int collapsed = 0, rounds = 0;
    for (;;) {
      const int tail = ex.uniform(&w.sc[F_QTAIL]);
      if (head >= tail) break;
      const int cnt = tail - head;
      ++rounds;
      PT_FOR(i, cnt) {
        const int q = w.hscr[head + i];
        int p = -1, num = 0, minl = 0x7fffffff;
        bool ok = true;
        PT_NOUNROLL for (int k = 0; k < outdeg(q); ++k) {
          const int e = out_arc(q, k);
          if (w.at[e] < 0) continue;
          const int l = w.hlab[w.ah[e]];
          if (l < 0) { num = -1000000; continue; }  // (cannot happen for a complete count; never collapse on a doubt)
          const int pp = w.tpar[w.l2t[l]];
          if (num == 0) p = pp; else if (pp != p) ok = false;
          ++num;
          if (l < minl) minl = l;
        }
        if (num < 2) continue;
        if (!ok || p < 0 || w.tdeg[p] != num) { w.sc[F_FAIL] = 1; continue; }
        PT_NOUNROLL for (int k = 0; k < outdeg(q); ++k) {
          const int e = out_arc(q, k);
          if (w.at[e] < 0) continue;
          const int c = w.ah[e], l = w.hlab[c], t = w.l2t[l];
          w.at[e] = (I)-1; w.hlab[c] = (I)-1;
          w.tlab[t] = (I)-1; w.tpar[t] = (I)TDEAD;
          if (l != minl) { w.l2h[l] = (I)-1; w.l2t[l] = (I)-1; }
        }
        w.hlab[q] = (I)minl; w.l2h[minl] = (I)q;
        w.tlab[p] = (I)minl; w.l2t[minl] = (I)p; w.tdeg[p] = (I)0;
        ++collapsed;
        notify_parent(q);
      }
      head = tail;
    }
    if (!rounds) return 0;
    if (ex.uniform(&w.sc[F_FAIL])) return -1;
    st.cherries += (unsigned)ex.reduce_add(collapsed);
    return 1;
  }

  // ---- RC: reticulated cherries ---------------------------------------------------------------------------
  // One thread per label x whose guest leaf sits in a binary cherry {x, y}: host leaf x hangs on q, and y (or the
  // out-degree-1 reticulation above y) has q among its parents -> keep that in-arc, delete the others.  (The mirrored case
  // is handled by the thread of y; both cannot apply at once: it would need q above z and z above q.)
  // returns bit 0: some multi-parent node got its parent fixed (needs a clean-up pass), bit 1: some reticulated cherry was
  // collapsed on the spot (q had exactly the two children; q's parent is told like after a true cherry).
  // Tolerates a CSR that is stale by arc deletions: a dead arc has at = -1 and never matches q, degrees only over-count.
  PT_NI int rule_ret_cherry() {
    ex.phase = 9;
    int fixed = 0, collapsed = 0;
    PT_FOR(x, L) {
      const int hx = w.l2h[x];
      if (hx < 0 || indeg(hx) != 1) continue;
      const int y = cherry_mate(x);
      if (y < 0) continue;
      const int hy = w.l2h[y];
      const int ex_arc = in_arc(hx, 0);
      const int q = w.at[ex_arc];
      if (q < 0 || hy < 0) continue;
      int c = -1;  // the multi-parent node whose parent gets fixed to q
      if (indeg(hy) >= 2) c = hy;
      else if (indeg(hy) == 1) { const int z = w.at[in_arc(hy, 0)]; if (z >= 0 && outdeg(z) == 1 && indeg(z) >= 2) c = z; }
      if (c < 0) continue;
      int keep = -1;
      PT_NOUNROLL for (int k = 0; k < indeg(c); ++k) if (w.at[in_arc(c, k)] == q) keep = in_arc(c, k);
      if (keep < 0) continue;
      PT_NOUNROLL for (int k = 0; k < indeg(c); ++k) { const int e = in_arc(c, k); if (e != keep) w.at[e] = (I)-1; }
      bool both = false;  // q's children are exactly hx and c (both arcs alive)
      if (outdeg(q) == 2) {
        const int e0 = out_arc(q, 0), e1 = out_arc(q, 1);
        both = (e0 == ex_arc && e1 == keep) || (e1 == ex_arc && e0 == keep);
      }
      if (both) {
        // collapse q -> leaf carrying the smaller label; the guest cherry -> leaf carrying the same label
        const int s_lab = x < y ? x : y, t_lab = x < y ? y : x;
        const int ts = w.l2t[s_lab], tt = w.l2t[t_lab], p = w.tpar[ts];
        w.at[ex_arc] = (I)-1; w.at[keep] = (I)-1;
        if (c != hy) w.at[in_arc(hy, 0)] = (I)-1;
        w.hlab[hx] = (I)-1; w.hlab[hy] = (I)-1;
        w.tlab[ts] = (I)-1; w.tpar[ts] = (I)TDEAD; w.tlab[tt] = (I)-1; w.tpar[tt] = (I)TDEAD;
        w.hlab[q] = (I)s_lab; w.l2h[s_lab] = (I)q; w.l2h[t_lab] = (I)-1;
        w.tlab[p] = (I)s_lab; w.l2t[s_lab] = (I)p; w.l2t[t_lab] = (I)-1; w.tdeg[p] = (I)0;
        notify_parent(q);
        ++collapsed;
      } else ++fixed;
    }
    const int nc = ex.reduce_add(collapsed), nf = ex.reduce_add(fixed);
    st.retcherries += (unsigned)(nc + nf);
    st.cherries += (unsigned)nc;
    return (nf ? 1 : 0) | (nc ? 2 : 0);
  }

  // ---- SC: sibling constraints ----------------------------------------------------------------------------
  // haux[v] = constraint of single-parent node v: -1 none, lam >= 0 "live leaves below v are a subset of {lam}",
  // -2 "nothing below v may stay alive".  hcnt[v] = 1 once v has been expanded.
  PT_HD void constrain_arc(int e, int lam) {
    const int c = w.ah[e];
    const bool reach = lam >= 0 && lbit(c, lam);
    if (indeg(c) >= 2) {
      if (!reach) { w.at[e] = (I)-1; ex.atomic_add(&w.sc[F_TMP2], 1); w.sc[F_FLAG] = 1; }
      return;
    }
    if (outdeg(c) == 0) { if (!reach) w.sc[F_FAIL] = 1; return; }  // a labelled single-parent leaf that must die
    const int want = reach ? lam : -2;
    if (w.haux[c] != want) { w.haux[c] = (I)want; w.sc[F_TMP] = 1; }
  }
  PT_NI int rule_sibling() {
    ex.phase = 10;
    PT_FOR(v, n + 1) { w.haux[v] = (I)-1; w.hcnt[v] = (I)0; }
    ex.sync();
    PT_FOR(x, L) {
      const int hx = w.l2h[x];
      if (hx < 0 || indeg(hx) != 1) continue;
      const int y = cherry_mate(x);
      if (y < 0) continue;
      const int q = w.at[in_arc(hx, 0)];
      PT_NOUNROLL for (int k = 0; k < outdeg(q); ++k) { const int e = out_arc(q, k); if (w.ah[e] != hx) constrain_arc(e, y); }
    }
    // propagate through single-parent nodes
    for (;;) {
      if (!take_flag(F_TMP)) break;
      PT_FOR(v, n) {
        if (w.haux[v] != -1 && !w.hcnt[v]) {
          w.hcnt[v] = (I)1;
          PT_NOUNROLL for (int k = 0; k < outdeg(v); ++k) constrain_arc(out_arc(v, k), w.haux[v]);
        }
      }
    }
    if (ex.uniform(&w.sc[F_FAIL])) return -1;
    st.arcs_deleted += (unsigned)w.sc[F_TMP2];
    ex.sync();
    set_scalar(F_TMP2, 0);
    return take_flag(F_FLAG) ? 1 : 0;
  }

  // ---- branching: lowest multi-parent node ------------------------------------------------------------------
  PT_NI int pick_branch_node() {
    ex.phase = 11;
    set_scalar(F_BEST, 0x7fffffff);
    PT_FOR(v, n) { if (indeg(v) >= 2) ex.atomic_min(&w.sc[F_BEST], (int)w.lvl[v]); }
    const int lo = ex.uniform(&w.sc[F_BEST]);
    if (lo == 0x7fffffff) return -1;
    set_scalar(F_BEST, 0x7fffffff);
    PT_FOR(v, n) { if (indeg(v) >= 2 && (int)w.lvl[v] == lo) ex.atomic_min(&w.sc[F_BEST], v); }
    const int b = ex.uniform(&w.sc[F_BEST]);
    return b;
  }

  // write the current (clean) state, with only in-arc `keep` of node x kept, as a compact instance.
  // Label ids are kept (no renumbering), so L and the bitset width stay those of the parent.
  template <class S>
  PT_NI void emit_child(int x, int keep, const TcEmitT<S>& out) {
    typedef TcLayout<S> Lay;
    ex.phase = 12;
    const int root = ex.uniform(&w.sc[F_ROOT]);
    // live host nodes (root, or any live in-arc) -> new ids by a scan of the flags (lvl is free after pick_branch_node)
    PT_FOR(v, n + 1) {
      const int live = (v < n && (v == root || indeg(v) > 0)) ? 1 : 0;
      int d = 0;
      if (live) PT_NOUNROLL for (int k = 0; k < outdeg(v); ++k) { const int e = out_arc(v, k); if (w.ah[e] != x || e == keep) ++d; }
      w.hcnt[v] = (I)live; w.haux[v] = (I)d;
    }
    ex.sync();
    ex.exclusive_scan(w.hcnt, w.lvl, n);   // lvl[v] = new id of v, lvl[n] = n2
    ex.exclusive_scan(w.haux, w.hscr, n);  // hscr[v] = first arc position of v in the child
    const int n2 = w.lvl[n], m2 = w.hscr[n];
    PT_FOR(v, n) {
      if (w.hcnt[v]) {
        const int id = w.lvl[v];
        out.succ_off[id] = Lay::kDegrees ? (S)w.haux[v] : (S)w.hscr[v];
        out.hlabel[id] = Lay::enc(w.hlab[v]);
        int pos = w.hscr[v];
        PT_NOUNROLL for (int k = 0; k < outdeg(v); ++k) { const int e = out_arc(v, k); if (w.ah[e] != x || e == keep) out.succ_adj[pos++] = (S)w.lvl[w.ah[e]]; }
      }
    }
    // guest
    ex.sync();  // the host part is done with hcnt / haux
    PT_FOR(t, nt + 1) { w.hcnt[t] = (I)((t < nt && w.tpar[t] != TDEAD) ? 1 : 0); }
    ex.sync();
    ex.exclusive_scan(w.hcnt, w.haux, nt);
    const int nt2 = w.haux[nt];
    PT_FOR(t, nt) {
      if (w.hcnt[t]) { const int id = w.haux[t]; out.tparent[id] = Lay::enc(w.tpar[t] >= 0 ? (int)w.haux[w.tpar[t]] : -1); out.tlabel[id] = Lay::enc(w.tlab[t]); }
    }
    if (ex.tid() == 0) { if (!Lay::kDegrees) out.succ_off[n2] = (S)m2; out.dims[0] = n2; out.dims[1] = m2; out.dims[2] = nt2; out.dims[3] = L; }
    ex.sync();
  }

  // Kahn peel from the leaves: every node must be reached, exactly one node has no in-arc (the reference throws
  // std::logic_error for several roots, utils/storage_adj_common.hpp:103-109; a cycle would hang its DFS)
  PT_NI bool check_dag(bool trusted) {
    ex.phase = 5;
    int roots = 0;
    set_scalar(F_TMP, 0);
    PT_FOR(v, n) {
      w.hcnt[v] = (I)outdeg(v);
      if (indeg(v) == 0) { ++roots; w.sc[F_ROOT] = v; }
      if (!trusted && outdeg(v) == 0) w.hscr[ex.atomic_add(&w.sc[F_TMP], 1)] = (I)v;
    }
    const int nroots = ex.reduce_add(roots);
    if (n > 0 && nroots != 1) { set_scalar(F_TMP, 0); return false; }
    if (trusted) { set_scalar(F_TMP, 0); return true; }
    int head = 0;
    for (;;) {
      const int tail = ex.uniform(&w.sc[F_TMP]);
      if (head >= tail) break;
      const int cnt = tail - head;
      PT_FOR(i, cnt) {
        const int v = w.hscr[head + i];
        PT_NOUNROLL for (int k = 0; k < indeg(v); ++k) { const int p = w.at[in_arc(v, k)]; if (ex.atomic_add_idx(&w.hcnt[p], -1) == 1) w.hscr[ex.atomic_add(&w.sc[F_TMP], 1)] = (I)p; }
      }
      head = tail;
      ex.sync();
    }
    set_scalar(F_TMP, 0);
    return head == n;  // every node was released exactly once <=> no cycle
  }

  // ---- the small core: when the rules are stuck, what is left is tiny ------------------------------------------------------
  // Measured on BASELINE config 2 (l = 100, r = 20): an instance that needs branching has been reduced to 9 ... 40 live host nodes
  // with 2 ... 5 multi-parent nodes by then.  Instead of handing its children to other warps (emit, reload, a full rule loop per
  // child: ~100 us of latency per level, the tail of every launch), the group decides it on the spot from the definition: ONE
  // SWITCHING PER THREAD -- every thread walks the live nodes children first (the Kahn order the wavefront has just left in hscr),
  // forms the label set below every node as a 32-bit mask under its switching, and checks that the nodes with >= 2 live children
  // carry exactly the guest's clusters.  Exact (sets, not hashes) because at most 32 labels are alive.  Applicable when
  // <= kCoreNodes (64) live nodes, <= 32 live labels and <= kCoreSwitchings (64) switchings remain; otherwise -1 and the caller branches.
  // Needs: fresh CSR, wavefront order in hscr (both hold when rule_loop has just returned TC_BRANCH: the caller tries this first and
  // branches only on -1).  Touches scratch only.
  // (kCoreSwitchings: 89 % of the stuck states of config 2 have <= 32 switchings, 95 % <= 64; every 32 more are one more pass of
  // ~20 us on this one warp, and a 16-pass straggler at the end of a launch costs more than a level of branching)
  static constexpr int kCoreNodes = 64, kCoreSwitchings = 64;
  int core_switchings_;   // > 0: the last verdict came from the small core, over this many switchings
  bool allow_core_;       // false: always branch (pt_options.no_small_core; keeps the branching machinery testable)
  PT_NI int finish_small_core() {
    if (!allow_core_) return -1;
    ex.phase = 17;
    // Set-up by ONE thread, on purpose: a dozen small parallel passes with their barriers were 10 KB of code, and this kernel pays
    // for every KB with all its instances (instruction cache).  The loops are short: <= 64 nodes, <= 32 labels with their ancestors.
    if (ex.tid() == 0) {
      int K = 0, nlab = 0, P = 1, Gi = 0, ok = 1;
      PT_NOUNROLL for (int v = 0; v < n; ++v) K += w.lvl[v] != LVL_INF;
      if (K > kCoreNodes || K < 1) ok = 0;
      PT_NOUNROLL for (int i = 0; i < K && ok; ++i) {   // position in the order; leaves get a bit, multi-parent nodes a digit
        const int v = w.hscr[i], d = indeg(v);
        w.hcnt[v] = (I)i;
        const bool leaf = outdeg(v) == 0 && w.hlab[v] >= 0;
        w.haux[v] = (I)(leaf ? nlab++ : -1);
        if (d >= 2) { w.lvl[v] = (I)P; P *= d; if (P > kCoreSwitchings) ok = 0; } else w.lvl[v] = (I)0;
        if (leaf) {   // clear the cluster words of the guest leaf's ancestors (first of two walks)
          const int t = w.l2t[w.hlab[v]];
          if (t < 0 || w.tpar[t] == TDEAD) { ok = 0; continue; }
          int guard = 0;
          PT_NOUNROLL for (int p = w.tpar[t]; p >= 0 && guard++ < nt; p = w.tpar[p]) { if (p >= w.capN * w.W) { ok = 0; break; } w.lset[p] = 0u; }
        }
      }
      if (nlab > 32 || nlab < 1) ok = 0;
      PT_NOUNROLL for (int i = 0; i < K && ok; ++i) {   // guest clusters; an inner node is listed (behind the order in hscr) when first met
        const int v = w.hscr[i];
        if (w.haux[v] < 0) continue;
        const uint32_t bit = 1u << w.haux[v];
        int guard = 0;
        PT_NOUNROLL for (int p = w.tpar[w.l2t[w.hlab[v]]]; p >= 0 && guard++ < nt; p = w.tpar[p]) {
          if (w.lset[p] == 0u) {
            if (w.tdeg[p] < 2 || K + Gi > w.capN || Gi >= 31) { ok = 0; break; }   // a unary guest node: not clean, leave it to the rules
            w.hscr[K + Gi++] = (I)p;
          }
          w.lset[p] |= bit;
        }
      }
      w.sc[F_TMP] = ok ? (K | (nlab << 8) | (Gi << 16)) : -1;
      w.sc[F_TMP2] = P;
    }
    ex.sync();
    const int packed = ex.uniform(&w.sc[F_TMP]), P = ex.uniform(&w.sc[F_TMP2]);
    if (ex.tid() == 0) { w.sc[F_TMP] = 0; w.sc[F_TMP2] = 0; }
    ex.sync();
    if (packed < 0) return -1;
    const int K = packed & 255, nlab = (packed >> 8) & 255, Gi = packed >> 16;
    // one switching per thread
    ex.phase = 18;
    const uint32_t all = nlab == 32 ? 0xffffffffu : ((1u << nlab) - 1u);
    uint32_t cl[kCoreNodes];
    int found = 0;
    PT_NOUNROLL for (int base = 0; base < P && !found; base += ex.nth()) {
      const int sidx = base + ex.tid();
      int ok = sidx < P ? 1 : 0, real = 0;
      PT_NOUNROLL for (int i = 0; i < K && ok; ++i) {
        const int v = w.hscr[i], od = outdeg(v);
        uint32_t mask = w.haux[v] >= 0 ? (1u << w.haux[v]) : 0u;
        int kids = 0;
        PT_NOUNROLL for (int k = 0; k < od; ++k) {
          const int e = out_arc(v, k), c = w.ah[e], dc = indeg(c);
          if (dc >= 2 && in_arc(c, (sidx / (int)w.lvl[c]) % dc) != e) continue;
          const uint32_t mc = cl[w.hcnt[c]];
          if (mc) { mask |= mc; ++kids; }
        }
        cl[i] = mask;
        if (kids >= 2) {
          ++real;
          int hit = 0;
          PT_NOUNROLL for (int j = 0; j < Gi; ++j) hit |= (w.lset[w.hscr[K + j]] == mask);
          ok = hit;
        }
      }
      if (ok && (real != Gi || cl[K - 1] != all)) ok = 0;
      found = ex.reduce_or(ok);
    }
    core_switchings_ = P;
    return found ? 1 : 0;
  }

  // ---- branching in place: keep only in-arc `keep` of node x and go on with the same workspace (first child of a branching;
  // the siblings are emitted as compact instances before this is called) ---------------------------------------------------
  PT_NI void keep_only(int x, int keep) {
    PT_FOR(k, indeg(x)) { const int e = in_arc(x, k); if (e != keep) w.at[e] = (I)-1; }
    ex.sync();
  }

  // ---- the rule loop (apply_rules, containment.hpp:1054-1083) -----------------------------------------------
  // begin(): load + validate + guest clean-up; returns TC_CONTINUE (go on with rule_loop), or TC_ERROR with *status set.
  // rule_loop(): returns TC_NOT / TC_DISPLAYED / TC_ERROR (*status holds the TCS_*) / TC_BRANCH (*branch_node = the node to
  // branch on; the caller emits the children and / or calls keep_only and rule_loop again).
  int missing_;    // a guest label without a host leaf; reported after the input has been validated
  int dag_state_;  // see clean_host
  template <class S>
  PT_HD int begin(const TcInst& in, bool trusted, int* status) {
    *status = TCS_OK;
    int s = load<S>(in);
    if (s != TCS_OK) { *status = s; return TC_ERROR; }
    missing_ = ex.uniform(&w.sc[F_FAIL]);
    clean_guest();
    dag_state_ = trusted ? 2 : 0;
    return TC_CONTINUE;
  }
  PT_HD int rule_loop(int* status, int* branch_node) {
    *branch_node = -1;
    for (;;) {
      ++st.passes;
      clean_host(dag_state_, status);
      if (*status != TCS_OK) return TC_ERROR;
      if (missing_ || ex.uniform(&w.sc[F_FAIL])) return TC_NOT;
      ex.phase = 14;
      if (count_labels() <= 2) return TC_DISPLAYED;  // containment.hpp:1061,1208
      // true cherries and on-the-spot reticulated cherries alternate without rebuilding the CSR: pendant subtrees that match
      // collapse level by level; only a fixed-but-not-collapsed reticulation (or nothing at all) sends us back to clean-up
      int r, progress = 0, head = 0;
      for (;;) {
        r = rule_true_cherry(head);
        if (r < 0) break;
        if (r > 0) progress = 1;
        r = rule_ret_cherry();
        if (r) progress = 1;
        if (!(r & 2)) break;
      }
      if (r < 0) return TC_NOT;
      if (progress) continue;
#ifdef PT_AONLY   // measuring aid (make aonly): how fast is the kernel when the code behind this point does not exist?  (wrong verdicts)
      *branch_node = 0; return TC_BRANCH;
#endif
      if (!wavefront()) { *status = TCS_BAD_GRAPH; return TC_ERROR; }
      r = rule_sibling();
      if (r < 0) return TC_NOT;
      if (r > 0) continue;
      const int x = pick_branch_node();
      if (x < 0) { *status = TCS_BAD_GRAPH; return TC_ERROR; }  // a clean tree host always has a true cherry
      // (the CALLER tries finish_small_core() before it branches: inlined here, inside the loop nest, the same code made every
      // instance 7 % slower -- 7.87 -> 8.45 ms per 10^5 -- although 92 % of them never reach it, and as an out-of-line call 35 %
      // slower: the registers live across the call get spilled all over the loop)
      *branch_node = x;
      return TC_BRANCH;
    }
  }
  template <class S>
  PT_HD int solve(const TcInst& in, bool trusted, int* status, int* branch_node) {
    *branch_node = -1;
    int c = begin<S>(in, trusted, status);
    if (c != TC_CONTINUE) return c;
    c = rule_loop(status, branch_node);
    if (c == TC_BRANCH) { const int core = finish_small_core(); if (core >= 0) c = core ? TC_DISPLAYED : TC_NOT; }
    return c;
  }
};

}  // namespace ptgpu
