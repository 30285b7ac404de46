/* oracle/pt_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * Plain-C CPU restatement of the reference's hot path, used exclusively as the CHECKER by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg.  Nothing in phylo_tools_b200/ (the product)
 * links, imports or calls this file.
 *
 * What is restated, and from where (file:line relative to the reference checkout):
 *   orc_tc_bruteforce   -- the containment SEMANTICS that utils/containment.hpp:955-957 (TreeInNetContainment)
 *                          decides: "N displays T iff some switching of N, cleaned up, equals T".  Cleaning =
 *                          NodeSuppresser::clean_up_node containment.hpp:756-824 (drop unlabelled dangling
 *                          leaves, suppress in=out=1 nodes); label clean-up = containment.hpp:1110-1125
 *                          (label only in T => not displayed; label only in N => leaf removed);
 *                          <=2 common labels => displayed (containment.hpp:1061,1208).
 *                          The reference decides this with reduction rules + branching (and is unreliable,
 *                          SURVEY F4); the oracle enumerates all prod(in-degree) switchings, which is the
 *                          definition.  PARITY PINNING: no golden verdicts exist in the reference
 *                          (SURVEY 4.2); pinned instead against outputs of the reference itself run in the
 *                          build container (tests/golden/tc_ref_verdicts.json, made by oracle/make_golden.py).
 *   orc_lca_naive       -- Tree::naiveLCA utils/tree.hpp:202-219 (alternating parent walk with a seen set).
 *   orc_lca_euler_*     -- rmq/lca.hpp:61-106: Euler tour (node emitted on entry and after each child),
 *                          depth array, node -> LAST Euler index, RMQ over depths.
 *   orc_sparse_rmq_*    -- rmq/sparse_rmq.hpp:45-90: table[d][i] = index of min of [i, i+2^d); on ties the
 *                          RIGHT operand wins (val(x) < val(y) ? x : y), so the rightmost minimum is returned.
 *                          Pinned by the known answers of rmq/rmq_test.hpp:51-109, rmq/pm_rmq_test.hpp:41-44
 *                          and rmq/lca.cpp:46-49 (tests/test_oracle.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAXW 16 /* up to 1024 common labels */

typedef struct { uint64_t w[ORC_MAXW]; } bits_t;

static int cmp_bits_W;
static int cmp_bits(const void* a, const void* b) { return memcmp(a, b, (size_t)cmp_bits_W * 8); }

/* reverse topological order of a DAG given as arc list; returns 0 on success, -1 if cyclic */
static int topo_order(int n, int m, const int32_t* tail, const int32_t* head, int32_t* order) {
  int32_t* indeg = (int32_t*)calloc((size_t)n, 4); int32_t* off = (int32_t*)calloc((size_t)n + 1, 4); int32_t* adj = (int32_t*)malloc((size_t)(m > 0 ? m : 1) * 4);
  for (int e = 0; e < m; ++e) { indeg[head[e]]++; off[tail[e] + 1]++; }
  for (int v = 0; v < n; ++v) off[v + 1] += off[v];
  int32_t* pos = (int32_t*)malloc((size_t)n * 4); memcpy(pos, off, (size_t)n * 4);
  for (int e = 0; e < m; ++e) adj[pos[tail[e]]++] = head[e];
  int cnt = 0;
  for (int v = 0; v < n; ++v) if (!indeg[v]) order[cnt++] = v;
  for (int i = 0; i < cnt; ++i) { int v = order[i]; for (int k = off[v]; k < off[v + 1]; ++k) if (--indeg[adj[k]] == 0) order[cnt++] = adj[k]; }
  free(indeg); free(off); free(adj); free(pos);
  return cnt == n ? 0 : -1;
}

/* Sorted set of clusters (size>=2) of the cleaned tree obtained from a DAG in which every node has
 * at most ONE active in-arc (act[e] != 0).  lab[v] = bit index of the leaf label, or -1. */
static int clusters_of(int n, int m, const int32_t* tail, const int32_t* head, const uint8_t* act,
                       const int32_t* lab, const int32_t* order, int W, bits_t* cl, int32_t* nalive, bits_t* out) {
  for (int v = 0; v < n; ++v) { memset(&cl[v], 0, sizeof(bits_t)); nalive[v] = 0; if (lab[v] >= 0) cl[v].w[lab[v] >> 6] |= 1ull << (lab[v] & 63); }
  /* process nodes children-first: walk the topological order backwards, pushing each node's cluster up its active in-arc */
  int32_t* up = (int32_t*)malloc((size_t)n * 4);
  for (int v = 0; v < n; ++v) up[v] = -1;
  for (int e = 0; e < m; ++e) if (act[e]) up[head[e]] = tail[e];
  int cnt = 0;
  for (int i = n - 1; i >= 0; --i) {
    int v = order[i]; int nz = 0;
    for (int k = 0; k < W; ++k) nz |= cl[v].w[k] != 0;
    if (!nz) continue;                      /* dead: unlabelled, nothing alive below */
    if (nalive[v] >= 2) out[cnt++] = cl[v]; /* a branching node of the cleaned tree */
    int p = up[v];
    if (p >= 0) { for (int k = 0; k < W; ++k) cl[p].w[k] |= cl[v].w[k]; nalive[p]++; }
  }
  free(up);
  cmp_bits_W = W; qsort(out, (size_t)cnt, sizeof(bits_t), cmp_bits);
  return cnt;
}

/* Exhaustive-switching containment.
 * Host: nh nodes, mh arcs (htail/hhead), hlab[v] = label id (any non-negative int) or -1 for unlabelled.
 * Guest: nt nodes, mt arcs, tlab likewise; guest must be a tree.  Labels on nodes with out-degree > 0 are ignored.
 * Returns 1 displayed, 0 not displayed, <0 error (-1 cyclic, -2 duplicate leaf label, -3 too many labels,
 * -4 guest not a tree, -5 too many switchings (> max_switchings)).
 * n_displaying (optional) receives the number of switchings that display the guest (only when the full
 * enumeration runs, i.e. stop_at_first == 0). first_index (optional) receives the first displaying mixed-radix index. */
int orc_tc_bruteforce(int nh, int mh, const int32_t* htail, const int32_t* hhead, const int32_t* hlab,
                      int nt, int mt, const int32_t* ttail, const int32_t* thead, const int32_t* tlab,
                      int stop_at_first, uint64_t max_switchings, uint64_t* n_displaying, uint64_t* first_index) {
  if (n_displaying) *n_displaying = 0;
  if (first_index) *first_index = UINT64_MAX;
  if (mt != nt - 1 && !(nt == 0 && mt == 0)) return -4;
  int32_t* hout = (int32_t*)calloc((size_t)nh + 1, 4); int32_t* tout = (int32_t*)calloc((size_t)nt + 1, 4); int32_t* tin = (int32_t*)calloc((size_t)nt + 1, 4);
  for (int e = 0; e < mh; ++e) hout[htail[e]]++;
  for (int e = 0; e < mt; ++e) { tout[ttail[e]]++; if (++tin[thead[e]] > 1) { free(hout); free(tout); free(tin); return -4; } }
  /* dictionary of leaf labels */
  int32_t maxlab = -1;
  for (int v = 0; v < nh; ++v) if (hlab[v] > maxlab) maxlab = hlab[v];
  for (int v = 0; v < nt; ++v) if (tlab[v] > maxlab) maxlab = tlab[v];
  int32_t* in_h = (int32_t*)malloc((size_t)(maxlab + 2) * 4); int32_t* in_t = (int32_t*)malloc((size_t)(maxlab + 2) * 4);
  for (int i = 0; i <= maxlab; ++i) in_h[i] = in_t[i] = -1;
  int rc = 1, dup = 0;
  for (int v = 0; v < nh; ++v) if (hlab[v] >= 0 && hout[v] == 0) { if (in_h[hlab[v]] >= 0) dup = 1; in_h[hlab[v]] = v; }
  for (int v = 0; v < nt; ++v) if (tlab[v] >= 0 && tout[v] == 0) { if (in_t[tlab[v]] >= 0) dup = 1; in_t[tlab[v]] = v; }
  int32_t* hl = (int32_t*)malloc((size_t)(nh > 0 ? nh : 1) * 4); int32_t* tl = (int32_t*)malloc((size_t)(nt > 0 ? nt : 1) * 4);
  for (int v = 0; v < nh; ++v) hl[v] = -1;
  for (int v = 0; v < nt; ++v) tl[v] = -1;
  int ncommon = 0, t_only = 0;
  for (int i = 0; i <= maxlab; ++i) {
    if (in_t[i] >= 0 && in_h[i] < 0) t_only = 1;                       /* containment.hpp:1116 */
    if (in_t[i] >= 0 && in_h[i] >= 0) { hl[in_h[i]] = ncommon; tl[in_t[i]] = ncommon; ncommon++; } /* host-only labels stay -1: :1117-1121 */
  }
  free(in_h); free(in_t); free(hout); free(tout); free(tin);
  if (dup) { free(hl); free(tl); return -2; }
  if (t_only) { free(hl); free(tl); return 0; }
  if (ncommon <= 2) { free(hl); free(tl); if (n_displaying) *n_displaying = 0; if (first_index) *first_index = 0; return 1; } /* :1061,1208 */
  int W = (ncommon + 63) / 64;
  if (W > ORC_MAXW) { free(hl); free(tl); return -3; }
  int32_t* horder = (int32_t*)malloc((size_t)nh * 4); int32_t* torder = (int32_t*)malloc((size_t)nt * 4);
  if (topo_order(nh, mh, htail, hhead, horder) || topo_order(nt, mt, ttail, thead, torder)) { free(hl); free(tl); free(horder); free(torder); return -1; }
  int nmax = nh > nt ? nh : nt;
  bits_t* cl = (bits_t*)malloc((size_t)nmax * sizeof(bits_t)); int32_t* nal = (int32_t*)malloc((size_t)nmax * 4);
  bits_t* tcl = (bits_t*)malloc((size_t)nt * sizeof(bits_t)); bits_t* hcl = (bits_t*)malloc((size_t)nh * sizeof(bits_t));
  uint8_t* tact = (uint8_t*)malloc((size_t)(mt > 0 ? mt : 1)); memset(tact, 1, (size_t)(mt > 0 ? mt : 1));
  int tcnt = clusters_of(nt, mt, ttail, thead, tact, tl, torder, W, cl, nal, tcl);
  /* reticulations and their in-arcs */
  int32_t* indeg = (int32_t*)calloc((size_t)nh, 4);
  for (int e = 0; e < mh; ++e) indeg[hhead[e]]++;
  int nr = 0; for (int v = 0; v < nh; ++v) if (indeg[v] >= 2) nr++;
  int32_t* rnode = (int32_t*)malloc((size_t)(nr > 0 ? nr : 1) * 4); int32_t* ridx = (int32_t*)malloc((size_t)nh * 4);
  nr = 0; for (int v = 0; v < nh; ++v) { ridx[v] = -1; if (indeg[v] >= 2) { ridx[v] = nr; rnode[nr++] = v; } }
  int32_t* roff = (int32_t*)calloc((size_t)nr + 1, 4);
  for (int i = 0; i < nr; ++i) roff[i + 1] = roff[i] + indeg[rnode[i]];
  int32_t* rarc = (int32_t*)malloc((size_t)(roff[nr] > 0 ? roff[nr] : 1) * 4); int32_t* fill = (int32_t*)calloc((size_t)(nr > 0 ? nr : 1), 4);
  for (int e = 0; e < mh; ++e) { int r = ridx[hhead[e]]; if (r >= 0) rarc[roff[r] + fill[r]++] = e; }
  uint64_t total = 1; int overflow = 0;
  for (int i = 0; i < nr; ++i) { uint64_t d = (uint64_t)indeg[rnode[i]]; if (total > max_switchings / d) { overflow = 1; break; } total *= d; }
  uint8_t* act = (uint8_t*)malloc((size_t)(mh > 0 ? mh : 1));
  uint64_t ndisp = 0, first = UINT64_MAX;
  if (overflow) rc = -5;
  else {
    for (uint64_t s = 0; s < total; ++s) {
      for (int e = 0; e < mh; ++e) act[e] = ridx[hhead[e]] < 0;
      uint64_t x = s;
      for (int i = 0; i < nr; ++i) { int d = indeg[rnode[i]]; act[rarc[roff[i] + (int)(x % (uint64_t)d)]] = 1; x /= (uint64_t)d; }
      int hcnt = clusters_of(nh, mh, htail, hhead, act, hl, horder, W, cl, nal, hcl);
      if (hcnt == tcnt && memcmp(hcl, tcl, (size_t)hcnt * sizeof(bits_t)) == 0) {
        /* memcmp over whole bits_t is fine: unused words are zero on both sides */
        if (first == UINT64_MAX) first = s;
        ndisp++;
        if (stop_at_first) break;
      }
    }
    rc = ndisp > 0;
  }
  if (n_displaying) *n_displaying = ndisp;
  if (first_index) *first_index = first;
  free(hl); free(tl); free(horder); free(torder); free(cl); free(nal); free(tcl); free(hcl); free(tact); free(indeg);
  free(rnode); free(ridx); free(roff); free(rarc); free(fill); free(act);
  return rc;
}

/* ---- LCA: utils/tree.hpp:202-219 (naiveLCA): alternate parent steps from x and y, first node seen twice wins ---- */
int32_t orc_lca_naive(const int32_t* parent, int32_t n, int32_t x, int32_t y, uint8_t* seen /* n bytes, zeroed; restored on return */) {
  int32_t touched_cap = 64, nt = 0; int32_t* touched = (int32_t*)malloc((size_t)touched_cap * 4); int32_t res = -1;
  (void)n;
  while (x != y) {
    int32_t* zs[2] = {&x, &y}; int done = 0;
    for (int k = 0; k < 2 && !done; ++k) {
      int32_t* z = zs[k];
      if (parent[*z] < 0) continue;               /* update_for_LCA: root never moves (tree.hpp:214) */
      if (seen[*z]) { res = *z; done = 1; break; }
      seen[*z] = 1; if (nt == touched_cap) { touched_cap *= 2; touched = (int32_t*)realloc(touched, (size_t)touched_cap * 4); } touched[nt++] = *z;
      *z = parent[*z];
    }
    if (done) break;
  }
  if (res < 0) res = x;
  for (int i = 0; i < nt; ++i) seen[touched[i]] = 0;
  free(touched);
  return res;
}
void orc_lca_naive_batch(const int32_t* parent, int32_t n, const int32_t* u, const int32_t* v, int32_t* out, int64_t q) {
  uint8_t* seen = (uint8_t*)calloc((size_t)n, 1);
  for (int64_t i = 0; i < q; ++i) out[i] = orc_lca_naive(parent, n, u[i], v[i], seen);
  free(seen);
}

/* ---- sparse_rmq: rmq/sparse_rmq.hpp:45-90 ---- */
typedef struct { int64_t n; int levels; const int32_t* val; int64_t** arr; } orc_sparse_rmq;
static int floor_log2_i64(int64_t x) { int l = 0; while ((x >> (l + 1)) > 0) ++l; return l; }
orc_sparse_rmq* orc_sparse_rmq_build(const int32_t* val, int64_t n) {
  orc_sparse_rmq* R = (orc_sparse_rmq*)malloc(sizeof(*R)); R->n = n; R->val = val;
  int logn = n >= 2 ? floor_log2_i64(n) : 1; if (logn < 1) logn = 1;      /* max(1, floor(log2 n)) :69-70 */
  R->levels = logn + 1; R->arr = (int64_t**)calloc((size_t)R->levels, sizeof(int64_t*));
  R->arr[0] = (int64_t*)malloc((size_t)(n > 0 ? n : 1) * 8);
  for (int64_t i = 0; i < n; ++i) R->arr[0][i] = i;                        /* :48 */
  for (int d = 0; d < logn; ++d) {                                          /* :52-65 */
    int64_t width = (int64_t)1 << d, cnt = n - 2 * width + 1 + 0;           /* |arr[d]| - width = (n-2^d+1) - 2^d */
    int64_t len_d = n - width + 1; cnt = len_d - width; if (cnt < 0) cnt = 0;
    R->arr[d + 1] = (int64_t*)malloc((size_t)(cnt > 0 ? cnt : 1) * 8);
    for (int64_t i = 0; i < cnt; ++i) { int64_t x = R->arr[d][i], y = R->arr[d][i + width]; R->arr[d + 1][i] = val[x] < val[y] ? x : y; }
  }
  return R;
}
/* query over the half-open range [u, v), u < v (rmq/rmq.hpp:56-63) */
int64_t orc_sparse_rmq_query(const orc_sparse_rmq* R, int64_t u, int64_t v) {
  int depth = (v - u - 1) >= 1 ? floor_log2_i64(v - u - 1) : 0;             /* max(0, floor(log2(v-u-1))) :82-83 */
  int64_t px = R->arr[depth][u], py = R->arr[depth][v - ((int64_t)1 << depth)];
  return R->val[px] < R->val[py] ? px : py;                                 /* :89 ties -> right */
}
void orc_sparse_rmq_free(orc_sparse_rmq* R) { for (int d = 0; d < R->levels; ++d) free(R->arr[d]); free(R->arr); free(R); }

/* ---- lca: rmq/lca.hpp:61-106 (Euler tour + depth + last index) over a sparse_rmq ---- */
typedef struct { int64_t n, E; int32_t* euler; int32_t* depth; int64_t* last; orc_sparse_rmq* rmq; } orc_lca;
orc_lca* orc_lca_euler_build(const int32_t* parent, int32_t n) {
  orc_lca* L = (orc_lca*)malloc(sizeof(*L)); L->n = n; L->E = 2 * (int64_t)n - 1;
  L->euler = (int32_t*)malloc((size_t)L->E * 4); L->depth = (int32_t*)malloc((size_t)L->E * 4); L->last = (int64_t*)malloc((size_t)n * 8);
  int32_t* off = (int32_t*)calloc((size_t)n + 1, 4); int32_t* adj = (int32_t*)malloc((size_t)n * 4); int32_t root = -1;
  for (int32_t v = 0; v < n; ++v) { if (parent[v] < 0) root = v; else off[parent[v] + 1]++; }
  for (int32_t v = 0; v < n; ++v) off[v + 1] += off[v];
  int32_t* pos = (int32_t*)malloc((size_t)n * 4); memcpy(pos, off, (size_t)n * 4);
  for (int32_t v = 0; v < n; ++v) if (parent[v] >= 0) adj[pos[parent[v]]++] = v;
  /* iterative version of preprocess(): emit on entry and after each child (:61-71) */
  int32_t* stk = (int32_t*)malloc((size_t)n * 4); int32_t* it = (int32_t*)malloc((size_t)n * 4); int sp = 0; int64_t k = 0; int32_t d = 0;
  stk[sp] = root; it[sp] = off[root]; sp++; L->euler[k] = root; L->depth[k] = 0; L->last[root] = k; k++;
  while (sp > 0) {
    int32_t v = stk[sp - 1];
    if (it[sp - 1] < off[v + 1]) { int32_t c = adj[it[sp - 1]++]; d++; stk[sp] = c; it[sp] = off[c]; sp++; L->euler[k] = c; L->depth[k] = d; L->last[c] = k; k++; }
    else { sp--; d--; if (sp > 0) { int32_t p = stk[sp - 1]; L->euler[k] = p; L->depth[k] = d; L->last[p] = k; k++; } }
  }
  L->rmq = orc_sparse_rmq_build(L->depth, L->E);
  free(off); free(adj); free(pos); free(stk); free(it);
  return L;
}
int32_t orc_lca_euler_query(const orc_lca* L, int32_t u, int32_t v) {
  int64_t a = L->last[u], b = L->last[v]; if (b < a) { int64_t t = a; a = b; b = t; }    /* :95-99 */
  return L->euler[orc_sparse_rmq_query(L->rmq, a, b + 1)];
}
void orc_lca_euler_batch(const orc_lca* L, const int32_t* u, const int32_t* v, int32_t* out, int64_t q) { for (int64_t i = 0; i < q; ++i) out[i] = orc_lca_euler_query(L, u[i], v[i]); }
void orc_lca_euler_free(orc_lca* L) { orc_sparse_rmq_free(L->rmq); free(L->euler); free(L->depth); free(L->last); free(L); }

/* ---- who_displays for a single-labelled tree host: TreeInTreeContainment::who_displays utils/containment.hpp:72-177 ----
 * The reference's DP says: host node v displays guest node u iff the host subtree below v, RESTRICTED to the labels below u
 * and cleaned, equals the guest subtree below u, and v is the lowest such node (compute_possibilities :128-171 builds the
 * subtree induced by the children's images and matches children to distinct branches).  Restated from that definition with
 * label bitsets, independently of the LCA machinery the product uses: for every guest node u, intersect every host
 * cluster with X_u = labels(T_u), take the deepest host node whose restricted cluster is X_u, and compare the two sets of
 * branching clusters.  out[u] = that host node or -1.  Pinned against the reference itself (tests/golden/who_ref.json). */
int orc_who_displays(int n, const int32_t* hparent, const int32_t* hlabel, int nt, const int32_t* tparent, const int32_t* tlabel, int32_t* out) {
  int32_t* hdeg = (int32_t*)calloc((size_t)n, 4); int32_t* tdeg = (int32_t*)calloc((size_t)nt, 4);
  for (int v = 0; v < n; ++v) if (hparent[v] >= 0) hdeg[hparent[v]]++;
  for (int t = 0; t < nt; ++t) if (tparent[t] >= 0) tdeg[tparent[t]]++;
  int32_t maxlab = -1;
  for (int v = 0; v < n; ++v) if (hlabel[v] > maxlab) maxlab = hlabel[v];
  for (int t = 0; t < nt; ++t) if (tlabel[t] > maxlab) maxlab = tlabel[t];
  int32_t* l2h = (int32_t*)malloc((size_t)(maxlab + 2) * 4);
  for (int i = 0; i <= maxlab; ++i) l2h[i] = -1;
  for (int v = 0; v < n; ++v) if (hlabel[v] >= 0 && hdeg[v] == 0) { if (l2h[hlabel[v]] >= 0) { free(hdeg); free(tdeg); free(l2h); return -2; } l2h[hlabel[v]] = v; }
  const int W = (maxlab + 64) / 64 > 0 ? (maxlab + 64) / 64 : 1;
  if (W > ORC_MAXW) { free(hdeg); free(tdeg); free(l2h); return -3; }
  /* children-first orders by depth */
  int32_t* hdepth = (int32_t*)calloc((size_t)n, 4); int32_t* tdepth = (int32_t*)calloc((size_t)nt, 4);
  for (int v = 0; v < n; ++v) { int d = 0, x = v; while (hparent[x] >= 0) { x = hparent[x]; ++d; } hdepth[v] = d; }
  for (int t = 0; t < nt; ++t) { int d = 0, x = t; while (tparent[x] >= 0) { x = tparent[x]; ++d; } tdepth[t] = d; }
  int maxhd = 0, maxtd = 0;
  for (int v = 0; v < n; ++v) if (hdepth[v] > maxhd) maxhd = hdepth[v];
  for (int t = 0; t < nt; ++t) if (tdepth[t] > maxtd) maxtd = tdepth[t];
  bits_t* X = (bits_t*)calloc((size_t)nt, sizeof(bits_t)); uint8_t* bad = (uint8_t*)calloc((size_t)nt, 1); int32_t* talive = (int32_t*)calloc((size_t)nt, 4);
  for (int d = maxtd; d >= 0; --d) for (int t = 0; t < nt; ++t) if (tdepth[t] == d) {
    if (tdeg[t] == 0 && tlabel[t] >= 0) { X[t].w[tlabel[t] >> 6] |= 1ull << (tlabel[t] & 63); if (l2h[tlabel[t]] < 0) bad[t] = 1; }
    int nz = 0; for (int k = 0; k < W; ++k) nz |= X[t].w[k] != 0;
    if (nz && tparent[t] >= 0) { for (int k = 0; k < W; ++k) X[tparent[t]].w[k] |= X[t].w[k]; talive[tparent[t]]++; bad[tparent[t]] |= bad[t]; }
  }
  bits_t* C = (bits_t*)malloc((size_t)n * sizeof(bits_t)); int32_t* calive = (int32_t*)malloc((size_t)n * 4);
  bits_t* tset = (bits_t*)malloc((size_t)nt * sizeof(bits_t)); bits_t* hset = (bits_t*)malloc((size_t)n * sizeof(bits_t));
  for (int u = 0; u < nt; ++u) {
    out[u] = -1;
    int nz = 0; for (int k = 0; k < W; ++k) nz |= X[u].w[k] != 0;
    if (!nz || bad[u]) continue;
    /* host clusters restricted to X_u */
    for (int v = 0; v < n; ++v) { memset(&C[v], 0, sizeof(bits_t)); calive[v] = 0; if (hdeg[v] == 0 && hlabel[v] >= 0) { const int l = hlabel[v]; if ((X[u].w[l >> 6] >> (l & 63)) & 1ull) C[v].w[l >> 6] |= 1ull << (l & 63); } }
    for (int d = maxhd; d >= 0; --d) for (int v = 0; v < n; ++v) if (hdepth[v] == d) {
      int z = 0; for (int k = 0; k < W; ++k) z |= C[v].w[k] != 0;
      if (z && hparent[v] >= 0) { for (int k = 0; k < W; ++k) C[hparent[v]].w[k] |= C[v].w[k]; calive[hparent[v]]++; }
    }
    int best = -1;
    for (int v = 0; v < n; ++v) if (memcmp(&C[v], &X[u], sizeof(bits_t)) == 0 && (best < 0 || hdepth[v] > hdepth[best])) best = v;
    if (best < 0) continue;
    /* branching clusters of T_u and of the restricted host subtree below `best` */
    int tc = 0, hc = 0;
    for (int t = 0; t < nt; ++t) if (talive[t] >= 2) { int in = 1; for (int k = 0; k < W; ++k) if (X[t].w[k] & ~X[u].w[k]) in = 0; int x = t; while (x != u && x >= 0) x = tparent[x]; if (in && x == u) tset[tc++] = X[t]; }
    for (int v = 0; v < n; ++v) if (calive[v] >= 2) { int x = v; while (x != best && x >= 0) x = hparent[x]; if (x == best) hset[hc++] = C[v]; }
    if (tc != hc) continue;
    cmp_bits_W = ORC_MAXW; qsort(tset, (size_t)tc, sizeof(bits_t), cmp_bits); qsort(hset, (size_t)hc, sizeof(bits_t), cmp_bits);
    if (memcmp(tset, hset, (size_t)tc * sizeof(bits_t)) == 0) out[u] = best;
  }
  free(hdeg); free(tdeg); free(l2h); free(hdepth); free(tdepth); free(X); free(bad); free(talive); free(C); free(calive); free(tset); free(hset);
  return 0;
}

/* ---- who_displays for MULTI-LABELLED tree hosts: TreeInTreeContainment::who_displays utils/containment.hpp:72-230 --------
 * The reference keeps, per guest node u, the preorder-sorted list of host nodes that MINIMALLY display the guest subtree below
 * u: base case = the host leaves carrying u's label (:115-126); a node with one child maps where the child maps (:167); for
 * d >= 2 children it builds the host subtree induced by the children's candidates (utils/induced_tree.hpp:65-168), lets every
 * candidate tell its induced ancestors which guest child it can display (:140-149) and accepts, lowest first, every induced
 * node at which the guest children can be matched to DISTINCT induced children (bipartite matching, utils/matching.hpp; :152-163).
 * Restated from the DEFINITION that algorithm computes: A(u) = host nodes whose subtree displays T_u (leaves not needed may be
 * ignored: a MUL-tree displays T if SOME choice of one leaf per label does), A'(u) = nodes x with an assignment of u's children
 * to distinct children y of x, y in A(child); A(u) = ancestors-or-self of A'(u); the answer M(u) = the minimal nodes of A'(u).
 * Every host node is tried, no induced tree, no LCA.  Unlabelled guest leaves are cleaned away (as everywhere in this repo).
 * out_off[nt+1] / out_adj: M(u) in ascending node id (the reference sorts by ITS preorder: compare as sets).
 * returns the total number of entries, or -1 if out_adj (capacity cap) is too small, -3 on more than 4096 guest or host children.
 * (adjacency of the matching: adj[y * words + i / 64] bit i % 64 = "guest child i can be displayed below host child y") */
typedef struct { int nl, nr, words; const uint64_t* adj; int matchR[4096]; uint8_t seen[4096]; } orc_match_ctx;
static int orc_try_match(orc_match_ctx* c, int i) {
  for (int y = 0; y < c->nr; ++y) if (((c->adj[(size_t)y * c->words + (i >> 6)] >> (i & 63)) & 1ull) && !c->seen[y]) {
    c->seen[y] = 1;
    if (c->matchR[y] < 0 || orc_try_match(c, c->matchR[y])) { c->matchR[y] = i; return 1; }
  }
  return 0;
}
int orc_mul_who_displays(int n, const int32_t* hparent, const int32_t* hlabel, int nt, const int32_t* tparent, const int32_t* tlabel,
                         int32_t* out_off, int32_t* out_adj, int cap) {
  int32_t* hdeg = (int32_t*)calloc((size_t)n, 4); int32_t* tdeg = (int32_t*)calloc((size_t)nt, 4);
  for (int v = 0; v < n; ++v) if (hparent[v] >= 0) hdeg[hparent[v]]++;
  for (int t = 0; t < nt; ++t) if (tparent[t] >= 0) tdeg[tparent[t]]++;
  int32_t* tdepth = (int32_t*)calloc((size_t)nt, 4); int maxtd = 0;
  for (int t = 0; t < nt; ++t) { int d = 0, x = t; while (tparent[x] >= 0) { x = tparent[x]; ++d; } tdepth[t] = d; if (d > maxtd) maxtd = d; }
  /* host children lists */
  int32_t* hoff = (int32_t*)calloc((size_t)n + 1, 4); int32_t* hadj = (int32_t*)malloc((size_t)(n + 1) * 4);
  for (int v = 0; v < n; ++v) if (hparent[v] >= 0) hoff[hparent[v] + 1]++;
  for (int v = 0; v < n; ++v) hoff[v + 1] += hoff[v];
  { int32_t* pos = (int32_t*)malloc((size_t)(n + 1) * 4); memcpy(pos, hoff, (size_t)(n + 1) * 4); for (int v = 0; v < n; ++v) if (hparent[v] >= 0) hadj[pos[hparent[v]]++] = v; free(pos); }
  /* A[u] (n bytes per guest node), state: 0 computed, 1 dead (cleaned away) */
  uint8_t* A = (uint8_t*)calloc((size_t)nt * (size_t)n, 1); uint8_t* M = (uint8_t*)calloc((size_t)nt * (size_t)n, 1); uint8_t* dead = (uint8_t*)calloc((size_t)nt, 1);
  int rc = 0;
  orc_match_ctx* mc = (orc_match_ctx*)malloc(sizeof(orc_match_ctx));
  uint64_t* adj = (uint64_t*)malloc((size_t)4096 * 64 * 8);
  int* kids = (int*)malloc((size_t)4096 * sizeof(int));
  for (int d = maxtd; d >= 0 && rc == 0; --d) for (int u = 0; u < nt && rc == 0; ++u) if (tdepth[u] == d) {
    uint8_t* Au = A + (size_t)u * n; uint8_t* Mu = M + (size_t)u * n;
    if (tdeg[u] == 0) {
      if (tlabel[u] < 0) { dead[u] = 1; continue; }
      for (int v = 0; v < n; ++v) if (hdeg[v] == 0 && hlabel[v] == tlabel[u]) Mu[v] = 1;
    } else {
      int nk = 0, empty = 0;
      for (int c = 0; c < nt; ++c) if (tparent[c] == u && !dead[c]) {
        if (nk == 4096) { rc = -3; break; }
        kids[nk++] = c;
        int any = 0; for (int v = 0; v < n; ++v) any |= M[(size_t)c * n + v];
        if (!any) empty = 1;
      }
      if (rc) break;
      if (nk == 0) { dead[u] = 1; continue; }
      if (empty) continue;
      if (nk == 1) memcpy(Mu, M + (size_t)kids[0] * n, (size_t)n);
      else {
        uint8_t* Ap = (uint8_t*)calloc((size_t)n, 1);
        for (int x = 0; x < n; ++x) {
          const int r = hoff[x + 1] - hoff[x];
          if (r < nk || r > 4096) continue;
          const int words = (nk + 63) / 64;
          memset(adj, 0, (size_t)r * words * 8);
          for (int k = 0; k < r; ++k) { const int y = hadj[hoff[x] + k]; for (int i = 0; i < nk; ++i) if (A[(size_t)kids[i] * n + y]) adj[(size_t)k * words + (i >> 6)] |= 1ull << (i & 63); }
          mc->nl = nk; mc->nr = r; mc->words = words; mc->adj = adj;
          for (int k = 0; k < r; ++k) mc->matchR[k] = -1;
          int ok = 1;
          for (int i = 0; i < nk && ok; ++i) { memset(mc->seen, 0, (size_t)r); ok = orc_try_match(mc, i); }
          Ap[x] = (uint8_t)ok;
        }
        /* minimal elements of A' */
        for (int x = 0; x < n; ++x) if (Ap[x]) {
          int below = 0;
          for (int w = 0; w < n && !below; ++w) if (w != x && Ap[w]) { int z = hparent[w]; while (z >= 0 && z != x) z = hparent[z]; if (z == x) below = 1; }
          if (!below) Mu[x] = 1;
        }
        free(Ap);
      }
    }
    /* A(u) = ancestors-or-self of M(u) */
    for (int v = 0; v < n; ++v) if (Mu[v]) { int z = v; while (z >= 0 && !Au[z]) { Au[z] = 1; z = hparent[z]; } }
  }
  int total = 0;
  if (rc == 0) {
    for (int u = 0; u < nt; ++u) {
      out_off[u] = total;
      for (int v = 0; v < n; ++v) if (M[(size_t)u * n + v]) { if (total >= cap) { rc = -1; break; } out_adj[total++] = v; }
      if (rc) break;
    }
    out_off[nt] = total;
  }
  free(hdeg); free(tdeg); free(tdepth); free(hoff); free(hadj); free(A); free(M); free(dead); free(mc); free(adj); free(kids);
  return rc ? rc : total;
}

/* highest_displayed_ancestor over LISTS (the general table): as orc_highest_displayed, "displayed by the host root alone" =
 * the list of that guest node is exactly { host_root } (utils/containment.hpp:340) */
int orc_highest_displayed_lists(int nt, const int32_t* tparent, const int32_t* off, const int32_t* adj, int32_t host_root, int32_t* out) {
  for (int s = 0; s < nt; ++s) {
    int v = s;
    if (tparent[v] < 0) { out[s] = v; continue; }
    int pv = tparent[v];
    for (;;) {
      if (off[pv + 1] == off[pv]) break;
      v = pv;
      if (tparent[v] < 0) break;
      if (off[v + 1] - off[v] == 1 && adj[off[v]] == host_root) break;
      pv = tparent[v];
    }
    out[s] = v;
  }
  return 0;
}

/* ---- unzip_retis: TreeInComponent::unzip_retis utils/containment.hpp:262-312 ------------------------------------------------
 * The reference walks the host below u with an EDGE traversal that keeps no seen-set, so everything below a reticulation is
 * visited once per way of reaching it: the network unfolds into a multi-labelled tree.  Arcs leaving a node of out-degree 1 are
 * skipped (:281) and a head of out-degree 1 is replaced by the end of its out-degree-1 chain (:284), i.e. unary nodes
 * (reticulations, suppressible nodes) disappear.  New node ids are handed out in visiting order (a preorder; the visiting order
 * of siblings is the storage's and not contractual); only labels of out-degree-0 nodes are kept (leaf_labels_only).
 * Restated as a plain recursive unfolding.  succ_off/succ_adj: host CSR; out_parent / out_label / out_origin (the host node a
 * copy stands for): capacity cap; returns the number of nodes of the MUL-tree, or -1 when it exceeds cap. */
typedef struct { const int32_t *off, *adj, *lab; int32_t *par, *olab, *orig; int cap, count; } orc_unzip_ctx;
static int orc_unzip_rec(orc_unzip_ctx* c, int x, int st_x) {
  if (c->off[x + 1] - c->off[x] == 1) return 0;                       /* never entered: heads are chain ends */
  for (int k = c->off[x]; k < c->off[x + 1]; ++k) {
    int y = c->adj[k];
    while (c->off[y + 1] - c->off[y] == 1) y = c->adj[c->off[y]];
    if (c->count >= c->cap) return -1;
    const int st_y = c->count++;
    c->par[st_y] = st_x; c->orig[st_y] = y;
    c->olab[st_y] = (c->off[y + 1] == c->off[y]) ? c->lab[y] : -1;
    if (orc_unzip_rec(c, y, st_y) < 0) return -1;
  }
  return 0;
}
int orc_unzip(int n, const int32_t* succ_off, const int32_t* succ_adj, const int32_t* label, int u, int32_t* out_parent, int32_t* out_label, int32_t* out_origin, int cap) {
  (void)n;
  orc_unzip_ctx c = {succ_off, succ_adj, label, out_parent, out_label, out_origin, cap, 0};
  if (cap < 1) return -1;
  /* the root: the reference translates u itself to 0 (:274) whatever its out-degree; if u is unary its arcs are skipped and the
   * subtree is the single node 0 -- callers pass a tree-component root, which is never unary after the clean-up */
  c.par[0] = -1; c.orig[0] = u; c.olab[0] = (succ_off[u + 1] == succ_off[u]) ? label[u] : -1; c.count = 1;
  if (orc_unzip_rec(&c, u, 0) < 0) return -1;
  return c.count;
}

/* ---- highest_displayed_ancestor: TreeInComponent::highest_displayed_ancestor utils/containment.hpp:328-344 -----------------
 * The reference's loop, one guest node at a time, over the who_displays table (who[u] = the one host node displaying u, or -1):
 *   pv = parent(v); forever { if who[pv] is empty return v; v = pv; if v is the guest root return v;
 *                             if who[v] == { host root } return v; pv = parent(v); }
 * out[root] = root (the reference is never called on the root).  PARITY UNPINNED: the function is private to TreeInComponent and
 * cannot be driven from outside the reference's network solver; this is a restatement of the quoted lines only. */
int orc_highest_displayed(int nt, const int32_t* tparent, const int32_t* who, int32_t host_root, int32_t* out) {
  for (int s = 0; s < nt; ++s) {
    int v = s;
    if (tparent[v] < 0) { out[s] = v; continue; }
    int pv = tparent[v];
    for (;;) {
      if (who[pv] < 0) break;
      v = pv;
      if (tparent[v] < 0) break;
      if (who[v] == host_root) break;
      pv = tparent[v];
    }
    out[s] = v;
  }
  return 0;
}

/* ---- bridges: utils/bridges.hpp:42-118 (BridgeFinder) -------------------------------------------------------------------
 * The reference finds the arcs whose removal disconnects the underlying undirected graph with two DFS passes over DFS
 * intervals (is_bridge_head :27-30).  Restated from the DEFINITION (remove the arc, test connectivity), which is
 * independent of both the reference's interval trick and the product's LCA + difference-counting method.
 * is_bridge[e] for arcs given as (tail[e], head[e]); returns the number of bridges.  Pinned against the reference itself
 * (tests/golden/bridges_ref.json). */
int orc_bridges(int n, int m, const int32_t* tail, const int32_t* head, uint8_t* is_bridge) {
  int32_t* off = (int32_t*)calloc((size_t)n + 1, 4); int32_t* adj = (int32_t*)malloc((size_t)(2 * m + 1) * 4); int32_t* aid = (int32_t*)malloc((size_t)(2 * m + 1) * 4);
  for (int e = 0; e < m; ++e) { off[tail[e] + 1]++; off[head[e] + 1]++; }
  for (int v = 0; v < n; ++v) off[v + 1] += off[v];
  int32_t* pos = (int32_t*)malloc((size_t)(n + 1) * 4); memcpy(pos, off, (size_t)(n + 1) * 4);
  for (int e = 0; e < m; ++e) { adj[pos[tail[e]]] = head[e]; aid[pos[tail[e]]++] = e; adj[pos[head[e]]] = tail[e]; aid[pos[head[e]]++] = e; }
  uint8_t* seen = (uint8_t*)malloc((size_t)n); int32_t* stk = (int32_t*)malloc((size_t)n * 4);
  int count = 0;
  for (int e = 0; e < m; ++e) {
    memset(seen, 0, (size_t)n);
    int sp = 0; stk[sp++] = tail[e]; seen[tail[e]] = 1;
    while (sp) { const int v = stk[--sp]; for (int k = off[v]; k < off[v + 1]; ++k) if (aid[k] != e && !seen[adj[k]]) { seen[adj[k]] = 1; stk[sp++] = adj[k]; } }
    is_bridge[e] = !seen[head[e]];
    count += is_bridge[e];
  }
  free(off); free(adj); free(aid); free(pos); free(seen); free(stk);
  return count;
}

/* ---- network isomorphism: utils/isomorphism.hpp:33-324 (IsomorphismMapper::check_isomorph), examples/iso.cpp:96-112 ------
 * The reference keeps a set of possible images per node, propagates along arcs and branches on a node with several
 * possibilities.  Restated from the DEFINITION: is there a bijection phi of the nodes with (u,v) an arc <=> (phi u, phi v) an
 * arc, and label(u) == label(phi u) for every node whose label counts under `flags` (0x01 leaves, 0x02 internal tree nodes,
 * 0x04 internal reticulations: FLAG_MAP_* :10-13)?  Plain backtracking over the nodes in an order in which every node but
 * the first follows a neighbour, candidates filtered by degrees, counted label and the arcs to already mapped nodes.
 * lab*: dense label ids shared by the two networks, -1 = unlabelled.  returns 1 / 0. */
typedef struct {
  int n, m; const int32_t *t1, *h1, *l1, *t2, *h2, *l2; int flags;
  int32_t *so1, *sa1, *po1, *pa1, *so2, *sa2, *po2, *pa2, *order, *map, *inv;
} orc_iso_ctx;
static void iso_csr(int n, int m, const int32_t* key, const int32_t* val, int32_t* off, int32_t* adj) {
  memset(off, 0, (size_t)(n + 1) * 4);
  for (int e = 0; e < m; ++e) off[key[e] + 1]++;
  for (int v = 0; v < n; ++v) off[v + 1] += off[v];
  int32_t* pos = (int32_t*)malloc((size_t)(n + 1) * 4); memcpy(pos, off, (size_t)(n + 1) * 4);
  for (int e = 0; e < m; ++e) adj[pos[key[e]]++] = val[e];
  free(pos);
}
static int iso_has(const int32_t* off, const int32_t* adj, int u, int v) { int c = 0; for (int k = off[u]; k < off[u + 1]; ++k) c += adj[k] == v; return c; }
static int iso_label(const orc_iso_ctx* c, int side, int v) {
  const int32_t* so = side ? c->so2 : c->so1; const int32_t* po = side ? c->po2 : c->po1; const int32_t* l = side ? c->l2 : c->l1;
  const int od = so[v + 1] - so[v], id = po[v + 1] - po[v];
  const int counts = od == 0 ? (c->flags & 1) : (id >= 2 ? (c->flags & 4) : (c->flags & 2));
  return counts ? l[v] : -1;
}
static int iso_rec(orc_iso_ctx* c, int k) {
  if (k == c->n) return 1;
  const int u = c->order[k];
  for (int v = 0; v < c->n; ++v) {
    if (c->inv[v] >= 0) continue;
    if (c->so1[u + 1] - c->so1[u] != c->so2[v + 1] - c->so2[v] || c->po1[u + 1] - c->po1[u] != c->po2[v + 1] - c->po2[v]) continue;
    if (iso_label(c, 0, u) != iso_label(c, 1, v)) continue;
    int ok = 1;
    for (int j = c->so1[u]; ok && j < c->so1[u + 1]; ++j) { const int w = c->sa1[j]; if (c->map[w] >= 0 && iso_has(c->so1, c->sa1, u, w) != iso_has(c->so2, c->sa2, v, c->map[w])) ok = 0; }
    for (int j = c->po1[u]; ok && j < c->po1[u + 1]; ++j) { const int w = c->pa1[j]; if (c->map[w] >= 0 && iso_has(c->so1, c->sa1, w, u) != iso_has(c->so2, c->sa2, c->map[w], v)) ok = 0; }
    /* arcs of v towards already used images must have a preimage arc too (degrees are equal, so counting one direction suffices
       once every neighbour is mapped; checked here for early pruning) */
    for (int j = c->so2[v]; ok && j < c->so2[v + 1]; ++j) { const int w = c->inv[c->sa2[j]]; if (w >= 0 && !iso_has(c->so1, c->sa1, u, w)) ok = 0; }
    for (int j = c->po2[v]; ok && j < c->po2[v + 1]; ++j) { const int w = c->inv[c->pa2[j]]; if (w >= 0 && !iso_has(c->so1, c->sa1, w, u)) ok = 0; }
    if (!ok) continue;
    c->map[u] = v; c->inv[v] = u;
    if (iso_rec(c, k + 1)) return 1;
    c->map[u] = -1; c->inv[v] = -1;
  }
  return 0;
}
int orc_iso(int n1, int m1, const int32_t* tail1, const int32_t* head1, const int32_t* lab1,
            int n2, int m2, const int32_t* tail2, const int32_t* head2, const int32_t* lab2, int flags) {
  if (n1 != n2 || m1 != m2) return 0;
  if (n1 == 0) return 1;
  const int n = n1, m = m1;
  orc_iso_ctx c; c.n = n; c.m = m; c.t1 = tail1; c.h1 = head1; c.l1 = lab1; c.t2 = tail2; c.h2 = head2; c.l2 = lab2; c.flags = flags;
  int32_t* buf = (int32_t*)malloc((size_t)(4 * (n + 1) + 4 * (m + 1) + 3 * n + 8) * 4);
  int32_t* p = buf;
  c.so1 = p; p += n + 1; c.po1 = p; p += n + 1; c.so2 = p; p += n + 1; c.po2 = p; p += n + 1;
  c.sa1 = p; p += m + 1; c.pa1 = p; p += m + 1; c.sa2 = p; p += m + 1; c.pa2 = p; p += m + 1;
  c.order = p; p += n; c.map = p; p += n; c.inv = p;
  iso_csr(n, m, tail1, head1, c.so1, c.sa1); iso_csr(n, m, head1, tail1, c.po1, c.pa1);
  iso_csr(n, m, tail2, head2, c.so2, c.sa2); iso_csr(n, m, head2, tail2, c.po2, c.pa2);
  /* order: breadth-first over the undirected graph from every yet unseen node */
  int cnt = 0; uint8_t* seen = (uint8_t*)calloc((size_t)n, 1);
  for (int s = 0; s < n; ++s) {
    if (seen[s]) continue;
    seen[s] = 1; c.order[cnt++] = s;
    for (int q = cnt - 1; q < cnt; ++q) {
      const int u = c.order[q];
      for (int j = c.so1[u]; j < c.so1[u + 1]; ++j) if (!seen[c.sa1[j]]) { seen[c.sa1[j]] = 1; c.order[cnt++] = c.sa1[j]; }
      for (int j = c.po1[u]; j < c.po1[u + 1]; ++j) if (!seen[c.pa1[j]]) { seen[c.pa1[j]] = 1; c.order[cnt++] = c.pa1[j]; }
    }
  }
  for (int v = 0; v < n; ++v) { c.map[v] = -1; c.inv[v] = -1; }
  const int r = iso_rec(&c, 0);
  free(seen); free(buf);
  return r;
}
