// tests/host_sim/sim_tc.cpp -- TEST INFRASTRUCTURE ONLY.
// Single-threaded SIMULATOR of the CUDA reducer: instantiates csrc/tc_core.cuh (the very code the kernel runs)
// with an executor whose parallel-for runs sequentially in a permuted order.  It exists so that the rule logic
// can be checked against the oracle on a machine without a GPU and so that order dependence between "parallel"
// iterations shows up as a verdict that changes with the permutation seed.  It is not linked into libptgpu.so
// and nothing in phylo_tools_b200/ calls it.
//
//   sim_tc gen <count> <leaves> <retis> <seed> <kind> <perm>   -> one char per instance (1 / 0 / E) on stdout,
//                                                               a JSON line of statistics on stderr
//   sim_tc file <path> <perm>                                  (path: pairs of Newick lines)
//   environment SIM_LAYOUT = 4 | 2 | 1 : element type of the arrays the solver loads and emits (int32, int16, uint8 with degrees)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>
#include "../../include/pt_gpu.h"
#include "../../phylo_tools_b200/csrc/tc_core.cuh"
#include "../../phylo_tools_b200/include/pt/packer.hpp"

using namespace ptgpu;

struct HostEx {
  int perm;
  int phase = 0;  // 0 identity, 1 reverse, >=2 affine permutation seeded by perm
  int tid() const { return 0; }
  int nth() const { return 1; }
  void sync() const {}
  static int gcd(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }
  int map(int i, int n) const {
    if (perm == 0 || n <= 1) return i;
    if (perm == 1) return n - 1 - i;
    int a = (int)((unsigned)perm * 2654435761u % (unsigned)n); if (a == 0) a = 1;
    while (gcd(a, n) != 1) ++a;
    const int b = (int)((unsigned)perm * 40503u % (unsigned)n);
    return (int)(((long long)a * i + b) % n);
  }
  int uniform(const int32_t* p) const { return *p; }
  template <class S> int ldg(const S* p) const { return (int)*p; }
  int reduce_or(int v) const { return v; }
  int reduce_add(int v) const { return v; }
  template <class T> T atomic_add(T* p, T v) const { T o = *p; *p = (T)(o + v); return o; }
  template <class T> T atomic_min(T* p, T v) const { T o = *p; if (v < o) *p = v; return o; }
  template <class T> T atomic_max(T* p, T v) const { T o = *p; if (v > o) *p = v; return o; }
  template <class T> T atomic_or(T* p, T v) const { T o = *p; *p = (T)(o | v); return o; }
  template <class T> T atomic_xor(T* p, T v) const { T o = *p; *p = (T)(o ^ v); return o; }
  template <class T> T atomic_cas(T* p, T c, T v) const { T o = *p; if (o == c) *p = v; return o; }
  template <class T> int atomic_cas_idx(T* p, int c, int v) const { int o = *p; if (o == c) *p = (T)v; return o; }
  template <class T> int compact_arcs(T* at, T* ah, int m, T*, T*) const {
    int k = 0;
    for (int e = 0; e < m; ++e) if (at[e] >= 0) { at[k] = at[e]; ah[k] = ah[e]; ++k; }
    return k;
  }
  template <class T> int atomic_add_idx(T* p, int v) const { const int o = *p; *p = (T)(o + v); return o; }
  template <class T> void atomic_xor_idx(T* p, int v) const { *p = (T)(*p ^ v); }
  template <class C, class T> void exclusive_scan(const C* in, T* out, int n) const { int acc = 0; for (int i = 0; i < n; ++i) { out[i] = (T)acc; acc += in[i]; } out[n] = (T)acc; }
};

struct Item { std::vector<int32_t> succ_off, succ_adj, hlabel, tparent, tlabel; int n, m, nt; };

struct Stats { long items = 0, branched = 0, cherries = 0, ret = 0, arcs = 0, passes = 0, max_items = 0, rules_only = 0; };

static int g_layout = 4;  // element size of the arrays the solver reads: 4 (int32), 2 (int16), 1 (uint8: out-degrees instead of offsets)

// an instance in the layout of element type S (what the kernel's load<S> reads and emit_child<S> writes)
template <class S>
struct ItemS {
  std::vector<S> succ_off, succ_adj, hlabel, tparent, tlabel; int n, m, nt;
  explicit ItemS(int cn, int cm, int cnt) : succ_off(cn + 1), succ_adj(cm + 1), hlabel(cn + 1), tparent(cnt + 1), tlabel(cnt + 1), n(0), m(0), nt(0) {}
  explicit ItemS(const Item& it) : n(it.n), m(it.m), nt(it.nt) {
    typedef TcLayout<S> Lay;
    if (Lay::kDegrees) { for (int v = 0; v < n; ++v) succ_off.push_back((S)(it.succ_off[v + 1] - it.succ_off[v])); }
    else for (int v = 0; v <= n; ++v) succ_off.push_back((S)it.succ_off[v]);
    for (int e = 0; e < m; ++e) succ_adj.push_back((S)it.succ_adj[e]);
    for (int v = 0; v < n; ++v) hlabel.push_back(Lay::enc(it.hlabel[v]));
    for (int t = 0; t < nt; ++t) { tparent.push_back(Lay::enc(it.tparent[t])); tlabel.push_back(Lay::enc(it.tlabel[t])); }
  }
};

// the kernel's work loop (tc_persistent_kernel): children 1 .. d-1 of a branching go to the pool (here: a stack) as compact
// instances of element type S, child 0 continues IN PLACE on the same workspace
template <class S>
static int solve_instance_t(const Item& root, int capN, int capM, int capNT, int capL, int perm, Stats& St, int* status) {
  using I = int16_t;
  std::vector<char> mem(TcWork<I>::bytes(capN, capM, capNT, capL) + 64);
  std::vector<ItemS<S>> stack; stack.emplace_back(root);
  bool first = true; long items = 0; int verdict = 0; *status = 0;
  while (!stack.empty() && verdict == 0) {
    ItemS<S> it = std::move(stack.back()); stack.pop_back();
    ++items;
    HostEx ex{perm};
    TcWork<I> w; w.carve(mem.data(), capN, capM, capNT, capL);
    TcSolver<HostEx, I> sv(ex, w);
    if (getenv("SIM_NO_CORE")) sv.allow_core_ = false;   // always branch (both ways of finishing are tested)
    TcInst in{it.n, it.m, it.nt, it.succ_off.data(), it.succ_adj.data(), it.hlabel.data(), it.tparent.data(), it.tlabel.data()};
    int st = 0, x = -1;
    int code = sv.template begin<S>(in, !first, &st);
    first = false;
    while (code == TC_CONTINUE) {
      code = sv.rule_loop(&st, &x);
      if (code != TC_BRANCH) break;
      if (getenv("SIM_DUMP_STUCK")) {   // the state the rules could not reduce (for looking at what a further rule would have to see)
        int live_n = 0, retis = 0; std::vector<int> ind(sv.n, 0), outd(sv.n, 0);
        for (int e = 0; e < sv.m; ++e) if (w.at[e] >= 0) { ++outd[w.at[e]]; ++ind[w.ah[e]]; }
        for (int v = 0; v < sv.n; ++v) { live_n += (ind[v] + outd[v]) > 0; retis += ind[v] > 1; }
        int gl = 0; for (int t = 0; t < sv.nt; ++t) gl += w.tlab[t] >= 0 && w.tpar[t] != TcSolver<HostEx, I>::TDEAD;
        long P = 1; for (int v = 0; v < sv.n; ++v) if (ind[v] > 1) P *= ind[v];
        fprintf(stderr, "STUCK items=%ld live_nodes=%d retis=%d switchings=%ld labels=%d branch_on=%d indeg=%d\n  host:", items, live_n, retis, P, sv.L, x, sv.indeg(x));
        for (int e = 0; e < sv.m; ++e) if (w.at[e] >= 0) fprintf(stderr, " %d>%d", (int)w.at[e], (int)w.ah[e]);
        fprintf(stderr, "\n  hlab:"); for (int v = 0; v < sv.n; ++v) if (w.hlab[v] >= 0 && ind[v] + outd[v] > 0) fprintf(stderr, " %d=%d", v, (int)w.hlab[v]);
        fprintf(stderr, "\n  guest:"); for (int t = 0; t < sv.nt; ++t) if (w.tpar[t] != TcSolver<HostEx, I>::TDEAD) fprintf(stderr, " %d<%d", t, (int)w.tpar[t]);
        fprintf(stderr, "\n  tlab:"); for (int t = 0; t < sv.nt; ++t) if (w.tlab[t] >= 0 && w.tpar[t] != TcSolver<HostEx, I>::TDEAD) fprintf(stderr, " %d=%d", t, (int)w.tlab[t]);
        fprintf(stderr, "\n");
      }
      { const int core = sv.finish_small_core(); if (core >= 0) { code = core ? TC_DISPLAYED : TC_NOT; break; } }
      const int d = sv.indeg(x);
      std::vector<int> arcs; for (int k = 0; k < d; ++k) arcs.push_back(sv.in_arc(x, k));
      for (int k = 1; k < d; ++k) {
        ItemS<S> ch(sv.n, sv.m, sv.nt);
        int32_t dims[4];
        TcEmitT<S> out{ch.succ_off.data(), ch.succ_adj.data(), ch.hlabel.data(), ch.tparent.data(), ch.tlabel.data(), dims};
        sv.emit_child(x, arcs[k], out);
        ch.n = dims[0]; ch.m = dims[1]; ch.nt = dims[2];
        stack.push_back(std::move(ch));
      }
      sv.keep_only(x, arcs[0]);
      ++items;
      code = TC_CONTINUE;
    }
    St.cherries += sv.st.cherries; St.ret += sv.st.retcherries; St.arcs += sv.st.arcs_deleted; St.passes += sv.st.passes;
    if (code == TC_ERROR) { *status = st; verdict = -1; break; }
    if (code == TC_DISPLAYED) { verdict = 1; break; }
  }
  St.items += items; if (items > 1) St.branched++; else St.rules_only++;
  if (items > St.max_items) St.max_items = items;
  return verdict;
}
static int solve_instance(const Item& root, int capN, int capM, int capNT, int capL, int perm, Stats& S, int* status) {
  if (g_layout == 1) return solve_instance_t<uint8_t>(root, capN, capM, capNT, capL, perm, S, status);
  if (g_layout == 2) return solve_instance_t<int16_t>(root, capN, capM, capNT, capL, perm, S, status);
  return solve_instance_t<int32_t>(root, capN, capM, capNT, capL, perm, S, status);
}

int main(int argc, char** argv) {
  if (argc < 2) return 1;
  if (const char* e = getenv("SIM_LAYOUT")) g_layout = atoi(e);  // 4 / 2 / 1: element size of the arrays the solver reads
  PT::BatchPacker P; int perm = 0;
  if (!strcmp(argv[1], "gen") && argc >= 8) {
    const long count = atol(argv[2]); const int l = atoi(argv[3]), r = atoi(argv[4]); const uint64_t seed = strtoull(argv[5], 0, 0); const int kind = atoi(argv[6]); perm = atoi(argv[7]);
    for (long i = 0; i < count; ++i) P.add_generated(l, r, seed + (uint64_t)i, kind);
  } else if (!strcmp(argv[1], "file") && argc >= 4) {
    std::ifstream in(argv[2]); std::string a, b; perm = atoi(argv[3]);
    while (std::getline(in, a) && std::getline(in, b)) P.add_newick(a, b);
  } else if (!strcmp(argv[1], "arrays") && argc >= 4) {
    // text format, one instance per block: "n m nt" / m lines "tail head" (sorted by tail) / n host labels / nt guest parents / nt guest labels
    std::ifstream in(argv[2]); perm = atoi(argv[3]);
    long n, m, nt; Stats S2; std::string out2;
    while (in >> n >> m >> nt) {
      Item it; it.n = (int)n; it.m = (int)m; it.nt = (int)nt;
      it.succ_off.assign(n + 1, 0); it.succ_adj.resize(m); it.hlabel.resize(n); it.tparent.resize(nt); it.tlabel.resize(nt);
      for (long e = 0; e < m; ++e) { long a, b; in >> a >> b; it.succ_off[a + 1]++; it.succ_adj[e] = (int32_t)b; }
      for (long v = 0; v < n; ++v) it.succ_off[v + 1] += it.succ_off[v];
      int maxl = 0;
      for (auto& x : it.hlabel) { in >> x; maxl = std::max(maxl, x + 1); }
      for (auto& x : it.tparent) in >> x;
      for (auto& x : it.tlabel) { in >> x; maxl = std::max(maxl, x + 1); }
      int st = 0;
      const int v = solve_instance(it, (int)n, (int)std::max<long>(m, 1), (int)nt, std::max(maxl, 1), perm, S2, &st);
      out2.push_back(v < 0 ? (char)('A' + st) : (char)('0' + v));
    }
    printf("%s\n", out2.c_str());
    return 0;
  } else return 1;
  Stats S; std::string out;
  for (size_t i = 0; i < P.size(); ++i) {
    Item it; const int64_t hb = P.host_node_off[i], ab = P.host_arc_off[i], gb = P.guest_node_off[i];
    it.n = (int)(P.host_node_off[i + 1] - hb); it.m = (int)(P.host_arc_off[i + 1] - ab); it.nt = (int)(P.guest_node_off[i + 1] - gb);
    it.succ_off.assign(P.host_succ_off.begin() + hb + (int64_t)i, P.host_succ_off.begin() + hb + (int64_t)i + it.n + 1);
    it.succ_adj.assign(P.host_succ_adj.begin() + ab, P.host_succ_adj.begin() + ab + it.m);
    it.hlabel.assign(P.host_label.begin() + hb, P.host_label.begin() + hb + it.n);
    it.tparent.assign(P.guest_parent.begin() + gb, P.guest_parent.begin() + gb + it.nt);
    it.tlabel.assign(P.guest_label.begin() + gb, P.guest_label.begin() + gb + it.nt);
    int st = 0;
    const int v = solve_instance(it, P.max_host_nodes, P.max_host_arcs, P.max_guest_nodes, P.max_labels > 0 ? P.max_labels : 1, perm, S, &st);
    out.push_back(v < 0 ? (char)('A' + st) : (char)('0' + v));
  }
  printf("%s\n", out.c_str());
  fprintf(stderr, "{\"instances\": %zu, \"items\": %ld, \"rules_only\": %ld, \"branched\": %ld, \"max_items\": %ld, \"cherries\": %ld, \"ret_cherries\": %ld, \"arcs_deleted\": %ld, \"passes\": %ld}\n",
          P.size(), S.items, S.rules_only, S.branched, S.max_items, S.cherries, S.ret, S.arcs, S.passes);
  return 0;
}
