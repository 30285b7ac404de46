// tests/host_sim/sim_tc_mt.cpp -- TEST INFRASTRUCTURE ONLY.
// Multi-threaded simulator of the GROUP executors (warp / CTA): T real threads run csrc/tc_core.cuh in SPMD fashion over
// one shared workspace, with a real barrier for ex.sync() and real atomics.  A barrier that is not reached by every thread
// (divergent control flow) deadlocks here exactly as __syncthreads would on the GPU -- a watchdog reports the barrier
// counts per thread -- and ThreadSanitizer sees data races between passes.  Input as sim_tc: "file <path>" or "gen ...".
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "../../include/pt_gpu.h"
#include "../../phylo_tools_b200/csrc/tc_core.cuh"
#include "../../phylo_tools_b200/include/pt/packer.hpp"

using namespace ptgpu;

struct Barrier {
  std::mutex mu; std::condition_variable cv; int n, waiting = 0; long gen = 0; bool broken = false;
  explicit Barrier(int n_) : n(n_) {}
  void wait() {
    std::unique_lock<std::mutex> lk(mu);
    const long g = gen;
    if (++waiting == n) { waiting = 0; ++gen; cv.notify_all(); return; }
    if (!cv.wait_for(lk, std::chrono::seconds(20), [&] { return gen != g || broken; })) { broken = true; cv.notify_all(); fprintf(stderr, "DEADLOCK: barrier generation %ld, %d of %d threads arrived\n", g, waiting, n); std::abort(); }
  }
};

struct Shared { Barrier bar; std::vector<int> scratch; std::atomic<int> red_or{0}, red_add{0}; explicit Shared(int t) : bar(t), scratch(t + 1) {} };

struct GroupEx {
  int t, T; Shared* sh; long nsync = 0; int phase = 0;
  int tid() const { return t; }
  int nth() const { return T; }
  void sync() { ++nsync; sh->bar.wait(); }
  int map(int i, int) const { return i; }
  template <class S> int ldg(const S* p) const { return (int)*p; }
  int uniform(const int32_t* p) { sync(); if (t == 0) sh->scratch[0] = __atomic_load_n(p, __ATOMIC_SEQ_CST); sync(); const int v = sh->scratch[0]; sync(); return v; }
  int reduce_or(int v) { sync(); if (t == 0) sh->red_or = 0; sync(); if (v) sh->red_or.fetch_or(v); sync(); return sh->red_or.load(); }
  int reduce_add(int v) { sync(); if (t == 0) sh->red_add = 0; sync(); if (v) sh->red_add.fetch_add(v); sync(); return sh->red_add.load(); }
  template <class X> X atomic_add(X* p, X v) const { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
  template <class X> X atomic_min(X* p, X v) const { X o = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (v < o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }
  template <class X> X atomic_max(X* p, X v) const { X o = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (v > o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }
  template <class X> X atomic_or(X* p, X v) const { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
  template <class X> X atomic_xor(X* p, X v) const { return __atomic_fetch_xor(p, v, __ATOMIC_SEQ_CST); }
  template <class X> X atomic_cas(X* p, X c, X v) const { __atomic_compare_exchange_n(p, &c, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return c; }
  template <class X> int atomic_cas_idx(X* p, int c, int v) const { X e = (X)c; __atomic_compare_exchange_n(p, &e, (X)v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return (int)e; }
  template <class X> int compact_arcs(X* at, X* ah, int m, X*, X*) {
    sync();
    if (t == 0) { int k = 0; for (int e = 0; e < m; ++e) if (at[e] >= 0) { at[k] = at[e]; ah[k] = ah[e]; ++k; } sh->scratch[0] = k; }
    sync();
    const int k = sh->scratch[0];
    sync();
    return k;
  }
  template <class X> int atomic_add_idx(X* p, int v) const { return (int)__atomic_fetch_add(p, (X)v, __ATOMIC_SEQ_CST); }
  template <class X> void atomic_xor_idx(X* p, int v) const { __atomic_fetch_xor(p, (X)v, __ATOMIC_SEQ_CST); }
  template <class C, class X> void exclusive_scan(const C* in, X* out, int n) { sync(); if (t == 0) { int acc = 0; for (int i = 0; i < n; ++i) { out[i] = (X)acc; acc += in[i]; } out[n] = (X)acc; } sync(); }
};

struct Item { std::vector<int32_t> succ_off, succ_adj, hlabel, tparent, tlabel; int n, m, nt; };

template <class I>
static int solve_instance(const Item& root, int capN, int capM, int capNT, int capL, int T) {
  std::vector<char> mem(TcWork<I>::bytes(capN, capM, capNT, capL) + 64);
  std::vector<Item> stack{root};
  bool first = true; int verdict = 0;
  while (!stack.empty()) {
    Item it = std::move(stack.back()); stack.pop_back();
    Shared sh(T);
    std::vector<int> codes(T), xs(T), sts(T);
    std::vector<Item> children;
    std::vector<std::thread> th;
    std::mutex cm;
    for (int t = 0; t < T; ++t) th.emplace_back([&, t] {
      GroupEx ex{t, T, &sh};
      TcWork<I> w; w.carve(mem.data(), capN, capM, capNT, capL);
      TcSolver<GroupEx, I> sv(ex, w);
      if (getenv("SIM_NO_CORE")) sv.allow_core_ = false;   // always branch (both ways of finishing are tested)
      TcInst in{it.n, it.m, it.nt, it.succ_off.data(), it.succ_adj.data(), it.hlabel.data(), it.tparent.data(), it.tlabel.data()};
      int st = 0, x = -1;
      const int code = sv.template solve<int32_t>(in, !first, &st, &x);
      codes[t] = code; xs[t] = x; sts[t] = st;
      if (code == TC_BRANCH) {
        const int d = sv.indeg(x);
        std::vector<int> arcs; for (int k = 0; k < d; ++k) arcs.push_back(sv.in_arc(x, k));
        ex.sync();
        for (int k = 0; k < d; ++k) {
          if (t == 0) { std::lock_guard<std::mutex> lk(cm); children.emplace_back(); Item& ch = children.back(); ch.succ_off.resize(sv.n + 1); ch.succ_adj.resize(sv.m + 1); ch.hlabel.resize(sv.n + 1); ch.tparent.resize(sv.nt + 1); ch.tlabel.resize(sv.nt + 1); }
          ex.sync();
          Item& ch = children.back();
          static thread_local int32_t dims[4];
          int32_t* dimp = t == 0 ? (int32_t*)&ch.n - 0 : dims;  // thread 0 writes n,m,nt below
          (void)dimp;
          std::vector<int32_t>* dummy = nullptr; (void)dummy;
          static int32_t shared_dims[4];
          TcEmit out{ch.succ_off.data(), ch.succ_adj.data(), ch.hlabel.data(), ch.tparent.data(), ch.tlabel.data(), shared_dims};
          sv.emit_child(x, arcs[k], out);
          if (t == 0) { ch.n = shared_dims[0]; ch.m = shared_dims[1]; ch.nt = shared_dims[2]; }
          ex.sync();
        }
      }
    });
    for (auto& x : th) x.join();
    first = false;
    for (int t = 1; t < T; ++t) if (codes[t] != codes[0] || xs[t] != xs[0]) { fprintf(stderr, "NON-UNIFORM result across threads: code %d vs %d, x %d vs %d\n", codes[0], codes[t], xs[0], xs[t]); std::abort(); }
    if (codes[0] == TC_ERROR) return -1 - sts[0];
    if (codes[0] == TC_DISPLAYED) { verdict = 1; break; }
    for (auto& c : children) stack.push_back(std::move(c));
  }
  return verdict;
}

int main(int argc, char** argv) {
  if (argc < 2) return 1;
  PT::BatchPacker P; int T = 4; bool i32 = false;
  if (!strcmp(argv[1], "gen") && argc >= 8) {
    const long count = atol(argv[2]); const int l = atoi(argv[3]), r = atoi(argv[4]); const uint64_t seed = strtoull(argv[5], 0, 0); const int kind = atoi(argv[6]); T = atoi(argv[7]);
    if (argc > 8) i32 = atoi(argv[8]) != 0;
    for (long i = 0; i < count; ++i) P.add_generated(l, r, seed + (uint64_t)i, kind);
  } else if (!strcmp(argv[1], "file") && argc >= 4) {
    std::ifstream in(argv[2]); std::string a, b; T = atoi(argv[3]); if (argc > 4) i32 = atoi(argv[4]) != 0;
    while (std::getline(in, a) && std::getline(in, b)) P.add_newick(a, b);
  } else return 1;
  std::string out;
  for (size_t i = 0; i < P.size(); ++i) {
    Item it; const int64_t hb = P.host_node_off[i], ab = P.host_arc_off[i], gb = P.guest_node_off[i];
    it.n = (int)(P.host_node_off[i + 1] - hb); it.m = (int)(P.host_arc_off[i + 1] - ab); it.nt = (int)(P.guest_node_off[i + 1] - gb);
    it.succ_off.assign(P.host_succ_off.begin() + hb + (int64_t)i, P.host_succ_off.begin() + hb + (int64_t)i + it.n + 1);
    it.succ_adj.assign(P.host_succ_adj.begin() + ab, P.host_succ_adj.begin() + ab + it.m);
    it.hlabel.assign(P.host_label.begin() + hb, P.host_label.begin() + hb + it.n);
    it.tparent.assign(P.guest_parent.begin() + gb, P.guest_parent.begin() + gb + it.nt);
    it.tlabel.assign(P.guest_label.begin() + gb, P.guest_label.begin() + gb + it.nt);
    const int L = P.max_labels > 0 ? P.max_labels : 1;
    const int v = i32 ? solve_instance<int32_t>(it, P.max_host_nodes, P.max_host_arcs, P.max_guest_nodes, L, T)
                      : solve_instance<int16_t>(it, P.max_host_nodes, P.max_host_arcs, P.max_guest_nodes, L, T);
    out.push_back(v < 0 ? (char)('A' - 1 - v) : (char)('0' + v));
  }
  printf("%s\n", out.c_str());
  return 0;
}
