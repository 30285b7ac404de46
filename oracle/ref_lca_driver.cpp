// oracle/ref_lca_driver.cpp -- TEST INFRASTRUCTURE ONLY.
// Drives the UNMODIFIED reference rmq/ sub-project (rmq/lca.hpp, rmq/sparse_rmq.hpp, rmq/pm_rmq.hpp,
// included from /root/reference at build time; nothing is copied) to produce LCA / RMQ answers for
// parity tests and as the "reference" CPU baseline of the LCA leg.
//
//   ref_lca lca   < in  : "n q" then n parents (-1 = root), then q pairs "u v"   -> q lines of LCA node ids
//   ref_lca rmq   < in  : "n q" then n ints, then q pairs "l r" (half-open [l,r)) -> q lines "idx value" (sparse_rmq)
//   ref_lca pmrmq < in  : same, array must be +-1                                  -> q lines "idx value" (pm_rmq)
//   ref_lca bench_lca n q seed : random binary tree with n leaves, q random queries; prints build/query seconds
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <string>
#include <vector>
#include <cstdint>
#include "lca.hpp"
// sparse_rmq.hpp has no include guard; it arrives via pm_rmq.hpp
#include "pm_rmq.hpp"
#include "tree.hpp"

using T = tree<int>;

// build the reference's recursive tree<int> bottom-up (children vectors are moved in)
static T build(int v, const std::vector<std::vector<int>>& ch) {
  // iterative post-order to avoid deep recursion on the build side
  struct Fr { int v; size_t i; std::vector<T> kids; };
  std::vector<Fr> st; st.push_back({v, 0, {}});
  while (true) {
    Fr& f = st.back();
    if (f.i < ch[f.v].size()) { int c = ch[f.v][f.i++]; st.push_back({c, 0, {}}); continue; }
    T node = f.kids.empty() ? T((unsigned)f.v, f.v) : T((unsigned)f.v, f.v, std::move(f.kids));
    st.pop_back();
    if (st.empty()) return node;
    st.back().kids.push_back(std::move(node));
  }
}
static void index_nodes(const T& t, std::vector<const T*>& by_id) {
  std::vector<const T*> st{&t};
  while (!st.empty()) { const T* x = st.back(); st.pop_back(); by_id[x->get_id()] = x; for (auto it = x->cbegin(); it != x->cend(); ++it) st.push_back(&*it); }
}
static uint64_t sm64(uint64_t& s) { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }

int main(int argc, char** argv) {
  std::string mode = argc > 1 ? argv[1] : "";
  if (mode == "lca") {
    long n, q; if (scanf("%ld %ld", &n, &q) != 2) return 2;
    std::vector<std::vector<int>> ch(n); int root = -1;
    for (long i = 0; i < n; ++i) { long p; if (scanf("%ld", &p) != 1) return 2; if (p < 0) root = (int)i; else ch[p].push_back((int)i); }
    T t = build(root, ch); std::vector<const T*> by_id(n); index_nodes(t, by_id);
    lca<T> L(t, (unsigned)n);
    for (long i = 0; i < q; ++i) { long u, v; if (scanf("%ld %ld", &u, &v) != 2) return 2; printf("%u\n", L.query(*by_id[u], *by_id[v])->get_id()); }
    return 0;
  }
  if (mode == "rmq" || mode == "pmrmq") {
    long n, q; if (scanf("%ld %ld", &n, &q) != 2) return 2;
    std::vector<int> a(n); for (auto& x : a) if (scanf("%d", &x) != 1) return 2;
    using It = std::vector<int>::const_iterator;
    std::unique_ptr<rmq<It>> R;
    if (mode == "rmq") R.reset(new sparse_rmq<It>(a.begin(), a.end())); else R.reset(new pm_rmq<It>(a.begin(), a.end()));
    for (long i = 0; i < q; ++i) { long l, r; if (scanf("%ld %ld", &l, &r) != 2) return 2; auto idx = R->query_offset(l, r); printf("%ld %d\n", (long)idx, a[idx]); }
    return 0;
  }
  if (mode == "bench_lca") {
    long leaves = atol(argv[2]), q = atol(argv[3]); uint64_t seed = strtoull(argv[4], 0, 0);
    long n = 2 * leaves - 1; std::vector<std::vector<int>> ch(n);
    // same process as phylo_tools_b200 gen: join two uniformly chosen roots until one remains
    std::vector<int> roots(leaves); for (long i = 0; i < leaves; ++i) roots[i] = (int)i;
    long next = leaves, nr = leaves;
    while (nr > 1) { long i = sm64(seed) % nr; std::swap(roots[i], roots[nr - 1]); long j = sm64(seed) % (nr - 1); int a = roots[nr - 1], b = roots[j]; ch[next] = {a, b}; roots[j] = (int)next++; --nr; }
    auto t0 = std::chrono::steady_clock::now();
    T t = build(roots[0], ch); std::vector<const T*> by_id(n); index_nodes(t, by_id);
    lca<T> L(t, (unsigned)n);
    auto t1 = std::chrono::steady_clock::now();
    uint64_t acc = 0;
    for (long i = 0; i < q; ++i) { long u = sm64(seed) % n, v = sm64(seed) % n; acc += L.query(*by_id[u], *by_id[v])->get_id(); }
    auto t2 = std::chrono::steady_clock::now();
    printf("{\"leaves\": %ld, \"queries\": %ld, \"build_s\": %.6f, \"query_s\": %.6f, \"checksum\": %llu}\n", leaves, q,
           std::chrono::duration<double>(t1 - t0).count(), std::chrono::duration<double>(t2 - t1).count(), (unsigned long long)acc);
    return 0;
  }
  fprintf(stderr, "usage: ref_lca lca|rmq|pmrmq|bench_lca\n"); return 1;
}
