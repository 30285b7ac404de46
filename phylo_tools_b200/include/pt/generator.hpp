// pt/generator.hpp -- random instance generators.
//
// Seeded restatement of utils/generator.hpp:19-46,175-245 and the pieces of utils/random.hpp:15-44 it uses.
// The reference draws from an UNSEEDED rand() and from the iteration order of a libstdc++ unordered_map
// (get_random_iterator(dangling), dangling.begin(); generator.hpp:207,237), so its exact output is not
// reproducible across platforms; this version draws the same process from a splitmix64 stream:
//   * nodes 1..t+r-1 are attached below a uniformly chosen node that still has a free out-slot
//     ("dangling"; uniform over NODES, not over slots, like get_random_iterator on the map),
//   * node i becomes a reticulation with probability (r - rc)/(t + r - i) when at least two nodes were dangling
//     (throw_bw_die, generator.hpp:214-216); its second parent is uniform among the OTHER dangling nodes,
//   * leaves fill the remaining slots (generator.hpp:235-244) and are named by sequential_taxon_name.
// New (the reference has no such function): random_displayed_tree() = the tree displayed under a
// uniform random switching, cleaned up, children shuffled.
#pragma once
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>
#include "types.hpp"

namespace PT {

struct SplitMix64 {
  uint64_t s;
  explicit SplitMix64(uint64_t seed) : s(seed) {}
  uint64_t next() { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
  uint32_t die(uint32_t sides) { return (uint32_t)(next() % sides); }  // throw_die, utils/random.hpp:15-18
};

// utils/generator.hpp:19-29
inline uint32_t l_from_nr(uint32_t n, uint32_t r) {
  if (n % 2 == 0) throw std::logic_error("binary networks must have an odd number of vertices");
  if (n < 2 * r + 1) throw std::logic_error("need at least " + std::to_string(2 * r + 1) + " nodes in a binary network with " + std::to_string(r) + " reticulations");
  return (n - 2 * r + 1) / 2;
}
inline uint32_t n_from_rl(uint32_t r, uint32_t l) { if (l == 0) throw std::logic_error("networks should have leaves"); return 2 * r + 2 * l - 1; }

// utils/generator.hpp:31-46: a..z, ba, bb, ...
struct sequential_taxon_name {
  uint32_t count = 0;
  operator std::string() { return to_string(count++); }
  std::string operator()(uint32_t x) const { return to_string(x); }
  static std::string to_string(uint32_t x) {
    std::string s;
    for (;;) { s.insert(s.begin(), (char)('a' + x % 26)); if (x < 26) break; x /= 26; }
    return s;
  }
};

// utils/generator.hpp:175-245.  names: anything with emplace(Node, std::string).
template <class EdgeContainer, class NameContainer>
void generate_random_binary_edgelist_trl(EdgeContainer& edges, NameContainer& names, uint32_t num_tree_nodes, uint32_t num_retis,
                                         uint32_t num_leaves, SplitMix64& rng) {
  if (num_leaves == 0) throw std::logic_error("cannot construct network without leaves");
  if (num_tree_nodes == 0) throw std::logic_error("cannot construct network without tree nodes");
  const uint32_t num_internal = num_tree_nodes + num_retis, num_nodes = num_internal + num_leaves;
  if (2 * num_tree_nodes + num_retis != (num_tree_nodes - 1) + 2 * num_retis + num_leaves)
    throw std::logic_error("there is no binary network with " + std::to_string(num_tree_nodes) + " tree nodes, " + std::to_string(num_retis) +
                           " reticulations, and " + std::to_string(num_leaves) + " leaves");
  for (;;) {  // the reference throws when the draw runs out of tree nodes (generator.hpp:225); we redraw instead
    const size_t edges_before = edges.size();
    std::vector<std::pair<uint32_t, uint8_t>> dangling;  // (node, free out-slots)
    dangling.emplace_back(0u, (uint8_t)2);
    auto take = [&](size_t k) { const uint32_t u = dangling[k].first; if (--dangling[k].second == 0) { dangling[k] = dangling.back(); dangling.pop_back(); return std::make_pair(u, true); } return std::make_pair(u, false); };
    uint32_t reti_count = 0, tree_count = 0; bool ok = true;
    for (uint32_t i = 1; i < num_internal && ok; ++i) {
      const size_t unsatisfied = dangling.size();
      const size_t k = rng.die((uint32_t)unsatisfied);
      const auto pr = take(k);
      edges.emplace_back((Node)pr.first, (Node)i);
      if (reti_count < num_retis && unsatisfied > 1 && rng.die(num_internal - i) < num_retis - reti_count) {
        size_t k2;
        if (pr.second) k2 = rng.die((uint32_t)dangling.size());
        else { k2 = rng.die((uint32_t)dangling.size() - 1); if (k2 >= k) ++k2; }
        edges.emplace_back((Node)take(k2).first, (Node)i);
        dangling.emplace_back(i, (uint8_t)1); ++reti_count;
      } else {
        if (tree_count == num_tree_nodes) { ok = false; break; }
        dangling.emplace_back(i, (uint8_t)2); ++tree_count;
      }
    }
    if (ok && reti_count == num_retis) {
      for (uint32_t i = num_internal; i < num_nodes; ++i) {
        if (dangling.empty()) { ok = false; break; }
        edges.emplace_back((Node)take(0).first, (Node)i);
        names.emplace((Node)i, sequential_taxon_name::to_string(i - num_internal));
      }
      if (ok) return;
    }
    edges.resize(edges_before);
    for (uint32_t i = num_internal; i < num_nodes; ++i) names.erase((Node)i);
  }
}
template <class EdgeContainer, class NameContainer>
void generate_random_binary_edgelist_rl(EdgeContainer& edges, NameContainer& names, uint32_t num_retis, uint32_t num_leaves, SplitMix64& rng) {
  const uint32_t n = n_from_rl(num_retis, num_leaves);
  generate_random_binary_edgelist_trl(edges, names, n - num_retis - num_leaves, num_retis, num_leaves, rng);
}
// from the number of nodes and reticulations / nodes and leaves (utils/generator.hpp:247-269; like the reference, the _nl form hands
// its second argument on as the number of RETICULATIONS -- generator.hpp:268)
template <class EdgeContainer, class NameContainer>
void generate_random_binary_edgelist_nr(EdgeContainer& edges, NameContainer& names, uint32_t num_nodes, uint32_t num_retis, SplitMix64& rng) {
  const uint32_t l = l_from_nr(num_nodes, num_retis);
  generate_random_binary_edgelist_trl(edges, names, num_nodes - num_retis - l, num_retis, l, rng);
}
template <class EdgeContainer, class NameContainer>
void generate_random_binary_edgelist_nl(EdgeContainer& edges, NameContainer& names, uint32_t num_nodes, uint32_t num_leaves, SplitMix64& rng) {
  generate_random_binary_edgelist_nr(edges, names, num_nodes, num_leaves, rng);
}


// utils/generator.hpp:50-92: random (not necessarily binary) tree; leaves are 0..l-1 and named, the root is the LAST node
template <class EdgeContainer, class NameContainer>
void generate_random_tree(EdgeContainer& edges, NameContainer& names, uint32_t num_internal, uint32_t num_leaves, SplitMix64& rng) {
  if (num_leaves == 0) throw std::logic_error("cannot construct tree without leaves");
  if (num_internal == 0) throw std::logic_error("cannot construct tree without internal nodes");
  const uint32_t num_nodes = num_leaves + num_internal;
  if (num_leaves + num_internal - 1 < 2 * num_internal)
    throw std::logic_error("there is no tree with " + std::to_string(num_internal) + " internal nodes and " + std::to_string(num_leaves) + " leaves");
  std::vector<uint32_t> free_nodes;
  for (uint32_t u = 0; u < num_leaves; ++u) { free_nodes.push_back(u); names.emplace((Node)u, sequential_taxon_name::to_string(u)); }
  for (uint32_t u = num_leaves; u < num_nodes; ++u) {
    const uint32_t nodes_left = num_nodes - u;
    const uint32_t max_degree = (uint32_t)free_nodes.size() - (nodes_left - 1);
    const uint32_t min_degree = (u == num_nodes - 1) ? max_degree : 2;
    const uint32_t degree = min_degree + rng.die(max_degree - min_degree + 1);
    for (uint32_t i = 0; i < degree; ++i) {
      const size_t k = rng.die((uint32_t)free_nodes.size());
      edges.emplace_back((Node)u, (Node)free_nodes[k]);
      free_nodes[k] = free_nodes.back(); free_nodes.pop_back();
    }
    free_nodes.push_back(u);
  }
}

// utils/generator.hpp:97-171 for the case tc -r uses (new_tree_nodes == new_reticulations == num_edges): every new arc joins
// two fresh nodes that subdivide two distinct random arcs, oriented so that no cycle appears.  n grows by 2 per arc.
template <class EdgeContainer>
void add_random_edges(EdgeContainer& edges, size_t& n, uint32_t num_edges, SplitMix64& rng) {
  if (num_edges == 0) return;
  if (edges.size() < 2) throw std::logic_error("cannot add edges to a tree/network with less than 2 edges");
  auto has_path = [&](Node from, Node to) {
    std::vector<std::vector<Node>> adj(n);
    for (const auto& e : edges) adj[e.first].push_back(e.second);
    std::vector<char> seen(n, 0); std::vector<Node> st{from}; seen[from] = 1;
    while (!st.empty()) { const Node v = st.back(); st.pop_back(); if (v == to) return true; for (Node c : adj[v]) if (!seen[c]) { seen[c] = 1; st.push_back(c); } }
    return false;
  };
  for (uint32_t k = 0; k < num_edges; ++k) {
    const size_t i = rng.die((uint32_t)edges.size()); size_t j = rng.die((uint32_t)edges.size() - 1); if (j >= i) ++j;
    const Node u = edges[i].first, v = edges[i].second, x = edges[j].first, y = edges[j].second;
    Node s = n++, t = n++;
    edges[i] = {u, s}; edges.emplace_back(s, v);   // subdivide (u,v) with s
    edges[j] = {x, t}; edges.emplace_back(t, y);   // subdivide (x,y) with t
    if (has_path(y, u) || y == u) std::swap(s, t);
    edges.emplace_back(s, t);
  }
}

// Tree displayed by a uniform random switching of the network given as an edge list over nodes 0..n-1:
// keep one in-arc per node, drop unlabelled dangling nodes, suppress in=out=1 nodes, drop an out-degree-1 root.
// Output: edges of the tree with nodes renumbered 0..nt-1 in preorder (root 0), children shuffled; labels carried over.
template <class EdgeContainer, class NameContainer>
size_t random_displayed_tree(const EdgeContainer& net_edges, size_t n, const NameContainer& net_names, SplitMix64& rng,
                             EdgeVec& tree_edges, NameContainer& tree_names) {
  std::vector<std::vector<uint32_t>> in(n), kids(n);
  for (const auto& e : net_edges) in[e.second].push_back((uint32_t)e.first);
  size_t root = n;
  for (size_t v = 0; v < n; ++v) {
    if (in[v].empty()) { root = v; continue; }
    const uint32_t p = in[v].size() == 1 ? in[v][0] : in[v][rng.die((uint32_t)in[v].size())];
    kids[p].push_back((uint32_t)v);
  }
  // post-order clean-up: alive = labelled leaf below; collapse unary chains
  std::vector<uint32_t> order; order.reserve(n);
  { std::vector<uint32_t> st{(uint32_t)root}; while (!st.empty()) { uint32_t v = st.back(); st.pop_back(); order.push_back(v); for (uint32_t c : kids[v]) st.push_back(c); } }
  std::vector<uint32_t> rep(n, UINT32_MAX);  // representative of the cleaned subtree rooted at v (UINT32_MAX = dead)
  std::vector<std::vector<uint32_t>> ckids(n);
  for (size_t i = order.size(); i-- > 0;) {
    const uint32_t v = order[i];
    if (kids[v].empty()) { if (net_names.count((Node)v)) rep[v] = v; continue; }
    std::vector<uint32_t> alive;
    for (uint32_t c : kids[v]) if (rep[c] != UINT32_MAX) alive.push_back(rep[c]);
    if (alive.empty()) continue;
    if (alive.size() == 1) { rep[v] = alive[0]; continue; }
    for (size_t a = alive.size(); a > 1; --a) std::swap(alive[a - 1], alive[rng.die((uint32_t)a)]);
    ckids[v] = std::move(alive); rep[v] = v;
  }
  tree_edges.clear();
  if (root == n || rep[root] == UINT32_MAX) return 0;
  size_t next = 0;
  std::vector<std::pair<uint32_t, uint32_t>> st{{rep[root], UINT32_MAX}};  // (old node, new parent)
  // preorder, first child first: push children in reverse
  while (!st.empty()) {
    auto [v, np] = st.back(); st.pop_back();
    const uint32_t id = (uint32_t)next++;
    if (np != UINT32_MAX) tree_edges.emplace_back((Node)np, (Node)id);
    if (ckids[v].empty()) { auto it = net_names.find((Node)v); if (it != net_names.end()) tree_names.emplace((Node)id, it->second); }
    for (size_t k = ckids[v].size(); k-- > 0;) st.push_back({ckids[v][k], id});
  }
  return next;
}

// random rooted binary tree on the given leaf labels: join two uniformly chosen roots until one remains
template <class NameContainer>
size_t random_binary_tree(const std::vector<std::string>& leaf_labels, SplitMix64& rng, EdgeVec& tree_edges, NameContainer& tree_names) {
  const size_t l = leaf_labels.size();
  tree_edges.clear();
  if (l == 0) return 0;
  std::vector<std::vector<uint32_t>> ck(2 * l - 1);
  std::vector<uint32_t> roots(l);
  for (size_t i = 0; i < l; ++i) roots[i] = (uint32_t)i;
  uint32_t next = (uint32_t)l;
  while (roots.size() > 1) {
    size_t i = rng.die((uint32_t)roots.size()); std::swap(roots[i], roots.back()); const uint32_t a = roots.back(); roots.pop_back();
    size_t j = rng.die((uint32_t)roots.size()); const uint32_t b = roots[j];
    ck[next] = {a, b}; roots[j] = next++;
  }
  size_t cnt = 0;
  std::vector<std::pair<uint32_t, uint32_t>> st{{roots[0], UINT32_MAX}};
  while (!st.empty()) {
    auto [v, np] = st.back(); st.pop_back();
    const uint32_t id = (uint32_t)cnt++;
    if (np != UINT32_MAX) tree_edges.emplace_back((Node)np, (Node)id);
    if (v < l) tree_names.emplace((Node)id, leaf_labels[v]);
    for (size_t k = ck[v].size(); k-- > 0;) st.push_back({ck[v][k], id});
  }
  return cnt;
}

}  // namespace PT
