// pt/bridges.hpp -- bridges and the bridge-separated components of a network, over pt_net_bridges.  Drop-in names for
//   list_bridges(N, root) / BridgeFinder                 utils/bridges.hpp:42-128   (Network::get_bridges utils/network.hpp:85-88)
//   BiconnectedComponents / BiconnectedComponentIter     utils/biconnected_comps.hpp:15-160
// The device computes, for every arc, whether it is a bridge and, for every node, the top node of its bridge-free component;
// this header only puts them into the reference's ORDER: bridges in post-order of the depth-first search that follows
// children() from the root (bridges.hpp:73-103), components in the order the reference iterator emits them (the component
// hanging below each bridge, then -- when trivial components are enumerated -- the bridge itself, the root component last).
#pragma once
#include <stdexcept>
#include <vector>
#include "../../../include/pt_gpu.h"
#include "containment.hpp"
#include "network.hpp"

namespace PT {

struct BridgeAnalysis {
  std::vector<uint8_t> is_bridge;   // per arc, CSR position (children(u) order)
  std::vector<int32_t> component;   // per node: top node of its bridge-free component
  template <class Net>
  explicit BridgeAnalysis(const Net& N, void* stream = nullptr) : is_bridge(N.num_edges() + 1, 0), component(N.num_nodes() + 1, -1) {
    throw_status(pt_net_bridges((int64_t)N.num_nodes(), (int64_t)N.num_edges(), N.succ_off().data(), N.succ_adj().data(), is_bridge.data(), component.data(),
                                PT_MEM_HOST, stream), "pt_net_bridges");
  }
};

// bridges of N in post-order (the order of BridgeFinder::bridge_collector_DFS)
template <class Net>
std::vector<Edge> list_bridges(const Net& N, const BridgeAnalysis& A) {
  std::vector<Edge> out;
  const size_t n = N.num_nodes();
  if (n == 0) return out;
  const auto& off = N.succ_off(); const auto& adj = N.succ_adj();
  std::vector<uint8_t> seen(n, 0);
  struct Frame { Node v; int32_t next; int32_t via; };  // via = CSR position of the arc we came in by (-1 at the root)
  std::vector<Frame> st{{N.root(), off[N.root()], -1}};
  seen[N.root()] = 1;
  while (!st.empty()) {
    Frame& f = st.back();
    if (f.next < off[f.v + 1]) {
      const int32_t e = f.next++;
      const Node c = (Node)adj[e];
      if (!seen[c]) { seen[c] = 1; st.push_back({c, off[c], e}); }
    } else {
      const Frame done = f;
      st.pop_back();
      if (done.via >= 0 && A.is_bridge[(size_t)done.via]) out.emplace_back(st.back().v, done.v);
    }
  }
  return out;
}
template <class Net>
std::vector<Edge> list_bridges(const Net& N) { return list_bridges(N, BridgeAnalysis(N)); }

// the components the reference iterator yields, as edge lists (biconnected_comps.hpp:34-96)
template <class Net>
class BiconnectedComponents {
  std::vector<EdgeVec> comps_;

 public:
  explicit BiconnectedComponents(const Net& N, bool enumerate_trivial = true) {
    const BridgeAnalysis A(N);
    const auto bridges = list_bridges(N, A);
    const auto& off = N.succ_off(); const auto& adj = N.succ_adj();
    // non-bridge arcs grouped by the component of their tail
    std::vector<EdgeVec> by_top(N.num_nodes());
    for (Node u = 0; u < N.num_nodes(); ++u)
      for (int32_t e = off[u]; e < off[u + 1]; ++e)
        if (!A.is_bridge[(size_t)e]) by_top[(size_t)A.component[u]].emplace_back(u, (Node)adj[e]);
    for (const auto& uv : bridges) {
      EdgeVec& below = by_top[(size_t)A.component[uv.second]];
      if (!below.empty()) { comps_.push_back(std::move(below)); below.clear(); }
      if (enumerate_trivial) comps_.push_back(EdgeVec{uv});
    }
    if (N.num_nodes() > 0) comps_.push_back(std::move(by_top[(size_t)A.component[N.root()]]));  // the root component (may be empty, as in the reference)
  }
  auto begin() const { return comps_.begin(); }
  auto end() const { return comps_.end(); }
  size_t size() const { return comps_.size(); }
  const EdgeVec& operator[](size_t i) const { return comps_[i]; }
};

}  // namespace PT
