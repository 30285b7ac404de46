// pt/network.hpp -- tree / network storage of the facade.
//
// One class over an immutable CSR adjacency replaces the reference's storage family on the containment path:
//   ConsecutiveTreeAdjacencyStorage / ConsecutiveNetworkAdjacencyStorage   utils/storage_adj_immutable.hpp:63-261
//   RootedAdjacencyStorage query API (children/parents/in_degree/out_degree/root/leaves/nodes/edges)
//                                                                           utils/storage_adj_common.hpp:118-184
//   _Tree / Network and the aliases ROTree, RWTree, RONetwork, RWNetwork, Compatible*   utils/tree.hpp:619-676,
//                                                                           utils/network.hpp:11-285
// The reference's *mutable* hash-map storage (utils/storage_adj_mutable.hpp) exists only so that its reduction
// rules can rewrite the graph one node at a time; here the rewriting happens inside the CUDA kernels on flat
// arrays, so the host side never mutates and the RW aliases map to the same immutable class.
//
// Layout: succ_off[n+1] / succ_adj[m] and pred_off[n+1] / pred_adj[m] as int32 -- exactly the arrays handed to
// pt_batch_create (include/pt_gpu.h).  Children of a node are stored in REVERSE edge-list order like the
// reference does (storage_adj_immutable.hpp:119,215 fill with --deg).
#pragma once
#include <algorithm>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>
#include "types.hpp"

namespace PT {

using LabelMapDefault = std::unordered_map<Node, std::string>;

// a read-only view of a slice of int32 ids that behaves like a container of Node
class NodeSlice {
  const int32_t* b_; const int32_t* e_;
 public:
  struct iterator {
    const int32_t* p;
    Node operator*() const { return (Node)*p; }
    iterator& operator++() { ++p; return *this; }
    bool operator!=(const iterator& o) const { return p != o.p; }
    bool operator==(const iterator& o) const { return p == o.p; }
  };
  NodeSlice(const int32_t* b, const int32_t* e) : b_(b), e_(e) {}
  iterator begin() const { return {b_}; }
  iterator end() const { return {e_}; }
  size_t size() const { return (size_t)(e_ - b_); }
  bool empty() const { return b_ == e_; }
  Node operator[](size_t i) const { return (Node)b_[i]; }
  Node front() const { return (Node)*b_; }
};

// _MultiLabel: the reference's multi_label_tag (utils/tree.hpp:641-645, ROMulTree / RWMulTree): several leaves may carry one name
template <class _LabelMap = LabelMapDefault, bool _MultiLabel = false>
class Phylogeny {
 public:
  using LabelMap = _LabelMap;
  static constexpr bool is_single_labeled = !_MultiLabel;
  using LabelType = std::string;
  using Edge = PT::Edge;

 protected:
  size_t n_ = 0, m_ = 0;
  Node root_ = NoNode;
  std::vector<int32_t> succ_off_, succ_adj_, pred_off_, pred_adj_;
  std::vector<Node> leaves_;
  std::shared_ptr<LabelMap> labels_;
  static const std::string& empty_label() { static const std::string e; return e; }

  template <class EdgeContainer>
  void build(const EdgeContainer& edges, size_t num_nodes_hint) {
    size_t n = num_nodes_hint;
    for (const auto& e : edges) n = std::max<size_t>(n, (size_t)std::max(e.first, e.second) + 1);
    if (n == 0 && edges.empty()) n = labels_ && !labels_->empty() ? 1 : 0;
    n_ = n; m_ = edges.size();
    succ_off_.assign(n + 1, 0); pred_off_.assign(n + 1, 0);
    for (const auto& e : edges) { succ_off_[e.first + 1]++; pred_off_[e.second + 1]++; }
    for (size_t v = 0; v < n; ++v) { succ_off_[v + 1] += succ_off_[v]; pred_off_[v + 1] += pred_off_[v]; }
    succ_adj_.resize(m_); pred_adj_.resize(m_);
    std::vector<int32_t> sfill(succ_off_.begin() + 1, succ_off_.end()), pfill(pred_off_.begin() + 1, pred_off_.end());
    for (const auto& e : edges) { succ_adj_[--sfill[e.first]] = (int32_t)e.second; pred_adj_[--pfill[e.second]] = (int32_t)e.first; }
    size_t roots = 0;
    for (size_t v = 0; v < n; ++v) {
      if (pred_off_[v + 1] == pred_off_[v]) { root_ = v; ++roots; }
      if (succ_off_[v + 1] == succ_off_[v]) leaves_.push_back(v);
    }
    // utils/storage_adj_common.hpp:103-109
    if (n > 0 && roots != 1) throw std::logic_error(roots == 0 ? "cannot create tree/network without a root" : "cannot create tree/network with multiple roots");
  }

 public:
  Phylogeny() : labels_(std::make_shared<LabelMap>()) {}
  // ctor shapes of utils/tree.hpp:387-411: (edges, labels[, tag])
  template <class EdgeContainer, class Tag = consecutive_tag>
  Phylogeny(const EdgeContainer& edges, std::shared_ptr<LabelMap> labels, Tag = Tag(), size_t num_nodes = 0) : labels_(std::move(labels)) { build(edges, num_nodes); }
  template <class EdgeContainer, class Tag = consecutive_tag>
  Phylogeny(const EdgeContainer& edges, const LabelMap& labels, Tag = Tag(), size_t num_nodes = 0) : labels_(std::make_shared<LabelMap>(labels)) { build(edges, num_nodes); }

  // utils/storage_adj_common.hpp:118-184
  size_t num_nodes() const { return n_; }
  size_t num_edges() const { return m_; }
  bool edgeless() const { return m_ == 0; }
  Node root() const { return root_; }
  NodeSlice children(Node u) const { return {succ_adj_.data() + succ_off_[u], succ_adj_.data() + succ_off_[u + 1]}; }
  NodeSlice parents(Node u) const { return {pred_adj_.data() + pred_off_[u], pred_adj_.data() + pred_off_[u + 1]}; }
  Node parent(Node u) const { return (Node)pred_adj_[pred_off_[u]]; }
  Degree out_degree(Node u) const { return (Degree)(succ_off_[u + 1] - succ_off_[u]); }
  Degree in_degree(Node u) const { return (Degree)(pred_off_[u + 1] - pred_off_[u]); }
  bool is_leaf(Node u) const { return out_degree(u) == 0; }
  bool is_reti(Node u) const { return in_degree(u) > 1; }
  bool is_tree_node(Node u) const { return out_degree(u) > 0 && in_degree(u) <= 1; }
  const std::vector<Node>& leaves() const { return leaves_; }
  std::vector<Node> reticulations() const { std::vector<Node> r; for (Node v = 0; v < n_; ++v) if (is_reti(v)) r.push_back(v); return r; }
  bool is_tree() const { return n_ == 0 || m_ == n_ - 1; }  // io/io.hpp:22
  EdgeVec edges() const { EdgeVec r; r.reserve(m_); for (Node u = 0; u < n_; ++u) for (Node c : children(u)) r.emplace_back(u, c); return r; }
  // utils/tree.hpp:138-148: "" = unlabelled
  const std::string& label(Node u) const { auto it = labels_->find(u); return it == labels_->end() ? empty_label() : it->second; }
  const LabelMap& labels() const { return *labels_; }
  std::shared_ptr<LabelMap> labels_ptr() const { return labels_; }

  // flat arrays for the C-ABI
  const std::vector<int32_t>& succ_off() const { return succ_off_; }
  const std::vector<int32_t>& succ_adj() const { return succ_adj_; }
  const std::vector<int32_t>& pred_off() const { return pred_off_; }
  const std::vector<int32_t>& pred_adj() const { return pred_adj_; }
  // parent array of a tree (-1 at the root)
  std::vector<int32_t> parent_array() const {
    if (!is_tree()) throw std::logic_error("parent_array() on a network with reticulations");
    std::vector<int32_t> p(n_, -1);
    for (Node v = 0; v < n_; ++v) if (in_degree(v)) p[v] = pred_adj_[pred_off_[v]];
    return p;
  }
  const Phylogeny& as_tree() const { return *this; }
  Phylogeny& as_tree() { return *this; }
};

// aliases of utils/tree.hpp:619-676 and utils/network.hpp:262-285
template <class LabelMap = LabelMapDefault> using ROTree = Phylogeny<LabelMap>;
template <class LabelMap = LabelMapDefault> using RWTree = Phylogeny<LabelMap>;
template <class LabelMap = LabelMapDefault> using RONetwork = Phylogeny<LabelMap>;
template <class LabelMap = LabelMapDefault> using RWNetwork = Phylogeny<LabelMap>;
template <class LabelMap = LabelMapDefault> using ROMulTree = Phylogeny<LabelMap, true>;
template <class LabelMap = LabelMapDefault> using RWMulTree = Phylogeny<LabelMap, true>;
template <class T> using CompatibleRWNetwork = Phylogeny<typename T::LabelMap>;
template <class T> using CompatibleROTree = Phylogeny<typename T::LabelMap>;

}  // namespace PT
