// pt/containment.hpp -- the containment entry points of the reference, over the CUDA engine.
//
// Drop-in names for utils/containment.hpp:
//   TreeInNetContainment<Host,Guest>(host, guest)   ctor overloads :1179-1182 (copy or move; we only read)
//       .displayed()                                 :1202
//       .force_parent(u, v)                          :1186-1199 (keep only the in-arc v->u of u)
//   TreeInTreeContainment<Host,Guest>(host, guest)  :50-54, .displayed() :83
// plus BatchContainment, the batched form the reference lacks: the engine's unit of work is a BATCH of instances
// (one warp per instance), so single-instance calls are batches of one.
//   TreeInTreeContainment::who_displays(u)          :72-81 (multi-labelled tree hosts included; all guest nodes in one GPU pass)
//   TreeInComponent(host, guest, label_match, u).highest_displayed_ancestor(v)   :241-347
//   TreeComponentInfos(N)                           utils/tree_components.hpp:41-234
//
// Errors follow the reference: duplicate labels and malformed graphs throw std::logic_error
// (utils/label_matching.hpp:51-52, utils/storage_adj_common.hpp:103-109); engine failures throw std::runtime_error.
// There is no CPU fallback: without a CUDA device displayed() throws.
#pragma once
#include <algorithm>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>
#include "../../../include/pt_gpu.h"
#include "network.hpp"
#include "packer.hpp"
#include "fast_newick.hpp"

namespace PT {

struct EngineError : public std::runtime_error { int code; EngineError(int c, const std::string& what) : std::runtime_error(what), code(c) {} };

inline void throw_status(int rc, const char* where) {
  if (rc == PT_OK) return;
  const std::string msg = std::string(where) + ": " + pt_last_error();
  if (rc == PT_ERR_INVALID) throw std::logic_error(msg);
  if (rc == PT_ERR_PARSE) throw MalformedNewick(std::string(), 0, msg);   // io/newick.hpp:15-26
  throw EngineError(rc, msg);
}

// many pairs given as TEXT (two extended-Newick lines per instance): parsed on the GPU into a device-resident batch
// (pt_batch_create_from_newick) and solved there; no flat arrays on the host at all
class NewickTextContainment {
  std::vector<uint8_t> status_;
  std::vector<bool> verdict_;
  pt_counters counters_{};

 public:
  explicit NewickTextContainment(const std::string& text, void* stream = nullptr) {
    pt_batch* h = nullptr;
    int64_t pairs = 0;
    throw_status(pt_batch_create_from_newick(&h, text.data(), (int64_t)text.size(), PT_MEM_HOST, &pairs, stream), "pt_batch_create_from_newick");
    std::vector<uint32_t> bits((size_t)(pairs + 31) / 32 + 1, 0);
    status_.assign((size_t)pairs + 1, 0);
    const int rc = pt_batch_contain(h, nullptr, bits.data(), status_.data(), PT_MEM_HOST, &counters_, stream);
    pt_batch_free(h);
    throw_status(rc, "pt_batch_contain");
    verdict_.resize((size_t)pairs);
    for (size_t i = 0; i < (size_t)pairs; ++i) verdict_[i] = (bits[i >> 5] >> (i & 31)) & 1u;
  }
  size_t size() const { return verdict_.size(); }
  bool displayed(size_t i) const { return verdict_.at(i); }
  int status(size_t i) const { return status_.at(i); }
  const pt_counters& counters() const { return counters_; }
};

// many (host, guest) pairs at once
class BatchContainment {
  BatchPacker packer_;
  std::vector<uint8_t> status_;
  std::vector<bool> verdict_;
  pt_counters counters_{};
  bool done_ = false;

 public:
  template <class Host, class Guest>
  size_t add(const Host& N, const Guest& T) { done_ = false; packer_.add(N, T); return packer_.size() - 1; }
  size_t add_newick(const std::string& first, const std::string& second) { done_ = false; packer_.add_newick(first, second); return packer_.size() - 1; }
  // two lines per instance, parsed in parallel (pt/fast_newick.hpp); returns the number of instances added
  size_t add_newick_text(const std::string& text, int threads = 0) { done_ = false; return PT::add_newick_text(packer_, text.data(), text.size(), threads); }
  size_t size() const { return packer_.size(); }
  const BatchPacker& packer() const { return packer_; }

  void run(void* stream = nullptr) {
    const size_t B = packer_.size();
    std::vector<uint32_t> bits((B + 31) / 32 + 1, 0);
    status_.assign(B + 1, 0);
    const pt_batch_desc d = packer_.desc();
    pt_batch* h = nullptr;
    throw_status(pt_batch_create(&h, &d, stream), "pt_batch_create");
    const int rc = pt_batch_contain(h, nullptr, bits.data(), status_.data(), PT_MEM_HOST, &counters_, stream);
    pt_batch_free(h);
    throw_status(rc, "pt_batch_contain");
    verdict_.resize(B);
    for (size_t i = 0; i < B; ++i) verdict_[i] = (bits[i >> 5] >> (i & 31)) & 1u;
    done_ = true;
  }
  // the same on every GPU of the node (pt_batch_contain_multi: contiguous shards, one host thread per device)
  void run_all_gpus(int num_devices = 0) {
    const size_t B = packer_.size();
    std::vector<uint32_t> bits((B + 31) / 32 + 1, 0);
    status_.assign(B + 1, 0);
    const pt_batch_desc d = packer_.desc();
    throw_status(pt_batch_contain_multi(&d, num_devices, nullptr, bits.data(), status_.data(), &counters_), "pt_batch_contain_multi");
    verdict_.resize(B);
    for (size_t i = 0; i < B; ++i) verdict_[i] = (bits[i >> 5] >> (i & 31)) & 1u;
    done_ = true;
  }
  // verdict of instance i; throws like the reference would for a malformed instance
  bool displayed(size_t i) {
    if (!done_) run();
    switch (status_.at(i)) {
      case PT_INST_OK: return verdict_[i];
      case PT_INST_MULTILABEL: throw std::logic_error("single-label map for multi-labeled tree/network");
      case PT_INST_BAD_GRAPH: throw std::logic_error("cannot create tree/network: not a rooted DAG with a single root");
      case PT_INST_TOO_LARGE: throw EngineError(PT_ERR_UNSUPPORTED, "instance exceeds the limits of the on-chip solver (pt_get_limits)");
      default: throw EngineError(PT_ERR_UNSUPPORTED, "branching pool exhausted; raise pt_options.pool_slots");
    }
  }
  uint8_t status(size_t i) { if (!done_) run(); return status_.at(i); }
  const pt_counters& counters() { if (!done_) run(); return counters_; }
};

template <class Host, class Guest>
class TreeInNetContainment {
  Host host_;
  Guest guest_;
  bool have_ = false, result_ = false;

 public:
  TreeInNetContainment(const Host& N, const Guest& T) : host_(N), guest_(T) {}
  TreeInNetContainment(Host&& N, const Guest& T) : host_(std::move(N)), guest_(T) {}
  TreeInNetContainment(const Host& N, Guest&& T) : host_(N), guest_(std::move(T)) {}
  TreeInNetContainment(Host&& N, Guest&& T) : host_(std::move(N)), guest_(std::move(T)) {}

  bool displayed() {
    if (!have_) { BatchContainment b; b.add(host_, guest_); result_ = b.displayed(0); have_ = true; }
    return result_;
  }
  // utils/containment.hpp:1186-1199: make v the only parent of u
  void force_parent(const Node u, const Node v) {
    EdgeVec e;
    bool found = false;
    for (const auto& uv : host_.edges()) { if (uv.head() == u) { if (uv.tail() == v) { found = true; e.push_back(uv); } } else e.push_back(uv); }
    if (!found) throw std::logic_error("force_parent: " + std::to_string(v) + " is not a parent of " + std::to_string(u));
    host_ = Host(e, host_.labels_ptr(), consecutive_tag(), host_.num_nodes());
    have_ = false;
  }
  const Host& host() const { return host_; }
  const Guest& guest() const { return guest_; }
};
template <class Host, class Guest> TreeInNetContainment(const Host&, const Guest&) -> TreeInNetContainment<Host, Guest>;
template <class Host, class Guest> TreeInNetContainment(Host&&, Guest&&) -> TreeInNetContainment<std::decay_t<Host>, std::decay_t<Guest>>;

// ---- vocabulary of the tree-in-tree DP (utils/induced_tree.hpp:11-35, utils/label_matching.hpp:23-82) ---------------------------
struct InducedSubtreeInfo { size_t dist_to_root = 0, order_number = 0; InducedSubtreeInfo(size_t d = 0, size_t o = 0) : dist_to_root(d), order_number(o) {} };
template <class Tree> using InducedSubtreeInfoMap = std::unordered_map<Node, InducedSubtreeInfo>;
// label -> (host nodes, guest nodes)
using LabelMatchingMap = std::unordered_map<std::string, std::pair<std::vector<Node>, std::vector<Node>>>;

// shared by the checkers below: the flat arrays of a host / guest pair with one label dictionary
struct DpArrays {
  std::vector<int32_t> hp, hl, gp, gl;
  template <class Host, class Guest>
  DpArrays(const Host& H, const Guest& G) : hp(H.parent_array()), hl(H.num_nodes()), gp(G.parent_array()), gl(G.num_nodes()) {
    std::unordered_map<std::string, int32_t> dict;
    auto id_of = [&](const std::string& s) -> int32_t { if (s.empty()) return -1; auto it = dict.find(s); if (it != dict.end()) return it->second; const int32_t id = (int32_t)dict.size(); dict.emplace(s, id); return id; };
    for (Node v = 0; v < H.num_nodes(); ++v) hl[v] = id_of(H.label(v));
    for (Node v = 0; v < G.num_nodes(); ++v) gl[v] = id_of(G.label(v));
  }
};

// TreeInTreeContainment (utils/containment.hpp:27-233): the host may be MULTI-LABELLED (the reference's ROMulTree: the label map
// may give the same name to several leaves); the guest is a single-labelled tree.  who_displays(u) (:72-81) = the host nodes that
// minimally display the guest subtree below u, sorted by host preorder; the whole table is computed by ONE pass of the GPU DP
// (pt_who_build) at the first question.  displayed() (:83) = the root's list is not empty.
// The 5-argument constructor of the reference (:50-54) is kept: node infos that are handed in empty are filled (make_node_infos,
// induced_tree.hpp:54-59: depth and preorder number of every host node), a label matching is accepted and not needed.
template <class Host, class Guest>
class TreeInTreeContainment {
 public:
  using NodeList = std::vector<Node>;
  using NodeInfos = InducedSubtreeInfoMap<Host>;
  using LabelMatching = LabelMatchingMap;

 private:
  Host host_;
  Guest guest_;
  std::shared_ptr<NodeInfos> node_infos_;
  std::vector<NodeList> table_;
  std::vector<int32_t> hda_;
  bool have_ = false;

  void fill_table() {
    const DpArrays a(host_, guest_);
    // a host that is declared single-labelled must be: the reference throws when it builds the label matching
    // (utils/label_matching.hpp:51-52,61-62); the guest is always single-labelled
    auto check_single = [](const auto& T, const std::vector<int32_t>& lab) {
      int32_t mx = -1;
      for (const int32_t l : lab) mx = std::max(mx, l);
      std::vector<char> seen((size_t)(mx + 2), 0);
      for (Node v = 0; v < T.num_nodes(); ++v) if (T.is_leaf(v) && lab[v] >= 0) { if (seen[(size_t)lab[v]]) throw std::logic_error("single-label map for multi-labeled tree/network"); seen[(size_t)lab[v]] = 1; }
    };
    if (Host::is_single_labeled) check_single(host_, a.hl);
    check_single(guest_, a.gl);
    pt_who* h = nullptr;
    throw_status(pt_who_build(&h, a.hp.data(), a.hl.data(), (int64_t)a.hp.size(), a.gp.data(), a.gl.data(), (int64_t)a.gp.size(), PT_MEM_HOST, nullptr), "pt_who_build");
    int64_t total = 0;
    pt_who_total(h, &total);
    std::vector<int32_t> off(a.gp.size() + 1), adj((size_t)std::max<int64_t>(total, 1));
    int rc = pt_who_lists(h, off.data(), adj.data(), PT_MEM_HOST, nullptr);
    hda_.assign(a.gp.size(), -1);
    if (rc == PT_OK) rc = pt_who_highest_displayed(h, (int32_t)host_.root(), hda_.data(), PT_MEM_HOST, nullptr);
    pt_who_free(h);
    throw_status(rc, "pt_who_lists");
    table_.assign(a.gp.size(), NodeList());
    for (size_t u = 0; u < a.gp.size(); ++u) table_[u].assign(adj.begin() + off[u], adj.begin() + off[u + 1]);
    have_ = true;
  }
  void fill_node_infos() {
    if (!node_infos_ || !node_infos_->empty()) return;
    pt_lca* L = nullptr;
    const std::vector<int32_t> hp = host_.parent_array();
    throw_status(pt_lca_build(&L, hp.data(), (int64_t)hp.size(), PT_MEM_HOST, nullptr), "pt_lca_build");
    std::vector<int32_t> d(hp.size()), o(hp.size());
    const int rc = pt_lca_node_info(L, d.data(), o.data(), PT_MEM_HOST, nullptr);
    pt_lca_free(L);
    throw_status(rc, "pt_lca_node_info");
    for (size_t v = 0; v < hp.size(); ++v) node_infos_->emplace((Node)v, InducedSubtreeInfo((size_t)d[v], (size_t)o[v]));
  }

 public:
  TreeInTreeContainment(const Host& H, const Guest& G, std::shared_ptr<NodeInfos> node_infos = {}, std::shared_ptr<LabelMatching> = {}, const bool = false)
      : host_(H), guest_(G), node_infos_(std::move(node_infos)) { fill_node_infos(); }
  TreeInTreeContainment(Host&& H, Guest&& G) : host_(std::move(H)), guest_(std::move(G)) {}

  const NodeList& who_displays(const Node u) { if (!have_) fill_table(); return table_.at(u); }
  bool displayed() { return !who_displays(guest_.root()).empty(); }
  // TreeInComponent::highest_displayed_ancestor(v) (:328-344) with this host as the component
  Node highest_displayed_ancestor(const Node v) { if (!have_) fill_table(); return (Node)hda_.at(v); }
  const Host& host() const { return host_; }
  const Guest& guest() const { return guest_; }
};
template <class Host, class Guest> TreeInTreeContainment(const Host&, const Guest&) -> TreeInTreeContainment<Host, Guest>;
template <class Host, class Guest> TreeInTreeContainment(Host&&, Guest&&) -> TreeInTreeContainment<std::decay_t<Host>, std::decay_t<Guest>>;

// TreeInComponent (utils/containment.hpp:241-347): the part of a host NETWORK below node u is unfolded into a multi-labelled tree
// (unzip_retis :262-312 -> pt_net_unzip) and the tree-in-tree DP answers, for a guest node v, the highest ancestor of v that this
// component still displays (:328-344).  subtree() / origin() expose the unfolding: node 0 is u, ids are a preorder.
template <class MulSubtree, class Guest, bool leaf_labels_only = true>
class TreeInComponent {
  std::vector<Node> origin_;   // (declared first: unzip() fills it while subtree_ is being constructed)
  MulSubtree subtree_;
  TreeInTreeContainment<MulSubtree, Guest> display_;

  template <class Host>
  static MulSubtree unzip(const Host& host, const Node u, std::vector<Node>& origin) {
    std::unordered_map<std::string, int32_t> dict; std::vector<std::string> names;
    std::vector<int32_t> lab(host.num_nodes(), -1);
    for (Node v = 0; v < host.num_nodes(); ++v) {
      const std::string& s = host.label(v);
      if (s.empty() || (leaf_labels_only && !host.is_leaf(v))) continue;
      auto it = dict.find(s);
      if (it == dict.end()) { it = dict.emplace(s, (int32_t)names.size()).first; names.push_back(s); }
      lab[v] = it->second;
    }
    int64_t count = 0;
    const int64_t m = (int64_t)host.num_edges();
    throw_status(pt_net_unzip((int64_t)host.num_nodes(), m, host.succ_off().data(), host.succ_adj().data(), lab.data(), (int32_t)u, nullptr, nullptr, nullptr, 0, &count, PT_MEM_HOST, nullptr), "pt_net_unzip");
    std::vector<int32_t> par((size_t)count), ol((size_t)count), org((size_t)count);
    throw_status(pt_net_unzip((int64_t)host.num_nodes(), m, host.succ_off().data(), host.succ_adj().data(), lab.data(), (int32_t)u, par.data(), ol.data(), org.data(), count, &count, PT_MEM_HOST, nullptr), "pt_net_unzip");
    EdgeVec edges; typename MulSubtree::LabelMap labels;
    // children in reverse order, so that the immutable storage (which reverses once more) keeps the unfolding's child order
    for (size_t v = (size_t)count; v-- > 1;) edges.emplace_back((Node)par[v], (Node)v);
    for (size_t v = 0; v < (size_t)count; ++v) if (ol[v] >= 0) labels.emplace((Node)v, names[(size_t)ol[v]]);
    origin.assign(org.begin(), org.end());
    return MulSubtree(edges, labels, consecutive_tag(), (size_t)count);
  }

 public:
  template <class Host, class HG_Label_Matching>
  TreeInComponent(const Host& host, const Guest& guest, const HG_Label_Matching&, const Node u) : subtree_(unzip(host, u, origin_)), display_(subtree_, guest) {}
  template <class Host>
  TreeInComponent(const Host& host, const Guest& guest, const Node u) : subtree_(unzip(host, u, origin_)), display_(subtree_, guest) {}

  Node highest_displayed_ancestor(const Node v) { return display_.highest_displayed_ancestor(v); }
  const std::vector<Node>& who_displays(const Node u) { return display_.who_displays(u); }
  const MulSubtree& subtree() const { return subtree_; }
  Node origin(const Node copy) const { return origin_.at(copy); }
};

// TreeComponentInfos (utils/tree_components.hpp:41-234) of a fresh network: comp_root / visible_leaf per node (NoNode where the
// reference leaves it unset) and the component DAG as an edge list (root above, root below), computed by pt_net_tree_components
struct ComponentData { Node comp_root = NoNode; Node visible_leaf = NoNode; };
template <class Network>
class TreeComponentInfos {
  std::vector<ComponentData> data_;
  EdgeVec dag_;

 public:
  explicit TreeComponentInfos(const Network& N) {
    const int64_t n = (int64_t)N.num_nodes(), m = (int64_t)N.num_edges();
    std::vector<int32_t> comp((size_t)n), vis((size_t)n);
    std::vector<int64_t> e((size_t)std::max<int64_t>(4 * m, 16));
    int64_t cnt = 0;
    throw_status(pt_net_tree_components(n, m, N.succ_off().data(), N.succ_adj().data(), comp.data(), vis.data(), e.data(), (int64_t)e.size(), &cnt, PT_MEM_HOST, nullptr), "pt_net_tree_components");
    if (cnt > (int64_t)e.size()) {
      e.assign((size_t)cnt, 0);
      throw_status(pt_net_tree_components(n, m, N.succ_off().data(), N.succ_adj().data(), comp.data(), vis.data(), e.data(), (int64_t)e.size(), &cnt, PT_MEM_HOST, nullptr), "pt_net_tree_components");
    }
    e.resize((size_t)cnt);
    std::sort(e.begin(), e.end());
    e.erase(std::unique(e.begin(), e.end()), e.end());
    for (const int64_t x : e) dag_.emplace_back((Node)(x >> 32), (Node)(x & 0xffffffff));
    data_.resize((size_t)n);
    for (int64_t v = 0; v < n; ++v) { data_[(size_t)v].comp_root = comp[(size_t)v] < 0 ? NoNode : (Node)comp[(size_t)v]; data_[(size_t)v].visible_leaf = vis[(size_t)v] < 0 ? NoNode : (Node)vis[(size_t)v]; }
  }
  const ComponentData& operator[](const Node v) const { return data_.at(v); }
  Node comp_root(const Node v) const { return data_.at(v).comp_root; }
  Node visible_leaf(const Node v) const { return data_.at(v).visible_leaf; }
  const EdgeVec& comp_DAG_edges() const { return dag_; }
};

// VisibleComponentRule (utils/containment.hpp:603-725) as a query over pt_net_visible_component: which tree component of the host the
// rule would treat, the leaf that sees it, and the guest node treat_comp_root would match it with.  The reference applies the rule inside
// its solver's loop; here the solver needs no such rule (the batch solver is exact without it), so this is the caller-visible answer to
// "which part of the host displays which subtree of the guest".  Labels must be cleaned as TreeInNetContainment::init does (:1110-1125).
template <class Host, class Guest>
class VisibleComponentRule {
  std::vector<uint8_t> elig_;
  Node root_ = NoNode, leaf_ = NoNode, guest_node_ = NoNode;

 public:
  VisibleComponentRule(const Host& host, const Guest& guest, const Node want_root = NoNode, const Node want_leaf = NoNode) : elig_(host.num_nodes(), 0) {
    std::unordered_map<std::string, int32_t> dict;
    auto id_of = [&](const std::string& s) -> int32_t { if (s.empty()) return -1; return dict.emplace(s, (int32_t)dict.size()).first->second; };
    std::vector<int32_t> hl(host.num_nodes(), -1), gl(guest.num_nodes(), -1);
    for (Node v = 0; v < host.num_nodes(); ++v) if (host.is_leaf(v)) hl[v] = id_of(host.label(v));
    for (Node v = 0; v < guest.num_nodes(); ++v) if (guest.is_leaf(v)) gl[v] = id_of(guest.label(v));
    const std::vector<int32_t> gp = guest.parent_array();
    int32_t r = -1, l = -1, g = -1;
    throw_status(pt_net_visible_component((int64_t)host.num_nodes(), (int64_t)host.num_edges(), host.succ_off().data(), host.succ_adj().data(), hl.data(), (int64_t)guest.num_nodes(),
                                          gp.data(), gl.data(), want_root == NoNode ? -1 : (int32_t)want_root, want_leaf == NoNode ? -1 : (int32_t)want_leaf, 0, elig_.data(), &r, &l, &g, nullptr),
                 "pt_net_visible_component");
    if (r >= 0) root_ = (Node)r;
    if (l >= 0) leaf_ = (Node)l;
    if (g >= 0) guest_node_ = (Node)g;
  }
  bool is_half_eligible(const Node u) const { return elig_.at(u) & 1; }        // :652-687
  bool is_eligible(const Node u) const { return elig_.at(u) & 2; }
  Node get_eligible_component_root() const { return root_; }                   // :689-704 (NoNode: the rule does not apply)
  Node visible_leaf() const { return leaf_; }
  Node matched_guest_node() const { return guest_node_; }                      // the v of match_nodes(u, v) in treat_comp_root :623-644
  bool apply() const { return root_ != NoNode && guest_node_ != NoNode; }
};

}  // namespace PT
