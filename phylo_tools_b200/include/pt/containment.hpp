// pt/containment.hpp -- the containment entry points of the reference, over the CUDA engine.
//
// Drop-in names for utils/containment.hpp:
//   TreeInNetContainment<Host,Guest>(host, guest)   ctor overloads :1179-1182 (copy or move; we only read)
//       .displayed()                                 :1202
//       .force_parent(u, v)                          :1186-1199 (keep only the in-arc v->u of u)
//   TreeInTreeContainment<Host,Guest>(host, guest)  :50-54, .displayed() :83
// plus BatchContainment, the batched form the reference lacks: the engine's unit of work is a BATCH of instances
// (one warp per instance), so single-instance calls are batches of one.
//   TreeInTreeContainment::who_displays(u)          :72-81 (single-labelled tree hosts; all guest nodes in one GPU pass)
//
// Errors follow the reference: duplicate labels and malformed graphs throw std::logic_error
// (utils/label_matching.hpp:51-52, utils/storage_adj_common.hpp:103-109); engine failures throw std::runtime_error.
// There is no CPU fallback: without a CUDA device displayed() throws.
#pragma once
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

// a tree host is a network without reticulations: same engine (the true-cherry rule alone decides it).
// who_displays(u) (containment.hpp:72-81): host nodes that minimally display the guest subtree below u, sorted by host
// preorder -- for a single-labelled host at most one node.  Computed for all guest nodes at once (pt_tree_who_displays).
template <class Host, class Guest>
class TreeInTreeContainment : public TreeInNetContainment<Host, Guest> {
  std::vector<int32_t> table_, hda_;
  void fill_table() {
    const Host& H = this->host(); const Guest& G = this->guest();
    std::unordered_map<std::string, int32_t> dict;
    auto id_of = [&](const std::string& s) -> int32_t { if (s.empty()) return -1; auto it = dict.find(s); if (it != dict.end()) return it->second; const int32_t id = (int32_t)dict.size(); dict.emplace(s, id); return id; };
    std::vector<int32_t> hp = H.parent_array(), gp = G.parent_array(), hl(H.num_nodes()), gl(G.num_nodes());
    for (Node v = 0; v < H.num_nodes(); ++v) hl[v] = id_of(H.label(v));
    for (Node v = 0; v < G.num_nodes(); ++v) gl[v] = id_of(G.label(v));
    table_.assign(G.num_nodes(), -1);
    throw_status(pt_tree_who_displays(hp.data(), hl.data(), (int64_t)hp.size(), gp.data(), gl.data(), (int64_t)gp.size(), table_.data(), PT_MEM_HOST, nullptr),
                 "pt_tree_who_displays");
  }

 public:
  using TreeInNetContainment<Host, Guest>::TreeInNetContainment;
  std::vector<Node> who_displays(const Node u) {
    if (table_.empty()) fill_table();
    std::vector<Node> r;
    if (table_.at(u) >= 0) r.push_back((Node)table_[u]);
    return r;
  }
  // TreeInComponent::highest_displayed_ancestor(v) (containment.hpp:328-344) with this host as the component: the highest
  // ancestor of v the climb reaches; all guest nodes are answered by one pt_tree_highest_displayed call and cached
  Node highest_displayed_ancestor(const Node v) {
    if (table_.empty()) fill_table();
    if (hda_.empty()) {
      const std::vector<int32_t> gp = this->guest().parent_array();
      hda_.assign(gp.size(), -1);
      throw_status(pt_tree_highest_displayed(gp.data(), table_.data(), (int64_t)gp.size(), (int32_t)this->host().root(), hda_.data(), PT_MEM_HOST, nullptr),
                   "pt_tree_highest_displayed");
    }
    return (Node)hda_.at(v);
  }
};
template <class Host, class Guest> TreeInTreeContainment(const Host&, const Guest&) -> TreeInTreeContainment<Host, Guest>;
template <class Host, class Guest> TreeInTreeContainment(Host&&, Guest&&) -> TreeInTreeContainment<std::decay_t<Host>, std::decay_t<Guest>>;

}  // namespace PT
