// pt/packer.hpp -- packs (host network, guest tree) instances into the flat int32 layout of include/pt_gpu.h.
//
// This is the host half of the boundary: what the reference does per instance when it copies a parsed edge
// list into ConsecutiveNetworkAdjacencyStorage (utils/storage_adj_immutable.hpp:173-261) and builds a
// LabelMatching (utils/label_matching.hpp:40-66), done once for a whole batch.  Label strings are dictionary-
// encoded per instance (host and guest share the dictionary); "" = unlabelled = -1 (utils/tree.hpp:138-148).
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>
#include "../../../include/pt_gpu.h"
#include "generator.hpp"
#include "network.hpp"
#include "newick.hpp"

namespace PT {

class BatchPacker {
 public:
  std::vector<int64_t> host_node_off{0}, host_arc_off{0}, guest_node_off{0};
  std::vector<int32_t> host_succ_off, host_succ_adj, host_pred_off, host_pred_adj, host_label;
  std::vector<int32_t> guest_parent, guest_child_off, guest_child_adj, guest_label;
  int32_t max_host_nodes = 0, max_host_arcs = 0, max_guest_nodes = 0, max_labels = 0;

  size_t size() const { return host_node_off.size() - 1; }

  // host: any DAG with one root; guest: must be a tree.  Throws std::logic_error like the reference's
  // storage constructors (storage_adj_common.hpp:103-109, storage_adj_immutable.hpp:219).
  template <class LabelMapH, class LabelMapG>
  void add(const EdgeVec& hedges, size_t hn, const LabelMapH& hlabels, const EdgeVec& gedges, size_t gn, const LabelMapG& glabels) {
    Phylogeny<LabelMapH> H(hedges, hlabels, consecutive_tag(), hn);
    Phylogeny<LabelMapG> G(gedges, glabels, consecutive_tag(), gn);
    add(H, G);
  }
  template <class Host, class Guest>
  void add(const Host& H, const Guest& G) {
    if (!G.is_tree()) throw std::logic_error("guest must be a tree");
    for (Node v = 0; v < G.num_nodes(); ++v) if (G.in_degree(v) > 1) throw std::logic_error("guest must be a tree");
    const size_t hn = H.num_nodes(), hm = H.num_edges(), gn = G.num_nodes();
    std::unordered_map<std::string, int32_t> dict;
    auto id_of = [&](const std::string& s) -> int32_t { if (s.empty()) return -1; auto it = dict.find(s); if (it != dict.end()) return it->second; const int32_t id = (int32_t)dict.size(); dict.emplace(s, id); return id; };
    host_succ_off.insert(host_succ_off.end(), H.succ_off().begin(), H.succ_off().end());
    host_succ_adj.insert(host_succ_adj.end(), H.succ_adj().begin(), H.succ_adj().end());
    host_pred_off.insert(host_pred_off.end(), H.pred_off().begin(), H.pred_off().end());
    host_pred_adj.insert(host_pred_adj.end(), H.pred_adj().begin(), H.pred_adj().end());
    for (Node v = 0; v < hn; ++v) host_label.push_back(id_of(H.label(v)));
    const std::vector<int32_t> par = G.parent_array();
    guest_parent.insert(guest_parent.end(), par.begin(), par.end());
    guest_child_off.insert(guest_child_off.end(), G.succ_off().begin(), G.succ_off().end());
    guest_child_adj.insert(guest_child_adj.end(), G.succ_adj().begin(), G.succ_adj().end());
    for (Node v = 0; v < gn; ++v) guest_label.push_back(id_of(G.label(v)));
    host_node_off.push_back(host_node_off.back() + (int64_t)hn);
    host_arc_off.push_back(host_arc_off.back() + (int64_t)hm);
    guest_node_off.push_back(guest_node_off.back() + (int64_t)gn);
    max_host_nodes = std::max<int32_t>(max_host_nodes, (int32_t)hn);
    max_host_arcs = std::max<int32_t>(max_host_arcs, (int32_t)hm);
    max_guest_nodes = std::max<int32_t>(max_guest_nodes, (int32_t)gn);
    max_labels = std::max<int32_t>(max_labels, (int32_t)dict.size());
    std::vector<std::string> names(dict.size());
    for (auto& kv : dict) names[(size_t)kv.second] = kv.first;
    dict_names.push_back(std::move(names));
  }
  // examples/tc.cpp:98-115: two Newick lines; the network is the host, or the first if both are trees
  void add_newick(const std::string& first, const std::string& second) {
    EdgeVec e0, e1; LabelMapDefault l0, l1;
    const size_t n0 = parse_newick(first, e0, l0), n1 = parse_newick(second, e1, l1);
    const bool t0 = e0.size() + 1 == n0, t1 = e1.size() + 1 == n1;
    if (t0 && !t1) add(e1, n1, l1, e0, n0, l0); else add(e0, n0, l0, e1, n1, l1);
  }
  // kind 0: displayed guest; 1: displayed guest with two leaf labels swapped; 2: random guest on the same labels
  void add_generated(int32_t num_leaves, int32_t num_retis, uint64_t seed, int kind) {
    SplitMix64 rng(seed);
    EdgeVec he, ge; LabelMapDefault hl, gl;
    generate_random_binary_edgelist_rl(he, hl, (uint32_t)num_retis, (uint32_t)num_leaves, rng);
    const size_t hn = n_from_rl((uint32_t)num_retis, (uint32_t)num_leaves);
    size_t gn;
    if (kind == 2) { std::vector<std::string> names; for (Node v = 0; v < hn; ++v) { auto it = hl.find(v); if (it != hl.end()) names.push_back(it->second); } gn = random_binary_tree(names, rng, ge, gl); }
    else {
      gn = random_displayed_tree(he, hn, hl, rng, ge, gl);
      if (kind == 1 && gl.size() >= 2) {
        std::vector<Node> leaves; for (auto& p : gl) leaves.push_back(p.first);
        std::sort(leaves.begin(), leaves.end());
        const size_t a = rng.die((uint32_t)leaves.size()); size_t b = rng.die((uint32_t)leaves.size() - 1); if (b >= a) ++b;
        std::swap(gl[leaves[a]], gl[leaves[b]]);
      }
    }
    add(he, hn, hl, ge, gn, gl);
  }

  pt_batch_desc desc() const {
    pt_batch_desc d{};
    d.num_instances = (int64_t)size(); d.mem = PT_MEM_HOST;
    d.host_node_off = host_node_off.data(); d.host_arc_off = host_arc_off.data(); d.guest_node_off = guest_node_off.data();
    d.host_succ_off = host_succ_off.data(); d.host_succ_adj = host_succ_adj.data(); d.host_label = host_label.data();
    d.guest_parent = guest_parent.data(); d.guest_label = guest_label.data();
    d.host_pred_off = host_pred_off.data(); d.host_pred_adj = host_pred_adj.data();
    d.guest_child_off = guest_child_off.data(); d.guest_child_adj = guest_child_adj.data();
    d.max_host_nodes = max_host_nodes; d.max_host_arcs = max_host_arcs; d.max_guest_nodes = max_guest_nodes; d.max_labels = max_labels;
    return d;
  }

  // the same batch with int16 data arrays (pt_batch_desc.index_bytes = 2): half the bytes to upload
  bool compact_ok() const { return max_host_nodes < 32768 && max_host_arcs < 32768 && max_guest_nodes < 32768 && max_labels < 32768; }
  pt_batch_desc desc_compact() {
    if (!compact_ok()) throw std::logic_error("an instance is too large for 16-bit ids");
    auto narrow = [](const std::vector<int32_t>& a, std::vector<int16_t>& b) { b.resize(a.size()); for (size_t i = 0; i < a.size(); ++i) b[i] = (int16_t)a[i]; };
    narrow(host_succ_off, c_succ_off); narrow(host_succ_adj, c_succ_adj); narrow(host_label, c_host_label);
    narrow(guest_parent, c_guest_parent); narrow(guest_label, c_guest_label);
    pt_batch_desc d = desc();
    d.host_succ_off = reinterpret_cast<const int32_t*>(c_succ_off.data()); d.host_succ_adj = reinterpret_cast<const int32_t*>(c_succ_adj.data());
    d.host_label = reinterpret_cast<const int32_t*>(c_host_label.data()); d.guest_parent = reinterpret_cast<const int32_t*>(c_guest_parent.data());
    d.guest_label = reinterpret_cast<const int32_t*>(c_guest_label.data());
    d.host_pred_off = d.host_pred_adj = d.guest_child_off = d.guest_child_adj = nullptr;
    d.index_bytes = 2;
    return d;
  }
  std::vector<int16_t> c_succ_off, c_succ_adj, c_host_label, c_guest_parent, c_guest_label;

  // the same batch with uint8 data arrays (pt_batch_desc.index_bytes = 1): a quarter of the bytes to upload.  host_succ_off
  // then holds the n OUT-DEGREES of an instance (at host_node_off[i], no "+ i") instead of n + 1 offsets -- arc positions can
  // exceed 255 (config 2: m = 258) while every node id, label id and degree still fits a byte; 255 stands for -1.
  bool bytes_ok() const { return max_host_nodes <= 255 && max_guest_nodes <= 255 && max_labels <= 255; }
  pt_batch_desc desc_bytes() {
    if (!bytes_ok()) throw std::logic_error("an instance is too large for 8-bit ids");
    const size_t B = size();
    b_degree.resize((size_t)host_node_off[B]);
    for (size_t i = 0; i < B; ++i)
      for (int64_t v = 0, n = host_node_off[i + 1] - host_node_off[i]; v < n; ++v) {
        const int32_t* off = &host_succ_off[(size_t)(host_node_off[i] + (int64_t)i + v)];
        if (off[1] - off[0] > 255) throw std::logic_error("an out-degree is too large for the 8-bit layout");
        b_degree[(size_t)(host_node_off[i] + v)] = (uint8_t)(off[1] - off[0]);
      }
    auto narrow = [](const std::vector<int32_t>& a, std::vector<uint8_t>& b) { b.resize(a.size()); for (size_t i = 0; i < a.size(); ++i) b[i] = (uint8_t)(a[i] < 0 ? 255 : a[i]); };
    narrow(host_succ_adj, b_succ_adj); narrow(host_label, b_host_label); narrow(guest_parent, b_guest_parent); narrow(guest_label, b_guest_label);
    pt_batch_desc d = desc();
    d.host_succ_off = reinterpret_cast<const int32_t*>(b_degree.data()); d.host_succ_adj = reinterpret_cast<const int32_t*>(b_succ_adj.data());
    d.host_label = reinterpret_cast<const int32_t*>(b_host_label.data()); d.guest_parent = reinterpret_cast<const int32_t*>(b_guest_parent.data());
    d.guest_label = reinterpret_cast<const int32_t*>(b_guest_label.data());
    d.host_pred_off = d.host_pred_adj = d.guest_child_off = d.guest_child_adj = nullptr;
    d.index_bytes = 1;
    return d;
  }
  std::vector<uint8_t> b_degree, b_succ_adj, b_host_label, b_guest_parent, b_guest_label;

  std::vector<std::vector<std::string>> dict_names;  // per instance: label id -> string

  // extended Newick (host, guest) of instance i, rebuilt from the flat arrays
  std::pair<std::string, std::string> newick_of(size_t i) const {
    if (dict_names[i].empty() && max_labels > 0) { bool any = false; for (int64_t v = host_node_off[i]; v < host_node_off[i + 1]; ++v) any |= host_label[(size_t)v] >= 0; if (any) throw std::logic_error("label strings were not kept for this instance (batch reader)"); }
    const int64_t hb = host_node_off[i], hn = host_node_off[i + 1] - hb, ab = host_arc_off[i];
    const int64_t gb = guest_node_off[i], gn = guest_node_off[i + 1] - gb;
    EdgeVec he, ge; LabelMapDefault hl, gl;
    for (int64_t v = 0; v < hn; ++v) {
      const int32_t* off = &host_succ_off[(size_t)(hb + (int64_t)i + v)];
      for (int32_t k = off[1]; k-- > off[0];) he.emplace_back((Node)v, (Node)host_succ_adj[(size_t)(ab + k)]);
      const int32_t l = host_label[(size_t)(hb + v)]; if (l >= 0) hl.emplace((Node)v, dict_names[i][(size_t)l]);
    }
    for (int64_t v = 0; v < gn; ++v) {
      const int32_t* off = &guest_child_off[(size_t)(gb + (int64_t)i + v)];
      for (int32_t k = off[1]; k-- > off[0];) ge.emplace_back((Node)v, (Node)guest_child_adj[(size_t)(gb - (int64_t)i + k)]);
      const int32_t l = guest_label[(size_t)(gb + v)]; if (l >= 0) gl.emplace((Node)v, dict_names[i][(size_t)l]);
    }
    Phylogeny<> H(he, hl, consecutive_tag(), (size_t)hn), G(ge, gl, consecutive_tag(), (size_t)gn);
    return {get_extended_newick(H), get_extended_newick(G)};
  }
};

}  // namespace PT
