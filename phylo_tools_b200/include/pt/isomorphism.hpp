// pt/isomorphism.hpp -- network isomorphism over pt_iso_batch.  Drop-in names for
//   make_iso_mapper(N0, N1, flags), IsomorphismMapper::check_isomorph()   utils/isomorphism.hpp:33-324
//   FLAG_MAP_LEAF_LABELS / _TREE_LABELS / _RETI_LABELS / _ALL_LABELS       utils/isomorphism.hpp:10-13
// plus BatchIsomorphism for many pairs per call.  Labels are compared as strings: the two networks of a pair get a common
// dictionary here (like LabelMatchingFromNets, utils/label_matching.hpp).
#pragma once
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>
#include "../../../include/pt_gpu.h"
#include "containment.hpp"
#include "network.hpp"

#define FLAG_MAP_LEAF_LABELS 0x01
#define FLAG_MAP_TREE_LABELS 0x02
#define FLAG_MAP_RETI_LABELS 0x04
#define FLAG_MAP_ALL_LABELS 0x07

namespace PT {

// two extended-Newick lines per pair, parsed AND compared on the GPU (pt_iso_from_newick): the `iso --batch` route
class NewickTextIsomorphism {
  std::vector<uint8_t> verdict_, status_;

 public:
  explicit NewickTextIsomorphism(const std::string& text, unsigned char flags = FLAG_MAP_LEAF_LABELS, void* stream = nullptr) {
    const size_t cap = text.size() / 4 + 2;
    std::vector<uint32_t> bits((cap + 31) / 32 + 1, 0);
    status_.assign(cap + 1, 0);
    int64_t pairs = 0;
    throw_status(pt_iso_from_newick(text.data(), (int64_t)text.size(), PT_MEM_HOST, flags, bits.data(), status_.data(), PT_MEM_HOST, &pairs, stream), "pt_iso_from_newick");
    status_.resize((size_t)pairs);
    verdict_.resize((size_t)pairs);
    for (size_t i = 0; i < (size_t)pairs; ++i) verdict_[i] = (bits[i >> 5] >> (i & 31)) & 1u;
  }
  size_t size() const { return verdict_.size(); }
  bool isomorph(size_t i) const { return verdict_.at(i) != 0; }
  int status(size_t i) const { return status_.at(i); }
};

class BatchIsomorphism {
  std::vector<int64_t> node_off_{0}, arc_off_{0};
  std::vector<int32_t> succ_off_, succ_adj_, pred_off_, pred_adj_, label_;
  std::vector<uint8_t> verdict_, status_;
  int32_t max_nodes_ = 0;
  bool done_ = false;
  template <class Net>
  void append(const Net& N, std::unordered_map<std::string, int32_t>& dict) {
    succ_off_.insert(succ_off_.end(), N.succ_off().begin(), N.succ_off().end());
    succ_adj_.insert(succ_adj_.end(), N.succ_adj().begin(), N.succ_adj().end());
    pred_off_.insert(pred_off_.end(), N.pred_off().begin(), N.pred_off().end());
    pred_adj_.insert(pred_adj_.end(), N.pred_adj().begin(), N.pred_adj().end());
    for (Node v = 0; v < N.num_nodes(); ++v) {
      const std::string& s = N.label(v);
      label_.push_back(s.empty() ? -1 : dict.emplace(s, (int32_t)dict.size()).first->second);
    }
    node_off_.push_back(node_off_.back() + (int64_t)N.num_nodes());
    arc_off_.push_back(arc_off_.back() + (int64_t)N.num_edges());
    if ((int32_t)N.num_nodes() > max_nodes_) max_nodes_ = (int32_t)N.num_nodes();
  }

 public:
  template <class NetA, class NetB>
  size_t add(const NetA& N0, const NetB& N1) {
    std::unordered_map<std::string, int32_t> dict;
    append(N0, dict); append(N1, dict);
    done_ = false;
    return size() - 1;
  }
  size_t size() const { return (node_off_.size() - 1) / 2; }
  void run(unsigned char flags = FLAG_MAP_LEAF_LABELS, void* stream = nullptr) {
    const size_t P = size();
    std::vector<uint32_t> bits((P + 31) / 32 + 1, 0);
    status_.assign(P + 1, 0);
    pt_iso_desc d{};
    d.num_pairs = (int64_t)P; d.mem = PT_MEM_HOST; d.node_off = node_off_.data(); d.arc_off = arc_off_.data();
    static const int32_t zero = 0;  // empty batches still need non-null arrays
    d.succ_off = succ_off_.empty() ? &zero : succ_off_.data(); d.succ_adj = succ_adj_.empty() ? &zero : succ_adj_.data();
    d.pred_off = pred_off_.empty() ? &zero : pred_off_.data(); d.pred_adj = pred_adj_.empty() ? &zero : pred_adj_.data();
    d.label = label_.empty() ? &zero : label_.data();
    d.flags = flags; d.max_nodes = max_nodes_;
    throw_status(pt_iso_batch(&d, bits.data(), status_.data(), PT_MEM_HOST, stream), "pt_iso_batch");
    verdict_.resize(P);
    for (size_t i = 0; i < P; ++i) verdict_[i] = (bits[i >> 5] >> (i & 31)) & 1u;
    done_ = true;
  }
  bool isomorph(size_t i) {
    if (!done_) run();
    if (status_.at(i) != PT_INST_OK) throw std::logic_error(status_[i] == PT_INST_TOO_LARGE ? "network too large for the isomorphism kernel" : "malformed network");
    return verdict_.at(i);
  }
};

// make_iso_mapper(N0, N1, flags).check_isomorph()  (utils/isomorphism.hpp:309-321, :195-240)
template <class NetA, class NetB>
class IsomorphismMapper {
  const NetA& N0; const NetB& N1; const unsigned char flags;
 public:
  IsomorphismMapper(const NetA& a, const NetB& b, unsigned char f) : N0(a), N1(b), flags(f) {}
  bool check_isomorph() const { BatchIsomorphism b; b.add(N0, N1); b.run(flags); return b.isomorph(0); }
};
template <class NetA, class NetB>
IsomorphismMapper<NetA, NetB> make_iso_mapper(const NetA& N0, const NetB& N1, const unsigned char flags) { return IsomorphismMapper<NetA, NetB>(N0, N1, flags); }

}  // namespace PT
