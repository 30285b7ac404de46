// pt/lca.hpp -- LCA over the CUDA engine.  Drop-in names for
//   Tree::LCA(x, y) / naiveLCA        utils/tree.hpp:202-223
//   lca<node_t>::lca(root), query()   rmq/lca.hpp:79-106
// plus the batched query the engine is built for.
#pragma once
#include <stdexcept>
#include <vector>
#include "../../../include/pt_gpu.h"
#include "containment.hpp"
#include "network.hpp"

namespace PT {

class LCAOracle {
  pt_lca* h_ = nullptr;
  LCAOracle(const LCAOracle&) = delete;
  LCAOracle& operator=(const LCAOracle&) = delete;

 public:
  // from a parent array (-1 at the root)
  explicit LCAOracle(const std::vector<int32_t>& parent, void* stream = nullptr) { throw_status(pt_lca_build(&h_, parent.data(), (int64_t)parent.size(), PT_MEM_HOST, stream), "pt_lca_build"); }
  // from a tree of the facade
  template <class Tree>
  explicit LCAOracle(const Tree& T, void* stream = nullptr) : LCAOracle(T.parent_array(), stream) {}
  ~LCAOracle() { pt_lca_free(h_); }

  Node query(Node u, Node v) const { const int32_t a = (int32_t)u, b = (int32_t)v; int32_t r = -1; throw_status(pt_lca_query(h_, &a, &b, &r, 1, PT_MEM_HOST, nullptr), "pt_lca_query"); return r < 0 ? NoNode : (Node)r; }
  Node LCA(Node u, Node v) const { return query(u, v); }
  std::vector<int32_t> query(const std::vector<int32_t>& u, const std::vector<int32_t>& v) const {
    if (u.size() != v.size()) throw std::logic_error("LCAOracle::query: size mismatch");
    std::vector<int32_t> out(u.size());
    throw_status(pt_lca_query(h_, u.data(), v.data(), out.data(), (int64_t)u.size(), PT_MEM_HOST, nullptr), "pt_lca_query");
    return out;
  }
  pt_lca_info info() const { pt_lca_info i{}; throw_status(pt_lca_get_info(h_, &i), "pt_lca_get_info"); return i; }
  const pt_lca* handle() const { return h_; }
};

}  // namespace PT
