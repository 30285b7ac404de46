// pt/lca.hpp -- LCA over the CUDA engine.  Drop-in names for
//   Tree::LCA(x, y) / naiveLCA        utils/tree.hpp:202-223
//   lca<node_t>::lca(root), query()   rmq/lca.hpp:79-106
// plus the batched query the engine is built for.
#pragma once
#include <iterator>
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
  // LCAs of consecutive entries of a preorder-sorted candidate list (utils/induced_tree.hpp:128-138): out[i] = LCA(nodes[i], nodes[i+1])
  std::vector<int32_t> query_adjacent(const std::vector<int32_t>& nodes) const {
    std::vector<int32_t> out(nodes.size() > 1 ? nodes.size() - 1 : 0);
    if (!out.empty()) throw_status(pt_lca_query_adjacent(h_, nodes.data(), (int64_t)nodes.size(), out.data(), PT_MEM_HOST, nullptr), "pt_lca_query_adjacent");
    return out;
  }
  pt_lca_info info() const { pt_lca_info i{}; throw_status(pt_lca_get_info(h_, &i), "pt_lca_get_info"); return i; }
  const pt_lca* handle() const { return h_; }
};

// the reference's class names (rmq/lca.hpp:28-107, rmq/rmq.hpp:17-75, rmq/sparse_rmq.hpp:19-91) over the same engine
//   lca<Tree> L(T);  L.query(u, v) -> Node                         (rmq/lca.hpp:79, :91)
//   sparse_rmq<It> R(begin, end);  R.query(u, v) -> It,  R.query_offset(u, v) -> index of the RIGHTMOST minimum of [u, v)   (rmq/rmq.hpp:63-72)
template <class Tree>
class lca {
  LCAOracle o_;
 public:
  explicit lca(const Tree& T, size_t = 0) : o_(T) {}
  Node query(Node u, Node v) const { return o_.query(u, v); }
};
template <class It>
class rmq {
 protected:
  It begin_;
  pt_rmq* h_ = nullptr;
 public:
  rmq(It b, It e) : begin_(b) {
    std::vector<int32_t> v;
    for (It i = b; i != e; ++i) v.push_back((int32_t)*i);
    throw_status(pt_rmq_build(&h_, v.data(), (int64_t)v.size(), PT_MEM_HOST, nullptr), "pt_rmq_build");
  }
  rmq(const rmq&) = delete;
  ~rmq() { pt_rmq_free(h_); }
  size_t query_offset(size_t u, size_t v) const { const int64_t l = (int64_t)u, r = (int64_t)v; int64_t o = -1; throw_status(pt_rmq_query(h_, &l, &r, &o, 1, PT_MEM_HOST, nullptr), "pt_rmq_query"); return (size_t)o; }
  It query(size_t u, size_t v) const { It i = begin_; std::advance(i, (long)query_offset(u, v)); return i; }
};
template <class It> using sparse_rmq = rmq<It>;
template <class It> using pm_rmq = rmq<It>;

// Tree::LCA(x, y) / naiveLCA (utils/tree.hpp:202-223) as a call on the tree: LCA(T, x, y); the structure is built per call, so
// use LCAOracle (or lca<Tree>) for more than a handful of queries on the same tree
template <class Tree> Node LCA(const Tree& T, Node x, Node y) { return LCAOracle(T).query(x, y); }
template <class Tree> Node naiveLCA(const Tree& T, Node x, Node y) { return LCAOracle(T).query(x, y); }

}  // namespace PT
