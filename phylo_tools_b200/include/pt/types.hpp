// pt/types.hpp -- basic vocabulary of the facade.
// Mirrors the names of the reference's utils/types.hpp:58-68 (Node, NoNode, Degree) and utils/edge.hpp:57-183
// (Edge<> with tail()/head(), EdgeVec) so that code written against igel-kun/phylo_tools reads the same.
#pragma once
#include <cstdint>
#include <exception>
#include <limits>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace PT {

using Node = uintptr_t;                                   // utils/types.hpp:58
constexpr Node NoNode = std::numeric_limits<Node>::max(); // utils/types.hpp:59
using Degree = uint_fast32_t;                             // utils/types.hpp:68
using Index = uintptr_t;                                  // io/newick.hpp:11

// Edge<> = pair<tail, head>  (utils/edge.hpp:57-143)
struct Edge : public std::pair<Node, Node> {
  using std::pair<Node, Node>::pair;
  Edge() : std::pair<Node, Node>(NoNode, NoNode) {}
  Node tail() const { return first; }
  Node head() const { return second; }
  std::pair<Node, Node> as_pair() const { return {first, second}; }
};
using EdgeVec = std::vector<Edge>;  // utils/edge.hpp:183

struct consecutive_tag {};      // nodes are 0..n-1 (utils/storage_adj_common.hpp)
struct non_consecutive_tag {};

// io/newick.hpp:15-26
struct MalformedNewick : public std::exception {
  const long pos;
  const std::string msg;
  MalformedNewick(const std::string&, long _pos, const std::string& _msg = "unknown error")
      : pos(_pos), msg(_msg + " (position " + std::to_string(_pos) + ")") {}
  const char* what() const noexcept override { return msg.c_str(); }
};
// io/edgelist.hpp:10-15
struct MalformedEdgeVec : public std::exception {
  const char* what() const noexcept override { return "error reading edgelist"; }
};

}  // namespace PT
