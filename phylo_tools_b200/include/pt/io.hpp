// pt/io.hpp -- file readers.  Drop-in for io/edgelist.hpp:19-86 (EdgeVecParser, parse_edgelist) and
// io/io.hpp:12-62 (EdgesAndNodeLabels, read_edges, read_edgelists).
#pragma once
#include <fstream>
#include <istream>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>
#include "network.hpp"
#include "newick.hpp"

namespace PT {

// io/edgelist.hpp:41-72: white-space separated name pairs, one arc per line; ids by first appearance;
// EVERY node is labelled with its name.
template <class EdgeList, class LabelMap>
size_t parse_edgelist(std::istream& in, EdgeList& el, LabelMap& names) {
  names.clear(); el.clear();
  std::unordered_map<std::string, Node> ids;
  auto get_id = [&](const std::string& name) {
    auto it = ids.find(name);
    if (it != ids.end()) return it->second;
    const Node id = (Node)ids.size();
    ids.emplace(name, id);
    std::string copy = name;
    detail::put_label(names, id, std::move(copy), 0);
    return id;
  };
  while (!in.eof()) {
    std::string a, b;
    in >> a;
    if (in.bad() || in.eof() || in.fail() || !std::isblank(in.peek())) throw MalformedEdgeVec();
    const Node u = get_id(a);
    in >> b;
    if (in.bad() || (!in.eof() && in.fail()) || !std::isspace(in.peek())) throw MalformedEdgeVec();
    const Node v = get_id(b);
    el.emplace_back(u, v);
    while (in.peek() == 10) in.get();
  }
  return ids.size();
}
template <class EdgeList, class LabelMap>
size_t parse_edgelist(std::istream& in, EdgeList& el, std::shared_ptr<LabelMap>& names) { return parse_edgelist(in, el, *names); }

// io/io.hpp:12-23
template <class Network, class LabelMap = typename Network::LabelMap, class _EdgeVec = EdgeVec>
struct EdgesAndNodeLabels {
  _EdgeVec edges;
  std::shared_ptr<LabelMap> labels = std::make_shared<LabelMap>();
  size_t num_nodes = 0;
  void from_newick(std::istream& in) { num_nodes = parse_newick(in, edges, *labels); }
  void from_edgelist(std::istream& in) { num_nodes = parse_edgelist(in, edges, *labels); }
  void clear() { edges.clear(); labels->clear(); }
  bool is_tree() const { return edges.size() == num_nodes - 1; }
};

// io/io.hpp:26-43: Newick first, then the whole stream as an edge list
template <class ENL>
bool read_edges(std::ifstream& in, ENL& el) {
  try { el.from_newick(in); }
  catch (const MalformedNewick&) {
    try { in.clear(); in.seekg(0); el.clear(); el.from_edgelist(in); }
    catch (const MalformedEdgeVec&) { return false; }
  }
  return true;
}
// io/io.hpp:52-62: one network per line
template <class ENL>
void read_edgelists(const std::string& filename, std::vector<ENL>& out) {
  std::ifstream in(filename);
  while (in.good() && !in.eof()) {
    out.emplace_back();
    if (!read_edges(in, out.back()) || out.back().num_nodes == 0) out.pop_back();
    while (std::isspace(in.peek())) in.get();
  }
}
template <class ENL>
void read_edgelists(const std::vector<std::string>& filenames, std::vector<ENL>& out) { for (const auto& f : filenames) read_edgelists(f, out); }

inline bool file_exists(const std::string& fn) { return std::ifstream(fn).good(); }

}  // namespace PT
