// pt/newick.hpp -- extended-Newick reader and writer.
//
// Drop-in for io/newick.hpp of the reference: parse_newick(string|istream, EdgeList&, LabelMap&) -> num nodes
// (io/newick.hpp:274-290) and get_extended_newick(N) (io/newick.hpp:29-55).
//
// Behaviour kept identical to the reference reader, because node ids are caller-visible:
//   * the text is consumed BACK TO FRONT; a node gets the next free id when its name token is met, so ids are
//     a preorder of the right-to-left spanning tree and the root is 0            (newick.hpp:59-61,164)
//   * a name containing '#' is a hybrid: the integer after the first digit following the last '#' keys it; the
//     first occurrence (from the right) creates the node and keeps the text before '#' as its label, later
//     occurrences only add an in-arc                                              (newick.hpp:172-182,235-247)
//   * the arc parent->child is appended after the child's whole subtree           (newick.hpp:226-232)
//   * ":<len>" is accepted and dropped; white space is skipped around tokens only (newick.hpp:141-144,250-258)
//   * a repeated arc is an error ("read double edge")                             (newick.hpp:216-217)
// Differences, on purpose: the reader is iterative (an explicit stack instead of recursion, so a 10^6-node
// caterpillar parses), it does not print, and a one-node tree "a;" keeps its label (the reference drops it
// through an off-by-one in read_name, newick.hpp:262-270).
#pragma once
#include <cctype>
#include <cstdlib>
#include <istream>
#include <memory>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include "types.hpp"

namespace PT {

namespace detail {
// LabelMap adaptors: anything with emplace(Node, string) works (std::unordered_map, std::map, ...);
// std::vector<std::string> is indexed by node.
template <class LabelMap>
inline auto put_label(LabelMap& m, Node u, std::string&& s, int) -> decltype(m.emplace(u, std::move(s)), void()) { m.emplace(u, std::move(s)); }
inline void put_label(std::vector<std::string>& m, Node u, std::string&& s, long) { if (m.size() <= u) m.resize(u + 1); m[u] = std::move(s); }
}  // namespace detail

template <class EL, class LabelMap>
class NewickParser {
  const std::string& text;
  EL& edges;
  LabelMap& names;
  long back;
  size_t node_count = 0;
  struct Hyb { Index node; Degree extra_in; };
  std::unordered_map<long, Hyb> hybrids;
  std::unordered_set<uint64_t> arcs_seen;

  void skip_ws() { while (back >= 0 && std::isspace((unsigned char)text[(size_t)back])) --back; }
  [[noreturn]] void fail(const std::string& what) const { throw MalformedNewick(text, back, what); }
  char cur() const { if (back < 0) throw MalformedNewick(text, back, "unexpected start of string"); return text[(size_t)back]; }

  // ":<number>" directly left of the cursor?  (drop it)
  void skip_length() {
    long i = back;
    while (i >= 0) { const char c = text[(size_t)i]; if ((c >= '0' && c <= '9') || c == '.' || c == '-' || c == '+' || c == 'E' || c == 'e') --i; else break; }
    if (i >= 0 && text[(size_t)i] == ':') back = i - 1;
  }
  // the token between the previous one of "()," and the cursor
  std::string take_name() {
    long i = back;
    while (i >= 0) { const char c = text[(size_t)i]; if (c == '(' || c == ')' || c == ',') break; --i; }
    std::string name = text.substr((size_t)(i + 1), (size_t)(back - i));
    back = i;
    return name;
  }
  Index resolve(std::string&& name) {
    const size_t sharp = name.rfind('#');
    if (sharp != std::string::npos) {
      const size_t digit = name.find_first_of("0123456789", sharp);
      if (digit == std::string::npos) fail("found '#' but no hybrid number: \"" + name + "\"");
      const long key = std::atol(name.c_str() + digit);
      auto it = hybrids.find(key);
      if (it != hybrids.end()) { ++it->second.extra_in; return it->second.node; }
      const Index id = node_count++;
      hybrids.emplace(key, Hyb{id, 0});
      name.resize(sharp);
      if (!name.empty()) detail::put_label(names, (Node)id, std::move(name), 0);
      return id;
    }
    const Index id = node_count++;
    if (!name.empty()) detail::put_label(names, (Node)id, std::move(name), 0);
    return id;
  }

  void run() {
    skip_ws();
    if (back < 0) return;
    if (text[(size_t)back] != ';') fail("expected ';'");
    --back;
    std::vector<Index> open;  // nodes whose "(...)" is being read
    bool want_subtree = true;
    Index done = 0;
    for (;;) {
      if (want_subtree) {
        skip_ws();
        const Index id = resolve(take_name());
        if (back > 0 && text[(size_t)back] == ')') { --back; open.push_back(id); skip_length(); continue; }
        done = id; want_subtree = false;
      }
      skip_ws();
      if (open.empty()) break;
      const Index parent = open.back();
      if (!arcs_seen.insert(((uint64_t)parent << 32) | (uint64_t)done).second)
        fail("read double edge " + std::to_string(parent) + " --> " + std::to_string(done));
      edges.emplace_back((Node)parent, (Node)done);
      const char c = cur();
      if (c == ',') { --back; skip_length(); want_subtree = true; }
      else if (c == '(') { --back; open.pop_back(); done = parent; }
      else fail(std::string("expected ',' or '(' but got '") + c + "'");
    }
  }

 public:
  NewickParser(const std::string& _text, EL& _edges, LabelMap& _names)
      : text(_text), edges(_edges), names(_names), back((long)_text.length() - 1) { run(); }
  bool is_tree() const { return hybrids.empty(); }
  size_t num_nodes() const { return node_count; }
};

template <class EdgeList, class LabelMap>
size_t parse_newick(const std::string& in, EdgeList& el, LabelMap& names) { return NewickParser<EdgeList, LabelMap>(in, el, names).num_nodes(); }
template <class EdgeList, class LabelMap>
size_t parse_newick(const std::string& in, EdgeList& el, std::shared_ptr<LabelMap>& names) { return parse_newick(in, el, *names); }
template <class EdgeList, class LabelMap>
size_t parse_newick(std::istream& in, EdgeList& el, LabelMap& names) { std::string line; std::getline(in, line); return parse_newick(line, el, names); }

// get_extended_newick (io/newick.hpp:29-55): "(children)label#H<nodeid>" at the first visit of a reticulation,
// "label#H<nodeid>" afterwards.  Iterative.  N needs root(), children(u), in_degree(u), label(u).
template <class Net>
std::string get_extended_newick(const Net& N) {
  if (N.num_nodes() == 0) return ";";
  std::string out;
  std::vector<char> seen(N.num_nodes(), 0);
  struct Fr { Node u; size_t next; bool expanded; };
  std::vector<Fr> st;
  auto close = [&](Node u) { out += N.label(u); if (N.in_degree(u) > 1) { out += "#H" + std::to_string(u); seen[u] = 1; } };
  st.push_back({N.root(), 0, false});
  while (!st.empty()) {
    Fr& f = st.back();
    const auto& ch = N.children(f.u);
    if (f.next == 0 && !f.expanded) {
      f.expanded = true;
      if (ch.empty() || (N.in_degree(f.u) > 1 && seen[f.u])) { const Node u = f.u; st.pop_back(); close(u); if (!st.empty()) out += ","; continue; }
      out += "(";
    }
    if (f.next < ch.size()) { const Node c = ch[f.next++]; st.push_back({c, 0, false}); continue; }
    out.pop_back();  // trailing ','
    out += ")";
    const Node u = f.u; st.pop_back(); close(u);
    if (!st.empty()) out += ",";
  }
  return out + ";";
}

}  // namespace PT
