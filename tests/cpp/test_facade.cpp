// tests/cpp/test_facade.cpp -- exercises the C++ facade (phylo_tools_b200/include/pt/) the way code written against the
// reference's headers would: parse two Newick strings, build MyNet / MyTree like examples/tc.cpp:50-51,110-114, ask
// TreeInNetContainment / TreeInTreeContainment / LCA.  Needs a GPU; exits 0 on success.
#include <cassert>
#include <iostream>
#include <sstream>
#include "../../phylo_tools_b200/include/pt/containment.hpp"
#include "../../phylo_tools_b200/include/pt/io.hpp"
#include "../../phylo_tools_b200/include/pt/lca.hpp"
#include "../../phylo_tools_b200/include/pt/bridges.hpp"
#include "../../phylo_tools_b200/include/pt/isomorphism.hpp"
#include "../../phylo_tools_b200/include/pt/generator.hpp"
#include <algorithm>
#include <fstream>

using namespace PT;
using MyTree = ROTree<>;
using MyNet = CompatibleRWNetwork<MyTree>;

static bool tc(const std::string& host, const std::string& guest) {
  EdgeVec he, ge; auto hl = std::make_shared<LabelMapDefault>(); auto gl = std::make_shared<LabelMapDefault>();
  const size_t hn = parse_newick(host, he, hl), gn = parse_newick(guest, ge, gl);
  MyNet N(he, hl, consecutive_tag(), hn); MyTree T(ge, gl, consecutive_tag(), gn);
  if (N.is_tree()) { TreeInTreeContainment c(N, T); return c.displayed(); }
  TreeInNetContainment c(std::move(N), std::move(T));
  return c.displayed();
}

// `test_facade bcc <file>`: one network per line; prints the bridges in post-order ("B tail head") and the components of
// BiconnectedComponents ("C" + sorted edges), then "--": the format of oracle/_ref/ref_tc bcc, compared by tests/test_cli.py
static int bcc_mode(const char* path) {
  std::ifstream in(path);
  std::string line;
  while (std::getline(in, line)) {
    if (line.empty()) continue;
    EdgeVec e; auto l = std::make_shared<LabelMapDefault>();
    const size_t n = parse_newick(line, e, l);
    RONetwork<> N(e, l, consecutive_tag(), n);
    for (const auto& uv : list_bridges(N)) std::cout << "B " << uv.first << " " << uv.second << "\n";
    for (const auto& comp : BiconnectedComponents<RONetwork<>>(N)) {
      std::vector<std::pair<Node, Node>> ed(comp.begin(), comp.end());
      std::sort(ed.begin(), ed.end());
      std::cout << "C";
      for (const auto& p : ed) std::cout << " " << p.first << "," << p.second;
      std::cout << "\n";
    }
    std::cout << "--\n";
  }
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 3 && std::string(argv[1]) == "bcc") return bcc_mode(argv[2]);
  // generators by (nodes, reticulations) / (reticulations, leaves) (utils/generator.hpp:247-281): BASELINE config 2 is n = 239, r = 20, l = 100
  {
    EdgeVec e1, e2; LabelMapDefault n1, n2; SplitMix64 r1(7), r2(7);
    generate_random_binary_edgelist_nr(e1, n1, 239, 20, r1);
    generate_random_binary_edgelist_rl(e2, n2, 20, 100, r2);
    assert(e1.size() == 258 && n1.size() == 100 && e1.size() == e2.size() && n1.size() == n2.size());
    bool threw = false;
    try { generate_random_binary_edgelist_nr(e1, n1, 240, 20, r1); } catch (const std::logic_error&) { threw = true; }   // an even number of nodes
    assert(threw);
  }
  // data/test.nw of the reference: ground truth "not displayed" (the reference itself crashes on it)
  assert(!tc("((a,(((b)#H1,(c)#H2),(#H1,#H2))),d);", "((a,b),(c,d));"));
  assert(tc("((a,(b)#H1),(#H1,c));", "(a,(b,c));"));
  assert(tc("((a,(b)#H1),(#H1,c));", "((a,b),c);"));
  assert(!tc("((a,(b)#H1),(#H1,c));", "((a,c),b);"));
  assert(tc("((a,b),(c,d));", "((d,c),(b,a));"));
  assert(tc("a;", "a;"));   // one node on both sides: no arc at all (the adjacency array of the batch is empty / null)
  assert(!tc("((a,b),(c,d));", "((a,c),(b,d));"));
  // force_parent (containment.hpp:1186): H1 = node 4 in reader numbering of "((a,(b)#H1),(#H1,c));": edges (1,2)(1,3)... check via parents()
  {
    EdgeVec he, ge; LabelMapDefault hl, gl;
    const size_t hn = parse_newick(std::string("((a,(b)#H1),(#H1,c));"), he, hl), gn = parse_newick(std::string("((a,b),c);"), ge, gl);
    MyNet N(he, hl, consecutive_tag(), hn); MyTree T(ge, gl, consecutive_tag(), gn);
    Node reti = NoNode;
    for (Node v = 0; v < N.num_nodes(); ++v) if (N.is_reti(v)) reti = v;
    assert(reti != NoNode && N.in_degree(reti) == 2);
    int shown = 0;
    for (size_t k = 0; k < 2; ++k) { TreeInNetContainment c(N, T); c.force_parent(reti, N.parents(reti)[k]); shown += c.displayed(); }
    assert(shown == 1);  // exactly one of the two switchings displays ((a,b),c)
  }
  // duplicate labels throw std::logic_error like label_matching.hpp:51-52
  bool threw = false;
  try { tc("((a,a),(c,d));", "((a,b),(c,d));"); } catch (const std::logic_error&) { threw = true; }
  assert(threw);
  // batch form
  BatchContainment b;
  b.add_newick("((a,b),(c,d));", "((a,b),(c,d));");
  b.add_newick("((a,b),(c,d));", "(((a)#H1,b),(#H1,(c,d)));");  // tree first, network second -> network is the host
  b.run();
  assert(b.displayed(0) && b.displayed(1) && b.counters().instances == 2);
  {  // the same two pairs as text, parsed on the GPU
    NewickTextContainment t("((a,b),(c,d));\n((a,b),(c,d));\n\n((a,b),(c,d));\n(((a)#H1,b),(#H1,(c,d)));\n");
    assert(t.size() == 2 && t.displayed(0) && t.displayed(1) && t.status(0) == 0 && t.counters().instances == 2);
    bool bad = false;
    try { NewickTextContainment u("((a,b),(c,d));\n((a,b),(c,d))\n"); } catch (const MalformedNewick&) { bad = true; }
    assert(bad);
  }
  b.run_all_gpus();   // every GPU of the node, one host thread each
  assert(b.displayed(0) && b.displayed(1) && b.counters().instances == 2);
  // LCA: the tree of rmq/lca.cpp:24-43 and its four known answers
  LCAOracle L(std::vector<int32_t>{3, 3, 3, 8, 5, 7, 7, 8, -1});
  assert(L.query(8, 8) == 8 && L.query(3, 7) == 8 && L.query(0, 2) == 3 && L.query(4, 6) == 7);
  { const std::vector<int32_t> adj = L.query_adjacent({0, 2, 3, 7, 4, 6}); assert((adj == std::vector<int32_t>{3, 3, 8, 7, 7})); }
  MyTree T2 = [] { EdgeVec e; LabelMapDefault l; const size_t n = parse_newick(std::string("((a,b),(c,(d,e)));"), e, l); return MyTree(e, l, consecutive_tag(), n); }();
  LCAOracle L2(T2);
  for (Node u = 0; u < T2.num_nodes(); ++u) assert(L2.LCA(u, T2.root()) == T2.root());
  // who_displays (containment.hpp:72): the pairs printed by the reference itself, see tests/golden/who_ref.json items 1-2
  {
    EdgeVec he, ge; LabelMapDefault hl, gl;
    const size_t hn = parse_newick(std::string("((a,b),(c,d));"), he, hl), gn = parse_newick(std::string("((a,c),(b,d));"), ge, gl);
    MyTree H(he, hl, consecutive_tag(), hn), G(ge, gl, consecutive_tag(), gn);
    TreeInTreeContainment c(H, G);
    const int exp[7] = {-1, 0, 2, 5, 0, 3, 6};
    for (Node u = 0; u < 7; ++u) { const auto w = c.who_displays(u); assert(exp[u] < 0 ? w.empty() : (w.size() == 1 && w[0] == (Node)exp[u])); }
    const int hda[7] = {0, 1, 1, 1, 4, 4, 4};   // highest_displayed_ancestor (containment.hpp:328-344), traced by hand
    for (Node u = 0; u < 7; ++u) assert(c.highest_displayed_ancestor(u) == (Node)hda[u]);
    assert(!c.displayed());
  }
  // multi-labelled host: data/test_tree.nw of the reference; the lists printed by the reference itself (tests/golden/whomul_ref.json item 0)
  {
    EdgeVec he, ge; LabelMapDefault hl, gl;
    const size_t hn = parse_newick(std::string("((((a,b),(c,d)),(e,d)),((a,b),(e,d)));"), he, hl), gn = parse_newick(std::string("(e,((a,c),(b,d)));"), ge, gl);
    using MulTree = ROMulTree<>;   // utils/tree.hpp:641-645
    MulTree H(he, hl, consecutive_tag(), hn); MyTree G(ge, gl, consecutive_tag(), gn);
    auto infos = std::make_shared<TreeInTreeContainment<MulTree, MyTree>::NodeInfos>();
    TreeInTreeContainment<MulTree, MyTree> c(H, G, infos);       // the reference's 5-argument constructor (containment.hpp:50-54), defaults for the rest
    assert(infos->size() == hn && infos->at(H.root()).dist_to_root == 0 && infos->at(H.root()).order_number == 0);
    const std::vector<std::vector<Node>> exp = {{}, {0}, {1, 12}, {3, 10, 14}, {6, 17}, {12}, {15}, {7, 18}, {4, 11}};
    for (Node u = 0; u < gn; ++u) { std::vector<Node> w = c.who_displays(u); std::sort(w.begin(), w.end()); assert(w == exp[u]); }
    for (Node u = 0; u < gn; ++u) { const auto& w = c.who_displays(u); for (size_t k = 1; k < w.size(); ++k) assert(infos->at(w[k - 1]).order_number < infos->at(w[k]).order_number); }
    assert(!c.displayed());
  }
  // TreeInComponent (containment.hpp:241-347) at the root of data/test.nw: the unfolding has 11 nodes and highest_displayed_ancestor is what
  // the reference prints (tests/golden/hda_ref.json item 0); TreeComponentInfos of the same network (comps_ref.json item 0)
  {
    EdgeVec he, ge; LabelMapDefault hl, gl;
    const size_t hn = parse_newick(std::string("((a,(((b)#H1,(c)#H2),(#H1,#H2))),d);"), he, hl), gn = parse_newick(std::string("((a,b),(c,d));"), ge, gl);
    MyNet N(he, hl, consecutive_tag(), hn); MyTree G(ge, gl, consecutive_tag(), gn);
    TreeInComponent<ROMulTree<>, MyTree> tic(N, G, LabelMatchingMap(), N.root());
    assert(tic.subtree().num_nodes() == 11 && tic.origin(0) == N.root());
    const int hda[7] = {-1, 1, 1, 1, 4, 4, 4};
    for (Node v = 1; v < 7; ++v) assert(tic.highest_displayed_ancestor(v) == (Node)hda[v]);
    TreeComponentInfos<MyNet> ci(N);
    for (Node v = 0; v < hn; ++v) assert(ci.comp_root(v) == 0);
    assert(ci.visible_leaf(0) != NoNode && N.is_leaf(ci.visible_leaf(0)) && ci.comp_DAG_edges().empty());
  }
  // VisibleComponentRule (containment.hpp:603-725) on the pair `ref_tc vcr` prints as "R 9 11 2" (tests/golden/vcr_ref.json item 3): the tree
  // component below the reticulation is 1/2-eligible and seen by leaf 11, the root (two host paths to it) is not; guest node 2 = (x,y)
  {
    EdgeVec he, ge; LabelMapDefault hl, gl;
    const size_t hn = parse_newick(std::string("(((a,((x,y))#H1),b),((c,#H1),d));"), he, hl), gn = parse_newick(std::string("((a,b),((c,d),(x,y)));"), ge, gl);
    MyNet N(he, hl, consecutive_tag(), hn); MyTree G(ge, gl, consecutive_tag(), gn);
    VisibleComponentRule<MyNet, MyTree> vc(N, G);
    assert(vc.apply() && vc.get_eligible_component_root() == 9 && vc.matched_guest_node() == 2 && vc.is_half_eligible(9) && !vc.is_half_eligible(0));
    assert(N.is_leaf(vc.visible_leaf()));
    VisibleComponentRule<MyNet, MyTree> vc2(N, G, 9, 11);
    assert(vc2.visible_leaf() == 11 && vc2.matched_guest_node() == 2);
  }
  // the reference's class names for LCA / RMQ (rmq/lca.hpp:79-106, rmq/rmq.hpp:63-72): lca<Tree>, sparse_rmq<It>::query / query_offset
  {
    lca<MyTree> L3(T2);
    assert(L3.query(T2.root(), 3) == T2.root() && LCA(T2, 2, 3) == L2.query(2, 3) && naiveLCA(T2, 5, 6) == L2.query(5, 6));
    const std::vector<int> a = {5, 3, 8, 3, 9, 1, 7};
    sparse_rmq<std::vector<int>::const_iterator> R(a.begin(), a.end());
    assert(R.query_offset(0, 5) == 3 && *R.query(0, 5) == 3 && R.query_offset(0, 7) == 5 && R.query_offset(2, 3) == 2);   // ties -> right (sparse_rmq.hpp:63,89)
  }
  // isomorphism (utils/isomorphism.hpp:309-321): data/nets.nw of the reference, and what the flags change
  {
    auto net = [](const char* nw) { EdgeVec e; auto l = std::make_shared<LabelMapDefault>(); const size_t n = parse_newick(std::string(nw), e, l); return RONetwork<>(e, l, consecutive_tag(), n); };
    const auto A = net("(a,(((b)#H1,(c)#H2),(#H1,#H2)),d);"), B = net("(d,(((b)#H1,(c)#H2),(#H1,#H2)),a);");
    assert(make_iso_mapper(A, B, FLAG_MAP_LEAF_LABELS).check_isomorph());
    const auto T1 = net("((a,b),(c,d));"), T2 = net("((a,c),(b,d));");
    assert(!make_iso_mapper(T1, T2, FLAG_MAP_LEAF_LABELS).check_isomorph());
    assert(make_iso_mapper(T1, T2, 0).check_isomorph());   // -il: same shape
    BatchIsomorphism bi;
    bi.add(A, B); bi.add(T1, T2); bi.add(T1, T1);
    bi.run();
    assert(bi.isomorph(0) && !bi.isomorph(1) && bi.isomorph(2));
  }
  std::cout << "facade ok\n";
  return 0;
}
