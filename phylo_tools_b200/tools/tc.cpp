// tools/tc.cpp -- replacement for examples/tc.cpp of the reference, same command line and the same protocol:
//   tc [-v] <file1> [file2]      two networks (2 lines of extended Newick in file1, or one per file); the network is the
//                                host, or the first if both are trees (examples/tc.cpp:110-111); guest must be a tree
//   tc -r <x> <y> <z>            random tree with x internal nodes and y leaves plus z extra arcs; checks that the tree is displayed
//   tc --batch <file>            NEW: many pairs (2 lines each); one verdict per line
// The verdict is the LAST line of stdout: "displayed" / "not displayed" /
// "sorry, can't check network-network containment yet..." (examples/tc.cpp:135-144); exit code 0 on a verdict,
// 1 on usage or file errors (examples/tc.cpp:29-45,103-106).
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <iterator>
#include <string>
#include <vector>
#include "../include/pt/containment.hpp"
#include "../include/pt/generator.hpp"
#include "../include/pt/io.hpp"

using namespace PT;
using MyTree = ROTree<>;
using MyNet = CompatibleRWNetwork<MyTree>;
using ENL = EdgesAndNodeLabels<MyTree>;

static int usage(const char* argv0) {
  std::cerr << argv0 << " [-v] <file1> [file2]\n\tfile1 and file2 describe two networks (either file1 contains 2 lines of extended newick or both file1 and file2 describe a network in extended newick or edgelist format)\n"
            << "\tUnless the first network is a tree and the second is not, we try to embed the second network in the first.\n\n"
            << argv0 << " -r <x> <y> <z>\n\trandomize a tree with x internal nodes + y leaves and add z additional edges, then check containment of the tree in the network\n\n"
            << argv0 << " --batch <file>\n\tcheck every pair of lines of <file>; one verdict per line\n";
  return EXIT_FAILURE;
}

int main(int argc, char** argv) {
  bool verbose = false; std::vector<std::string> files, rargs; std::string batch;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "-v") verbose = true;
    else if (a == "-h" || a == "--help") { usage(argv[0]); return EXIT_SUCCESS; }
    else if (a == "-r") { for (int k = 0; k < 3 && i + 1 < argc; ++k) rargs.push_back(argv[++i]); if (rargs.size() != 3) return usage(argv[0]); }
    else if (a == "--batch") { if (i + 1 >= argc) return usage(argv[0]); batch = argv[++i]; }
    else if (!a.empty() && a[0] == '-') return usage(argv[0]);
    else files.push_back(a);
  }
  try {
    if (!batch.empty()) {
      if (!file_exists(batch)) { std::cerr << batch << " cannot be opened for reading" << std::endl; return EXIT_FAILURE; }
      std::ifstream in(batch, std::ios::binary); std::string text((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
      NewickTextContainment bc(text);   // parsed and solved on the GPU (pt_batch_create_from_newick)
      for (size_t i = 0; i < bc.size(); ++i) std::cout << (bc.status(i) ? "error" : (bc.displayed(i) ? "displayed" : "not displayed")) << "\n";
      return EXIT_SUCCESS;
    }
    std::vector<ENL> el;
    if (!rargs.empty()) {
      const int x = std::atoi(rargs[0].c_str()), y = std::atoi(rargs[1].c_str()), z = std::atoi(rargs[2].c_str());
      if (x == 0) { std::cerr << "cannot construct tree with " << rargs[0] << " nodes\n"; return EXIT_FAILURE; }
      if (y <= x) { std::cerr << "cannot construct tree with " << rargs[0] << " internal nodes & " << rargs[1] << " leaves\n"; return EXIT_FAILURE; }
      SplitMix64 rng(0xB200ull + (uint64_t)x * 1000003u + (uint64_t)y * 1009u + (uint64_t)z);
      el.resize(2);
      generate_random_tree(el[1].edges, *el[1].labels, (uint32_t)x, (uint32_t)y, rng);
      el[1].num_nodes = (size_t)x + y;
      el[0].edges = el[1].edges; *el[0].labels = *el[1].labels; el[0].num_nodes = el[1].num_nodes;
      add_random_edges(el[0].edges, el[0].num_nodes, (uint32_t)z, rng);
    } else {
      if (files.empty() || files.size() > 2) return usage(argv[0]);
      for (const auto& f : files) if (!file_exists(f)) { std::cerr << f << " cannot be opened for reading" << std::endl; return EXIT_FAILURE; }
      read_edgelists(files, el);
      if (el.size() < 2) { std::cerr << "could not read 2 networks from files" << std::endl; return EXIT_FAILURE; }
    }
    const bool host_idx = (el[0].is_tree() && !el[1].is_tree());
    MyNet N(el[host_idx].edges, el[host_idx].labels, consecutive_tag(), el[host_idx].num_nodes);
    MyTree T(el[1 - host_idx].edges, el[1 - host_idx].labels, consecutive_tag(), el[1 - host_idx].num_nodes);
    if (verbose) std::cout << "N:\n" << get_extended_newick(N) << "\nT:\n" << get_extended_newick(T) << "\n";
    std::cout << "\n\n starting the containment engine...\n\n";
    if (T.is_tree()) {
      if (N.is_tree()) { TreeInTreeContainment tc(N, T); std::cout << (tc.displayed() ? "displayed\n" : "not displayed\n"); }
      else { TreeInNetContainment tc(N, T); std::cout << (tc.displayed() ? "displayed\n" : "not displayed\n"); }
    } else std::cout << "sorry, can't check network-network containment yet...\n";
  } catch (const MalformedNewick& e) { std::cerr << e.what() << std::endl; return EXIT_FAILURE;
  } catch (const std::exception& e) { std::cerr << "error: " << e.what() << std::endl; return EXIT_FAILURE; }
  return EXIT_SUCCESS;
}
