// tools/gen.cpp -- replacement for examples/gen.cpp: random binary network as extended Newick.
//   gen [-n nodes] [-r retis] [-l leaves] [-a] [--seed s] [--with-tree] [file]
//   gen --pairs count -l leaves -r retis [--seed s] [--kind k] [file] : `count` (network, tree) pairs, two lines each, exactly the
//       instances pt_packer_add_generated(count, leaves, retis, seed, kind) makes (bench.py's workload; this program does not
//       load libptgpu.so, so the reference arm of the benchmark can make its input without the product library)
// Node-count arithmetic as in examples/gen.cpp:37-84 (n = 2r + 2l - 1; defaults n=99, r=10; with one of -n/-r/-l given the
// network has ~10% reticulations).  Unlike the reference (unseeded rand(): every run prints the same network) the draw is
// seeded; --with-tree appends a tree displayed by a random switching (the reference has no such option).
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include "../include/pt/generator.hpp"
#include "../include/pt/network.hpp"
#include "../include/pt/newick.hpp"
#include "../include/pt/packer.hpp"

using namespace PT;

int main(int argc, char** argv) {
  long n = -1, r = -1, l = -1, pairs = -1; int kind = 0; bool append_mode = false, with_tree = false; uint64_t seed = 0xB200; std::string file;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto next = [&]() -> long { if (i + 1 >= argc) { std::cerr << "missing value for " << a << "\n"; std::exit(EXIT_FAILURE); } return std::atol(argv[++i]); };
    if (a == "-n") n = next(); else if (a == "-r") r = next(); else if (a == "-l") l = next(); else if (a == "-a") append_mode = true;
    else if (a == "--seed") seed = (uint64_t)next(); else if (a == "--pairs") pairs = next(); else if (a == "--kind") kind = (int)next(); else if (a == "--with-tree") with_tree = true; else if (a == "-v") {} else if (a == "-h") { std::cout << argv[0] << " [-n nodes] [-r retis] [-l leaves] [-a] [--seed s] [--with-tree] [file]\n"; return 0; }
    else file = a;
  }
  try {
    if (pairs >= 0) {
      if (l <= 0 || r < 0 || kind < 0 || kind > 2) throw std::logic_error("--pairs needs -l, -r and a kind in 0..2");
      std::ofstream fo; if (!file.empty()) fo.open(file, append_mode ? std::ios::app : std::ios::out);
      std::ostream& o = file.empty() ? std::cout : fo;
      for (long i = 0; i < pairs; ++i) {
        BatchPacker P;
        P.add_generated((int32_t)l, (int32_t)r, seed + (uint64_t)i, kind);
        const auto nw = P.newick_of(0);
        o << nw.first << "\n" << nw.second << "\n";
      }
      return 0;
    }
    const int given = (n >= 0) + (r >= 0) + (l >= 0);
    if (given == 0) { n = 99; r = 10; l = l_from_nr((uint32_t)n, (uint32_t)r); }
    else if (given == 1) {
      if (n >= 0) { r = n / 10; l = l_from_nr((uint32_t)n, (uint32_t)r); }
      else if (r >= 0) { n = 10 * r + 1; l = l_from_nr((uint32_t)n, (uint32_t)r); }
      else { r = (2 * l - 1) / 8; n = n_from_rl((uint32_t)r, (uint32_t)l); }
    } else if (given == 2) {
      if (l < 0) l = l_from_nr((uint32_t)n, (uint32_t)r); else if (n < 0) n = n_from_rl((uint32_t)r, (uint32_t)l); else r = l_from_nr((uint32_t)n, (uint32_t)l);
    } else if ((long)l_from_nr((uint32_t)n, (uint32_t)r) != l) throw std::logic_error("there is no binary network with these numbers of vertices, reticulations and leaves");
    const long t = n - r - l;
    if (n < 0 || r < 0 || l < 0 || t < 0) throw std::logic_error("network geometry implied by your parameters is invalid");
    std::cerr << "constructing network with " << n << " vertices: " << t << " tree nodes, " << r << " reticulations and " << l << " leaves" << std::endl;
    SplitMix64 rng(seed);
    EdgeVec el; LabelMapDefault names;
    generate_random_binary_edgelist_trl(el, names, (uint32_t)t, (uint32_t)r, (uint32_t)l, rng);
    RONetwork<> N(el, names, consecutive_tag(), (size_t)n);
    std::string out = get_extended_newick(N) + "\n";
    if (with_tree) { EdgeVec te; LabelMapDefault tn; const size_t tnodes = random_displayed_tree(el, (size_t)n, names, rng, te, tn); ROTree<> T(te, tn, consecutive_tag(), tnodes); out += get_extended_newick(T) + "\n"; }
    if (!file.empty()) { std::ofstream o(file, append_mode ? std::ios::app : std::ios::out); o << out; } else std::cout << out;
  } catch (const std::exception& e) { std::cerr << "cannot generate such a network: " << e.what() << std::endl; return EXIT_FAILURE; }
  return 0;
}
