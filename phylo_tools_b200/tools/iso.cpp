// tools/iso.cpp -- replacement for examples/iso.cpp of the reference, same command line and output:
//   iso [-v] [-mr] [-mt] [-ma] [-il] <file1> [file2]
// two networks (2 lines of extended Newick in file1, or one per file); prints "reading networks...", "checking isomorphism..."
// and, as the LAST line, "isomorph!" or "not isomorph!" (examples/iso.cpp:66-112); exit code 1 on usage or file errors.
//   -mr / -mt / -ma : labels of reticulations / internal tree nodes / all nodes have to match;  -il : leaf labels need not match
//   iso --batch <file> : NEW: many pairs (2 lines each), one verdict per line
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>
#include "../include/pt/io.hpp"
#include "../include/pt/isomorphism.hpp"

using namespace PT;
using Net = RONetwork<>;
using ENL = EdgesAndNodeLabels<Net>;

static int usage(const char* argv0) {
  std::cerr << "usage: " << argv0 << " [flags] <file1> [file2]\n"
            << "  two networks in extended Newick or edge-list format: both in file1 (one per line), or one in each file\n"
            << "  -v    print both networks before checking\n"
            << "  -mr   reticulation labels must agree\n"
            << "  -mt   labels of internal tree nodes must agree\n"
            << "  -ma   all labels must agree (= -mr -mt, and leaf labels even with -il)\n"
            << "  -il   ignore leaf labels\n"
            << "usage: " << argv0 << " [flags] --batch <file>\n"
            << "  every two lines of <file> are one pair; prints one verdict per line\n";
  return EXIT_FAILURE;
}

int main(int argc, char** argv) {
  bool verbose = false, mr = false, mt = false, ma = false, il = false; std::vector<std::string> files; std::string batch;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "-v") verbose = true; else if (a == "-mr") mr = true; else if (a == "-mt") mt = true; else if (a == "-ma") ma = true; else if (a == "-il") il = true;
    else if (a == "-h" || a == "--help") { usage(argv[0]); return EXIT_SUCCESS; }
    else if (a == "--batch") { if (i + 1 >= argc) return usage(argv[0]); batch = argv[++i]; }
    else if (!a.empty() && a[0] == '-') return usage(argv[0]);
    else files.push_back(a);
  }
  const unsigned char flags = (unsigned char)((!il ? FLAG_MAP_LEAF_LABELS : 0) | (mt ? FLAG_MAP_TREE_LABELS : 0) | (mr ? FLAG_MAP_RETI_LABELS : 0) | (ma ? FLAG_MAP_ALL_LABELS : 0));
  try {
    std::vector<ENL> el;
    if (!batch.empty()) {
      if (!file_exists(batch)) { std::cerr << batch << " cannot be opened for reading" << std::endl; return EXIT_FAILURE; }
      std::ifstream in(batch, std::ios::binary); std::string text((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
      NewickTextIsomorphism b(text, flags);   // parsed and compared on the GPU (pt_iso_from_newick)
      for (size_t i = 0; i < b.size(); ++i) std::cout << (b.status(i) ? "error" : (b.isomorph(i) ? "isomorph!" : "not isomorph!")) << "\n";
      return EXIT_SUCCESS;
    }
    if (files.empty() || files.size() > 2) return usage(argv[0]);
    for (const auto& f : files) if (!file_exists(f)) { std::cerr << f << " cannot be opened for reading" << std::endl; return EXIT_FAILURE; }
    std::cout << "reading networks..." << std::endl;
    read_edgelists(files[0], el);
    if (el.empty()) { std::cerr << "could not read any network from " << files[0] << std::endl; return EXIT_FAILURE; }
    if (el.size() < 2) {
      if (files.size() < 2) { std::cerr << "could not read 2 networks from " << files[0] << " but no other useable source was found" << std::endl; return EXIT_FAILURE; }
      read_edgelists(files[1], el);
      if (el.size() < 2) { std::cerr << "could not read any network from " << files[1] << std::endl; return EXIT_FAILURE; }
    }
    Net N0(el[0].edges, el[0].labels, consecutive_tag(), el[0].num_nodes), N1(el[1].edges, el[1].labels, consecutive_tag(), el[1].num_nodes);
    if (verbose) std::cout << "N0: " << std::endl << get_extended_newick(N0) << std::endl << "N1:" << std::endl << get_extended_newick(N1) << std::endl;
    std::cout << "checking isomorphism..." << std::endl;
    std::cout << (make_iso_mapper(N0, N1, flags).check_isomorph() ? "isomorph!" : "not isomorph!") << std::endl;
  } catch (const std::exception& e) {
    std::cerr << "error: " << e.what() << std::endl;
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}
