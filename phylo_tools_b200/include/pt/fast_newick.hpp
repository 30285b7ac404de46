// pt/fast_newick.hpp -- batch reader: a text buffer of extended-Newick lines (two per instance: host, guest) straight into
// the flat arrays of include/pt_gpu.h, in parallel on all host cores.
//
// once the kernels take 15 ms per 10^5 instances, reading the text is what a caller waits for.  The
// per-instance reader of pt/newick.hpp (= the reference's io/newick.hpp behaviour) builds an edge vector, a label map and two
// storage objects per network; this one does the same right-to-left scan (same node numbering: root 0, ids in the order the
// name tokens are met from the right; same hybrid rule; same ":length" and white-space handling; same double-edge check;
// io/newick.hpp:159-272) but writes nodes, arcs and label spans into per-thread flat buffers with no per-node allocation,
// then the per-thread pieces are concatenated.  The output is bit-identical to BatchPacker::add_newick on the same lines
// (tests/test_host.py), including the child order of the CSR (reverse arc order, utils/storage_adj_immutable.hpp:119) and
// the per-instance label ids (first appearance over host nodes, then guest nodes).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include "packer.hpp"

namespace PT {

namespace fastnw {

struct Span { uint32_t off, len; };
struct Parsed {  // one network, storage reused between calls
  std::vector<int32_t> tail, head;       // arcs in reference order
  std::vector<Span> label;               // per node (len 0 = unlabelled)
  std::vector<std::pair<long, int32_t>> hybrids;
  std::vector<int32_t> open;
  int32_t n = 0, narcs = 0;  // valid prefix of label / tail,head
  bool has_hybrid = false;
};

// parses line s[0, len) ; returns nullptr on success or a static error text.  One backward pass; the per-node and per-arc
// outputs are written through raw pointers into buffers sized once per line (a node costs at least one character).
inline const char* parse_line(const char* s, long len, Parsed& P) {
  P.hybrids.clear(); P.open.clear(); P.n = 0; P.has_hybrid = false;
  const size_t cap = (size_t)len + 2;
  if (P.label.size() < cap) { P.label.resize(cap); P.tail.resize(cap); P.head.resize(cap); }
  Span* lab = P.label.data(); int32_t* tl = P.tail.data(); int32_t* hd = P.head.data();
  int32_t n = 0, m = 0;
  auto finish = [&](const char* err) { P.n = n; P.narcs = m; return err; };
  long back = len - 1;
#define FNW_SKIP_WS() while (back >= 0 && (s[back] == ' ' || (unsigned)(s[back] - 9) < 5u)) --back
#define FNW_SKIP_LEN() do { long i_ = back; while (i_ >= 0) { const char c_ = s[i_]; if ((c_ >= '0' && c_ <= '9') || c_ == '.' || c_ == '-' || c_ == '+' || c_ == 'E' || c_ == 'e') --i_; else break; } if (i_ >= 0 && s[i_] == ':') back = i_ - 1; } while (0)
  FNW_SKIP_WS();
  if (back < 0) return finish(nullptr);
  if (s[back] != ';') return finish("expected ';'");
  --back;
  bool want_subtree = true;
  int32_t done = 0;
  for (;;) {
    if (want_subtree) {
      FNW_SKIP_WS();
      long i = back, sharp = -1;
      while (i >= 0) { const char c = s[i]; if (c == '(' || c == ')' || c == ',') break; if (c == '#' && sharp < 0) sharp = i; --i; }
      const long nb = i + 1, nl = back - i;  // name = s[nb, nb+nl)
      back = i;
      int32_t id;
      if (sharp >= 0) {
        long d = sharp;
        while (d < nb + nl && !(s[d] >= '0' && s[d] <= '9')) ++d;
        if (d == nb + nl) return finish("found '#' but no hybrid number");
        long key = 0;
        while (d < nb + nl && s[d] >= '0' && s[d] <= '9') { key = key * 10 + (s[d] - '0'); ++d; }
        id = -1;
        for (auto& h : P.hybrids) if (h.first == key) { id = h.second; break; }
        if (id < 0) { id = n; P.hybrids.emplace_back(key, id); lab[n++] = Span{(uint32_t)nb, (uint32_t)(sharp - nb)}; }
        P.has_hybrid = true;
      } else { id = n; lab[n++] = Span{(uint32_t)nb, (uint32_t)nl}; }
      if (back > 0 && s[back] == ')') { --back; P.open.push_back(id); FNW_SKIP_LEN(); continue; }
      done = id; want_subtree = false;
    }
    FNW_SKIP_WS();
    if (P.open.empty()) break;
    const int32_t parent = P.open.back();
    tl[m] = parent; hd[m++] = done;
    if (back < 0) return finish("unexpected start of string");
    const char c = s[back];
    if (c == ',') { --back; FNW_SKIP_LEN(); want_subtree = true; }
    else if (c == '(') { --back; P.open.pop_back(); done = parent; }
    else return finish("expected ',' or '('");
  }
#undef FNW_SKIP_WS
#undef FNW_SKIP_LEN
  return finish(nullptr);
}

struct Chunk {  // flat output of one thread
  std::vector<int64_t> hn, hm, gn;  // per instance sizes
  std::vector<int32_t> succ_off, succ_adj, pred_off, pred_adj, hlabel, gparent, gchild_off, gchild_adj, glabel;
  int32_t max_hn = 0, max_hm = 0, max_gn = 0, max_l = 0;
  std::string error; int64_t error_pair = -1;
};

// CSR of one network appended to off/adj (children in reverse arc order like Phylogeny::build); returns false on a double edge
inline bool append_csr(const Parsed& P, std::vector<int32_t>& off, std::vector<int32_t>& adj, std::vector<int32_t>& scratch, bool by_head) {
  const int32_t n = P.n, m = P.narcs;
  const size_t o0 = off.size(), a0 = adj.size();
  off.resize(o0 + n + 1, 0); adj.resize(a0 + m);
  int32_t* O = off.data() + o0; int32_t* A = adj.data() + a0;
  const std::vector<int32_t>& key = by_head ? P.head : P.tail; const std::vector<int32_t>& val = by_head ? P.tail : P.head;
  for (int32_t e = 0; e < m; ++e) O[key[e] + 1]++;
  for (int32_t v = 0; v < n; ++v) O[v + 1] += O[v];
  scratch.assign(O + 1, O + n + 1);
  for (int32_t e = 0; e < m; ++e) A[--scratch[key[e]]] = val[e];
  if (!by_head) {  // double-edge check (io/newick.hpp:216-217): children lists are short
    for (int32_t v = 0; v < n; ++v) for (int32_t a = O[v]; a < O[v + 1]; ++a) for (int32_t b = a + 1; b < O[v + 1]; ++b) if (A[a] == A[b]) return false;
  }
  return true;
}

}  // namespace fastnw

// appends every pair of lines of [text, text+len) to the packer; returns the number of instances added.
// Throws MalformedNewick / std::logic_error naming the first bad pair (nothing is added then).
inline size_t add_newick_text(BatchPacker& pk, const char* text, size_t len, int num_threads = 0) {
  using namespace fastnw;
  // 1. line index (non-blank lines)
  std::vector<std::pair<size_t, size_t>> lines;
  for (size_t p = 0; p < len;) {
    const void* nlp = memchr(text + p, '\n', len - p);
    const size_t e = nlp ? (size_t)((const char*)nlp - text) : len;
    size_t b = p, ee = e;
    while (b < ee && (text[b] == ' ' || text[b] == '\t' || text[b] == '\r')) ++b;
    while (ee > b && (text[ee - 1] == ' ' || text[ee - 1] == '\t' || text[ee - 1] == '\r')) --ee;
    if (ee > b) lines.emplace_back(b, ee - b);
    p = e + 1;
  }
  const size_t npairs = lines.size() / 2;
  if (npairs == 0) return 0;
  int T = num_threads > 0 ? num_threads : (int)std::thread::hardware_concurrency();
  T = std::max(1, std::min<int>(T, (int)((npairs + 255) / 256)));
  std::vector<Chunk> chunks(T);
  auto work = [&](int t) {
    Chunk& C = chunks[t];
    const size_t p0 = npairs * t / T, p1 = npairs * (t + 1) / T;
    Parsed A, B; std::vector<int32_t> scratch;
    {  // one node per ~2.7 text bytes in generated networks; reserving avoids the doubling copies
      size_t bytes = 0; for (size_t pi = p0; pi < p1; ++pi) bytes += lines[2 * pi].second + lines[2 * pi + 1].second;
      const size_t est = bytes / 2 + 64 * (p1 - p0);
      C.succ_off.reserve(est); C.succ_adj.reserve(est); C.pred_off.reserve(est); C.pred_adj.reserve(est); C.hlabel.reserve(est);
      C.gparent.reserve(est); C.gchild_off.reserve(est); C.gchild_adj.reserve(est); C.glabel.reserve(est);
      C.hn.reserve(p1 - p0); C.hm.reserve(p1 - p0); C.gn.reserve(p1 - p0);
    }
    struct Ent { uint32_t hash; Span sp; int32_t id; const char* base; };
    std::vector<Ent> table; std::vector<int32_t> slots;
    for (size_t pi = p0; pi < p1; ++pi) {
      const char* l0 = text + lines[2 * pi].first; const char* l1 = text + lines[2 * pi + 1].first;
      const char* err = parse_line(l0, (long)lines[2 * pi].second, A);
      if (!err) err = parse_line(l1, (long)lines[2 * pi + 1].second, B);
      if (err) { C.error = std::string("MalformedNewick: ") + err; C.error_pair = (int64_t)pi; return; }
      // examples/tc.cpp:110-111: the network is the host, or the first line if both are trees
      const bool t0 = A.narcs + 1 == A.n, t1 = B.narcs + 1 == B.n;
      const bool swap = t0 && !t1;
      const Parsed& H = swap ? B : A; const Parsed& G = swap ? A : B;
      const char* hbase = swap ? l1 : l0; const char* gbase = swap ? l0 : l1;
      if (G.narcs + 1 != G.n || G.has_hybrid) { C.error = "guest must be a tree"; C.error_pair = (int64_t)pi; return; }
      // one root each (utils/storage_adj_common.hpp:103-109): node 0 is the root by construction; every other node has an in-arc
      // label dictionary: open addressing over (hash, span)
      const int32_t nl_max = H.n + G.n;
      size_t cap = 16; while (cap < (size_t)nl_max * 2) cap <<= 1;
      slots.assign(cap, -1); table.clear();
      auto id_of = [&](const char* base, Span sp) -> int32_t {
        if (sp.len == 0) return -1;
        uint32_t h = 2166136261u;
        for (uint32_t k = 0; k < sp.len; ++k) h = (h ^ (unsigned char)base[sp.off + k]) * 16777619u;
        size_t i = h & (cap - 1);
        for (;;) {
          const int32_t s = slots[i];
          if (s < 0) { slots[i] = (int32_t)table.size(); table.push_back(Ent{h, sp, (int32_t)table.size(), base}); return (int32_t)table.size() - 1; }
          const Ent& en = table[s];
          if (en.hash == h && en.sp.len == sp.len && memcmp(en.base + en.sp.off, base + sp.off, sp.len) == 0) return en.id;
          i = (i + 1) & (cap - 1);
        }
      };
      if (!append_csr(H, C.succ_off, C.succ_adj, scratch, false)) { C.error = "MalformedNewick: read double edge"; C.error_pair = (int64_t)pi; return; }
      append_csr(H, C.pred_off, C.pred_adj, scratch, true);
      for (int32_t v = 0; v < H.n; ++v) C.hlabel.push_back(id_of(hbase, H.label[v]));
      const size_t gp0 = C.gparent.size();
      C.gparent.resize(gp0 + G.n, -1);
      for (int32_t e = 0; e < G.narcs; ++e) C.gparent[gp0 + G.head[e]] = G.tail[e];
      if (!append_csr(G, C.gchild_off, C.gchild_adj, scratch, false)) { C.error = "MalformedNewick: read double edge"; C.error_pair = (int64_t)pi; return; }
      for (int32_t v = 0; v < G.n; ++v) C.glabel.push_back(id_of(gbase, G.label[v]));
      C.hn.push_back(H.n); C.hm.push_back((int64_t)H.narcs); C.gn.push_back(G.n);
      C.max_hn = std::max(C.max_hn, H.n); C.max_hm = std::max(C.max_hm, H.narcs); C.max_gn = std::max(C.max_gn, G.n);
      C.max_l = std::max(C.max_l, (int32_t)table.size());
    }
  };
  if (T == 1) work(0);
  else { std::vector<std::thread> th; for (int t = 0; t < T; ++t) th.emplace_back(work, t); for (auto& x : th) x.join(); }
  for (const Chunk& C : chunks) if (C.error_pair >= 0) {
    const std::string msg = "pair " + std::to_string(C.error_pair) + ": " + C.error;
    if (C.error.rfind("MalformedNewick", 0) == 0) throw MalformedNewick(std::string(), (long)C.error_pair, msg);
    throw std::logic_error(msg);
  }
  // 2. concatenate: one resize per array, then every thread copies its own piece to its final place
  struct Base { size_t so, sa, po, pa, hl, gp, gco, gca, gl, inst; };
  std::vector<Base> base(T + 1);
  base[0] = Base{pk.host_succ_off.size(), pk.host_succ_adj.size(), pk.host_pred_off.size(), pk.host_pred_adj.size(), pk.host_label.size(), pk.guest_parent.size(),
                 pk.guest_child_off.size(), pk.guest_child_adj.size(), pk.guest_label.size(), pk.size()};
  for (int t = 0; t < T; ++t) {
    const Chunk& C = chunks[t]; const Base& b = base[t];
    base[t + 1] = Base{b.so + C.succ_off.size(), b.sa + C.succ_adj.size(), b.po + C.pred_off.size(), b.pa + C.pred_adj.size(), b.hl + C.hlabel.size(), b.gp + C.gparent.size(),
                       b.gco + C.gchild_off.size(), b.gca + C.gchild_adj.size(), b.gl + C.glabel.size(), b.inst + C.hn.size()};
    pk.max_host_nodes = std::max(pk.max_host_nodes, C.max_hn); pk.max_host_arcs = std::max(pk.max_host_arcs, C.max_hm);
    pk.max_guest_nodes = std::max(pk.max_guest_nodes, C.max_gn); pk.max_labels = std::max(pk.max_labels, C.max_l);
  }
  const Base& E = base[T];
  const int64_t hn0 = pk.host_node_off.back(), hm0 = pk.host_arc_off.back(), gn0 = pk.guest_node_off.back();
  pk.host_succ_off.resize(E.so); pk.host_succ_adj.resize(E.sa); pk.host_pred_off.resize(E.po); pk.host_pred_adj.resize(E.pa); pk.host_label.resize(E.hl);
  pk.guest_parent.resize(E.gp); pk.guest_child_off.resize(E.gco); pk.guest_child_adj.resize(E.gca); pk.guest_label.resize(E.gl);
  pk.host_node_off.resize(E.inst + 1); pk.host_arc_off.resize(E.inst + 1); pk.guest_node_off.resize(E.inst + 1);
  pk.dict_names.resize(E.inst);  // names are not kept on the fast path (newick_of() is unavailable for these instances)
  std::vector<int64_t> hn_base(T + 1, hn0), hm_base(T + 1, hm0), gn_base(T + 1, gn0);
  for (int t = 0; t < T; ++t) {
    int64_t a = 0, b = 0, c = 0;
    for (size_t i = 0; i < chunks[t].hn.size(); ++i) { a += chunks[t].hn[i]; b += chunks[t].hm[i]; c += chunks[t].gn[i]; }
    hn_base[t + 1] = hn_base[t] + a; hm_base[t + 1] = hm_base[t] + b; gn_base[t + 1] = gn_base[t] + c;
  }
  auto place = [&](int t) {
    const Chunk& C = chunks[t]; const Base& b = base[t];
    auto cp = [](std::vector<int32_t>& dst, size_t at, const std::vector<int32_t>& src) { if (!src.empty()) memcpy(dst.data() + at, src.data(), src.size() * 4); };
    cp(pk.host_succ_off, b.so, C.succ_off); cp(pk.host_succ_adj, b.sa, C.succ_adj); cp(pk.host_pred_off, b.po, C.pred_off); cp(pk.host_pred_adj, b.pa, C.pred_adj);
    cp(pk.host_label, b.hl, C.hlabel); cp(pk.guest_parent, b.gp, C.gparent); cp(pk.guest_child_off, b.gco, C.gchild_off); cp(pk.guest_child_adj, b.gca, C.gchild_adj);
    cp(pk.guest_label, b.gl, C.glabel);
    int64_t a = hn_base[t], m = hm_base[t], g = gn_base[t];
    for (size_t i = 0; i < C.hn.size(); ++i) { a += C.hn[i]; m += C.hm[i]; g += C.gn[i]; pk.host_node_off[b.inst + i + 1] = a; pk.host_arc_off[b.inst + i + 1] = m; pk.guest_node_off[b.inst + i + 1] = g; }
  };
  if (T == 1) place(0);
  else { std::vector<std::thread> th; for (int t = 0; t < T; ++t) th.emplace_back(place, t); for (auto& x : th) x.join(); }
  return npairs;
}

}  // namespace PT
