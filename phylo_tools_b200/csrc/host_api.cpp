// csrc/host_api.cpp -- host-only entry points of the C-ABI (include/pt_gpu.h): readers, packer, generators.
// Thin C wrappers over the C++ facade in phylo_tools_b200/include/pt/ so that ctypes callers (tests, bench.py)
// and C++ callers share one implementation.  No CUDA in this file.
#include <cstring>
#include <new>
#include <string>
#include <vector>
#include "../../include/pt_gpu.h"
#include "../include/pt/generator.hpp"
#include "../include/pt/newick.hpp"
#include "../include/pt/fast_newick.hpp"
#include "../include/pt/packer.hpp"
#include "../include/pt/io.hpp"
#include <sstream>
#include "pt_error.h"

namespace ptgpu {
char* error_buffer() { static thread_local char buf[512] = {0}; return buf; }
}
using ptgpu::set_error;

struct pt_packer { PT::BatchPacker p; };

extern "C" {

const char* pt_last_error(void) { return ptgpu::error_buffer(); }
int pt_abi_version(void) { return PT_GPU_ABI_VERSION; }

int pt_parse_newick(const char* text, int64_t* num_nodes, int64_t* num_edges, int32_t* edges, int64_t* num_labels,
                    int32_t* label_node, char* label_buf, int64_t* label_buf_len) {
  if (!text || !num_nodes || !num_edges) return set_error(PT_ERR_INVALID, "pt_parse_newick: null argument");
  try {
    PT::EdgeVec el; std::vector<std::string> names;
    const std::string s(text);
    const size_t n = PT::parse_newick(s, el, names);
    *num_nodes = (int64_t)n; *num_edges = (int64_t)el.size();
    int64_t nl = 0, bytes = 0;
    for (size_t v = 0; v < names.size(); ++v) if (!names[v].empty()) { ++nl; bytes += (int64_t)names[v].size() + 1; }
    if (num_labels) *num_labels = nl;
    const int64_t cap = label_buf_len ? *label_buf_len : 0;
    if (label_buf_len) *label_buf_len = bytes;
    if (edges) for (size_t e = 0; e < el.size(); ++e) { edges[2 * e] = (int32_t)el[e].tail(); edges[2 * e + 1] = (int32_t)el[e].head(); }
    if (label_node && label_buf) {
      if (cap < bytes) return set_error(PT_ERR_INVALID, "pt_parse_newick: label buffer too small (%lld < %lld)", (long long)cap, (long long)bytes);
      int64_t k = 0, off = 0;
      for (size_t v = 0; v < names.size(); ++v) if (!names[v].empty()) { label_node[k++] = (int32_t)v; memcpy(label_buf + off, names[v].c_str(), names[v].size() + 1); off += (int64_t)names[v].size() + 1; }
    }
    return PT_OK;
  } catch (const PT::MalformedNewick& e) { return set_error(PT_ERR_PARSE, "MalformedNewick: %s", e.what());
  } catch (const std::bad_alloc&) { return set_error(PT_ERR_NOMEM, "out of host memory");
  } catch (const std::exception& e) { return set_error(PT_ERR_INVALID, "%s", e.what()); }
}

int pt_parse_edgelist(const char* text, int64_t* num_nodes, int64_t* num_edges, int32_t* edges, int64_t* num_labels,
                      int32_t* label_node, char* label_buf, int64_t* label_buf_len) {
  if (!text || !num_nodes || !num_edges) return set_error(PT_ERR_INVALID, "pt_parse_edgelist: null argument");
  try {
    PT::EdgeVec el; std::vector<std::string> names;
    std::istringstream in{std::string(text)};
    const size_t n = PT::parse_edgelist(in, el, names);
    *num_nodes = (int64_t)n; *num_edges = (int64_t)el.size();
    int64_t nl = 0, bytes = 0;
    for (size_t v = 0; v < names.size(); ++v) if (!names[v].empty()) { ++nl; bytes += (int64_t)names[v].size() + 1; }
    if (num_labels) *num_labels = nl;
    const int64_t cap = label_buf_len ? *label_buf_len : 0;
    if (label_buf_len) *label_buf_len = bytes;
    if (edges) for (size_t e = 0; e < el.size(); ++e) { edges[2 * e] = (int32_t)el[e].tail(); edges[2 * e + 1] = (int32_t)el[e].head(); }
    if (label_node && label_buf) {
      if (cap < bytes) return set_error(PT_ERR_INVALID, "pt_parse_edgelist: label buffer too small (%lld < %lld)", (long long)cap, (long long)bytes);
      int64_t k = 0, off = 0;
      for (size_t v = 0; v < names.size(); ++v) if (!names[v].empty()) { label_node[k++] = (int32_t)v; memcpy(label_buf + off, names[v].c_str(), names[v].size() + 1); off += (int64_t)names[v].size() + 1; }
    }
    return PT_OK;
  } catch (const PT::MalformedEdgeVec& e) { return set_error(PT_ERR_PARSE, "MalformedEdgeVec: %s", e.what());
  } catch (const std::bad_alloc&) { return set_error(PT_ERR_NOMEM, "out of host memory");
  } catch (const std::exception& e) { return set_error(PT_ERR_INVALID, "%s", e.what()); }
}

int pt_packer_create(pt_packer** out) {
  if (!out) return set_error(PT_ERR_INVALID, "pt_packer_create: null argument");
  *out = new (std::nothrow) pt_packer();
  return *out ? PT_OK : set_error(PT_ERR_NOMEM, "out of host memory");
}
int pt_packer_add_newick(pt_packer* p, const char* host_newick, const char* guest_newick) {
  if (!p || !host_newick || !guest_newick) return set_error(PT_ERR_INVALID, "pt_packer_add_newick: null argument");
  try { p->p.add_newick(host_newick, guest_newick); return PT_OK; }
  catch (const PT::MalformedNewick& e) { return set_error(PT_ERR_PARSE, "MalformedNewick: %s", e.what()); }
  catch (const std::bad_alloc&) { return set_error(PT_ERR_NOMEM, "out of host memory"); }
  catch (const std::exception& e) { return set_error(PT_ERR_INVALID, "%s", e.what()); }
}
int pt_packer_add_newick_text(pt_packer* p, const char* text, int64_t len, int num_threads, int64_t* num_added) {
  if (!p || !text || len < 0) return set_error(PT_ERR_INVALID, "pt_packer_add_newick_text: bad argument");
  try { const size_t k = PT::add_newick_text(p->p, text, (size_t)len, num_threads); if (num_added) *num_added = (int64_t)k; return PT_OK; }
  catch (const PT::MalformedNewick& e) { return set_error(PT_ERR_PARSE, "%s", e.what()); }
  catch (const std::bad_alloc&) { return set_error(PT_ERR_NOMEM, "out of host memory"); }
  catch (const std::exception& e) { return set_error(PT_ERR_INVALID, "%s", e.what()); }
}
int pt_packer_add_generated(pt_packer* p, int64_t count, int32_t num_leaves, int32_t num_retis, uint64_t seed, int kind) {
  if (!p || count < 0 || num_leaves <= 0 || num_retis < 0 || kind < 0 || kind > 2) return set_error(PT_ERR_INVALID, "pt_packer_add_generated: bad argument");
  try { for (int64_t i = 0; i < count; ++i) p->p.add_generated(num_leaves, num_retis, seed + (uint64_t)i, kind); return PT_OK; }
  catch (const std::bad_alloc&) { return set_error(PT_ERR_NOMEM, "out of host memory"); }
  catch (const std::exception& e) { return set_error(PT_ERR_INVALID, "%s", e.what()); }
}
int pt_packer_desc(pt_packer* p, pt_batch_desc* desc) {
  if (!p || !desc) return set_error(PT_ERR_INVALID, "pt_packer_desc: null argument");
  *desc = p->p.desc();
  return PT_OK;
}
int pt_packer_desc_compact(pt_packer* p, pt_batch_desc* desc) {
  if (!p || !desc) return set_error(PT_ERR_INVALID, "pt_packer_desc_compact: null argument");
  try { *desc = p->p.desc_compact(); return PT_OK; }
  catch (const std::bad_alloc&) { return set_error(PT_ERR_NOMEM, "out of host memory"); }
  catch (const std::exception& e) { return set_error(PT_ERR_INVALID, "pt_packer_desc_compact: %s", e.what()); }
}
int pt_packer_desc_bytes(pt_packer* p, pt_batch_desc* desc) {
  if (!p || !desc) return set_error(PT_ERR_INVALID, "pt_packer_desc_bytes: null argument");
  try { *desc = p->p.desc_bytes(); return PT_OK; }
  catch (const std::bad_alloc&) { return set_error(PT_ERR_NOMEM, "out of host memory"); }
  catch (const std::exception& e) { return set_error(PT_ERR_INVALID, "pt_packer_desc_bytes: %s", e.what()); }
}
int pt_packer_get_newick(pt_packer* p, int64_t i, char* host_buf, int64_t* host_len, char* guest_buf, int64_t* guest_len) {
  if (!p || !host_len || !guest_len || i < 0 || (size_t)i >= p->p.size()) return set_error(PT_ERR_INVALID, "pt_packer_get_newick: bad argument");
  try {
    const auto nw = p->p.newick_of((size_t)i);
    const int64_t hc = *host_len, gc = *guest_len;
    *host_len = (int64_t)nw.first.size() + 1; *guest_len = (int64_t)nw.second.size() + 1;
    if (host_buf && guest_buf) {
      if (hc < *host_len || gc < *guest_len) return set_error(PT_ERR_INVALID, "pt_packer_get_newick: buffer too small");
      memcpy(host_buf, nw.first.c_str(), nw.first.size() + 1); memcpy(guest_buf, nw.second.c_str(), nw.second.size() + 1);
    }
    return PT_OK;
  } catch (const std::exception& e) { return set_error(PT_ERR_INVALID, "%s", e.what()); }
}
int pt_packer_free(pt_packer* p) { delete p; return PT_OK; }

int pt_gen_random_tree(int64_t num_leaves, uint64_t seed, int32_t* parent_out) {
  if (num_leaves <= 0 || !parent_out || 2 * num_leaves - 1 > INT32_MAX) return set_error(PT_ERR_INVALID, "pt_gen_random_tree: bad argument");
  PT::SplitMix64 rng(seed);
  std::vector<int32_t> roots((size_t)num_leaves);
  for (int64_t i = 0; i < num_leaves; ++i) roots[(size_t)i] = (int32_t)i;
  int64_t nr = num_leaves; int32_t next = (int32_t)num_leaves;
  while (nr > 1) {
    const int64_t i = (int64_t)(rng.next() % (uint64_t)nr); std::swap(roots[(size_t)i], roots[(size_t)nr - 1]);
    const int64_t j = (int64_t)(rng.next() % (uint64_t)(nr - 1));
    const int32_t a = roots[(size_t)nr - 1], b = roots[(size_t)j];
    parent_out[a] = next; parent_out[b] = next; roots[(size_t)j] = next++; --nr;
  }
  parent_out[roots[0]] = -1;
  return PT_OK;
}

}  // extern "C"
