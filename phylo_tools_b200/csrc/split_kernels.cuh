// csrc/split_kernels.cuh -- bridge splitting inside the containment engine (included by lca_kernels.cu).
//
// north_star subsystem (1): "bridge and biconnected-component splitting" for instances that do not fit one CTA's shared memory
// (BASELINE config 4: n = 10^6, r = 10^3).  Semantic spec: utils/bridges.hpp:42-118, utils/biconnected_comps.hpp:33-88 (the reference
// never wires them into tc; its rule loop would need days at this size).
//
// An arc whose removal disconnects the network (a bridge / cut arc) u -> v pins down a cluster: everything below v hangs on that
// one arc in EVERY switching, so the guest must contain a node t_v with exactly the labels below v, and the containment question
// falls apart into independent questions -- one per bridge-free component, with everything below its outgoing bridges shrunk to a
// pendant leaf on both sides.  A config-4 network is a tree with ~10^3 small blobs: 98 % of its arcs are bridges, and a component
// that is a single tree node is decided by one comparison (the guest nodes of its children must hang on its own guest node).
//
// The whole batch is treated as ONE forest: a super-root above the instance roots (those arcs are bridges, so instance i is paired
// with guest i by the same cluster test).  Steps, all data-parallel over the ~B * n nodes of the batch:
//   1. concatenate (ids rebased), degrees, label tables; the split path only takes CLEAN batches (every leaf labelled, every label
//      on both sides, no unary guest node, one root per instance) -- anything else is left to the global-memory executors
//   2. host: spanning tree (first in-arc) -> pt_lca build -> cycle counting -> bridges + component tops      (the pt_net_bridges kernels)
//   3. guest: pt_lca build; t_v = LCA_T(leftmost, rightmost guest leaf among the labels below v) by two range queries over the host
//      preorder (pt_rmq machinery); cluster test = equal leaf counts
//   4. single-node components: parent_T(t_w) == t_v for every child w
//   5. the other components ("blobs") become a batch of small instances -- blob nodes + one pendant leaf per outgoing bridge, guest =
//      the subtree of T induced by the pendants' guest nodes (stack sweep over their preorder, one LCA per consecutive pair) -- and
//      are solved by the shared-memory executor (the persistent warp kernel); an instance is displayed iff no test and no blob fails
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace ptgpu {

// every instance's ranges must lie inside the arrays (their lengths are the last entries of the offset tables) and follow each other
// without gaps or overlaps: the concatenation below writes by these offsets.  out[0..2] = the three lengths, out[3] = 1 if not
__global__ void sp_offsets_check_kernel(int64_t B, const int64_t* __restrict__ hnode_off, const int64_t* __restrict__ harc_off, const int64_t* __restrict__ gnode_off, int64_t* __restrict__ out) {
  const int64_t NH = hnode_off[B], MH = harc_off[B], NG = gnode_off[B];
  if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = NH; out[1] = MH; out[2] = NG; if (hnode_off[0] != 0 || harc_off[0] != 0 || gnode_off[0] != 0 || NH < 0 || MH < 0 || NG < 0) out[3] = 1; }
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x)
    if (hnode_off[i + 1] < hnode_off[i] || harc_off[i + 1] < harc_off[i] || gnode_off[i + 1] < gnode_off[i] || hnode_off[i] < 0 || harc_off[i] < 0 || gnode_off[i] < 0 ||
        hnode_off[i + 1] > NH || harc_off[i + 1] > MH || gnode_off[i + 1] > NG)
      out[3] = 1;
}

template <class S>
__global__ void sp_concat_kernel(int64_t B, const int64_t* __restrict__ hnode_off, const int64_t* __restrict__ harc_off, const int64_t* __restrict__ gnode_off,
                                 const S* __restrict__ succ_off, const S* __restrict__ succ_adj, const S* __restrict__ hlabel, const S* __restrict__ tparent, const S* __restrict__ tlabel,
                                 int32_t* __restrict__ so, int32_t* __restrict__ sa, int32_t* __restrict__ hl, int32_t* __restrict__ tp, int32_t* __restrict__ tl,
                                 int32_t* __restrict__ inst_h, int32_t* __restrict__ inst_t, int32_t* __restrict__ flags) {
  const int64_t i = blockIdx.y;
  const int64_t hb = hnode_off[i], n = hnode_off[i + 1] - hb, ab = harc_off[i], m = harc_off[i + 1] - ab, gb = gnode_off[i], nt = gnode_off[i + 1] - gb;
  const int64_t NG = gnode_off[B];
  const int32_t labbase = (int32_t)(hb + gb);
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = tid; v < n; v += nth) {
    const int32_t o = (int32_t)succ_off[hb + i + v], o1 = (int32_t)succ_off[hb + i + v + 1];
    if (o < 0 || o1 < o || o1 > m) flags[0] = 1;
    so[hb + v] = (int32_t)(ab + o);
    const int32_t l = (int32_t)hlabel[hb + v];
    if (l >= n + nt) flags[0] = 1;
    hl[hb + v] = l < 0 ? -1 : labbase + l;
    inst_h[hb + v] = (int32_t)i;
  }
  for (int64_t e = tid; e < m; e += nth) { const int32_t h = (int32_t)succ_adj[ab + e]; if (h < 0 || h >= n) flags[0] = 1; sa[ab + e] = (int32_t)(hb + (h < 0 ? 0 : (h >= n ? n - 1 : h))); }
  for (int64_t t = tid; t < nt; t += nth) {
    const int32_t p = (int32_t)tparent[gb + t], l = (int32_t)tlabel[gb + t];
    if (p >= nt || l >= n + nt) flags[0] = 1;
    tp[gb + t] = p < 0 ? (int32_t)NG : (int32_t)(gb + p);
    tl[gb + t] = l < 0 ? -1 : labbase + l;
    inst_t[gb + t] = (int32_t)i;
  }
}
// degrees, roots, label tables (l2h / l2t over the global label space); flags[0] bad graph, [1] not clean (fall back)
__global__ void sp_degrees_kernel(int64_t NH, int64_t MH, const int32_t* __restrict__ so, const int32_t* __restrict__ sa, int32_t* __restrict__ indeg) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < MH; e += (int64_t)gridDim.x * blockDim.x) atomicAdd(&indeg[sa[e]], 1);
}
__global__ void sp_host_tables_kernel(int64_t NH, const int32_t* __restrict__ so, const int32_t* __restrict__ indeg, const int32_t* __restrict__ hl, const int32_t* __restrict__ inst_h,
                                      int32_t* __restrict__ l2h, int32_t* __restrict__ root_of, int32_t* __restrict__ nroots, int32_t* __restrict__ flags) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < NH; v += (int64_t)gridDim.x * blockDim.x) {
    const bool leaf = so[v + 1] == so[v];
    if (indeg[v] == 0) { root_of[inst_h[v]] = (int32_t)v; atomicAdd(&nroots[inst_h[v]], 1); }
    if (leaf) { if (hl[v] < 0) flags[1] = 1; else if (atomicCAS(&l2h[hl[v]], -1, (int32_t)v) != -1) flags[1] = 1; }
  }
}
__global__ void sp_guest_tables_kernel(int64_t NG, const int32_t* __restrict__ tp, const int32_t* __restrict__ tl, int32_t* __restrict__ tdeg) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < NG; t += (int64_t)gridDim.x * blockDim.x) atomicAdd(&tdeg[tp[t]], 1);
}
__global__ void sp_guest_check_kernel(int64_t B, int64_t NG, const int64_t* __restrict__ gnode_off, const int32_t* __restrict__ tp, const int32_t* __restrict__ tl, const int32_t* __restrict__ tdeg,
                                      const int32_t* __restrict__ inst_t, int32_t* __restrict__ l2t, int32_t* __restrict__ troots, int32_t* __restrict__ flags) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < NG; t += (int64_t)gridDim.x * blockDim.x) {
    if (tp[t] == (int32_t)NG) atomicAdd(&troots[inst_t[t]], 1);
    if (tdeg[t] == 1) flags[1] = 1;                                     // unary guest node: not clean
    if (tdeg[t] == 0) { if (tl[t] < 0) flags[1] = 1; else if (atomicCAS(&l2t[tl[t]], -1, (int32_t)t) != -1) flags[1] = 1; }
  }
}
__global__ void sp_labels_check_kernel(int64_t NL, const int32_t* __restrict__ l2h, const int32_t* __restrict__ l2t, int32_t* __restrict__ flags) {
  for (int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; l < NL; l += (int64_t)gridDim.x * blockDim.x) if ((l2h[l] >= 0) != (l2t[l] >= 0)) flags[1] = 1;
}
__global__ void sp_roots_check_kernel(int64_t B, int64_t NH, int64_t MH, const int32_t* __restrict__ root_of, const int32_t* __restrict__ nroots, const int32_t* __restrict__ troots,
                                      int32_t* __restrict__ so, int32_t* __restrict__ sa, int32_t* __restrict__ flags) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
    if (nroots[i] != 1 || troots[i] != 1) { flags[0] = 1; sa[MH + i] = 0; } else sa[MH + i] = root_of[i];
    if (i == 0) { so[NH] = (int32_t)MH; so[NH + 1] = (int32_t)(MH + B); }
  }
}

// range minimum of int32 values over [l, r] (inclusive) through the pt_rmq structure; returns the index of the rightmost minimum
__device__ __forceinline__ int64_t rmq_one(const LcaRecord* __restrict__ rec, const u64* __restrict__ keypos, const u64* __restrict__ table, const int64_t* __restrict__ level_base, int64_t l, int64_t r) {
  const LcaRecord a = rec[l], b = rec[r];
  const uint32_t bl = (uint32_t)(l >> 5), br = (uint32_t)(r >> 5);
  u64 best;
  if (bl == br) { const uint32_t m = b.mask >> (l & 31); best = keypos[((int64_t)bl << 5) + (l & 31) + __ffs(m) - 1]; }
  else {
    best = a.suf < b.pre ? a.suf : b.pre;
    if (br - bl >= 2) { const uint32_t span = br - bl - 1; const int k = 31 - __clz(span); const u64* lvl = table + level_base[k]; const u64 t1 = lvl[bl + 1], t2 = lvl[br - (1u << k)]; const u64 t = t1 < t2 ? t1 : t2; best = t < best ? t : best; }
  }
  return (int64_t)(0xFFFFFFFFu - (uint32_t)(best & 0xffffffffull));
}

// A[host position] = guest preorder position of the leaf's label (min array: INT_MAX elsewhere; max array holds the negation)
__global__ void sp_leafpos_kernel(int64_t N, const int32_t* __restrict__ nodeat_h, const int32_t* __restrict__ so, const int32_t* __restrict__ hl, const int32_t* __restrict__ l2t,
                                  const LcaRecord* __restrict__ grec, int32_t* __restrict__ amin, int32_t* __restrict__ amax, int32_t* __restrict__ isleaf_h) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < N; p += (int64_t)gridDim.x * blockDim.x) {
    const int32_t v = nodeat_h[p];
    int32_t lo = INT32_MAX, hi = INT32_MAX, leaf = 0;
    if (v < N - 1 && so[v + 1] == so[v] && hl[v] >= 0 && l2t[hl[v]] >= 0) { const int32_t tp = (int32_t)grec[l2t[hl[v]]].pos; lo = tp; hi = -tp; leaf = 1; }
    amin[p] = lo; amax[p] = hi; isleaf_h[p] = leaf;
  }
}
__global__ void sp_tleaf_kernel(int64_t NT, const int32_t* __restrict__ nodeat_t, const int32_t* __restrict__ tdeg, int32_t* __restrict__ isleaf_t) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < NT; p += (int64_t)gridDim.x * blockDim.x) isleaf_t[p] = tdeg[nodeat_t[p]] == 0 ? 1 : 0;
}
// component tops: t_v and the cluster test (leaf counts); tmap[v] = guest node of top v, -1 for the other nodes
__global__ void sp_tops_kernel(int64_t N, const int32_t* __restrict__ up, const LcaRecord* __restrict__ hrec, const int32_t* __restrict__ hsize, const int32_t* __restrict__ leafpre_h,
                               const LcaRecord* __restrict__ rmin, const u64* __restrict__ kmin, const u64* __restrict__ tmin, const int64_t* __restrict__ lbmin,
                               const LcaRecord* __restrict__ rmax, const u64* __restrict__ kmax, const u64* __restrict__ tmax, const int64_t* __restrict__ lbmax,
                               const int32_t* __restrict__ amin, const LcaRecord* __restrict__ grec, const u64* __restrict__ gkeypos, const u64* __restrict__ gtable, const int64_t* __restrict__ glb,
                               const int32_t* __restrict__ nodeat_t, const int32_t* __restrict__ gsize, const int32_t* __restrict__ leafpre_t, const int32_t* __restrict__ inst_h, int64_t NHreal,
                               int32_t* __restrict__ tmap, int32_t* __restrict__ fail) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < N; v += (int64_t)gridDim.x * blockDim.x) {
    if (up[v] != (int32_t)v) { tmap[v] = -1; continue; }
    const int64_t l = hrec[v].pos, r = l + hsize[v] - 1;
    const int64_t imin = rmq_one(rmin, kmin, tmin, lbmin, l, r), imax = rmq_one(rmax, kmax, tmax, lbmax, l, r);
    const int32_t plo = amin[imin], phi = amin[imax];
    const int32_t t = lca_one(grec, gkeypos, gtable, glb, nodeat_t[plo], nodeat_t[phi]);
    tmap[v] = t;
    const int32_t cnth = leafpre_h[r + 1] - leafpre_h[l];
    const int64_t tl = grec[t].pos, tr = tl + gsize[t];
    const int32_t cntt = leafpre_t[tr] - leafpre_t[tl];
    if (cnth != cntt && v < NHreal) fail[inst_h[v]] = 1;
  }
}
__global__ void sp_comp_sizes_kernel(int64_t N, const int32_t* __restrict__ up, int32_t* __restrict__ nb) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < N; v += (int64_t)gridDim.x * blockDim.x) atomicAdd(&nb[up[v]], 1);
}
// single-node components with >= 2 children: the children's guest nodes must hang on the node's own guest node
__global__ void sp_trivial_kernel(int64_t NHreal, const int32_t* __restrict__ so, const int32_t* __restrict__ sa, const int32_t* __restrict__ up, const int32_t* __restrict__ nb,
                                  const int32_t* __restrict__ tmap, const int32_t* __restrict__ tp, const int32_t* __restrict__ inst_h, int32_t* __restrict__ fail) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < NHreal; v += (int64_t)gridDim.x * blockDim.x) {
    if (up[v] != (int32_t)v || nb[v] != 1) continue;
    const int32_t d = so[v + 1] - so[v];
    if (d < 2) continue;
    const int32_t t = tmap[v];
    for (int32_t e = so[v]; e < so[v + 1]; ++e) if (tp[tmap[sa[e]]] != t) { fail[inst_h[v]] = 1; break; }
  }
}
// nodes of non-trivial components: sort keys (component top << 32 | node)
__global__ void sp_blob_keys_kernel(int64_t NHreal, const int32_t* __restrict__ up, const int32_t* __restrict__ nb, u64* __restrict__ keys, int32_t* __restrict__ count) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < NHreal; v += (int64_t)gridDim.x * blockDim.x)
    if (nb[up[v]] >= 2) keys[atomicAdd(count, 1)] = ((u64)(uint32_t)up[v] << 32) | (u64)(uint32_t)v;
}
// per sorted blob node j: out-degree; local_of[v] = j
__global__ void sp_blob_deg_kernel(int32_t nbn, const u64* __restrict__ keys, const int32_t* __restrict__ so, int32_t* __restrict__ odeg, int32_t* __restrict__ local_of) {
  for (int32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < nbn; j += gridDim.x * blockDim.x) {
    const int32_t v = (int32_t)(keys[j] & 0xffffffffull);
    odeg[j] = so[v + 1] - so[v]; local_of[v] = j;
  }
}
// the arcs of the blob nodes in a compact list: an internal arc names the sorted index of its head, a bridge the guest node of the
// pendant part it leads to (with its guest preorder position); a blob node that is itself a labelled leaf names its guest leaf
__global__ void sp_blob_arcs_kernel(int32_t nbn, const u64* __restrict__ keys, const int32_t* __restrict__ so, const int32_t* __restrict__ sa, const uint8_t* __restrict__ is_bridge,
                                    const int32_t* __restrict__ odeg_pre, const int32_t* __restrict__ local_of, const int32_t* __restrict__ tmap, const LcaRecord* __restrict__ grec,
                                    const int32_t* __restrict__ hl, const int32_t* __restrict__ l2t, const int32_t* __restrict__ inst_h,
                                    int32_t* __restrict__ arc_head, int32_t* __restrict__ arc_gnode, int32_t* __restrict__ arc_gpos, int32_t* __restrict__ node_gnode, int32_t* __restrict__ node_gpos,
                                    int32_t* __restrict__ node_inst) {
  for (int32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < nbn; j += gridDim.x * blockDim.x) {
    const int32_t v = (int32_t)(keys[j] & 0xffffffffull);
    int32_t pos = odeg_pre[j];
    for (int32_t e = so[v]; e < so[v + 1]; ++e, ++pos) {
      const int32_t w = sa[e];
      if (is_bridge[e]) { const int32_t t = tmap[w]; arc_head[pos] = -1; arc_gnode[pos] = t; arc_gpos[pos] = (int32_t)grec[t].pos; }
      else { arc_head[pos] = local_of[w]; arc_gnode[pos] = -1; arc_gpos[pos] = -1; }
    }
    int32_t gn = -1, gp = -1;
    if (so[v + 1] == so[v] && hl[v] >= 0 && l2t[hl[v]] >= 0) { gn = l2t[hl[v]]; gp = (int32_t)grec[gn].pos; }
    node_gnode[j] = gn; node_gpos[j] = gp; node_inst[j] = inst_h[v];
  }
}
__global__ void sp_node_posdepth_kernel(int64_t q, const int32_t* __restrict__ nodes, const LcaRecord* __restrict__ grec, int32_t* __restrict__ pos, int32_t* __restrict__ depth) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < q; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t x = nodes[i];
    if (x < 0) { pos[i] = -1; depth[i] = -1; continue; }
    pos[i] = (int32_t)grec[x].pos; depth[i] = (int32_t)(grec[x].key >> 32);
  }
}
__global__ void sp_result_kernel(int64_t B, const int32_t* __restrict__ fail, uint32_t* __restrict__ bits) {
  for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < (B + 31) / 32; w += (int64_t)gridDim.x * blockDim.x) {
    uint32_t x = 0;
    for (int k = 0; k < 32; ++k) { const int64_t i = w * 32 + k; if (i < B && !fail[i]) x |= 1u << k; }
    bits[w] = x;
  }
}

}  // namespace ptgpu

// returns PT_OK with *applied = 1 when the batch was decided here (bits / status written, DEVICE memory), *applied = 0 when the
// batch is not clean enough for the split path (the caller runs the global-memory executors), or an error
extern "C" int ptgpu_split_contain(int64_t B, const int64_t* d_hnode_off, const int64_t* d_harc_off, const int64_t* d_gnode_off, const void* d_succ_off, const void* d_succ_adj,
                                   const void* d_hlabel, const void* d_tparent, const void* d_tlabel, int idx_bytes, uint32_t* d_bits, uint8_t* d_status, uint64_t* blobs_solved,
                                   int* applied, void* stream) {
  using namespace ptgpu;
  *applied = 0;
  if (blobs_solved) *blobs_solved = 0;
  if (idx_bytes != 4 && idx_bytes != 2) return PT_OK;
  cudaStream_t s = (cudaStream_t)stream;
  DeviceProps dp;
  if (int rc = device_props(&dp)) return rc;
  std::vector<void*> owned;
  pt_lca *H = nullptr, *G = nullptr; pt_rmq *Rmin = nullptr, *Rmax = nullptr;
  auto cleanup = [&]() { cudaStreamSynchronize(s); for (void* p : owned) dev_free(p); if (H) pt_lca_free(H); if (G) pt_lca_free(G); if (Rmin) pt_rmq_free(Rmin); if (Rmax) pt_rmq_free(Rmax); };
  auto dalloc = [&](size_t bytes) -> void* { void* p = nullptr; if (lca_malloc(&p, bytes ? bytes : 4) != cudaSuccess) { cudaGetLastError(); return nullptr; } owned.push_back(p); return p; };
#define SP_TRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup(); return set_error(PT_ERR_CUDA, "%s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); } } while (0)
#define SP_PTR(p) do { if (!(p)) { cleanup(); return set_error(PT_ERR_NOMEM, "split path: out of device memory"); } } while (0)
  // PT_SPLIT_TRACE=1: wall time of every phase on stderr (a synchronise per mark: a profiling aid, not the normal path)
  static const bool trace = getenv("PT_SPLIT_TRACE") != nullptr;
  auto t_last = std::chrono::steady_clock::now();
  auto mark = [&](const char* what) {
    if (!trace) return;
    cudaStreamSynchronize(s);
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[split] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
    t_last = now;
  };
  int64_t tails[4] = {0, 0, 0, 0};
  int64_t* d_tails = (int64_t*)dalloc(32);
  SP_PTR(d_tails);
  SP_TRY(cudaMemsetAsync(d_tails, 0, 32, s));
  sp_offsets_check_kernel<<<std::max(1, (int)std::min<int64_t>(dp.sms * 4, (B + 255) / 256)), 256, 0, s>>>(B, d_hnode_off, d_harc_off, d_gnode_off, d_tails);
  SP_TRY(cudaMemcpyAsync(tails, d_tails, 32, cudaMemcpyDeviceToHost, s));
  SP_TRY(cudaStreamSynchronize(s));
  if (tails[3]) { cleanup(); return PT_OK; }   // malformed offset tables: the executors report the instances concerned (PT_INST_BAD_GRAPH)
  const int64_t NH = tails[0], MH = tails[1], NG = tails[2], N = NH + 1, M = MH + B, NT = NG + 1, NL = NH + NG;
  if (N > INT32_MAX / 2 || M > INT32_MAX / 2 || NT > INT32_MAX / 2 || B > 65535) { cleanup(); return PT_OK; }
  const int grid = dp.sms * 8;
  int32_t* so = (int32_t*)dalloc((size_t)(N + 1) * 4); int32_t* sa = (int32_t*)dalloc((size_t)M * 4); int32_t* hl = (int32_t*)dalloc((size_t)N * 4);
  int32_t* tp = (int32_t*)dalloc((size_t)NT * 4); int32_t* tl = (int32_t*)dalloc((size_t)NT * 4); int32_t* inst_h = (int32_t*)dalloc((size_t)N * 4); int32_t* inst_t = (int32_t*)dalloc((size_t)NT * 4);
  int32_t* indeg = (int32_t*)dalloc((size_t)N * 4); int32_t* tdeg = (int32_t*)dalloc((size_t)(NT + 1) * 4); int32_t* l2h = (int32_t*)dalloc((size_t)NL * 4); int32_t* l2t = (int32_t*)dalloc((size_t)NL * 4);
  int32_t* root_of = (int32_t*)dalloc((size_t)B * 4); int32_t* nroots = (int32_t*)dalloc((size_t)B * 4); int32_t* troots = (int32_t*)dalloc((size_t)B * 4); int32_t* fail = (int32_t*)dalloc((size_t)B * 4);
  int32_t* flags = (int32_t*)dalloc(64);
  SP_PTR(so); SP_PTR(sa); SP_PTR(hl); SP_PTR(tp); SP_PTR(tl); SP_PTR(inst_h); SP_PTR(inst_t); SP_PTR(indeg); SP_PTR(tdeg); SP_PTR(l2h); SP_PTR(l2t); SP_PTR(root_of); SP_PTR(nroots);
  SP_PTR(troots); SP_PTR(fail); SP_PTR(flags);
  SP_TRY(cudaMemsetAsync(indeg, 0, (size_t)N * 4, s)); SP_TRY(cudaMemsetAsync(tdeg, 0, (size_t)(NT + 1) * 4, s)); SP_TRY(cudaMemsetAsync(l2h, 0xff, (size_t)NL * 4, s));
  SP_TRY(cudaMemsetAsync(l2t, 0xff, (size_t)NL * 4, s)); SP_TRY(cudaMemsetAsync(nroots, 0, (size_t)B * 4, s)); SP_TRY(cudaMemsetAsync(troots, 0, (size_t)B * 4, s));
  SP_TRY(cudaMemsetAsync(fail, 0, (size_t)B * 4, s)); SP_TRY(cudaMemsetAsync(flags, 0, 64, s)); SP_TRY(cudaMemsetAsync(root_of, 0, (size_t)B * 4, s));
  {
    const dim3 g((unsigned)std::max<int64_t>(1, std::min<int64_t>(dp.sms * 8 / std::max<int64_t>(B, 1) + 1, 1024)), (unsigned)B);
    if (idx_bytes == 2) sp_concat_kernel<int16_t><<<g, 256, 0, s>>>(B, d_hnode_off, d_harc_off, d_gnode_off, (const int16_t*)d_succ_off, (const int16_t*)d_succ_adj, (const int16_t*)d_hlabel, (const int16_t*)d_tparent, (const int16_t*)d_tlabel, so, sa, hl, tp, tl, inst_h, inst_t, flags);
    else sp_concat_kernel<int32_t><<<g, 256, 0, s>>>(B, d_hnode_off, d_harc_off, d_gnode_off, (const int32_t*)d_succ_off, (const int32_t*)d_succ_adj, (const int32_t*)d_hlabel, (const int32_t*)d_tparent, (const int32_t*)d_tlabel, so, sa, hl, tp, tl, inst_h, inst_t, flags);
  }
  const int32_t superlab = -1, minus1 = -1; (void)superlab;
  SP_TRY(cudaMemcpyAsync(hl + NH, &minus1, 4, cudaMemcpyHostToDevice, s)); SP_TRY(cudaMemcpyAsync(tp + NG, &minus1, 4, cudaMemcpyHostToDevice, s)); SP_TRY(cudaMemcpyAsync(tl + NG, &minus1, 4, cudaMemcpyHostToDevice, s));
  sp_degrees_kernel<<<grid, 256, 0, s>>>(NH, MH, so, sa, indeg);
  sp_host_tables_kernel<<<grid, 256, 0, s>>>(NH, so, indeg, hl, inst_h, l2h, root_of, nroots, flags);
  sp_guest_tables_kernel<<<grid, 256, 0, s>>>(NG, tp, tl, tdeg);
  sp_guest_check_kernel<<<grid, 256, 0, s>>>(B, NG, d_gnode_off, tp, tl, tdeg, inst_t, l2t, troots, flags);
  sp_labels_check_kernel<<<grid, 256, 0, s>>>(NL, l2h, l2t, flags);
  sp_roots_check_kernel<<<grid, 256, 0, s>>>(B, NH, MH, root_of, nroots, troots, so, sa, flags);
  int32_t hf[16];
  SP_TRY(cudaMemcpyAsync(hf, flags, 64, cudaMemcpyDeviceToHost, s));
  SP_TRY(cudaStreamSynchronize(s));
  mark("concatenate + tables");
  if (hf[0] || hf[1]) { cleanup(); return PT_OK; }   // malformed or not clean: the general executors decide (and report per-instance statuses)
  // ---- host side: spanning tree, LCA structure, bridges, component tops (the kernels of pt_net_bridges) ----
  int32_t* tree_arc = (int32_t*)dalloc((size_t)N * 4); int32_t* arc_tail = (int32_t*)dalloc((size_t)M * 4); int32_t* parent = (int32_t*)dalloc((size_t)N * 4);
  int32_t* cntpos = (int32_t*)dalloc((size_t)(N + 1) * 4); int32_t* prefix = (int32_t*)dalloc((size_t)(N + 1) * 4); int32_t* up = (int32_t*)dalloc((size_t)N * 4);
  uint8_t* is_bridge = (uint8_t*)dalloc((size_t)M + 1);
  int64_t* scan_tmp = (int64_t*)dalloc(((size_t)((std::max(N, NT) + SCAN_TILE - 1) / SCAN_TILE) + 2) * 8);
  SP_PTR(tree_arc); SP_PTR(arc_tail); SP_PTR(parent); SP_PTR(cntpos); SP_PTR(prefix); SP_PTR(up); SP_PTR(is_bridge); SP_PTR(scan_tmp);
  SP_TRY(cudaMemsetAsync(tree_arc, 0x7f, (size_t)N * 4, s)); SP_TRY(cudaMemsetAsync(parent, 0xff, (size_t)N * 4, s)); SP_TRY(cudaMemsetAsync(cntpos, 0, (size_t)(N + 1) * 4, s));
  SP_TRY(cudaMemsetAsync(is_bridge, 0, (size_t)M + 1, s));
  br_first_parent_kernel<<<grid, 256, 0, s>>>(N, M, so, sa, tree_arc);
  br_tree_parent_kernel<<<grid, 256, 0, s>>>(N, so, sa, tree_arc, arc_tail, parent);
  if (int rc = lca_build_impl(&H, parent, N, PT_MEM_DEVICE, stream, true)) { cleanup(); if (rc == PT_ERR_INVALID) return PT_OK; return rc; }   // a cycle: the general path reports it
  mark("host spanning tree + LCA");
  br_cover_kernel<<<grid, 256, 0, s>>>(M, arc_tail, sa, tree_arc, H->rec, H->keypos, H->table, H->d_level_base, cntpos);
  if (int rc = device_exclusive_scan(cntpos, N, prefix, scan_tmp, s)) { cleanup(); return rc; }
  br_mark_kernel<<<grid, 256, 0, s>>>(N, parent, tree_arc, H->rec, H->size, prefix, is_bridge, up);
  for (int it = 0; it < 40; ++it) {
    int32_t h = 0;
    SP_TRY(cudaMemsetAsync(flags, 0, 4, s));
    br_jump_kernel<<<grid, 256, 0, s>>>(N, up, flags);
    SP_TRY(cudaMemcpyAsync(&h, flags, 4, cudaMemcpyDeviceToHost, s));
    SP_TRY(cudaStreamSynchronize(s));
    if (!h) break;
  }
  mark("bridges + components");
  // ---- guest: LCA structure; guest positions of the labels along the host preorder; cluster tests ----
  if (int rc = lca_build_impl(&G, tp, NT, PT_MEM_DEVICE, stream, true)) { cleanup(); if (rc == PT_ERR_INVALID) return PT_OK; return rc; }
  mark("guest LCA");
  int32_t* amin = (int32_t*)dalloc((size_t)N * 4); int32_t* amax = (int32_t*)dalloc((size_t)N * 4); int32_t* leaf_h = (int32_t*)dalloc((size_t)(N + 1) * 4); int32_t* leafpre_h = (int32_t*)dalloc((size_t)(N + 1) * 4);
  int32_t* leaf_t = (int32_t*)dalloc((size_t)(NT + 1) * 4); int32_t* leafpre_t = (int32_t*)dalloc((size_t)(NT + 1) * 4); int32_t* tmap = (int32_t*)dalloc((size_t)N * 4); int32_t* nb = (int32_t*)dalloc((size_t)N * 4);
  SP_PTR(amin); SP_PTR(amax); SP_PTR(leaf_h); SP_PTR(leafpre_h); SP_PTR(leaf_t); SP_PTR(leafpre_t); SP_PTR(tmap); SP_PTR(nb);
  sp_leafpos_kernel<<<grid, 256, 0, s>>>(N, H->nodeat, so, hl, l2t, G->rec, amin, amax, leaf_h);
  sp_tleaf_kernel<<<grid, 256, 0, s>>>(NT, G->nodeat, tdeg, leaf_t);
  if (int rc = device_exclusive_scan(leaf_h, N, leafpre_h, scan_tmp, s)) { cleanup(); return rc; }
  if (int rc = device_exclusive_scan(leaf_t, NT, leafpre_t, scan_tmp, s)) { cleanup(); return rc; }
  if (int rc = pt_rmq_build(&Rmin, amin, N, PT_MEM_DEVICE, stream)) { cleanup(); return rc; }
  if (int rc = pt_rmq_build(&Rmax, amax, N, PT_MEM_DEVICE, stream)) { cleanup(); return rc; }
  mark("leaf positions + 2 RMQ builds");
  sp_tops_kernel<<<grid, 256, 0, s>>>(N, up, H->rec, H->size, leafpre_h, Rmin->rec, Rmin->keypos, Rmin->table, Rmin->d_level_base, Rmax->rec, Rmax->keypos, Rmax->table, Rmax->d_level_base,
                                      amin, G->rec, G->keypos, G->table, G->d_level_base, G->nodeat, G->size, leafpre_t, inst_h, NH, tmap, fail);
  SP_TRY(cudaMemsetAsync(nb, 0, (size_t)N * 4, s));
  sp_comp_sizes_kernel<<<grid, 256, 0, s>>>(N, up, nb);
  sp_trivial_kernel<<<grid, 256, 0, s>>>(NH, so, sa, up, nb, tmap, tp, inst_h, fail);
  mark("cluster tests + trivial comps");
  // ---- the blobs: compact their nodes, sort by (component, node), pull their arcs ----
  u64* keys = (u64*)dalloc((size_t)N * 8); u64* keys2 = (u64*)dalloc((size_t)N * 8); int32_t* cnt = (int32_t*)dalloc(64);
  SP_PTR(keys); SP_PTR(keys2); SP_PTR(cnt);
  SP_TRY(cudaMemsetAsync(cnt, 0, 64, s));
  sp_blob_keys_kernel<<<grid, 256, 0, s>>>(NH, up, nb, keys, cnt);
  int32_t nbn = 0;
  SP_TRY(cudaMemcpyAsync(&nbn, cnt, 4, cudaMemcpyDeviceToHost, s));
  SP_TRY(cudaStreamSynchronize(s));
  uint64_t nblobs_solved = 0;
  if (nbn > 0) {
    size_t tb = 0;
    SP_TRY(cub::DeviceRadixSort::SortKeys(nullptr, tb, keys, keys2, nbn, 0, 64, s));
    void* tmp = dalloc(tb); SP_PTR(tmp);
    SP_TRY(cub::DeviceRadixSort::SortKeys(tmp, tb, keys, keys2, nbn, 0, 64, s));
    int32_t* odeg = (int32_t*)dalloc((size_t)(nbn + 1) * 4); int32_t* odeg_pre = (int32_t*)dalloc((size_t)(nbn + 1) * 4); int32_t* local_of = (int32_t*)dalloc((size_t)N * 4);
    SP_PTR(odeg); SP_PTR(odeg_pre); SP_PTR(local_of);
    sp_blob_deg_kernel<<<grid, 256, 0, s>>>(nbn, keys2, so, odeg, local_of);
    if (int rc = device_exclusive_scan(odeg, nbn, odeg_pre, scan_tmp, s)) { cleanup(); return rc; }
    int32_t narcs = 0;
    SP_TRY(cudaMemcpyAsync(&narcs, odeg_pre + nbn, 4, cudaMemcpyDeviceToHost, s));
    SP_TRY(cudaStreamSynchronize(s));
    int32_t* arc_head = (int32_t*)dalloc((size_t)std::max(narcs, 1) * 4); int32_t* arc_gnode = (int32_t*)dalloc((size_t)std::max(narcs, 1) * 4); int32_t* arc_gpos = (int32_t*)dalloc((size_t)std::max(narcs, 1) * 4);
    int32_t* node_gnode = (int32_t*)dalloc((size_t)nbn * 4); int32_t* node_gpos = (int32_t*)dalloc((size_t)nbn * 4); int32_t* node_inst = (int32_t*)dalloc((size_t)nbn * 4);
    SP_PTR(arc_head); SP_PTR(arc_gnode); SP_PTR(arc_gpos); SP_PTR(node_gnode); SP_PTR(node_gpos); SP_PTR(node_inst);
    sp_blob_arcs_kernel<<<grid, 256, 0, s>>>(nbn, keys2, so, sa, is_bridge, odeg_pre, local_of, tmap, G->rec, hl, l2t, inst_h, arc_head, arc_gnode, arc_gpos, node_gnode, node_gpos, node_inst);
    std::vector<u64> hkeys((size_t)nbn); std::vector<int32_t> hpre((size_t)nbn + 1), hhead((size_t)narcs), hgnode((size_t)narcs), hgpos((size_t)narcs), hngn((size_t)nbn), hngp((size_t)nbn), hinst((size_t)nbn);
    SP_TRY(cudaMemcpyAsync(hkeys.data(), keys2, (size_t)nbn * 8, cudaMemcpyDeviceToHost, s)); SP_TRY(cudaMemcpyAsync(hpre.data(), odeg_pre, (size_t)(nbn + 1) * 4, cudaMemcpyDeviceToHost, s));
    if (narcs) { SP_TRY(cudaMemcpyAsync(hhead.data(), arc_head, (size_t)narcs * 4, cudaMemcpyDeviceToHost, s)); SP_TRY(cudaMemcpyAsync(hgnode.data(), arc_gnode, (size_t)narcs * 4, cudaMemcpyDeviceToHost, s));
                 SP_TRY(cudaMemcpyAsync(hgpos.data(), arc_gpos, (size_t)narcs * 4, cudaMemcpyDeviceToHost, s)); }
    SP_TRY(cudaMemcpyAsync(hngn.data(), node_gnode, (size_t)nbn * 4, cudaMemcpyDeviceToHost, s)); SP_TRY(cudaMemcpyAsync(hngp.data(), node_gpos, (size_t)nbn * 4, cudaMemcpyDeviceToHost, s));
    SP_TRY(cudaMemcpyAsync(hinst.data(), node_inst, (size_t)nbn * 4, cudaMemcpyDeviceToHost, s));
    SP_TRY(cudaStreamSynchronize(s));
    mark("blob compaction + download");
    if (trace) fprintf(stderr, "[split] %d blob nodes, %d arcs\n", nbn, narcs);
    // ---- host: one small instance per blob (the blobs together hold a few thousand nodes) ----
    struct Cand { int32_t gpos, gnode, label; };
    std::vector<int32_t> bstart;   // first sorted index of every blob
    for (int32_t j = 0; j < nbn; ++j) if (j == 0 || (hkeys[(size_t)j] >> 32) != (hkeys[(size_t)j - 1] >> 32)) bstart.push_back(j);
    const size_t nblobs = bstart.size();
    bstart.push_back(nbn);
    std::vector<int64_t> b_hn{0}, b_ha{0}, b_gn{0};
    std::vector<int32_t> b_so, b_sa, b_hl, b_gp, b_gl, b_inst;
    std::vector<std::vector<Cand>> cands(nblobs);
    std::vector<int32_t> cand_nodes;   // all candidates' guest nodes, blob after blob, each blob sorted by guest preorder
    std::vector<size_t> cand_off{0};
    for (size_t b = 0; b < nblobs; ++b) {
      const int32_t j0 = bstart[b], j1 = bstart[b + 1], nreal = j1 - j0;
      int32_t pend = 0;
      for (int32_t a = hpre[(size_t)j0]; a < hpre[(size_t)j1]; ++a) pend += hhead[(size_t)a] < 0;
      const int32_t nprime = nreal + pend, mprime = hpre[(size_t)j1] - hpre[(size_t)j0];
      const size_t so0 = b_so.size(), sa0 = b_sa.size(), hl0 = b_hl.size();
      b_so.resize(so0 + (size_t)nprime + 1); b_sa.resize(sa0 + (size_t)mprime); b_hl.resize(hl0 + (size_t)nprime, -1);
      int32_t pk = 0, lab = 0;
      std::vector<Cand>& C = cands[b];
      for (int32_t j = j0; j < j1; ++j) {
        b_so[so0 + (size_t)(j - j0)] = hpre[(size_t)j] - hpre[(size_t)j0];
        for (int32_t a = hpre[(size_t)j]; a < hpre[(size_t)j + 1]; ++a) {
          int32_t head;
          if (hhead[(size_t)a] < 0) { head = nreal + pk; b_hl[hl0 + (size_t)head] = lab; C.push_back({hgpos[(size_t)a], hgnode[(size_t)a], lab}); ++lab; ++pk; }
          else head = hhead[(size_t)a] - j0;
          b_sa[sa0 + (size_t)(a - hpre[(size_t)j0])] = head;
        }
        if (hngn[(size_t)j] >= 0) { b_hl[hl0 + (size_t)(j - j0)] = lab; C.push_back({hngp[(size_t)j], hngn[(size_t)j], lab}); ++lab; }
      }
      for (int32_t k = nreal; k <= nprime; ++k) b_so[so0 + (size_t)k] = mprime;
      std::sort(C.begin(), C.end(), [](const Cand& x, const Cand& y) { return x.gpos < y.gpos; });
      for (const Cand& c : C) cand_nodes.push_back(c.gnode);
      cand_off.push_back(cand_nodes.size());
      b_hn.push_back(b_hn.back() + nprime); b_ha.push_back(b_ha.back() + mprime);
      b_inst.push_back(hinst[(size_t)j0]);
    }
    // LCAs of consecutive candidates (utils/induced_tree.hpp:128-138), their positions and depths: one batched device call
    const size_t nc = cand_nodes.size();
    std::vector<int32_t> lnode(std::max<size_t>(nc, 1)), lpos(std::max<size_t>(nc, 1)), ldepth(std::max<size_t>(nc, 1)), cdepth(std::max<size_t>(nc, 1));
    if (nc >= 1) {
      int32_t* d_c = (int32_t*)dalloc(nc * 4); int32_t* d_l = (int32_t*)dalloc(nc * 4); int32_t* d_p = (int32_t*)dalloc(nc * 4); int32_t* d_d = (int32_t*)dalloc(nc * 4);
      SP_PTR(d_c); SP_PTR(d_l); SP_PTR(d_p); SP_PTR(d_d);
      SP_TRY(cudaMemcpyAsync(d_c, cand_nodes.data(), nc * 4, cudaMemcpyHostToDevice, s));
      SP_TRY(cudaMemsetAsync(d_l, 0xff, nc * 4, s));
      if (nc >= 2) { if (int rc = pt_lca_query_adjacent(G, d_c, (int64_t)nc, d_l, PT_MEM_DEVICE, stream)) { cleanup(); return rc; } }
      sp_node_posdepth_kernel<<<std::max(1, (int)std::min<size_t>((nc + 255) / 256, 1024)), 256, 0, s>>>((int64_t)nc, d_l, G->rec, d_p, d_d);
      SP_TRY(cudaMemcpyAsync(lnode.data(), d_l, nc * 4, cudaMemcpyDeviceToHost, s)); SP_TRY(cudaMemcpyAsync(lpos.data(), d_p, nc * 4, cudaMemcpyDeviceToHost, s));
      SP_TRY(cudaMemcpyAsync(ldepth.data(), d_d, nc * 4, cudaMemcpyDeviceToHost, s));
      sp_node_posdepth_kernel<<<std::max(1, (int)std::min<size_t>((nc + 255) / 256, 1024)), 256, 0, s>>>((int64_t)nc, d_c, G->rec, d_p, d_d);
      SP_TRY(cudaMemcpyAsync(cdepth.data(), d_d, nc * 4, cudaMemcpyDeviceToHost, s));
      SP_TRY(cudaStreamSynchronize(s));
    }
    // guest side of every blob: the subtree of T induced by its candidates (stack sweep; node ids in creation order)
    for (size_t b = 0; b < nblobs; ++b) {
      const std::vector<Cand>& C = cands[b];
      const size_t c0 = cand_off[b], K = C.size();
      const size_t g0 = b_gp.size();
      struct VN { int32_t pos, depth, parent, label; };
      std::vector<VN> vn; std::vector<int32_t> st;
      for (size_t k = 0; k < K; ++k) {
        const int32_t w = (int32_t)vn.size();
        vn.push_back({C[k].gpos, cdepth[c0 + k], -1, C[k].label});
        if (!st.empty()) {
          const int32_t lp = lpos[c0 + k - 1], ld = ldepth[c0 + k - 1];   // LCA of candidates k-1 and k = LCA of the stack top's subtree and w
          if (lp != vn[(size_t)st.back()].pos) {
            while (st.size() >= 2 && vn[(size_t)st[st.size() - 2]].depth >= ld) { vn[(size_t)st.back()].parent = st[st.size() - 2]; st.pop_back(); }
            if (vn[(size_t)st.back()].pos != lp) { const int32_t L = (int32_t)vn.size(); vn.push_back({lp, ld, -1, -1}); vn[(size_t)st.back()].parent = L; st.back() = L; }
          }
        }
        st.push_back(w);
      }
      while (st.size() >= 2) { vn[(size_t)st.back()].parent = st[st.size() - 2]; st.pop_back(); }
      if (vn.empty()) vn.push_back({0, 0, -1, -1});
      for (const VN& x : vn) { b_gp.push_back(x.parent); b_gl.push_back(x.label); }
      b_gn.push_back((int64_t)b_gp.size());
      (void)g0;
    }
    mark("blob instances on the host");
    // solve the blobs with the shared-memory executor
    pt_batch_desc d{};
    d.num_instances = (int64_t)nblobs; d.mem = PT_MEM_HOST; d.host_node_off = b_hn.data(); d.host_arc_off = b_ha.data(); d.guest_node_off = b_gn.data();
    d.host_succ_off = b_so.data(); d.host_succ_adj = b_sa.data(); d.host_label = b_hl.data(); d.guest_parent = b_gp.data(); d.guest_label = b_gl.data(); d.index_bytes = 4;
    pt_batch* bb = nullptr;
    if (int rc = pt_batch_create(&bb, &d, stream)) { cleanup(); return rc; }
    std::vector<uint32_t> bbits((nblobs + 31) / 32 + 1, 0); std::vector<uint8_t> bstat(nblobs + 1, 0);
    pt_options bo{}; bo.path = 4;   // 4 = "do not split again" (internal): warp executor, or the global-memory ones for a huge blob
    const int rc2 = pt_batch_contain(bb, &bo, bbits.data(), bstat.data(), PT_MEM_HOST, nullptr, stream);
    pt_batch_free(bb);
    mark("blob batch solved");
    if (trace) fprintf(stderr, "[split] %zu blobs, largest host %lld nodes\n", nblobs, (long long)[&]() { int64_t mx = 0; for (size_t b = 0; b < nblobs; ++b) mx = std::max(mx, b_hn[b + 1] - b_hn[b]); return mx; }());
    if (rc2) { cleanup(); return rc2; }
    std::vector<int32_t> hfail((size_t)B, 0);
    bool any = false;
    for (size_t b = 0; b < nblobs; ++b) {
      if (bstat[b] != 0) { cleanup(); return PT_OK; }   // a blob the solver could not take (overflow ...): leave the whole batch to the general path
      if (!((bbits[b >> 5] >> (b & 31)) & 1u)) { hfail[(size_t)b_inst[b]] = 1; any = true; }
    }
    if (any) {
      int32_t* d_f2 = (int32_t*)dalloc((size_t)B * 4); SP_PTR(d_f2);
      std::vector<int32_t> cur((size_t)B);
      SP_TRY(cudaMemcpyAsync(cur.data(), fail, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
      SP_TRY(cudaStreamSynchronize(s));
      for (int64_t i = 0; i < B; ++i) cur[(size_t)i] |= hfail[(size_t)i];
      SP_TRY(cudaMemcpyAsync(fail, cur.data(), (size_t)B * 4, cudaMemcpyHostToDevice, s));
      SP_TRY(cudaStreamSynchronize(s));
    }
    nblobs_solved = nblobs;
  }
  sp_result_kernel<<<std::max(1, (int)std::min<int64_t>(((B + 31) / 32 + 255) / 256, 1024)), 256, 0, s>>>(B, fail, d_bits);
  if (d_status) SP_TRY(cudaMemsetAsync(d_status, 0, (size_t)B, s));
  SP_TRY(cudaGetLastError());
  SP_TRY(cudaStreamSynchronize(s));
#undef SP_TRY
#undef SP_PTR
  cleanup();
  if (blobs_solved) *blobs_solved = nblobs_solved;
  *applied = 1;
  return PT_OK;
}
