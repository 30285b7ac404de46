// csrc/lca_kernels.cu -- batched LCA on sm_100a.
//
// Entry points pt_lca_build / pt_lca_query / pt_lca_node_info / pt_lca_free (include/pt_gpu.h) stand in for
//   Tree::LCA / naiveLCA          utils/tree.hpp:202-223   (parent walk, O(depth) per query)
//   lca<node_t>::lca(), query()   rmq/lca.hpp:61-106       (Euler tour + depth array + RMQ)
//   sparse_rmq                    rmq/sparse_rmq.hpp:45-90 (O(n log n) index table)
// Answers are node ids, hence unique: bit-exact against either reference path.
//
// Layout.  Instead of the reference's 2n-1 Euler tour with a 26-level index table (3.9 GB at
// n = 2*10^7, seven dependent random gathers per query) the build orders nodes by PREORDER position and uses
//     lca(u,v) = parent( argmin depth over positions (pos[u], pos[v]] ),   pos[u] < pos[v]
// Positions are cut into blocks of 32.  Per NODE one 32-byte record (one DRAM sector):
//     { pos, in-block stack mask, min key over (pos, block end], min key over [block start, pos], own key }
// with key = depth << 32 | parent.  Whole blocks between the two end blocks are answered by a sparse table over
// packed block minima (n/32 entries per level, ~100 MB at n = 2*10^7: L2 resident on B200's 126 MB L2).
// A query is therefore: 8 B of (u,v) streamed, TWO random 32-byte sector reads, two table reads that hit L2,
// 4 B streamed out.
#include <cooperative_groups.h>
#include <cub/device/device_radix_sort.cuh>
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <utility>
#include <vector>
#include "cuda_common.cuh"

// Device memory comes from the caching allocator (cuda_common.cuh): raw cudaMalloc / cudaFree cost tens of milliseconds per call
// once the process holds a few GB (measured: pt_tree_who_displays 23 ms -> 151 ms after the text legs of bench.py had run).
// dev_free does not synchronise the way cudaFree does, so every release below is preceded by a stream / device synchronise.
static inline cudaError_t lca_malloc(void** p, size_t bytes) { return ptgpu::dev_alloc(p, bytes) == PT_OK ? cudaSuccess : cudaErrorMemoryAllocation; }
using ptgpu::dev_free;

namespace cg = cooperative_groups;

namespace ptgpu {

typedef unsigned long long u64;
constexpr u64 KEY_INF = ~0ull;

struct __align__(32) LcaRecord { uint32_t pos, mask; u64 suf, pre, key; };
static_assert(sizeof(LcaRecord) == 32, "one record = one 32-byte sector");

// ---- build step 1: children counts + root ---------------------------------------------------------------------
__global__ void lca_count_kernel(const int32_t* __restrict__ parent, int64_t n, int32_t* __restrict__ deg, int32_t* __restrict__ info /*[0]=nroots [1]=root [2]=bad*/) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
    const int32_t p = parent[v];
    if (p < 0) { atomicAdd(&info[0], 1); info[1] = (int32_t)v; }
    else if (p >= n || p == v) info[2] = 1;
    else atomicAdd(&deg[p], 1);
  }
}

// ---- exclusive scan (three passes, tiles of 2048) ---------------------------------------------------------------
constexpr int SCAN_TILE = 2048;
__global__ void scan_tile_sums_kernel(const int32_t* __restrict__ in, int64_t n, int64_t* __restrict__ tile_sums) {
  __shared__ int64_t wsum[8];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  int64_t s = 0;
  for (int i = threadIdx.x; i < SCAN_TILE; i += 256) { const int64_t k = base + i; if (k < n) s += in[k]; }
  for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { int64_t t = 0; for (int k = 0; k < 8; ++k) t += wsum[k]; tile_sums[blockIdx.x] = t; }
}
__global__ void scan_tile_offsets_kernel(int64_t* __restrict__ tile_sums, int64_t ntiles) {
  // single CTA: sequential over chunks of 1024 with a block scan inside
  __shared__ int64_t buf[1024];
  __shared__ int64_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t b = 0; b < ntiles; b += 1024) {
    const int64_t k = b + threadIdx.x;
    const int64_t v = k < ntiles ? tile_sums[k] : 0;
    buf[threadIdx.x] = v;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) { int64_t t = threadIdx.x >= d ? buf[threadIdx.x - d] : 0; __syncthreads(); buf[threadIdx.x] += t; __syncthreads(); }
    if (k < ntiles) tile_sums[k] = carry + buf[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += buf[1023];
    __syncthreads();
  }
}
__global__ void scan_apply_kernel(const int32_t* __restrict__ in, int64_t n, const int64_t* __restrict__ tile_off, int32_t* __restrict__ out /* n+1 */) {
  __shared__ int32_t wtot[8];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  // each thread owns 8 consecutive elements
  int32_t v[8]; int32_t s = 0;
  const int64_t t0 = base + (int64_t)threadIdx.x * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) { v[j] = (t0 + j < n) ? in[t0 + j] : 0; s += v[j]; }
  int32_t incl = s;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const int32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
  if (lane == 31) wtot[wid] = incl;
  __syncthreads();
  int32_t woff = 0;
  for (int k = 0; k < wid; ++k) woff += wtot[k];
  int64_t acc = tile_off[blockIdx.x] + woff + incl - s;
#pragma unroll
  for (int j = 0; j < 8; ++j) { if (t0 + j < n) out[t0 + j] = (int32_t)acc; acc += v[j]; }
  if (t0 <= n - 1 && n - 1 < t0 + 8) out[n] = (int32_t)acc;  // the thread holding the last element writes the total
}

// ---- build step 3: fill children lists (order inside a list is arbitrary: LCA answers do not depend on it) -------
__global__ void lca_fill_children_kernel(const int32_t* __restrict__ parent, int64_t n, const int32_t* __restrict__ child_off, int32_t* __restrict__ cursor,
                                         int32_t* __restrict__ child_adj) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
    const int32_t p = parent[v];
    if (p >= 0) child_adj[child_off[p] + atomicAdd(&cursor[p], 1)] = (int32_t)v;
  }
}

// children lists in ascending node id: the atomic fill above leaves them in arrival order, which would make the preorder (and
// with it the order of every who_displays list) differ from build to build; lists are short, one thread sorts one list
__global__ void lca_sort_children_kernel(int64_t n, const int32_t* __restrict__ child_off, int32_t* __restrict__ child_adj, int32_t* __restrict__ long_lists) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
    const int32_t b = child_off[v], e = child_off[v + 1];
    if (e - b == 2) { const int32_t x = child_adj[b], y = child_adj[b + 1]; if (x > y) { child_adj[b] = y; child_adj[b + 1] = x; } continue; }
    if (e - b <= 64) { for (int32_t i = b + 1; i < e; ++i) { const int32_t x = child_adj[i]; int32_t j = i; while (j > b && child_adj[j - 1] > x) { child_adj[j] = child_adj[j - 1]; --j; } child_adj[j] = x; } continue; }
    *long_lists = 1;   // more than 64 children: left to the radix sort of (parent, child) keys below (a star of 10^6 leaves in one thread took 3 s)
  }
}
__global__ void lca_child_keys_kernel(int64_t n, const int32_t* __restrict__ parent, u64* __restrict__ keys) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x)
    keys[v] = parent[v] < 0 ? ~0ull : (((u64)(uint32_t)parent[v] << 32) | (u64)(uint32_t)v);
}
__global__ void lca_child_from_keys_kernel(int64_t n, const u64* __restrict__ keys, int32_t* __restrict__ child_adj) {
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n - 1; j += (int64_t)gridDim.x * blockDim.x) child_adj[j] = (int32_t)(keys[j] & 0xffffffffull);
}

// ---- build step 4: three level-synchronous sweeps in ONE cooperative kernel -------------------------------------
//   top-down   : BFS order + depth (wavefront, children appended with one atomic per parent)
//   bottom-up  : subtree sizes
//   top-down   : preorder positions
// level_off[l] = first index of level l in order[]; ctr[3] rotating append counters; info[3] receives the level count,
// info[4] the number of nodes reached.
__global__ void __launch_bounds__(256) lca_sweeps_kernel(int64_t n, int32_t root, const int32_t* __restrict__ child_off, const int32_t* __restrict__ child_adj,
                                                         int32_t* __restrict__ order, int32_t* __restrict__ depth, int32_t* __restrict__ size,
                                                         int32_t* __restrict__ pre, int32_t* __restrict__ level_off, int32_t* __restrict__ ctr, int32_t* __restrict__ info) {
  cg::grid_group grid = cg::this_grid();
  const int64_t gtid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, gsize = (int64_t)gridDim.x * blockDim.x;
  if (gtid == 0) { order[0] = root; depth[root] = 0; pre[root] = 0; level_off[0] = 0; ctr[0] = ctr[1] = ctr[2] = 0; }
  grid.sync();
  int lev = 0; int64_t b = 0, e = 1;
  while (b < e) {
    int32_t* c = &ctr[lev % 3];
    if (gtid == 0) { ctr[(lev + 1) % 3] = 0; level_off[lev + 1] = (int32_t)e; }
    for (int64_t i = b + gtid; i < e; i += gsize) {
      const int32_t v = order[i];
      const int32_t cb = child_off[v], ce = child_off[v + 1];
      if (ce > cb) {
        const int64_t base = e + atomicAdd(c, ce - cb);
        for (int32_t j = cb; j < ce; ++j) { const int32_t ch = child_adj[j]; order[base + (j - cb)] = ch; depth[ch] = lev + 1; }
      }
    }
    grid.sync();
    const int64_t cnt = *(volatile int32_t*)c;
    b = e; e = e + cnt; ++lev;
  }
  const int levels = lev;  // number of non-empty levels
  if (gtid == 0) { info[3] = levels; info[4] = (int32_t)e; }
  if (e != n) return;      // unreachable nodes (cycle / several components): host reports the error; uniform exit
  for (int l = levels - 1; l >= 0; --l) {
    const int64_t lb = level_off[l], le = level_off[l + 1];
    for (int64_t i = lb + gtid; i < le; i += gsize) {
      const int32_t v = order[i];
      int32_t s = 1;
      for (int32_t j = child_off[v]; j < child_off[v + 1]; ++j) s += size[child_adj[j]];
      size[v] = s;
    }
    grid.sync();
  }
  for (int l = 0; l < levels; ++l) {
    const int64_t lb = level_off[l], le = level_off[l + 1];
    for (int64_t i = lb + gtid; i < le; i += gsize) {
      const int32_t v = order[i];
      int32_t run = pre[v] + 1;
      for (int32_t j = child_off[v]; j < child_off[v + 1]; ++j) { const int32_t ch = child_adj[j]; pre[ch] = run; run += size[ch]; }
    }
    grid.sync();
  }
}

// ---- build step 4': depth / preorder / subtree size by EULER TOUR + LIST RANKING (replaces the level-synchronous sweeps above
// in pt_lca_build: they pay three grid-wide barriers per LEVEL, so a path of 10^6 nodes took seconds and even a random tree of
// 2*10^7 nodes spent 8 of its 13 ms waiting at barriers and on dependent loads).  The reference's lca does the same walk
// recursively (rmq/lca.hpp:61-71: "emit node on entry and after each child").
//   arcs: 2c = the step DOWN into node c, 2c+1 = the step UP out of c (c != root); succ[] links them into the tour
//   ranking: every 64th arc (by a hash) is a SPLITTER; one thread walks from each splitter to the next one, summing length and
//   (down - up) weight of its sublist; the ~n/32 splitters are ranked by pointer jumping; a second walk hands every arc its
//   index idx and prefix weight W.   depth(c) = W(2c), preorder(c) = (idx(2c) + 1 + W(2c)) / 2, size(c) = (idx(2c+1) - idx(2c) + 1) / 2.
// Work and traffic are O(n) whatever the shape of the tree.
constexpr uint32_t ET_END = 0x7fffffffu;
__device__ __forceinline__ bool et_is_splitter(uint32_t a) { return ((a * 2654435761u) >> 26) == 0u; }
// succ word: bit 31 = "the successor is a splitter (or the end)", bits 0..30 = successor arc
__global__ void et_next_sibling_kernel(int64_t n, const int32_t* __restrict__ parent, const int32_t* __restrict__ child_off, const int32_t* __restrict__ child_adj, int32_t* __restrict__ next_sib) {
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n - 1; j += (int64_t)gridDim.x * blockDim.x) {
    const int32_t c = child_adj[j], p = parent[c];
    next_sib[c] = (j + 1 < child_off[p + 1]) ? child_adj[j + 1] : -1;
  }
}
__global__ void et_succ_kernel(int64_t n, int32_t root, const int32_t* __restrict__ parent, const int32_t* __restrict__ child_off, const int32_t* __restrict__ child_adj,
                               const int32_t* __restrict__ next_sib, uint32_t head, uint32_t* __restrict__ succ, int32_t* __restrict__ slot, int32_t* __restrict__ nsplit) {
  for (int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; a < 2 * n; a += (int64_t)gridDim.x * blockDim.x) {
    const int32_t c = (int32_t)(a >> 1);
    uint32_t nx = ET_END;
    if (c != root) {
      if (!(a & 1)) nx = child_off[c + 1] > child_off[c] ? 2u * (uint32_t)child_adj[child_off[c]] : (uint32_t)a + 1u;
      else { const int32_t sb = next_sib[c]; const int32_t p = parent[c]; nx = sb >= 0 ? 2u * (uint32_t)sb : (p == root ? ET_END : 2u * (uint32_t)p + 1u); }
    }
    const bool nsp = nx == ET_END || et_is_splitter(nx) || nx == head;
    succ[a] = nx | (nsp ? 0x80000000u : 0u);
    const bool me = c != root && (et_is_splitter((uint32_t)a) || (uint32_t)a == head);
    slot[a] = me ? atomicAdd(nsplit, 1) : -1;
  }
}
// first walk: sublist [a, next splitter): length and weight; next splitter's slot
__global__ void et_walk1_kernel(int64_t n, const uint32_t* __restrict__ succ, const int32_t* __restrict__ slot, int32_t* __restrict__ sp_arc, int32_t* __restrict__ nxt,
                                int32_t* __restrict__ len, int32_t* __restrict__ wsum) {
  for (int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; a < 2 * n; a += (int64_t)gridDim.x * blockDim.x) {
    const int32_t i = slot[a];
    if (i < 0) continue;
    uint32_t x = (uint32_t)a; int32_t l = 0, w = 0;
    for (int64_t guard = 0; guard <= 2 * n; ++guard) {
      ++l; w += (x & 1u) ? -1 : 1;
      const uint32_t sw = succ[x];
      x = sw & 0x7fffffffu;
      if (sw & 0x80000000u) break;
    }
    sp_arc[i] = (int32_t)a; len[i] = l; wsum[i] = w;
    nxt[i] = x == ET_END ? -1 : slot[x];
  }
}
// pointer jumping over the splitters: suffix sums of (len, wsum)
__global__ void et_jump_kernel(int32_t S, const int32_t* __restrict__ nxt_in, const int32_t* __restrict__ len_in, const int32_t* __restrict__ w_in,
                               int32_t* __restrict__ nxt_out, int32_t* __restrict__ len_out, int32_t* __restrict__ w_out) {
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < S; i += gridDim.x * blockDim.x) {
    const int32_t j = nxt_in[i];
    if (j < 0) { nxt_out[i] = -1; len_out[i] = len_in[i]; w_out[i] = w_in[i]; }
    else { nxt_out[i] = nxt_in[j]; len_out[i] = len_in[i] + len_in[j]; w_out[i] = w_in[i] + w_in[j]; }
  }
}
// second walk: idx and prefix weight of every arc of the sublist; idx_down -> pre[], idx_up -> size[], W(down) -> depth[]
__global__ void et_walk2_kernel(int32_t S, int64_t n, const uint32_t* __restrict__ succ, const int32_t* __restrict__ sp_arc, const int32_t* __restrict__ nxt_final,
                                const int32_t* __restrict__ suf_len, const int32_t* __restrict__ suf_w, int32_t head_slot, int32_t* __restrict__ depth, int32_t* __restrict__ pre, int32_t* __restrict__ size) {
  const int32_t total = suf_len[head_slot], totw = suf_w[head_slot];
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < S; i += gridDim.x * blockDim.x) {
    if (nxt_final[i] != -1) continue;   // a splitter whose list never ends: it sits on a cycle beside the tree (the build reports it)
    uint32_t x = (uint32_t)sp_arc[i];
    int32_t idx = total - suf_len[i], w = totw - suf_w[i];
    if (idx < 0) continue;
    for (int64_t guard = 0; guard <= 2 * n; ++guard) {
      const int32_t c = (int32_t)(x >> 1);
      if (x & 1u) { w -= 1; size[c] = idx; } else { w += 1; pre[c] = idx; depth[c] = w; }
      ++idx;
      const uint32_t sw = succ[x];
      x = sw & 0x7fffffffu;
      if (sw & 0x80000000u) break;
    }
  }
}
__global__ void et_finish_kernel(int64_t n, int32_t root, int32_t* __restrict__ depth, int32_t* __restrict__ pre, int32_t* __restrict__ size, int32_t* __restrict__ info /* [3] levels [4] reached */) {
  int32_t mx = 0, reached = 0;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < n; c += (int64_t)gridDim.x * blockDim.x) {
    if (c == root) { depth[c] = 0; pre[c] = 0; size[c] = (int32_t)n; ++reached; continue; }
    const int32_t id = pre[c], iu = size[c], d = depth[c];
    if (id < 0 || iu < id) { continue; }   // never visited
    pre[c] = (id + 1 + d) / 2; size[c] = (iu - id + 1) / 2;
    mx = max(mx, d); ++reached;
  }
  for (int o = 16; o; o >>= 1) { mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o)); reached += __shfl_xor_sync(0xffffffffu, reached, o); }
  if ((threadIdx.x & 31) == 0) { atomicMax(&info[3], mx + 1); atomicAdd(&info[4], reached); }
}

// ---- build step 5: keys by position ------------------------------------------------------------------------------
__global__ void lca_keys_kernel(int64_t n, const int32_t* __restrict__ parent, const int32_t* __restrict__ depth, const int32_t* __restrict__ pre,
                                u64* __restrict__ keypos, int32_t* __restrict__ nodeat) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
    const int32_t p = pre[v];
    keypos[p] = ((u64)(uint32_t)depth[v] << 32) | (u64)(uint32_t)parent[v];
    nodeat[p] = (int32_t)v;
  }
}

// ---- build step 6: one warp per block of 32 positions: prefix / suffix minima, stack masks, block minimum ----------
template <bool BY_NODE>
__global__ void __launch_bounds__(256) lca_blocks_kernel(int64_t n, int64_t nblocks, const u64* __restrict__ keypos, const int32_t* __restrict__ nodeat,
                                                         LcaRecord* __restrict__ rec, u64* __restrict__ table0) {
  const int lane = threadIdx.x & 31;
  for (int64_t blk = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; blk < nblocks; blk += ((int64_t)gridDim.x * blockDim.x) >> 5) {
    const int64_t pos = blk * 32 + lane;
    const bool valid = pos < n;
    const u64 key = valid ? keypos[pos] : KEY_INF;
    u64 pre = key;  // inclusive prefix minimum
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const u64 t = __shfl_up_sync(0xffffffffu, pre, d); if (lane >= d && t < pre) pre = t; }
    u64 sufi = key;  // inclusive suffix minimum
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const u64 t = __shfl_down_sync(0xffffffffu, sufi, d); if (lane + d < 32 && t < sufi) sufi = t; }
    u64 suf = __shfl_down_sync(0xffffffffu, sufi, 1);  // LCA: exclusive (pos, block end];  RMQ: inclusive [pos, block end]
    if (lane == 31) suf = KEY_INF;
    if (!BY_NODE) suf = sufi;
    // mask of lane r: bit j set iff key[j] <= min(key[j+1..r]); walk j downwards with a running minimum
    uint32_t mask = 0; u64 run = KEY_INF;
    for (int j = 31; j >= 0; --j) {
      const u64 kj = __shfl_sync(0xffffffffu, key, j);
      if (j <= lane && kj <= run) { mask |= 1u << j; run = kj; }
    }
    if (valid) {
      LcaRecord r; r.pos = (uint32_t)pos; r.mask = mask; r.suf = suf; r.pre = pre; r.key = key;
      uint4* dst = reinterpret_cast<uint4*>(&rec[BY_NODE ? (int64_t)nodeat[pos] : pos]);
      const uint4* srcv = reinterpret_cast<const uint4*>(&r);
      dst[0] = srcv[0]; dst[1] = srcv[1];
    }
    const u64 bmin = __shfl_sync(0xffffffffu, pre, 31);
    if (lane == 0) table0[blk] = bmin;
  }
}

// ---- build step 7: one sparse-table level (sparse_rmq.hpp:52-65, on packed keys instead of indices) ----------------
__global__ void lca_table_level_kernel(const u64* __restrict__ prev, u64* __restrict__ next, int64_t count, int64_t half) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const u64 a = prev[i], b = prev[i + half];
    next[i] = a < b ? a : b;
  }
}

// ---- query ------------------------------------------------------------------------------------------------------------
template <int QPT>
__global__ void __launch_bounds__(256) lca_query_kernel(int64_t q, int64_t n, const int32_t* __restrict__ qu, const int32_t* __restrict__ qv, int32_t* __restrict__ out,
                                                        const LcaRecord* __restrict__ rec, const u64* __restrict__ keypos, const u64* __restrict__ table,
                                                        const int64_t* __restrict__ level_base /* device array: offset of level k in table */) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const u64 keep = l2_policy_evict_last();
  for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < q; i0 += stride * QPT) {
    int32_t u[QPT], v[QPT];
    // record words: w0 = pos | mask << 32, w1 = suf, w2 = pre, w3 = key
    u64 a0[QPT], a1[QPT], a2[QPT], a3[QPT], b0[QPT], b1[QPT], b2[QPT], b3[QPT];
    bool ok[QPT];
#pragma unroll
    for (int j = 0; j < QPT; ++j) {
      const int64_t i = i0 + j * stride;
      ok[j] = i < q;
      u[j] = ok[j] ? ld_stream(qu + i) : 0; v[j] = ok[j] ? ld_stream(qv + i) : 0;
    }
#pragma unroll
    for (int j = 0; j < QPT; ++j) {
      ok[j] = ok[j] && (uint64_t)(uint32_t)u[j] < (uint64_t)n && (uint64_t)(uint32_t)v[j] < (uint64_t)n;
      if (ok[j]) { ld_sector_evict_first(rec + u[j], a0[j], a1[j], a2[j], a3[j]); ld_sector_evict_first(rec + v[j], b0[j], b1[j], b2[j], b3[j]); }
    }
#pragma unroll
    for (int j = 0; j < QPT; ++j) {
      const int64_t i = i0 + j * stride;
      if (i >= q) continue;
      int32_t ans = -1;
      if (ok[j]) {
        if (u[j] == v[j]) ans = u[j];
        else {
          // left = smaller position
          const bool swap = (uint32_t)a0[j] > (uint32_t)b0[j];
          const u64 l0 = swap ? b0[j] : a0[j], lsuf = swap ? b1[j] : a1[j];
          const u64 r0 = swap ? a0[j] : b0[j], rpre = swap ? a2[j] : b2[j];
          const uint32_t lp = (uint32_t)l0, rp = (uint32_t)r0;
          const uint32_t bl = lp >> 5, br = rp >> 5;
          u64 best;
          if (bl == br) {
            const uint32_t m = (uint32_t)(r0 >> 32) >> ((lp & 31) + 1);        // stack entries strictly right of lp
            const uint32_t jpos = (lp & 31) + 1 + (uint32_t)__ffs(m) - 1;        // m != 0: bit (rp & 31) is always set
            best = jpos == (rp & 31) ? (swap ? a3[j] : b3[j]) : keypos[((int64_t)bl << 5) + jpos];  // the right end itself: its own key is in the record
          } else {
            best = lsuf < rpre ? lsuf : rpre;      // (lp, end of its block]  and  [start of rp's block, rp]
            if (br - bl >= 2) {
              const uint32_t span = br - bl - 1;   // whole blocks bl+1 .. br-1
              const int k = 31 - __clz(span);
              const u64* lvl = table + level_base[k];
              const u64 t1 = ld_u64_hint(lvl + (bl + 1), keep), t2 = ld_u64_hint(lvl + (br - (1u << k)), keep);
              const u64 t = t1 < t2 ? t1 : t2;
              best = t < best ? t : best;
            }
          }
          ans = (int32_t)(uint32_t)(best & 0xffffffffull);
        }
      }
      st_stream(out + i, ans);
    }
  }
}

// ---- consecutive pairs of a list: out[i] = lca(nodes[i], nodes[i+1]) -- the induced-subtree step of the reference
// (utils/induced_tree.hpp:128-138: LCAs of consecutive candidates of a preorder-sorted list).  Every record is loaded ONCE: a
// lane gets the record of its right neighbour by shuffle (the last lane of a warp loads the one record beyond the warp).  For a
// list sorted by preorder on a tree whose node ids follow a preorder (what the reference's readers produce) the record loads
// stream: 4 B in + 32 B record + 4 B out per query = SURVEY's 40-byte model, moved once.
template <int QPT>
__global__ void __launch_bounds__(256) lca_adjacent_kernel(int64_t q /* = count - 1 */, int64_t n, const int32_t* __restrict__ nodes, int32_t* __restrict__ out,
                                                           const LcaRecord* __restrict__ rec, const u64* __restrict__ keypos, const u64* __restrict__ table,
                                                           const int64_t* __restrict__ level_base) {
  const int lane = threadIdx.x & 31;
  const u64 keep = l2_policy_evict_last();
  const int64_t tile = (int64_t)blockDim.x * QPT;
  for (int64_t t0 = blockIdx.x * tile; t0 < q; t0 += (int64_t)gridDim.x * tile) {
    int32_t u[QPT]; u64 a0[QPT], a1[QPT], a2[QPT], a3[QPT];
#pragma unroll
    for (int j = 0; j < QPT; ++j) {
      const int64_t i = t0 + (int64_t)j * blockDim.x + threadIdx.x;   // element i of the list (i <= q is a valid element)
      u[j] = i <= q ? ld_stream(nodes + i) : -1;
    }
#pragma unroll
    for (int j = 0; j < QPT; ++j) {
      a0[j] = a1[j] = a2[j] = a3[j] = 0;
      if ((uint64_t)(uint32_t)u[j] < (uint64_t)n) { const uint4* p = reinterpret_cast<const uint4*>(rec + u[j]); const uint4 x = __ldg(p), y = __ldg(p + 1);
        a0[j] = ((u64)x.y << 32) | x.x; a1[j] = ((u64)x.w << 32) | x.z; a2[j] = ((u64)y.y << 32) | y.x; a3[j] = ((u64)y.w << 32) | y.z; }
    }
#pragma unroll
    for (int j = 0; j < QPT; ++j) {
      const int64_t i = t0 + (int64_t)j * blockDim.x + threadIdx.x;
      // right neighbour: element i + 1
      int32_t v = __shfl_down_sync(0xffffffffu, u[j], 1);
      u64 b0 = __shfl_down_sync(0xffffffffu, a0[j], 1), b2 = __shfl_down_sync(0xffffffffu, a2[j], 1), b3 = __shfl_down_sync(0xffffffffu, a3[j], 1), b1 = __shfl_down_sync(0xffffffffu, a1[j], 1);
      if (lane == 31 && i < q) {
        v = ld_stream(nodes + i + 1);
        if ((uint64_t)(uint32_t)v < (uint64_t)n) { const LcaRecord r = rec[v]; b0 = ((u64)r.mask << 32) | r.pos; b1 = r.suf; b2 = r.pre; b3 = r.key; }
      }
      if (i >= q) continue;
      int32_t ans = -1;
      if ((uint64_t)(uint32_t)u[j] < (uint64_t)n && (uint64_t)(uint32_t)v < (uint64_t)n) {
        if (u[j] == v) ans = v;
        else {
          const bool swap = (uint32_t)a0[j] > (uint32_t)b0;
          const u64 l0 = swap ? b0 : a0[j], lsuf = swap ? b1 : a1[j];
          const u64 r0 = swap ? a0[j] : b0, rpre = swap ? a2[j] : b2, rkey = swap ? a3[j] : b3;
          const uint32_t lp = (uint32_t)l0, rp = (uint32_t)r0, bl = lp >> 5, br = rp >> 5;
          u64 best;
          if (bl == br) {
            const uint32_t m = (uint32_t)(r0 >> 32) >> ((lp & 31) + 1);
            const uint32_t jpos = (lp & 31) + 1 + (uint32_t)__ffs(m) - 1;
            best = jpos == (rp & 31) ? rkey : keypos[((int64_t)bl << 5) + jpos];
          } else {
            best = lsuf < rpre ? lsuf : rpre;
            if (br - bl >= 2) {
              const uint32_t span = br - bl - 1;
              const int k = 31 - __clz(span);
              const u64* lvl = table + level_base[k];
              const u64 t1 = ld_u64_hint(lvl + (bl + 1), keep), t2 = ld_u64_hint(lvl + (br - (1u << k)), keep);
              const u64 t = t1 < t2 ? t1 : t2;
              best = t < best ? t : best;
            }
          }
          ans = (int32_t)(uint32_t)(best & 0xffffffffull);
        }
      }
      st_stream(out + i, ans);
    }
  }
}

__global__ void lca_node_info_kernel(int64_t n, const LcaRecord* __restrict__ rec, int32_t* __restrict__ depth, int32_t* __restrict__ preorder) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
    if (depth) depth[v] = (int32_t)(rec[v].key >> 32);
    if (preorder) preorder[v] = (int32_t)rec[v].pos;
  }
}

// ---- generic RMQ over int32 values (rmq/sparse_rmq.hpp semantics: index of the RIGHTMOST minimum of [l, r)) ------------
// key = (value biased to unsigned) << 32 | (0xFFFFFFFF - index): the smallest key is the smallest value, largest index.
__global__ void rmq_keys_kernel(int64_t n, const int32_t* __restrict__ values, u64* __restrict__ keypos) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    keypos[i] = ((u64)((uint32_t)values[i] ^ 0x80000000u) << 32) | (u64)(0xFFFFFFFFu - (uint32_t)i);
}
__global__ void __launch_bounds__(256) rmq_query_kernel(int64_t q, int64_t n, const int64_t* __restrict__ ql, const int64_t* __restrict__ qr, int64_t* __restrict__ out,
                                                        const LcaRecord* __restrict__ rec, const u64* __restrict__ keypos, const u64* __restrict__ table,
                                                        const int64_t* __restrict__ level_base) {
  const u64 keep = l2_policy_evict_last();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < q; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t l = ql[i], r = qr[i] - 1;  // inclusive [l, r]
    int64_t ans = -1;
    if (l >= 0 && l <= r && r < n) {
      u64 a0, a1, a2, a3, b0, b1, b2, b3;
      ld_sector_evict_first(rec + l, a0, a1, a2, a3);
      ld_sector_evict_first(rec + r, b0, b1, b2, b3);
      const uint32_t bl = (uint32_t)(l >> 5), br = (uint32_t)(r >> 5);
      u64 best;
      if (bl == br) {
        const uint32_t m = (uint32_t)(b0 >> 32) >> (l & 31);          // stack entries at or right of l
        best = keypos[((int64_t)bl << 5) + (l & 31) + __ffs(m) - 1];    // m != 0: bit (r & 31) is always set
      } else {
        best = a1 < b2 ? a1 : b2;                                       // [l, end of block]  and  [start of block, r]
        if (br - bl >= 2) {
          const uint32_t span = br - bl - 1;
          const int k = 31 - __clz(span);
          const u64* lvl = table + level_base[k];
          const u64 t1 = ld_u64_hint(lvl + (bl + 1), keep), t2 = ld_u64_hint(lvl + (br - (1u << k)), keep);
          const u64 t = t1 < t2 ? t1 : t2;
          best = t < best ? t : best;
        }
      }
      ans = (int64_t)(0xFFFFFFFFu - (uint32_t)(best & 0xffffffffull));
    }
    out[i] = ans;
  }
}

// ---- one LCA query from device code (same arithmetic as lca_query_kernel) ------------------------------------------------
__device__ __forceinline__ int32_t lca_one(const LcaRecord* __restrict__ rec, const u64* __restrict__ keypos, const u64* __restrict__ table,
                                           const int64_t* __restrict__ level_base, int32_t u, int32_t v) {
  if (u == v) return u;
  LcaRecord a = rec[u], b = rec[v];
  if (a.pos > b.pos) { const LcaRecord t = a; a = b; b = t; }
  const uint32_t lp = a.pos, rp = b.pos, bl = lp >> 5, br = rp >> 5;
  u64 best;
  if (bl == br) {
    const uint32_t m = b.mask >> ((lp & 31) + 1);
    best = keypos[((int64_t)bl << 5) + (lp & 31) + 1 + __ffs(m) - 1];
  } else {
    best = a.suf < b.pre ? a.suf : b.pre;
    if (br - bl >= 2) {
      const uint32_t span = br - bl - 1;
      const int k = 31 - __clz(span);
      const u64* lvl = table + level_base[k];
      const u64 t1 = lvl[bl + 1], t2 = lvl[br - (1u << k)];
      const u64 t = t1 < t2 ? t1 : t2;
      best = t < best ? t : best;
    }
  }
  return (int32_t)(uint32_t)(best & 0xffffffffull);
}

// ---- who_displays for a single-labelled tree host (TreeInTreeContainment::who_displays, utils/containment.hpp:72-177) ------
// The reference computes, per guest node u, the host nodes that minimally display the guest subtree below u: children
// possibilities are merged in host preorder, the induced subtree is built from the LCAs of consecutive candidates
// (utils/induced_tree.hpp:128-138) and a bipartite matching assigns children to distinct branches.  With a single-labelled
// host every list has at most one entry, the induced subtree of k candidates needs exactly the k-1 consecutive LCAs, and
// the matching degenerates to "all consecutive LCAs are the same node v and no candidate is v itself".  Each guest node = (small
// sort of its children's images by host preorder) + k-1 LCA queries.
constexpr int WD_MAX_CHILDREN = 64;   // up to here the images are sorted in registers / local memory; beyond, in a global scratch span

// in-place heap sort of 32- or 64-bit keys (one thread)
template <typename K>
__device__ inline void dp_heapsort(K* a, int64_t n) {
  for (int64_t start = n / 2 - 1; start >= 0; --start) {
    int64_t r = start; const K v = a[r];
    for (;;) { int64_t c = 2 * r + 1; if (c >= n) break; if (c + 1 < n && a[c + 1] > a[c]) ++c; if (a[c] <= v) break; a[r] = a[c]; r = c; }
    a[r] = v;
  }
  for (int64_t end = n - 1; end > 0; --end) {
    const K v = a[end]; a[end] = a[0];
    int64_t r = 0;
    for (;;) { int64_t c = 2 * r + 1; if (c >= end) break; if (c + 1 < end && a[c + 1] > a[c]) ++c; if (a[c] <= v) break; a[r] = a[c]; r = c; }
    a[r] = v;
  }
}

// a guest node with more than WD_MAX_CHILDREN children (a star, a wide multifurcation): the same test over (position, node) keys
// sorted in a span of `big` -- every guest node is the child of one parent, so all spans together need at most nt keys
__device__ __noinline__ int32_t wd_node_big(const int32_t cb, const int32_t ce, const int32_t* __restrict__ child_adj, const LcaRecord* __restrict__ rec, const u64* __restrict__ keypos,
                                            const u64* __restrict__ table, const int64_t* __restrict__ level_base, const int32_t* disp, u64* __restrict__ big, unsigned long long* __restrict__ big_top) {
  int64_t k = 0;
  for (int32_t j = cb; j < ce; ++j) {
    const int32_t c = __ldcg(&disp[child_adj[j]]);
    if (c == -2) continue;
    if (c == -1) return -1;
    ++k;
  }
  if (k == 0) return -2;
  u64* keys = big + atomicAdd(big_top, (unsigned long long)k);
  int64_t i = 0;
  for (int32_t j = cb; j < ce; ++j) {
    const int32_t c = __ldcg(&disp[child_adj[j]]);
    if (c >= 0) keys[i++] = ((u64)rec[c].pos << 32) | (u64)(uint32_t)c;
  }
  if (k == 1) return (int32_t)(uint32_t)keys[0];
  dp_heapsort(keys, k);
  const int32_t v = lca_one(rec, keypos, table, level_base, (int32_t)(uint32_t)keys[0], (int32_t)(uint32_t)keys[k - 1]);
  for (int64_t a = 0; a < k; ++a) if ((int32_t)(uint32_t)keys[a] == v) return -1;
  for (int64_t a = 0; a + 1 < k; ++a)   // (two children with the same image have that image as their LCA, which is not v)
    if (lca_one(rec, keypos, table, level_base, (int32_t)(uint32_t)keys[a], (int32_t)(uint32_t)keys[a + 1]) != v) return -1;
  return v;
}
__device__ __forceinline__ int32_t wd_node(const int32_t u, const int32_t* __restrict__ child_off, const int32_t* __restrict__ child_adj, const int32_t* __restrict__ tlabel,
                                           const int32_t* __restrict__ l2h, int64_t nlab, const LcaRecord* __restrict__ rec, const u64* __restrict__ keypos,
                                           const u64* __restrict__ table, const int64_t* __restrict__ level_base, const int32_t* disp, u64* __restrict__ big, unsigned long long* __restrict__ big_top) {
  const int32_t cb = child_off[u], ce = child_off[u + 1];
  if (cb == ce) {  // guest leaf: base case construct_base_cases :115-126
    const int32_t lab = tlabel[u];
    return lab < 0 ? -2 : ((int64_t)lab < nlab ? l2h[lab] : -1);
  }
  if (ce - cb > WD_MAX_CHILDREN) return wd_node_big(cb, ce, child_adj, rec, keypos, table, level_base, disp, big, big_top);
  int32_t img[WD_MAX_CHILDREN]; uint32_t ps[WD_MAX_CHILDREN];
  int k = 0;
  for (int32_t j = cb; j < ce; ++j) {
    const int32_t c = __ldcg(&disp[child_adj[j]]);   // (written by another SM during the same launch: through L2)
    if (c == -2) continue;               // unlabelled dangling subtree: cleaned away
    if (c == -1) return -1;              // merge_child_poss :180-230: an empty child list empties the parent
    // insertion sort by host preorder position (sort_by_order :108-111)
    const uint32_t p = rec[c].pos;
    int a = k++;
    while (a > 0 && ps[a - 1] > p) { ps[a] = ps[a - 1]; img[a] = img[a - 1]; --a; }
    ps[a] = p; img[a] = c;
  }
  if (k == 0) return -2;
  if (k == 1) return img[0];             // :167: a node with one child maps where the child maps
  const int32_t v = lca_one(rec, keypos, table, level_base, img[0], img[k - 1]);
  bool ok = true;
  for (int a = 0; a < k && ok; ++a) ok = img[a] != v;
  for (int a = 0; a + 1 < k && ok; ++a) ok = lca_one(rec, keypos, table, level_base, img[a], img[a + 1]) == v;
  return ok ? v : -1;
}
// DATAFLOW over the guest tree (no levels, no grid barrier: see dp_who_flow_kernel): a thread starts at every guest leaf; whoever
// completes the last child of a node goes on with that node
__global__ void __launch_bounds__(256) who_displays_kernel(int64_t nt, const int32_t* __restrict__ gparent, int32_t* __restrict__ pending, const int32_t* __restrict__ child_off,
                                                           const int32_t* __restrict__ child_adj, const int32_t* __restrict__ tlabel, const int32_t* __restrict__ l2h, int64_t nlab,
                                                           const LcaRecord* __restrict__ rec, const u64* __restrict__ keypos, const u64* __restrict__ table,
                                                           const int64_t* __restrict__ level_base, int32_t* disp /* -2 dead, -1 not displayed, >= 0 host node */,
                                                           u64* __restrict__ big /* nt keys: sort space of the nodes with many children */, unsigned long long* __restrict__ processed /* [1]: top of big */) {
  unsigned long long done = 0;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < nt; idx += (int64_t)gridDim.x * blockDim.x) {
    int32_t u = (int32_t)idx;
    if (child_off[u] != child_off[u + 1]) continue;
    for (;;) {
      disp[u] = wd_node(u, child_off, child_adj, tlabel, l2h, nlab, rec, keypos, table, level_base, disp, big, processed + 1);
      ++done;
      const int32_t p = gparent[u];
      if (p < 0 || p >= nt) break;
      __threadfence();
      if (atomicSub(&pending[p], 1) != 1) break;
      __threadfence();
      u = p;
    }
  }
  if (done) atomicAdd(processed, done);
}

__global__ void wd_label_map_kernel(int64_t n, const int32_t* __restrict__ deg, const int32_t* __restrict__ hlabel, int64_t nlab, int32_t* __restrict__ l2h, int32_t* __restrict__ flags) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
    const int32_t lab = hlabel[v];
    if (lab >= 0 && deg[v] == 0) { if (lab >= nlab) flags[2] = 1; else if (atomicCAS(&l2h[lab], -1, (int32_t)v) != -1) flags[1] = 1; }
  }
}
__global__ void wd_finish_kernel(int64_t nt, int32_t* __restrict__ disp) {
  for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < nt; u += (int64_t)gridDim.x * blockDim.x) if (disp[u] < 0) disp[u] = -1;
}

// device exclusive scan of int32 -> int32 (out has n+1 entries); tmp holds ceil(n/SCAN_TILE) int64
static int device_exclusive_scan(const int32_t* in, int64_t n, int32_t* out, int64_t* tmp, cudaStream_t s) {
  const int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  if (tiles == 0) { PT_CUDA(cudaMemsetAsync(out, 0, 4, s)); return PT_OK; }
  scan_tile_sums_kernel<<<(unsigned)tiles, 256, 0, s>>>(in, n, tmp);
  scan_tile_offsets_kernel<<<1, 1024, 0, s>>>(tmp, tiles);
  scan_apply_kernel<<<(unsigned)tiles, 256, 0, s>>>(in, n, tmp, out);
  PT_CUDA(cudaGetLastError());
  return PT_OK;
}

}  // namespace ptgpu

using namespace ptgpu;

struct pt_lca {
  int64_t n = 0, nblocks = 0, levels = 0, table_levels = 0, device_bytes = 0;
  float build_ms = 0.f;
  LcaRecord* rec = nullptr; u64* keypos = nullptr; u64* table = nullptr; int64_t* d_level_base = nullptr;
  int32_t* size = nullptr;    // subtree sizes (kept only when requested: bridges need preorder intervals)
  int32_t* nodeat = nullptr;  // node at every preorder position (kept with `size`: the tree-in-tree DP works on positions)
  int32_t *d_u = nullptr, *d_v = nullptr, *d_out = nullptr; int64_t qcap = 0;  // staging for HOST queries
};

static void lca_release(pt_lca* h) {
  if (!h) return;
  cudaDeviceSynchronize();
  dev_free(h->rec); dev_free(h->keypos); dev_free(h->table); dev_free(h->d_level_base); dev_free(h->d_u); dev_free(h->d_v); dev_free(h->d_out); dev_free(h->size); dev_free(h->nodeat);
  delete h;
}

extern "C" {

static int lca_build_impl(pt_lca** out, const int32_t* parent, int64_t n, pt_mem parent_mem, void* stream, bool keep_size) {
  if (!out || !parent || n <= 0 || n > (int64_t)INT32_MAX - 64) return set_error(PT_ERR_INVALID, "pt_lca_build: bad argument (n must be in [1, 2^31-65])");
  *out = nullptr;
  if (int rc = require_device()) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  DeviceProps dp;
  if (int rc = device_props(&dp)) return rc;
  pt_lca* h = new (std::nothrow) pt_lca();
  if (!h) return set_error(PT_ERR_NOMEM, "out of host memory");
  h->n = n; h->nblocks = (n + 31) / 32;
  // scratch (freed before returning)
  int32_t *d_parent = nullptr, *deg = nullptr, *child_off = nullptr, *child_adj = nullptr, *order = nullptr, *depth = nullptr, *size = nullptr, *pre = nullptr,
          *level_off = nullptr, *small = nullptr, *nodeat = nullptr;
  int64_t* scan_tmp = nullptr;
  void* et_owned[3] = {nullptr, nullptr, nullptr};   // scratch of the Euler-tour ranking
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int rc = PT_OK;
  auto cleanup = [&]() {
    cudaStreamSynchronize(s);
    for (void* p : et_owned) dev_free(p);
    if (parent_mem == PT_MEM_HOST) dev_free(d_parent);
    dev_free(deg); dev_free(child_off); dev_free(child_adj); dev_free(order); dev_free(depth); if (size != h->size) dev_free(size); dev_free(pre); dev_free(level_off); dev_free(small);
    if (nodeat != h->nodeat) dev_free(nodeat);
    dev_free(scan_tmp);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
  };
#define LCA_TRY(call)                                                                                                     \
  do {                                                                                                                    \
    cudaError_t e__ = (call);                                                                                             \
    if (e__ != cudaSuccess) {                                                                                             \
      rc = set_error(e__ == cudaErrorMemoryAllocation ? PT_ERR_NOMEM : PT_ERR_CUDA, "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      cleanup(); lca_release(h); return rc;                                                                               \
    }                                                                                                                     \
  } while (0)
  static const bool trace = getenv("PT_SPLIT_TRACE") != nullptr;   // profiling aid: wall time of the build's phases on stderr
  auto t_last = std::chrono::steady_clock::now();
  auto mark = [&](const char* what) {
    if (!trace) return;
    cudaStreamSynchronize(s);
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[lca build n=%lld] %-22s %8.3f ms\n", (long long)n, what, std::chrono::duration<double, std::milli>(now - t_last).count());
    t_last = now;
  };
  const size_t n4 = (size_t)n * 4;
  if (parent_mem == PT_MEM_HOST) { LCA_TRY(lca_malloc((void**)&d_parent, n4)); LCA_TRY(cudaMemcpyAsync(d_parent, parent, n4, cudaMemcpyHostToDevice, s)); }
  else d_parent = (int32_t*)parent;
  LCA_TRY(cudaEventCreate(&ev0)); LCA_TRY(cudaEventCreate(&ev1));
  LCA_TRY(lca_malloc((void**)&deg, n4 + 4)); LCA_TRY(lca_malloc((void**)&child_off, n4 + 4)); LCA_TRY(lca_malloc((void**)&child_adj, n4));
  LCA_TRY(lca_malloc((void**)&order, n4)); LCA_TRY(lca_malloc((void**)&depth, n4)); LCA_TRY(lca_malloc((void**)&size, n4)); LCA_TRY(lca_malloc((void**)&pre, n4));
  LCA_TRY(lca_malloc((void**)&level_off, n4 + 8)); LCA_TRY(lca_malloc((void**)&small, 64)); LCA_TRY(lca_malloc((void**)&nodeat, n4));
  LCA_TRY(lca_malloc((void**)&scan_tmp, ((size_t)((n + SCAN_TILE - 1) / SCAN_TILE) + 1) * 8));
  LCA_TRY(lca_malloc((void**)&h->rec, (size_t)n * sizeof(LcaRecord)));
  LCA_TRY(lca_malloc((void**)&h->keypos, (size_t)n * 8));
  // table geometry: level k has nblocks - 2^k + 1 entries
  std::vector<int64_t> level_base; int64_t total = 0;
  for (int k = 0; ((int64_t)1 << k) <= h->nblocks; ++k) { level_base.push_back(total); total += h->nblocks - ((int64_t)1 << k) + 1; }
  h->table_levels = (int64_t)level_base.size();
  LCA_TRY(lca_malloc((void**)&h->table, (size_t)total * 8));
  LCA_TRY(lca_malloc((void**)&h->d_level_base, level_base.size() * 8));
  LCA_TRY(cudaMemcpyAsync(h->d_level_base, level_base.data(), level_base.size() * 8, cudaMemcpyHostToDevice, s));
  h->device_bytes = (int64_t)n * (int64_t)sizeof(LcaRecord) + n * 8 + total * 8;

  mark("allocations");
  LCA_TRY(cudaEventRecord(ev0, s));
  const int grid = dp.sms * 8;
  LCA_TRY(cudaMemsetAsync(deg, 0, n4 + 4, s)); LCA_TRY(cudaMemsetAsync(small, 0, 64, s));
  lca_count_kernel<<<grid, 256, 0, s>>>(d_parent, n, deg, small);
  if ((rc = device_exclusive_scan(deg, n, child_off, scan_tmp, s)) != PT_OK) { cleanup(); lca_release(h); return rc; }
  LCA_TRY(cudaMemsetAsync(deg, 0, n4, s));
  lca_fill_children_kernel<<<grid, 256, 0, s>>>(d_parent, n, child_off, deg, child_adj);
  lca_sort_children_kernel<<<grid, 256, 0, s>>>(n, child_off, child_adj, small + 5);
  int32_t info[8];
  LCA_TRY(cudaMemcpyAsync(info, small, 32, cudaMemcpyDeviceToHost, s));
  LCA_TRY(cudaStreamSynchronize(s));
  if (info[5]) {   // some node has more than 64 children: the whole children array by one radix sort of (parent << 32 | child)
    u64 *k1 = nullptr, *k2 = nullptr; void* tmp = nullptr; size_t tb = 0;
    LCA_TRY(lca_malloc((void**)&k1, (size_t)n * 8)); et_owned[0] = k1;
    LCA_TRY(lca_malloc((void**)&k2, (size_t)n * 8)); et_owned[1] = k2;
    lca_child_keys_kernel<<<grid, 256, 0, s>>>(n, d_parent, k1);
    LCA_TRY(cub::DeviceRadixSort::SortKeys(nullptr, tb, k1, k2, (int)n, 0, 64, s));
    LCA_TRY(lca_malloc(&tmp, tb)); et_owned[2] = tmp;
    LCA_TRY(cub::DeviceRadixSort::SortKeys(tmp, tb, k1, k2, (int)n, 0, 64, s));
    lca_child_from_keys_kernel<<<grid, 256, 0, s>>>(n, k2, child_adj);
    LCA_TRY(cudaStreamSynchronize(s));
    for (void*& p : et_owned) { dev_free(p); p = nullptr; }
  }
  mark("children CSR");
  if (info[0] != 1 || info[2]) {  // utils/storage_adj_common.hpp:103-109: exactly one root
    rc = set_error(PT_ERR_INVALID, "pt_lca_build: parent array is not a rooted tree (%d roots%s)", info[0], info[2] ? ", parent index out of range" : "");
    cleanup(); lca_release(h); return rc;
  }
  static const bool use_sweeps = getenv("PT_LCA_SWEEPS") != nullptr;   // A/B aid: the level-synchronous sweeps this build used before
  if (use_sweeps) {
    int occ = 0;
    LCA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lca_sweeps_kernel, 256, 0));
    const int cgrid = dp.sms * std::max(1, std::min(occ, 4));
    int32_t root = info[1];
    int64_t nn = n;
    int32_t* ctr = small + 8; int32_t* inf = small;
    void* args[] = {&nn, &root, &child_off, &child_adj, &order, &depth, &size, &pre, &level_off, &ctr, &inf};
    LCA_TRY(cudaLaunchCooperativeKernel((void*)lca_sweeps_kernel, dim3(cgrid), dim3(256), args, 0, s));
  } else if (n == 1) {
    const int32_t one[2] = {1, 1};
    LCA_TRY(cudaMemsetAsync(depth, 0, 4, s)); LCA_TRY(cudaMemsetAsync(pre, 0, 4, s)); LCA_TRY(cudaMemcpyAsync(size, one, 4, cudaMemcpyHostToDevice, s));
    LCA_TRY(cudaMemcpyAsync(small + 3, one, 8, cudaMemcpyHostToDevice, s));
  } else {
    // Euler tour + list ranking.  Scratch: order[] and level_off[] (free in this path) hold next_sib and the splitter slots' arcs.
    const int32_t root = info[1];
    uint32_t* succ = nullptr; int32_t* slot = nullptr;
    LCA_TRY(lca_malloc((void**)&succ, (size_t)n * 8)); et_owned[0] = succ;
    LCA_TRY(lca_malloc((void**)&slot, (size_t)n * 8)); et_owned[1] = slot;
    int32_t* next_sib = order;
    et_next_sibling_kernel<<<grid, 256, 0, s>>>(n, d_parent, child_off, child_adj, next_sib);
    int32_t first_child = 0;
    {
      int32_t ro[2] = {0, 0};
      LCA_TRY(cudaMemcpyAsync(ro, child_off + root, 8, cudaMemcpyDeviceToHost, s));
      LCA_TRY(cudaStreamSynchronize(s));
      if (ro[1] == ro[0]) {   // n > 1 and nothing hangs on the root: the other nodes form cycles
        rc = set_error(PT_ERR_INVALID, "pt_lca_build: parent array is not a tree (1 of %lld nodes reachable from the root: cycle)", (long long)n);
        cleanup(); lca_release(h); return rc;
      }
      LCA_TRY(cudaMemcpyAsync(&first_child, child_adj + ro[0], 4, cudaMemcpyDeviceToHost, s));
      LCA_TRY(cudaStreamSynchronize(s));
    }
    const uint32_t head = 2u * (uint32_t)first_child;
    LCA_TRY(cudaMemsetAsync(small + 8, 0, 4, s));
    et_succ_kernel<<<grid, 256, 0, s>>>(n, root, d_parent, child_off, child_adj, next_sib, head, succ, slot, small + 8);
    int32_t S = 0;
    LCA_TRY(cudaMemcpyAsync(&S, small + 8, 4, cudaMemcpyDeviceToHost, s));
    LCA_TRY(cudaStreamSynchronize(s));
    int32_t* sp = nullptr;   // sp_arc | nxt A/B | len A/B | w A/B : 7 arrays of S
    LCA_TRY(lca_malloc((void**)&sp, (size_t)std::max(S, 1) * 7 * 4)); et_owned[2] = sp;
    int32_t* sp_arc = sp; int32_t* nxtA = sp + (size_t)S; int32_t* nxtB = sp + 2 * (size_t)S; int32_t* lenA = sp + 3 * (size_t)S; int32_t* lenB = sp + 4 * (size_t)S;
    int32_t* wA = sp + 5 * (size_t)S; int32_t* wB = sp + 6 * (size_t)S;
    et_walk1_kernel<<<grid, 256, 0, s>>>(n, succ, slot, sp_arc, nxtA, lenA, wA);
    int rounds = 1;
    while (((int64_t)1 << rounds) < (int64_t)S) ++rounds;
    const int jgrid = (int)std::max<int64_t>(1, std::min<int64_t>(grid, ((int64_t)S + 255) / 256));
    for (int r = 0; r <= rounds; ++r) {
      et_jump_kernel<<<jgrid, 256, 0, s>>>(S, nxtA, lenA, wA, nxtB, lenB, wB);
      std::swap(nxtA, nxtB); std::swap(lenA, lenB); std::swap(wA, wB);
    }
    int32_t head_slot = 0;
    LCA_TRY(cudaMemcpyAsync(&head_slot, slot + head, 4, cudaMemcpyDeviceToHost, s));
    LCA_TRY(cudaStreamSynchronize(s));
    LCA_TRY(cudaMemsetAsync(pre, 0xff, n4, s)); LCA_TRY(cudaMemsetAsync(size, 0xff, n4, s));
    et_walk2_kernel<<<jgrid, 256, 0, s>>>(S, n, succ, sp_arc, nxtA, lenA, wA, head_slot, depth, pre, size);
    et_finish_kernel<<<grid, 256, 0, s>>>(n, root, depth, pre, size, small);
  }
  LCA_TRY(cudaMemcpyAsync(info, small, 32, cudaMemcpyDeviceToHost, s));
  LCA_TRY(cudaStreamSynchronize(s));
  if (info[4] != (int32_t)n) {
    rc = set_error(PT_ERR_INVALID, "pt_lca_build: parent array is not a tree (%d of %lld nodes reachable from the root: cycle)", info[4], (long long)n);
    cleanup(); lca_release(h); return rc;
  }
  h->levels = info[3];
  mark(use_sweeps ? "sweeps" : "Euler tour + ranking");
  if (trace) fprintf(stderr, "[lca build n=%lld] %d levels\n", (long long)n, info[3]);
  lca_keys_kernel<<<grid, 256, 0, s>>>(n, d_parent, depth, pre, h->keypos, nodeat);
  lca_blocks_kernel<true><<<grid, 256, 0, s>>>(n, h->nblocks, h->keypos, nodeat, h->rec, h->table);
  for (size_t k = 1; k < level_base.size(); ++k) {
    const int64_t count = h->nblocks - ((int64_t)1 << k) + 1;
    const int g = (int)std::min<int64_t>(grid, (count + 255) / 256);
    lca_table_level_kernel<<<std::max(g, 1), 256, 0, s>>>(h->table + level_base[k - 1], h->table + level_base[k], count, (int64_t)1 << (k - 1));
  }
  LCA_TRY(cudaGetLastError());
  LCA_TRY(cudaEventRecord(ev1, s));
  LCA_TRY(cudaStreamSynchronize(s));
  LCA_TRY(cudaEventElapsedTime(&h->build_ms, ev0, ev1));
  mark("records + table");
#undef LCA_TRY
  if (keep_size) { h->size = size; h->nodeat = nodeat; }
  cleanup();
  *out = h;
  return PT_OK;
}

int pt_lca_build(pt_lca** out, const int32_t* parent, int64_t n, pt_mem parent_mem, void* stream) { return lca_build_impl(out, parent, n, parent_mem, stream, false); }

int pt_lca_query(const pt_lca* hc, const int32_t* u, const int32_t* v, int32_t* out, int64_t q, pt_mem mem, void* stream) {
  pt_lca* h = const_cast<pt_lca*>(hc);
  if (!h || q < 0 || (q > 0 && (!u || !v || !out))) return set_error(PT_ERR_INVALID, "pt_lca_query: bad argument");
  if (q == 0) return PT_OK;
  if (int rc = require_device()) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  DeviceProps dp;
  if (int rc = device_props(&dp)) return rc;
  const int32_t *du = u, *dv = v; int32_t* dout = out;
  if (mem == PT_MEM_HOST) {
    if (h->qcap < q) {
      cudaStreamSynchronize(s);
      dev_free(h->d_u); dev_free(h->d_v); dev_free(h->d_out); h->d_u = h->d_v = h->d_out = nullptr; h->qcap = 0;
      PT_CUDA(lca_malloc((void**)&h->d_u, (size_t)q * 4)); PT_CUDA(lca_malloc((void**)&h->d_v, (size_t)q * 4)); PT_CUDA(lca_malloc((void**)&h->d_out, (size_t)q * 4));
      h->qcap = q;
    }
    PT_CUDA(cudaMemcpyAsync(h->d_u, u, (size_t)q * 4, cudaMemcpyHostToDevice, s));
    PT_CUDA(cudaMemcpyAsync(h->d_v, v, (size_t)q * 4, cudaMemcpyHostToDevice, s));
    du = h->d_u; dv = h->d_v; dout = h->d_out;
  }
  constexpr int QPT = 4;
  const int64_t threads = (q + QPT - 1) / QPT;
  const int grid = (int)std::min<int64_t>((threads + 255) / 256, (int64_t)dp.sms * 8);
  lca_query_kernel<QPT><<<std::max(grid, 1), 256, 0, s>>>(q, h->n, du, dv, dout, h->rec, h->keypos, h->table, h->d_level_base);
  PT_CUDA(cudaGetLastError());
  if (mem == PT_MEM_HOST) {
    PT_CUDA(cudaMemcpyAsync(out, dout, (size_t)q * 4, cudaMemcpyDeviceToHost, s));
    PT_CUDA(cudaStreamSynchronize(s));
  }
  return PT_OK;
}

int pt_lca_query_adjacent(const pt_lca* hc, const int32_t* nodes, int64_t count, int32_t* out, pt_mem mem, void* stream) {
  pt_lca* h = const_cast<pt_lca*>(hc);
  if (!h || count < 0 || (count > 1 && (!nodes || !out))) return set_error(PT_ERR_INVALID, "pt_lca_query_adjacent: bad argument");
  if (count < 2) return PT_OK;
  if (int rc = require_device()) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  DeviceProps dp;
  if (int rc = device_props(&dp)) return rc;
  const int64_t q = count - 1;
  const int32_t* dn = nodes; int32_t* dout = out;
  if (mem == PT_MEM_HOST) {
    if (h->qcap < count) {
      cudaStreamSynchronize(s);
      dev_free(h->d_u); dev_free(h->d_v); dev_free(h->d_out); h->d_u = h->d_v = h->d_out = nullptr; h->qcap = 0;
      PT_CUDA(lca_malloc((void**)&h->d_u, (size_t)count * 4)); PT_CUDA(lca_malloc((void**)&h->d_v, (size_t)count * 4)); PT_CUDA(lca_malloc((void**)&h->d_out, (size_t)count * 4));
      h->qcap = count;
    }
    PT_CUDA(cudaMemcpyAsync(h->d_u, nodes, (size_t)count * 4, cudaMemcpyHostToDevice, s));
    dn = h->d_u; dout = h->d_out;
  }
  constexpr int QPT = 4;
  const int grid = (int)std::min<int64_t>((q + 256 * QPT - 1) / (256 * QPT), (int64_t)dp.sms * 8);
  lca_adjacent_kernel<QPT><<<std::max(grid, 1), 256, 0, s>>>(q, h->n, dn, dout, h->rec, h->keypos, h->table, h->d_level_base);
  PT_CUDA(cudaGetLastError());
  if (mem == PT_MEM_HOST) {
    PT_CUDA(cudaMemcpyAsync(out, dout, (size_t)q * 4, cudaMemcpyDeviceToHost, s));
    PT_CUDA(cudaStreamSynchronize(s));
  }
  return PT_OK;
}

int pt_lca_get_info(const pt_lca* h, pt_lca_info* info) {
  if (!h || !info) return set_error(PT_ERR_INVALID, "pt_lca_get_info: null argument");
  info->num_nodes = h->n; info->num_levels = h->levels; info->num_blocks = h->nblocks; info->table_levels = h->table_levels;
  info->device_bytes = h->device_bytes; info->build_ms = h->build_ms;
  return PT_OK;
}

int pt_lca_node_info(const pt_lca* h, int32_t* depth, int32_t* preorder, pt_mem mem, void* stream) {
  if (!h) return set_error(PT_ERR_INVALID, "pt_lca_node_info: null argument");
  if (int rc = require_device()) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  int32_t *dd = depth, *dp = preorder;
  const size_t n4 = (size_t)h->n * 4;
  if (mem == PT_MEM_HOST) { dd = dp = nullptr; if (depth) PT_CUDA(lca_malloc((void**)&dd, n4)); if (preorder) PT_CUDA(lca_malloc((void**)&dp, n4)); }
  lca_node_info_kernel<<<592, 256, 0, s>>>(h->n, h->rec, dd, dp);
  PT_CUDA(cudaGetLastError());
  if (mem == PT_MEM_HOST) {
    if (depth) PT_CUDA(cudaMemcpyAsync(depth, dd, n4, cudaMemcpyDeviceToHost, s));
    if (preorder) PT_CUDA(cudaMemcpyAsync(preorder, dp, n4, cudaMemcpyDeviceToHost, s));
    PT_CUDA(cudaStreamSynchronize(s));
    dev_free(dd); dev_free(dp);
  }
  return PT_OK;
}

int pt_lca_free(pt_lca* h) { lca_release(h); return PT_OK; }

}  // extern "C"

// ---- RMQ API --------------------------------------------------------------------------------------------------------
struct pt_rmq {
  int64_t n = 0, nblocks = 0;
  LcaRecord* rec = nullptr; u64* keypos = nullptr; u64* table = nullptr; int64_t* d_level_base = nullptr;
};
static void rmq_release(pt_rmq* h) { if (!h) return; cudaDeviceSynchronize(); dev_free(h->rec); dev_free(h->keypos); dev_free(h->table); dev_free(h->d_level_base); delete h; }

extern "C" {

int pt_rmq_build(pt_rmq** out, const int32_t* values, int64_t n, pt_mem mem, void* stream) {
  if (!out || !values || n <= 0 || n >= ((int64_t)1 << 32) - 64) return set_error(PT_ERR_INVALID, "pt_rmq_build: bad argument");
  *out = nullptr;
  if (int rc = require_device()) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  DeviceProps dp;
  if (int rc = device_props(&dp)) return rc;
  pt_rmq* h = new (std::nothrow) pt_rmq();
  if (!h) return set_error(PT_ERR_NOMEM, "out of host memory");
  h->n = n; h->nblocks = (n + 31) / 32;
  int32_t* dvals = nullptr;
  auto fail = [&](int code) { cudaStreamSynchronize(s); if (mem == PT_MEM_HOST) dev_free(dvals); rmq_release(h); return code; };
#define RMQ_TRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return fail(set_error(e__ == cudaErrorMemoryAllocation ? PT_ERR_NOMEM : PT_ERR_CUDA, "%s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e__))); } while (0)
  if (mem == PT_MEM_HOST) { RMQ_TRY(lca_malloc((void**)&dvals, (size_t)n * 4)); RMQ_TRY(cudaMemcpyAsync(dvals, values, (size_t)n * 4, cudaMemcpyHostToDevice, s)); }
  else dvals = (int32_t*)values;
  std::vector<int64_t> level_base; int64_t total = 0;
  for (int k = 0; ((int64_t)1 << k) <= h->nblocks; ++k) { level_base.push_back(total); total += h->nblocks - ((int64_t)1 << k) + 1; }
  RMQ_TRY(lca_malloc((void**)&h->rec, (size_t)n * sizeof(LcaRecord)));
  RMQ_TRY(lca_malloc((void**)&h->keypos, (size_t)n * 8));
  RMQ_TRY(lca_malloc((void**)&h->table, (size_t)total * 8));
  RMQ_TRY(lca_malloc((void**)&h->d_level_base, level_base.size() * 8));
  RMQ_TRY(cudaMemcpyAsync(h->d_level_base, level_base.data(), level_base.size() * 8, cudaMemcpyHostToDevice, s));
  const int grid = dp.sms * 8;
  rmq_keys_kernel<<<grid, 256, 0, s>>>(n, dvals, h->keypos);
  lca_blocks_kernel<false><<<grid, 256, 0, s>>>(n, h->nblocks, h->keypos, nullptr, h->rec, h->table);
  for (size_t k = 1; k < level_base.size(); ++k) {
    const int64_t count = h->nblocks - ((int64_t)1 << k) + 1;
    const int g = (int)std::min<int64_t>(grid, (count + 255) / 256);
    lca_table_level_kernel<<<std::max(g, 1), 256, 0, s>>>(h->table + level_base[k - 1], h->table + level_base[k], count, (int64_t)1 << (k - 1));
  }
  RMQ_TRY(cudaGetLastError());
  RMQ_TRY(cudaStreamSynchronize(s));
#undef RMQ_TRY
  if (mem == PT_MEM_HOST) dev_free(dvals);
  *out = h;
  return PT_OK;
}

int pt_rmq_query(const pt_rmq* h, const int64_t* l, const int64_t* r, int64_t* out_index, int64_t q, pt_mem mem, void* stream) {
  if (!h || q < 0 || (q > 0 && (!l || !r || !out_index))) return set_error(PT_ERR_INVALID, "pt_rmq_query: bad argument");
  if (q == 0) return PT_OK;
  if (int rc = require_device()) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t *dl = l, *dr = r; int64_t* dout = out_index;
  int64_t* buf = nullptr;
  if (mem == PT_MEM_HOST) {
    PT_CUDA(lca_malloc((void**)&buf, (size_t)q * 24));
    PT_CUDA(cudaMemcpyAsync(buf, l, (size_t)q * 8, cudaMemcpyHostToDevice, s));
    PT_CUDA(cudaMemcpyAsync(buf + q, r, (size_t)q * 8, cudaMemcpyHostToDevice, s));
    dl = buf; dr = buf + q; dout = buf + 2 * q;
  }
  const int grid = (int)std::min<int64_t>((q + 255) / 256, 148 * 16);
  rmq_query_kernel<<<std::max(grid, 1), 256, 0, s>>>(q, h->n, dl, dr, dout, h->rec, h->keypos, h->table, h->d_level_base);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && mem == PT_MEM_HOST) { e = cudaMemcpyAsync(out_index, dout, (size_t)q * 8, cudaMemcpyDeviceToHost, s); if (e == cudaSuccess) e = cudaStreamSynchronize(s); }
  if (mem == PT_MEM_HOST) { cudaStreamSynchronize(s); dev_free(buf); }
  if (e != cudaSuccess) return set_error(PT_ERR_CUDA, "pt_rmq_query: %s", cudaGetErrorString(e));
  return PT_OK;
}

int pt_rmq_free(pt_rmq* h) { rmq_release(h); return PT_OK; }

}  // extern "C"

// ---- who_displays API ------------------------------------------------------------------------------------------------
extern "C" int pt_tree_who_displays(const int32_t* host_parent, const int32_t* host_label, int64_t n, const int32_t* guest_parent,
                                    const int32_t* guest_label, int64_t nt, int32_t* out_host_node, pt_mem mem, void* stream) {
  if (!host_parent || !host_label || !guest_parent || !guest_label || !out_host_node || n <= 0 || nt <= 0 || n > INT32_MAX / 2 || nt > INT32_MAX / 2)
    return set_error(PT_ERR_INVALID, "pt_tree_who_displays: bad argument");
  if (int rc = require_device()) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  DeviceProps dp;
  if (int rc = device_props(&dp)) return rc;
  std::vector<void*> owned;
  pt_lca* H = nullptr;
  auto cleanup = [&]() { cudaStreamSynchronize(s); for (void* p : owned) dev_free(p); if (H) pt_lca_free(H); };
  auto dalloc = [&](size_t bytes) -> void* { void* p = nullptr; if (lca_malloc(&p, bytes ? bytes : 4) != cudaSuccess) { cudaGetLastError(); return nullptr; } owned.push_back(p); return p; };
#define WD_TRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup(); return set_error(PT_ERR_CUDA, "%s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); } } while (0)
#define WD_PTR(p) do { if (!(p)) { cleanup(); return set_error(PT_ERR_NOMEM, "pt_tree_who_displays: out of device memory"); } } while (0)
  const int32_t *hp = host_parent, *hl = host_label, *gp = guest_parent, *gl = guest_label;
  if (mem == PT_MEM_HOST) {
    int32_t* a = (int32_t*)dalloc((size_t)n * 4); int32_t* b = (int32_t*)dalloc((size_t)n * 4); int32_t* c = (int32_t*)dalloc((size_t)nt * 4); int32_t* d = (int32_t*)dalloc((size_t)nt * 4);
    WD_PTR(a); WD_PTR(b); WD_PTR(c); WD_PTR(d);
    WD_TRY(cudaMemcpyAsync(a, host_parent, (size_t)n * 4, cudaMemcpyHostToDevice, s)); WD_TRY(cudaMemcpyAsync(b, host_label, (size_t)n * 4, cudaMemcpyHostToDevice, s));
    WD_TRY(cudaMemcpyAsync(c, guest_parent, (size_t)nt * 4, cudaMemcpyHostToDevice, s)); WD_TRY(cudaMemcpyAsync(d, guest_label, (size_t)nt * 4, cudaMemcpyHostToDevice, s));
    hp = a; hl = b; gp = c; gl = d;
  }
  if (int rc = pt_lca_build(&H, hp, n, PT_MEM_DEVICE, stream)) { cleanup(); return rc; }
  const int grid = dp.sms * 8;
  const int64_t nlab = n + nt;
  int32_t* hdeg = (int32_t*)dalloc((size_t)n * 4 + 4); int32_t* small = (int32_t*)dalloc(128); int32_t* l2h = (int32_t*)dalloc((size_t)nlab * 4);
  int32_t* tdeg = (int32_t*)dalloc((size_t)nt * 4 + 4); int32_t* toff = (int32_t*)dalloc((size_t)nt * 4 + 4); int32_t* tadj = (int32_t*)dalloc((size_t)nt * 4);
  int32_t* disp = (int32_t*)dalloc((size_t)nt * 4);
  int64_t* scan_tmp = (int64_t*)dalloc(((size_t)((std::max(n, nt) + SCAN_TILE - 1) / SCAN_TILE) + 1) * 8);
  WD_PTR(hdeg); WD_PTR(small); WD_PTR(l2h); WD_PTR(tdeg); WD_PTR(toff); WD_PTR(tadj); WD_PTR(disp); WD_PTR(scan_tmp);
  WD_TRY(cudaMemsetAsync(hdeg, 0, (size_t)n * 4 + 4, s)); WD_TRY(cudaMemsetAsync(tdeg, 0, (size_t)nt * 4 + 4, s)); WD_TRY(cudaMemsetAsync(small, 0, 128, s));
  WD_TRY(cudaMemsetAsync(l2h, 0xff, (size_t)nlab * 4, s));
  lca_count_kernel<<<grid, 256, 0, s>>>(hp, n, hdeg, small);           // host out-degrees (leaf test); host validity was checked by pt_lca_build
  lca_count_kernel<<<grid, 256, 0, s>>>(gp, nt, tdeg, small + 8);      // guest children counts + root
  wd_label_map_kernel<<<grid, 256, 0, s>>>(n, hdeg, hl, nlab, l2h, small + 16);
  if (int rc = device_exclusive_scan(tdeg, nt, toff, scan_tmp, s)) { cleanup(); return rc; }
  WD_TRY(cudaMemsetAsync(tdeg, 0, (size_t)nt * 4, s));
  lca_fill_children_kernel<<<grid, 256, 0, s>>>(gp, nt, toff, tdeg, tadj);
  int32_t info[32];
  WD_TRY(cudaMemcpyAsync(info, small, 128, cudaMemcpyDeviceToHost, s));
  WD_TRY(cudaStreamSynchronize(s));
  if (info[8] != 1 || info[10]) { cleanup(); return set_error(PT_ERR_INVALID, "pt_tree_who_displays: guest is not a rooted tree (%d roots)", info[8]); }
  if (info[17]) { cleanup(); return set_error(PT_ERR_INVALID, "single-label map for multi-labeled tree/network"); }  // label_matching.hpp:51-52
  if (info[18]) { cleanup(); return set_error(PT_ERR_INVALID, "pt_tree_who_displays: label id out of range (must be < n + nt)"); }
  {
    int32_t* pending = (int32_t*)dalloc((size_t)nt * 4 + 4); unsigned long long* processed = (unsigned long long*)dalloc(16);
    u64* big = (u64*)dalloc((size_t)nt * 8);   // sort space of the guest nodes with more than WD_MAX_CHILDREN children
    WD_PTR(pending); WD_PTR(processed); WD_PTR(big);
    WD_TRY(cudaMemcpyAsync(pending, tdeg, (size_t)nt * 4, cudaMemcpyDeviceToDevice, s));   // children still to come, per guest node
    WD_TRY(cudaMemsetAsync(processed, 0, 16, s));
    who_displays_kernel<<<(unsigned)std::max<int64_t>(1, std::min<int64_t>((int64_t)dp.sms * 16, (nt + 255) / 256)), 256, 0, s>>>(
        nt, gp, pending, toff, tadj, gl, l2h, nlab, H->rec, H->keypos, H->table, H->d_level_base, disp, big, processed);
    WD_TRY(cudaGetLastError());
    unsigned long long hp2 = 0;
    WD_TRY(cudaMemcpyAsync(&hp2, processed, 8, cudaMemcpyDeviceToHost, s));
    WD_TRY(cudaStreamSynchronize(s));
    if (hp2 != (unsigned long long)nt) { cleanup(); return set_error(PT_ERR_INVALID, "pt_tree_who_displays: guest parent array has a cycle"); }
  }
  wd_finish_kernel<<<grid, 256, 0, s>>>(nt, disp);
  WD_TRY(cudaMemcpyAsync(info, small, 128, cudaMemcpyDeviceToHost, s));
  if (mem == PT_MEM_HOST) WD_TRY(cudaMemcpyAsync(out_host_node, disp, (size_t)nt * 4, cudaMemcpyDeviceToHost, s));
  else WD_TRY(cudaMemcpyAsync(out_host_node, disp, (size_t)nt * 4, cudaMemcpyDeviceToDevice, s));
  WD_TRY(cudaStreamSynchronize(s));
#undef WD_TRY
#undef WD_PTR
  cleanup();
  return PT_OK;
}

// ---- highest_displayed_ancestor for every guest node at once (TreeInComponent::highest_displayed_ancestor, containment.hpp:328-344)
// The reference climbs from v: parent not displayed -> stop at v; else move up and stop there if that node is the guest root or is
// displayed by the host root alone.  "First ancestor-or-self x of parent(v) with stop(x)" is pointer jumping: O(log depth) rounds.
namespace ptgpu {
__global__ void hda_init_kernel(int64_t nt, const int32_t* __restrict__ gparent, const int32_t* __restrict__ who, int32_t host_root, int32_t* __restrict__ jump) {
  for (int64_t x = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; x < nt; x += (int64_t)gridDim.x * blockDim.x) {
    const int32_t px = gparent[x];
    const bool stop = px < 0 || who[x] == host_root || who[px] < 0;   // a climb that reached x ends here
    jump[x] = stop ? (int32_t)x : px;
  }
}
__global__ void hda_jump_kernel(int64_t nt, const int32_t* __restrict__ in, int32_t* __restrict__ out, int32_t* __restrict__ changed) {
  bool any = false;
  for (int64_t x = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; x < nt; x += (int64_t)gridDim.x * blockDim.x) {
    const int32_t j = in[x], jj = in[j];
    out[x] = jj; any |= jj != j;
  }
  if (any) *changed = 1;
}
__global__ void hda_finish_kernel(int64_t nt, const int32_t* __restrict__ gparent, const int32_t* __restrict__ who, const int32_t* __restrict__ jump, int32_t* __restrict__ out) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nt; v += (int64_t)gridDim.x * blockDim.x) {
    const int32_t pv = gparent[v];
    out[v] = (pv < 0 || who[pv] < 0) ? (int32_t)v : jump[pv];
  }
}
}  // namespace ptgpu

extern "C" int pt_tree_highest_displayed(const int32_t* guest_parent, const int32_t* who, int64_t nt, int32_t host_root, int32_t* out, pt_mem mem, void* stream) {
  if (!guest_parent || !who || !out || nt <= 0 || nt > INT32_MAX / 2 || host_root < 0) return set_error(PT_ERR_INVALID, "pt_tree_highest_displayed: bad argument");
  if (int rc = require_device()) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  DeviceProps dp;
  if (int rc = device_props(&dp)) return rc;
  std::vector<void*> owned;
  auto cleanup = [&]() { cudaStreamSynchronize(s); for (void* p : owned) dev_free(p); };
  auto dalloc = [&](size_t bytes) -> void* { void* p = nullptr; if (lca_malloc(&p, bytes ? bytes : 4) != cudaSuccess) { cudaGetLastError(); return nullptr; } owned.push_back(p); return p; };
#define HD_TRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup(); return set_error(PT_ERR_CUDA, "%s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); } } while (0)
  const size_t bytes = (size_t)nt * 4;
  int32_t *gp = (int32_t*)guest_parent, *wd = (int32_t*)who, *res = out;
  int32_t* a = (int32_t*)dalloc(bytes); int32_t* b = (int32_t*)dalloc(bytes); int32_t* flag = (int32_t*)dalloc(4);
  if (!a || !b || !flag) { cleanup(); return set_error(PT_ERR_NOMEM, "pt_tree_highest_displayed: out of device memory"); }
  if (mem == PT_MEM_HOST) {
    gp = (int32_t*)dalloc(bytes); wd = (int32_t*)dalloc(bytes); res = (int32_t*)dalloc(bytes);
    if (!gp || !wd || !res) { cleanup(); return set_error(PT_ERR_NOMEM, "pt_tree_highest_displayed: out of device memory"); }
    HD_TRY(cudaMemcpyAsync(gp, guest_parent, bytes, cudaMemcpyHostToDevice, s)); HD_TRY(cudaMemcpyAsync(wd, who, bytes, cudaMemcpyHostToDevice, s));
  }
  const int grid = (int)std::min<int64_t>((nt + 255) / 256, (int64_t)dp.sms * 8);
  ptgpu::hda_init_kernel<<<grid, 256, 0, s>>>(nt, gp, wd, host_root, a);
  for (int round = 0; round < 40; ++round) {   // the distance covered doubles every round
    int32_t changed = 0;
    HD_TRY(cudaMemsetAsync(flag, 0, 4, s));
    ptgpu::hda_jump_kernel<<<grid, 256, 0, s>>>(nt, a, b, flag);
    HD_TRY(cudaMemcpyAsync(&changed, flag, 4, cudaMemcpyDeviceToHost, s));
    HD_TRY(cudaStreamSynchronize(s));
    std::swap(a, b);
    if (!changed) break;
  }
  ptgpu::hda_finish_kernel<<<grid, 256, 0, s>>>(nt, gp, wd, a, res);
  HD_TRY(cudaGetLastError());
  if (mem == PT_MEM_HOST) HD_TRY(cudaMemcpyAsync(out, res, bytes, cudaMemcpyDeviceToHost, s));
  HD_TRY(cudaStreamSynchronize(s));
#undef HD_TRY
  cleanup();
  return PT_OK;
}

// ---- bridges / 2-edge-connected components of a rooted network -------------------------------------------------------------
// BridgeFinder utils/bridges.hpp:42-118 (two recursive DFS passes over DFS intervals) re-designed as data-parallel steps:
//   1. spanning tree: every node keeps its first in-arc (a rooted DAG with one root gives a rooted tree)
//   2. pt_lca on that tree (preorder positions, subtree sizes)
//   3. every other arc (u,v) closes a cycle through a = lca(u,v): cnt[u]++, cnt[v]++, cnt[a] -= 2 (one batched LCA pass)
//   4. prefix sums of cnt in preorder: the tree arc into w is covered iff the sum over w's preorder interval is > 0
//   5. a tree arc that no cycle covers is a bridge; components = pointer jumping to the nearest ancestor entered by a bridge
namespace ptgpu {
__global__ void br_first_parent_kernel(int64_t n, int64_t m, const int32_t* __restrict__ succ_off, const int32_t* __restrict__ succ_adj, int32_t* __restrict__ tree_arc) {
  // tree_arc[v] = smallest arc id entering v
  for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < n; u += (int64_t)gridDim.x * blockDim.x)
    for (int32_t e = succ_off[u]; e < succ_off[u + 1]; ++e) atomicMin(&tree_arc[succ_adj[e]], e);
}
__global__ void br_tree_parent_kernel(int64_t n, const int32_t* __restrict__ succ_off, const int32_t* __restrict__ succ_adj, const int32_t* __restrict__ tree_arc,
                                      int32_t* __restrict__ arc_tail, int32_t* __restrict__ parent) {
  for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < n; u += (int64_t)gridDim.x * blockDim.x) {
    for (int32_t e = succ_off[u]; e < succ_off[u + 1]; ++e) { arc_tail[e] = (int32_t)u; if (tree_arc[succ_adj[e]] == e) parent[succ_adj[e]] = (int32_t)u; }
  }
}
__global__ void br_cover_kernel(int64_t m, const int32_t* __restrict__ arc_tail, const int32_t* __restrict__ succ_adj, const int32_t* __restrict__ tree_arc,
                                const LcaRecord* __restrict__ rec, const u64* __restrict__ keypos, const u64* __restrict__ table, const int64_t* __restrict__ level_base,
                                int32_t* __restrict__ cntpos) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < m; e += (int64_t)gridDim.x * blockDim.x) {
    const int32_t u = arc_tail[e], v = succ_adj[e];
    if (tree_arc[v] == (int32_t)e) continue;
    const int32_t a = lca_one(rec, keypos, table, level_base, u, v);
    atomicAdd(&cntpos[rec[u].pos], 1); atomicAdd(&cntpos[rec[v].pos], 1); atomicAdd(&cntpos[rec[a].pos], -2);
  }
}
__global__ void br_mark_kernel(int64_t n, const int32_t* __restrict__ parent, const int32_t* __restrict__ tree_arc, const LcaRecord* __restrict__ rec,
                               const int32_t* __restrict__ size, const int32_t* __restrict__ prefix, uint8_t* __restrict__ is_bridge, int32_t* __restrict__ up) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
    if (parent[v] < 0) { up[v] = (int32_t)v; continue; }
    const uint32_t p = rec[v].pos;
    const bool covered = prefix[p + size[v]] - prefix[p] > 0;
    if (!covered) is_bridge[tree_arc[v]] = 1;
    up[v] = covered ? parent[v] : (int32_t)v;
  }
}
__global__ void br_jump_kernel(int64_t n, int32_t* __restrict__ up, int32_t* __restrict__ changed) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
    const int32_t a = up[v], b = up[a];
    if (a != b) { up[v] = b; *changed = 1; }
  }
}
}  // namespace ptgpu

extern "C" int pt_net_bridges(int64_t n, int64_t m, const int32_t* succ_off, const int32_t* succ_adj, uint8_t* is_bridge, int32_t* component, pt_mem mem, void* stream) {
  if (n <= 0 || m < 0 || !succ_off || (!succ_adj && m > 0) || !is_bridge || n > INT32_MAX / 2 || m > INT32_MAX / 2) return set_error(PT_ERR_INVALID, "pt_net_bridges: bad argument");
  if (int rc = require_device()) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  DeviceProps dp;
  if (int rc = device_props(&dp)) return rc;
  std::vector<void*> owned; pt_lca* H = nullptr;
  auto cleanup = [&]() { cudaStreamSynchronize(s); for (void* p : owned) dev_free(p); if (H) pt_lca_free(H); };
  auto dalloc = [&](size_t bytes) -> void* { void* p = nullptr; if (lca_malloc(&p, bytes ? bytes : 4) != cudaSuccess) { cudaGetLastError(); return nullptr; } owned.push_back(p); return p; };
#define BR_TRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup(); return set_error(PT_ERR_CUDA, "%s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); } } while (0)
#define BR_PTR(p) do { if (!(p)) { cleanup(); return set_error(PT_ERR_NOMEM, "pt_net_bridges: out of device memory"); } } while (0)
  const int32_t *so = succ_off, *sa = succ_adj;
  if (mem == PT_MEM_HOST) {
    int32_t* a = (int32_t*)dalloc((size_t)(n + 1) * 4); int32_t* b = (int32_t*)dalloc((size_t)m * 4); BR_PTR(a); BR_PTR(b);
    BR_TRY(cudaMemcpyAsync(a, succ_off, (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, s)); if (m) BR_TRY(cudaMemcpyAsync(b, succ_adj, (size_t)m * 4, cudaMemcpyHostToDevice, s));
    so = a; sa = b;
  }
  const int grid = dp.sms * 8;
  int32_t* tree_arc = (int32_t*)dalloc((size_t)n * 4); int32_t* arc_tail = (int32_t*)dalloc((size_t)m * 4); int32_t* parent = (int32_t*)dalloc((size_t)n * 4);
  int32_t* cntpos = (int32_t*)dalloc((size_t)(n + 1) * 4); int32_t* prefix = (int32_t*)dalloc((size_t)(n + 1) * 4); int32_t* up = (int32_t*)dalloc((size_t)n * 4);
  uint8_t* d_br = (uint8_t*)dalloc((size_t)m + 1); int32_t* flag = (int32_t*)dalloc(16);
  int64_t* scan_tmp = (int64_t*)dalloc(((size_t)((n + SCAN_TILE - 1) / SCAN_TILE) + 1) * 8);
  BR_PTR(tree_arc); BR_PTR(arc_tail); BR_PTR(parent); BR_PTR(cntpos); BR_PTR(prefix); BR_PTR(up); BR_PTR(d_br); BR_PTR(flag); BR_PTR(scan_tmp);
  BR_TRY(cudaMemsetAsync(tree_arc, 0x7f, (size_t)n * 4, s)); BR_TRY(cudaMemsetAsync(parent, 0xff, (size_t)n * 4, s)); BR_TRY(cudaMemsetAsync(cntpos, 0, (size_t)(n + 1) * 4, s));
  BR_TRY(cudaMemsetAsync(d_br, 0, (size_t)m + 1, s));
  br_first_parent_kernel<<<grid, 256, 0, s>>>(n, m, so, sa, tree_arc);
  br_tree_parent_kernel<<<grid, 256, 0, s>>>(n, so, sa, tree_arc, arc_tail, parent);
  BR_TRY(cudaGetLastError());
  if (int rc = lca_build_impl(&H, parent, n, PT_MEM_DEVICE, stream, true)) { cleanup(); return rc; }  // also validates: one root, no cycle
  br_cover_kernel<<<grid, 256, 0, s>>>(m, arc_tail, sa, tree_arc, H->rec, H->keypos, H->table, H->d_level_base, cntpos);
  if (int rc = device_exclusive_scan(cntpos, n, prefix, scan_tmp, s)) { cleanup(); return rc; }
  br_mark_kernel<<<grid, 256, 0, s>>>(n, parent, tree_arc, H->rec, H->size, prefix, d_br, up);
  for (int it = 0; it < 40; ++it) {
    int32_t h = 0;
    BR_TRY(cudaMemsetAsync(flag, 0, 4, s));
    br_jump_kernel<<<grid, 256, 0, s>>>(n, up, flag);
    BR_TRY(cudaMemcpyAsync(&h, flag, 4, cudaMemcpyDeviceToHost, s));
    BR_TRY(cudaStreamSynchronize(s));
    if (!h) break;
  }
  const cudaMemcpyKind kind = mem == PT_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  if (m) BR_TRY(cudaMemcpyAsync(is_bridge, d_br, (size_t)m, kind, s));
  if (component) BR_TRY(cudaMemcpyAsync(component, up, (size_t)n * 4, kind, s));
  BR_TRY(cudaStreamSynchronize(s));
#undef BR_TRY
#undef BR_PTR
  cleanup();
  return PT_OK;
}

#include "dp_kernels.cuh"
#include "split_kernels.cuh"
