// csrc/newick_kernels.cu -- extended-Newick text -> device-resident batch, parsed ON THE GPU (pt_batch_create_from_newick).
//
// SURVEY 8(f) rank 1.  The reference reads one network at a time (NewickParser io/newick.hpp:64-272, read_edgelists
// io/io.hpp:52-62); the host batch reader (pt/fast_newick.hpp) does 2.6e5 pairs/s on 16 cores, fifty times slower than the
// containment kernel consumes them.  Here the text is uploaded once and
//   1. the newlines are located (nw_newline_kernel, two passes), the lines trimmed and the blank ones dropped (one thread per line),
//   2. every line is parsed by ONE WARP (nw_parse_line_warp; odd or malformed lines by one thread, nw_parse_line_seq) with the
//      result of the reference's right-to-left scan (same node numbering: root 0, ids in the order the name tokens are met from
//      the right; hybrid "#H<k>": first occurrence from the right creates the node; ":length" and white space skipped) into
//      scratch arrays addressed by the line's text offset,
//   3. one warp per pair picks host and guest (the network is the host, or the first line if both are trees,
//      examples/tc.cpp:110-111), gives the labels dense ids by first appearance (host nodes, then guest nodes) through an
//      open-addressing table kept in the pair's own scratch span, and publishes the sizes,
//   4. prefix sums over the pairs give the offset tables,
//   5. one warp per pair writes the CSR (children in reverse arc order, utils/storage_adj_immutable.hpp:119), the label
//      ids and the guest parent array -- directly in the int16 layout when the batch allows it.
// The arrays are bit-identical to BatchPacker::add_newick_text on the same text (tests/test_gpu_newick.py); the handle
// owns them (no host copy ever exists).  Malformed input: PT_ERR_PARSE / PT_ERR_INVALID naming the first bad pair.
#include <cub/device/device_scan.cuh>
#include <cub/block/block_scan.cuh>
#include <algorithm>
#include <vector>
#include "cuda_common.cuh"

extern "C" int ptgpu_batch_adopt(pt_batch** out, int64_t B, void* const arrays[8], int index_bytes, const int32_t maxima[4], void* stream);

namespace ptgpu {

// ---- the newline index: two passes over the text, 64 bytes per thread as four 128-bit loads (the text is 16-byte aligned and
// zero-padded).  Pass 1 counts per block of 16 KB, a scan over the block counts gives each block its place, pass 2 writes the
// positions in order.  (A CUB select over 64-bit positions took 0.42 ms per 121 MB and needed 8 bytes of output space per character.)
constexpr int NL_THREADS = 256, NL_BYTES = 64;
__device__ __forceinline__ unsigned nl_mask16(uint4 w) {  // bit i = byte i of the 16 is '\n'
  unsigned m = 0;
  const unsigned v[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const unsigned x = v[k] ^ 0x0a0a0a0au;                            // zero bytes where '\n'
    const unsigned z = ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x | 0x7f7f7f7fu);   // 0x80 in every zero byte (exact, no carries across bytes)
    m |= (((z >> 7) & 1u) | ((z >> 14) & 2u) | ((z >> 21) & 4u) | ((z >> 28) & 8u)) << (4 * k);
  }
  return m;
}
template <bool WRITE>
__global__ void __launch_bounds__(NL_THREADS) nw_newline_kernel(const char* __restrict__ text, int64_t len, int32_t* __restrict__ block_count,
                                                                const int32_t* __restrict__ block_off, int64_t* __restrict__ nlpos) {
  typedef cub::BlockScan<int, NL_THREADS> Scan;
  __shared__ typename Scan::TempStorage tmp;
  const int64_t base = ((int64_t)blockIdx.x * NL_THREADS + threadIdx.x) * NL_BYTES;
  unsigned m[NL_BYTES / 16];
  int c = 0;
#pragma unroll
  for (int k = 0; k < NL_BYTES / 16; ++k) {
    const int64_t o = base + 16 * k;
    m[k] = 0;
    if (o < len) {
      m[k] = nl_mask16(__ldg(reinterpret_cast<const uint4*>(text + o)));
      if (o + 16 > len) m[k] &= (1u << (int)(len - o)) - 1u;   // the padding is zeros, but stay exact
      c += __popc(m[k]);
    }
  }
  int before = 0, total = 0;
  Scan(tmp).ExclusiveSum(c, before, total);
  if (!WRITE) { if (threadIdx.x == 0) block_count[blockIdx.x] = total; return; }
  int64_t* out = nlpos + block_off[blockIdx.x] + before;
#pragma unroll
  for (int k = 0; k < NL_BYTES / 16; ++k) {
    unsigned mm = m[k];
    while (mm) { const int b = __ffs(mm) - 1; mm &= mm - 1; *out++ = base + 16 * k + b; }
  }
}

// raw line r = text between newline r-1 and newline r (virtual newlines at -1 and len); trimmed like pt/fast_newick.hpp
__global__ void nw_trim_kernel(const char* __restrict__ text, int64_t len, const int64_t* __restrict__ nlpos, int64_t num_nl,
                               int64_t* __restrict__ lbeg, int32_t* __restrict__ llen, int32_t* __restrict__ flag) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r > num_nl) return;
  int64_t b = r == 0 ? 0 : nlpos[r - 1] + 1, e = r == num_nl ? len : nlpos[r];
  while (b < e && (text[b] == ' ' || text[b] == '\t' || text[b] == '\r')) ++b;
  while (e > b && (text[e - 1] == ' ' || text[e - 1] == '\t' || text[e - 1] == '\r')) --e;
  lbeg[r] = b; llen[r] = (int32_t)min(e - b, (int64_t)INT32_MAX); flag[r] = e > b;
}
__global__ void nw_compact_kernel(int64_t num_raw, const int64_t* __restrict__ lbeg, const int32_t* __restrict__ llen, const int32_t* __restrict__ flag,
                                  const int32_t* __restrict__ pos, int64_t* __restrict__ obeg, int32_t* __restrict__ olen) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= num_raw || !flag[r]) return;
  obeg[pos[r]] = lbeg[r]; olen[pos[r]] = llen[r];
}

enum { NW_OK = 0, NW_SEMI = 1, NW_HYBRID = 2, NW_START = 3, NW_SEP = 4, NW_GUEST = 5, NW_DOUBLE = 6 };
__host__ const char* nw_error_text(int code) {
  switch (code) {
    case NW_SEMI: return "MalformedNewick: expected ';'";
    case NW_HYBRID: return "MalformedNewick: found '#' but no hybrid number";
    case NW_START: return "MalformedNewick: unexpected start of string";
    case NW_SEP: return "MalformedNewick: expected ',' or '('";
    case NW_GUEST: return "guest must be a tree";
    case NW_DOUBLE: return "MalformedNewick: read double edge";
    default: return "ok";
  }
}

struct NwScratch {  // addressed by region index (region of line k starts at lbeg[k] + 2k, holds llen[k] + 2 entries)
  int2 *arc;  // (tail, head) in reference arc order          -- one 8-byte store per arc: every thread writes its own region, so
  int2 *lab;  // (offset, length) of each node's name in the line   a store costs a 32-byte sector whatever its width
  int32_t *st, *hk, *hi;
};
__device__ __forceinline__ bool nw_ws(char c) { return c == ' ' || (unsigned)(c - 9) < 5u; }
__device__ __forceinline__ bool nw_digit(char c) { return c >= '0' && c <= '9'; }

// the scan of pt/fast_newick.hpp parse_line (= io/newick.hpp:159-272), one thread, one line.  This is the DEFINITION of the numbering
// and of every error; nw_parse_kernel runs it for the lines its warp-parallel path does not take.
__device__ __noinline__ void nw_parse_line_seq(const char* __restrict__ text, int64_t k, const int64_t* __restrict__ lbeg, const int32_t* __restrict__ llen,
                                               NwScratch S, int32_t* __restrict__ ln, int32_t* __restrict__ lm, int32_t* __restrict__ lhyb, int32_t* __restrict__ lerr) {
  // the line is read through a 16-byte register window (one aligned 128-bit load per 16 characters of the backward scan
  // instead of one byte load per character); `text` is 16-byte aligned and padded by the caller
  const int64_t line0 = lbeg[k];
  const int len = llen[k];
  int64_t win_blk = -1;
  uint4 win = make_uint4(0, 0, 0, 0);
  auto CH = [&](int i) -> char {
    const int64_t abs_i = line0 + i, blk = abs_i >> 4;
    if (blk != win_blk) { win = __ldg(reinterpret_cast<const uint4*>(text) + blk); win_blk = blk; }
    const unsigned word = (abs_i & 8) ? ((abs_i & 4) ? win.w : win.z) : ((abs_i & 4) ? win.y : win.x);
    return (char)((word >> ((abs_i & 3) * 8)) & 0xffu);
  };
  const int64_t reg = lbeg[k] + 2 * k;
  int2 *arc = S.arc + reg, *lab = S.lab + reg;
  int32_t *st = S.st + reg, *hk = S.hk + reg, *hi = S.hi + reg;
  int n = 0, m = 0, nh = 0, sp = 0, err = NW_OK, has_h = 0;
  int back = len - 1;
#define NW_SKIP_WS() while (back >= 0 && nw_ws(CH(back))) --back
#define NW_SKIP_LEN() do { int i_ = back; while (i_ >= 0) { const char c_ = CH(i_); if (nw_digit(c_) || c_ == '.' || c_ == '-' || c_ == '+' || c_ == 'E' || c_ == 'e') --i_; else break; } if (i_ >= 0 && CH(i_) == ':') back = i_ - 1; } while (0)
  NW_SKIP_WS();
  if (back >= 0) {
    if (CH(back) != ';') err = NW_SEMI;
    else {
      --back;
      bool want = true;
      int done = 0;
      for (;;) {
        if (want) {
          NW_SKIP_WS();
          int i = back, sharp = -1;
          while (i >= 0) { const char c = CH(i); if (c == '(' || c == ')' || c == ',') break; if (c == '#' && sharp < 0) sharp = i; --i; }
          const int nb = i + 1, nl = back - i;
          back = i;
          int id;
          if (sharp >= 0) {
            int d = sharp;
            while (d < nb + nl && !nw_digit(CH(d))) ++d;
            if (d == nb + nl) { err = NW_HYBRID; break; }
            int key = 0;
            while (d < nb + nl && nw_digit(CH(d))) { key = key * 10 + (CH(d) - '0'); ++d; }
            id = -1;
            for (int h = 0; h < nh; ++h) if (hk[h] == key) { id = hi[h]; break; }
            if (id < 0) { id = n; hk[nh] = key; hi[nh++] = id; lab[n] = make_int2(nb, sharp - nb); ++n; }
            has_h = 1;
          } else { id = n; lab[n] = make_int2(nb, nl); ++n; }
          if (back > 0 && CH(back) == ')') { --back; st[sp++] = id; NW_SKIP_LEN(); continue; }
          done = id; want = false;
        }
        NW_SKIP_WS();
        if (sp == 0) break;
        const int parent = st[sp - 1];
        arc[m++] = make_int2(parent, done);
        if (back < 0) { err = NW_START; break; }
        const char c = CH(back);
        if (c == ',') { --back; NW_SKIP_LEN(); want = true; }
        else if (c == '(') { --back; --sp; done = parent; }
        else { err = NW_SEP; break; }
      }
    }
  }
#undef NW_SKIP_WS
#undef NW_SKIP_LEN
  ln[k] = n; lm[k] = m; lhyb[k] = has_h; lerr[k] = err;
}

// ONE WARP per line, for well-formed lines (everything else -- any error, text left of the tree, a ')' at the very start -- goes to
// nw_parse_line_seq, so errors and oddities keep their sequential definition).  The reader's right-to-left scan is a sequence of
// TOKEN SLOTS (one left of the ';', one left of every ')' and ','; each makes a node or names an existing hybrid) and of ARCS (one
// per ',' and '(': from the owner of the open group to the subtree just finished).  Both are counted with ballots over 32
// characters at a time; "the owner of depth d" is the nearest ')' to the right that opened depth d -- found inside the chunk with
// match.any, else in the per-depth array st[] the previous chunks left behind.  Then one lane per slot scans its own short token
// with the sequential rules (branch length, blanks, '#'), repeated hybrid numbers are folded onto their first slot, ids and names
// are compacted in slot order, and the arcs renamed.  Returns false if the line has to go the sequential way.
__device__ __forceinline__ bool nw_parse_line_warp(const char* __restrict__ T, int len, int2* arc, int2* lab, int32_t* st, int32_t* hk, int32_t* hi, int lane,
                                                   int& n_out, int& m_out, int& hyb_out) {
  const unsigned FULL = 0xffffffffu, lt = (1u << lane) - 1u;
  enum { D_SEMI, D_CL, D_OP, D_CM };
  int sp = 0, nslots = 1, narcs = 0, lastD = D_SEMI;
  if (lane == 0) hi[0] = len - 2;  // slot 0: the root's token, left of the ';'
  for (int r0 = 1; r0 < len; r0 += 32) {
    const int r = r0 + lane, p = len - 1 - r;   // lane order = reading order (right to left)
    const bool act = r < len;
    const char c = act ? T[p] : ' ';
    const bool isCL = c == ')', isOP = c == '(', isCM = c == ',';
    const unsigned cl = __ballot_sync(FULL, isCL), op = __ballot_sync(FULL, isOP), cm = __ballot_sync(FULL, isCM), dm = cl | op | cm;
    int rD = lastD;  // the delimiter met before this character
    if (dm & lt) { const unsigned b = 1u << (31 - __clz(dm & lt)); rD = (cl & b) ? D_CL : (op & b) ? D_OP : D_CM; }
    const int spb = sp + __popc(cl & lt) - __popc(op & lt);   // open groups when this character is met
    const int ls = nslots + __popc((cl | cm) & lt) - 1;       // the slot opened last
    const bool want = isOP || isCM;
    bool bad = false;
    if (isCL) bad = p == 0 || rD == D_OP;       // the reader never opens a group at the first character; ")" right after "(" is an error
    else if (want) bad = spb < 1;               // the tree is finished, what follows is ignored or an error
    else bad = act && rD == D_OP && !nw_ws(c);  // only blanks between "(" and the next "," or "("
    if (__any_sync(FULL, bad)) return false;
    int memP = 0, memC = 0;
    const bool wantC = want && rD == D_OP;      // the subtree just finished is the group that "(" closed: its owner sits at depth spb
    if (want) { memP = st[spb - 1]; if (wantC) memC = st[spb]; }
    __syncwarp();
    const unsigned gA = __match_any_sync(FULL, isCL ? spb : want ? spb - 1 : -1 - lane);
    const unsigned candA = gA & cl & lt;
    const int vA = __shfl_sync(FULL, ls, candA ? 31 - __clz(candA) : lane);
    const unsigned gB = __match_any_sync(FULL, isCL ? spb : wantC ? spb : -1 - lane);
    const unsigned candB = gB & cl & lt;
    const int vB = __shfl_sync(FULL, ls, candB ? 31 - __clz(candB) : lane);
    if (isCL && !(gA & cl & ~lt & ~(1u << lane))) st[spb] = ls;   // the last ")" of this depth in the chunk stays
    if (isCL || isCM) hi[ls + 1] = p - 1;                          // where the slot this delimiter opens begins
    if (want) arc[narcs + __popc((cm | op) & lt)] = make_int2(candA ? vA : memP, wantC ? (candB ? vB : memC) : ls);
    sp += __popc(cl) - __popc(op); nslots += __popc(cl | cm); narcs += __popc(cm | op);
    if (dm) { const unsigned b = 1u << (31 - __clz(dm)); lastD = (cl & b) ? D_CL : (op & b) ? D_OP : D_CM; }
    __syncwarp();
  }
  if (sp != 0) return false;
  // tokens: names and hybrid numbers, by slot
  int H = 0;
  for (int s0 = 0; s0 < nslots; s0 += 32) {
    const int s = s0 + lane;
    bool isH = false, tbad = false;
    int key = 0;
    int2 span = make_int2(0, 0);
    if (s < nslots) {
      int back = hi[s];
      if (s != 0) {  // ":<number>" right of the name (the first slot has none: io/newick.hpp reads the root's name straight after ';')
        int i_ = back;
        while (i_ >= 0) { const char c_ = T[i_]; if (nw_digit(c_) || c_ == '.' || c_ == '-' || c_ == '+' || c_ == 'E' || c_ == 'e') --i_; else break; }
        if (i_ >= 0 && T[i_] == ':') back = i_ - 1;
      }
      while (back >= 0 && nw_ws(T[back])) --back;
      int i = back, sharp = -1;
      while (i >= 0) { const char c = T[i]; if (c == '(' || c == ')' || c == ',') break; if (c == '#' && sharp < 0) sharp = i; --i; }
      const int nb = i + 1, nl = back - i;
      if (sharp >= 0) {
        int d = sharp;
        while (d < nb + nl && !nw_digit(T[d])) ++d;
        if (d == nb + nl) tbad = true;
        else {
          while (d < nb + nl && nw_digit(T[d])) { key = key * 10 + (T[d] - '0'); ++d; }
          isH = true; span = make_int2(nb, sharp - nb);
        }
      } else span = make_int2(nb, nl);
    }
    if (__any_sync(FULL, tbad)) return false;
    const unsigned hm = __ballot_sync(FULL, isH);
    __syncwarp();
    if (s < nslots) { lab[s] = span; hk[s] = key; hi[s] = s; }
    if (isH) st[H + __popc(hm & lt)] = s;
    H += __popc(hm);
  }
  __syncwarp();
  // a hybrid number met before names the node of its first slot: hi[s] = that slot
  for (int i0 = 0; i0 < H; i0 += 32) {
    const int i = i0 + lane;
    if (i < H) {
      const int s = st[i], key = hk[s];
      for (int j = 0; j < i; ++j) { const int sj = st[j]; if (hk[sj] == key) { hi[s] = sj; break; } }
    }
  }
  __syncwarp();
  // node ids = slots without the repeats, in slot order; names move to their node's place; hk[s] = node of slot s
  int ndup = 0;
  for (int s0 = 0; s0 < nslots; s0 += 32) {
    const int s = s0 + lane;
    const bool in = s < nslots;
    const int rep = in ? hi[s] : 0;
    const bool dup = in && rep != s;
    const int2 span = in ? lab[s] : make_int2(0, 0);
    const unsigned dmk = __ballot_sync(FULL, dup);
    const int id = s - ndup - __popc(dmk & lt);
    __syncwarp();
    if (in && !dup) { hk[s] = id; lab[id] = span; }
    __syncwarp();
    if (dup) hk[s] = hk[rep];
    ndup += __popc(dmk);
    __syncwarp();
  }
  for (int e = lane; e < narcs; e += 32) { const int2 a = arc[e]; arc[e] = make_int2(hk[a.x], hk[a.y]); }
  __syncwarp();
  n_out = nslots - ndup; m_out = narcs; hyb_out = H > 0;
  return true;
}

__global__ void __launch_bounds__(256) nw_parse_kernel(const char* __restrict__ text, int64_t num_lines, const int64_t* __restrict__ lbeg, const int32_t* __restrict__ llen,
                                NwScratch S, int32_t* __restrict__ ln, int32_t* __restrict__ lm, int32_t* __restrict__ lhyb, int32_t* __restrict__ lerr) {
  const int64_t k = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (k >= num_lines) return;
  const int64_t line0 = lbeg[k], reg = line0 + 2 * k;
  const int len = llen[k];
  const char* T = text + line0;
  int n = 0, m = 0, hyb = 0;
  bool ok = len >= 1 && T[len - 1] == ';';
  if (ok) ok = nw_parse_line_warp(T, len, S.arc + reg, S.lab + reg, S.st + reg, S.hk + reg, S.hi + reg, lane, n, m, hyb);
  if (ok) { if (lane == 0) { ln[k] = n; lm[k] = m; lhyb[k] = hyb; lerr[k] = NW_OK; } }
  else { __syncwarp(); if (lane == 0) nw_parse_line_seq(text, k, lbeg, llen, S, ln, lm, lhyb, lerr); }
}

// per pair: host / guest, label ids (written over the hybrid-key scratch of each node), sizes, maxima, first error
// ONE WARP per pair.  Label ids must come out in first-appearance order (host line, then guest line), so the nodes are taken in
// chunks of 32 in that order: inside a chunk every lane hashes and probes its own name, a slot keeps the SMALLEST node index of its
// label (claim by CAS, lower by atomicMin), and the lanes that ended up owning a slot number the new labels by lane order.
__device__ __forceinline__ void nw_label_chunks(const char* __restrict__ text, const int64_t* __restrict__ lbeg, NwScratch S, int64_t L, int n, int64_t reg_a, int64_t reg_b,
                                                int64_t la, int64_t lb, int32_t* slots, int cap, int& count, int lane) {
  const int64_t reg = lbeg[L] + 2 * L;
  const char* base = text + lbeg[L];
  for (int v0 = 0; v0 < n; v0 += 32) {
    const int v = v0 + lane;
    const int2 sp = v < n ? S.lab[reg + v] : make_int2(0, 0);
    const int me = (int)(reg + v - reg_a);
    const int chunk_end = (int)(reg - reg_a) + min(v0 + 32, n);  // nodes of THIS chunk of THIS line are [me - lane, chunk_end)
    int pos = -1;
    if (sp.y > 0) {
      unsigned h = 2166136261u;
      for (int k = 0; k < sp.y; ++k) h = (h ^ (unsigned char)base[sp.x + k]) * 16777619u;
      int i = (int)(h & (unsigned)(cap - 1));
      for (;;) {
        int sidx = __ldcg(&slots[i]);
        if (sidx < 0) { sidx = atomicCAS(&slots[i], -1, me); if (sidx < 0) { pos = i; break; } }
        const int64_t g = reg_a + sidx;
        const char* gb = text + (g >= reg_b ? lb : la);
        const int2 gs = S.lab[g];
        bool same = gs.y == sp.y;
        for (int k = 0; same && k < sp.y; ++k) same = gb[gs.x + k] == base[sp.x + k];
        if (same) { if (sidx > me && sidx < chunk_end) atomicMin(&slots[i], me); pos = i; break; }  // lower only inside this chunk: the host line may be the second one
        i = (i + 1) & (cap - 1);
      }
    }
    __syncwarp();
    const int rep = pos >= 0 ? __ldcg(&slots[pos]) : -1;
    const bool fresh = pos >= 0 && rep == me;
    const unsigned fm = __ballot_sync(0xffffffffu, fresh);
    int id = -1;
    if (fresh) { id = count + __popc(fm & ((1u << lane) - 1u)); S.hk[reg + v] = id; }
    __syncwarp();
    if (pos >= 0 && !fresh) id = S.hk[reg_a + rep];
    __syncwarp();
    if (v < n) S.hk[reg + v] = id;  // label id of node v (the hybrid keys are no longer needed)
    count += __popc(fm);
  }
  __syncwarp();
}

__global__ void __launch_bounds__(256) nw_pair_kernel(const char* __restrict__ text, int64_t num_pairs, const int64_t* __restrict__ lbeg, const int32_t* __restrict__ llen, NwScratch S,
                               const int32_t* __restrict__ ln, const int32_t* __restrict__ lm, const int32_t* __restrict__ lhyb, const int32_t* __restrict__ lerr,
                               int64_t* __restrict__ hn, int64_t* __restrict__ hm, int64_t* __restrict__ gn, int32_t* __restrict__ swapped,
                               int32_t* __restrict__ maxima /* n m nt labels */, unsigned long long* __restrict__ first_err /* pair << 8 | code */) {
  const int64_t p = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (p >= num_pairs) return;
  const int64_t a = 2 * p, b = 2 * p + 1;
  int err = lerr[a] ? lerr[a] : lerr[b];
  const bool t0 = lm[a] + 1 == ln[a], t1 = lm[b] + 1 == ln[b];
  const bool sw = t0 && !t1;
  const int64_t H = sw ? b : a, G = sw ? a : b;
  if (!err && (lm[G] + 1 != ln[G] || lhyb[G])) err = NW_GUEST;
  if (err) {
    if (lane == 0) { hn[p] = 0; hm[p] = 0; gn[p] = 0; swapped[p] = 0; atomicMin(first_err, ((unsigned long long)p << 8) | (unsigned)err); }
    return;
  }
  // label dictionary: open addressing over the pair's contiguous span of the `st` scratch (free after parsing)
  const int64_t reg_a = lbeg[a] + 2 * a, reg_b = lbeg[b] + 2 * b, span = reg_b + llen[b] + 2 - reg_a;
  int named = 0;
  for (int side = 0; side < 2; ++side) { const int64_t L = side ? G : H, reg = lbeg[L] + 2 * L; for (int v = lane; v < ln[L]; v += 32) named += S.lab[reg + v].y > 0; }
  for (int o = 16; o; o >>= 1) named += __shfl_xor_sync(0xffffffffu, named, o);
  int cap = 1; while ((int64_t)cap * 2 <= span && cap < 2 * named + 2) cap <<= 1;   // > named, and inside the pair's span (a name costs >= 2 characters)
  int32_t* slots = S.st + reg_a;  // entry = region index (relative to reg_a) of the first node with that label; -1 empty
  for (int i = lane; i < cap; i += 32) slots[i] = -1;
  __syncwarp();
  int count = 0;
  nw_label_chunks(text, lbeg, S, H, ln[H], reg_a, reg_b, lbeg[a], lbeg[b], slots, cap, count, lane);
  nw_label_chunks(text, lbeg, S, G, ln[G], reg_a, reg_b, lbeg[a], lbeg[b], slots, cap, count, lane);
  if (lane == 0) {
    hn[p] = ln[H]; hm[p] = lm[H]; gn[p] = ln[G]; swapped[p] = sw;
    atomicMax(&maxima[0], ln[H]); atomicMax(&maxima[1], lm[H]); atomicMax(&maxima[2], ln[G]); atomicMax(&maxima[3], count);
  }
}

// one warp: CSR of n nodes from m arcs (keyed by tail, or by head when BY_HEAD), neighbours in REVERSE arc order like the sequential
// `A[--cursor[key]] = other` over ascending arcs.  Arcs are taken 32 at a time in arc order, lanes sharing a key rank themselves with
// match.any, and the group's lowest lane moves the key's counter / cursor -- no atomics, the order is exact.  cur: n ints of scratch.
template <class OUT, bool BY_HEAD>
__device__ __forceinline__ void nw_warp_csr(const int2* __restrict__ arc, int m, int n, int32_t* cur, OUT* __restrict__ O, OUT* __restrict__ A, int lane) {
  const unsigned lt = (1u << lane) - 1u;
  for (int v = lane; v < n; v += 32) cur[v] = 0;
  __syncwarp();
  for (int e0 = 0; e0 < m; e0 += 32) {
    const int e = e0 + lane;
    const int t = e < m ? (BY_HEAD ? arc[e].y : arc[e].x) : -1 - lane;
    const unsigned grp = __match_any_sync(0xffffffffu, t);
    if (e < m && !(grp & lt)) cur[t] += __popc(grp);
    __syncwarp();
  }
  int carry = 0;
  for (int v0 = 0; v0 < n; v0 += 32) {
    const int v = v0 + lane;
    int x = v < n ? cur[v] : 0;
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (v < n) { O[v + 1] = (OUT)(carry + x); cur[v] = carry + x; }
    carry += __shfl_sync(0xffffffffu, x, 31);
  }
  if (lane == 0) O[0] = 0;
  __syncwarp();
  for (int e0 = 0; e0 < m; e0 += 32) {
    const int e = e0 + lane;
    int2 a2 = e < m ? arc[e] : make_int2(-1 - lane, 0);
    if (BY_HEAD && e < m) a2 = make_int2(a2.y, a2.x);
    const unsigned grp = __match_any_sync(0xffffffffu, a2.x);
    int c = 0;
    if (e < m) { c = cur[a2.x]; A[c - 1 - __popc(grp & lt)] = (OUT)a2.y; }
    __syncwarp();
    if (e < m && !(grp & lt)) cur[a2.x] = c - __popc(grp);
    __syncwarp();
  }
}

// ONE WARP per pair: the final arrays, written with coalesced stores.  OUT = int32_t or int16_t (pt_batch_desc.index_bytes).
// Children come out in reverse arc order (Phylogeny::build), see nw_warp_csr.
template <class OUT>
__global__ void __launch_bounds__(256) nw_emit_kernel(int64_t num_pairs, const int64_t* __restrict__ lbeg, NwScratch S, const int32_t* __restrict__ ln, const int32_t* __restrict__ lm,
                               const int32_t* __restrict__ swapped, const int64_t* __restrict__ hoff, const int64_t* __restrict__ aoff, const int64_t* __restrict__ goff,
                               OUT* __restrict__ succ_off, OUT* __restrict__ succ_adj, OUT* __restrict__ hlabel, OUT* __restrict__ gparent, OUT* __restrict__ glabel,
                               unsigned long long* __restrict__ first_err) {
  const int64_t p = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (p >= num_pairs) return;
  const int64_t a = 2 * p, b = 2 * p + 1, H = swapped[p] ? b : a, G = swapped[p] ? a : b;
  const int64_t rh = lbeg[H] + 2 * H, rg = lbeg[G] + 2 * G;
  const int n = ln[H], m = lm[H], nt = ln[G], mt = lm[G];
  OUT* O = succ_off + hoff[p] + p; OUT* A = succ_adj + aoff[p];
  const int2* arc = S.arc + rh;
  int32_t* cur = S.hi + rh;  // counters, then cursors (the hybrid ids are no longer needed)
  nw_warp_csr<OUT, false>(arc, m, n, cur, O, A, lane);
  bool dbl = false;
  for (int v = lane; v < n; v += 32) {
    const int x0 = v ? (int)O[v] : 0, x1 = (int)O[v + 1];
    for (int x = x0; x < x1; ++x) for (int y = x + 1; y < x1; ++y) dbl |= A[x] == A[y];
  }
  if (__any_sync(0xffffffffu, dbl) && lane == 0) atomicMin(first_err, ((unsigned long long)p << 8) | (unsigned)NW_DOUBLE);
  OUT* HL = hlabel + hoff[p];
  for (int v = lane; v < n; v += 32) HL[v] = (OUT)S.hk[rh + v];
  OUT* GP = gparent + goff[p]; OUT* GL = glabel + goff[p];
  for (int v = lane; v < nt; v += 32) { GP[v] = (OUT)-1; GL[v] = (OUT)S.hk[rg + v]; }
  __syncwarp();
  const int2* garc = S.arc + rg;
  for (int e = lane; e < mt; e += 32) { const int2 a2 = garc[e]; GP[a2.y] = (OUT)a2.x; }
}

// ---- pairs of NETWORKS for pt_iso_from_newick: no host / guest roles, both lines keep their order ------------------------------
// label ids shared by the two networks of a pair (first appearance: first line, then second), sizes per graph g = 2 * pair + side
__global__ void __launch_bounds__(256) nw_iso_pair_kernel(const char* __restrict__ text, int64_t num_pairs, const int64_t* __restrict__ lbeg, const int32_t* __restrict__ llen, NwScratch S,
                                   const int32_t* __restrict__ ln, const int32_t* __restrict__ lm, const int32_t* __restrict__ lerr,
                                   int64_t* __restrict__ gn, int64_t* __restrict__ gm, int32_t* __restrict__ max_nodes, unsigned long long* __restrict__ first_err) {
  const int64_t p = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (p >= num_pairs) return;
  const int64_t a = 2 * p, b = 2 * p + 1;
  if (lane == 0) { gn[a] = ln[a]; gn[b] = ln[b]; gm[a] = lm[a]; gm[b] = lm[b]; }
  const int err = lerr[a] ? lerr[a] : lerr[b];
  if (err) { if (lane == 0) atomicMin(first_err, ((unsigned long long)p << 8) | (unsigned)err); return; }
  const int64_t reg_a = lbeg[a] + 2 * a, reg_b = lbeg[b] + 2 * b, span = reg_b + llen[b] + 2 - reg_a;
  int named = 0;
  for (int side = 0; side < 2; ++side) { const int64_t L = side ? b : a, reg = lbeg[L] + 2 * L; for (int v = lane; v < ln[L]; v += 32) named += S.lab[reg + v].y > 0; }
  for (int o = 16; o; o >>= 1) named += __shfl_xor_sync(0xffffffffu, named, o);
  int cap = 1; while ((int64_t)cap * 2 <= span && cap < 2 * named + 2) cap <<= 1;
  int32_t* slots = S.st + reg_a;
  for (int i = lane; i < cap; i += 32) slots[i] = -1;
  __syncwarp();
  int count = 0;
  nw_label_chunks(text, lbeg, S, a, ln[a], reg_a, reg_b, lbeg[a], lbeg[b], slots, cap, count, lane);
  nw_label_chunks(text, lbeg, S, b, ln[b], reg_a, reg_b, lbeg[a], lbeg[b], slots, cap, count, lane);
  if (lane == 0) atomicMax(max_nodes, max(ln[a], ln[b]));
}
// one warp per GRAPH: successor and predecessor CSR (the order inside a list does not matter to pt_iso_batch) and label ids
__global__ void __launch_bounds__(256) nw_iso_emit_kernel(int64_t num_graphs, const int64_t* __restrict__ lbeg, NwScratch S, const int32_t* __restrict__ ln, const int32_t* __restrict__ lm,
                                   const int64_t* __restrict__ noff, const int64_t* __restrict__ aoff, int32_t* __restrict__ succ_off, int32_t* __restrict__ succ_adj,
                                   int32_t* __restrict__ pred_off, int32_t* __restrict__ pred_adj, int32_t* __restrict__ label) {
  const int64_t g = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (g >= num_graphs) return;
  const int64_t reg = lbeg[g] + 2 * g;
  const int n = ln[g], m = lm[g];
  const int2* arc = S.arc + reg;
  int32_t* cur = S.hi + reg;
  nw_warp_csr<int32_t, false>(arc, m, n, cur, succ_off + noff[g] + g, succ_adj + aoff[g], lane);
  nw_warp_csr<int32_t, true>(arc, m, n, cur, pred_off + noff[g] + g, pred_adj + aoff[g], lane);
  int32_t* Lb = label + noff[g];
  for (int v = lane; v < n; v += 32) Lb[v] = S.hk[reg + v];
}

__global__ void nw_set_first_kernel(int64_t* a, int64_t* b, int64_t* c) { a[0] = 0; b[0] = 0; c[0] = 0; }

}  // namespace ptgpu

namespace ptgpu {

// what both text entry points share: the device copy of the text, its non-blank lines, and every line parsed into scratch
struct NwFront {
  const char* d_text = nullptr;
  int64_t num_lines = 0, P = 0;
  int64_t* d_lbeg = nullptr; int32_t* d_llen = nullptr;
  NwScratch S{};
  int32_t *d_ln = nullptr, *d_lm = nullptr, *d_lh = nullptr, *d_le = nullptr;
};
// every allocation goes to `tmp` (the caller frees it, also on failure)
static int nw_front(const char* text, int64_t len, pt_mem mem, cudaStream_t s, std::vector<void*>& tmp, NwFront& F) {
  auto alloc = [&](void** p, size_t bytes, bool) -> int { const int rc = dev_alloc(p, std::max<size_t>(bytes, 16)); if (rc == PT_OK) tmp.push_back(*p); return rc; };
#define NW_TRY(expr) do { const int rc__ = (expr); if (rc__ != PT_OK) return rc__; } while (0)
#define NW_CUDA(call) do { const cudaError_t e__ = (call); if (e__ != cudaSuccess) return set_error(PT_ERR_CUDA, "newick reader: %s: %s", #call, cudaGetErrorString(e__)); } while (0)
  // own, 16-byte aligned copy of the text with 16 bytes of padding (the parser reads aligned 128-bit words)
  const char* d_text = nullptr;
  {
    void* p = nullptr;
    NW_TRY(alloc(&p, (size_t)len + 32, true));
    if (len) NW_CUDA(cudaMemcpyAsync(p, text, (size_t)len, mem == PT_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, s));
    NW_CUDA(cudaMemsetAsync(static_cast<char*>(p) + len, 0, 32, s));
    d_text = (const char*)p;
  }
  // ---- 1. lines ----
  int64_t* d_nlpos = nullptr;
  int64_t num_nl = 0;
  if (len > 0) {
    const int64_t nblocks = (len + (int64_t)NL_THREADS * NL_BYTES - 1) / ((int64_t)NL_THREADS * NL_BYTES);
    int32_t *d_bcnt = nullptr, *d_boff = nullptr;
    NW_TRY(alloc((void**)&d_bcnt, (size_t)(nblocks + 1) * 4, true)); NW_TRY(alloc((void**)&d_boff, (size_t)(nblocks + 1) * 4, true));
    nw_newline_kernel<false><<<(unsigned)nblocks, NL_THREADS, 0, s>>>(d_text, len, d_bcnt, nullptr, nullptr);
    NW_CUDA(cudaMemsetAsync(d_bcnt + nblocks, 0, 4, s));
    size_t tb = 0;
    NW_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_bcnt, d_boff, nblocks + 1, s));
    void* d_tmp = nullptr;
    NW_TRY(alloc(&d_tmp, tb, true));
    NW_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_bcnt, d_boff, nblocks + 1, s));
    int32_t num_nl32 = 0;
    NW_CUDA(cudaMemcpyAsync(&num_nl32, d_boff + nblocks, 4, cudaMemcpyDeviceToHost, s));
    NW_CUDA(cudaStreamSynchronize(s));
    num_nl = num_nl32;
    NW_TRY(alloc((void**)&d_nlpos, (size_t)(num_nl + 1) * 8, true));
    nw_newline_kernel<true><<<(unsigned)nblocks, NL_THREADS, 0, s>>>(d_text, len, nullptr, d_boff, d_nlpos);
  } else {
    NW_TRY(alloc((void**)&d_nlpos, 16, true));
  }
  const int64_t num_raw = num_nl + 1;
  int64_t *d_rbeg = nullptr, *d_lbeg = nullptr; int32_t *d_rlen = nullptr, *d_flag = nullptr, *d_pos = nullptr, *d_llen = nullptr;
  NW_TRY(alloc((void**)&d_rbeg, (size_t)num_raw * 8, true)); NW_TRY(alloc((void**)&d_rlen, (size_t)num_raw * 4, true));
  NW_TRY(alloc((void**)&d_flag, (size_t)(num_raw + 1) * 4, true)); NW_TRY(alloc((void**)&d_pos, (size_t)(num_raw + 1) * 4, true));
  nw_trim_kernel<<<(unsigned)((num_raw + 255) / 256), 256, 0, s>>>(d_text, len, d_nlpos, num_nl, d_rbeg, d_rlen, d_flag);
  NW_CUDA(cudaMemsetAsync(d_flag + num_raw, 0, 4, s));
  {
    size_t tb = 0;
    NW_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_flag, d_pos, num_raw + 1, s));
    void* d_tmp = nullptr;
    NW_TRY(alloc(&d_tmp, tb, true));
    NW_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_flag, d_pos, num_raw + 1, s));
  }
  int32_t num_lines32 = 0;
  NW_CUDA(cudaMemcpyAsync(&num_lines32, d_pos + num_raw, 4, cudaMemcpyDeviceToHost, s));
  NW_CUDA(cudaStreamSynchronize(s));
  const int64_t num_lines = num_lines32, P = num_lines / 2;
  NW_TRY(alloc((void**)&d_lbeg, (size_t)(num_lines + 1) * 8, true)); NW_TRY(alloc((void**)&d_llen, (size_t)(num_lines + 1) * 4, true));
  nw_compact_kernel<<<(unsigned)((num_raw + 255) / 256), 256, 0, s>>>(num_raw, d_rbeg, d_rlen, d_flag, d_pos, d_lbeg, d_llen);
  // ---- 2. parse every line ----
  const size_t S_len = (size_t)len + 2 * (size_t)num_lines + 8;
  NwScratch S{};
  NW_TRY(alloc((void**)&S.arc, S_len * 8, true)); NW_TRY(alloc((void**)&S.lab, S_len * 8, true));
  int32_t** sarr[3] = {&S.st, &S.hk, &S.hi};
  for (auto pp : sarr) NW_TRY(alloc((void**)pp, S_len * 4, true));
  int32_t *d_ln = nullptr, *d_lm = nullptr, *d_lh = nullptr, *d_le = nullptr;
  for (int32_t** pp : {&d_ln, &d_lm, &d_lh, &d_le}) NW_TRY(alloc((void**)pp, (size_t)(num_lines + 1) * 4, true));
  if (2 * P > 0) nw_parse_kernel<<<(unsigned)((2 * P + 7) / 8), 256, 0, s>>>(d_text, 2 * P, d_lbeg, d_llen, S, d_ln, d_lm, d_lh, d_le);
  F.d_text = d_text; F.num_lines = num_lines; F.P = P; F.d_lbeg = d_lbeg; F.d_llen = d_llen; F.S = S; F.d_ln = d_ln; F.d_lm = d_lm; F.d_lh = d_lh; F.d_le = d_le;
  return PT_OK;
#undef NW_TRY
#undef NW_CUDA
}

}  // namespace ptgpu

using namespace ptgpu;

extern "C" int pt_batch_create_from_newick(pt_batch** out, const char* text, int64_t len, pt_mem mem, int64_t* num_pairs, void* stream) {
  if (!out || (!text && len > 0) || len < 0) return set_error(PT_ERR_INVALID, "pt_batch_create_from_newick: bad argument");
  *out = nullptr;
  if (num_pairs) *num_pairs = 0;
  if (int rc = require_device()) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  std::vector<void*> tmp;  // freed on return
  void* keep[8] = {};       // owned by the handle on success
  auto cleanup = [&](bool ok) { cudaStreamSynchronize(s); for (void* p : tmp) dev_free(p); if (!ok) for (void* p : keep) dev_free(p); };
  auto alloc = [&](void** p, size_t bytes, bool temp) -> int { const int rc = dev_alloc(p, std::max<size_t>(bytes, 16)); if (rc == PT_OK && temp) tmp.push_back(*p); return rc; };
#define NW_TRY(expr) do { const int rc__ = (expr); if (rc__ != PT_OK) { cleanup(false); return rc__; } } while (0)
#define NW_CUDA(call) do { const cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup(false); return set_error(PT_ERR_CUDA, "pt_batch_create_from_newick: %s: %s", #call, cudaGetErrorString(e__)); } } while (0)
  NwFront F;
  NW_TRY(nw_front(text, len, mem, s, tmp, F));
  const char* d_text = F.d_text; const int64_t P = F.P; int64_t* d_lbeg = F.d_lbeg; int32_t* d_llen = F.d_llen; NwScratch S = F.S;
  int32_t *d_ln = F.d_ln, *d_lm = F.d_lm, *d_lh = F.d_lh, *d_le = F.d_le;
  // ---- 3. pairs ----
  int64_t *d_hn = nullptr, *d_hm = nullptr, *d_gn = nullptr; int32_t *d_sw = nullptr, *d_max = nullptr; unsigned long long* d_err = nullptr;
  for (int k = 0; k < 3; ++k) NW_TRY(alloc(&keep[k], (size_t)(P + 1) * 8, false));
  NW_TRY(alloc((void**)&d_hn, (size_t)(P + 1) * 8, true)); NW_TRY(alloc((void**)&d_hm, (size_t)(P + 1) * 8, true)); NW_TRY(alloc((void**)&d_gn, (size_t)(P + 1) * 8, true));
  NW_TRY(alloc((void**)&d_sw, (size_t)(P + 1) * 4, true)); NW_TRY(alloc((void**)&d_max, 16, true)); NW_TRY(alloc((void**)&d_err, 8, true));
  NW_CUDA(cudaMemsetAsync(d_max, 0, 16, s)); NW_CUDA(cudaMemsetAsync(d_err, 0xff, 8, s));
  if (P > 0) nw_pair_kernel<<<(unsigned)((P + 7) / 8), 256, 0, s>>>(d_text, P, d_lbeg, d_llen, S, d_ln, d_lm, d_lh, d_le, d_hn, d_hm, d_gn, d_sw, d_max, d_err);
  // ---- 4. offset tables ----
  int64_t* off[3] = {(int64_t*)keep[0], (int64_t*)keep[1], (int64_t*)keep[2]};
  nw_set_first_kernel<<<1, 1, 0, s>>>(off[0], off[1], off[2]);
  if (P > 0) {
    size_t tb = 0;
    NW_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tb, d_hn, off[0] + 1, P, s));
    void* d_tmp = nullptr;
    NW_TRY(alloc(&d_tmp, tb, true));
    NW_CUDA(cub::DeviceScan::InclusiveSum(d_tmp, tb, d_hn, off[0] + 1, P, s));
    NW_CUDA(cub::DeviceScan::InclusiveSum(d_tmp, tb, d_hm, off[1] + 1, P, s));
    NW_CUDA(cub::DeviceScan::InclusiveSum(d_tmp, tb, d_gn, off[2] + 1, P, s));
  }
  int64_t totals[3] = {0, 0, 0}; int32_t maxima[4] = {0, 0, 0, 0}; unsigned long long first_err = ~0ull;
  for (int k = 0; k < 3; ++k) NW_CUDA(cudaMemcpyAsync(&totals[k], off[k] + P, 8, cudaMemcpyDeviceToHost, s));
  NW_CUDA(cudaMemcpyAsync(maxima, d_max, 16, cudaMemcpyDeviceToHost, s));
  NW_CUDA(cudaMemcpyAsync(&first_err, d_err, 8, cudaMemcpyDeviceToHost, s));
  NW_CUDA(cudaStreamSynchronize(s));
  auto report = [&](unsigned long long fe) {
    const int code = (int)(fe & 0xff);
    const long long pair = (long long)(fe >> 8);
    cleanup(false);
    return set_error(code == NW_GUEST ? PT_ERR_INVALID : PT_ERR_PARSE, "pair %lld: %s", pair, nw_error_text(code));
  };
  if (first_err != ~0ull) return report(first_err);
  // ---- 5. final arrays (int16 when every id fits) ----
  const bool idx16 = maxima[0] < 32768 && maxima[1] < 32768 && maxima[2] < 32768 && maxima[3] < 32768;
  const size_t esz = idx16 ? 2 : 4;
  const size_t counts[5] = {(size_t)(totals[0] + P), (size_t)totals[1], (size_t)totals[0], (size_t)totals[2], (size_t)totals[2]};
  for (int k = 0; k < 5; ++k) NW_TRY(alloc(&keep[3 + k], counts[k] * esz, false));
  if (P > 0) {
    if (idx16) nw_emit_kernel<int16_t><<<(unsigned)((P + 7) / 8), 256, 0, s>>>(P, d_lbeg, S, d_ln, d_lm, d_sw, off[0], off[1], off[2], (int16_t*)keep[3], (int16_t*)keep[4],
                                                                                   (int16_t*)keep[5], (int16_t*)keep[6], (int16_t*)keep[7], d_err);
    else nw_emit_kernel<int32_t><<<(unsigned)((P + 7) / 8), 256, 0, s>>>(P, d_lbeg, S, d_ln, d_lm, d_sw, off[0], off[1], off[2], (int32_t*)keep[3], (int32_t*)keep[4],
                                                                             (int32_t*)keep[5], (int32_t*)keep[6], (int32_t*)keep[7], d_err);
    NW_CUDA(cudaGetLastError());
    NW_CUDA(cudaMemcpyAsync(&first_err, d_err, 8, cudaMemcpyDeviceToHost, s));
    NW_CUDA(cudaStreamSynchronize(s));
    if (first_err != ~0ull) return report(first_err);
  }
  pt_batch* h = nullptr;
  NW_TRY(ptgpu_batch_adopt(&h, P, keep, idx16 ? 2 : 4, maxima, stream));
  cleanup(true);
  *out = h;
  if (num_pairs) *num_pairs = P;
  return PT_OK;
#undef NW_TRY
#undef NW_CUDA
}

// Isomorphism straight from text: every two non-blank lines are one pair of networks (the `iso` tool's two-line file, examples/iso.cpp),
// parsed on the GPU like pt_batch_create_from_newick and handed to pt_iso_batch without ever leaving the device.
extern "C" int pt_iso_from_newick(const char* text, int64_t len, pt_mem mem, int32_t flags, uint32_t* result_bits, uint8_t* status, pt_mem out_mem,
                                  int64_t* num_pairs, void* stream) {
  if ((!text && len > 0) || len < 0 || !result_bits) return set_error(PT_ERR_INVALID, "pt_iso_from_newick: bad argument");
  if (num_pairs) *num_pairs = 0;
  if (int rc = require_device()) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  std::vector<void*> tmp;
  auto cleanup = [&]() { cudaStreamSynchronize(s); for (void* p : tmp) dev_free(p); };
  auto alloc = [&](void** p, size_t bytes) -> int { const int rc = dev_alloc(p, std::max<size_t>(bytes, 16)); if (rc == PT_OK) tmp.push_back(*p); return rc; };
#define NW_TRY(expr) do { const int rc__ = (expr); if (rc__ != PT_OK) { cleanup(); return rc__; } } while (0)
#define NW_CUDA(call) do { const cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup(); return set_error(PT_ERR_CUDA, "pt_iso_from_newick: %s: %s", #call, cudaGetErrorString(e__)); } } while (0)
  NwFront F;
  NW_TRY(nw_front(text, len, mem, s, tmp, F));
  if (F.num_lines & 1) { cleanup(); return set_error(PT_ERR_PARSE, "pt_iso_from_newick: odd number of networks (%lld)", (long long)F.num_lines); }
  const int64_t P = F.P, G = 2 * P;
  int64_t *d_gn = nullptr, *d_gm = nullptr, *d_noff = nullptr, *d_aoff = nullptr; int32_t* d_max = nullptr; unsigned long long* d_err = nullptr;
  NW_TRY(alloc((void**)&d_gn, (size_t)(G + 1) * 8)); NW_TRY(alloc((void**)&d_gm, (size_t)(G + 1) * 8));
  NW_TRY(alloc((void**)&d_noff, (size_t)(G + 1) * 8)); NW_TRY(alloc((void**)&d_aoff, (size_t)(G + 1) * 8));
  NW_TRY(alloc((void**)&d_max, 16)); NW_TRY(alloc((void**)&d_err, 8));
  NW_CUDA(cudaMemsetAsync(d_max, 0, 16, s)); NW_CUDA(cudaMemsetAsync(d_err, 0xff, 8, s));
  if (P > 0) nw_iso_pair_kernel<<<(unsigned)((P + 7) / 8), 256, 0, s>>>(F.d_text, P, F.d_lbeg, F.d_llen, F.S, F.d_ln, F.d_lm, F.d_le, d_gn, d_gm, d_max, d_err);
  nw_set_first_kernel<<<1, 1, 0, s>>>(d_noff, d_aoff, d_noff);
  if (G > 0) {
    size_t tb = 0;
    NW_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tb, d_gn, d_noff + 1, G, s));
    void* d_tmp = nullptr;
    NW_TRY(alloc(&d_tmp, tb));
    NW_CUDA(cub::DeviceScan::InclusiveSum(d_tmp, tb, d_gn, d_noff + 1, G, s));
    NW_CUDA(cub::DeviceScan::InclusiveSum(d_tmp, tb, d_gm, d_aoff + 1, G, s));
  }
  int64_t totals[2] = {0, 0}; int32_t max_nodes = 0; unsigned long long first_err = ~0ull;
  NW_CUDA(cudaMemcpyAsync(&totals[0], d_noff + G, 8, cudaMemcpyDeviceToHost, s)); NW_CUDA(cudaMemcpyAsync(&totals[1], d_aoff + G, 8, cudaMemcpyDeviceToHost, s));
  NW_CUDA(cudaMemcpyAsync(&max_nodes, d_max, 4, cudaMemcpyDeviceToHost, s)); NW_CUDA(cudaMemcpyAsync(&first_err, d_err, 8, cudaMemcpyDeviceToHost, s));
  NW_CUDA(cudaStreamSynchronize(s));
  if (first_err != ~0ull) {
    cleanup();
    return set_error(PT_ERR_PARSE, "pair %lld: %s", (long long)(first_err >> 8), nw_error_text((int)(first_err & 0xff)));
  }
  int32_t *d_so = nullptr, *d_sa = nullptr, *d_po = nullptr, *d_pa = nullptr, *d_lab = nullptr;
  NW_TRY(alloc((void**)&d_so, (size_t)(totals[0] + G) * 4)); NW_TRY(alloc((void**)&d_po, (size_t)(totals[0] + G) * 4));
  NW_TRY(alloc((void**)&d_sa, (size_t)totals[1] * 4)); NW_TRY(alloc((void**)&d_pa, (size_t)totals[1] * 4)); NW_TRY(alloc((void**)&d_lab, (size_t)totals[0] * 4));
  if (G > 0) nw_iso_emit_kernel<<<(unsigned)((G + 7) / 8), 256, 0, s>>>(G, F.d_lbeg, F.S, F.d_ln, F.d_lm, d_noff, d_aoff, d_so, d_sa, d_po, d_pa, d_lab);
  NW_CUDA(cudaGetLastError());
  pt_iso_desc d{};
  d.num_pairs = P; d.mem = PT_MEM_DEVICE; d.node_off = d_noff; d.arc_off = d_aoff; d.succ_off = d_so; d.succ_adj = d_sa; d.pred_off = d_po; d.pred_adj = d_pa;
  d.label = d_lab; d.flags = flags; d.max_nodes = std::max(max_nodes, 1);
  const int rc = P > 0 ? pt_iso_batch(&d, result_bits, status, out_mem, stream) : PT_OK;
  cleanup();
  if (rc == PT_OK && num_pairs) *num_pairs = P;
  return rc;
#undef NW_TRY
#undef NW_CUDA
}
