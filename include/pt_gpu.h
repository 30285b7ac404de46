/* pt_gpu.h -- C-ABI of libptgpu.so, the B200-native batched tree-containment + LCA engine.
 *
 * The reference (igel-kun/phylo_tools) is header-only C++17 with NO plugin / FFI layer; the
 * boundary a maintainer would bind is therefore its template API.  Every entry point below names the
 * reference interface it stands in for (file:line relative to the reference checkout).  The C++ facade in
 * phylo_tools_b200/include/pt/ keeps the reference's names on top of this ABI; INTEGRATION.md shows the
 * reference-side stubs.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ or torch types cross this boundary.
 *   - every function returns PT_OK (0) or a negative pt_status; pt_last_error() gives the text
 *     (thread-local).  The reference throws std::logic_error / MalformedNewick instead
 *     (utils/storage_adj_common.hpp:103-109, utils/label_matching.hpp:51-52, io/newick.hpp:15-26);
 *     the facade re-throws those types from the status codes.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All device work of a
 *     call is enqueued on it.  Calls with HOST output buffers synchronise the stream before returning;
 *     calls with DEVICE output buffers are asynchronous.
 *   - there is NO CPU fallback: without a usable CUDA device every compute entry point returns
 *     PT_ERR_NO_DEVICE.
 *   - handles are single-owner; distinct handles may be used from distinct host threads.
 */
#ifndef PT_GPU_H_
#define PT_GPU_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PT_GPU_ABI_VERSION 3

typedef enum pt_status {
  PT_OK = 0,
  PT_ERR_INVALID = -1,     /* malformed arguments / structurally invalid input                        */
  PT_ERR_CUDA = -2,        /* a CUDA runtime call failed (text in pt_last_error)                      */
  PT_ERR_NOMEM = -3,       /* host or device allocation failed                                        */
  PT_ERR_NO_DEVICE = -4,   /* no CUDA device: the engine has no CPU path                              */
  PT_ERR_UNSUPPORTED = -5, /* instance exceeds an engine limit (see pt_limits)                        */
  PT_ERR_PARSE = -6        /* MalformedNewick / MalformedEdgeVec (io/newick.hpp:15, io/edgelist.hpp:10) */
} pt_status;

typedef enum pt_mem { PT_MEM_HOST = 0, PT_MEM_DEVICE = 1 } pt_mem;

const char* pt_last_error(void);
int pt_abi_version(void);
/* number of visible CUDA devices (0 without a GPU; never fails) */
int pt_device_count(void);
/* select the device used by subsequent calls of this thread (one process per GPU: call once with LOCAL_RANK) */
int pt_set_device(int device);

/* ------------------------------------------------------------------------------------------------
 * LCA  -- replaces  Tree::LCA / naiveLCA            utils/tree.hpp:202-223
 *                   lca<node_t>::lca(root), query   rmq/lca.hpp:79-106
 * Input is the tree as a parent array (parent[root] < 0; exactly one root); answers are node ids and
 * therefore unique -- bit-exact against either reference implementation.
 * Internals: preorder positions instead of the 2n-1 Euler tour, 32-wide blocks, a sparse
 * table over packed (depth,parent) block minima (rmq/sparse_rmq.hpp:45-90 re-designed), one 32-byte record
 * per node so a query is two random sector reads plus two table reads that stay in L2.
 * ---------------------------------------------------------------------------------------------- */
typedef struct pt_lca pt_lca;

typedef struct pt_lca_info {
  int64_t num_nodes;
  int64_t num_levels;      /* depth of the tree + 1 (number of wavefront steps of the build)           */
  int64_t num_blocks;      /* 32-position blocks                                                       */
  int64_t table_levels;    /* levels of the block sparse table                                         */
  int64_t device_bytes;    /* bytes kept resident for queries                                          */
  float build_ms;          /* device time of the build (CUDA events)                                   */
} pt_lca_info;

int pt_lca_build(pt_lca** out, const int32_t* parent, int64_t n, pt_mem parent_mem, void* stream);
int pt_lca_query(const pt_lca* h, const int32_t* u, const int32_t* v, int32_t* out, int64_t q,
                 pt_mem mem, void* stream);
/* out[i] = lca(nodes[i], nodes[i+1]) for i < count - 1: the induced-subtree step of the reference, which takes the LCAs of
 * CONSECUTIVE candidates of a preorder-sorted list (compute_inner_nodes, utils/induced_tree.hpp:128-138).  Any list is accepted;
 * each node's record is read once (the right neighbour's arrives by a warp shuffle), so a list sorted by preorder on a tree whose
 * ids follow a preorder -- what the reference's own readers produce (io/newick.hpp:164, containment.hpp:292) -- streams. */
int pt_lca_query_adjacent(const pt_lca* h, const int32_t* nodes, int64_t count, int32_t* out, pt_mem mem, void* stream);
int pt_lca_get_info(const pt_lca* h, pt_lca_info* info);
/* depth[v] and preorder position of every node (DEVICE or HOST output, n entries each; either may be NULL);
 * the order_number / dist_to_root of InducedSubtreeInfo  utils/induced_tree.hpp:11-35 */
int pt_lca_node_info(const pt_lca* h, int32_t* depth, int32_t* preorder, pt_mem mem, void* stream);
int pt_lca_free(pt_lca* h);

/* ------------------------------------------------------------------------------------------------
 * who_displays -- replaces  TreeInTreeContainment<Host,Guest>::who_displays(u)   utils/containment.hpp:72-177
 * for a SINGLE-LABELLED TREE host: out_host_node[u] = the host node that minimally displays the guest subtree below u
 * (the one entry of the reference's list), or -1 if none does.  The guest is displayed iff out_host_node[guest root] >= 0
 * (containment.hpp:83).  Internally the induced-subtree step of the reference (k-1 LCAs of consecutive candidates,
 * utils/induced_tree.hpp:128-138) runs as a dataflow over the guest tree (leaves first, a node as soon as its last child is done)
 * on top of pt_lca.  Any number of children per guest node (more than 64: sorted in device scratch by the one thread that owns the node).
 * Labels: dense ids < n + nt shared by both trees, -1 = unlabelled; only labels on leaves count.  Multi-labelled hosts
 * (the reference's MUL-trees) return PT_ERR_INVALID: use pt_who_build.
 * ---------------------------------------------------------------------------------------------- */
int pt_tree_who_displays(const int32_t* host_parent, const int32_t* host_label, int64_t n, const int32_t* guest_parent,
                         const int32_t* guest_label, int64_t nt, int32_t* out_host_node, pt_mem mem, void* stream);
/* replaces  TreeInComponent::highest_displayed_ancestor(v)   utils/containment.hpp:328-344, for ALL guest nodes at once, on top of
 * the table pt_tree_who_displays returned (`who`, -1 = not displayed): out[v] = the highest ancestor-or-self of v reached by the
 * reference's climb -- stay at v if its parent is not displayed; otherwise move to the parent and stop there if it is the guest
 * root or is displayed by `host_root` (the root of the host tree / tree component), else keep climbing.  out[root] = root (the
 * reference never asks).  guest_parent[root] = -1.  Pointer jumping, O(nt log depth). */
int pt_tree_highest_displayed(const int32_t* guest_parent, const int32_t* who, int64_t nt, int32_t host_root, int32_t* out, pt_mem mem, void* stream);

/* ------------------------------------------------------------------------------------------------
 * The tree-in-tree DP in its general form -- replaces
 *     TreeInTreeContainment<Host, Guest>::who_displays(u), displayed()    utils/containment.hpp:27-233
 *       (merge_child_poss :180-230, the induced subtree utils/induced_tree.hpp:65-168, the child-set registration :140-149,
 *        perfect_child_matching :174-177 = bipartite_matching utils/matching.hpp:8-93)
 *     TreeInComponent::unzip_retis / highest_displayed_ancestor            utils/containment.hpp:262-312, 328-344
 *     TreeComponentInfos (comp_root, visible_leaf, component DAG)           utils/tree_components.hpp:41-234
 * The host tree may be MULTI-LABELLED (the reference's ROMulTree: a leaf label may occur several times); the guest is a
 * single-labelled tree.  pt_who_build runs the whole DP (a dataflow over the guest tree on top of pt_lca) and keeps, per
 * guest node u, the list of host nodes that MINIMALLY display the guest subtree below u, ascending in host preorder -- what
 * who_displays(u) returns.  The guest is displayed iff the list of its root is not empty (:83).  A guest node whose
 * children all have ONE candidate (always, over a single-labelled host) may have any number of children; one whose children have several
 * candidates each is matched over ceil(d / 64)-word child sets by one thread and is limited to 1024 children (more: PT_ERR_UNSUPPORTED).
 * Labels: dense ids < n + nt shared by both trees, -1 = none; only leaf labels count.
 *   pt_who_total : number of entries over all lists;  pt_who_lists : off[nt + 1] and the concatenated lists (HOST or DEVICE)
 *   pt_who_highest_displayed : highest_displayed_ancestor of EVERY guest node over those lists ("displayed by the host root
 *       alone" = the list is exactly { host_root }, :340); out[guest root] = guest root
 *   pt_net_unzip : the multi-labelled tree the reference unfolds below node `root` of a network (edge traversal without a
 *       seen-set, arcs leaving out-degree-1 nodes skipped, out-degree-1 heads replaced by the end of their chain): node 0 = root,
 *       ids in preorder with the children in CSR order; out_origin[i] = the network node copy i stands for; out_label = the
 *       label of out-degree-0 origins.  *count receives the size; if it exceeds `capacity` (or the outputs are NULL) nothing
 *       is written and PT_ERR_UNSUPPORTED (PT_OK for NULL outputs) is returned: call again with room for *count nodes.
 *   pt_net_tree_components : comp_root[v] (-1 = NoNode) and visible_leaf[v] (for component roots: a leaf whose comp_root is v,
 *       the largest such id; else -1) as TreeComponentInfos computes them for a fresh network, and the arcs of the component
 *       DAG as (root above << 32 | root below), duplicates included (*dag_count of them; at most dag_capacity are written).
 *   pt_net_visible_component : VisibleComponentRule (utils/containment.hpp:603-725) as a query on a label-cleaned pair (HOST arrays):
 *       eligibility[v] (optional, n bytes) = bit 0: component root v is "1/2-eligible" (is_half_eligible :652-687: every root below it in
 *       the component DAG is, and the host has at most ONE path from v's tree component to each of them), bit 1: and a leaf sees it
 *       (eligible), bit 2: and no eligible root lies below it (what get_eligible_component_root :689-704 can return).  *root = the root
 *       the rule treats (smallest id with bit 2; want_root >= 0 overrides; -1 = the rule does not apply), *leaf = its visible leaf
 *       (want_leaf >= 0 overrides), *guest_node = highest_displayed_ancestor of the guest leaf carrying that label in
 *       TreeInComponent(host, guest, root) (treat_comp_root :623-644): the pair (root, guest_node) is what the reference hands to
 *       match_nodes.  Path counts are taken once per component-DAG arc; the reference's own counters depend on the iteration order
 *       of its hash containers (every root it calls half-eligible is half-eligible here).  unzip_cap bounds the unfolding (0 = 2^26).
 * ---------------------------------------------------------------------------------------------- */
typedef struct pt_who pt_who;
int pt_who_build(pt_who** out, const int32_t* host_parent, const int32_t* host_label, int64_t n, const int32_t* guest_parent,
                 const int32_t* guest_label, int64_t nt, pt_mem mem, void* stream);
int pt_who_total(const pt_who* h, int64_t* total_entries);
int pt_who_lists(const pt_who* h, int32_t* out_off, int32_t* out_adj, pt_mem mem, void* stream);
int pt_who_highest_displayed(const pt_who* h, int32_t host_root, int32_t* out, pt_mem mem, void* stream);
int pt_who_free(pt_who* h);
int pt_net_unzip(int64_t n, int64_t m, const int32_t* succ_off, const int32_t* succ_adj, const int32_t* label, int32_t root,
                 int32_t* out_parent, int32_t* out_label, int32_t* out_origin, int64_t capacity, int64_t* count, pt_mem mem, void* stream);
int pt_net_tree_components(int64_t n, int64_t m, const int32_t* succ_off, const int32_t* succ_adj, int32_t* comp_root, int32_t* visible_leaf,
                           int64_t* dag_edges, int64_t dag_capacity, int64_t* dag_count, pt_mem mem, void* stream);
int pt_net_visible_component(int64_t n, int64_t m, const int32_t* succ_off, const int32_t* succ_adj, const int32_t* host_label, int64_t nt,
                             const int32_t* guest_parent, const int32_t* guest_label, int32_t want_root, int32_t want_leaf, int64_t unzip_cap,
                             uint8_t* eligibility, int32_t* root, int32_t* leaf, int32_t* guest_node, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Network isomorphism -- replaces  make_iso_mapper(N0, N1, flags).check_isomorph()
 *                                  utils/isomorphism.hpp:33-324 (examples/iso.cpp:96-112, the `iso` tool)
 * for P pairs of networks at once.  Graph g = 2 * pair + side owns the slices given by node_off / arc_off
 * ([2P+1] prefix sums); succ_off / pred_off have n_g + 1 entries per graph starting at node_off[g] + g (LOCAL arc
 * positions), succ_adj / pred_adj m_g LOCAL node ids (children / parents), label a dense id shared by the two
 * networks of a pair or -1.  flags = which labels have to match (FLAG_MAP_* utils/isomorphism.hpp:10-13); the `iso`
 * tool's default is PT_ISO_LEAF_LABELS.  Result bit p = 1 iff the two networks of pair p are isomorphic (arc- and
 * label-preserving bijection): exact -- positives are verified arc by arc, negatives follow from invariants.
 * status[p]: PT_INST_OK, PT_INST_BAD_GRAPH (index out of range), PT_INST_TOO_LARGE (more than 4 096 nodes per
 * network), PT_INST_POOL_OVERFLOW (too many tied sub-problems).
 * ---------------------------------------------------------------------------------------------- */
#define PT_ISO_LEAF_LABELS 0x01
#define PT_ISO_TREE_LABELS 0x02
#define PT_ISO_RETI_LABELS 0x04
#define PT_ISO_ALL_LABELS 0x07
typedef struct pt_iso_desc {
  int64_t num_pairs;
  pt_mem mem;                /* where the arrays below live */
  const int64_t* node_off;   /* [2P+1] */
  const int64_t* arc_off;    /* [2P+1] */
  const int32_t* succ_off;
  const int32_t* succ_adj;
  const int32_t* pred_off;
  const int32_t* pred_adj;
  const int32_t* label;
  int32_t flags;
  int32_t max_nodes;         /* largest network of the batch; 0 = compute (HOST input only) */
} pt_iso_desc;
int pt_iso_batch(const pt_iso_desc* d, uint32_t* result_bits, uint8_t* status, pt_mem out_mem, void* stream);
/* the same from TEXT: every two non-blank lines of text[0, len) are one pair of networks in extended Newick (the two-line file of
 * examples/iso.cpp); parsed on the GPU (see pt_batch_create_from_newick) and solved there.  result_bits / status need room
 * for len / 4 pairs at most (a pair takes at least four characters); *num_pairs receives the number found.  A malformed
 * line or an odd number of lines: PT_ERR_PARSE (naming the first bad pair), nothing is written. */
int pt_iso_from_newick(const char* text, int64_t len, pt_mem mem, int32_t flags, uint32_t* result_bits, uint8_t* status, pt_mem out_mem,
                       int64_t* num_pairs, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Bridges -- replaces  BridgeFinder / list_bridges(N, root)   utils/bridges.hpp:42-128
 *            (Network::get_bridges utils/network.hpp:85-88; consumer: the bridge-separated components that
 *             utils/biconnected_comps.hpp:15-120 iterates)
 * Host network as the same CSR as pt_batch_desc (succ_off[n+1], succ_adj[m]).  is_bridge[e] = 1 iff removing arc e
 * (CSR position e) disconnects the underlying undirected graph; component[v] (optional) = id of the bridge-free
 * component of v, named by its topmost node (v's nearest ancestor-or-self that is the root or is entered by a bridge).
 * ---------------------------------------------------------------------------------------------- */
int pt_net_bridges(int64_t n, int64_t m, const int32_t* succ_off, const int32_t* succ_adj, uint8_t* is_bridge,
                   int32_t* component, pt_mem mem, void* stream);

/* ------------------------------------------------------------------------------------------------
 * RMQ -- replaces  rmq<It>::query / query_offset    rmq/rmq.hpp:63-74
 *                  sparse_rmq                        rmq/sparse_rmq.hpp:45-90
 * query over the half-open range [l, r), l < r; returns the index of the RIGHTMOST minimum, which is what
 * sparse_rmq returns (ties -> right operand, sparse_rmq.hpp:63,89).
 * ---------------------------------------------------------------------------------------------- */
typedef struct pt_rmq pt_rmq;
int pt_rmq_build(pt_rmq** out, const int32_t* values, int64_t n, pt_mem mem, void* stream);
int pt_rmq_query(const pt_rmq* h, const int64_t* l, const int64_t* r, int64_t* out_index, int64_t q,
                 pt_mem mem, void* stream);
int pt_rmq_free(pt_rmq* h);

/* ------------------------------------------------------------------------------------------------
 * Batched tree containment -- replaces
 *     TreeInNetContainment(host, guest).displayed()   utils/containment.hpp:1179-1182,1202
 *     TreeInTreeContainment(host, guest).displayed()  utils/containment.hpp:50-54,83
 *     the `tc` dispatch                               examples/tc.cpp:130-144
 * for B independent (host network, guest tree) instances at once.
 *
 * Flat layout (all int32 unless noted; instance i owns the slices given by the three offset tables).  With
 * index_bytes = 2 the five data arrays (succ_off, succ_adj, both label arrays, guest_parent) hold int16_t instead
 * (pass the pointers cast to const int32_t*): every instance then has fewer than 32 768 nodes, arcs and labels, and
 * HOST input costs half the PCIe traffic -- the end-to-end path is bound by exactly that copy.
 * With index_bytes = 1 they hold uint8_t, 255 standing for -1, and host_succ_off holds the n_i OUT-DEGREES of instance i
 * (sum(n_i) entries, instance i at host_node_off[i]) instead of n_i + 1 offsets: arc positions may exceed 255 (BASELINE
 * config 2 has m = 258) while every node id, label id and degree fits a byte -- 1 134 B per config-2 instance instead of
 * 2 294 B (int16) or 4 564 B (int32).  Needs at most 255 host nodes, guest nodes and labels per instance.
 *   host_node_off[B+1], host_arc_off[B+1], guest_node_off[B+1]   (int64 prefix sums of n_i, m_i, nt_i; first entry 0, last entry
 *                    = the length of the arrays below.  Every instance's range is checked against those lengths by the kernel
 *                    that loads it: a malformed entry makes the instances it touches PT_INST_BAD_GRAPH, nothing else)
 *   host_succ_off  : sum(n_i + 1) entries; instance i starts at host_node_off[i] + i; values are LOCAL arc
 *                    positions 0..m_i  (the immutable CSR of utils/storage_adj_immutable.hpp:173-261)
 *   host_succ_adj  : sum(m_i) entries at host_arc_off[i]; LOCAL node ids (children)
 *   host_label     : sum(n_i) entries at host_node_off[i]; per-instance dense label id >= 0, or -1 = unlabelled
 *                    ("" in the reference, utils/tree.hpp:138-148).  Host and guest of one instance share
 *                    the dictionary; ids must be < n_i + nt_i.
 *   guest_parent   : sum(nt_i) entries at guest_node_off[i]; LOCAL parent id, -1 for the root
 *   guest_label    : sum(nt_i) entries
 *   OPTIONAL (may be NULL; accepted for layout compatibility with SURVEY 8(b), not read by the kernels):
 *   host_pred_off / host_pred_adj (same shapes as succ), guest_child_off (sum(nt_i+1)), guest_child_adj (sum(nt_i-1))
 *
 * Semantics (the definition the reference intends, containment.hpp:955-957,1110-1125,1061):
 *   only labels on out-degree-0 nodes count; a label only in the guest => not displayed; a label only in
 *   the host => that leaf is deleted; <= 2 common labels => displayed; otherwise displayed iff some
 *   switching of the host (one in-arc kept per node), after removing unlabelled dangling nodes and
 *   suppressing in=out=1 nodes on both sides, equals the guest.
 * ---------------------------------------------------------------------------------------------- */
typedef struct pt_batch pt_batch;

typedef struct pt_batch_desc {
  int64_t num_instances;
  pt_mem mem;                      /* where ALL arrays below live                                     */
  const int64_t* host_node_off;
  const int64_t* host_arc_off;
  const int64_t* guest_node_off;
  const int32_t* host_succ_off;
  const int32_t* host_succ_adj;
  const int32_t* host_label;
  const int32_t* guest_parent;
  const int32_t* guest_label;
  const int32_t* host_pred_off;    /* optional */
  const int32_t* host_pred_adj;    /* optional */
  const int32_t* guest_child_off;  /* optional */
  const int32_t* guest_child_adj;  /* optional */
  /* per-batch maxima; 0 = let the library compute them (one extra pass + sync for DEVICE input) */
  int32_t max_host_nodes, max_host_arcs, max_guest_nodes, max_labels;
  int32_t index_bytes;             /* 0 or 4: data arrays are int32_t; 2: int16_t; 1: uint8_t with out-degrees (see above) */
} pt_batch_desc;

/* per-instance status written next to the verdict bit */
typedef enum pt_inst_status {
  PT_INST_OK = 0,
  PT_INST_MULTILABEL = 1,   /* duplicate leaf label: std::logic_error in label_matching.hpp:51-52,61-62 */
  PT_INST_BAD_GRAPH = 2,    /* index out of range, cycle, no/multiple roots (storage_adj_common.hpp:103-109) */
  PT_INST_TOO_LARGE = 3,    /* pt_options.path = 1 was forced and the instance exceeds pt_limits          */
  PT_INST_POOL_OVERFLOW = 4,/* branching pool exhausted (raise pt_options.pool_slots)                   */
  PT_INST_ROUND_LIMIT = 5   /* pt_options.max_rounds ended the branching before a verdict was reached   */
} pt_inst_status;

typedef struct pt_counters {
  uint64_t instances;
  uint64_t displayed;
  uint64_t not_displayed;
  uint64_t errors;               /* instances whose status != PT_INST_OK                               */
  uint64_t resolved_by_rules;    /* decided without branching                                          */
  uint64_t branched;             /* original instances that needed >= 1 branching                      */
  uint64_t work_items;           /* (sub)instances solved over all rounds                              */
  uint64_t rounds;               /* kernel launches (1 on the shared-memory executor: one persistent launch)  */
  uint64_t cherry_reductions;    /* true (multi-)cherries collapsed        (CherryPicker  :500-600)    */
  uint64_t reticulated_cherries; /* reticulation parents fixed                                         */
  uint64_t arcs_deleted;         /* arcs removed by the sibling-constraint rule                        */
  uint64_t rule_passes;          /* rule-loop iterations summed over work items (apply_rules :1054)    */
} pt_counters;

typedef struct pt_options {
  int32_t pool_slots;        /* capacity of the branching pool per round (0 = default)                  */
  int32_t max_rounds;        /* global-memory executors only: bound on the branching depth (0 = none: the depth is
                                bounded by the node count anyway); instances still pending then get PT_INST_ROUND_LIMIT */
  int32_t path;              /* 0 auto; 1 force the shared-memory executor (one warp per instance, int16 ids);
                                2 force the global-memory executor with one CTA per instance (int32 ids);
                                3 the same with a thread-block cluster per instance (what auto picks for the
                                  global-memory executor when a launch has fewer instances than half the SMs);
                                4 auto, never split (the default behaves like it);
                                5 split the instances along their bridges first (utils/bridges.hpp:42-118, the splitting of
                                  utils/biconnected_comps.hpp:33-88): tree parts are decided by cluster comparisons, every
                                  bridge-free component becomes a small instance of its own; batches that are not clean
                                  (unlabelled leaves, labels on one side only, unary guest nodes) fall back to auto.
                                  Pays for networks whose bridge-free components are small; generator networks have one
                                  giant component, there 3 is faster */
  int32_t no_small_core;     /* 0 (default): an instance on which the rules are stuck and of which at most 64 host nodes, 32 labels and
                                512 switchings are left is decided on the spot by enumerating those switchings, one per thread, with exact
                                32-bit cluster sets (what is left of a config-2 instance at that point has 9 ... 40 nodes); 1: always branch
                                through the pool (keeps the branching machinery testable; slower) */
  int32_t reserved[4];
} pt_options;

typedef struct pt_limits {
  int32_t max_host_nodes, max_host_arcs, max_guest_nodes, max_labels; /* shared-memory executor; beyond: global-memory executor */
} pt_limits;
int pt_get_limits(pt_limits* out);

/* HOST input is copied to the device here (the h2d leg of the end-to-end path); DEVICE input is referenced,
 * not copied, and must stay valid until pt_batch_free. */
int pt_batch_create(pt_batch** out, const pt_batch_desc* desc, void* stream);
/* result_bits: ceil(B/32) uint32 words, bit (i & 31) of word (i >> 5) = instance i displayed.
 * status: B bytes (pt_inst_status), may be NULL.  counters may be NULL (always HOST memory).
 * out_mem says where result_bits/status live. */
int pt_batch_contain(pt_batch* b, const pt_options* opt, uint32_t* result_bits, uint8_t* status,
                     pt_mem out_mem, pt_counters* counters, void* stream);
/* number of kernel launches issued by the last pt_batch_contain on this handle (for bench.py gpu_launches) */
int64_t pt_batch_last_launches(const pt_batch* b);
/* device time (CUDA events on `stream`) of the reduce-kernel launch of branching round `round` of the last
 * pt_batch_contain (round 0 = the launch over the original instances: the dominant kernel); PT_ERR_INVALID if that
 * round did not run */
int pt_batch_round_ms(const pt_batch* b, int round, float* ms);
/* Device time of the persistent launch of call number `launch` (0-based, counted per handle by pt_batch_launch_count) for the last 64
 * calls: a caller that queues many pt_batch_contain steps without waiting in between reads every step's kernel time afterwards. */
int64_t pt_batch_launch_count(const pt_batch* b);
int pt_batch_launch_ms(const pt_batch* b, int64_t launch, float* ms);
/* bytes copied host->device by pt_batch_create (0 for DEVICE input) */
int64_t pt_batch_h2d_bytes(const pt_batch* b);
/* Waits for what THIS handle has enqueued (an event behind its last call), not for everything the caller has queued on the stream
 * since: a caller may keep the next batch's pt_batch_create / pt_batch_contain queued behind this one (double buffering: the next
 * upload overlaps this solve) and free this one as soon as its own result has arrived. */
int pt_batch_free(pt_batch* b);

/* Extended-Newick text straight to a device-resident batch, PARSED ON THE GPU: every two non-blank lines of text[0, len) are
 * one instance (the network is the host, or the first line if both are trees, examples/tc.cpp:110-111).  Same node numbering,
 * child order and label ids as parse_newick (io/newick.hpp:274-290) + the packer, i.e. the arrays are bit-identical to
 * pt_packer_add_newick_text on the same text -- but no host copy of them ever exists: the text is uploaded (mem = where it
 * lives), one warp parses each line, one warp per pair writes the CSR (int16 layout whenever every id fits).  On malformed
 * input nothing is created and PT_ERR_PARSE / PT_ERR_INVALID names the first bad pair.  Use with pt_batch_contain / _free. */
int pt_batch_create_from_newick(pt_batch** out, const char* text, int64_t len, pt_mem mem, int64_t* num_pairs, void* stream);
/* copy one of the batch's device arrays to the host (tests, debugging): which = 0 host_node_off, 1 host_arc_off,
 * 2 guest_node_off (int64), 3 host_succ_off, 4 host_succ_adj, 5 host_label, 6 guest_parent, 7 guest_label (*index_bytes wide);
 * dst == NULL only queries *bytes; maxima = {max_host_nodes, max_host_arcs, max_guest_nodes, max_labels} */
int pt_batch_download(pt_batch* b, int which, void* dst, int64_t capacity_bytes, int64_t* bytes, int32_t* index_bytes, int32_t maxima[4]);

/* The whole node in one call, without torch: the instances of a HOST batch are split into contiguous ranges (boundaries
 * multiples of 32), one per GPU (num_devices <= 0: all visible ones; more shards than GPUs share them round-robin), each solved by pt_batch_create / pt_batch_contain on its
 * own host thread; result bits, status and counters land in the caller's buffers as for pt_batch_contain with PT_MEM_HOST
 * (counters.rounds = the largest branching depth of any shard).  SURVEY 8(e): the path shards with no collective. */
int pt_batch_contain_multi(const pt_batch_desc* d, int num_devices, const pt_options* opt, uint32_t* result_bits, uint8_t* status,
                           pt_counters* counters);

/* ------------------------------------------------------------------------------------------------
 * One process per GPU: the instances of a batch are cut into contiguous ranges whose boundaries are multiples of
 * 32, so that every rank owns whole words of the packed result vector; the path's ONLY collective is one all-reduce(SUM) of the
 * result words of the whole batch (disjoint words: SUM == OR) plus one of the counters -- issued here, inside the library,
 * through NCCL (dlopen of libnccl.so.2: the NCCL the host process already uses, e.g. torch's).  The reference has no
 * counterpart (single-threaded, no communication backend: SURVEY 5).
 *   pt_comm_unique_id : rank 0 makes the 128-byte id and hands it to the other ranks by the host program's own bootstrap
 *   pt_comm_init      : every rank, same id; nranks == 1 needs no id and no NCCL
 *   pt_batch_contain_sharded : `b` = this rank's range [first_instance, first_instance + local_instances) of a batch of
 *       total_instances (first_instance a multiple of 32).  full_bits: DEVICE, ceil(total / 32) words; full_status: DEVICE,
 *       total bytes padded to a multiple of 4, or NULL; both hold the WHOLE batch on every rank afterwards.  The local kernel
 *       writes straight into this rank's words.  total (HOST, optional): counters summed over the ranks (rounds: the maximum);
 *       the call synchronises the stream only when it is given.
 *   pt_comm_allreduce : in-place all-reduce of `count` DEVICE words of elem_bytes 4 or 8, op 0 = sum, 1 = max (the enumerator's
 *       found word)
 * ---------------------------------------------------------------------------------------------- */
typedef struct pt_comm pt_comm;
int pt_comm_unique_id(void* id128);
int pt_comm_init(pt_comm** out, int nranks, int rank, const void* id128);
int pt_comm_rank(const pt_comm* c);
/* 1 if the ranks exchange result words through each other's memory (CUDA IPC mappings over NVLink peer access, set up and agreed on
 * by all ranks in pt_comm_init), 0 if through NCCL (no peer access, more than 64 ranks, or PT_COMM_P2P=0 in the environment) */
int pt_comm_peer_exchange(const pt_comm* c);
int pt_comm_size(const pt_comm* c);
int pt_comm_allreduce(pt_comm* c, void* device_words, int64_t count, int elem_bytes, int op, void* stream);
int pt_batch_contain_sharded(pt_batch* b, pt_comm* c, const pt_options* opt, int64_t first_instance, int64_t total_instances, int64_t local_instances,
                             uint32_t* full_bits, uint8_t* full_status, pt_counters* total, void* stream);
int pt_comm_free(pt_comm* c);

/* ------------------------------------------------------------------------------------------------
 * Exhaustive switching enumeration (one instance, a range of the mixed-radix switching index).
 * The reference has no counterpart: it branches on one reticulation at a time (containment.hpp:1225-1266);
 * the verdict semantics are identical (exists a switching).  Used for general networks that do not reduce
 * and to shard [0, prod in-degree) across GPUs.
 * Instance arrays as in pt_batch_desc but for ONE instance (HOST memory).
 * n_switchings receives the size of the full index space; n_displaying the number of indices in
 * [begin, end) whose switching displays the guest; first_index the smallest such index or UINT64_MAX.
 * end == 0 means "to the end".
 * HOW equality is decided: the kernel compares a canonical 64-bit Merkle hash of the display tree with the guest's (collision
 * probability about 2^-64 per comparison), so n_displaying is exact up to that probability; first_index -- the WITNESS -- is
 * confirmed exactly before it is returned (the switching is applied to the host and the exact solver of pt_batch_contain decides;
 * a collision would be skipped and counted out).
 * pt_enum_create / pt_enum_run / pt_enum_free: the same with a persistent handle (host preprocessing, upload and accumulators once;
 * a run is one kernel launch per chunk of the range).  With a pt_comm and stop_at_first the found word is all-reduced (MAX)
 * across the ranks after every chunk of `chunk` switchings (0 = 2^20): every rank stops as soon as any rank has a witness
 * (SURVEY 8(e) row 2); *stopped_early reports it.  Every rank of the communicator must call pt_enum_run together.
 * ---------------------------------------------------------------------------------------------- */
int pt_enum_switchings(int32_t n, int32_t m, const int32_t* succ_off, const int32_t* succ_adj, const int32_t* host_label,
                       int32_t nt, const int32_t* guest_parent, const int32_t* guest_label,
                       uint64_t begin, uint64_t end, int stop_at_first,
                       uint64_t* n_switchings, uint64_t* n_displaying, uint64_t* first_index, void* stream);
typedef struct pt_enum pt_enum;
int pt_enum_create(pt_enum** out, int32_t n, int32_t m, const int32_t* succ_off, const int32_t* succ_adj, const int32_t* host_label, int32_t nt,
                   const int32_t* guest_parent, const int32_t* guest_label, uint64_t* n_switchings, void* stream);
int pt_enum_run(pt_enum* h, uint64_t begin, uint64_t end, int stop_at_first, pt_comm* comm, uint64_t chunk, uint64_t* n_displaying,
                uint64_t* first_index, int* stopped_early, void* stream);
int pt_enum_free(pt_enum* h);

/* ------------------------------------------------------------------------------------------------
 * Host-side readers / generators (no GPU needed).  C wrappers over the C++ facade so that non-C++ callers
 * (the ctypes tests, bench.py) reach the same code.
 * ---------------------------------------------------------------------------------------------- */
/* parse_newick  io/newick.hpp:274-290 : node ids = right-to-left preorder, root 0; edges in the
 * reference's (post-order) sequence.  Two-call protocol: call with edges == NULL to get the sizes.
 * labels: concatenated NUL-terminated strings, label_node[k] = node of the k-th string. */
int pt_parse_newick(const char* text, int64_t* num_nodes, int64_t* num_edges, int32_t* edges /* 2*num_edges */,
                    int64_t* num_labels, int32_t* label_node, char* label_buf, int64_t* label_buf_len);

/* parse_edgelist  io/edgelist.hpp:41-86 : white-space separated name pairs, one arc per line; node ids by first appearance; EVERY
 * node is labelled with its name.  Same two-call protocol and outputs as pt_parse_newick; MalformedEdgeVec -> PT_ERR_PARSE. */
int pt_parse_edgelist(const char* text, int64_t* num_nodes, int64_t* num_edges, int32_t* edges /* 2*num_edges */,
                      int64_t* num_labels, int32_t* label_node, char* label_buf, int64_t* label_buf_len);

/* Opaque packer: accumulates (host newick, guest newick) pairs into the flat layout of pt_batch_desc. */
typedef struct pt_packer pt_packer;
int pt_packer_create(pt_packer** out);
/* returns PT_ERR_PARSE / PT_ERR_INVALID for a malformed pair (nothing is added) */
int pt_packer_add_newick(pt_packer* p, const char* host_newick, const char* guest_newick);
/* batch reader: every two non-blank lines of text[0, len) are one instance (host line, guest line; the network is the host,
 * or the first line if both are trees, examples/tc.cpp:110-111).  Parsed on num_threads host threads (0 = all cores)
 * straight into the flat arrays; same node numbering and label ids as pt_packer_add_newick.  On a malformed pair nothing
 * is added and PT_ERR_PARSE / PT_ERR_INVALID names the pair. */
int pt_packer_add_newick_text(pt_packer* p, const char* text, int64_t len, int num_threads, int64_t* num_added);
/* adds `count` synthetic instances: the generate_random_binary_edgelist_trl process (utils/generator.hpp:175-245)
 * with a seeded PRNG (seed + instance index), guest = tree displayed by a uniform random switching with shuffled
 * child order; kind 0 = positive, 1 = guest with two leaf labels swapped, 2 = random guest tree on the same labels */
int pt_packer_add_generated(pt_packer* p, int64_t count, int32_t num_leaves, int32_t num_retis, uint64_t seed, int kind);
/* fills `desc` with pointers into the packer's HOST buffers (valid until the next add / free) */
int pt_packer_desc(pt_packer* p, pt_batch_desc* desc);
/* the same batch with int16 data arrays (index_bytes = 2); PT_ERR_INVALID if some instance is too large for that.
 * The arrays belong to the packer and stay valid until the next add_* / free. */
int pt_packer_desc_compact(pt_packer* p, pt_batch_desc* desc);
/* the same batch with uint8 data arrays (index_bytes = 1, host_succ_off = out-degrees); PT_ERR_INVALID if some instance has more
 * than 255 host nodes, guest nodes or labels.  Same ownership as pt_packer_desc_compact. */
int pt_packer_desc_bytes(pt_packer* p, pt_batch_desc* desc);
/* extended Newick text of instance i (host, then guest); two-call protocol on buffer lengths */
int pt_packer_get_newick(pt_packer* p, int64_t i, char* host_buf, int64_t* host_len, char* guest_buf, int64_t* guest_len);
int pt_packer_free(pt_packer* p);

/* random rooted binary tree with `num_leaves` leaves as a parent array of 2*num_leaves-1 entries
 * ("repeatedly join two uniformly chosen roots"), seeded */
int pt_gen_random_tree(int64_t num_leaves, uint64_t seed, int32_t* parent_out);

#ifdef __cplusplus
}
#endif
#endif /* PT_GPU_H_ */
