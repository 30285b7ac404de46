"""ctypes binding of libptgpu.so (include/pt_gpu.h).  The library is the product; this file only declares it.
There is no Python or CPU fallback: if the shared library is missing the import fails loudly."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PT_GPU_LIB") or os.path.join(_HERE, "libptgpu.so")  # PT_GPU_LIB: profiling builds only

PT_OK, PT_ERR_INVALID, PT_ERR_CUDA, PT_ERR_NOMEM, PT_ERR_NO_DEVICE, PT_ERR_UNSUPPORTED, PT_ERR_PARSE = 0, -1, -2, -3, -4, -5, -6
PT_MEM_HOST, PT_MEM_DEVICE = 0, 1
INST_OK, INST_MULTILABEL, INST_BAD_GRAPH, INST_TOO_LARGE, INST_POOL_OVERFLOW, INST_ROUND_LIMIT = 0, 1, 2, 3, 4, 5

i32p, i64p, u32p, u64p, u8p = (C.POINTER(t) for t in (C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_uint8))


class BatchDesc(C.Structure):
    _fields_ = [("num_instances", C.c_int64), ("mem", C.c_int),
                ("host_node_off", C.c_void_p), ("host_arc_off", C.c_void_p), ("guest_node_off", C.c_void_p),
                ("host_succ_off", C.c_void_p), ("host_succ_adj", C.c_void_p), ("host_label", C.c_void_p),
                ("guest_parent", C.c_void_p), ("guest_label", C.c_void_p),
                ("host_pred_off", C.c_void_p), ("host_pred_adj", C.c_void_p), ("guest_child_off", C.c_void_p), ("guest_child_adj", C.c_void_p),
                ("max_host_nodes", C.c_int32), ("max_host_arcs", C.c_int32), ("max_guest_nodes", C.c_int32), ("max_labels", C.c_int32),
                ("index_bytes", C.c_int32)]


class IsoDesc(C.Structure):
    _fields_ = [("num_pairs", C.c_int64), ("mem", C.c_int), ("node_off", C.c_void_p), ("arc_off", C.c_void_p), ("succ_off", C.c_void_p),
                ("succ_adj", C.c_void_p), ("pred_off", C.c_void_p), ("pred_adj", C.c_void_p), ("label", C.c_void_p),
                ("flags", C.c_int32), ("max_nodes", C.c_int32)]


class Counters(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("instances", "displayed", "not_displayed", "errors", "resolved_by_rules", "branched", "work_items", "rounds",
                                          "cherry_reductions", "reticulated_cherries", "arcs_deleted", "rule_passes")]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class Options(C.Structure):
    _fields_ = [("pool_slots", C.c_int32), ("max_rounds", C.c_int32), ("path", C.c_int32), ("no_small_core", C.c_int32), ("reserved", C.c_int32 * 4)]


class Limits(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("max_host_nodes", "max_host_arcs", "max_guest_nodes", "max_labels")]


class LcaInfo(C.Structure):
    _fields_ = [("num_nodes", C.c_int64), ("num_levels", C.c_int64), ("num_blocks", C.c_int64), ("table_levels", C.c_int64),
                ("device_bytes", C.c_int64), ("build_ms", C.c_float)]


def load():
    if not os.path.exists(LIB_PATH):
        raise ImportError("phylo_tools_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(or `make -C phylo_tools_b200/csrc`). There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    sig = {
        "pt_last_error": (C.c_char_p, []),
        "pt_abi_version": (C.c_int, []),
        "pt_device_count": (C.c_int, []),
        "pt_set_device": (C.c_int, [C.c_int]),
        "pt_lca_build": (C.c_int, [C.POINTER(vp), vp, C.c_int64, C.c_int, vp]),
        "pt_lca_query": (C.c_int, [vp, vp, vp, vp, C.c_int64, C.c_int, vp]),
        "pt_lca_get_info": (C.c_int, [vp, C.POINTER(LcaInfo)]),
        "pt_lca_node_info": (C.c_int, [vp, vp, vp, C.c_int, vp]),
        "pt_lca_free": (C.c_int, [vp]),
        "pt_tree_who_displays": (C.c_int, [vp, vp, C.c_int64, vp, vp, C.c_int64, vp, C.c_int, vp]),
        "pt_tree_highest_displayed": (C.c_int, [vp, vp, C.c_int64, C.c_int32, vp, C.c_int, vp]),
        "pt_net_bridges": (C.c_int, [C.c_int64, C.c_int64, vp, vp, vp, vp, C.c_int, vp]),
        "pt_rmq_build": (C.c_int, [C.POINTER(vp), vp, C.c_int64, C.c_int, vp]),
        "pt_rmq_query": (C.c_int, [vp, vp, vp, vp, C.c_int64, C.c_int, vp]),
        "pt_rmq_free": (C.c_int, [vp]),
        "pt_get_limits": (C.c_int, [C.POINTER(Limits)]),
        "pt_batch_create": (C.c_int, [C.POINTER(vp), C.POINTER(BatchDesc), vp]),
        "pt_batch_contain": (C.c_int, [vp, C.POINTER(Options), vp, vp, C.c_int, C.POINTER(Counters), vp]),
        "pt_batch_contain_multi": (C.c_int, [C.POINTER(BatchDesc), C.c_int, C.POINTER(Options), vp, vp, C.POINTER(Counters)]),
        "pt_batch_create_from_newick": (C.c_int, [C.POINTER(vp), vp, C.c_int64, C.c_int, i64p, vp]),
        "pt_batch_download": (C.c_int, [vp, C.c_int, vp, C.c_int64, i64p, i32p, i32p]),
        "pt_batch_last_launches": (C.c_int64, [vp]),
        "pt_batch_round_ms": (C.c_int, [vp, C.c_int, C.POINTER(C.c_float)]),
        "pt_batch_launch_count": (C.c_int64, [vp]),
        "pt_batch_launch_ms": (C.c_int, [vp, C.c_int64, C.POINTER(C.c_float)]),
        "pt_batch_h2d_bytes": (C.c_int64, [vp]),
        "pt_batch_free": (C.c_int, [vp]),
        "pt_enum_switchings": (C.c_int, [C.c_int32, C.c_int32, vp, vp, vp, C.c_int32, vp, vp, C.c_uint64, C.c_uint64, C.c_int, u64p, u64p, u64p, vp]),
        "pt_enum_create": (C.c_int, [C.POINTER(vp), C.c_int32, C.c_int32, vp, vp, vp, C.c_int32, vp, vp, u64p, vp]),
        "pt_enum_run": (C.c_int, [vp, C.c_uint64, C.c_uint64, C.c_int, vp, C.c_uint64, u64p, u64p, C.POINTER(C.c_int), vp]),
        "pt_enum_free": (C.c_int, [vp]),
        "pt_parse_newick": (C.c_int, [C.c_char_p, i64p, i64p, i32p, i64p, i32p, C.c_char_p, i64p]),
        "pt_parse_edgelist": (C.c_int, [C.c_char_p, i64p, i64p, i32p, i64p, i32p, C.c_char_p, i64p]),
        "pt_packer_create": (C.c_int, [C.POINTER(vp)]),
        "pt_packer_add_newick": (C.c_int, [vp, C.c_char_p, C.c_char_p]),
        "pt_packer_add_newick_text": (C.c_int, [vp, C.c_char_p, C.c_int64, C.c_int, i64p]),
        "pt_packer_add_generated": (C.c_int, [vp, C.c_int64, C.c_int32, C.c_int32, C.c_uint64, C.c_int]),
        "pt_packer_desc": (C.c_int, [vp, C.POINTER(BatchDesc)]),
        "pt_packer_desc_compact": (C.c_int, [vp, C.POINTER(BatchDesc)]),
        "pt_packer_desc_bytes": (C.c_int, [vp, C.POINTER(BatchDesc)]),
        "pt_packer_get_newick": (C.c_int, [vp, C.c_int64, C.c_char_p, i64p, C.c_char_p, i64p]),
        "pt_packer_free": (C.c_int, [vp]),
        "pt_gen_random_tree": (C.c_int, [C.c_int64, C.c_uint64, i32p]),
        "pt_iso_batch": (C.c_int, [C.POINTER(IsoDesc), vp, vp, C.c_int, vp]),
        "pt_iso_from_newick": (C.c_int, [vp, C.c_int64, C.c_int, C.c_int32, vp, vp, C.c_int, i64p, vp]),
        "pt_who_build": (C.c_int, [C.POINTER(vp), vp, vp, C.c_int64, vp, vp, C.c_int64, C.c_int, vp]),
        "pt_who_total": (C.c_int, [vp, i64p]),
        "pt_who_lists": (C.c_int, [vp, vp, vp, C.c_int, vp]),
        "pt_who_highest_displayed": (C.c_int, [vp, C.c_int32, vp, C.c_int, vp]),
        "pt_who_free": (C.c_int, [vp]),
        "pt_net_unzip": (C.c_int, [C.c_int64, C.c_int64, vp, vp, vp, C.c_int32, vp, vp, vp, C.c_int64, i64p, C.c_int, vp]),
        "pt_net_tree_components": (C.c_int, [C.c_int64, C.c_int64, vp, vp, vp, vp, vp, C.c_int64, i64p, C.c_int, vp]),
        "pt_net_visible_component": (C.c_int, [C.c_int64, C.c_int64, vp, vp, vp, C.c_int64, vp, vp, C.c_int32, C.c_int32, C.c_int64, vp, vp, vp, vp, vp]),
        "pt_lca_query_adjacent": (C.c_int, [vp, vp, C.c_int64, vp, C.c_int, vp]),
        "pt_comm_peer_exchange": (C.c_int, [vp]),
        "pt_comm_unique_id": (C.c_int, [vp]),
        "pt_comm_init": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int, vp]),
        "pt_comm_rank": (C.c_int, [vp]),
        "pt_comm_size": (C.c_int, [vp]),
        "pt_comm_allreduce": (C.c_int, [vp, vp, C.c_int64, C.c_int, C.c_int, vp]),
        "pt_batch_contain_sharded": (C.c_int, [vp, vp, vp, C.c_int64, C.c_int64, C.c_int64, vp, vp, vp, vp]),
        "pt_comm_free": (C.c_int, [vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)  # AttributeError here = the library does not export what include/pt_gpu.h declares
        f.restype, f.argtypes = res, args
    return L


EXPORTED = ["pt_last_error", "pt_abi_version", "pt_device_count", "pt_set_device", "pt_lca_build", "pt_lca_query", "pt_lca_get_info",
            "pt_lca_node_info", "pt_lca_free", "pt_tree_who_displays", "pt_tree_highest_displayed", "pt_net_bridges", "pt_rmq_build", "pt_rmq_query", "pt_rmq_free", "pt_get_limits", "pt_batch_create",
            "pt_batch_contain", "pt_batch_contain_multi", "pt_batch_create_from_newick", "pt_batch_download", "pt_batch_last_launches", "pt_batch_round_ms", "pt_batch_launch_count", "pt_batch_launch_ms", "pt_batch_h2d_bytes", "pt_batch_free", "pt_enum_switchings", "pt_enum_create", "pt_enum_run", "pt_enum_free", "pt_parse_newick", "pt_parse_edgelist", "pt_packer_create",
            "pt_packer_add_newick", "pt_packer_add_newick_text", "pt_packer_add_generated", "pt_packer_desc", "pt_packer_desc_compact", "pt_packer_desc_bytes", "pt_packer_get_newick", "pt_packer_free", "pt_gen_random_tree", "pt_iso_batch", "pt_iso_from_newick",
            "pt_lca_query_adjacent", "pt_who_build", "pt_who_total", "pt_who_lists", "pt_who_highest_displayed", "pt_who_free", "pt_net_unzip", "pt_net_tree_components", "pt_net_visible_component", "pt_comm_unique_id", "pt_comm_init", "pt_comm_rank", "pt_comm_peer_exchange", "pt_comm_size", "pt_comm_allreduce", "pt_batch_contain_sharded", "pt_comm_free"]
