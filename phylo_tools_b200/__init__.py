"""phylo_tools_b200 -- B200-native batched tree containment + LCA behind the API surface of igel-kun/phylo_tools.

The product is the C-ABI shared library phylo_tools_b200/libptgpu.so (include/pt_gpu.h; CUDA kernels for sm_100a in
csrc/) plus the C++ facade in include/pt/.  This Python package is a thin ctypes layer used by the tests, bench.py and
Python callers; it contains no algorithmic code and no CPU fallback."""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import (PT_MEM_DEVICE, PT_MEM_HOST, BatchDesc, Counters, LcaInfo, Options)

_L = _capi.load()


class PtError(RuntimeError):
    def __init__(self, code, where):
        self.code = code
        super().__init__("%s failed (%d): %s" % (where, code, (_L.pt_last_error() or b"").decode(errors="replace")))


class MalformedNewick(ValueError):
    """io/newick.hpp:15-26"""


def _check(rc, where):
    if rc == _capi.PT_ERR_PARSE:
        raise MalformedNewick((_L.pt_last_error() or b"").decode(errors="replace"))
    if rc != 0:
        raise PtError(rc, where)


def lib():
    return _L


def device_count():
    return _L.pt_device_count()


def set_device(i):
    _check(_L.pt_set_device(int(i)), "pt_set_device")


def _ptr(x):
    """raw address of a numpy array / torch tensor / int"""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    return x.data_ptr()  # torch tensor


def _is_device(x):
    return hasattr(x, "is_cuda") and x.is_cuda


def parse_newick(text):
    """parse_newick (io/newick.hpp:274-290) -> (num_nodes, edges[(tail, head)], {node: label})"""
    s = text.encode()
    n, m, nl, bl = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64(0)
    _check(_L.pt_parse_newick(s, C.byref(n), C.byref(m), None, C.byref(nl), None, None, C.byref(bl)), "pt_parse_newick")
    edges = (C.c_int32 * max(1, 2 * m.value))()
    nodes = (C.c_int32 * max(1, nl.value))()
    buf = C.create_string_buffer(max(1, bl.value))
    _check(_L.pt_parse_newick(s, C.byref(n), C.byref(m), edges, C.byref(nl), nodes, buf, C.byref(bl)), "pt_parse_newick")
    names = buf.raw[:bl.value].split(b"\0")[:-1] if bl.value else []
    return n.value, [(edges[2 * i], edges[2 * i + 1]) for i in range(m.value)], {nodes[i]: names[i].decode() for i in range(nl.value)}


class MalformedEdgeVec(ValueError):
    """io/edgelist.hpp:10-15"""


def parse_edgelist(text):
    """parse_edgelist (io/edgelist.hpp:77-86) -> (num_nodes, edges[(tail, head)], {node: label}); every node carries its name"""
    s = text.encode()
    n, m, nl, bl = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64(0)
    rc = _L.pt_parse_edgelist(s, C.byref(n), C.byref(m), None, C.byref(nl), None, None, C.byref(bl))
    if rc == _capi.PT_ERR_PARSE:
        raise MalformedEdgeVec((_L.pt_last_error() or b"").decode(errors="replace"))
    _check(rc, "pt_parse_edgelist")
    edges = (C.c_int32 * max(1, 2 * m.value))()
    nodes = (C.c_int32 * max(1, nl.value))()
    buf = C.create_string_buffer(max(1, bl.value))
    _check(_L.pt_parse_edgelist(s, C.byref(n), C.byref(m), edges, C.byref(nl), nodes, buf, C.byref(bl)), "pt_parse_edgelist")
    names = buf.raw[:bl.value].split(b"\0")[:-1] if bl.value else []
    return n.value, [(edges[2 * i], edges[2 * i + 1]) for i in range(m.value)], {nodes[i]: names[i].decode() for i in range(nl.value)}


def random_tree(num_leaves, seed):
    """parent array (int32, 2*num_leaves-1 entries) of a random rooted binary tree"""
    out = np.empty(2 * num_leaves - 1, dtype=np.int32)
    _check(_L.pt_gen_random_tree(num_leaves, seed, out.ctypes.data_as(_capi.i32p)), "pt_gen_random_tree")
    return out


class Packer:
    """host-side batch builder (pt/packer.hpp): Newick pairs or generated instances -> flat arrays"""

    def __init__(self):
        self._h = C.c_void_p()
        _check(_L.pt_packer_create(C.byref(self._h)), "pt_packer_create")

    def add_newick(self, host, guest):
        _check(_L.pt_packer_add_newick(self._h, host.encode(), guest.encode()), "pt_packer_add_newick")

    def add_newick_text(self, text, threads=0):
        """batch reader (pt/fast_newick.hpp): two lines per instance, parsed in parallel; returns the number added"""
        data = text if isinstance(text, (bytes, bytearray)) else text.encode()
        k = C.c_int64(0)
        _check(_L.pt_packer_add_newick_text(self._h, bytes(data), len(data), threads, C.byref(k)), "pt_packer_add_newick_text")
        return k.value

    def add_generated(self, count, leaves, retis, seed, kind=0):
        _check(_L.pt_packer_add_generated(self._h, count, leaves, retis, seed, kind), "pt_packer_add_generated")

    def desc(self):
        d = BatchDesc()
        _check(_L.pt_packer_desc(self._h, C.byref(d)), "pt_packer_desc")
        return d

    def arrays(self, compact=False):
        """numpy views (copied) of the flat arrays, keyed like pt_batch_desc; compact=True or 2: int16 data arrays
        (index_bytes = 2); compact=1: uint8 data arrays with out-degrees in host_succ_off (index_bytes = 1)"""
        compact = 2 if compact is True else int(compact)
        if compact == 2:
            d = BatchDesc()
            _check(_L.pt_packer_desc_compact(self._h, C.byref(d)), "pt_packer_desc_compact")
        elif compact == 1:
            d = BatchDesc()
            _check(_L.pt_packer_desc_bytes(self._h, C.byref(d)), "pt_packer_desc_bytes")
        else:
            d = self.desc()
        B = d.num_instances
        it, nt = {2: (C.c_int16, np.int16), 1: (C.c_uint8, np.uint8)}.get(compact, (C.c_int32, np.int32))

        def view(ptr, count, ctype, dtype):
            if not ptr or count == 0:
                return np.zeros(0, dtype=dtype)
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(count,)).copy()
        hn = view(d.host_node_off, B + 1, C.c_int64, np.int64)
        ha = view(d.host_arc_off, B + 1, C.c_int64, np.int64)
        gn = view(d.guest_node_off, B + 1, C.c_int64, np.int64)
        NH, MH, NG = (int(hn[-1]), int(ha[-1]), int(gn[-1])) if B else (0, 0, 0)
        out = dict(num_instances=B, host_node_off=hn, host_arc_off=ha, guest_node_off=gn,
                   host_succ_off=view(d.host_succ_off, NH + (0 if compact == 1 else B), it, nt), host_succ_adj=view(d.host_succ_adj, MH, it, nt),
                   host_label=view(d.host_label, NH, it, nt), guest_parent=view(d.guest_parent, NG, it, nt),
                   guest_label=view(d.guest_label, NG, it, nt), index_bytes=compact if compact in (1, 2) else 4,
                   host_pred_off=view(d.host_pred_off, NH + B, C.c_int32, np.int32), host_pred_adj=view(d.host_pred_adj, MH, C.c_int32, np.int32),
                   guest_child_off=view(d.guest_child_off, NG + B, C.c_int32, np.int32), guest_child_adj=view(d.guest_child_adj, max(NG - B, 0), C.c_int32, np.int32),
                   max_host_nodes=d.max_host_nodes, max_host_arcs=d.max_host_arcs, max_guest_nodes=d.max_guest_nodes, max_labels=d.max_labels)
        return out

    def newick(self, i):
        hl, gl = C.c_int64(0), C.c_int64(0)
        _check(_L.pt_packer_get_newick(self._h, i, None, C.byref(hl), None, C.byref(gl)), "pt_packer_get_newick")
        hb, gb = C.create_string_buffer(hl.value), C.create_string_buffer(gl.value)
        _check(_L.pt_packer_get_newick(self._h, i, hb, C.byref(hl), gb, C.byref(gl)), "pt_packer_get_newick")
        return hb.value.decode(), gb.value.decode()

    def __len__(self):
        return int(self.desc().num_instances)

    def __del__(self):
        if getattr(self, "_h", None) and _L is not None:
            _L.pt_packer_free(self._h)
            self._h = None


_ARRAY_KEYS = ["host_node_off", "host_arc_off", "guest_node_off", "host_succ_off", "host_succ_adj", "host_label", "guest_parent", "guest_label"]
_OPT_KEYS = ["host_pred_off", "host_pred_adj", "guest_child_off", "guest_child_adj"]


def make_desc(arrays, mem):
    d = BatchDesc()
    d.num_instances = int(arrays["num_instances"])
    d.mem = mem
    for k in _ARRAY_KEYS:
        setattr(d, k, _ptr(arrays[k]))
    for k in _OPT_KEYS:
        setattr(d, k, _ptr(arrays.get(k)) if arrays.get(k) is not None and (not hasattr(arrays[k], "__len__") or len(arrays[k])) else None)
    for k in ("max_host_nodes", "max_host_arcs", "max_guest_nodes", "max_labels", "index_bytes"):
        setattr(d, k, int(arrays.get(k, 0)))
    return d


class Batch:
    """pt_batch_* : batched TreeInNetContainment(...).displayed() (utils/containment.hpp:1179-1202)"""

    def __init__(self, arrays, stream=None):
        """arrays: dict keyed like pt_batch_desc; values numpy arrays (HOST) or torch CUDA tensors (DEVICE), all of one kind"""
        self._keep = arrays
        self.num_instances = int(arrays["num_instances"])
        dev = _is_device(arrays["host_label"])
        self._h = C.c_void_p()
        d = make_desc(arrays, PT_MEM_DEVICE if dev else PT_MEM_HOST)
        _check(_L.pt_batch_create(C.byref(self._h), C.byref(d), stream), "pt_batch_create")

    @classmethod
    def from_newick_text(cls, text, stream=None):
        """pt_batch_create_from_newick: extended-Newick text (two lines per instance) parsed ON THE GPU into a device-resident
        batch; text: bytes / str (HOST) or a torch CUDA uint8 tensor (DEVICE)"""
        self = cls.__new__(cls)
        self._h = C.c_void_p()
        n = C.c_int64(0)
        if _is_device(text):
            self._keep = text
            rc = _L.pt_batch_create_from_newick(C.byref(self._h), text.data_ptr(), int(text.numel()), PT_MEM_DEVICE, C.byref(n), stream)
        else:
            data = text if isinstance(text, (bytes, bytearray)) else text.encode()
            self._keep = data
            rc = _L.pt_batch_create_from_newick(C.byref(self._h), C.cast(C.c_char_p(bytes(data)), C.c_void_p), len(data), PT_MEM_HOST, C.byref(n), stream)
        if rc == _capi.PT_ERR_PARSE:
            raise MalformedNewick((_L.pt_last_error() or b"").decode(errors="replace"))
        _check(rc, "pt_batch_create_from_newick")
        self.num_instances = int(n.value)
        return self

    def download(self):
        """the batch's device arrays copied back to numpy (pt_batch_download), keyed like Packer.arrays()"""
        keys = ["host_node_off", "host_arc_off", "guest_node_off", "host_succ_off", "host_succ_adj", "host_label", "guest_parent", "guest_label"]
        out = dict(num_instances=self.num_instances)
        ib, mx = C.c_int32(0), (C.c_int32 * 4)()
        for which, k in enumerate(keys):
            nb = C.c_int64(0)
            _check(_L.pt_batch_download(self._h, which, None, 0, C.byref(nb), C.byref(ib), mx), "pt_batch_download")
            dt = np.int64 if which < 3 else {2: np.int16, 1: np.uint8}.get(ib.value, np.int32)
            a = np.zeros(max(nb.value // np.dtype(dt).itemsize, 0), dtype=dt)
            if nb.value:
                _check(_L.pt_batch_download(self._h, which, a.ctypes.data, nb.value, C.byref(nb), None, None), "pt_batch_download")
            out[k] = a
        out.update(index_bytes=int(ib.value), max_host_nodes=mx[0], max_host_arcs=mx[1], max_guest_nodes=mx[2], max_labels=mx[3])
        return out

    def contain(self, result_bits=None, status=None, want_counters=True, stream=None, pool_slots=0, max_rounds=0, path=0, no_small_core=0):
        """returns (verdicts bool[B] or None when device outputs are given, status uint8[B] or None, counters dict or None)"""
        B = self.num_instances
        words = (B + 31) // 32
        dev_out = result_bits is not None and _is_device(result_bits)
        if result_bits is None:
            result_bits = np.zeros(max(words, 1), dtype=np.uint32)
            status = np.zeros(max(B, 1), dtype=np.uint8)
        opt = Options(pool_slots, max_rounds, path, no_small_core)
        cnt = Counters() if want_counters else None
        _check(_L.pt_batch_contain(self._h, C.byref(opt), _ptr(result_bits), _ptr(status), PT_MEM_DEVICE if dev_out else PT_MEM_HOST,
                                   C.byref(cnt) if cnt is not None else None, stream), "pt_batch_contain")
        if dev_out:
            return None, None, (cnt.as_dict() if cnt is not None else None)
        bits = np.unpackbits(result_bits[:words].view(np.uint8), bitorder="little")[:B].astype(bool)
        return bits, (status[:B] if status is not None else None), (cnt.as_dict() if cnt is not None else None)

    def last_launches(self):
        return int(_L.pt_batch_last_launches(self._h))

    def round_ms(self, rnd=0):
        ms = C.c_float()
        _check(_L.pt_batch_round_ms(self._h, rnd, C.byref(ms)), "pt_batch_round_ms")
        return float(ms.value)

    def launch_count(self):
        return int(_L.pt_batch_launch_count(self._h))

    def launch_ms(self, launch):
        """device time of the persistent launch of call number `launch` of this handle (the last 64 calls are kept)"""
        ms = C.c_float()
        _check(_L.pt_batch_launch_ms(self._h, int(launch), C.byref(ms)), "pt_batch_launch_ms")
        return float(ms.value)

    def h2d_bytes(self):
        return int(_L.pt_batch_h2d_bytes(self._h))

    def free(self):
        if getattr(self, "_h", None) and _L is not None:
            _L.pt_batch_free(self._h)
            self._h = None

    __del__ = free


def contain_multi(arrays, num_devices=0, pool_slots=0, max_rounds=0):
    """pt_batch_contain_multi: one call for all GPUs of the node (HOST arrays, one host thread per device, no collective)
    -> (verdicts bool[B], status uint8[B], counters dict)"""
    B = int(arrays["num_instances"])
    d = make_desc(arrays, PT_MEM_HOST)
    bits = np.zeros((B + 31) // 32 + 1, dtype=np.uint32)
    status = np.zeros(B + 1, dtype=np.uint8)
    opt, cnt = Options(pool_slots, max_rounds, 0), Counters()
    _check(_L.pt_batch_contain_multi(C.byref(d), int(num_devices), C.byref(opt), _ptr(bits), _ptr(status), C.byref(cnt)), "pt_batch_contain_multi")
    return np.unpackbits(bits.view(np.uint8), bitorder="little")[:B].astype(bool), status[:B], cnt.as_dict()


class Comm:
    """pt_comm_*: the library's own NCCL communicator, one process per GPU (include/pt_gpu.h).  `bootstrap(id_bytes_or_None)`
    must hand rank 0's 128-byte id to every rank (bench.py: a torch.distributed broadcast) and return it."""

    def __init__(self, nranks, rank, bootstrap=None):
        self.nranks, self.rank = int(nranks), int(rank)
        ident = None
        if self.nranks > 1:
            buf = C.create_string_buffer(128)
            if self.rank == 0:
                _check(_L.pt_comm_unique_id(buf), "pt_comm_unique_id")
            ident = bootstrap(bytes(buf.raw) if self.rank == 0 else None)
            ident = C.create_string_buffer(bytes(ident), 128)
        self._h = C.c_void_p()
        _check(_L.pt_comm_init(C.byref(self._h), self.nranks, self.rank, ident), "pt_comm_init")

    def peer_exchange(self):
        """True if the ranks exchange result words through each other's memory (NVLink peer access), False if through NCCL"""
        return bool(_L.pt_comm_peer_exchange(self._h))

    def allreduce(self, words, op="sum", stream=None):
        """in place on a torch CUDA tensor of 4- or 8-byte elements"""
        _check(_L.pt_comm_allreduce(self._h, words.data_ptr(), int(words.numel()), int(words.element_size()), 0 if op == "sum" else 1, stream), "pt_comm_allreduce")

    def contain_sharded(self, batch, first_instance, total_instances, full_bits, full_status=None, want_counters=False, stream=None, pool_slots=0):
        """pt_batch_contain_sharded: solve this rank's range and all-reduce the packed result words of the whole batch
        (full_bits / full_status: torch CUDA tensors for the WHOLE batch) -> counters dict summed over ranks, or None"""
        opt = Options(pool_slots, 0, 0)
        cnt = Counters() if want_counters else None
        _check(_L.pt_batch_contain_sharded(batch._h, self._h, C.byref(opt), int(first_instance), int(total_instances), int(batch.num_instances),
                                           full_bits.data_ptr(), full_status.data_ptr() if full_status is not None else None,
                                           C.byref(cnt) if cnt is not None else None, stream), "pt_batch_contain_sharded")
        return cnt.as_dict() if cnt is not None else None

    def free(self):
        if getattr(self, "_h", None) and _L is not None:
            _L.pt_comm_free(self._h)
            self._h = None

    __del__ = free


def contain_newick(pairs):
    """[(host_newick, guest_newick)] -> (bool verdicts, status) ; the `tc` protocol of examples/tc.cpp:110-142 for a list"""
    p = Packer()
    for h, g in pairs:
        p.add_newick(h, g)
    b = Batch(p.arrays())
    try:
        v, s, _ = b.contain()
    finally:
        b.free()
    return v, s


class Lca:
    """pt_lca_* : Tree::LCA (utils/tree.hpp:202-223) / lca<>::query (rmq/lca.hpp:91-106) for batches of queries"""

    def __init__(self, parent, stream=None):
        self._h = C.c_void_p()
        self._keep = parent
        self.n = int(parent.shape[0]) if hasattr(parent, "shape") else len(parent)
        if not _is_device(parent):
            parent = np.ascontiguousarray(parent, dtype=np.int32)
            self._keep = parent
        _check(_L.pt_lca_build(C.byref(self._h), _ptr(parent), self.n, PT_MEM_DEVICE if _is_device(parent) else PT_MEM_HOST, stream), "pt_lca_build")

    def query(self, u, v, out=None, stream=None):
        dev = _is_device(u)
        if not dev:
            u = np.ascontiguousarray(u, dtype=np.int32)
            v = np.ascontiguousarray(v, dtype=np.int32)
            if out is None:
                out = np.empty(len(u), dtype=np.int32)
        q = int(u.shape[0])
        _check(_L.pt_lca_query(self._h, _ptr(u), _ptr(v), _ptr(out), q, PT_MEM_DEVICE if dev else PT_MEM_HOST, stream), "pt_lca_query")
        return out

    def query_adjacent(self, nodes, out=None, stream=None):
        """pt_lca_query_adjacent: out[i] = lca(nodes[i], nodes[i+1]) (utils/induced_tree.hpp:128-138)"""
        dev = _is_device(nodes)
        if not dev:
            nodes = np.ascontiguousarray(nodes, dtype=np.int32)
            if out is None:
                out = np.empty(max(len(nodes) - 1, 0), dtype=np.int32)
        _check(_L.pt_lca_query_adjacent(self._h, _ptr(nodes), int(nodes.shape[0]), _ptr(out), PT_MEM_DEVICE if dev else PT_MEM_HOST, stream), "pt_lca_query_adjacent")
        return out

    def info(self):
        i = LcaInfo()
        _check(_L.pt_lca_get_info(self._h, C.byref(i)), "pt_lca_get_info")
        return {k: getattr(i, k) for k, _ in i._fields_}

    def node_info(self):
        d, p = np.empty(self.n, dtype=np.int32), np.empty(self.n, dtype=np.int32)
        _check(_L.pt_lca_node_info(self._h, d.ctypes.data, p.ctypes.data, PT_MEM_HOST, None), "pt_lca_node_info")
        return d, p

    def free(self):
        if getattr(self, "_h", None) and _L is not None:
            _L.pt_lca_free(self._h)
            self._h = None

    __del__ = free


def who_displays(host_parent, host_label, guest_parent, guest_label):
    """pt_tree_who_displays: int32[nt], the host node minimally displaying each guest subtree or -1
    (TreeInTreeContainment::who_displays, utils/containment.hpp:72, for a single-labelled tree host)"""
    a = [np.ascontiguousarray(x, dtype=np.int32) for x in (host_parent, host_label, guest_parent, guest_label)]
    out = np.empty(len(a[2]), dtype=np.int32)
    _check(_L.pt_tree_who_displays(a[0].ctypes.data, a[1].ctypes.data, len(a[0]), a[2].ctypes.data, a[3].ctypes.data, len(a[2]), out.ctypes.data, PT_MEM_HOST, None),
           "pt_tree_who_displays")
    return out


class Who:
    """pt_who_*: TreeInTreeContainment<Host, Guest>::who_displays for a (possibly multi-labelled) tree host
    (utils/containment.hpp:27-233): per guest node the host nodes that minimally display it, ascending in host preorder"""

    def __init__(self, host_parent, host_label, guest_parent, guest_label):
        a = [np.ascontiguousarray(x, dtype=np.int32) for x in (host_parent, host_label, guest_parent, guest_label)]
        self.nt = len(a[2])
        self._h = C.c_void_p()
        _check(_L.pt_who_build(C.byref(self._h), a[0].ctypes.data, a[1].ctypes.data, len(a[0]), a[2].ctypes.data, a[3].ctypes.data, len(a[2]), PT_MEM_HOST, None), "pt_who_build")

    def lists(self):
        """(off[nt+1], adj)"""
        t = C.c_int64(0)
        _check(_L.pt_who_total(self._h, C.byref(t)), "pt_who_total")
        off, adj = np.zeros(self.nt + 1, dtype=np.int32), np.zeros(max(t.value, 1), dtype=np.int32)
        _check(_L.pt_who_lists(self._h, off.ctypes.data, adj.ctypes.data, PT_MEM_HOST, None), "pt_who_lists")
        return off, adj[:t.value]

    def who_displays(self, u):
        off, adj = self.lists()
        return adj[off[u]:off[u + 1]]

    def highest_displayed(self, host_root):
        out = np.empty(self.nt, dtype=np.int32)
        _check(_L.pt_who_highest_displayed(self._h, int(host_root), out.ctypes.data, PT_MEM_HOST, None), "pt_who_highest_displayed")
        return out

    def free(self):
        if getattr(self, "_h", None) and _L is not None:
            _L.pt_who_free(self._h)
            self._h = None

    __del__ = free


def net_unzip(succ_off, succ_adj, label, root):
    """pt_net_unzip (TreeInComponent::unzip_retis, utils/containment.hpp:262-312) -> (parent, label, origin) of the MUL-tree"""
    so = np.ascontiguousarray(succ_off, dtype=np.int32); sa = np.ascontiguousarray(succ_adj, dtype=np.int32); lab = np.ascontiguousarray(label, dtype=np.int32)
    n, m = len(so) - 1, len(sa)
    cnt = C.c_int64(0)
    _check(_L.pt_net_unzip(n, m, so.ctypes.data, sa.ctypes.data if m else None, lab.ctypes.data, int(root), None, None, None, 0, C.byref(cnt), PT_MEM_HOST, None), "pt_net_unzip")
    k = cnt.value
    par, ol, org = (np.empty(max(k, 1), dtype=np.int32) for _ in range(3))
    _check(_L.pt_net_unzip(n, m, so.ctypes.data, sa.ctypes.data if m else None, lab.ctypes.data, int(root), par.ctypes.data, ol.ctypes.data, org.ctypes.data, k, C.byref(cnt), PT_MEM_HOST, None), "pt_net_unzip")
    return par[:k], ol[:k], org[:k]


def net_tree_components(succ_off, succ_adj):
    """pt_net_tree_components (TreeComponentInfos, utils/tree_components.hpp:41-234) -> (comp_root, visible_leaf, sorted unique
    component-DAG arcs [(root above, root below)])"""
    so = np.ascontiguousarray(succ_off, dtype=np.int32); sa = np.ascontiguousarray(succ_adj, dtype=np.int32)
    n, m = len(so) - 1, len(sa)
    comp, vis = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int32)
    cap = max(16, 4 * m)
    while True:
        edges = np.zeros(cap, dtype=np.int64)
        cnt = C.c_int64(0)
        _check(_L.pt_net_tree_components(n, m, so.ctypes.data, sa.ctypes.data if m else None, comp.ctypes.data, vis.ctypes.data, edges.ctypes.data, cap, C.byref(cnt), PT_MEM_HOST, None), "pt_net_tree_components")
        if cnt.value <= cap:
            break
        cap = int(cnt.value)
    e = np.unique(edges[:cnt.value])
    return comp, vis, [(int(x >> 32), int(x & 0xFFFFFFFF)) for x in e]


def net_visible_component(succ_off, succ_adj, host_label, guest_parent, guest_label, root=-1, leaf=-1, unzip_cap=0):
    """pt_net_visible_component (VisibleComponentRule, utils/containment.hpp:603-725) -> (eligibility uint8[n], root, visible leaf,
    guest node); eligibility bits: 1 half-eligible, 2 eligible, 4 eligible and lowest; -1 where the rule has no answer"""
    so = np.ascontiguousarray(succ_off, dtype=np.int32); sa = np.ascontiguousarray(succ_adj, dtype=np.int32)
    hl = np.ascontiguousarray(host_label, dtype=np.int32); gp = np.ascontiguousarray(guest_parent, dtype=np.int32); gl = np.ascontiguousarray(guest_label, dtype=np.int32)
    n, m, nt = len(so) - 1, len(sa), len(gp)
    if len(hl) != n or len(gl) != nt:
        raise ValueError("label arrays do not match the node counts")
    el = np.zeros(n, dtype=np.uint8)
    r, lf, g = C.c_int32(-1), C.c_int32(-1), C.c_int32(-1)
    _check(_L.pt_net_visible_component(n, m, so.ctypes.data, sa.ctypes.data if m else None, hl.ctypes.data, nt, gp.ctypes.data, gl.ctypes.data, int(root), int(leaf), int(unzip_cap),
                                       el.ctypes.data, C.byref(r), C.byref(lf), C.byref(g), None), "pt_net_visible_component")
    return el, r.value, lf.value, g.value


def highest_displayed(guest_parent, who, host_root):
    """pt_tree_highest_displayed: int32[nt], TreeInComponent::highest_displayed_ancestor (utils/containment.hpp:328-344) of every
    guest node, from the table who_displays returned"""
    gp = np.ascontiguousarray(guest_parent, dtype=np.int32); wd = np.ascontiguousarray(who, dtype=np.int32)
    if len(gp) != len(wd):
        raise ValueError("guest_parent and who differ in length")
    out = np.empty(len(gp), dtype=np.int32)
    _check(_L.pt_tree_highest_displayed(gp.ctypes.data, wd.ctypes.data, len(gp), int(host_root), out.ctypes.data, PT_MEM_HOST, None), "pt_tree_highest_displayed")
    return out


def net_bridges(succ_off, succ_adj):
    """pt_net_bridges -> (bool[m] is_bridge per CSR arc, int32[n] component id); utils/bridges.hpp"""
    so = np.ascontiguousarray(succ_off, dtype=np.int32)
    sa = np.ascontiguousarray(succ_adj, dtype=np.int32)
    n, m = len(so) - 1, len(sa)
    br = np.zeros(max(m, 1), dtype=np.uint8)
    comp = np.empty(n, dtype=np.int32)
    _check(_L.pt_net_bridges(n, m, so.ctypes.data, sa.ctypes.data, br.ctypes.data, comp.ctypes.data, PT_MEM_HOST, None), "pt_net_bridges")
    return br[:m].astype(bool), comp


class Rmq:
    """pt_rmq_* : sparse_rmq (rmq/sparse_rmq.hpp:45-90): index of the rightmost minimum of [l, r)"""

    def __init__(self, values):
        values = np.ascontiguousarray(values, dtype=np.int32)
        self._h = C.c_void_p()
        _check(_L.pt_rmq_build(C.byref(self._h), values.ctypes.data, len(values), PT_MEM_HOST, None), "pt_rmq_build")

    def query(self, l, r):
        l = np.ascontiguousarray(l, dtype=np.int64)
        r = np.ascontiguousarray(r, dtype=np.int64)
        out = np.empty(len(l), dtype=np.int64)
        _check(_L.pt_rmq_query(self._h, l.ctypes.data, r.ctypes.data, out.ctypes.data, len(l), PT_MEM_HOST, None), "pt_rmq_query")
        return out

    def free(self):
        if getattr(self, "_h", None) and _L is not None:
            _L.pt_rmq_free(self._h)
            self._h = None

    __del__ = free


def enum_switchings(succ_off, succ_adj, host_label, guest_parent, guest_label, begin=0, end=0, stop_at_first=False):
    """pt_enum_switchings -> (n_switchings, n_displaying, first_index or None)"""
    a = [np.ascontiguousarray(x, dtype=np.int32) for x in (succ_off, succ_adj, host_label, guest_parent, guest_label)]
    tot, nd, fi = C.c_uint64(), C.c_uint64(), C.c_uint64()
    _check(_L.pt_enum_switchings(len(a[2]), len(a[1]), a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, len(a[3]), a[3].ctypes.data, a[4].ctypes.data,
                                 begin, end, 1 if stop_at_first else 0, C.byref(tot), C.byref(nd), C.byref(fi), None), "pt_enum_switchings")
    return tot.value, nd.value, (None if fi.value == 2**64 - 1 else fi.value)


class Enum:
    """pt_enum_*: the switching enumerator with a persistent handle (no allocation / upload per run) and, with a Comm, the
    cross-GPU found flag: every rank stops as soon as any rank has a witness"""

    def __init__(self, succ_off, succ_adj, host_label, guest_parent, guest_label):
        a = [np.ascontiguousarray(x, dtype=np.int32) for x in (succ_off, succ_adj, host_label, guest_parent, guest_label)]
        self._h = C.c_void_p()
        tot = C.c_uint64()
        _check(_L.pt_enum_create(C.byref(self._h), len(a[2]), len(a[1]), a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, len(a[3]), a[3].ctypes.data, a[4].ctypes.data,
                                 C.byref(tot), None), "pt_enum_create")
        self.total = tot.value

    def run(self, begin=0, end=0, stop_at_first=False, comm=None, chunk=0):
        """-> (n_displaying, first_index or None, stopped_early)"""
        nd, fi, early = C.c_uint64(), C.c_uint64(), C.c_int(0)
        _check(_L.pt_enum_run(self._h, begin, end, 1 if stop_at_first else 0, comm._h if comm is not None else None, chunk, C.byref(nd), C.byref(fi), C.byref(early), None), "pt_enum_run")
        return nd.value, (None if fi.value == 2**64 - 1 else fi.value), bool(early.value)

    def free(self):
        if getattr(self, "_h", None) and _L is not None:
            _L.pt_enum_free(self._h)
            self._h = None

    __del__ = free


def _csr(n, keys, vals):
    """CSR of `vals` grouped by `keys` (stable): (off[n+1], adj[m]) as int32"""
    keys, vals = np.asarray(keys, dtype=np.int64), np.asarray(vals, dtype=np.int32)
    off = np.zeros(n + 1, dtype=np.int32)
    if len(keys):
        np.add.at(off, keys + 1, 1)
    off = np.cumsum(off, dtype=np.int64).astype(np.int32)
    order = np.argsort(keys, kind="stable")
    return off, vals[order].astype(np.int32)


def iso_arrays(networks):
    """[(n, edges, labels) ...] with labels as dense ids (or -1), two consecutive entries per pair -> dict keyed like pt_iso_desc"""
    node_off, arc_off = [0], [0]
    so, sa, po, pa, lab = [], [], [], [], []
    for n, edges, labels in networks:
        e = np.asarray(edges, dtype=np.int32).reshape(-1, 2)
        o, a = _csr(n, e[:, 0], e[:, 1])
        so.append(o); sa.append(a)
        o, a = _csr(n, e[:, 1], e[:, 0])
        po.append(o); pa.append(a)
        lab.append(np.asarray(labels, dtype=np.int32))
        node_off.append(node_off[-1] + n); arc_off.append(arc_off[-1] + len(e))
    cat = lambda xs: np.ascontiguousarray(np.concatenate(xs) if xs else np.zeros(0, np.int32), dtype=np.int32)
    return dict(num_pairs=len(networks) // 2, node_off=np.asarray(node_off, dtype=np.int64), arc_off=np.asarray(arc_off, dtype=np.int64),
                succ_off=cat(so), succ_adj=cat(sa), pred_off=cat(po), pred_adj=cat(pa), label=cat(lab),
                max_nodes=max([n for n, _, _ in networks], default=0))


def iso_batch(arrays, flags=1):
    """pt_iso_batch: make_iso_mapper(N0, N1, flags).check_isomorph() (utils/isomorphism.hpp) for a batch of pairs
    -> (bool verdicts[P], status uint8[P]).  arrays: dict from iso_arrays (numpy = HOST)"""
    P = int(arrays["num_pairs"])
    d = _capi.IsoDesc()
    d.num_pairs = P; d.mem = PT_MEM_HOST; d.flags = int(flags); d.max_nodes = int(arrays.get("max_nodes", 0))
    keep = {}
    for k in ("node_off", "arc_off", "succ_off", "succ_adj", "pred_off", "pred_adj", "label"):
        a = arrays[k]
        if len(a) == 0:
            a = np.zeros(1, dtype=a.dtype)
        keep[k] = a
        setattr(d, k, a.ctypes.data)
    bits = np.zeros((P + 31) // 32 + 1, dtype=np.uint32)
    status = np.zeros(P + 1, dtype=np.uint8)
    _check(_L.pt_iso_batch(C.byref(d), bits.ctypes.data, status.ctypes.data, PT_MEM_HOST, None), "pt_iso_batch")
    return np.unpackbits(bits.view(np.uint8), bitorder="little")[:P].astype(bool), status[:P]


def iso_newick(pairs, flags=1):
    """[(newick0, newick1)] -> (bool verdicts, status): the `iso` tool of examples/iso.cpp for a list of pairs (label strings get
    dense ids shared by the two networks of a pair)"""
    nets = []
    for a, b in pairs:
        ids = {}
        for s in (a, b):
            n, edges, labels = parse_newick(s)
            lab = np.full(n, -1, dtype=np.int32)
            for v, name in labels.items():
                if name != "":
                    lab[v] = ids.setdefault(name, len(ids))
            nets.append((n, edges, lab))
    return iso_batch(iso_arrays(nets), flags)


def iso_newick_text(text, flags=1):
    """pt_iso_from_newick: two extended-Newick lines per pair, parsed and compared on the GPU -> (bool verdicts, status);
    text: bytes / str (HOST) or a torch CUDA uint8 tensor (DEVICE: verdict words and status then stay on the device until read here)"""
    n = C.c_int64(0)
    if _is_device(text):
        import torch
        size = int(text.numel())
        cap = size // 4 + 2
        d_bits = torch.zeros((cap + 31) // 32 + 1, dtype=torch.int32, device=text.device)
        d_stat = torch.zeros(cap + 1, dtype=torch.uint8, device=text.device)
        rc = _L.pt_iso_from_newick(text.data_ptr(), size, PT_MEM_DEVICE, int(flags), d_bits.data_ptr(), d_stat.data_ptr(), PT_MEM_DEVICE, C.byref(n), None)
        bits, status = d_bits.cpu().numpy().view(np.uint32), d_stat.cpu().numpy()
    else:
        data = text if isinstance(text, (bytes, bytearray)) else text.encode()
        cap = len(data) // 4 + 2
        bits = np.zeros((cap + 31) // 32 + 1, dtype=np.uint32)
        status = np.zeros(cap + 1, dtype=np.uint8)
        rc = _L.pt_iso_from_newick(C.cast(C.c_char_p(bytes(data)), C.c_void_p), len(data), PT_MEM_HOST, int(flags), bits.ctypes.data, status.ctypes.data, PT_MEM_HOST, C.byref(n), None)
    if rc == _capi.PT_ERR_PARSE:
        raise MalformedNewick((_L.pt_last_error() or b"").decode(errors="replace"))
    _check(rc, "pt_iso_from_newick")
    P = int(n.value)
    return np.unpackbits(bits.view(np.uint8), bitorder="little")[:P].astype(bool), status[:P]
