"""ctypes binding of the C-ABI (include/mpgpu.h) -- the same calls the reference-side shim
makes from sprparsimony.cpp / iqtree.cpp.  Used by tests/ and bench.py.

There is no CPU fallback: importing this module needs the built library, and creating an
engine needs a CUDA device.  Nothing here touches oracle/."""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libmpgpu.so")

RNG_FN = C.CFUNCTYPE(C.c_double, C.c_void_p)

_lib = None


class MpGpuError(RuntimeError):
    pass


def lib():
    """Load libmpgpu.so (built in-tree by __graft_entry__.build() / mpboot_b200/csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MpGpuError("libmpgpu.so is not built (%s): run `python -c 'import __graft_entry__ as g; g.build()'`; "
                         "there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u32p = C.c_void_p, C.c_int, C.c_int64, C.c_void_p
    L.mpgpu_last_error.restype = C.c_char_p
    L.mpgpu_device_count.restype = i32
    L.mpgpu_create.argtypes = [C.POINTER(vp), i32, vp, i32, i32]
    L.mpgpu_destroy.argtypes = [vp]
    L.mpgpu_set_allreduce.argtypes = [vp, vp, vp]
    L.mpgpu_peer_prepare.argtypes = [vp, i64, vp]
    L.mpgpu_peer_connect.argtypes = [vp, vp]
    L.mpgpu_peer_stats.argtypes = [vp, vp, vp, vp]
    L.mpgpu_search_info.argtypes = [vp, vp, vp, vp]
    L.mpgpu_split_table.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, i32]
    L.mpgpu_stream.restype = vp
    L.mpgpu_stream.argtypes = [vp]
    L.mpgpu_synchronize.argtypes = [vp]
    L.mpgpu_launch_count.restype = i64
    L.mpgpu_launch_count.argtypes = [vp]
    L.mpgpu_load_alignment.argtypes = [vp, i32, i32, i32, vp, vp, i32]
    L.mpgpu_set_weights.argtypes = [vp, vp]
    L.mpgpu_get_layout.argtypes = [vp, vp, vp, vp, vp, vp]
    L.mpgpu_get_tip_planes.argtypes = [vp, i32, vp]
    L.mpgpu_set_tree.argtypes = [vp, vp, vp]
    L.mpgpu_get_view_counts_partial.argtypes = [vp, vp]
    L.mpgpu_set_view_counts.argtypes = [vp, vp]
    L.mpgpu_view_length.argtypes = [vp, i32, i32, vp]
    L.mpgpu_get_view_planes.argtypes = [vp, i32, i32, vp]
    L.mpgpu_tree_score.argtypes = [vp, vp]
    L.mpgpu_edge_mismatch_partial.argtypes = [vp, i32, i32, vp]
    L.mpgpu_pattern_parsimony.argtypes = [vp, vp, vp]
    L.mpgpu_visit_order.argtypes = [vp, vp]
    L.mpgpu_scan_visits.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, i32, vp]
    L.mpgpu_scan_plan.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp]
    L.mpgpu_scan_launch.argtypes = [vp, vp]
    L.mpgpu_scan_plan_bytes.restype = i64
    L.mpgpu_scan_plan_bytes.argtypes = [vp]
    L.mpgpu_scan_finish.argtypes = [vp, vp, vp, vp, vp, i32]
    L.mpgpu_optimize_spr.argtypes = [vp, vp, vp, i32, i32, vp, vp, vp, vp]
    L.mpgpu_stepwise_addition.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp, vp]
    L.mpgpu_refine_replicates.argtypes = [vp, i32, vp, i32, vp, vp, i32, i32, vp, vp, vp, vp]
    L.mpgpu_load_replicates.argtypes = [vp, i32, vp, i32, vp, i32]
    L.mpgpu_load_replicates2.argtypes = [vp, i32, vp, i32, vp, i32, vp]
    L.mpgpu_reps_info.argtypes = [vp, vp, vp, vp]
    L.mpgpu_reps_timing.argtypes = [vp, vp, vp, vp, vp]
    L.mpgpu_int8_peak.argtypes = [vp, C.c_int, vp]
    L.mpgpu_set_replicate_shards.argtypes = [vp, C.c_int, C.c_int]
    L.mpgpu_set_option.argtypes = [vp, C.c_char_p, i32]
    L.mpgpu_set_cost_matrix.argtypes = [vp, vp, i32, vp, i32, vp]
    L.mpgpu_sankoff_layout.argtypes = [vp, vp, vp, vp, i32]
    L.mpgpu_sankoff_view.argtypes = [vp, i32, i32, vp]
    L.mpgpu_scan_bounds.argtypes = [vp, vp, i32]
    L.mpgpu_sankoff_reps_stats.argtypes = [vp, vp, vp]
    L.mpgpu_reps_current_tree.argtypes = [vp, vp]
    L.mpgpu_reps_candidates.argtypes = [vp, vp, i32, vp]
    L.mpgpu_reps_candidates_device.argtypes = [vp, vp, i32, vp, vp]
    L.mpgpu_optimize_spr_bb.argtypes = [vp, vp, vp, i32, i32, vp, vp, vp, vp]
    L.mpgpu_splitmix64_double.restype = C.c_double
    L.mpgpu_splitmix64_double.argtypes = [vp]
    L.mpgpu_treels_create.restype = vp
    L.mpgpu_treels_create.argtypes = [i32]
    L.mpgpu_treels_destroy.argtypes = [vp]
    L.mpgpu_treels_size.restype = i64
    L.mpgpu_treels_size.argtypes = [vp]
    L.mpgpu_treels_logl.argtypes = [vp, vp]
    L.mpgpu_treels_num_materialized.restype = i64
    L.mpgpu_treels_num_materialized.argtypes = [vp]
    L.mpgpu_treels_materialized.argtypes = [vp, vp]
    L.mpgpu_treels_hooks.argtypes = [vp, vp, vp, vp]
    L.mpgpu_treels_mulhits.restype = i64
    L.mpgpu_treels_mulhits.argtypes = [vp, i32, vp, vp, i64]
    L.mpgpu_treels_toplists.restype = i64
    L.mpgpu_treels_toplists.argtypes = [vp, i32, vp, vp, i64]
    L.mpgpu_treels_topiters.restype = i64
    L.mpgpu_treels_topiters.argtypes = [vp, i32, vp, i64]
    L.mpgpu_set_remain_bounds.argtypes = [vp, vp, C.c_int]
    L.mpgpu_reps_prefix_max.argtypes = [vp, i32, vp, C.c_int, vp]
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class BBHooks(C.Structure):
    """mpgpu_bb_hooks (include/mpgpu.h)"""
    _fields_ = [("user", C.c_void_p), ("random_double", C.c_void_p), ("push_tree_logl", C.c_void_p),
                ("materialize", C.c_void_p), ("mulhit", C.c_void_p), ("tophit", C.c_void_p), ("disthit", C.c_void_p)]


class BBState(C.Structure):
    """mpgpu_bb_state (include/mpgpu.h)"""
    _fields_ = [("B", C.c_int32), ("boot_logl", C.c_void_p), ("boot_counts", C.c_void_p), ("boot_trees", C.c_void_p),
                ("logl_cutoff", C.c_double), ("ufboot_epsilon", C.c_double), ("n_calls", C.c_int64), ("n_reps", C.c_int64),
                ("ratchet", C.c_int32), ("ratchet_pattern_pars", C.c_void_p), ("ratchet_last_score", C.c_int32),
                ("policy", C.c_int32), ("top_n", C.c_int32), ("top_count", C.c_void_p), ("boot_threshold", C.c_void_p),
                ("cur_it", C.c_int32), ("updates_off", C.c_int32), ("boot_tree_orig_logl", C.c_void_p)]


class HostRng:
    """The library's splitmix64 random_double (mpgpu_splitmix64_double) with its state: pass .fn and
    .user wherever an mpgpu_rng_fn / user pair is expected."""

    def __init__(self, seed):
        self.state = C.c_uint64(seed)
        self.fn = C.cast(lib().mpgpu_splitmix64_double, C.c_void_p).value
        self.user = C.addressof(self.state)


class Treels:
    """mpgpu_treels: the host-side treels / treels_logl container of include/mpgpu.h."""

    def __init__(self, ntaxa):
        self.L = lib()
        self.h = C.c_void_p(self.L.mpgpu_treels_create(ntaxa))

    def __del__(self):
        try:
            if self.h:
                self.L.mpgpu_treels_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def hooks(self, rng_fn_ptr, rng_user=None):
        hk = BBHooks()
        self.L.mpgpu_treels_hooks(self.h, C.c_void_p(rng_fn_ptr), C.c_void_p(rng_user) if rng_user else None, C.byref(hk))
        return hk

    def logl(self):
        k = self.L.mpgpu_treels_size(self.h)
        out = np.zeros(max(k, 1), dtype=np.float64)
        self.L.mpgpu_treels_logl(self.h, _p(out))
        return out[:k]

    def mulhits(self, nsamples):
        """boot_trees_parsimony (-mulhits): (sizes[nsamples], members ascending, concatenated)"""
        sizes = np.zeros(nsamples, dtype=np.int32)
        tot = self.L.mpgpu_treels_mulhits(self.h, nsamples, _p(sizes), None, 0)
        flat = np.zeros(max(tot, 1), dtype=np.int32)
        self.L.mpgpu_treels_mulhits(self.h, nsamples, _p(sizes), _p(flat), tot)
        return sizes, flat[:tot]

    def toplists(self, nsamples):
        """boot_trees_parsimony_top (-mulhits -topboot): (sizes[nsamples], (tree_index, rell) pairs in list order)"""
        sizes = np.zeros(nsamples, dtype=np.int32)
        tot = self.L.mpgpu_treels_toplists(self.h, nsamples, _p(sizes), None, 0)
        flat = np.zeros((max(tot, 1), 2), dtype=np.int32)
        self.L.mpgpu_treels_toplists(self.h, nsamples, _p(sizes), _p(flat), tot)
        return sizes, flat[:tot]

    def topiters(self, nsamples):
        """boot_trees_parsimony_top_iter (-distinct_iter_top_boot), in the order of toplists' pairs"""
        tot = self.L.mpgpu_treels_topiters(self.h, nsamples, None, 0)
        flat = np.zeros(max(tot, 1), dtype=np.int32)
        self.L.mpgpu_treels_topiters(self.h, nsamples, _p(flat), tot)
        return flat[:tot]

    def materialized(self):
        k = self.L.mpgpu_treels_num_materialized(self.h)
        out = np.zeros((max(k, 1), 4), dtype=np.int64)
        self.L.mpgpu_treels_materialized(self.h, _p(out))
        return out[:k]


def device_count():
    return lib().mpgpu_device_count()


class Engine:
    """One mpgpu context (one GPU, one word-slice of the alignment)."""

    def __init__(self, device=0, stream=None, shard_rank=0, shard_count=1):
        self.L = lib()
        h = C.c_void_p()
        rc = self.L.mpgpu_create(C.byref(h), device, stream, shard_rank, shard_count)
        if rc:
            raise MpGpuError(self.L.mpgpu_last_error().decode())
        self.h = h
        self.n = self.P = 0
        self.shard_count = shard_count

    def set_allreduce(self, callback):
        """callback: a ctypes function object of type mpgpu_allreduce_fn (see mpboot_b200.sharded)."""
        self._allreduce_cb = callback                    # keep it alive
        self._ck(self.L.mpgpu_set_allreduce(self.h, C.cast(callback, C.c_void_p), None))

    def peer_prepare(self, capacity=1 << 16):
        """mpgpu_peer_prepare: this shard's exchange region for vectors of up to `capacity` int32; returns its CUDA IPC
        handle (64 bytes) for the other shards."""
        h = np.zeros(64, dtype=np.uint8)
        self._ck(self.L.mpgpu_peer_prepare(self.h, int(capacity), _p(h)))
        return h

    def peer_connect(self, handles):
        """mpgpu_peer_connect: handles = uint8 [shard_count][64], the IPC handles of all shards in shard order."""
        h = np.ascontiguousarray(handles, dtype=np.uint8)
        group = self.shard_count if self.shard_count > 1 else getattr(self, "rep_count", 1)      # pattern or replicate shards
        assert h.shape == (group, 64)
        self._ck(self.L.mpgpu_peer_connect(self.h, _p(h)))

    def peer_stats(self):
        calls, elems, err = C.c_int64(0), C.c_int64(0), C.c_int(0)
        self._ck(self.L.mpgpu_peer_stats(self.h, C.byref(calls), C.byref(elems), C.byref(err)))
        return calls.value, elems.value, err.value

    def split_table(self, ntaxa, tokens, token_begin, tree_weight):
        """mpgpu_split_table: (split_bits [U][W] uint32, split_weight [U], emit_unique [tokens]) in first-seen order."""
        tokens = np.ascontiguousarray(tokens, dtype=np.int32)
        token_begin = np.ascontiguousarray(token_begin, dtype=np.int64)
        tree_weight = np.ascontiguousarray(tree_weight, dtype=np.int32)
        ntrees = len(tree_weight)
        assert len(token_begin) == ntrees + 1
        nemit = int(token_begin[-1] - token_begin[0])
        W = (ntaxa + 31) // 32
        nu = C.c_int32(0)
        bits = np.zeros((nemit, W), dtype=np.uint32)
        wgt = np.zeros(nemit, dtype=np.int32)
        eu = np.zeros(nemit, dtype=np.int32)
        self._ck(self.L.mpgpu_split_table(self.h, int(ntaxa), ntrees, _p(tokens), _p(token_begin), _p(tree_weight), C.byref(nu),
                                          _p(bits), _p(wgt), _p(eu), nemit))
        return bits[: nu.value].copy(), wgt[: nu.value].copy(), eu

    def search_info(self):
        """(score of the start tree, moves applied, scan batches) of the last SPR search on this context."""
        s, m, b = C.c_uint32(0), C.c_int64(0), C.c_int64(0)
        self._ck(self.L.mpgpu_search_info(self.h, C.byref(s), C.byref(m), C.byref(b)))
        return s.value, m.value, b.value

    def _ck(self, rc):
        if rc:
            raise MpGpuError(self.L.mpgpu_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.mpgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- R1
    def load_alignment(self, codes, weights, datatype, sort_alignment=True):
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        weights = np.ascontiguousarray(weights, dtype=np.int32)
        self.n, self.P = codes.shape
        self._ck(self.L.mpgpu_load_alignment(self.h, self.n, self.P, datatype, _p(codes), _p(weights), int(sort_alignment)))
        lay = self.layout()
        self.S, self.ref_words, self.shard_words, self.n_inf, self.n_sites = lay
        return lay

    def set_weights(self, weights):
        weights = np.ascontiguousarray(weights, dtype=np.int32)
        self._ck(self.L.mpgpu_set_weights(self.h, _p(weights)))
        self.S, self.ref_words, self.shard_words, self.n_inf, self.n_sites = self.layout()

    def layout(self):
        s, rw, sw, ni = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        ns = C.c_int64()
        self._ck(self.L.mpgpu_get_layout(self.h, C.byref(s), C.byref(rw), C.byref(sw), C.byref(ni), C.byref(ns)))
        return s.value, rw.value, sw.value, ni.value, ns.value

    def tip_planes(self, tip):
        out = np.zeros((self.S, self.ref_words), dtype=np.uint32)
        self._ck(self.L.mpgpu_get_tip_planes(self.h, tip, _p(out)))
        return out

    # -- R2/R3/R4
    def set_tree(self, bn, bs):
        bn = np.ascontiguousarray(bn, dtype=np.int32); bs = np.ascontiguousarray(bs, dtype=np.int32)
        self._ck(self.L.mpgpu_set_tree(self.h, _p(bn), _p(bs)))

    def view_counts_partial(self):
        out = np.zeros(4 * self.n - 6, dtype=np.uint32)
        self._ck(self.L.mpgpu_get_view_counts_partial(self.h, _p(out)))
        return out

    def set_view_counts(self, counts):
        counts = np.ascontiguousarray(counts, dtype=np.uint32)
        self._ck(self.L.mpgpu_set_view_counts(self.h, _p(counts)))

    def view_length(self, node, slot):
        v = C.c_uint32()
        self._ck(self.L.mpgpu_view_length(self.h, node, slot, C.byref(v)))
        return v.value

    def view_planes(self, node, slot):
        out = np.zeros((self.S, self.ref_words), dtype=np.uint32)
        self._ck(self.L.mpgpu_get_view_planes(self.h, node, slot, _p(out)))
        return out

    def tree_score(self):
        v = C.c_uint32()
        self._ck(self.L.mpgpu_tree_score(self.h, C.byref(v)))
        return v.value

    def edge_mismatch_partial(self, node, slot):
        v = C.c_uint32()
        self._ck(self.L.mpgpu_edge_mismatch_partial(self.h, node, slot, C.byref(v)))
        return v.value

    # -- R5
    def pattern_parsimony(self):
        upper = self.n_inf if getattr(self, "_sort", True) else self.P
        out = np.zeros(max(self.P, 1) + 16, dtype=np.uint16)
        s = C.c_int32()
        self._ck(self.L.mpgpu_pattern_parsimony(self.h, _p(out), C.byref(s)))
        return out, s.value

    # -- R6
    def visit_order(self):
        out = np.zeros(2 * self.n - 1, dtype=np.int32)
        self._ck(self.L.mpgpu_visit_order(self.h, _p(out)))
        return out

    def scan_visits(self, order, first, count, mintrav=1, maxtrav=6, capacity=None, reuse=False):
        """reuse=True: the output arrays of the previous call with the same shape are written again (what a C caller does with its
        own buffers); the returned views are then only valid until the next such call."""
        if not (isinstance(order, np.ndarray) and order.dtype == np.int32 and order.flags.c_contiguous):
            order = np.ascontiguousarray(order, dtype=np.int32)
        if capacity is None:
            capacity = count * (8 << min(maxtrav, 10)) + 16
        key = (count, capacity)
        if reuse and getattr(self, "_scan_out_key", None) == key:
            vb, mp, cr, cp, ptrs = self._scan_out
        else:
            vb = np.zeros(count + 1, dtype=np.int32)
            mp = np.zeros(capacity, dtype=np.uint32)
            cr = np.zeros(capacity, dtype=np.int32)
            cp = np.zeros(capacity, dtype=np.int32)
            ptrs = (_p(vb), _p(mp), _p(cr), _p(cp))
            if reuse:
                self._scan_out_key, self._scan_out = key, (vb, mp, cr, cp, ptrs)
        nc = C.c_int()
        self._ck(self.L.mpgpu_scan_visits(self.h, _p(order), first, count, mintrav, maxtrav,
                                          ptrs[0], ptrs[1], ptrs[2], ptrs[3], capacity, C.byref(nc)))
        k = nc.value
        return vb, mp[:k], cr[:k], cp[:k]

    def scan_plan(self, order, first, count, mintrav=1, maxtrav=6):
        order = np.ascontiguousarray(order, dtype=np.int32)
        nc, nt = C.c_int(), C.c_int()
        self._ck(self.L.mpgpu_scan_plan(self.h, _p(order), first, count, mintrav, maxtrav, C.byref(nc), C.byref(nt)))
        return nc.value, nt.value

    def scan_plan_bytes(self):
        return self.L.mpgpu_scan_plan_bytes(self.h)

    def scan_launch(self):
        ptr = C.c_void_p()
        self._ck(self.L.mpgpu_scan_launch(self.h, C.byref(ptr)))
        return ptr.value

    def scan_finish(self, n_cand, count):
        vb = np.zeros(count + 1, dtype=np.int32)
        mp = np.zeros(n_cand + 1, dtype=np.uint32)
        cr = np.zeros(n_cand + 1, dtype=np.int32)
        cp = np.zeros(n_cand + 1, dtype=np.int32)
        self._ck(self.L.mpgpu_scan_finish(self.h, _p(vb), _p(mp), _p(cr), _p(cp), n_cand + 1))
        return vb, mp[:n_cand], cr[:n_cand], cp[:n_cand]

    def optimize_spr(self, bn, bs, rng_fn_ptr, mintrav=1, maxtrav=6, rng_user=None):
        """pllOptimizeSprParsimony.  rng_fn_ptr: address of a `double f(void*)` (the host's
        random_double).  Returns (startMP, back_node, back_slot, insertions scored)."""
        bn = np.array(bn, dtype=np.int32, copy=True); bs = np.array(bs, dtype=np.int32, copy=True)
        best = C.c_uint32(); nins = C.c_int64()
        self._ck(self.L.mpgpu_optimize_spr(self.h, _p(bn), _p(bs), mintrav, maxtrav,
                                           C.c_void_p(rng_fn_ptr), C.c_void_p(rng_user) if rng_user else None,
                                           C.byref(best), C.byref(nins)))
        return best.value, bn, bs, nins.value

    def refine_replicates(self, boot, trees_bn, trees_bs, rng_fn_ptr, mintrav=1, maxtrav=6, rng_user=None):
        """The default-policy loop of IQTree::optimizeBootTrees: replicate b = frequencies boot[b] over the
        resident codes + its tree (ring tables, rows of trees_bn/trees_bs).  Returns (scores, trees_bn,
        trees_bs, insertions scored)."""
        boot = np.ascontiguousarray(boot, dtype=np.uint16)
        B, stride = boot.shape
        tbn = np.array(trees_bn, dtype=np.int32, copy=True).reshape(B, -1)
        tbs = np.array(trees_bs, dtype=np.int32, copy=True).reshape(B, -1)
        assert tbn.shape[1] == 3 * (2 * self.n - 1)
        scores = np.zeros(B, dtype=np.uint32); nins = C.c_int64()
        self._ck(self.L.mpgpu_refine_replicates(self.h, B, _p(boot), stride, _p(tbn), _p(tbs), mintrav, maxtrav,
                                                C.c_void_p(rng_fn_ptr), C.c_void_p(rng_user) if rng_user else None,
                                                _p(scores), C.byref(nins)))
        return scores, tbn, tbs, nins.value

    # -- R11 (-cost)
    def set_cost_matrix(self, cost, segment_upper):
        """pllCostMatrix + pllSegmentUpper: switches the context to Sankoff weighted parsimony (None: back to Fitch).
        Returns highest_cost."""
        if cost is None:
            self._ck(self.L.mpgpu_set_cost_matrix(self.h, None, 0, None, 0, None))
            return 0
        cost = np.ascontiguousarray(cost, dtype=np.uint32)
        seg = np.ascontiguousarray(segment_upper, dtype=np.int32)
        hi = C.c_uint32()
        self._ck(self.L.mpgpu_set_cost_matrix(self.h, _p(cost), cost.shape[0], _p(seg), len(seg), C.byref(hi)))
        return hi.value

    def sankoff_layout(self):
        L, nb = C.c_int(), C.c_int()
        self._ck(self.L.mpgpu_sankoff_layout(self.h, C.byref(L), C.byref(nb), None, 0))
        lb = np.zeros(max(nb.value, 1), dtype=np.uint32)
        self._ck(self.L.mpgpu_sankoff_layout(self.h, None, None, _p(lb), nb.value))
        return L.value, lb[:nb.value]

    def sankoff_view(self, node, slot=0):
        L, _ = self.sankoff_layout()
        out = np.zeros((L, self.S), dtype=np.uint16)
        self._ck(self.L.mpgpu_sankoff_view(self.h, node, slot, _p(out)))
        return out

    def sankoff_reps_stats(self):
        """(chunks contracted on the tensor cores, chunks through the exact CUDA-core kernel) under -cost -bb"""
        a, b = C.c_int64(), C.c_int64()
        self._ck(self.L.mpgpu_sankoff_reps_stats(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def scan_bounds(self, n_cand):
        out = np.zeros(max(n_cand, 1), dtype=np.uint32)
        self._ck(self.L.mpgpu_scan_bounds(self.h, _p(out), max(n_cand, 1)))
        return out[:n_cand]

    # -- R7
    def stepwise_addition(self, seed, spr_dist, rng_fn_ptr, rng_user=None):
        """_pllComputeRandomizedStepwiseAdditionParsimonyTree.  Returns (bestParsimony, back_node,
        back_slot, insertions scored, updated seed)."""
        bn = np.zeros(3 * (2 * self.n - 1), dtype=np.int32); bs = np.zeros_like(bn)
        sd = C.c_int64(seed); best = C.c_uint32(); nins = C.c_int64()
        self._ck(self.L.mpgpu_stepwise_addition(self.h, C.byref(sd), spr_dist, C.c_void_p(rng_fn_ptr),
                                                C.c_void_p(rng_user) if rng_user else None, _p(bn), _p(bs),
                                                C.byref(best), C.byref(nins)))
        return best.value, bn, bs, nins.value, sd.value

    # -- R8 / -bb
    def set_option(self, name, value):
        self._ck(self.L.mpgpu_set_option(self.h, name.encode(), int(value)))

    def load_replicates(self, boot, segment_upper, original_sample=None):
        boot = np.ascontiguousarray(boot, dtype=np.uint16)
        seg = np.ascontiguousarray(segment_upper, dtype=np.int32)
        self.B = boot.shape[0]
        if original_sample is None:
            self._ck(self.L.mpgpu_load_replicates(self.h, boot.shape[0], _p(boot), boot.shape[1], _p(seg), len(seg)))
        else:
            orig = np.zeros(max(self.P, len(original_sample)), dtype=np.uint16)
            orig[: len(original_sample)] = original_sample
            self._ck(self.L.mpgpu_load_replicates2(self.h, boot.shape[0], _p(boot), boot.shape[1], _p(seg), len(seg), _p(orig)))

    def reps_info(self):
        g, e, t = C.c_int(), C.c_int(), C.c_int()
        self._ck(self.L.mpgpu_reps_info(self.h, C.byref(g), C.byref(e), C.byref(t)))
        return g.value, e.value, t.value

    def reps_timing(self):
        ms, rows, pat, sp = C.c_float(), C.c_int(), C.c_int(), C.c_int()
        self._ck(self.L.mpgpu_reps_timing(self.h, C.byref(ms), C.byref(rows), C.byref(pat), C.byref(sp)))
        return ms.value, rows.value, pat.value, sp.value

    def set_replicate_shards(self, rank, count):
        """multi-GPU -bb: this context keeps replicates [B*rank/count, B*(rank+1)/count) of what load_replicates is given"""
        self._ck(self.L.mpgpu_set_replicate_shards(self.h, int(rank), int(count)))
        self.rep_rank, self.rep_count = int(rank), int(count)

    def int8_peak(self, iters=4096):
        """measured tcgen05.mma kind::i8 issue rate of this device, int8 TOP/s"""
        t = C.c_double()
        self._ck(self.L.mpgpu_int8_peak(self.h, int(iters), C.byref(t)))
        return t.value

    def reps_current_tree(self):
        out = np.zeros(self.B, dtype=np.int32)
        self._ck(self.L.mpgpu_reps_current_tree(self.h, _p(out)))
        return out

    def reps_candidates(self, cand_idx):
        idx = np.ascontiguousarray(cand_idx, dtype=np.int32)
        out = np.zeros((len(idx), self.B), dtype=np.int32)
        self._ck(self.L.mpgpu_reps_candidates(self.h, _p(idx), len(idx), _p(out)))
        return out

    def reps_candidates_device(self, cand_idx):
        """Asynchronous: returns (device pointer, pitch) of the int32 [m][pitch] result."""
        idx = np.ascontiguousarray(cand_idx, dtype=np.int32)
        self._keep = idx
        ptr, pitch = C.c_void_p(), C.c_int()
        self._ck(self.L.mpgpu_reps_candidates_device(self.h, _p(idx), len(idx), C.byref(ptr), C.byref(pitch)))
        return ptr.value, pitch.value

    def set_remain_bounds(self, bounds):
        """boot_samples_pars_remain_bounds [B][nseg-1] (IQTree::pllComputeRellRemainBound) or None"""
        if bounds is None:
            self._ck(self.L.mpgpu_set_remain_bounds(self.h, None, 0))
            return
        b = np.ascontiguousarray(bounds, dtype=np.int32)
        assert b.ndim == 2 and b.shape[0] == self.B
        self._ck(self.L.mpgpu_set_remain_bounds(self.h, _p(b), b.shape[1]))

    def reps_prefix_max(self, cand_idx, samples):
        """max over the tested segments of (prefix of 16-bit segment sums + remain bound) for one candidate of the last scan
        batch (-1 = the current tree) and the listed replicates: the left side of the skip test of iqtree.cpp:3433-3445"""
        sm = np.ascontiguousarray(samples, dtype=np.int32)
        out = np.zeros(len(sm), dtype=np.int32)
        self._ck(self.L.mpgpu_reps_prefix_max(self.h, int(cand_idx), _p(sm), len(sm), _p(out)))
        return out

    def optimize_spr_bb(self, bn, bs, hooks, boot_logl, boot_counts, boot_trees, logl_cutoff=0.0, eps=0.5,
                        mintrav=1, maxtrav=6, ratchet_pattern_pars=None, mulhits=False, topboot=0, distinct=0, cur_it=1,
                        boot_threshold=None, updates_off=False, boot_tree_orig_logl=None):
        """pllOptimizeSprParsimony + saveCurrentTree (default policy).  hooks: BBHooks (e.g.
        Treels.hooks(rng)); boot_* arrays are updated in place.  Returns (startMP, back_node,
        back_slot, insertions scored, saveCurrentTree calls, REPS vectors used)."""
        bn = np.array(bn, dtype=np.int32, copy=True); bs = np.array(bs, dtype=np.int32, copy=True)
        assert boot_logl.dtype == np.float64 and boot_counts.dtype == np.int32 and boot_trees.dtype == np.int32
        st = BBState(len(boot_logl), boot_logl.ctypes.data, boot_counts.ctypes.data, boot_trees.ctypes.data,
                     float(logl_cutoff), float(eps), 0, 0, 0, None, 0, 1 if mulhits else 0, 0, None, None, int(cur_it),
                     1 if updates_off else 0, None if boot_tree_orig_logl is None else boot_tree_orig_logl.ctypes.data)
        if distinct:                                    # -distinct_iter_top_boot K: boot_threshold [B] int32 is the caller's, in/out
            assert boot_threshold is not None and boot_threshold.dtype == np.int32 and len(boot_threshold) == len(boot_logl)
            st.policy = 3; st.top_n = int(distinct); st.boot_threshold = boot_threshold.ctypes.data
        if topboot:                                     # -mulhits -topboot N
            self._top_count = np.zeros(len(boot_logl), dtype=np.int32)
            self._boot_threshold = np.full(len(boot_logl), -(2 ** 31 - 1), dtype=np.int32)
            st.policy = 2; st.top_n = int(topboot)
            st.top_count = self._top_count.ctypes.data; st.boot_threshold = self._boot_threshold.ctypes.data
        if ratchet_pattern_pars is not None:            # ratchet iteration (iqtree.cpp:3283-3294)
            rp = np.zeros(max(self.P, len(ratchet_pattern_pars)), dtype=np.uint16)
            rp[: len(ratchet_pattern_pars)] = ratchet_pattern_pars
            st.ratchet = 1; st.ratchet_pattern_pars = rp.ctypes.data
        self.last_bb_state = st
        best = C.c_uint32(); nins = C.c_int64()
        self._ck(self.L.mpgpu_optimize_spr_bb(self.h, _p(bn), _p(bs), mintrav, maxtrav, C.byref(hooks), C.byref(st),
                                              C.byref(best), C.byref(nins)))
        return best.value, bn, bs, nins.value, st.n_calls, st.n_reps

    def synchronize(self):
        self._ck(self.L.mpgpu_synchronize(self.h))

    def launch_count(self):
        return self.L.mpgpu_launch_count(self.h)

    def stream(self):
        return self.L.mpgpu_stream(self.h)
