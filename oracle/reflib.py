"""TEST INFRASTRUCTURE: ctypes binding of oracle/_ref/libmpref.so (the reference's own
parsimony engine, see oracle/ref_driver.cpp).  Only tests/, the golden-vector generator and
bench.py's reference/cpu_baseline legs may import this module."""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "_ref", "libmpref.so")


def available():
    return os.path.exists(PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(PATH)
        L.mpref_create.restype = C.c_void_p
        L.mpref_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.mpref_destroy.argtypes = [C.c_void_p]
        L.mpref_random_double.restype = C.c_double
        L.mpref_random_double.argtypes = [C.c_void_p]
        L.mpref_seed_rng.argtypes = [C.c_uint64]
        L.mpref_rng_draws.restype = C.c_uint64
        L.mpref_char_map.argtypes = [C.c_int, C.c_void_p]
        L.mpref_bitvector.argtypes = [C.c_int, C.c_void_p, C.c_int]
        L.mpref_set_ring.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.mpref_get_ring.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.mpref_get_nodep.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.mpref_set_weights.argtypes = [C.c_void_p, C.c_void_p]
        L.mpref_allocate.argtypes = [C.c_void_p, C.c_int]
        L.mpref_num_informative.argtypes = [C.c_void_p]
        L.mpref_get_parsvect.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.mpref_node_score.restype = C.c_uint
        L.mpref_node_score.argtypes = [C.c_void_p, C.c_int]
        L.mpref_evaluate_full.restype = C.c_uint
        L.mpref_evaluate_full.argtypes = [C.c_void_p, C.c_int]
        L.mpref_evaluate_at.restype = C.c_uint
        L.mpref_evaluate_at.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.mpref_pattern_parsimony.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.mpref_min_pars_pattern.argtypes = [C.c_void_p, C.c_int]
        L.mpref_record.argtypes = [C.c_void_p, C.c_int]
        L.mpref_saved_count.argtypes = [C.c_void_p]
        L.mpref_saved_mp.argtypes = [C.c_void_p, C.c_void_p]
        L.mpref_saved_ptn.argtypes = [C.c_void_p, C.c_void_p]
        L.mpref_rearrange.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_void_p]
        L.mpref_apply_move.argtypes = [C.c_void_p, C.c_int]
        L.mpref_node_rectifier.argtypes = [C.c_void_p]
        L.mpref_optimize_spr.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.mpref_ras.restype = C.c_uint
        L.mpref_ras.argtypes = [C.c_void_p, C.c_long, C.c_int]
        L.mpref_reps.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.mpref_sweep_count_insertions.restype = C.c_ulong
        L.mpref_sweep_count_insertions.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.mpref_set_cost_matrix.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.mpref_get_sankoff_vect.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.mpref_set_best.argtypes = [C.c_void_p, C.c_uint]
        L.mpref_remainder_bounds.argtypes = [C.c_void_p, C.c_void_p]
        _boot_protos(L, "mpref")
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _boot_protos(L, pre):
    vp, i = C.c_void_p, C.c_int
    getattr(L, pre + "_tree_fingerprint").restype = C.c_uint64
    getattr(L, pre + "_tree_fingerprint").argtypes = [vp]
    getattr(L, pre + "_boot_init").argtypes = [vp, i, vp, i, vp, i, C.c_double, C.c_double, vp]
    getattr(L, pre + "_boot_free").argtypes = [vp]
    getattr(L, pre + "_boot_set_cutoff").argtypes = [vp, C.c_double]
    getattr(L, pre + "_boot_set_mulhits").argtypes = [vp, i]
    getattr(L, pre + "_boot_mulhits").argtypes = [vp, vp, vp, i]
    getattr(L, pre + "_boot_set_topboot").argtypes = [vp, i]
    getattr(L, pre + "_boot_toplists").argtypes = [vp, vp, vp, vp, i]
    getattr(L, pre + "_boot_set_distinct").argtypes = [vp, i, i]
    getattr(L, pre + "_boot_topiters").argtypes = [vp, vp, i]
    getattr(L, pre + "_boot_set_ratchet").argtypes = [vp, vp, vp]
    getattr(L, pre + "_boot_set_state").argtypes = [vp, vp, vp, vp]
    getattr(L, pre + "_boot_get_state").argtypes = [vp, vp, vp, vp]
    getattr(L, pre + "_boot_counters").restype = C.c_long
    getattr(L, pre + "_boot_counters").argtypes = [vp, vp]
    getattr(L, pre + "_boot_treels").argtypes = [vp, vp]
    getattr(L, pre + "_boot_nmat").argtypes = [vp]
    getattr(L, pre + "_boot_mats").argtypes = [vp, vp]


class BootMixin:
    """-bb bookkeeping of IQTree::saveCurrentTree (default policy) attached to an engine: every
    saveCurrentTree up-call of the search then runs cutoff filter, REPS and the per-replicate
    best-tree update.  `bound` = per-pattern lower bound for the skip test (reference driver:
    ras_pars_score, combined with pllCalcMinParsScorePattern inside; C port: the final
    min_unit_pars) or None to run without the skip test."""
    _pre = None

    def _f(self, name):
        return getattr(self._L(), self._pre + name)

    def boot_init(self, boot, segment_upper, logl_cutoff=0.0, eps=0.5, bound=None):
        boot = np.ascontiguousarray(boot, dtype=np.uint16)
        seg = np.ascontiguousarray(segment_upper, dtype=np.int32)
        self._B = boot.shape[0]
        b = None if bound is None else np.ascontiguousarray(bound, dtype=np.int32)
        self._f("_boot_init")(self.h, boot.shape[0], _p(boot), boot.shape[1], _p(seg), len(seg),
                              float(logl_cutoff), float(eps), None if b is None else _p(b))

    def boot_set_ratchet(self, original_sample, initial_ptn):
        """Ratchet-iteration semantics (iqtree.cpp:3283-3294): cur_logl of every call is recomputed on the
        original frequencies from the previous call's pattern vector (initial_ptn before the first)."""
        a = np.zeros(self.P, dtype=np.uint16); a[: len(original_sample)] = original_sample
        b = np.zeros(self.P, dtype=np.uint16); b[: len(initial_ptn)] = initial_ptn
        self._f("_boot_set_ratchet")(self.h, _p(a), _p(b))

    def boot_free(self):
        self._f("_boot_free")(self.h)

    def boot_set_cutoff(self, c):
        self._f("_boot_set_cutoff")(self.h, float(c))

    def boot_set_mulhits(self, on=True):
        """params->multiple_hits (-mulhits without -topboot, iqtree.cpp:3498-3531)"""
        self._f("_boot_set_mulhits")(self.h, int(on))

    def boot_set_topboot(self, n):
        """params->store_top_boot_trees with -mulhits (iqtree.cpp:3536-3583)"""
        self._f("_boot_set_topboot")(self.h, int(n))

    def boot_set_distinct(self, k, cur_it):
        """params->distinct_iter_top_boot = k (iqtree.cpp:3587-3685) for the search to come in iteration cur_it (IQTree::curIt);
        the lists and thresholds persist across calls with the same k"""
        self._f("_boot_set_distinct")(self.h, int(k), int(cur_it))

    def boot_topiters(self):
        """boot_trees_parsimony_top_iter, in the order of boot_toplists' pairs"""
        tot = self._f("_boot_topiters")(self.h, None, 0)
        flat = np.zeros(max(tot, 1), dtype=np.int32)
        self._f("_boot_topiters")(self.h, _p(flat), tot)
        return flat[:tot]

    def boot_toplists(self):
        """boot_trees_parsimony_top: (sizes[B], boot_threshold[B], (tree_index, rell) pairs in list order, concatenated)"""
        sizes = np.zeros(self._B, dtype=np.int32); thr = np.zeros(self._B, dtype=np.int32)
        tot = self._f("_boot_toplists")(self.h, _p(sizes), _p(thr), None, 0)
        flat = np.zeros((max(tot, 1), 2), dtype=np.int32)
        self._f("_boot_toplists")(self.h, _p(sizes), _p(thr), _p(flat), tot)
        return sizes, thr, flat[:tot]

    def boot_mulhits(self):
        """boot_trees_parsimony: (sizes[B], members of every set ascending, concatenated)"""
        sizes = np.zeros(self._B, dtype=np.int32)
        tot = self._f("_boot_mulhits")(self.h, _p(sizes), None, 0)
        flat = np.zeros(max(tot, 1), dtype=np.int32)
        self._f("_boot_mulhits")(self.h, _p(sizes), _p(flat), tot)
        return sizes, flat[:tot]

    def boot_set_state(self, boot_logl, boot_counts, boot_trees):
        a = np.ascontiguousarray(boot_logl, dtype=np.float64)
        b = np.ascontiguousarray(boot_counts, dtype=np.int32)
        c = np.ascontiguousarray(boot_trees, dtype=np.int32)
        self._f("_boot_set_state")(self.h, _p(a), _p(b), _p(c))

    def boot_state(self):
        a = np.zeros(self._B, dtype=np.float64); b = np.zeros(self._B, dtype=np.int32); c = np.zeros(self._B, dtype=np.int32)
        self._f("_boot_get_state")(self.h, _p(a), _p(b), _p(c))
        return a, b, c

    def boot_counters(self):
        """(saveCurrentTree calls, treels_logl size, REPS rows, skipped replicates, bad pattern sums)"""
        out = np.zeros(4, dtype=np.int64)
        bad = self._f("_boot_counters")(self.h, _p(out))
        return int(out[0]), int(out[1]), int(out[2]), int(out[3]), int(bad)

    def boot_treels(self):
        out = np.zeros(max(self.boot_counters()[1], 1), dtype=np.float64)
        self._f("_boot_treels")(self.h, _p(out))
        return out[: self.boot_counters()[1]]

    def boot_mats(self):
        """rows (call index, pruned ref, insertion ref, tree_index, topology fingerprint) of every tree
        materialised because it won a replicate (iqtree.cpp:3692-3708); refs are 0 for the reference
        driver (it sees only the tree)."""
        k = self._f("_boot_nmat")(self.h)
        out = np.zeros((max(k, 1), 5), dtype=np.int64)
        self._f("_boot_mats")(self.h, _p(out))
        return out[:k]

    def fingerprint(self):
        return int(self._f("_tree_fingerprint")(self.h))


def char_map(datatype):
    out = np.zeros(256, dtype=np.uint8)
    assert lib().mpref_char_map(datatype, _p(out)) == 0
    return out


def bitvector(datatype, ncodes):
    out = np.zeros(ncodes, dtype=np.uint32)
    und = lib().mpref_bitvector(datatype, _p(out), ncodes)
    return out, und


class RefEngine(BootMixin):
    """One reference pllInstance + partitionList over an ASCII pattern matrix."""
    _pre = "mpref"

    def _L(self):
        return lib()

    def __init__(self, chars, weights, datatype, sort_alignment=True, n_informative=None):
        chars = np.ascontiguousarray(chars, dtype=np.uint8)
        weights = np.ascontiguousarray(weights, dtype=np.int32)
        self.n, self.P = chars.shape
        self.datatype = datatype
        ninf = self.P if n_informative is None else n_informative
        self.h = lib().mpref_create(self.n, self.P, datatype, _p(chars), _p(weights), int(sort_alignment), ninf)
        assert self.h
        self.W = None
        self.S = {0: 2, 1: 4, 2: 20, 6: 32}[datatype]

    def close(self):
        if self.h:
            lib().mpref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_ring(self, bn, bs):
        bn = np.ascontiguousarray(bn, dtype=np.int32); bs = np.ascontiguousarray(bs, dtype=np.int32)
        lib().mpref_set_ring(self.h, _p(bn), _p(bs))

    def get_ring(self):
        bn = np.zeros(3 * (2 * self.n - 1), dtype=np.int32); bs = np.zeros_like(bn)
        lib().mpref_get_ring(self.h, _p(bn), _p(bs))
        return bn, bs

    def get_nodep(self):
        a = np.zeros(2 * self.n - 1, dtype=np.int32); b = np.zeros_like(a)
        lib().mpref_get_nodep(self.h, _p(a), _p(b))
        return a, b

    def set_weights(self, w):
        w = np.ascontiguousarray(w, dtype=np.int32)
        lib().mpref_set_weights(self.h, _p(w))

    def allocate(self, per_site=False):
        self.W = lib().mpref_allocate(self.h, int(per_site))
        return self.W

    def num_informative(self):
        return lib().mpref_num_informative(self.h)

    # ---- Sankoff (-cost): the engine switches on pllCostMatrix (sprparsimony.cpp:556, 967, 2830)
    def set_cost_matrix(self, cost, segment_upper):
        """cost [S][S] (None = back to Fitch), segment_upper as IQTree::doSegmenting; returns highest_cost."""
        if cost is None:
            return lib().mpref_set_cost_matrix(self.h, None, None, 0)
        c = np.ascontiguousarray(cost, dtype=np.uint32); seg = np.ascontiguousarray(segment_upper, dtype=np.int32)
        assert c.shape == (self.S, self.S)
        return lib().mpref_set_cost_matrix(self.h, _p(c), _p(seg), len(seg))

    def set_sankoff_short(self, on):
        """False = -short_off: 32-bit Sankoff vectors and segment sums.  Before set_cost_matrix."""
        lib().mpref_set_sankoff_short(self.h, 1 if on else 0)

    def sankoff_vect(self, node):
        """u16 [parsimonyLength][S] cost vector of one node (de-blocked from [len/16][S][16])."""
        raw = np.zeros(self.W * self.S, dtype=np.uint16)
        L = lib().mpref_get_sankoff_vect(self.h, node, _p(raw))
        return raw.reshape(L // 16, self.S, 16).transpose(0, 2, 1).reshape(L, self.S)

    def set_best(self, best):
        lib().mpref_set_best(self.h, int(best))

    def remainder_bounds(self):
        out = np.zeros(65536, dtype=np.uint32)
        k = lib().mpref_remainder_bounds(self.h, _p(out))
        return out[:k].copy()

    def parsvect(self, node):
        out = np.zeros((self.S, self.W), dtype=np.uint32)
        lib().mpref_get_parsvect(self.h, node, _p(out))
        return out

    def node_score(self, node):
        return lib().mpref_node_score(self.h, node)

    def evaluate_full(self, per_site=False):
        return lib().mpref_evaluate_full(self.h, int(per_site))

    def evaluate_at(self, node, slot, full=False, per_site=False):
        return lib().mpref_evaluate_at(self.h, node, slot, int(full), int(per_site))

    def pattern_parsimony(self, count=None):
        out = np.zeros(self.P + 16, dtype=np.uint16)
        s = C.c_int(0)
        lib().mpref_pattern_parsimony(self.h, _p(out), C.byref(s))
        return out[: (self.P if count is None else count)], s.value

    def min_pars_pattern(self, site):
        return lib().mpref_min_pars_pattern(self.h, site)

    def record(self, ptn=False):
        lib().mpref_record(self.h, int(ptn))

    def saved(self, ptn=False):
        k = lib().mpref_saved_count(self.h)
        mp = np.zeros(k, dtype=np.int32)
        lib().mpref_saved_mp(self.h, _p(mp))
        if not ptn:
            return mp
        pt = np.zeros((k, self.P), dtype=np.uint16)
        lib().mpref_saved_ptn(self.h, _p(pt))
        return mp, pt

    def rearrange(self, i, mintrav, maxtrav, per_site, best_in):
        out = np.zeros(6, dtype=np.uint32)
        rc = lib().mpref_rearrange(self.h, i, mintrav, maxtrav, int(per_site), int(best_in), _p(out))
        return rc, out

    def apply_move(self, per_site=False):
        lib().mpref_apply_move(self.h, int(per_site))

    def node_rectifier(self):
        lib().mpref_node_rectifier(self.h)

    def optimize_spr(self, mintrav=1, maxtrav=6, bb=False, ratchet_realloc=False):
        return lib().mpref_optimize_spr(self.h, mintrav, maxtrav, int(bb), int(ratchet_realloc))

    def ras(self, seed, spr_dist):
        return lib().mpref_ras(self.h, seed, spr_dist)

    def sweep_count(self, mintrav, maxtrav, per_site, reps=1):
        return lib().mpref_sweep_count_insertions(self.h, mintrav, maxtrav, int(per_site), reps)


def reps(pars, boot, segment_upper):
    """REPS on the reference's Vec16us (iqtree.cpp:3424-3449). pars u16[P], boot u16[B][P]."""
    P = len(pars)
    Pp = (P + 15) // 16 * 16 + 16
    B = boot.shape[0]
    a = np.zeros(Pp, dtype=np.uint16); a[:P] = pars
    w = np.zeros((B, Pp), dtype=np.uint16); w[:, :P] = boot
    seg = np.ascontiguousarray(segment_upper, dtype=np.int32)
    out = np.zeros(B, dtype=np.int32)
    lib().mpref_reps(_p(a), _p(w), B, Pp, _p(seg), len(seg), _p(out))
    return out
