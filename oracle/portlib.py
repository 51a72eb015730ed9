"""TEST INFRASTRUCTURE: ctypes binding of oracle/libmporacle.so (the plain-C restatement,
oracle/mp_oracle.c).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.  Same method names as oracle/reflib.RefEngine so
tests can run either against the same assertions."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "libmporacle.so")

_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, PATH])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(PATH):
            build()
        L = C.CDLL(PATH)
        vp, i = C.c_void_p, C.c_int
        L.mporacle_create.restype = vp
        L.mporacle_create.argtypes = [i, i, i, vp, vp, i]
        L.mporacle_destroy.argtypes = [vp]
        L.mporacle_set_weights.argtypes = [vp, vp]
        L.mporacle_allocate.argtypes = [vp, i]
        L.mporacle_num_informative.argtypes = [vp]
        L.mporacle_get_parsvect.argtypes = [vp, i, vp]
        L.mporacle_node_score.restype = C.c_uint
        L.mporacle_node_score.argtypes = [vp, i]
        L.mporacle_set_ring.argtypes = [vp, vp, vp]
        L.mporacle_get_ring.argtypes = [vp, vp, vp]
        L.mporacle_get_nodep.argtypes = [vp, vp]
        L.mporacle_node_rectifier.argtypes = [vp]
        L.mporacle_evaluate_full.restype = C.c_uint
        L.mporacle_evaluate_full.argtypes = [vp, i]
        L.mporacle_pattern_parsimony.argtypes = [vp, vp, vp]
        L.mporacle_min_pars_pattern.argtypes = [vp, i]
        L.mporacle_record.argtypes = [vp, i]
        L.mporacle_saved_count.argtypes = [vp]
        L.mporacle_saved_mp.argtypes = [vp, vp]
        L.mporacle_saved_ptn.argtypes = [vp, vp]
        L.mporacle_rearrange.argtypes = [vp, i, i, i, i, C.c_uint, vp]
        L.mporacle_apply_move.argtypes = [vp, i]
        L.mporacle_optimize_spr.argtypes = [vp, i, i, i]
        L.mporacle_ras.restype = C.c_uint
        L.mporacle_ras.argtypes = [vp, C.c_long, i]
        L.mporacle_sweep_count_insertions.restype = C.c_ulong
        L.mporacle_sweep_count_insertions.argtypes = [vp, i, i, i, i]
        L.mporacle_random_double.restype = C.c_double
        L.mporacle_random_double.argtypes = [vp]
        L.mporacle_seed_rng.argtypes = [C.c_uint64]
        L.mporacle_rng_draws.restype = C.c_uint64
        L.mporacle_code_mask.restype = C.c_uint32
        L.mporacle_code_mask.argtypes = [i, i]
        L.mporacle_undetermined.argtypes = [i]
        L.mporacle_reps.argtypes = [vp, vp, i, i, vp, i, vp]
        L.mporacle_segments.argtypes = [vp, vp, i, i, vp]
        L.mporacle_set_cost_matrix.argtypes = [vp, vp, vp, i]
        L.mporacle_get_sankoff_vect.argtypes = [vp, i, vp]
        L.mporacle_set_best.argtypes = [vp, C.c_uint]
        L.mporacle_remainder_bounds.argtypes = [vp, vp]
        from .reflib import _boot_protos
        _boot_protos(L, "mporacle")
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def seed_rng(seed):
    lib().mporacle_seed_rng(seed)


def rng_draws():
    return lib().mporacle_rng_draws()


def rng_fn_address():
    """Address of `double mporacle_random_double(void*)`, to hand to mpgpu_optimize_spr as the host RNG."""
    return C.cast(lib().mporacle_random_double, C.c_void_p).value


from .reflib import BootMixin  # noqa: E402  (pure-Python mixin, loads no library)


class OracleEngine(BootMixin):
    _pre = "mporacle"

    def _L(self):
        return lib()

    def __init__(self, codes, weights, datatype, sort_alignment=True):
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        weights = np.ascontiguousarray(weights, dtype=np.int32)
        self.n, self.P = codes.shape
        self.datatype = datatype
        self.S = {0: 2, 1: 4, 2: 20, 6: 32}[datatype]
        self.h = lib().mporacle_create(self.n, self.P, datatype, _p(codes), _p(weights), int(sort_alignment))
        assert self.h
        self.W = None

    def close(self):
        if self.h:
            lib().mporacle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_ring(self, bn, bs):
        bn = np.ascontiguousarray(bn, dtype=np.int32); bs = np.ascontiguousarray(bs, dtype=np.int32)
        lib().mporacle_set_ring(self.h, _p(bn), _p(bs))

    def get_ring(self):
        bn = np.zeros(3 * (2 * self.n - 1), dtype=np.int32); bs = np.zeros_like(bn)
        lib().mporacle_get_ring(self.h, _p(bn), _p(bs))
        return bn, bs

    def get_nodep(self):
        refs = np.zeros(2 * self.n - 1, dtype=np.int32)
        lib().mporacle_get_nodep(self.h, _p(refs))
        return refs // 3, refs % 3

    def set_weights(self, w):
        w = np.ascontiguousarray(w, dtype=np.int32)
        lib().mporacle_set_weights(self.h, _p(w))

    def allocate(self, per_site=False):
        self.W = lib().mporacle_allocate(self.h, int(per_site))
        return self.W

    # ---- Sankoff (-cost)
    def set_cost_matrix(self, cost, segment_upper):
        if cost is None:
            return lib().mporacle_set_cost_matrix(self.h, None, None, 0)
        c = np.ascontiguousarray(cost, dtype=np.uint32); seg = np.ascontiguousarray(segment_upper, dtype=np.int32)
        return lib().mporacle_set_cost_matrix(self.h, _p(c), _p(seg), len(seg))

    def sankoff_vect(self, node):
        out = np.zeros((self.W, self.S), dtype=np.uint16)
        lib().mporacle_get_sankoff_vect(self.h, node, _p(out))
        return out

    def set_best(self, best):
        lib().mporacle_set_best(self.h, int(best))

    def remainder_bounds(self):
        out = np.zeros(65536, dtype=np.uint32)
        k = lib().mporacle_remainder_bounds(self.h, _p(out))
        return out[:k].copy()

    def num_informative(self):
        return lib().mporacle_num_informative(self.h)

    def parsvect(self, node):
        out = np.zeros((self.S, self.W), dtype=np.uint32)
        lib().mporacle_get_parsvect(self.h, node, _p(out))
        return out

    def node_score(self, node):
        return lib().mporacle_node_score(self.h, node)

    def evaluate_full(self, per_site=False):
        return lib().mporacle_evaluate_full(self.h, int(per_site))

    def pattern_parsimony(self, count=None):
        out = np.zeros(self.P + 16, dtype=np.uint16)
        s = C.c_int(0)
        lib().mporacle_pattern_parsimony(self.h, _p(out), C.byref(s))
        return out[: (self.P if count is None else count)], s.value

    def min_pars_pattern(self, site):
        return lib().mporacle_min_pars_pattern(self.h, site)

    def record(self, ptn=False):
        lib().mporacle_record(self.h, int(ptn))

    def saved(self, ptn=False):
        k = lib().mporacle_saved_count(self.h)
        mp = np.zeros(k, dtype=np.int32)
        if k:
            lib().mporacle_saved_mp(self.h, _p(mp))
        if not ptn:
            return mp
        pt = np.zeros((k, self.P), dtype=np.uint16)
        if k:
            lib().mporacle_saved_ptn(self.h, _p(pt))
        return mp, pt

    def saved_refs(self):
        """(pruned ref, insertion ref) per recorded call (0, 0 = the current tree), ref = 3 * node + slot"""
        k = lib().mporacle_saved_count(self.h)
        out = np.zeros((max(k, 1), 2), dtype=np.int32)
        if k:
            lib().mporacle_saved_refs(self.h, _p(out))
        return out[:k]

    def rearrange(self, i, mintrav, maxtrav, per_site, best_in):
        out = np.zeros(6, dtype=np.uint32)
        rc = lib().mporacle_rearrange(self.h, i, mintrav, maxtrav, int(per_site), int(best_in), _p(out))
        return rc, out

    def apply_move(self, per_site=False):
        lib().mporacle_apply_move(self.h, int(per_site))

    def node_rectifier(self):
        lib().mporacle_node_rectifier(self.h)

    def optimize_spr(self, mintrav=1, maxtrav=6, bb=False):
        return lib().mporacle_optimize_spr(self.h, mintrav, maxtrav, int(bb))

    def ras(self, seed, spr_dist):
        return lib().mporacle_ras(self.h, seed, spr_dist)

    def sweep_count(self, mintrav, maxtrav, per_site, reps=1):
        return lib().mporacle_sweep_count_insertions(self.h, mintrav, maxtrav, int(per_site), reps)


def code_mask(datatype, code):
    return lib().mporacle_code_mask(datatype, code)


def reps(pars, boot, segment_upper):
    P = len(pars)
    Pp = (P + 15) // 16 * 16 + 16
    B = boot.shape[0]
    a = np.zeros(Pp, dtype=np.uint16); a[:P] = pars
    w = np.zeros((B, Pp), dtype=np.uint16); w[:, :P] = boot
    seg = np.ascontiguousarray(segment_upper, dtype=np.int32)
    out = np.zeros(B, dtype=np.int32)
    lib().mporacle_reps(_p(a), _p(w), B, Pp, _p(seg), len(seg), _p(out))
    return out


def segments(ras_score, freq, n_informative):
    ras_score = np.ascontiguousarray(ras_score, dtype=np.int32)
    freq = np.ascontiguousarray(freq, dtype=np.int32)
    out = np.zeros(len(freq) + 1, dtype=np.int32)
    k = lib().mporacle_segments(_p(ras_score), _p(freq), len(freq), n_informative, _p(out))
    return out[:k]
