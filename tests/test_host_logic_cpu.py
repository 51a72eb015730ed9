"""CPU suite, host logic: the tree-walking half of the library (visit order, candidate enumeration, applying a move) runs
without a device through the mpgpu_host_* entry points and must reproduce the reference's order exactly -- checked against
the golden vectors (produced by the reference itself) and, call by call, against the C oracle's record of which (pruned
ref, insertion ref) pair every saveCurrentTree / testInsertParsimony call saw."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

from mpboot_b200 import engine
from oracle import portlib
from tests.helpers import make_case

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASE_FILES = sorted(f for f in glob.glob(os.path.join(GOLD, "*.npz"))
                    if not f.endswith("tables.npz") and not os.path.basename(f).startswith(("sankoff_", "mulhits", "skbb")))
IDS = [os.path.basename(p)[:-4] for p in CASE_FILES]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def host_visit_order(n, bn, bs):
    L = engine.lib()
    L.mpgpu_host_visit_order.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    bn = np.ascontiguousarray(bn, dtype=np.int32); bs = np.ascontiguousarray(bs, dtype=np.int32)
    out = np.zeros(2 * n - 1, dtype=np.int32)
    assert L.mpgpu_host_visit_order(n, _p(bn), _p(bs), _p(out)) == 0, L.mpgpu_last_error()
    return out


def host_enumerate(n, bn, bs, order, first, count, mintrav, maxtrav):
    L = engine.lib()
    L.mpgpu_host_enumerate.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    bn = np.ascontiguousarray(bn, dtype=np.int32); bs = np.ascontiguousarray(bs, dtype=np.int32)
    order = np.ascontiguousarray(order, dtype=np.int32)
    cap = count * (8 << min(maxtrav, 10)) + 16
    vb = np.zeros(count + 1, dtype=np.int32); cr = np.zeros(cap, dtype=np.int32); cp = np.zeros(cap, dtype=np.int32)
    nc = C.c_int()
    assert L.mpgpu_host_enumerate(n, _p(bn), _p(bs), _p(order), first, count, mintrav, maxtrav, _p(vb), _p(cr), _p(cp), cap,
                                  C.byref(nc)) == 0, L.mpgpu_last_error()
    return vb, cr[: nc.value], cp[: nc.value]


def host_apply_spr(n, bn, bs, remove_ref, insert_ref):
    L = engine.lib()
    L.mpgpu_host_apply_spr.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    bn = np.array(bn, dtype=np.int32, copy=True); bs = np.array(bs, dtype=np.int32, copy=True)
    assert L.mpgpu_host_apply_spr(n, _p(bn), _p(bs), int(remove_ref), int(insert_ref)) == 0, L.mpgpu_last_error()
    return bn, bs


@pytest.mark.parametrize("path", CASE_FILES, ids=IDS)
def test_visit_order_and_candidate_counts_match_golden(path):
    g = dict(np.load(path))
    n, mt = int(g["n"]), int(g["maxtrav"])
    order = host_visit_order(n, g["bn"], g["bs"])
    assert np.array_equal(order[1:], g["order"][1:])                          # nodeRectifierPars
    vb, cref, cprune = host_enumerate(n, g["bn"], g["bs"], order, 1, 2 * n - 2, 1, mt)
    assert np.array_equal(vb, g["visit_begin"]) and len(cref) == len(g["visit_mp"])


@pytest.mark.parametrize("n,L,dt,seed,mt", [(12, 200, 1, 1, 6), (33, 300, 1, 2, 6), (64, 200, 2, 3, 4), (20, 100, 0, 4, 3),
                                              (100, 150, 1, 5, 6), (7, 120, 1, 6, 6), (4, 80, 1, 7, 6), (48, 100, 6, 8, 12)])
def test_enumeration_matches_oracle_call_by_call(n, L, dt, seed, mt):
    """Every candidate of every node visit: same pruned ref and insertion ref, in the same order, as the oracle's
    rearrangeParsimony / addTraverseParsimony restatement saw them; also in pieces (batching must not matter)."""
    c = make_case(n, L, dt, seed, mu=0.2)
    o = portlib.OracleEngine(c["codes"], c["weights"], dt)
    o.set_ring(c["bn"], c["bs"]); o.allocate(True)
    s0 = o.evaluate_full(True)
    order = host_visit_order(n, c["bn"], c["bs"])
    rn, rs = o.get_nodep()
    assert np.array_equal((3 * rn + rs)[1:], order[1:])
    vb, cref, cprune = host_enumerate(n, c["bn"], c["bs"], order, 1, 2 * n - 2, 1, mt)
    for i in range(1, 2 * n - 1):
        o.record(False)
        o.rearrange(i, 1, mt, True, s0)
        refs = o.saved_refs()
        assert tuple(refs[0]) == (0, 0)                                        # the current tree first (:2286)
        assert np.array_equal(refs[1:, 0], cprune[vb[i - 1]: vb[i]]), i
        assert np.array_equal(refs[1:, 1], cref[vb[i - 1]: vb[i]]), i
    for first, count in ((1, 1), (2, 3), (n, n - 2), (2 * n - 2, 1)):
        if first + count > 2 * n - 1 or count < 1:
            continue
        vb1, cr1, cp1 = host_enumerate(n, c["bn"], c["bs"], order, first, count, 1, mt)
        a, b = vb[first - 1], vb[first - 1 + count]
        assert np.array_equal(vb1, vb[first - 1: first + count] - a)
        assert np.array_equal(cr1, cref[a:b]) and np.array_equal(cp1, cprune[a:b])


@pytest.mark.parametrize("n,seed", [(10, 11), (40, 12), (90, 13)])
def test_apply_move_matches_oracle(n, seed):
    """restoreTreeRearrangeParsimony (:2379): after applying the oracle's chosen move the ring tables agree, over a chain of moves."""
    c = make_case(n, 300, 1, seed, mu=0.15)
    o = portlib.OracleEngine(c["codes"], c["weights"], 1)
    o.set_ring(c["bn"], c["bs"]); o.allocate(False)
    s0 = o.evaluate_full(False)
    bn, bs = np.array(c["bn"], dtype=np.int32), np.array(c["bs"], dtype=np.int32)
    portlib.seed_rng(3)
    moved = 0
    for i in range(1, 2 * n - 1):
        rc, out = o.rearrange(i, 1, 5, False, s0)
        rem, ins = 3 * int(out[1]) + int(out[2]), 3 * int(out[3]) + int(out[4])
        if not out[1] or not out[3]:
            continue
        o.apply_move(False)
        bn, bs = host_apply_spr(n, bn, bs, rem, ins)
        wbn, wbs = o.get_ring()
        assert np.array_equal(bn[3:], wbn[3:]) and np.array_equal(bs[3:], wbs[3:]), i
        s0 = o.evaluate_full(False)
        moved += 1
        if moved >= 12:
            break
    assert moved >= 3


@pytest.mark.parametrize("n,seed", [(40, 3), (200, 5), (777, 9)])
def test_parallel_enumeration_is_the_sequential_plan(n, seed):
    """Large batches are enumerated by several host threads (ScanPlanner::add_parallel): the scan program -- tasks, view offsets,
    control words, candidate and visit tables -- must be the same bytes as the single-threaded one, for any thread count, piece
    count and visit range; called repeatedly (the pool's workers sleep and wake) and from a forked child (a pool of its own)."""
    import os
    from mpboot_b200 import synth
    L = engine.lib()
    L.mpgpu_last_error.restype = C.c_char_p
    L.mpgpu_host_plan_selftest.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 6
    bn, bs = synth.random_tree_rings(n, np.random.default_rng(seed))
    bn = np.ascontiguousarray(bn, dtype=np.int32); bs = np.ascontiguousarray(bs, dtype=np.int32)
    order = host_visit_order(n, bn, bs)
    for rep in range(3):
        for nthreads in (1, 2, 3, 4, 8):
            for pieces in (1, 2, 3):
                for first, count, mt in ((1, 2 * n - 2, 6), (n // 2, n, 3), (n + 1, n - 2, 8), (1, 2 * n - 2, 1)):
                    rc = L.mpgpu_host_plan_selftest(n, _p(bn), _p(bs), _p(order), first, count, 1, mt, nthreads, pieces)
                    assert rc == 0, (L.mpgpu_last_error().decode(), nthreads, pieces, first, count, mt)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", DeprecationWarning)      # (forking with the pool's helper threads around is the point)
        pid = os.fork()
    if pid == 0:
        rc = L.mpgpu_host_plan_selftest(n, _p(bn), _p(bs), _p(order), 1, 2 * n - 2, 1, 6, 4, 2)
        os._exit(0 if rc == 0 else 1)
    assert os.waitpid(pid, 0)[1] == 0


def test_disthit_hook_is_the_reference_list_update():
    """The -distinct_iter_top_boot list update (iqtree.cpp:3624-3677) as the library's treels container implements it (the `disthit`
    hook mpgpu_optimize_spr_bb calls on every accepted tree), against a line-by-line restatement on random call sequences: tree
    already listed -> nothing; this iteration has an entry -> the better one stays; room left -> append; full -> the worst goes;
    the threshold is the list's minimum."""
    from mpboot_b200.engine import Treels, BBHooks
    L = engine.lib()
    rng = np.random.default_rng(77)
    for K in (1, 2, 3, 5):
        tl = Treels(8)
        hk = tl.hooks(C.cast(L.mpgpu_splitmix64_double, C.c_void_p).value)
        fn = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32)(hk.disthit)
        B = 6
        top = [[] for _ in range(B)]; it_of = [[] for _ in range(B)]; thr = [-(2 ** 31 - 1)] * B
        for step in range(4000):
            b = int(rng.integers(B)); tree = int(rng.integers(40)); cur_it = 1 + step // 400
            # only calls the policy accepts reach the hook: rell >= threshold
            rell = int(max(thr[b], -300) + rng.integers(0, 4)) if thr[b] > -(2 ** 31 - 1) else int(-300 + rng.integers(0, 20))
            t = min(K, len(it_of[b]))
            new_thr = thr[b]
            if not any(top[b][c][0] == tree for c in range(t)):
                c = 0
                while c < t:
                    if it_of[b][c] == cur_it:
                        if rell > top[b][c][1]:
                            top[b][c] = [tree, rell]
                        break
                    c += 1
                if c == t and t < K:
                    it_of[b].append(cur_it); top[b].append([tree, rell])
                elif c == t and t == K:
                    worst = min(range(t), key=lambda d: (top[b][d][1], d))
                    top[b][worst] = [tree, rell]; it_of[b][worst] = cur_it
                new_thr = min(e[1] for e in top[b])
            got = fn(hk.user, b, tree, rell, cur_it, K, thr[b])
            assert got == new_thr, (K, step)
            thr[b] = new_thr
        sizes, flat = tl.toplists(B)
        assert list(sizes) == [len(x) for x in top]
        assert flat.tolist() == [e for x in top for e in x]
        assert tl.topiters(B).tolist() == [i for x in it_of for i in x]
