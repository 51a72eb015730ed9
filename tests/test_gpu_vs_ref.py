"""GPU parity against the reference's own engine (oracle/_ref/libmpref.so, built from
/root/reference by oracle/Makefile and shipped to the GPU box as a prebuilt .so)."""
import numpy as np
import pytest

from tests.helpers import make_case

pytestmark = pytest.mark.gpu

CASES = [
    (12, 300, 1, 7),       # tiny DNA
    (40, 2000, 1, 11),     # DNA
    (30, 600, 2, 5),       # AA
    (25, 500, 6, 9),       # 32-state
    (20, 400, 0, 3),       # binary
]


def _engines(case):
    from mpboot_b200.engine import Engine
    from oracle import reflib
    if not reflib.available():
        pytest.skip("oracle/_ref not built")
    ref = reflib.RefEngine(case["chars"], case["weights"], case["datatype"], n_informative=case["n_inf"])
    ref.set_ring(case["bn"], case["bs"])
    eng = Engine()
    eng.load_alignment(case["codes"], case["weights"], case["datatype"])
    eng.set_tree(case["bn"], case["bs"])
    return ref, eng


@pytest.mark.parametrize("n,L,dt,seed", CASES)
def test_tip_planes_and_score(n, L, dt, seed):
    case = make_case(n, L, dt, seed)
    ref, eng = _engines(case)
    W = ref.allocate(per_site=True)
    assert W == eng.ref_words
    assert ref.num_informative() == eng.n_inf
    for tip in (1, 2, n // 2, n):
        assert np.array_equal(ref.parsvect(tip), eng.tip_planes(tip))
    s_ref = ref.evaluate_full(per_site=True)
    assert eng.tree_score() == s_ref
    pp_ref, sum_ref = ref.pattern_parsimony(case["n_inf"])
    pp, sm = eng.pattern_parsimony()
    assert sm == sum_ref == s_ref
    assert np.array_equal(pp[: case["n_inf"]], pp_ref)


@pytest.mark.parametrize("n,L,dt,seed", CASES)
def test_scan_matches_reference_visit_by_visit(n, L, dt, seed):
    case = make_case(n, L, dt, seed)
    ref, eng = _engines(case)
    ref.allocate(per_site=True)
    s0 = ref.evaluate_full(per_site=True)
    rn, rs = ref.get_nodep()
    order = eng.visit_order()
    assert np.array_equal(order[1:], (3 * rn + rs)[1:])
    vb, mp, cref, cprune = eng.scan_visits(order, 1, 2 * n - 2, 1, 6)
    for i in range(1, 2 * n - 1):
        ref.record(False)
        rc, out = ref.rearrange(i, 1, 6, True, s0)
        saved = ref.saved()
        assert saved[0] == s0                      # evaluateParsimony(p) at :2285
        mine = mp[vb[i - 1]: vb[i]]
        assert np.array_equal(saved[1:], mine.astype(np.int32)), "visit %d" % i


@pytest.mark.parametrize("n,L,dt,seed", CASES[:3])
def test_optimize_spr_same_moves(n, L, dt, seed):
    import ctypes as C
    from oracle import reflib
    case = make_case(n, L, dt, seed)
    ref, eng = _engines(case)
    L_ = reflib.lib()
    L_.mpref_seed_rng(1234)
    r_ref = ref.optimize_spr(1, 6, bb=False)
    draws_ref = L_.mpref_rng_draws()
    bn_ref, bs_ref = ref.get_ring()
    L_.mpref_seed_rng(1234)
    fn = C.cast(L_.mpref_random_double, C.c_void_p).value
    r, bn, bs, nins = eng.optimize_spr(case["bn"], case["bs"], fn, 1, 6)
    assert r == r_ref
    assert L_.mpref_rng_draws() == draws_ref
    assert np.array_equal(bn[3:], bn_ref[3:]) and np.array_equal(bs[3:], bs_ref[3:])
