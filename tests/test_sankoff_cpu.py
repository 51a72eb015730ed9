"""CPU suite: the -cost (Sankoff, weighted parsimony) path, SURVEY 8a row R11.  The plain-C oracle
(oracle/mp_oracle.c: compressSankoffDNA :2637, newviewSankoff...SIMD :477, evaluateSankoff...SIMD :880 with
its per-segment u16 wrap and lower-bound early termination, findMstScore parstree.cpp:606) against the golden
vectors tools/make_golden_sankoff.py produced by running the reference itself, and -- when oracle/_ref is
present -- against the reference live on further seeded cases."""
import glob
import os

import numpy as np
import pytest

from oracle import portlib, reflib
from tests.helpers import make_case

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SANKOFF_FILES = sorted(glob.glob(os.path.join(GOLD, "sankoff_*.npz")))


def load(path):
    return dict(np.load(path))


def port_for(g):
    ora = portlib.OracleEngine(g["codes"], g["weights"], int(g["datatype"]))
    assert ora.set_cost_matrix(g["cost"], g["seg"]) == int(g["highest"])
    ora.set_ring(g["bn"], g["bs"])
    return ora


def test_sankoff_golden_present():
    assert len(SANKOFF_FILES) >= 5


@pytest.mark.parametrize("path", SANKOFF_FILES, ids=[os.path.basename(p)[:-4] for p in SANKOFF_FILES])
def test_port_sankoff_matches_golden(path):
    g = load(path)
    n, maxtrav, ninf = int(g["n"]), int(g["maxtrav"]), int(g["n_inf"])
    ora = port_for(g)
    assert ora.allocate(per_site=False) == int(g["L"])
    for t in range(1, n + 1):
        assert np.array_equal(ora.sankoff_vect(t), g["tips"][t - 1])
    assert ora.evaluate_full(per_site=False) == int(g["score"])
    assert np.array_equal(ora.remainder_bounds(), g["lower_bounds"])
    for k, node in enumerate(range(n + 1, 2 * n - 1)):
        assert np.array_equal(ora.sankoff_vect(node), g["node_vect"][k]), node
        assert ora.node_score(node) == int(g["node_score"][k])
    rn, rs = ora.get_nodep()
    assert np.array_equal((3 * rn + rs)[1:], g["order"][1:])
    # plain sweep: early termination on, decisions of every visit
    portlib.seed_rng(31337)
    for i in range(1, 2 * n - 1):
        rc, out = ora.rearrange(i, 1, maxtrav, False, int(g["score"]))
        assert np.array_equal(out, g["plain_visit_out"][i - 1]), i
    assert portlib.rng_draws() == int(g["plain_draws"])
    # per-pattern mode: exact score of every insertion
    ora.allocate(per_site=True)
    assert ora.evaluate_full(per_site=True) == int(g["score"])
    pp, sm = ora.pattern_parsimony(ninf)
    assert sm == int(g["ptn_sum"]) and np.array_equal(pp, g["ptn_pars"])
    portlib.seed_rng(31337)
    vb = g["visit_begin"]
    for i in range(1, 2 * n - 1):
        ora.record(False)
        rc, out = ora.rearrange(i, 1, maxtrav, True, int(g["score"]))
        assert np.array_equal(ora.saved()[1:], g["visit_mp"][vb[i - 1]: vb[i]]), i
        assert np.array_equal(out, g["visit_out"][i - 1])
    for tag, bb in (("plain", False), ("bb", True)):
        portlib.seed_rng(2024)
        ora.set_ring(g["bn"], g["bs"])
        ora.allocate(bb)
        ora.record(False)
        assert ora.optimize_spr(1, maxtrav, bb=bb) == int(g["opt_%s_ret" % tag])
        assert portlib.rng_draws() == int(g["opt_%s_draws" % tag])
        bn, bs = ora.get_ring()
        assert np.array_equal(bn[3:], g["opt_%s_bn" % tag][3:]) and np.array_equal(bs[3:], g["opt_%s_bs" % tag][3:])
        if bb:
            assert np.array_equal(ora.saved(), g["opt_bb_saved"])
    portlib.seed_rng(77)
    assert ora.ras(int(g["ras_seed"]), maxtrav) == int(g["ras_ret"])
    assert portlib.rng_draws() == int(g["ras_draws"])
    bn, bs = ora.get_ring()
    assert np.array_equal(bn[3:], g["ras_bn"][3:]) and np.array_equal(bs[3:], g["ras_bs"][3:])


def test_early_termination_is_exercised():
    """At least one golden sweep must contain visits whose reported best differs between the plain mode (lower
    bound early exit, :951-956) and the exact mode -- otherwise the fixtures would not pin that branch."""
    hit = 0
    for path in SANKOFF_FILES:
        g = load(path)
        if len(g["lower_bounds"]):
            hit += 1
    assert hit >= 3


def test_uniform_cost_equals_fitch():
    """With unit costs Sankoff and Fitch give the same tree length (sanity of both restatements)."""
    c = make_case(14, 400, 1, 12)
    ora = portlib.OracleEngine(c["codes"], c["weights"], 1)
    ora.set_ring(c["bn"], c["bs"])
    ora.allocate(False)
    fitch = ora.evaluate_full(False)
    ora.set_cost_matrix((1 - np.eye(4)).astype(np.uint32), np.array([c["n_inf"]], dtype=np.int32))
    ora.set_ring(c["bn"], c["bs"])
    ora.allocate(False)
    assert ora.evaluate_full(False) == fitch


@pytest.mark.skipif(not reflib.available(), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("n,L,dt,seed", [(18, 500, 1, 31), (11, 250, 2, 32), (13, 260, 6, 33)])
def test_port_sankoff_matches_reference_live(n, L, dt, seed):
    c = make_case(n, L, dt, seed)
    S = {0: 2, 1: 4, 2: 20, 6: 32}[dt]
    rng = np.random.default_rng(seed)
    cost = rng.integers(1, 6, size=(S, S)); cost = np.minimum(cost, cost.T); np.fill_diagonal(cost, 0)
    ninf = c["n_inf"]
    seg = np.array([s for s in (48, 112, 176) if s < ninf] + [ninf], dtype=np.int32)
    ref = reflib.RefEngine(c["chars"], c["weights"], dt, n_informative=ninf)
    ora = portlib.OracleEngine(c["codes"], c["weights"], dt)
    assert ref.set_cost_matrix(cost, seg) == ora.set_cost_matrix(cost, seg)
    for bb in (False, True):
        for e in (ref, ora):
            e.set_ring(c["bn"], c["bs"]); e.allocate(bb)
        reflib.lib().mpref_seed_rng(9); portlib.seed_rng(9)
        assert ref.optimize_spr(1, 6, bb=bb) == ora.optimize_spr(1, 6, bb=bb)
        assert reflib.lib().mpref_rng_draws() == portlib.rng_draws()
        assert np.array_equal(ref.get_ring()[0][3:], ora.get_ring()[0][3:])
