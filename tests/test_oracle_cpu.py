"""CPU suite, part 1: the plain-C oracle (oracle/mp_oracle.c) against the golden vectors that
tools/make_golden.py produced by running the reference itself, and -- when oracle/_ref is
present -- against the reference live on further seeded cases."""
import glob
import os

import numpy as np
import pytest

from mpboot_b200 import encoding
from oracle import portlib, reflib
from tests.helpers import make_case

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASE_FILES = sorted(f for f in glob.glob(os.path.join(GOLD, "*.npz"))
                    if not f.endswith("tables.npz") and not os.path.basename(f).startswith(("sankoff_", "mulhits", "skbb")))


def load(path):
    return dict(np.load(path))


def test_golden_files_present():
    assert len(CASE_FILES) >= 5 and os.path.exists(os.path.join(GOLD, "tables.npz"))


def test_char_maps_and_state_masks():
    t = load(os.path.join(GOLD, "tables.npz"))
    for dt in (0, 1, 2, 6):
        assert np.array_equal(encoding.CHAR_MAP[dt], t["char_map_%d" % dt])
        bv = t["bitvector_%d" % dt]
        assert portlib.lib().mporacle_undetermined(dt) == int(t["undetermined_%d" % dt])
        for code in range(len(bv)):
            assert portlib.code_mask(dt, code) == int(bv[code])


def test_reps_u16_wrap_semantics():
    t = load(os.path.join(GOLD, "tables.npz"))
    out = portlib.reps(t["reps_pars"], t["reps_boot"], t["reps_seg"])
    assert np.array_equal(out, t["reps_out"])
    out1 = portlib.reps(t["reps_pars"], t["reps_boot"], np.array([len(t["reps_pars"])], dtype=np.int32))
    assert np.array_equal(out1, t["reps_out_1seg"])
    # and the closed form the CUDA path uses: sum over segments of (exact sum mod 2^16)
    pars = t["reps_pars"].astype(np.int64); boot = t["reps_boot"].astype(np.int64)
    lo = 0; acc = np.zeros(boot.shape[0], dtype=np.int64)
    for up in t["reps_seg"]:
        acc += (boot[:, lo:up] * pars[lo:up]).sum(axis=1) % 65536
        lo = ((up + 15) // 16) * 16
    assert np.array_equal(acc, t["reps_out"])


def test_segments_rule():
    # doSegmenting (iqtree.cpp:3793): cut at multiples of 16 once the running sum exceeds 4095
    score = np.full(100, 30, dtype=np.int32); freq = np.full(100, 10, dtype=np.int32)
    seg = portlib.segments(score, freq, 90)
    assert list(seg) == [16, 32, 48, 64, 80, 96, 90]


@pytest.mark.parametrize("path", CASE_FILES, ids=[os.path.basename(p)[:-4] for p in CASE_FILES])
def test_port_matches_golden(path):
    g = load(path)
    n, dt, mt = int(g["n"]), int(g["datatype"]), int(g["maxtrav"])
    assert np.array_equal(encoding.encode(g["chars"], dt), g["codes"])
    o = portlib.OracleEngine(g["codes"], g["weights"], dt)
    o.set_ring(g["bn"], g["bs"])
    assert o.allocate(per_site=True) == int(g["W"])
    assert o.num_informative() == int(g["ref_n_inf"]) == int(g["n_inf"])
    for t in range(1, n + 1):
        assert np.array_equal(o.parsvect(t), g["tip_planes"][t - 1])
    assert o.evaluate_full(per_site=True) == int(g["score"])
    pp, sm = o.pattern_parsimony(int(g["n_inf"]))
    assert sm == int(g["ptn_sum"]) and np.array_equal(pp, g["ptn_pars"])
    on, os_ = o.get_nodep()
    assert np.array_equal((3 * on + os_)[1:], g["order"][1:])
    for i in range(int(g["n_inf"])):
        assert o.min_pars_pattern(i) == int(g["min_pars"][i])
    vb = g["visit_begin"]
    portlib.seed_rng(31337)
    for i in range(1, 2 * n - 1):
        o.record(i == 3)
        rc, out = o.rearrange(i, 1, mt, True, int(g["score"]))
        if i == 3:
            m, pt = o.saved(True)
            assert np.array_equal(pt[1:, : int(g["n_inf"])], g["visit3_ptn"][1:])
        else:
            m = o.saved()
        assert m[0] == int(g["score"])
        assert np.array_equal(m[1:], g["visit_mp"][vb[i - 1]: vb[i]])
        assert np.array_equal(out, g["visit_out"][i - 1])
    for tag, bb in (("plain", False), ("bb", True)):
        portlib.seed_rng(2024)
        o.set_ring(g["bn"], g["bs"])
        o.record(False)
        assert o.optimize_spr(1, mt, bb=bb) == int(g["opt_%s_ret" % tag])
        assert portlib.rng_draws() == int(g["opt_%s_draws" % tag])
        bn, bs = o.get_ring()
        assert np.array_equal(bn[3:], g["opt_%s_bn" % tag][3:]) and np.array_equal(bs[3:], g["opt_%s_bs" % tag][3:])
        if bb:
            assert np.array_equal(o.saved(), g["opt_bb_saved"])
    portlib.seed_rng(77)
    assert o.ras(4242 + {"dna12": 7, "dna40": 11, "aa24": 5, "morph20": 9, "bin16": 3, "dna30w": 21}[os.path.basename(path)[:-4]], mt) == int(g["ras_ret"])
    assert portlib.rng_draws() == int(g["ras_draws"])
    bn, bs = o.get_ring()
    assert np.array_equal(bn[3:], g["ras_bn"][3:]) and np.array_equal(bs[3:], g["ras_bs"][3:])


@pytest.mark.skipif(not reflib.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("n,L,dt,seed", [(15, 500, 1, 31), (35, 900, 2, 32), (18, 350, 6, 33), (50, 3000, 1, 34)])
def test_port_matches_reference_live(n, L, dt, seed):
    c = make_case(n, L, dt, seed)
    ref = reflib.RefEngine(c["chars"], c["weights"], dt, n_informative=c["n_inf"]); ref.set_ring(c["bn"], c["bs"])
    o = portlib.OracleEngine(c["codes"], c["weights"], dt); o.set_ring(c["bn"], c["bs"])
    assert ref.allocate(True) == o.allocate(True)
    s = ref.evaluate_full(True)
    assert s == o.evaluate_full(True)
    a, sa = ref.pattern_parsimony(c["n_inf"]); b, sb = o.pattern_parsimony(c["n_inf"])
    assert sa == sb == s and np.array_equal(a, b)
    reflib.lib().mpref_seed_rng(99); portlib.seed_rng(99)
    ref.record(False); o.record(False)
    assert ref.optimize_spr(1, 6, bb=True) == o.optimize_spr(1, 6, bb=True)
    assert reflib.lib().mpref_rng_draws() == portlib.rng_draws()
    assert np.array_equal(ref.saved(), o.saved())
    r1 = ref.get_ring(); r2 = o.get_ring()
    assert np.array_equal(r1[0][3:], r2[0][3:]) and np.array_equal(r1[1][3:], r2[1][3:])


@pytest.mark.skipif(not reflib.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("n,L,dt,seed,hi", [(20, 600, 1, 51, 6), (14, 300, 2, 52, 5), (12, 200, 6, 53, 4), (16, 300, 0, 54, 7)])
def test_port_matches_reference_on_asymmetric_cost_matrices(n, L, dt, seed, hi):
    """An asymmetric cost matrix makes Sankoff scores root-dependent; the port must follow the reference's own orientation
    (evaluateSankoff... :905-918, left = p->back) -- tree score and every insertion score of a sweep, against the reference live.
    This is what lets the GPU suite use the port as the oracle for asymmetric matrices."""
    c = make_case(n, L, dt, seed)
    S = {0: 2, 1: 4, 2: 20, 6: 32}[dt]
    rng = np.random.default_rng(seed)
    cost = rng.integers(1, hi, size=(S, S)); np.fill_diagonal(cost, 0)
    if np.array_equal(cost, cost.T):
        cost[0, 1] += 1
    assert not np.array_equal(cost, cost.T)
    ninf = c["n_inf"]
    seg = np.array([s for s in range(80, ninf, 80)] + [ninf], dtype=np.int32)
    got = []
    for o in (portlib.OracleEngine(c["codes"], c["weights"], dt), reflib.RefEngine(c["chars"], c["weights"], dt, n_informative=ninf)):
        o.set_cost_matrix(cost.astype(np.uint32), seg)
        o.set_ring(c["bn"], c["bs"]); o.allocate(per_site=True)
        s0 = o.evaluate_full(per_site=True)
        mps = []
        for i in range(1, 2 * n - 1):
            o.record(False); o.rearrange(i, 1, 5, True, s0)
            mps.append(o.saved()[1:].copy())
        got.append((s0, np.concatenate(mps)))
    assert got[0][0] == got[1][0] and np.array_equal(got[0][1], got[1][1])
