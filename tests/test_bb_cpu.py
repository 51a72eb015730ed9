"""CPU suite for the -bb bookkeeping oracle: the C port's restatement of IQTree::saveCurrentTree
(default policy) against the reference driver's re-typed copy (which runs the reference's own
search, pattern-score and Vec16us code), on whole -bb SPR searches: same treels_logl, same
per-replicate best trees/scores/counts, same RNG draws, same materialised topologies."""
import numpy as np
import pytest

from oracle import portlib, reflib
from tests.helpers import make_case, make_boot, segments_for, fingerprint_ring

needs_ref = pytest.mark.skipif(not reflib.available(), reason="oracle/_ref not built (needs /root/reference)")


def bb_setup(n, L, dt, seed, B, mu=0.05, art_segments=True, heavy=True):
    c = make_case(n, L, dt, seed, mu=mu)
    o = portlib.OracleEngine(c["codes"], c["weights"], dt)
    o.set_ring(c["bn"], c["bs"]); o.allocate(True)
    s0 = o.evaluate_full(True)
    pp, _ = o.pattern_parsimony(c["n_inf"])
    ninf = c["n_inf"]
    if art_segments and ninf > 64:          # several segments (multiples of 16, last = n_inf)
        seg = np.array([16, 48, 16 * (ninf // 32), ninf], dtype=np.int32)
        seg = np.unique(seg)
    else:
        seg = segments_for(pp, c["weights"], ninf)
    hv = None
    if heavy:                               # u16 lane products / sums that wrap, weights > 255
        hv = [(1, 3, 40000), (2, 5, 300), (2, 6, 65535)] + [(3, k, 900) for k in range(0, min(ninf, 40))]
    boot = make_boot(c, B, seed, heavy=hv)
    minp = np.array([o.min_pars_pattern(i) for i in range(ninf)], dtype=np.int32)
    P = c["codes"].shape[1]
    ras = np.zeros(P, dtype=np.int32); ras[:ninf] = pp
    bound = np.zeros(P, dtype=np.int32); bound[:ninf] = np.minimum(minp, pp)
    return c, o, s0, pp, seg, boot, ras, bound


def ratchet_setup(c, pp, seed):
    """What a ratchet iteration changes (alignment.cpp:1940-1963): half of the informative patterns get
    frequency + 1 for the search; saveCurrentTree still scores cur_logl on the original frequencies."""
    rng = np.random.default_rng(700 + seed)
    w2 = c["weights"].copy()
    w2[: c["n_inf"]] += (rng.random(c["n_inf"]) < 0.5).astype(w2.dtype)
    P = len(w2)
    orig = np.zeros(P, dtype=np.uint16); orig[: c["n_inf"]] = c["weights"][: c["n_inf"]]
    init = np.zeros(P, dtype=np.uint16); init[: c["n_inf"]] = pp
    return w2, orig, init


def run_bb(eng, c, boot, seg, cutoff, bound, is_ref, seed=2024, mt=6, ratchet=None, mulhits=False, cost=None, topboot=0):
    eng.set_cost_matrix(cost, seg if cost is not None else None)      # -cost: pllCostMatrix + pllSegmentUpper (None = Fitch)
    if ratchet is not None:
        eng.set_weights(ratchet[0])
    else:
        eng.set_weights(c["weights"])
    eng.set_ring(c["bn"], c["bs"])
    eng.allocate(per_site=True)
    eng.boot_init(boot, seg, cutoff, 0.5, bound)
    if ratchet is not None:
        eng.boot_set_ratchet(ratchet[1], ratchet[2])
    if mulhits:
        eng.boot_set_mulhits(True)
    if topboot:
        eng.boot_set_topboot(topboot)
    (reflib.lib().mpref_seed_rng if is_ref else portlib.seed_rng)(seed)
    eng.record(False)
    ret = eng.optimize_spr(1, mt, bb=True)
    draws = reflib.lib().mpref_rng_draws() if is_ref else portlib.rng_draws()
    return dict(ret=ret, draws=draws, ring=eng.get_ring(), state=eng.boot_state(), counters=eng.boot_counters(),
                treels=eng.boot_treels(), mats=eng.boot_mats(), saved=eng.saved(), mulhits=eng.boot_mulhits(),
                toplists=eng.boot_toplists())


def same(a, b):
    assert a["ret"] == b["ret"] and a["draws"] == b["draws"]
    assert all(np.array_equal(x, y) for x, y in zip(a["mulhits"], b["mulhits"]))
    assert all(np.array_equal(x, y) for x, y in zip(a["toplists"], b["toplists"]))
    assert all(np.array_equal(x, y) for x, y in zip(a["ring"], b["ring"]))
    assert all(np.array_equal(x, y) for x, y in zip(a["state"], b["state"]))
    assert a["counters"] == b["counters"] and a["counters"][4] == 0
    assert np.array_equal(a["treels"], b["treels"]) and np.array_equal(a["saved"], b["saved"])
    assert np.array_equal(a["mats"][:, [0, 3, 4]], b["mats"][:, [0, 3, 4]])


@needs_ref
@pytest.mark.parametrize("n,L,dt,seed,B,mu", [(12, 300, 1, 7, 50, 0.05), (24, 400, 2, 5, 40, 0.05),
                                               (30, 800, 1, 21, 64, 0.01), (20, 300, 6, 9, 30, 0.05)])
def test_port_bb_equals_reference(n, L, dt, seed, B, mu):
    c, o, s0, pp, seg, boot, ras, bound = bb_setup(n, L, dt, seed, B, mu)
    r = reflib.RefEngine(c["chars"], c["weights"], dt, n_informative=c["n_inf"])
    plain = run_bb(o, c, boot, seg, 0.0, None, False)
    same(plain, run_bb(r, c, boot, seg, 0.0, None, True))
    cutoff = -(plain["ret"] + 4.0)                      # keeps roughly the best candidates only
    a = run_bb(o, c, boot, seg, cutoff, bound, False)
    same(a, run_bb(r, c, boot, seg, cutoff, ras, True))
    assert a["counters"][1] < a["counters"][0]          # the cutoff dropped candidates
    # the skip test is decision-neutral
    b = run_bb(o, c, boot, seg, cutoff, None, False)
    for k in ("ret", "draws"):
        assert a[k] == b[k]
    assert all(np.array_equal(x, y) for x, y in zip(a["state"], b["state"]))
    assert np.array_equal(a["mats"], b["mats"])


@needs_ref
@pytest.mark.parametrize("n,L,dt,seed,B,mu", [(12, 300, 1, 7, 50, 0.05), (24, 400, 2, 5, 40, 0.05), (30, 800, 1, 21, 64, 0.01)])
def test_port_bb_ratchet_iteration_equals_reference(n, L, dt, seed, B, mu):
    c, o, s0, pp, seg, boot, ras, bound = bb_setup(n, L, dt, seed, B, mu)
    r = reflib.RefEngine(c["chars"], c["weights"], dt, n_informative=c["n_inf"])
    rt = ratchet_setup(c, pp, seed)
    a = run_bb(o, c, boot, seg, 0.0, None, False, ratchet=rt)
    same(a, run_bb(r, c, boot, seg, 0.0, None, True, ratchet=rt))
    # with a cutoff the chain of stale scores breaks at some call and nothing passes afterwards
    top = np.unique(-a["treels"])[::-1]               # original-frequency scores of the unfiltered run, worst first
    for worst in top[:2]:                             # the chain breaks right after the first call that scores `worst`
        cutoff = -(worst - 0.5)
        x = run_bb(o, c, boot, seg, cutoff, None, False, ratchet=rt)
        same(x, run_bb(r, c, boot, seg, cutoff, None, True, ratchet=rt))
        assert x["counters"][1] < a["counters"][1]


MULHITS_CASES = [(12, 300, 1, 7, 50, 0.05), (24, 400, 2, 5, 40, 0.05), (30, 800, 1, 21, 64, 0.01), (20, 300, 6, 9, 30, 0.05)]
MULHITS_GOLD = __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "golden", "mulhits.npz")


@needs_ref
@pytest.mark.parametrize("n,L,dt,seed,B,mu", MULHITS_CASES)
def test_port_bb_mulhits_equals_reference(n, L, dt, seed, B, mu):
    """-mulhits (params->multiple_hits, iqtree.cpp:3498-3531): every tree tying a replicate's best score is kept in
    boot_trees_parsimony[sample]; no tie-break draws from saveCurrentTree."""
    c, o, s0, pp, seg, boot, ras, bound = bb_setup(n, L, dt, seed, B, mu)
    r = reflib.RefEngine(c["chars"], c["weights"], dt, n_informative=c["n_inf"])
    a = run_bb(o, c, boot, seg, 0.0, None, False, mulhits=True)
    same(a, run_bb(r, c, boot, seg, 0.0, None, True, mulhits=True))
    assert a["mulhits"][0].min() >= 1                                    # every replicate has at least one best tree
    assert dt == 6 or a["mulhits"][0].max() > 1                          # ties are kept (the 32-state case has none)
    d = run_bb(o, c, boot, seg, 0.0, None, False)
    assert a["draws"] < d["draws"]                                       # the default policy's tie-break draws are gone
    cutoff = -(a["ret"] + 4.0)
    x = run_bb(o, c, boot, seg, cutoff, bound, False, mulhits=True)
    same(x, run_bb(r, c, boot, seg, cutoff, ras, True, mulhits=True))
    y = run_bb(o, c, boot, seg, cutoff, None, False, mulhits=True)      # the skip test is decision-neutral here too
    assert all(np.array_equal(p, q) for p, q in zip(x["mulhits"], y["mulhits"])) and np.array_equal(x["state"][0], y["state"][0])


@needs_ref
@pytest.mark.parametrize("n,L,dt,seed,B,mu", MULHITS_CASES[:3])
@pytest.mark.parametrize("N", [1, 3, 10])
def test_port_bb_topboot_equals_reference(n, L, dt, seed, B, mu, N):
    """-mulhits -topboot N (store_top_boot_trees, iqtree.cpp:3536-3583): per replicate the N best newly seen trees"""
    c, o, s0, pp, seg, boot, ras, bound = bb_setup(n, L, dt, seed, B, mu)
    r = reflib.RefEngine(c["chars"], c["weights"], dt, n_informative=c["n_inf"])
    a = run_bb(o, c, boot, seg, 0.0, None, False, topboot=N)
    same(a, run_bb(r, c, boot, seg, 0.0, None, True, topboot=N))
    assert a["toplists"][0].max() == N
    cutoff = -(a["ret"] + 4.0)
    x = run_bb(o, c, boot, seg, cutoff, bound, False, topboot=N)
    same(x, run_bb(r, c, boot, seg, cutoff, ras, True, topboot=N))


def run_bb_distinct(eng, c, boot, seg, bound, is_ref, K, iters=3, seed=2024, mt=6, cutoff=0.0):
    """-distinct_iter_top_boot K over several iterations (IQTree::curIt = 1, 2, ...): one bookkeeping state, every iteration an SPR
    search from another random tree, as doTreeSearch drives pllOptimizeSprParsimony once per iteration"""
    from mpboot_b200 import synth
    eng.set_cost_matrix(None, None)
    eng.set_weights(c["weights"])
    eng.set_ring(c["bn"], c["bs"])
    eng.allocate(per_site=True)
    eng.boot_init(boot, seg, cutoff, 0.5, bound)
    (reflib.lib().mpref_seed_rng if is_ref else portlib.seed_rng)(seed)
    rets = []
    for it in range(1, iters + 1):
        bn, bs = (c["bn"], c["bs"]) if it == 1 else synth.random_tree_rings(c["n"], np.random.default_rng(5000 + it))
        eng.set_ring(bn, bs)
        eng.allocate(per_site=True)
        eng.boot_set_distinct(K, it)
        eng.record(False)
        rets.append(eng.optimize_spr(1, mt, bb=True))
    draws = reflib.lib().mpref_rng_draws() if is_ref else portlib.rng_draws()
    return dict(ret=rets, draws=draws, ring=eng.get_ring(), state=eng.boot_state(), counters=eng.boot_counters(),
                treels=eng.boot_treels(), mats=eng.boot_mats(), toplists=eng.boot_toplists(), topiters=eng.boot_topiters())


def same_distinct(a, b):
    assert a["ret"] == b["ret"] and a["draws"] == b["draws"]
    assert all(np.array_equal(x, y) for x, y in zip(a["toplists"], b["toplists"])) and np.array_equal(a["topiters"], b["topiters"])
    assert all(np.array_equal(x, y) for x, y in zip(a["ring"], b["ring"]))
    assert all(np.array_equal(x, y) for x, y in zip(a["state"], b["state"]))
    assert a["counters"] == b["counters"] and a["counters"][4] == 0
    assert np.array_equal(a["treels"], b["treels"])
    assert np.array_equal(a["mats"][:, [0, 3, 4]], b["mats"][:, [0, 3, 4]])


DISTINCT_CASES = [(12, 300, 1, 7, 50, 0.05), (24, 400, 2, 5, 40, 0.05), (30, 800, 1, 21, 64, 0.01)]


@needs_ref
@pytest.mark.parametrize("n,L,dt,seed,B,mu", DISTINCT_CASES)
@pytest.mark.parametrize("K", [1, 3])
def test_port_bb_distinct_iter_equals_reference(n, L, dt, seed, B, mu, K):
    """-distinct_iter_top_boot K (iqtree.cpp:3587-3685): per replicate up to K trees from distinct iterations, accepted against
    boot_threshold; with the remain bounds the skip test (against boot_logl) changes decisions, so it is compared too."""
    c, o, s0, pp, seg, boot, ras, bound = bb_setup(n, L, dt, seed, B, mu)
    r = reflib.RefEngine(c["chars"], c["weights"], dt, n_informative=c["n_inf"])
    a = run_bb_distinct(o, c, boot, seg, None, False, K)
    same_distinct(a, run_bb_distinct(r, c, boot, seg, None, True, K))
    assert a["toplists"][0].max() == min(K, 2) and a["topiters"].max() > 1     # lists of several entries, entries of later iterations
    x = run_bb_distinct(o, c, boot, seg, bound, False, K)
    same_distinct(x, run_bb_distinct(r, c, boot, seg, ras, True, K))
    assert x["counters"][3] > 0                                          # replicates were skipped ...
    if K == 3 and n == 24:                                               # ... and that is not decision-neutral under this policy
        assert not all(np.array_equal(p, q) for p, q in zip(x["state"], a["state"]))


def mulhits_golden_case(g, k):
    n, L, dt, seed, B, mu = [x for x in MULHITS_CASES[k]]
    c, o, s0, pp, seg, boot, ras, bound = bb_setup(n, L, dt, seed, B, mu)
    assert np.array_equal(boot, g["c%d_boot" % k]) and np.array_equal(c["codes"], g["c%d_codes" % k])   # same seeded inputs
    return c, o, seg, boot, bound


TOPBOOT_NS = (1, 4)


def check_mulhits_golden(g, k, tag, r):
    p = "c%d_%s_" % (k, tag)
    assert r["ret"] == int(g[p + "ret"]) and r["draws"] == int(g[p + "draws"])
    assert np.array_equal(r["ring"][0][3:], g[p + "bn"][3:]) and np.array_equal(r["ring"][1][3:], g[p + "bs"][3:])
    assert np.array_equal(r["state"][0], g[p + "boot_logl"])
    assert np.array_equal(r["mulhits"][0], g[p + "sizes"]) and np.array_equal(r["mulhits"][1], g[p + "flat"])
    assert np.array_equal(r["treels"], g[p + "treels"])
    assert np.array_equal(r["mats_tf"], g[p + "mats"])                 # tree_index, topology fingerprint per materialised tree
    if p + "top_sizes" in g:
        assert np.array_equal(r["toplists"][0], g[p + "top_sizes"]) and np.array_equal(r["toplists"][1], g[p + "top_thr"])
        assert np.array_equal(r["toplists"][2], g[p + "top_flat"])


@pytest.mark.parametrize("k", range(len(MULHITS_CASES)))
def test_port_bb_mulhits_matches_golden(k):
    g = dict(np.load(MULHITS_GOLD))
    c, o, seg, boot, bound = mulhits_golden_case(g, k)
    for tag in ("all", "cut"):
        r = run_bb(o, c, boot, seg, float(g["c%d_%s_cutoff" % (k, tag)]), None, False, mulhits=True)
        r["mats_tf"] = r["mats"][:, [3, 4]]
        check_mulhits_golden(g, k, tag, r)
    for N in TOPBOOT_NS:                                               # -mulhits -topboot N
        tag = "top%d" % N
        r = run_bb(o, c, boot, seg, float(g["c%d_%s_cutoff" % (k, tag)]), None, False, topboot=N)
        r["mats_tf"] = r["mats"][:, [3, 4]]
        check_mulhits_golden(g, k, tag, r)


# ---- -cost together with -bb: saveCurrentTree on Sankoff pattern vectors (pllComputeSankoffPatternParsimony :3346) ----
SANKOFF_BB_CASES = [(16, 400, 1, 7, 40, 0.05), (14, 300, 2, 5, 30, 0.05), (22, 700, 1, 21, 50, 0.02), (12, 260, 6, 9, 24, 0.05)]
SANKOFF_BB_GOLD = __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "golden", "skbb.npz")


def sankoff_bb_cost(dt, seed):
    S = {0: 2, 1: 4, 2: 20, 6: 32}[dt]
    if S == 4:
        return np.array([[0, 2, 1, 2], [2, 0, 2, 1], [1, 2, 0, 2], [2, 1, 2, 0]], dtype=np.uint32)
    r = np.random.default_rng(300 + seed).integers(1, 5, size=(S, S))
    r = np.minimum(r, r.T); np.fill_diagonal(r, 0)
    return r.astype(np.uint32)


@needs_ref
@pytest.mark.parametrize("n,L,dt,seed,B,mu", SANKOFF_BB_CASES)
def test_port_sankoff_bb_equals_reference(n, L, dt, seed, B, mu):
    c, o, s0, pp, seg, boot, ras, bound = bb_setup(n, L, dt, seed, B, mu)
    cost = sankoff_bb_cost(dt, seed)
    r = reflib.RefEngine(c["chars"], c["weights"], dt, n_informative=c["n_inf"])
    a = run_bb(o, c, boot, seg, 0.0, None, False, cost=cost)
    same(a, run_bb(r, c, boot, seg, 0.0, None, True, cost=cost))
    cutoff = -(a["ret"] + 6.0)
    x = run_bb(o, c, boot, seg, cutoff, None, False, cost=cost)
    same(x, run_bb(r, c, boot, seg, cutoff, None, True, cost=cost))
    assert x["counters"][1] < x["counters"][0]
    m = run_bb(o, c, boot, seg, 0.0, None, False, cost=cost, mulhits=True)
    same(m, run_bb(r, c, boot, seg, 0.0, None, True, cost=cost, mulhits=True))
    for e in (o, r):
        e.set_cost_matrix(None, None)


def sankoff_bb_golden_case(g, k):
    n, L, dt, seed, B, mu = SANKOFF_BB_CASES[k]
    c, o, s0, pp, seg, boot, ras, bound = bb_setup(n, L, dt, seed, B, mu)
    assert np.array_equal(boot, g["c%d_boot" % k]) and np.array_equal(c["codes"], g["c%d_codes" % k])
    return c, o, seg, boot, g["c%d_cost" % k]


def check_sankoff_bb_golden(g, k, tag, r):
    p = "c%d_%s_" % (k, tag)
    assert r["ret"] == int(g[p + "ret"]) and r["draws"] == int(g[p + "draws"])
    assert np.array_equal(r["ring"][0][3:], g[p + "bn"][3:]) and np.array_equal(r["ring"][1][3:], g[p + "bs"][3:])
    for x, key in zip(r["state"], ("boot_logl", "boot_counts", "boot_trees")):
        assert np.array_equal(x, g[p + key]), key
    assert np.array_equal(r["treels"], g[p + "treels"])
    assert np.array_equal(r["mats_tf"], g[p + "mats"])


@pytest.mark.parametrize("k", range(len(SANKOFF_BB_CASES)))
def test_port_sankoff_bb_matches_golden(k):
    g = dict(np.load(SANKOFF_BB_GOLD))
    c, o, seg, boot, cost = sankoff_bb_golden_case(g, k)
    for tag in ("all", "cut"):
        r = run_bb(o, c, boot, seg, float(g["c%d_%s_cutoff" % (k, tag)]), None, False, cost=cost)
        r["mats_tf"] = r["mats"][:, [3, 4]]
        check_sankoff_bb_golden(g, k, tag, r)
    o.set_cost_matrix(None, None)


def test_fingerprint_matches_python_helper():
    c, o, s0, pp, seg, boot, ras, bound = bb_setup(16, 300, 1, 3, 8)
    res = run_bb(o, c, boot, seg, 0.0, None, False)
    bn, bs = res["ring"]
    assert fingerprint_ring(bn, bs, 16) == o.fingerprint() % (1 << 64)
    m = res["mats"]
    assert len(m) > 0 and (m[:, 3] >= 0).all()


# ---- committed golden vectors (produced by the reference driver, tools/make_golden.py) ----------
import glob
import os

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASE_FILES = sorted(f for f in glob.glob(os.path.join(GOLD, "*.npz"))
                    if not f.endswith("tables.npz") and not os.path.basename(f).startswith(("sankoff_", "mulhits", "skbb")))
IDS = [os.path.basename(p)[:-4] for p in CASE_FILES]


def check_against_golden(g, tag, res):
    """res: dict(ret, draws, ring, state, counters=(calls, trees, reps rows), treels, mats[call?, tree_index, fp])"""
    assert res["ret"] == int(g["bb_%s_ret" % tag]) and res["draws"] == int(g["bb_%s_draws" % tag])
    assert np.array_equal(res["ring"][0][3:], g["bb_%s_bn" % tag][3:]) and np.array_equal(res["ring"][1][3:], g["bb_%s_bs" % tag][3:])
    assert np.array_equal(res["state"][0], g["bb_%s_boot_logl" % tag])
    assert np.array_equal(res["state"][1], g["bb_%s_boot_counts" % tag])
    assert np.array_equal(res["state"][2], g["bb_%s_boot_trees" % tag])
    assert np.array_equal(res["treels"], g["bb_%s_treels" % tag])
    cnt = g["bb_%s_counters" % tag]
    assert res["counters"][0] == int(cnt[0]) and res["counters"][1] == int(cnt[1]) and res["counters"][2] == int(cnt[2])
    assert np.array_equal(res["mats"], g["bb_%s_mats" % tag][:, 1:])       # tree_index, topology fingerprint


@pytest.mark.parametrize("path", CASE_FILES, ids=IDS)
def test_port_bb_equals_golden(path):
    g = dict(np.load(path))
    n, dt, mt = int(g["n"]), int(g["datatype"]), int(g["maxtrav"])
    c = dict(n=n, datatype=dt, codes=g["codes"], weights=g["weights"], n_inf=int(g["n_inf"]), bn=g["bn"], bs=g["bs"])
    o = portlib.OracleEngine(g["codes"], g["weights"], dt)
    for tag in ("all", "cut", "rall", "rcut"):
        rt = (g["bb_ratchet_weights"], g["bb_ratchet_orig"], g["bb_ratchet_init"]) if tag[0] == "r" else None
        r = run_bb(o, c, g["bb_boot"], g["bb_seg"], float(g["bb_%s_cutoff" % tag]), None, False, mt=mt, ratchet=rt)
        r["mats"] = r["mats"][:, [3, 4]]
        check_against_golden(g, tag, r)
