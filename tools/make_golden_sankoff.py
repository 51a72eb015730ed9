"""Generates tests/golden/sankoff_*.npz by running the REFERENCE ITSELF (oracle/_ref/libmpref.so, the
unmodified sprparsimony.cpp with pllCostMatrix set, see oracle/ref_driver.cpp) on small seeded cases of
the -cost (Sankoff, weighted parsimony) path: SURVEY 8a row R11.
Run here (where /root/reference exists):  python tools/make_golden_sankoff.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import reflib  # noqa: E402
from tests.helpers import make_case  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
TSTV = np.array([[0, 2, 1, 2], [2, 0, 2, 1], [1, 2, 0, 2], [2, 1, 2, 0]], dtype=np.uint32)   # transitions 1, transversions 2


def sym_cost(S, seed, hi=5):
    rng = np.random.default_rng(seed)
    a = rng.integers(1, hi, size=(S, S))
    a = np.minimum(a, a.T)
    np.fill_diagonal(a, 0)
    return a.astype(np.uint32)


# (name, n, sites, datatype, seed, maxtrav, cost, interior segment bounds)
CASES = [
    ("sankoff_dna16", 16, 600, 1, 3, 6, TSTV, []),
    ("sankoff_dna24", 24, 900, 1, 4, 6, sym_cost(4, 4), [160, 320, 480]),
    ("sankoff_aa12", 12, 300, 2, 5, 5, sym_cost(20, 5), [64, 128]),
    ("sankoff_morph14", 14, 300, 6, 6, 5, sym_cost(32, 6), [96]),
    ("sankoff_bin10", 10, 200, 0, 7, 4, sym_cost(2, 7), []),
    ("sankoff_dna20w", 20, 700, 1, 8, 6, sym_cost(4, 8, hi=9), [48, 96, 160]),   # repeated columns -> pattern weights > 1
]


def one_case(name, n, L, dt, seed, maxtrav, cost, segs):
    mu = 0.01 if name.endswith("w") else 0.05
    c = make_case(n, L, dt, seed, mu=mu)
    ninf = c["n_inf"]
    seg = np.array([s for s in segs if s < ninf] + [ninf], dtype=np.int32)
    ref = reflib.RefEngine(c["chars"], c["weights"], dt, n_informative=ninf)
    g = dict(n=n, datatype=dt, maxtrav=maxtrav, chars=c["chars"], codes=c["codes"], weights=c["weights"], n_inf=ninf,
             bn=c["bn"], bs=c["bs"], cost=cost, seg=seg)
    g["highest"] = ref.set_cost_matrix(cost, seg)
    ref.set_ring(c["bn"], c["bs"])
    g["L"] = ref.allocate(per_site=False)
    g["tips"] = np.stack([ref.sankoff_vect(t) for t in range(1, n + 1)])
    g["score"] = ref.evaluate_full(per_site=False)
    g["lower_bounds"] = ref.remainder_bounds()
    g["node_vect"] = np.stack([ref.sankoff_vect(i) for i in range(n + 1, 2 * n - 1)])    # as oriented by evaluate(start, full)
    g["node_score"] = np.array([ref.node_score(i) for i in range(n + 1, 2 * n - 1)], dtype=np.uint32)
    rn, rs = ref.get_nodep()
    g["order"] = (3 * rn + rs).astype(np.int32)
    # one sweep in plain mode (early termination active: the decisions), moves not applied
    reflib.lib().mpref_seed_rng(31337)
    outs = []
    for i in range(1, 2 * n - 1):
        rc, out = ref.rearrange(i, 1, maxtrav, False, g["score"])
        outs.append(out)
    g["plain_visit_out"] = np.stack(outs).astype(np.uint32)
    g["plain_draws"] = reflib.lib().mpref_rng_draws()
    # the same sweep with per-pattern scores (no early termination: every insertion's exact score)
    ref.allocate(per_site=True)
    assert ref.evaluate_full(per_site=True) == g["score"]
    pp, sm = ref.pattern_parsimony(ninf)
    g["ptn_pars"] = pp.copy(); g["ptn_sum"] = sm
    reflib.lib().mpref_seed_rng(31337)
    vb = [0]; mps = []; outs = []
    for i in range(1, 2 * n - 1):
        ref.record(False)
        rc, out = ref.rearrange(i, 1, maxtrav, True, g["score"])
        m = ref.saved()
        assert m[0] == g["score"]
        mps.append(m[1:]); vb.append(vb[-1] + len(m) - 1); outs.append(out)
    g["visit_begin"] = np.array(vb, dtype=np.int32)
    g["visit_mp"] = np.concatenate(mps).astype(np.int32)
    g["visit_out"] = np.stack(outs).astype(np.uint32)
    for tag, bb in (("plain", False), ("bb", True)):
        reflib.lib().mpref_seed_rng(2024)
        ref.set_ring(c["bn"], c["bs"])
        ref.allocate(bb)
        ref.record(False)
        g["opt_%s_ret" % tag] = ref.optimize_spr(1, maxtrav, bb=bb)
        g["opt_%s_draws" % tag] = reflib.lib().mpref_rng_draws()
        bn, bs = ref.get_ring()
        g["opt_%s_bn" % tag] = bn; g["opt_%s_bs" % tag] = bs
        if bb:
            g["opt_bb_saved"] = ref.saved().astype(np.int32)
    reflib.lib().mpref_seed_rng(77)
    g["ras_seed"] = 4242 + seed
    g["ras_ret"] = ref.ras(4242 + seed, maxtrav)
    g["ras_draws"] = reflib.lib().mpref_rng_draws()
    bn, bs = ref.get_ring()
    g["ras_bn"] = bn; g["ras_bs"] = bs
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **g)
    print(name, "L", g["L"], "score", g["score"], "cands", len(g["visit_mp"]), "opt", g["opt_plain_ret"], g["opt_bb_ret"],
          "ras", g["ras_ret"], "segments", seg, "LB", g["lower_bounds"])


if __name__ == "__main__":
    assert reflib.available(), "build oracle/_ref first: make -C oracle ref"
    for c in CASES:
        one_case(*c)
