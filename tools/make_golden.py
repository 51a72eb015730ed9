"""Generates tests/golden/*.npz by running the REFERENCE ITSELF (oracle/_ref/libmpref.so =
/root/reference/sprparsimony.cpp compiled in place, see oracle/ref_driver.cpp) on small seeded
cases.  Run here (where /root/reference exists):  python tools/make_golden.py
The fixtures are what pins oracle/mp_oracle.c and, through it, the CUDA path."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import reflib  # noqa: E402
from tests.helpers import make_case, make_boot  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# (name, n, sites, datatype, seed, maxtrav)
CASES = [
    ("dna12", 12, 300, 1, 7, 6),
    ("dna40", 40, 1500, 1, 11, 6),
    ("aa24", 24, 400, 2, 5, 6),
    ("morph20", 20, 300, 6, 9, 5),
    ("bin16", 16, 300, 0, 3, 4),
    ("dna30w", 30, 800, 1, 21, 6),     # heavier pattern weights (low mu -> repeated columns)
]


def one_case(name, n, L, dt, seed, maxtrav):
    mu = 0.01 if name.endswith("w") else 0.05
    c = make_case(n, L, dt, seed, mu=mu)
    ref = reflib.RefEngine(c["chars"], c["weights"], dt, n_informative=c["n_inf"])
    ref.set_ring(c["bn"], c["bs"])
    W = ref.allocate(per_site=True)
    g = dict(n=n, datatype=dt, maxtrav=maxtrav, chars=c["chars"], codes=c["codes"], weights=c["weights"],
             n_inf=c["n_inf"], bn=c["bn"], bs=c["bs"], W=W, ref_n_inf=ref.num_informative())
    g["tip_planes"] = np.stack([ref.parsvect(t) for t in range(1, n + 1)])
    g["score"] = ref.evaluate_full(per_site=True)
    pp, sm = ref.pattern_parsimony(c["n_inf"])
    g["ptn_pars"] = pp.copy(); g["ptn_sum"] = sm
    rn, rs = ref.get_nodep()
    g["order"] = (3 * rn + rs).astype(np.int32)
    g["min_pars"] = np.array([ref.min_pars_pattern(i) for i in range(c["n_inf"])], dtype=np.int32)
    # every node visit of one sweep on the start tree, moves not applied
    vb = [0]; mps = []; outs = []; ptn_first = None
    reflib.lib().mpref_seed_rng(31337)               # ties inside a visit draw from the stream
    for i in range(1, 2 * n - 1):
        ref.record(i == 3)
        rc, out = ref.rearrange(i, 1, maxtrav, True, g["score"])
        if i == 3:
            m, pt = ref.saved(True)
            ptn_first = pt[:, : c["n_inf"]].copy()
        else:
            m = ref.saved()
        assert m[0] == g["score"]
        mps.append(m[1:]); vb.append(vb[-1] + len(m) - 1); outs.append(out)
    g["visit_begin"] = np.array(vb, dtype=np.int32)
    g["visit_mp"] = np.concatenate(mps).astype(np.int32)
    g["visit_out"] = np.stack(outs).astype(np.uint32)
    g["visit3_ptn"] = ptn_first                      # per-pattern vectors of every candidate of visit 3
    # the real search, plain and -bb, with a fixed RNG stream
    for tag, bb in (("plain", False), ("bb", True)):
        reflib.lib().mpref_seed_rng(2024)
        ref.set_ring(c["bn"], c["bs"])
        ref.record(False)
        g["opt_%s_ret" % tag] = ref.optimize_spr(1, maxtrav, bb=bb)
        g["opt_%s_draws" % tag] = reflib.lib().mpref_rng_draws()
        bn, bs = ref.get_ring()
        g["opt_%s_bn" % tag] = bn; g["opt_%s_bs" % tag] = bs
        g["opt_%s_score" % tag] = ref.evaluate_full(per_site=bb)
        if bb:
            g["opt_bb_saved"] = ref.saved().astype(np.int32)
    # the -bb search with the whole saveCurrentTree bookkeeping (re-typed default policy on top of the
    # reference's own search, pattern scores and Vec16us REPS, oracle/ref_driver.cpp): 40 replicates, a few
    # of them forced to wrap at 16 bits / exceed 255, artificial segment bounds, with and without cutoff
    ninf = c["n_inf"]
    seg = np.unique(np.array([16, 48, 16 * max(ninf // 32, 4), ninf], dtype=np.int32))
    seg = seg[seg <= ninf]
    hv = [(1, 3, 40000), (2, 5, 300), (2, 6, 65535)] + [(3, k, 900) for k in range(0, min(ninf, 40))]
    boot = make_boot(c, 40, seed, heavy=hv)
    g["bb_boot"] = boot; g["bb_seg"] = seg
    ras = np.zeros(len(c["weights"]), dtype=np.int32); ras[:ninf] = g["ptn_pars"]
    for tag, cutoff in (("all", 0.0), ("cut", -(float(g["opt_bb_ret"]) + 4.0))):
        ref.set_ring(c["bn"], c["bs"])
        ref.allocate(per_site=True)
        ref.boot_init(boot, seg, cutoff, 0.5, ras)
        reflib.lib().mpref_seed_rng(2024)
        ref.record(False)
        g["bb_%s_cutoff" % tag] = cutoff
        g["bb_%s_ret" % tag] = ref.optimize_spr(1, maxtrav, bb=True)
        g["bb_%s_draws" % tag] = reflib.lib().mpref_rng_draws()
        bn, bs = ref.get_ring()
        g["bb_%s_bn" % tag] = bn; g["bb_%s_bs" % tag] = bs
        bl, bc, bt = ref.boot_state()
        g["bb_%s_boot_logl" % tag] = bl; g["bb_%s_boot_counts" % tag] = bc; g["bb_%s_boot_trees" % tag] = bt
        g["bb_%s_counters" % tag] = np.array(ref.boot_counters(), dtype=np.int64)
        g["bb_%s_treels" % tag] = ref.boot_treels()
        g["bb_%s_mats" % tag] = ref.boot_mats()[:, [0, 3, 4]]          # call index, tree_index, topology fingerprint
    # a ratchet iteration (on_ratchet_hclimb1): search on perturbed frequencies, cur_logl re-scored on the
    # original ones from the previous call's vector (iqtree.cpp:3283-3294); cutoff chosen so that the chain of
    # passing calls breaks somewhere inside the search
    from tests.test_bb_cpu import ratchet_setup
    rt = ratchet_setup(c, g["ptn_pars"], seed)
    g["bb_ratchet_weights"] = rt[0]; g["bb_ratchet_orig"] = rt[1]; g["bb_ratchet_init"] = rt[2]
    cut = 0.0
    for tag in ("rall", "rcut"):
        ref.set_weights(rt[0])
        ref.set_ring(c["bn"], c["bs"])
        ref.allocate(per_site=True)
        ref.boot_init(boot, seg, cut, 0.5, None)
        ref.boot_set_ratchet(rt[1], rt[2])
        reflib.lib().mpref_seed_rng(2024)
        ref.record(False)
        g["bb_%s_cutoff" % tag] = cut
        g["bb_%s_ret" % tag] = ref.optimize_spr(1, maxtrav, bb=True)
        g["bb_%s_draws" % tag] = reflib.lib().mpref_rng_draws()
        bn, bs = ref.get_ring()
        g["bb_%s_bn" % tag] = bn; g["bb_%s_bs" % tag] = bs
        bl, bc, bt = ref.boot_state()
        g["bb_%s_boot_logl" % tag] = bl; g["bb_%s_boot_counts" % tag] = bc; g["bb_%s_boot_trees" % tag] = bt
        g["bb_%s_counters" % tag] = np.array(ref.boot_counters(), dtype=np.int64)
        g["bb_%s_treels" % tag] = ref.boot_treels()
        g["bb_%s_mats" % tag] = ref.boot_mats()[:, [0, 3, 4]]
        top = np.unique(-g["bb_rall_treels"])[::-1]
        cut = -(float(top[min(1, len(top) - 1)]) - 0.5)
    ref.set_weights(c["weights"])
    ref.boot_free()
    # randomized stepwise addition
    reflib.lib().mpref_seed_rng(77)
    g["ras_seed"] = 4242 + seed
    g["ras_ret"] = ref.ras(4242 + seed, maxtrav)
    g["ras_draws"] = reflib.lib().mpref_rng_draws()
    bn, bs = ref.get_ring()
    g["ras_bn"] = bn; g["ras_bs"] = bs
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **g)
    print(name, "W", W, "score", g["score"], "cands", len(g["visit_mp"]), "opt", g["opt_plain_ret"], g["opt_bb_ret"],
          "saved", len(g["opt_bb_saved"]), "ras", g["ras_ret"])


def tables():
    g = {}
    for dt in (0, 1, 2, 6):
        g["char_map_%d" % dt] = reflib.char_map(dt)
        nc = {0: 4, 1: 16, 2: 23, 6: 33}[dt]
        bv, und = reflib.bitvector(dt, nc)
        g["bitvector_%d" % dt] = bv; g["undetermined_%d" % dt] = und
    # REPS on the reference's Vec16us, including products and sums that wrap at 16 bits
    rng = np.random.default_rng(5)
    P = 203
    pars = rng.integers(0, 40, size=P).astype(np.uint16)
    boot = rng.integers(0, 6, size=(37, P)).astype(np.uint16)
    boot[3, 17] = 40000; pars[17] = 3          # lane product wraps
    boot[5, :64] = 900                         # lane sums wrap
    seg = np.array([64, 128, P], dtype=np.int32)
    g["reps_pars"] = pars; g["reps_boot"] = boot; g["reps_seg"] = seg
    g["reps_out"] = reflib.reps(pars, boot, seg)
    g["reps_out_1seg"] = reflib.reps(pars, boot, np.array([P], dtype=np.int32))
    np.savez_compressed(os.path.join(OUT, "tables.npz"), **g)
    print("tables ok; reps sample", g["reps_out"][:6])


if __name__ == "__main__":
    assert reflib.available(), "build oracle/_ref first: make -C oracle ref"
    os.makedirs(OUT, exist_ok=True)
    tables()
    for c in CASES:
        one_case(*c)
