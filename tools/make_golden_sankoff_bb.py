"""Generates tests/golden/skbb.npz by running the REFERENCE driver (oracle/_ref/libmpref.so: the reference's own
Sankoff kernels, search, pllComputeSankoffPatternParsimony and Vec16us REPS; saveCurrentTree's default policy re-typed
in oracle/ref_driver.cpp) on whole -bb SPR searches under -cost: SURVEY 8a rows R11 + R8/R9.
Run here (where /root/reference exists):  python tools/make_golden_sankoff_bb.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import reflib  # noqa: E402
from tests.test_bb_cpu import SANKOFF_BB_CASES, bb_setup, run_bb, sankoff_bb_cost  # noqa: E402

if __name__ == "__main__":
    assert reflib.available(), "build oracle/_ref first: make -C oracle ref"
    g = {}
    for k, (n, L, dt, seed, B, mu) in enumerate(SANKOFF_BB_CASES):
        c, o, s0, pp, seg, boot, ras, bound = bb_setup(n, L, dt, seed, B, mu)
        cost = sankoff_bb_cost(dt, seed)
        r = reflib.RefEngine(c["chars"], c["weights"], dt, n_informative=c["n_inf"])
        g["c%d_boot" % k] = boot; g["c%d_codes" % k] = c["codes"]; g["c%d_cost" % k] = cost
        first = run_bb(r, c, boot, seg, 0.0, None, True, cost=cost)
        for tag, cutoff in (("all", 0.0), ("cut", -(first["ret"] + 6.0))):
            x = run_bb(r, c, boot, seg, cutoff, None, True, cost=cost)
            p = "c%d_%s_" % (k, tag)
            g[p + "cutoff"] = cutoff; g[p + "ret"] = x["ret"]; g[p + "draws"] = x["draws"]
            g[p + "bn"], g[p + "bs"] = x["ring"]
            g[p + "boot_logl"], g[p + "boot_counts"], g[p + "boot_trees"] = x["state"]
            g[p + "treels"] = x["treels"]; g[p + "mats"] = x["mats"][:, [3, 4]]
            print(k, tag, "ret", x["ret"], "draws", x["draws"], "calls", x["counters"][0], "trees", len(x["treels"]),
                  "materialised", len(x["mats"]), "bad sums", x["counters"][4])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "skbb.npz"), **g)
