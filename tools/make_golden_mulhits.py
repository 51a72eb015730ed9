"""Generates tests/golden/mulhits.npz by running the REFERENCE driver (oracle/_ref/libmpref.so: the reference's own
search, pattern scores and Vec16us REPS, with saveCurrentTree's -mulhits branch iqtree.cpp:3498-3531 re-typed in
oracle/ref_driver.cpp) on whole -bb SPR searches under -mulhits: SURVEY 8a row R9.
Run here (where /root/reference exists):  python tools/make_golden_mulhits.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import reflib  # noqa: E402
from tests.test_bb_cpu import MULHITS_CASES, bb_setup, run_bb  # noqa: E402

if __name__ == "__main__":
    assert reflib.available(), "build oracle/_ref first: make -C oracle ref"
    g = {}
    for k, (n, L, dt, seed, B, mu) in enumerate(MULHITS_CASES):
        c, o, s0, pp, seg, boot, ras, bound = bb_setup(n, L, dt, seed, B, mu)
        r = reflib.RefEngine(c["chars"], c["weights"], dt, n_informative=c["n_inf"])
        g["c%d_boot" % k] = boot; g["c%d_codes" % k] = c["codes"]
        first = run_bb(r, c, boot, seg, 0.0, None, True, mulhits=True)
        for tag, cutoff in (("all", 0.0), ("cut", -(first["ret"] + 4.0))):
            x = run_bb(r, c, boot, seg, cutoff, ras if cutoff else None, True, mulhits=True)
            p = "c%d_%s_" % (k, tag)
            g[p + "cutoff"] = cutoff; g[p + "ret"] = x["ret"]; g[p + "draws"] = x["draws"]
            g[p + "bn"], g[p + "bs"] = x["ring"]
            g[p + "boot_logl"] = x["state"][0]
            g[p + "sizes"], g[p + "flat"] = x["mulhits"]
            g[p + "treels"] = x["treels"]; g[p + "mats"] = x["mats"][:, [3, 4]]
            print(k, tag, "ret", x["ret"], "draws", x["draws"], "calls", x["counters"][0], "trees", len(x["treels"]),
                  "set sizes", x["mulhits"][0].min(), x["mulhits"][0].max(), "materialised", len(x["mats"]))
    from tests.test_bb_cpu import TOPBOOT_NS
    for k, (n, L, dt, seed, B, mu) in enumerate(MULHITS_CASES):          # -mulhits -topboot N
        c, o, s0, pp, seg, boot, ras, bound = bb_setup(n, L, dt, seed, B, mu)
        r = reflib.RefEngine(c["chars"], c["weights"], dt, n_informative=c["n_inf"])
        for N in TOPBOOT_NS:
            x = run_bb(r, c, boot, seg, 0.0, None, True, topboot=N)
            p = "c%d_top%d_" % (k, N)
            g[p + "cutoff"] = 0.0; g[p + "ret"] = x["ret"]; g[p + "draws"] = x["draws"]
            g[p + "bn"], g[p + "bs"] = x["ring"]
            g[p + "boot_logl"] = x["state"][0]
            g[p + "sizes"], g[p + "flat"] = x["mulhits"]
            g[p + "treels"] = x["treels"]; g[p + "mats"] = x["mats"][:, [3, 4]]
            g[p + "top_sizes"], g[p + "top_thr"], g[p + "top_flat"] = x["toplists"]
            print(k, "topboot", N, "ret", x["ret"], "draws", x["draws"], "materialised", len(x["mats"]), "list sizes",
                  x["toplists"][0].min(), x["toplists"][0].max())
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "mulhits.npz"), **g)
