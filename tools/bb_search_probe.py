"""GPU probe of the whole -bb SPR search at bench size (the call bench.py's bb.search times): mpgpu_optimize_spr_bb from the
random tree, cutoff off, B = 1000.  MPGPU_PROFILE=1 adds the library's wall-clock breakdown.
Usage: python tools/bb_search_probe.py [workload] [repeats]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mpboot_b200 import engine  # noqa: E402
from mpboot_b200.engine import HostRng, Treels  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    B = 1000
    case = bench.build_case(wl, 1)
    n, ninf = case["n"], case["n_inf"]
    eng = engine.Engine()
    eng.load_alignment(case["codes"], case["weights"], case["datatype"])
    eng.set_tree(case["bn"], case["bs"])
    boot = bench.make_replicates(case, B)
    seg = bench.do_segmenting(eng.pattern_parsimony()[0][:ninf], case["weights"], ninf)
    eng.load_replicates(boot, seg)
    for k in range(reps):
        rs = HostRng(11)
        bl = np.full(B, -float(np.iinfo(np.int64).max)); bc = np.zeros(B, dtype=np.int32); bt = np.full(B, -1, dtype=np.int32)
        tl = Treels(n)
        l0 = eng.launch_count()
        t0 = time.time()
        ret, bn2, bs2, nins, ncalls, nreps = eng.optimize_spr_bb(case["bn"], case["bs"], tl.hooks(rs.fn, rs.user), bl, bc, bt, 0.0, 0.5, 1, 6)
        dt = time.time() - t0
        print("mpgpu_optimize_spr_bb: %.3f s, score %d, %d insertions, %d REPS vectors, %d launches, checksum %d"
              % (dt, ret, nins, nreps, eng.launch_count() - l0, int(bl.sum()) % 1000003), flush=True)


if __name__ == "__main__":
    main()
