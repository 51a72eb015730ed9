"""GPU probe of whole searches at bench size: mpgpu_optimize_spr (plain) and mpgpu_stepwise_addition from
the same start as the reference's pllOptimizeSprParsimony (oracle/_ref), same RNG stream -- checks that the
final tree, score and draw count are identical and prints both wall times.  MPGPU_PROFILE=1 adds the
library's own wall-clock breakdown.   Usage: python tools/search_probe.py [workload] [repeats]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mpboot_b200 import engine  # noqa: E402
from oracle import reflib  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    case = bench.build_case(wl, 1)
    L_ = reflib.lib()
    fn = C.cast(L_.mpref_random_double, C.c_void_p).value
    ref_result = None
    if os.environ.get("SEARCH_PROBE_REF", "1") == "1":
        ref = reflib.RefEngine(case["chars"], case["weights"], case["datatype"], n_informative=case["n_inf"])
        ref.set_ring(case["bn"], case["bs"])
        L_.mpref_seed_rng(1234)
        t0 = time.time()
        r_ref = ref.optimize_spr(1, 6, bb=False)
        dt = time.time() - t0
        bn_ref, bs_ref = ref.get_ring()
        ref_result = (r_ref, L_.mpref_rng_draws(), bn_ref, bs_ref)
        print("reference pllOptimizeSprParsimony: %.3f s, score %d, %d draws" % (dt, r_ref, ref_result[1]), flush=True)
    eng = engine.Engine()
    eng.load_alignment(case["codes"], case["weights"], case["datatype"])
    for k in range(reps):
        L_.mpref_seed_rng(1234)
        l0 = eng.launch_count()
        t0 = time.time()
        r, bn, bs, nins = eng.optimize_spr(case["bn"], case["bs"], fn, 1, 6)
        dt = time.time() - t0
        print("mpgpu_optimize_spr: %.3f s, score %d, %d insertions (%.2f M/s), %d draws, %d launches"
              % (dt, r, nins, nins / dt / 1e6, L_.mpref_rng_draws(), eng.launch_count() - l0), flush=True)
        if ref_result:
            assert r == ref_result[0] and L_.mpref_rng_draws() == ref_result[1]
            assert np.array_equal(bn[3:], ref_result[2][3:]) and np.array_equal(bs[3:], ref_result[3][3:])
            print("  identical to the reference (score, draws, final ring table)", flush=True)


if __name__ == "__main__":
    main()
