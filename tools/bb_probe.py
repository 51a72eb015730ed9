"""GPU probe of the -bb path at bench size: times mpgpu_reps_candidates over every candidate of
one SPR sweep (all candidates pass the cutoff = the worst case for REPS) and checks the tensor
path against the exact CUDA-core path on a sample.  Usage: python tools/bb_probe.py [workload] [B]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mpboot_b200 import engine  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    check = int(os.environ.get("BB_PROBE_CHECK", "1"))
    case = bench.build_case(wl, 1)
    n = case["n"]
    rng = np.random.default_rng(5)
    P = case["codes"].shape[1]
    w = case["weights"].astype(np.float64)
    boot = rng.multinomial(int(w.sum()), w / w.sum(), size=B).astype(np.uint16)
    ninf = case["n_inf"]
    eng = engine.Engine()
    eng.load_alignment(case["codes"], case["weights"], case["datatype"])
    eng.set_tree(case["bn"], case["bs"])
    seg = bench.do_segmenting(eng.pattern_parsimony()[0][:ninf], case["weights"], ninf)
    print("segments: %d (mean %.0f patterns)" % (len(seg), ninf / len(seg)), flush=True)
    t0 = time.time(); eng.load_replicates(boot, seg); t1 = time.time()
    print("load_replicates %.3f s, info (groups, exceptions, tensor) = %s" % (t1 - t0, eng.reps_info()), flush=True)
    order = eng.visit_order()
    vb, mp, cref, cprune = eng.scan_visits(order, 1, 2 * n - 2, 1, 6)
    ncand = len(mp)
    idx = np.arange(-1, ncand, dtype=np.int32)
    for rep in range(3):
        t0 = time.time(); res = eng.reps_candidates(idx); dt = time.time() - t0
        macs = float(len(idx)) * ninf * B
        print("reps_candidates: %d calls x %d patterns x %d replicates in %.2f ms  (%.1f TMAC/s incl. D2H of %d MB)"
              % (len(idx), ninf, B, dt * 1e3, macs / dt / 1e12, res.nbytes >> 20), flush=True)
    # device-only step: scan + rows + contraction + combine, nothing read back
    import torch
    eng.set_option("reps_timing", 1)
    calls = []
    for v in range(2 * n - 2):
        calls.append(-1); calls.extend(range(vb[v], vb[v + 1]))
    calls = np.array(calls, dtype=np.int32)
    eng.scan_plan(order, 1, 2 * n - 2, 1, 6)
    for rep in range(4):
        eng.synchronize(); t0 = time.time()
        eng.scan_launch(); eng.reps_candidates_device(calls)
        t_host = time.time() - t0
        eng.synchronize(); dt = time.time() - t0
        ms, rows, pat, sp = eng.reps_timing()
        print("bb step (device, %d calls): %.3f ms wall (host enqueue %.3f ms); k_reps_tc %.3f ms for %d rows x %d patterns x %d, splits %d -> %.2f POPS"
              % (len(calls), dt * 1e3, t_host * 1e3, ms, rows, pat, B, sp, 2.0 * rows * ninf * B / (ms * 1e-3) / 1e15), flush=True)
    if os.environ.get("BB_PROBE_SEARCH", "1") == "1":
        from oracle import portlib
        Bn = boot.shape[0]
        for cutoff_on in (False, True):
            bl = np.full(Bn, -float(np.iinfo(np.int64).max)); bc = np.zeros(Bn, dtype=np.int32); bt = np.full(Bn, -1, dtype=np.int32)
            tl = engine.Treels(n)
            portlib.seed_rng(1)
            cutoff = 0.0
            if cutoff_on:
                cutoff = -(float(final_score) + 10.0)
            t0 = time.time()
            ret, bn2, bs2, nins, ncalls, nreps = eng.optimize_spr_bb(case["bn"], case["bs"], tl.hooks(portlib.rng_fn_address()), bl, bc, bt, cutoff)
            dt = time.time() - t0
            final_score = ret
            print("optimize_spr_bb cutoff=%s: startMP %d, %d insertions, %d calls, %d REPS vectors, %d trees materialised, %.2f s -> %.2f M insertions/s"
                  % (cutoff, ret, nins, ncalls, nreps, len(tl.materialized()), dt, nins / dt / 1e6), flush=True)
    if check:
        eng2 = engine.Engine()
        eng2.set_option("reps_tensor", 0)
        eng2.load_alignment(case["codes"], case["weights"], case["datatype"])
        eng2.set_tree(case["bn"], case["bs"])
        eng2.load_replicates(boot, seg)
        eng2.scan_visits(order, 1, 2 * n - 2, 1, 6)
        sample = np.concatenate([[-1], rng.choice(ncand, size=min(300, ncand), replace=False)]).astype(np.int32)
        a = eng2.reps_candidates(sample)
        assert np.array_equal(a, res[sample + 1]), "tensor path differs from the exact CUDA-core path"
        print("tensor path == exact CUDA-core path on %d sampled calls" % len(sample))


if __name__ == "__main__":
    main()
