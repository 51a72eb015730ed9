#!/usr/bin/env python
"""bench.py -- throughput of the parsimony hot path (SPR insertion scoring) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one batch: every insertion that one SPR sweep
(rearrangeParsimony for all 2n-2 node visits, radius 1..6) tests on a fixed random tree is
scored by one launch of k_spr_scan.  Workload at N=1: BASELINE.json configs[1], synthetic DNA
200 taxa x 100 000 sites (every site its own pattern).  At N>1 the alignment is pattern-sharded
(SURVEY 8e): every rank holds a contiguous word slice of every bit plane, the slice has the
same 100 000 sites per rank (weak scaling), and the per-insertion int32 counts are all-reduced
over NCCL each step.

metric = Fitch site-node ops/s (one insertion = one node combine + one edge evaluation over all
expanded sites = 2*sites site-node ops); insertions/s is reported next to it.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (taxa, sites per GPU, datatype, mu, seed, tree seed)
    "c2": (200, 100000, 1, 0.05, 2, 102),
    "c3": (500, 50000, 2, 0.08, 3, 103),
    "c4": (1000, 1000000, 1, 0.03, 4, 104),
    "c5": (300, 20000, 6, 0.05, 5, 105),
    "tiny": (24, 2000, 1, 0.05, 6, 106),
}
METRIC = "fitch_site_node_ops_per_s"
UNIT = "site-node ops/s"


def build_case(name, nshards, scaling="weak"):
    """weak: `sites` per GPU (the alignment grows with the shard count); strong: `sites` in total."""
    from mpboot_b200 import hostprep, synth
    n, sites, dt, mu, seed, tseed = WORKLOADS[name]
    total = sites * nshards if scaling == "weak" else sites
    if n * total > (1 << 28):                                # C4-sized: bounded host memory
        chars = synth.evolve_alignment_blocked(n, total, dt, mu, seed)
    else:
        chars = synth.evolve_alignment(n, total, dt, mu, seed)
    prep = hostprep.prepare(chars, dt, compress=False)     # every site its own pattern
    bn, bs = synth.random_tree_rings(n, np.random.default_rng(tseed))
    prep.update(n=n, datatype=dt, bn=bn, bs=bs, sites=total)
    return prep


def do_segmenting(scores, weights, n_informative):
    """IQTree::doSegmenting (iqtree.cpp:3793-3819): a new REPS segment every time the running
    sum of ras_pars_score * frequency exceeds USHRT_MAX / 16 at a multiple of 16 patterns."""
    seg, run = [], 0
    prod = np.zeros(len(weights), dtype=np.int64)
    prod[: len(scores)] = np.asarray(scores, dtype=np.int64) * np.asarray(weights[: len(scores)], dtype=np.int64)
    for i in range(len(weights)):
        run += int(prod[i])
        if (i + 1) % 16 == 0 and run > 65535 // 16:
            seg.append(i + 1); run = 0
    if run:
        seg.append(n_informative)
    return np.array(seg, dtype=np.int32)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_traffic(workload, kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/traffic.json)."""
    try:
        return int(json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[workload][kernel]["bytes"])
    except Exception:
        return None


def bench_search(eng, case, args):
    """Whole plain SPR search (pllOptimizeSprParsimony) from the workload's random tree through the C-ABI with
    host ring tables in and out -- the call MPBoot makes -- next to the reference's own search on one host core
    (same start, same RNG stream; score, draw count and final tree must be identical)."""
    from mpboot_b200 import engine
    # Our arm draws from the library's own splitmix64 (mpgpu_splitmix64_double); the reference side below is seeded with
    # the same value on its own copy of the same generator, so both see the same stream.  Nothing under oracle/ runs
    # inside our timed calls.
    GOLD, INV = 0x9E3779B97F4A7C15, pow(0x9E3779B97F4A7C15, -1, 1 << 64)

    def draws_of(rng, seed):
        return ((rng.state.value - seed) * INV) % (1 << 64)

    best = None
    for _ in range(3):
        rng = engine.HostRng(1234)
        t0 = time.time()
        r, bn, bs, nins = eng.optimize_spr(case["bn"], case["bs"], rng.fn, 1, args.maxtrav, rng_user=rng.user)
        dt = time.time() - t0
        if best is None or dt < best:
            best = dt
    draws = int(draws_of(rng, 1234))
    out = {"what": "mpgpu_optimize_spr from the random tree to convergence, host buffers (e2e), best of 3",
           "wall_s": best, "insertions": int(nins), "insertions_per_s": nins / best, "final_score": int(r), "rng_draws": draws}
    # N1: the refinement loop of optimizeBootTrees over resident codes, from the search's final tree
    Bref = 20
    boot = make_replicates(case, Bref, seed=9)
    tbn = np.tile(bn, (Bref, 1)); tbs = np.tile(bs, (Bref, 1))
    rng = engine.HostRng(99)
    t0 = time.time()
    sc, _, _, rins = eng.refine_replicates(boot, tbn, tbs, rng.fn, 1, args.maxtrav, rng_user=rng.user)
    dt = time.time() - t0
    out["refine"] = {"what": "mpgpu_refine_replicates: %d bootstrap replicates re-weighted over the resident codes and hill-climbed from the search's final tree" % Bref,
                     "wall_s": dt, "replicates_per_s": Bref / dt, "insertions": int(rins), "insertions_per_s": rins / dt}
    eng.set_tree(case["bn"], case["bs"])
    if not args.no_cpu_baseline and args.workload in ("c2", "c3", "c5", "tiny"):
        from oracle import portlib, reflib
        use_ref = reflib.available()
        if use_ref:
            seed_fn, draws_fn = reflib.lib().mpref_seed_rng, reflib.lib().mpref_rng_draws
        else:
            seed_fn, draws_fn = portlib.seed_rng, portlib.rng_draws
        if use_ref:
            ref = reflib.RefEngine(case["chars"], case["weights"], case["datatype"], n_informative=case["n_inf"]); kind = "reference"
        else:
            ref = portlib.OracleEngine(case["codes"], case["weights"], case["datatype"]); kind = "port"
        ref.set_ring(case["bn"], case["bs"])
        seed_fn(1234)
        t0 = time.time()
        r_ref = ref.optimize_spr(1, args.maxtrav, bb=False)
        dt = time.time() - t0
        bn_ref, bs_ref = ref.get_ring()
        same = bool(r_ref == r and int(draws_fn()) == draws and np.array_equal(bn[3:], bn_ref[3:]) and np.array_equal(bs[3:], bs_ref[3:]))
        out["cpu_baseline"] = {"wall_s": dt, "cores": 1, "kind": kind, "final_score": int(r_ref), "identical_result": same,
                               "sample": "the whole search (pllOptimizeSprParsimony), single-threaded code"}
        out["speedup_vs_one_core"] = dt / best
        # MPBoot's own starting point (config[0], `-s`): one randomized-stepwise-addition tree + its SPR rounds
        # (_pllComputeRandomizedStepwiseAdditionParsimonyTree, sprparsimony.cpp:3224), same seeds on both sides
        rng = engine.HostRng(4321)
        t0 = time.time()
        rb, rbn, rbs, rins2, _ = eng.stepwise_addition(777, args.maxtrav, rng.fn, rng_user=rng.user)
        t_gpu = time.time() - t0
        d_gpu = int(draws_of(rng, 4321))
        seed_fn(4321)
        t0 = time.time()
        rb_ref = ref.ras(777, args.maxtrav)
        t_ref = time.time() - t0
        wbn, wbs = ref.get_ring()
        out["ras"] = {"what": "one RAS tree + SPR rounds (mpgpu_stepwise_addition, host buffers) vs the reference on one core",
                      "wall_s": t_gpu, "insertions": int(rins2), "final_score": int(rb), "reference_wall_s": t_ref,
                      "identical_result": bool(rb == rb_ref and d_gpu == int(draws_fn()) and np.array_equal(rbn[3:], wbn[3:])
                                               and np.array_equal(rbs[3:], wbs[3:])),
                      "speedup_vs_one_core": t_ref / t_gpu}
        eng.set_tree(case["bn"], case["bs"])
    return out


def measured_profile(workload, kernel):
    """ncu --set full numbers of `kernel` on `workload` (profiles/traffic.json): DRAM bytes and executed warp instructions
    per launch -- counts taken under the profiler, never timings."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[workload][kernel]
    except Exception:
        return {}


def roofline_of(workload, world, alg_bytes, ker_s, peak, peak_src, n_cand, words):
    """k_spr_scan is bounded by the ALU pipe (LOP3 / ISETP / SHF / IADD3 issue at one warp instruction per 2 cycles per SM
    sub-partition, B300_MICROARCH.md "Pipe rates"): its working set is L2-resident and every view is read once per prune
    task, so the canonical-bytes HBM figure exceeds 1 and says nothing about utilisation.  Reported: the instruction
    roofline (ALU-pipe warp instructions per launch from the committed ncu capture / measured kernel time, against
    SMs x 4 x 0.5 x SM clock) AND the canonical-bytes fraction the contract defines (8*S*W bytes per insertion)."""
    hbm = {"achieved": alg_bytes / ker_s / 1e9, "peak": peak, "unit": "GB/s", "frac": alg_bytes / ker_s / 1e9 / peak,
           "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes}
    prof = measured_profile(workload, "k_spr_scan") if world == 1 else {}
    out = {"kernel": "k_spr_scan", "kernel_ms": ker_s * 1e3, "traffic": prof.get("bytes"), "canonical_hbm": hbm}
    alu = prof.get("alu_inst")
    if alu:
        try:
            mhz = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["sm_max_mhz"])
        except Exception:
            mhz = 1965.0
        sms = 148
        peak_alu = sms * 4 * 0.5 * mhz * 1e6 / 1e9                 # G warp-instructions/s the ALU pipes can issue
        ach = alu / ker_s / 1e9
        chunks = n_cand * (words // 32)
        out.update({"bound": "alu", "achieved": ach, "peak": peak_alu, "unit": "G warp-inst/s (ALU pipe)", "frac": ach / peak_alu,
                    "peak_source": "148 SMs x 4 sub-partitions x 0.5 inst/clk x %.0f MHz (sm_max_mhz)" % mhz,
                    "alu_warp_inst_per_launch": alu, "warp_inst_per_launch": prof.get("inst"),
                    "warp_inst_per_insertion_chunk": (prof.get("inst") / chunks) if prof.get("inst") else None,
                    "profile_source": prof.get("source")})
    else:
        out.update({"bound": "hbm", "achieved": hbm["achieved"], "peak": peak, "unit": "GB/s", "frac": hbm["frac"], "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes})
    return out


BB1000_CASE = "c1_100x5000"


def start_bb1000_stock(args):
    """Section bb1000 (third metric of BASELINE.json: -bb 1000 wall time through the unchanged host): the unmodified
    reference program on one host core, started now so that it runs while the GPU sections do."""
    import subprocess as sp
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        import mpboot_dropin_check as dropin
    except Exception as e:
        return {"unavailable": "tools/mpboot_dropin_check.py: %r" % (e,)}
    stock, gpu = os.path.join(dropin.BIN, "mpboot-avx"), os.path.join(dropin.BIN, "mpboot-avx-gpu")
    if not (os.path.exists(stock) and os.path.exists(gpu)):
        return {"unavailable": "integration/_bin binaries are not built (integration/build.sh needs /root/reference)"}
    out = os.path.join(ROOT, "gpurun_out", "bb1000")
    os.makedirs(out, exist_ok=True)
    aln = dropin.make_alignment(BB1000_CASE, out)
    pre = os.path.join(out, BB1000_CASE + ".bb.stock")
    t0 = time.time()
    proc = sp.Popen([stock, "-s", aln, "-seed", "1", "-bb", "1000", "-pre", pre], stdout=sp.PIPE, stderr=sp.STDOUT, text=True)
    return {"dropin": dropin, "proc": proc, "t0": t0, "aln": aln, "out": out, "stock_prefix": pre, "gpu": gpu}


def finish_bb1000(job, args):
    import re
    if "unavailable" in job:
        return job
    dropin = job["dropin"]
    pre = os.path.join(job["out"], BB1000_CASE + ".bb.gpu")
    so, _ = job["proc"].communicate(timeout=3600)            # the stock run first: the patched program is then timed on a quiet host
    stock_wall = time.time() - job["t0"]
    runs = [dropin.run_binary(job["gpu"], job["aln"], pre, ["-bb", "1000"], 1800) for _ in range(2)]
    g = min(runs, key=lambda r: r["search_wall_s"] if r["search_wall_s"] else 1e9)
    m = re.search(r"Wall-clock time used for tree search: ([0-9.]+) sec", so)
    stock_search = float(m.group(1)) if m else None
    same = {}
    for ext in (".treefile", ".contree", ".splits.nex"):
        try:
            same[ext] = open(job["stock_prefix"] + ext, "rb").read() == open(pre + ext, "rb").read()
        except OSError:
            same[ext] = False
    n, L = dropin.CASES[BB1000_CASE][0], dropin.CASES[BB1000_CASE][1]
    return {"what": "mpboot -s <aln> -seed 1 -bb 1000 through the unchanged MPBoot host: integration/_bin/mpboot-avx-gpu (the reference "
                    "program + integration/mpboot_gpu.patch + libmpgpu) against integration/_bin/mpboot-avx (the unmodified reference, "
                    "AVX build, one core) on the same box; times are the program's own 'Wall-clock time used for tree search'",
            "workload": "synthetic DNA %d taxa x %d sites (C1's largest fixture)" % (n, L),
            "stock_search_wall_s": stock_search, "gpu_search_wall_s": g["search_wall_s"],
            "speedup": (stock_search / g["search_wall_s"]) if (stock_search and g["search_wall_s"]) else None,
            "gpu_search_wall_s_runs": [r["search_wall_s"] for r in runs],
            "gpu_process_wall_s": g["process_wall_s"], "stock_process_wall_s_upper_bound": stock_wall,
            "identical_outputs": same, "best_score": g["best_score"], "gpu_stats": g["stats"], "gpu_rc": g["rc"]}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class DevCounts:
    """Zero-copy torch view of the library's device count vector (for the NCCL all-reduce)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 2}


def cpu_baseline(case, maxtrav, budget_s=20.0):
    """The reference's own engine (oracle/_ref) -- or the C port when it is not built -- timed on
    one host core on a bounded sample of the same workload: as many whole node visits of the same
    sweep as fit in ~budget_s."""
    from oracle import portlib, reflib
    if reflib.available():
        eng = reflib.RefEngine(case["chars"], case["weights"], case["datatype"], n_informative=case["n_inf"])
        kind = "reference"
    else:
        eng = portlib.OracleEngine(case["codes"], case["weights"], case["datatype"])
        kind = "port"
    eng.set_ring(case["bn"], case["bs"])
    eng.allocate(per_site=True)                 # per-site mode only so that the recorder can count insertions
    s0 = eng.evaluate_full(per_site=True)
    n = case["n"]
    # count insertions per visit first (untimed, recorder on) ...
    counts = []
    t0 = time.time()
    for i in range(1, 2 * n - 1):
        eng.record(False)
        eng.rearrange(i, 1, maxtrav, True, s0)
        counts.append(len(eng.saved()) - 1)
        if time.time() - t0 > budget_s:
            break
    nvis = len(counts)
    # ... then time the same visits in plain mode (perSiteScores = 0, the plain-search kernel)
    eng.set_ring(case["bn"], case["bs"])
    eng.allocate(per_site=False)
    s0 = eng.evaluate_full(per_site=False)
    t0 = time.time()
    for i in range(1, nvis + 1):
        eng.rearrange(i, 1, maxtrav, False, s0)
    dt = time.time() - t0
    ins = int(sum(counts))
    return ins, dt, kind, "%d of %d node visits of the same sweep (%d insertions), plain mode" % (nvis, 2 * n - 2, ins)


def bench_cost(case, order, args, flush, stream, local):
    """the same sweep under -cost (Sankoff weighted parsimony, transitions 1 / transversions 2
    for DNA, random symmetric costs 1..4 otherwise) on a context of its own.  Device-timed step = memset of the
    per-(candidate, segment) sums + k_sk_scan (every insertion of the sweep); e2e = mpgpu_scan_visits with host
    buffers (plan, H2D, k_sk_scan, k_sk_finish, D2H).  CPU baseline = the reference's own Sankoff kernels
    (evaluateSankoff.../newviewSankoff...SIMD, Vec16us) on one host core, same visits, bounded sample."""
    import torch
    from mpboot_b200 import engine
    from oracle import portlib, reflib
    n, ninf, dt = case["n"], case["n_inf"], case["datatype"]
    eng = engine.Engine(device=local, stream=stream)
    eng.load_alignment(case["codes"], case["weights"], dt)
    eng.set_tree(case["bn"], case["bs"])
    S = eng.S
    seg = do_segmenting(eng.pattern_parsimony()[0][:ninf], case["weights"], ninf)
    if len(seg) == 0 or seg[-1] != ninf:
        seg = np.append(seg[seg < ninf], ninf).astype(np.int32)
    if S == 4:
        cost = np.array([[0, 2, 1, 2], [2, 0, 2, 1], [1, 2, 0, 2], [2, 1, 2, 0]], dtype=np.uint32)
    else:
        r = np.random.default_rng(7).integers(1, 5, size=(S, S)); cost = np.minimum(r, r.T); np.fill_diagonal(cost, 0)
        cost = cost.astype(np.uint32)
    eng.set_cost_matrix(cost, seg)
    score = eng.tree_score()
    nvis = 2 * n - 2
    n_cand, n_tasks = eng.scan_plan(order, 1, nvis, 1, args.maxtrav)
    steps = max(3, min(args.steps, 10))
    for _ in range(3):
        flush.zero_(); eng.scan_launch()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(steps):
        flush.zero_()
        ev[k][0].record(); eng.scan_launch(); ev[k][1].record()
    torch.cuda.synchronize()
    ker_s = sum(a.elapsed_time(b) for a, b in ev) * 1e-3 / steps
    vb, mp, _, _ = eng.scan_finish(n_cand, nvis)
    t0 = time.time()
    for _ in range(3):
        vb2, mp2, _, _ = eng.scan_visits(order, 1, nvis, 1, args.maxtrav, capacity=n_cand + 16)
    e2e_s = (time.time() - t0) / 3
    assert np.array_equal(mp, mp2)
    L, lb = eng.sankoff_layout()
    alg_bytes = 4.0 * S * L * n_cand                 # canonical: the Q and R cost vectors (u16 [L][S]) of every insertion
    peak, peak_src = measured_peak()
    pairs = (L + 63) // 64 * 32
    out = {"what": "the same sweep under -cost (Sankoff), %d segments" % len(seg), "tree_score": int(score),
           "insertions_per_step": int(n_cand), "ms_per_step": ker_s * 1e3, "insertions_per_s": n_cand / ker_s,
           "minplus_u16x2_ops_per_s": n_cand / ker_s * pairs * S * S,
           "roofline": {"bound": "hbm", "kernel": "k_sk_scan", "achieved": alg_bytes / ker_s / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": alg_bytes / ker_s / 1e9 / peak, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                        "kernel_ms": ker_s * 1e3, "traffic": measured_traffic(args.workload, "k_sk_scan"),
                        "note": "canonical accounting 4*S*L bytes per insertion; the kernel is bound by the DPX/ALU pipe "
                                "(S*S VIADDMNMX.U16x2 per pattern pair per insertion), see profiles/"},
           "e2e": {"insertions_per_s": n_cand / e2e_s, "ms_per_step": e2e_s * 1e3,
                   "h2d_bytes_per_step": int(eng.scan_plan_bytes()), "d2h_bytes_per_step": 8 * int(n_cand)}}
    if not args.no_search:
        rng = engine.HostRng(12345)
        t0 = time.time()
        ret, bn, bs, nins = eng.optimize_spr(case["bn"], case["bs"], rng.fn, 1, args.maxtrav, rng_user=rng.user)
        out["search"] = {"what": "mpgpu_optimize_spr under -cost from the random tree (early exit replayed)", "wall_s": time.time() - t0,
                         "insertions": int(nins), "final_score": int(ret)}
    if not args.no_bb:
        # -cost with -bb: every insertion's per-pattern cost vector (k_sk_scan<ROWS>) against B replicates (k_sk_reps, exact u16 weights)
        os.environ.setdefault("MPGPU_REPS_ROW_BYTES", str(6 << 30))
        boot = make_replicates(case, args.replicates)
        eng.set_option("reps_timing", 1)
        eng.load_replicates(boot, seg)
        eng.set_tree(case["bn"], case["bs"])
        eng.scan_plan(order, 1, nvis, 1, args.maxtrav)
        calls = []
        for v in range(nvis):
            calls.append(-1); calls.extend(range(vb[v], vb[v + 1]))
        calls = np.array(calls, dtype=np.int32)
        eng.scan_launch(); eng.reps_candidates_device(calls)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.zero_()
        e0.record(); eng.scan_launch(); eng.reps_candidates_device(calls); e1.record()
        torch.cuda.synchronize()
        bb_s = e0.elapsed_time(e1) * 1e-3
        macs = float(len(np.unique(calls))) * L * args.replicates
        tch, ech = eng.sankoff_reps_stats()
        out["bb"] = {"what": "the same sweep under -cost -bb, cutoff off: every insertion's per-pattern cost vector x %d replicates "
                             "(k_sk_scan<ROWS> + u8 pack + wrap proof + contraction)" % args.replicates,
                     "calls_per_step": int(len(calls)), "ms_per_step": bb_s * 1e3, "reps_vectors_per_s": len(calls) / bb_s,
                     "mac_per_s": macs / bb_s, "chunks_on_tensor_cores": int(tch), "chunks_on_exact_kernel": int(ech)}
        if tch:
            tc_ms, rows, npat, splits = eng.reps_timing()
            ops = 2.0 * rows * npat * args.replicates
            peak_tc = 2.0 * float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops", 1650.0)) \
                if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 3300.0
            peak_tc_src = "2 x MEASURED_PEAKS.json bf16_tflops (no int8 figure is measured)"
            try:
                i8 = max(eng.int8_peak(4096) for _ in range(3))
                if i8 > 0:
                    peak_tc, peak_tc_src = i8, "measured in this run: mpgpu_int8_peak (tcgen05.mma kind::i8 issued back to back, one CTA per SM)"
            except Exception:                                     # noqa: BLE001
                pass
            out["bb"]["roofline"] = {"bound": "tensor", "kernel": "k_reps_tc<BYTES>", "achieved": ops / (tc_ms * 1e-3) / 1e12,
                                     "peak": peak_tc, "unit": "TOP/s (int8, s32 accumulate)", "frac": ops / (tc_ms * 1e-3) / 1e12 / peak_tc,
                                     "kernel_ms": tc_ms, "shape": "%d rows x %d patterns x %d replicates, K splits %d"
                                                                  % (rows, npat, args.replicates, splits),
                                     "peak_source": peak_tc_src}
        if not args.no_cpu_baseline:
            use_ref = reflib.available()
            ref = (reflib.RefEngine(case["chars"], case["weights"], dt, n_informative=ninf) if use_ref
                   else portlib.OracleEngine(case["codes"], case["weights"], dt))
            ref.set_cost_matrix(cost, seg)
            ref.set_ring(case["bn"], case["bs"])
            ref.allocate(per_site=True)
            s0 = ref.evaluate_full(per_site=True)
            ref.boot_init(boot, seg, 0.0, 0.5, None)
            nv, t0 = 0, time.time()
            for i in range(1, nvis + 1):
                ref.rearrange(i, 1, args.maxtrav, True, s0)
                nv += 1
                if time.time() - t0 > args.cpu_budget:
                    break
            dt_s = time.time() - t0
            cnt = ref.boot_counters()
            out["bb"]["cpu_baseline"] = {"reps_vectors_per_s": cnt[2] / dt_s, "cores": 1, "kind": "reference" if use_ref else "port",
                                         "sample": "%d of %d node visits (%d REPS vectors x %d replicates) in %.1f s"
                                                   % (nv, nvis, cnt[2], args.replicates, dt_s)}
    if not args.no_cpu_baseline:
        use_ref = reflib.available()
        ref = (reflib.RefEngine(case["chars"], case["weights"], dt, n_informative=ninf) if use_ref
               else portlib.OracleEngine(case["codes"], case["weights"], dt))
        ref.set_cost_matrix(cost, seg)
        ref.set_ring(case["bn"], case["bs"])
        ref.allocate(per_site=True)
        s0 = ref.evaluate_full(per_site=True)
        assert s0 == score, "reference and device disagree on the -cost tree score"
        ins, nv = 0, 0
        t0 = time.time()
        for i in range(1, nvis + 1):
            ref.record(False)
            ref.rearrange(i, 1, args.maxtrav, True, s0)
            got = ref.saved()[1:]
            assert np.array_equal(got, mp[vb[i - 1]: vb[i]].astype(np.int32)), "insertion scores differ from the reference"
            ins += len(got); nv += 1
            if time.time() - t0 > args.cpu_budget:
                break
        dt_s = time.time() - t0
        out["cpu_baseline"] = {"insertions_per_s": ins / dt_s, "cores": 1, "kind": "reference" if use_ref else "port",
                               "sample": "%d of %d node visits of the same sweep (%d insertions, exact mode, every score checked "
                                         "against the device) in %.1f s" % (nv, nvis, ins, dt_s)}
    eng.close()
    return out


def make_replicates(case, B, seed=5):
    """boot_samples_pars as MPBoot draws them: multinomial resampling of the sites, counted per pattern
    (alignment.cpp:1985-1990, iqtree.cpp:285-313)."""
    rng = np.random.default_rng(seed)
    w = case["weights"].astype(np.float64)
    return rng.multinomial(int(w.sum()), w / w.sum(), size=B).astype(np.uint16)


def bench_bb_sharded(eng, case, order, vb, n_cand, args, flush, world, rank, local, stream, mode):
    """-bb at N > 1 (every rank calls this), the same sweep with the cutoff off, device-timed as the max over ranks, then a check
    of REPS vectors against an unsharded context on rank 0.
    mode "replicates" (north_star: replicates are sharded): every GPU holds the WHOLE C2 alignment and scores every candidate, but
      contracts only its B / N replicates (mpgpu_set_replicate_shards) -- strong scaling of the 1-GPU bb step, nothing exchanged
      on the device path (the searches exchange the rows of calls that can change a replicate);
    mode "patterns": the pattern-sharded context of the main line -- each shard contracts its slice of the patterns and the
      row x replicate partial sums (rows x 1024 s32) go through the exchange step."""
    import torch
    import torch.distributed as dist
    from mpboot_b200 import engine, sharded
    B = args.replicates
    if mode == "replicates":
        case = build_case(args.workload, 1, "weak")                  # the single-GPU configuration, on every GPU
        eng = engine.Engine(device=local, stream=stream)
        eng.set_replicate_shards(rank, world)
        sharded.connect_peers(eng)
        eng.load_alignment(case["codes"], case["weights"], case["datatype"])
        eng.set_tree(case["bn"], case["bs"])
        order = eng.visit_order()
        vb, mp, _, _ = eng.scan_visits(order, 1, 2 * case["n"] - 2, 1, args.maxtrav)
        n_cand = len(mp)
    n, ninf = case["n"], case["n_inf"]
    boot = make_replicates(case, B)
    seg = do_segmenting(eng.pattern_parsimony()[0][:ninf], case["weights"], ninf)
    eng.load_replicates(boot, seg)
    calls = []
    for v in range(2 * n - 2):
        calls.append(-1); calls.extend(range(vb[v], vb[v + 1]))
    calls = np.array(calls, dtype=np.int32)
    eng.scan_plan(order, 1, 2 * n - 2, 1, args.maxtrav)
    steps = max(3, min(args.steps, 10))
    for _ in range(2):
        eng.scan_launch(); eng.reps_candidates_device(calls)
    dist.barrier(); torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(steps):
        flush.zero_()
        ev[k][0].record(); eng.scan_launch(); eng.reps_candidates_device(calls); ev[k][1].record()
    dist.barrier(); torch.cuda.synchronize()
    tt = torch.tensor([sum(a.elapsed_time(b) for a, b in ev) / steps], device="cuda", dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    step_ms = float(tt[0])
    sample = calls[: min(len(calls), 400)]
    got = eng.reps_candidates(sample)                         # collective: every rank ends up with the complete vectors
    out = None
    if rank == 0:
        one = engine.Engine(device=local, stream=stream)
        one.load_alignment(case["codes"], case["weights"], case["datatype"])
        one.set_tree(case["bn"], case["bs"])
        one.load_replicates(boot, do_segmenting(one.pattern_parsimony()[0][:ninf], case["weights"], ninf))
        one.scan_plan(order, 1, 2 * n - 2, 1, args.maxtrav)
        one.scan_launch()
        want = one.reps_candidates(sample)
        one.close()
        what = ("the C2 sweep (200 x 100000, the 1-GPU configuration: strong scaling) under -bb with the %d replicates sharded over %d GPUs, cutoff off: "
                "every GPU scans all candidates and contracts its B / N replicates" % (B, world)) if mode == "replicates" else \
               ("the sweep of the main line under -bb on %d pattern shards, cutoff off: scan + delta rows + contraction of each shard's patterns + "
                "in-library exchange of the row x replicate sums + combine" % world)
        out = {"what": what, "replicates": B, "calls_per_step": int(len(calls)), "ms_per_step": step_ms,
               "reps_vectors_per_s": len(calls) / (step_ms * 1e-3), "insertions_per_s": n_cand / (step_ms * 1e-3),
               "check": {"reps_vectors_equal_unsharded": bool(np.array_equal(got, want)), "vectors_compared": int(len(sample))}}
        assert out["check"]["reps_vectors_equal_unsharded"], "sharded REPS vectors differ from the unsharded context"
    if mode == "replicates":
        eng.close()
    return out


def bench_bb(eng, case, order, vb, n_cand, args, flush):
    """The same sweep under -bb with the cutoff off: every scored insertion and, once per node visit, the
    current tree go through REPS against B = 1000 replicates (iqtree.cpp:3411-3449) -- the worst case for the
    replicate contraction.  Device-timed step = k_spr_scan + delta rows + tcgen05 contraction + combine;
    search = one whole mpgpu_optimize_spr_bb call (host bookkeeping, RNG replay and moves included)."""
    import torch
    from mpboot_b200 import engine
    n, B, ninf = case["n"], args.replicates, case["n_inf"]
    boot = make_replicates(case, B)
    seg = do_segmenting(eng.pattern_parsimony()[0][:ninf], case["weights"], ninf)
    eng.set_option("reps_timing", 1)
    eng.load_replicates(boot, seg)
    calls = []
    for v in range(2 * n - 2):
        calls.append(-1); calls.extend(range(vb[v], vb[v + 1]))
    calls = np.array(calls, dtype=np.int32)
    eng.scan_plan(order, 1, 2 * n - 2, 1, args.maxtrav)
    steps = max(3, min(args.steps, 10))
    for _ in range(3):
        eng.scan_launch(); eng.reps_candidates_device(calls)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    tc_ms = []
    for k in range(steps):
        flush.zero_()
        ev[k][0].record()
        eng.scan_launch(); eng.reps_candidates_device(calls)
        ev[k][1].record()
        torch.cuda.synchronize()
        ms, rows, pat, splits = eng.reps_timing()
        tc_ms.append(ms)
    step_ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    tc = float(np.median(tc_ms))
    ops = 2.0 * rows * ninf * B                                   # algorithmic int8 ops of the contraction
    peak_bf16 = None
    try:
        peak_bf16 = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
    except Exception:
        pass
    peak = 2.0 * (peak_bf16 if peak_bf16 else 1590.0)
    peak_src = "2 x %s bf16_tflops (no int8 figure is measured; nominal dense int8 is 4500)" % ("MEASURED_PEAKS.json" if peak_bf16 else "fallback")
    try:                                                          # the measured denominator: tcgen05.mma kind::i8 issued back to back
        i8 = max(eng.int8_peak(4096) for _ in range(3))
        if i8 > 0:
            peak_2bf16 = peak
            peak = i8
            peak_src = ("measured in this run: mpgpu_int8_peak (tcgen05.mma kind::i8 M128 N256 K32 back to back from resident "
                        "shared-memory tiles, one CTA per SM); 2 x bf16_tflops would be %.0f" % peak_2bf16)
    except Exception as e:                                        # noqa: BLE001
        peak_src += " [mpgpu_int8_peak failed: %s]" % e
    groups, exc, tensor = eng.reps_info()
    out = {
        "replicates": B, "patterns": ninf, "segments": int(len(seg)), "calls_per_step": int(len(calls)),
        "rows_per_step": int(rows), "ms_per_step": step_ms, "insertions_per_s": n_cand / (step_ms * 1e-3),
        "reps_vectors_per_s": len(calls) / (step_ms * 1e-3),
        "exact_path": {"column_groups": groups, "exception_patterns": exc, "tensor": tensor},
        "roofline": {"bound": "tensor", "kernel": "k_reps_tc", "achieved": ops / (tc * 1e-3) / 1e12, "peak": peak,
                     "unit": "TOP/s (int8, s32 accumulate)", "frac": ops / (tc * 1e-3) / 1e12 / peak,
                     "peak_source": peak_src,
                     "kernel_ms": tc, "shape": "%d rows x %d patterns x %d replicates, K splits %d" % (rows, pat, B, splits),
                     "algorithmic_ops_per_launch": ops, "traffic": measured_traffic(args.workload, "k_reps_tc")},
    }
    # the whole -bb SPR search from the same random tree, cutoff off (best of 2: the first call on a context also pays for the
    # growth of the row / staging buffers and the first launches of the ROWS kernels, which MPBoot amortises over its ~1000 searches)
    from mpboot_b200.engine import HostRng, Treels
    walls = []
    for _ in range(2):
        rs = HostRng(11)
        bl = np.full(B, -float(np.iinfo(np.int64).max)); bc = np.zeros(B, dtype=np.int32); bt = np.full(B, -1, dtype=np.int32)
        tl = Treels(n)
        t0 = time.time()
        ret, bn2, bs2, nins, ncalls, nreps = eng.optimize_spr_bb(case["bn"], case["bs"], tl.hooks(rs.fn, rs.user), bl, bc, bt, 0.0, 0.5, 1, args.maxtrav)
        walls.append(time.time() - t0)
    dt = min(walls)
    out["search"] = {"what": "mpgpu_optimize_spr_bb from the random tree, cutoff off, host buffers (e2e), best of 2", "wall_s": dt,
                     "first_call_wall_s": walls[0],
                     "insertions": int(nins), "savecurrenttree_calls": int(ncalls), "reps_vectors": int(nreps),
                     "insertions_per_s": nins / dt, "reps_vectors_per_s": nreps / dt, "final_score": int(ret),
                     "replicates_won": int((bt >= 0).sum()), "trees_materialised": int(len(tl.materialized()))}
    return out, boot, seg


def cpu_baseline_bb(case, boot, seg, maxtrav, budget_s=20.0):
    """Reference arm of the -bb step: rearrangeParsimony with per-site scores, every saveCurrentTree call
    through pllComputePatternParsimony + the Vec16us REPS loop + the default bookkeeping (oracle/_ref), on as
    many node visits of the same sweep as fit in ~budget_s on one core."""
    from oracle import portlib, reflib
    if reflib.available():
        eng = reflib.RefEngine(case["chars"], case["weights"], case["datatype"], n_informative=case["n_inf"]); kind = "reference"
    else:
        eng = portlib.OracleEngine(case["codes"], case["weights"], case["datatype"]); kind = "port"
    eng.set_ring(case["bn"], case["bs"])
    eng.allocate(per_site=True)
    s0 = eng.evaluate_full(per_site=True)
    eng.boot_init(boot, seg, 0.0, 0.5, None)
    n = case["n"]
    eng.record(False)
    t0 = time.time()
    nvis = 0
    for i in range(1, 2 * n - 1):
        eng.rearrange(i, 1, maxtrav, True, s0)
        nvis += 1
        if time.time() - t0 > budget_s:
            break
    dt = time.time() - t0
    calls = eng.boot_counters()[0]
    ins = calls - nvis
    return {"insertions_per_s": ins / dt, "reps_vectors_per_s": calls / dt, "cores": 1, "kind": kind,
            "sample": "%d of %d node visits of the same sweep (%d insertions, %d REPS vectors x %d replicates) in %.1f s"
                      % (nvis, 2 * n - 2, ins, calls, boot.shape[0], dt)}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle/_ref when built
    here, else the C port) on the same workload; one process per host core, each running the
    same bounded sample, because the reference path is single-threaded and not re-entrant."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    world = max(1, args.gpus)
    case = build_case(args.workload, world, args.scaling)        # the same alignment our arm shards over `world` GPUs
    ncores = max(1, len(os.sched_getaffinity(0)))
    nproc = min(ncores, 64)
    budget = 6.0
    # dlopen the reference library in THIS process before the fork, so that the driver's loaded-library record of the
    # reference arm shows oracle/_ref/libmpref.so (the workers inherit the mapping)
    from oracle import reflib
    if reflib.available():
        reflib.lib()
    with mp.get_context("fork").Pool(nproc) as pool:
        t0 = time.time()
        res = pool.starmap(_ref_worker, [(case, args.maxtrav, budget, args.steps, args.warmup)] * nproc)
        wall = time.time() - t0
    ins = sum(r[0] for r in res)
    elapsed = max(r[1] for r in res)
    kind, sample = res[0][2], res[0][3]
    sites = int(np.asarray(case["weights"][: case["n_inf"]], dtype=np.int64).sum())   # the engine works on the informative sites, expanded by weight (compressDNA)
    ins_per_s = ins / elapsed
    value = ins_per_s * 2.0 * sites
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / max(args.steps, 1),
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "insertions_per_s": ins_per_s,
        "config": {"workload": workload_name(args.workload, world, case), "maxtrav": args.maxtrav, "processes": nproc},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nproc, "kind": kind, "sample": sample + " per step, x%d processes" % nproc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": wall,
        "worker_kinds": sorted(set(r[2] for r in res)),
    }
    print(json.dumps(line), flush=True)


def _ref_worker(case, maxtrav, budget_s, steps, warmup):
    from oracle import portlib, reflib
    if reflib.available():
        eng = reflib.RefEngine(case["chars"], case["weights"], case["datatype"], n_informative=case["n_inf"]); kind = "reference"
    else:
        eng = portlib.OracleEngine(case["codes"], case["weights"], case["datatype"]); kind = "port"
    n = case["n"]
    # insertions per node visit depend on the topology only: count them on the first columns (cheap, recorder on)
    small = portlib.OracleEngine(np.ascontiguousarray(case["codes"][:, :256]), np.ascontiguousarray(case["weights"][:256]), case["datatype"])
    small.set_ring(case["bn"], case["bs"])
    small.allocate(per_site=True)
    ss = small.evaluate_full(per_site=True)
    all_counts = []
    for i in range(1, 2 * n - 1):
        small.record(False)
        small.rearrange(i, 1, maxtrav, True, ss)
        all_counts.append(len(small.saved()) - 1)
    eng.set_ring(case["bn"], case["bs"])
    eng.allocate(per_site=False)
    s0 = eng.evaluate_full(per_site=False)
    nvis = 0
    t0 = time.time()
    for i in range(1, 2 * n - 1):               # untimed warm-up pass that also sizes the bounded sample
        eng.rearrange(i, 1, maxtrav, False, s0)
        nvis += 1
        if time.time() - t0 > budget_s:
            break
    counts = all_counts[:nvis]
    t0 = time.time()
    for _ in range(steps):
        for i in range(1, nvis + 1):
            eng.rearrange(i, 1, maxtrav, False, s0)
    dt = time.time() - t0
    ins = int(sum(counts)) * steps
    return ins, dt, kind, "%d of %d node visits of the sweep (%d insertions)" % (nvis, 2 * n - 2, int(sum(counts)))


def workload_name(name, nshards, case):
    n, sites, dt, mu, seed, tseed = WORKLOADS[name]
    return "%s: synthetic %s %d taxa x %d sites (%d per GPU), SPR sweep radius 1..6 on a fixed random tree" % (
        name, {0: "BIN", 1: "DNA", 2: "AA", 6: "MORPH32"}[dt], n, case["sites"], case["sites"] // nshards)


def measure_sweep(args, workload, scaling, steps, world, rank, local, stream, flush, want_clocks):
    """One workload's SPR sweep on `world` GPUs: device-timed step (kernel + in-library exchange step), the kernel alone,
    the end-to-end call with host buffers, and -- at world > 1 -- a check of every reduced insertion score against an
    unsharded context that rank 0 builds over the whole alignment."""
    import torch
    import torch.distributed as dist
    from mpboot_b200 import engine, sharded

    case = build_case(workload, world, scaling)
    n = case["n"]
    eng = engine.Engine(device=local, stream=stream, shard_rank=rank, shard_count=world)
    if world > 1:
        sharded.connect_peers(eng)               # the exchange step lives in the library: one-shot all-reduce over NVLink peer memory
    eng.load_alignment(case["codes"], case["weights"], case["datatype"])
    eng.set_tree(case["bn"], case["bs"])
    order = eng.visit_order()
    nvis = 2 * n - 2
    n_cand, n_tasks = eng.scan_plan(order, 1, nvis, 1, args.maxtrav)
    S, W = eng.S, eng.shard_words

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(nsteps, exchange):
        """nsteps launches bracketed by CUDA events on the library's stream; exchange = False times the scan kernel alone
        (option "exchange" 0: the shard's partial counts stay partial)."""
        if world > 1:
            eng.set_option("exchange", 1 if exchange else 0)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nsteps)]
        barrier()
        t0 = time.time()
        for k in range(nsteps):
            flush.zero_()                        # L2 flush between timed iterations (outside the events)
            ev[k][0].record()
            eng.scan_launch()
            ev[k][1].record()
        barrier()
        wall = time.time() - t0
        ms = sum(a.elapsed_time(b) for a, b in ev)
        if world > 1:
            tt = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt[0])
            eng.set_option("exchange", 1)
        return ms, wall

    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        eng.scan_launch()
    barrier()
    launches0 = eng.launch_count()
    sampler = ClockSampler(local) if want_clocks else None
    if sampler:
        sampler.start()
    dev_ms, t_wall = timed(steps, True)
    clocks = sampler.stop() if sampler else None
    launches = eng.launch_count() - launches0
    ker_ms = timed(steps, False)[0] if world > 1 else dev_ms
    eng.scan_plan(order, 1, nvis, 1, args.maxtrav)                     # the timed loop without exchange left partial counts behind
    eng.scan_launch()
    vb, mp, cref, cprune = eng.scan_finish(n_cand, nvis)

    # e2e: the reference-facing call with host buffers (plan on host, H2D, kernel, exchange, D2H, finish)
    e2e_steps = max(3, steps)
    for _ in range(2):                           # untimed warm-up of the end-to-end path (page-locked buffers)
        eng.scan_visits(order, 1, nvis, 1, args.maxtrav, capacity=n_cand + 16, reuse=True)
    barrier()
    t0 = time.time()
    for _ in range(e2e_steps):                   # the caller's four output arrays are the same host buffers every step, as a C host's would be
        vb2, mp2, _, _ = eng.scan_visits(order, 1, nvis, 1, args.maxtrav, capacity=n_cand + 16, reuse=True)
    barrier()
    e2e_s = (time.time() - t0) / e2e_steps
    if world > 1:
        tt = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt[0])
    assert np.array_equal(mp, mp2), "device-resident and end-to-end paths disagree"

    check = None
    if world > 1:
        # every reduced score against an unsharded context over the whole alignment (rank 0), and the same vector on every rank
        h = torch.tensor([int(np.asarray(mp, dtype=np.int64).sum() % (1 << 62)), int(len(mp))], device="cuda", dtype=torch.int64)
        hs = [torch.empty_like(h) for _ in range(world)]
        dist.all_gather(hs, h)
        same_on_all = all(bool((x == hs[0]).all()) for x in hs)
        if rank == 0:
            one = engine.Engine(device=local, stream=stream)
            one.load_alignment(case["codes"], case["weights"], case["datatype"])
            one.set_tree(case["bn"], case["bs"])
            _, mp1, _, _ = one.scan_visits(order, 1, nvis, 1, args.maxtrav, capacity=n_cand + 16)
            s1 = one.tree_score()
            one.close()
            check = {"insertion_scores_equal_unsharded": bool(np.array_equal(mp1, mp)), "tree_score_equal_unsharded": bool(s1 == eng.tree_score()),
                     "identical_on_all_ranks": same_on_all, "insertions_compared": int(len(mp)),
                     "exchange_steps": int(eng.peer_stats()[0]), "exchange_timeouts": int(eng.peer_stats()[2])}
            assert check["insertion_scores_equal_unsharded"] and check["tree_score_equal_unsharded"] and same_on_all, check
        else:
            eng.tree_score()                     # the collective half of rank 0's call
    res = {"case": case, "eng": eng, "order": order, "n_cand": n_cand, "n_tasks": n_tasks, "S": S, "W": W, "nvis": nvis,
           "sites_total": eng.n_sites, "dev_ms": dev_ms, "ker_ms": ker_ms, "e2e_s": e2e_s, "launches": launches, "clocks": clocks,
           "wall_s": t_wall, "vb": vb, "mp": mp, "check": check, "barrier": barrier}
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # the drop-in's -bb 1000 run (section bb1000) starts the stock binary now, on a host core, so that its minutes
    # overlap the GPU sections below
    bb1000_job = start_bb1000_stock(args) if (rank == 0 and world == 1 and not args.no_bb1000) else None

    # a real (non-default) stream shared by torch and the library, so that torch.cuda.Event
    # brackets exactly the library's launches (the legacy default stream's handle is NULL,
    # which mpgpu_create takes as "make a private stream")
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    r = measure_sweep(args, args.workload, args.scaling, args.steps, world, rank, local, stream, flush, True)
    eng, case, order, n_cand, n_tasks, S, W, nvis = r["eng"], r["case"], r["order"], r["n_cand"], r["n_tasks"], r["S"], r["W"], r["nvis"]
    vb = r["vb"]
    sites_total = r["sites_total"]
    c4 = None
    if not args.no_c4 and args.workload == "c2":
        # north_star's scaling configuration: 1000 taxa x 1M sites, the SAME alignment sharded over the GPUs (strong)
        r4 = measure_sweep(args, "c4", "strong", max(3, args.steps // 2), world, rank, local, stream, flush, False)
        ops4 = 2.0 * r4["sites_total"]
        ms4 = r4["dev_ms"] / max(3, args.steps // 2)
        c4 = {"workload": workload_name("c4", world, r4["case"]), "scaling": "strong", "n_gpus": world,
              "ms_per_step": ms4, "kernel_ms": r4["ker_ms"] / max(3, args.steps // 2), "insertions_per_step": r4["n_cand"],
              "insertions_per_s": r4["n_cand"] / (ms4 * 1e-3), "value": r4["n_cand"] / (ms4 * 1e-3) * ops4, "unit": UNIT,
              "e2e_ms_per_step": r4["e2e_s"] * 1e3, "words_per_plane_per_gpu": r4["W"],
              "canonical_hbm_frac": 8.0 * r4["S"] * r4["W"] * r4["n_cand"] / (r4["ker_ms"] / max(3, args.steps // 2) * 1e-3) / 1e9 / measured_peak()[0],
              "check": r4["check"],
              "note": "speed-up over 1 GPU = this key's ms_per_step at n_gpus = 1 (same alignment) / ms_per_step here"}
        r4["eng"].close()
        del r4

    bb_sharded = bb_patterns = None
    if world > 1 and not args.no_bb:
        bb_sharded = bench_bb_sharded(eng, case, order, vb, n_cand, args, flush, world, rank, local, stream, "replicates")
        bb_patterns = bench_bb_sharded(eng, case, order, vb, n_cand, args, flush, world, rank, local, stream, "patterns")

    if rank == 0:
        ops_per_ins = 2.0 * sites_total
        ms_per_step = r["dev_ms"] / args.steps
        ins_per_s = n_cand / (ms_per_step * 1e-3)
        value = ins_per_s * ops_per_ins
        peak, peak_src = measured_peak()
        alg_bytes = 8.0 * S * W * n_cand                       # canonical 8*S*W bytes per insertion (this shard)
        ker_s = r["ker_ms"] * 1e-3 / args.steps
        achieved = alg_bytes / ker_s / 1e9
        e2e_s = r["e2e_s"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "insertions_per_s": ins_per_s, "insertions_per_step": n_cand,
            "config": {"workload": workload_name(args.workload, world, case), "maxtrav": args.maxtrav,
                       "states": S, "words_per_plane_per_gpu": W, "prune_tasks": n_tasks, "bb": False,
                       "l2": "flushed between timed steps (256 MiB memset)",
                       "parallelism": ("pattern-sharded x%d, in-library one-shot all-reduce of the int32 counts over NVLink peer memory "
                                       "(k_peer_allreduce, on the stream)" % world) if world > 1 else "single GPU"},
            "roofline": roofline_of(args.workload, world, alg_bytes, ker_s, peak, peak_src, n_cand, W),
            "e2e": {"value": (n_cand / e2e_s) * ops_per_ins, "unit": UNIT, "insertions_per_s": n_cand / e2e_s,
                    "h2d_bytes_per_step": None, "d2h_bytes_per_step": 4 * (n_cand + 2 * nvis), "ms_per_step": e2e_s * 1e3},
            "gpu_launches": int(r["launches"]), "clocks": r["clocks"], "wall_s": r["wall_s"],
        }
        if world > 1:
            line["exchange"] = {"what": "device time of the exchange step per sweep = step - scan kernel alone (max over ranks each)",
                                "ms_per_step": (r["dev_ms"] - r["ker_ms"]) / args.steps, "int32_per_step": n_cand + 2 * nvis,
                                "check": r["check"]}
        if c4:
            line["c4_strong"] = c4
        if bb_sharded:
            line["bb"] = bb_sharded
        if bb_patterns:
            line["bb_pattern_shards"] = bb_patterns
        line["e2e"]["h2d_bytes_per_step"] = int(eng.scan_plan_bytes())
        from mpboot_b200 import engine as _engine
        line["e2e"]["host_threads"] = int(_engine.lib().mpgpu_host_plan_threads())     # threads that enumerate the sweep's plan (MPGPU_PLAN_THREADS)
        if world == 1 and not args.no_search:
            line["search"] = bench_search(eng, case, args)
        if world == 1 and not args.no_bb:
            bb, boot, seg = bench_bb(eng, case, order, vb, n_cand, args, flush)
            if not args.no_cpu_baseline:
                bb["cpu_baseline"] = cpu_baseline_bb(case, boot, seg, args.maxtrav, budget_s=args.cpu_budget)
            line["bb"] = bb
        if world == 1 and not args.no_cost:
            line["cost"] = bench_cost(case, order, args, flush, stream, local)
        if world == 1 and not args.no_cpu_baseline:
            ins, dt, kind, sample = cpu_baseline(case, args.maxtrav, budget_s=args.cpu_budget)
            line["cpu_baseline"] = {"value": ins / dt * ops_per_ins, "unit": UNIT, "cores": 1, "kind": kind,
                                    "sample": sample, "insertions_per_s": ins / dt}
        if bb1000_job is not None:
            line["bb1000"] = finish_bb1000(bb1000_job, args)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--maxtrav", type=int, default=6)
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-bb", action="store_true", help="skip the -bb (replicate scoring) section")
    ap.add_argument("--no-cost", action="store_true", help="skip the -cost (Sankoff) section")
    ap.add_argument("--no-search", action="store_true", help="skip the whole-search (pllOptimizeSprParsimony) section")
    ap.add_argument("--no-c4", action="store_true", help="skip the C4 (1000 taxa x 1M sites, strong scaling) section")
    ap.add_argument("--no-bb1000", action="store_true", help="skip the drop-in section (MPBoot -bb 1000: patched vs stock binary)")
    ap.add_argument("--replicates", type=int, default=1000)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N>1: weak = the workload's sites per GPU (default), strong = the workload's sites in total, sharded")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
