"""Runs the two MPBoot binaries of integration/_bin -- the unmodified reference (mpboot-avx) and the same program
with the parsimony path on the device (mpboot-avx-gpu, integration/mpboot_gpu.patch) -- on seeded synthetic
alignments, compares .treefile / .contree / .splits.nex byte for byte and reports the program's own
"Wall-clock time used for tree search" (phyloanalysis.cpp:1647-1648) of each.

    python tools/mpboot_dropin_check.py [--cases c1_12x300,c1_17x1998] [--modes plain,bb] [--out gpurun_out/x1]
                                        [--golden tests/golden/mpboot]   (compare against committed outputs instead of
                                                                           running the stock binary)
One JSON line per (case, mode) on stdout.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpboot_b200 import synth  # noqa: E402

BIN = os.path.join(ROOT, "integration", "_bin")

# name -> (taxa, sites, datatype, mu, seed, extra CLI)
CASES = {
    "c1_12x300": (12, 300, synth.PLL_DNA_DATA, 0.05, 11, []),
    "c1_17x1998": (17, 1998, synth.PLL_DNA_DATA, 0.05, 1, []),
    "c1_100x5000": (100, 5000, synth.PLL_DNA_DATA, 0.05, 12, []),
    "aa_20x600": (20, 600, synth.PLL_AA_DATA, 0.08, 13, ["-st", "AA"]),
    "morph_16x400": (16, 400, synth.PLL_GENERIC_32, 0.05, 14, ["-st", "MORPH"]),
    "c2_200x100000": (200, 100000, synth.PLL_DNA_DATA, 0.05, 2, []),
    # the other replicate-bookkeeping policies of IQTree::saveCurrentTree (iqtree.cpp:3498-3583), same alignment as c1_17x1998
    "mulhits_17x1998": (17, 1998, synth.PLL_DNA_DATA, 0.05, 1, ["-mulhits"]),
    "topboot_17x1998": (17, 1998, synth.PLL_DNA_DATA, 0.05, 1, ["-mulhits", "-topboot", "3"]),
    "mulhits_aa_20x600": (20, 600, synth.PLL_AA_DATA, 0.08, 13, ["-st", "AA", "-mulhits"]),
    # -distinct_iter_top_boot K (iqtree.cpp:3587-3685) on saturated alignments (supports well below 100, so the policy shows in the
    # outputs: they differ from the default policy's): 2 REPS segments (no segment in the skip test's range) / 5 segments
    "distinct_20x800": (20, 800, synth.PLL_DNA_DATA, 0.25, 21, ["-distinct_iter_top_boot", "3"]),
    "distinct_30x1500": (30, 1500, synth.PLL_DNA_DATA, 0.35, 23, ["-distinct_iter_top_boot", "2"]),
    # -cutoff_from_btrees (logl_cutoff from IQTree::boot_tree_orig_logl, iqtree.cpp:1657-1661, filled at :3524 / :3618 / :3717) under the
    # default, the -mulhits and the distinct-iteration policy; -min_iter1_cand (iteration 1 only extends treels_logl, :3404)
    "cutoffbt_30x1500": (30, 1500, synth.PLL_DNA_DATA, 0.35, 23, ["-cutoff_from_btrees"]),
    "cutoffbt_mulhits_20x800": (20, 800, synth.PLL_DNA_DATA, 0.25, 21, ["-mulhits", "-cutoff_from_btrees"]),
    "cutoffbt_distinct_20x800": (20, 800, synth.PLL_DNA_DATA, 0.25, 21, ["-distinct_iter_top_boot", "2", "-cutoff_from_btrees"]),
    "miniter1_20x800": (20, 800, synth.PLL_DNA_DATA, 0.25, 21, ["-min_iter1_cand"]),
    # -do_first_rell: REPS over the first half of the patterns only (iqtree.cpp:3426-3429), alone and under the distinct-iteration policy
    "firstrell_30x1500": (30, 1500, synth.PLL_DNA_DATA, 0.35, 23, ["-do_first_rell"]),
    "firstrell_distinct_30x1500": (30, 1500, synth.PLL_DNA_DATA, 0.35, 23, ["-do_first_rell", "-distinct_iter_top_boot", "2"]),
    # -autovec: the unsegmented plain-int REPS loop (iqtree.cpp:3418-3423), no skip test
    "autovec_30x1500": (30, 1500, synth.PLL_DNA_DATA, 0.35, 23, ["-autovec"]),
    # -cost (Sankoff weighted parsimony, ParsTree): transitions 1 / transversions 2; "@tstv" = a cost file written next to the alignment
    "cost_17x1998": (17, 1998, synth.PLL_DNA_DATA, 0.05, 1, ["-cost", "@tstv"]),
    # an asymmetric matrix (obeys the triangle inequality, so ParsTree::initCostMatrix leaves it alone): scores depend on the root
    "costasym_17x1998": (17, 1998, synth.PLL_DNA_DATA, 0.05, 1, ["-cost", "@asym"]),
    # -short_off: 32-bit Sankoff vectors and segment sums (tools.cpp:2365)
    "costu32_17x1998": (17, 1998, synth.PLL_DNA_DATA, 0.05, 1, ["-cost", "@tstv", "-short_off"]),
}
COST_FILES = {"@tstv": "4\n0 2 1 2\n2 0 2 1\n1 2 0 2\n2 1 2 0\n",
              "@asym": "4\n0 3 1 2\n2 0 3 1\n2 2 0 3\n3 2 2 0\n"}
MODES = {"plain": [], "bb": ["-bb", "1000"]}
OUTPUTS = {"plain": [".treefile"], "bb": [".treefile", ".contree", ".splits.nex"]}


def write_phylip(path, chars):
    n, L = chars.shape
    with open(path, "w") as f:
        f.write("%d %d\n" % (n, L))
        for i in range(n):
            f.write("T%-9d %s\n" % (i, chars[i].tobytes().decode()))


def make_alignment(name, outdir):
    n, L, dt, mu, seed, _ = CASES[name]
    path = os.path.join(outdir, name + ".phy")
    if not os.path.exists(path):
        gen = synth.evolve_alignment if n * L <= 50_000_000 else synth.evolve_alignment_blocked
        write_phylip(path, gen(n, L, dt, mu, seed))
    return path


def run_binary(binary, aln, prefix, extra, timeout):
    extra = list(extra)
    for i, a in enumerate(extra):
        if a in COST_FILES:
            path = os.path.join(os.path.dirname(aln), a[1:] + ".cost")
            with open(path, "w") as f:
                f.write(COST_FILES[a])
            extra[i] = path
    cmd = [binary, "-s", aln, "-seed", "1", "-pre", prefix] + extra
    env = dict(os.environ, MPBOOT_GPU_STATS="1")
    t0 = time.time()
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout, env=env)
    wall = time.time() - t0
    with open(prefix + ".stdout", "w") as f:
        f.write(p.stdout)
        f.write("\n--- stderr ---\n")
        f.write(p.stderr)
    m = re.search(r"Wall-clock time used for tree search: ([0-9.]+) sec", p.stdout)
    b = re.search(r"BEST SCORE FOUND : (\d+)", p.stdout)
    stats = [ln for ln in p.stderr.splitlines() if ln.startswith("[mpgpu]")]
    return {"rc": p.returncode, "process_wall_s": round(wall, 3), "search_wall_s": float(m.group(1)) if m else None,
            "best_score": int(b.group(1)) if b else None, "stats": stats[-1] if stats else None,
            "tail": (p.stdout[-600:] + p.stderr[-600:]) if p.returncode else None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="c1_12x300,c1_17x1998")
    ap.add_argument("--modes", default="plain,bb")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "x1"))
    ap.add_argument("--golden", default=None)
    ap.add_argument("--skip-gpu", action="store_true", help="only run the stock binary (to produce golden outputs)")
    ap.add_argument("--skip-stock", action="store_true", help="only run the patched binary (workloads the stock binary needs hours for)")
    ap.add_argument("--timeout", type=int, default=3600)
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    ok = True
    for name in args.cases.split(","):
        aln = make_alignment(name, args.out)
        for mode in args.modes.split(","):
            extra = CASES[name][5] + MODES[mode]
            line = {"case": name, "mode": mode}
            ref_prefix = None
            if args.golden:
                ref_prefix = os.path.join(args.golden, "%s.%s" % (name, mode))
            elif args.skip_stock:
                ref_prefix = None
            else:
                ref_prefix = os.path.join(args.out, "%s.%s.stock" % (name, mode))
                line["stock"] = run_binary(os.path.join(BIN, "mpboot-avx"), aln, ref_prefix, extra, args.timeout)
            if not args.skip_gpu:
                gpu_prefix = os.path.join(args.out, "%s.%s.gpu" % (name, mode))
                line["gpu"] = run_binary(os.path.join(BIN, "mpboot-avx-gpu"), aln, gpu_prefix, extra, args.timeout)
                same = {}
                for ext in (OUTPUTS[mode] if ref_prefix else []):
                    try:
                        same[ext] = open(ref_prefix + ext, "rb").read() == open(gpu_prefix + ext, "rb").read()
                    except OSError:
                        same[ext] = False
                line["identical"] = same
                if (ref_prefix and not all(same.values())) or line["gpu"]["rc"] != 0:
                    ok = False
                if "stock" in line and line["stock"]["search_wall_s"] and line["gpu"]["search_wall_s"]:
                    line["search_speedup"] = round(line["stock"]["search_wall_s"] / max(line["gpu"]["search_wall_s"], 1e-9), 2)
            print(json.dumps(line), flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
