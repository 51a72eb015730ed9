"""X1: the drop-in inside the unchanged MPBoot host.

integration/_bin/mpboot-avx-gpu is the reference program (main, CLI, IQTree, SPRNG, candidate set, IQ-TREE-tree kernel,
summarizeBootstrap, output writers) built from /root/reference with integration/mpboot_gpu.patch, which routes the
sprparsimony entry points to libmpgpu.so.  Its .treefile / .contree / .splits.nex must be byte-identical to what the
unmodified reference (integration/_bin/mpboot-avx, same sources, same cmake line) writes for the same alignment and
-seed: tests/golden/mpboot/ holds those files (tools/make_golden_mpboot.sh).  This pins the replicate bookkeeping (R9)
and the supports to the real IQTree::saveCurrentTree / summarizeBootstrap, and -comppars pins the device score to the
IQ-TREE-tree kernel (R10, phylotree.cpp:692-1061)."""
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import mpboot_dropin_check as dropin  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "mpboot")
GPU_BIN = os.path.join(dropin.BIN, "mpboot-avx-gpu")

RUNS = [("c1_12x300", "plain"), ("c1_12x300", "bb"), ("c1_17x1998", "plain"), ("c1_17x1998", "bb"),
        ("aa_20x600", "plain"), ("aa_20x600", "bb"), ("morph_16x400", "plain"),
        ("c1_100x5000", "plain"), ("c1_100x5000", "bb"),
        ("mulhits_17x1998", "bb"), ("topboot_17x1998", "bb"), ("mulhits_aa_20x600", "bb"),
        ("distinct_20x800", "bb"), ("distinct_30x1500", "bb"),
        ("cutoffbt_30x1500", "bb"), ("cutoffbt_mulhits_20x800", "bb"), ("cutoffbt_distinct_20x800", "bb"), ("miniter1_20x800", "bb"),
        ("autovec_30x1500", "bb"), ("firstrell_30x1500", "bb"), ("firstrell_distinct_30x1500", "bb"),
        ("cost_17x1998", "plain"), ("cost_17x1998", "bb"), ("costasym_17x1998", "plain"), ("costasym_17x1998", "bb"),
        ("costu32_17x1998", "plain"), ("costu32_17x1998", "bb")]


def _need_binary():
    if not os.path.exists(GPU_BIN):
        pytest.fail("integration/_bin/mpboot-avx-gpu is missing: run integration/build.sh where /root/reference exists "
                    "(__graft_entry__.build() does)")


@pytest.mark.parametrize("case,mode", RUNS, ids=["%s-%s" % r for r in RUNS])
def test_outputs_identical_to_stock_mpboot(case, mode, tmp_path):
    _need_binary()
    aln = dropin.make_alignment(case, str(tmp_path))
    prefix = str(tmp_path / ("%s.%s.gpu" % (case, mode)))
    res = dropin.run_binary(GPU_BIN, aln, prefix, dropin.CASES[case][5] + dropin.MODES[mode], 1800)
    assert res["rc"] == 0, res["tail"]
    assert res["stats"] and "kernel launches 0" not in res["stats"]          # the device did the work
    for ext in dropin.OUTPUTS[mode]:
        want = open(os.path.join(GOLD, "%s.%s%s" % (case, mode, ext)), "rb").read()
        got = open(prefix + ext, "rb").read()
        assert got == want, "%s differs from the stock binary's output" % ext


@pytest.mark.parametrize("case", ["c1_17x1998", "aa_20x600", "morph_16x400"])
def test_comppars_two_kernels_one_score(case, tmp_path):
    """-comppars (sprparsimony.cpp:3613-3653): IQ-TREE's own kernel and the 'PLL kernel' -- here the device -- on a user tree."""
    _need_binary()
    aln = dropin.make_alignment(case, str(tmp_path))
    tree = os.path.join(GOLD, "%s.plain.treefile" % case)
    cmd = [GPU_BIN, "-s", aln, "-comppars", tree, "-pre", str(tmp_path / "cp")] + dropin.CASES[case][5]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-800:]
    iq = re.search(r"Parsimony score \(by IQTree kernel\) is: (\d+)", p.stdout)
    pll = re.search(r"Parsimony score \(by PLL kernel\) is: (\d+)", p.stdout)
    assert iq and pll, p.stdout[-800:]
    assert int(iq.group(1)) == int(pll.group(1)) > 0
