"""Shared builders for the parity tests: seeded synthetic cases in the form both the reference
driver (oracle/_ref), the C oracle and the CUDA library accept."""
import numpy as np

from mpboot_b200 import hostprep, synth


def make_case(n, nsites, datatype, seed, mu=0.05, gap=0.01, amb=0.001, tree_seed=None):
    chars = synth.evolve_alignment(n, nsites, datatype, mu, seed, gap=gap, amb=amb)
    prep = hostprep.prepare(chars, datatype)
    rng = np.random.default_rng(1000 + seed if tree_seed is None else tree_seed)
    bn, bs = synth.random_tree_rings(n, rng)
    return dict(n=n, datatype=datatype, chars=prep["chars"], codes=prep["codes"], weights=prep["weights"],
                n_inf=prep["n_inf"], bn=bn, bs=bs)


def make_boot(case, B, seed, ras_tree_scores=None, heavy=None):
    """Replicate pattern frequencies as MPBoot draws them (multinomial over all sites,
    alignment.cpp:1985-1990 -> boot_samples_pars[b][ptn], iqtree.cpp:285-313), u16 [B][P]."""
    rng = np.random.default_rng(9000 + seed)
    w = case["weights"].astype(np.float64)
    L = int(case["weights"].sum())
    boot = rng.multinomial(L, w / w.sum(), size=B).astype(np.uint16)
    if heavy is not None:                       # force u16 wraps / >255 weights in some replicates
        for (b, ptn, val) in heavy:
            boot[b, ptn] = val
    return boot


def segments_for(scores, weights, n_inf):
    """IQTree::doSegmenting (iqtree.cpp:3793) from per-pattern scores of the RAS tree."""
    from oracle import portlib
    P = len(weights)
    sc = np.zeros(P, dtype=np.int32); sc[: len(scores)] = scores
    return portlib.segments(sc, weights, n_inf)


def fingerprint_ring(bn, bs, n, move=None):
    """Topology fingerprint (same function as oracle/mp_oracle.c:mporacle_tree_fingerprint) of a
    ring table, optionally after regrafting pruned ref `move[0]` onto the branch at ref `move[1]`."""
    M = (1 << 64) - 1
    back = {}
    for node in range(1, 2 * n - 1):
        for s in range(1 if node <= n else 3):
            r = 3 * node + s
            back[r] = 3 * int(bn[r]) + int(bs[r])

    def nxt(r):
        node = r // 3
        return r if node <= n else 3 * node + (r % 3 + 1) % 3

    def hook(a, b):
        back[a] = b; back[b] = a

    if move is not None and move[0]:
        p, q = int(move[0]), int(move[1])
        a, b = back[nxt(p)], back[nxt(nxt(p))]
        hook(a, b)
        r = back[q]
        hook(nxt(p), q); hook(nxt(nxt(p)), r)

    def fin(z):
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        return z ^ (z >> 31)

    import sys
    sys.setrecursionlimit(max(10000, 4 * n))

    def th(r):
        if r // 3 <= n:
            return fin(r // 3)
        a, b = th(back[nxt(r)]), th(back[nxt(nxt(r))])
        if a > b:
            a, b = b, a
        return fin((a * 0x9E3779B97F4A7C15 + b + 0x632BE59BD9B4E019) & M)

    return fin(th(back[3]) ^ 0x1234567)
