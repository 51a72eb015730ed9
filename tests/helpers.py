"""Shared builders for the parity tests: seeded synthetic cases in the form both the reference
driver (oracle/_ref), the C oracle and the CUDA library accept."""
import numpy as np

from mpboot_b200 import encoding, synth


def informative_first(pat, w, codes, datatype):
    """Reorder patterns so the parsimony-informative ones form a prefix, as the reference host
    guarantees after optimizeAlignment's sort (phyloanalysis.cpp:2800-2816)."""
    und = {0: 3, 1: 15, 2: 22, 6: 32}[datatype]
    inf = np.zeros(pat.shape[1], dtype=bool)
    for j in range(pat.shape[1]):
        col = codes[:, j]
        inf[j] = len(np.unique(col[col < und])) >= 2
    order = np.concatenate([np.nonzero(inf)[0], np.nonzero(~inf)[0]])
    return pat[:, order], w[order], codes[:, order], int(inf.sum())


def make_case(n, nsites, datatype, seed, mu=0.05, gap=0.01, amb=0.001, tree_seed=None):
    chars = synth.evolve_alignment(n, nsites, datatype, mu, seed, gap=gap, amb=amb)
    pat, w = synth.compress_patterns(chars)
    codes = encoding.encode(pat, datatype)
    pat, w, codes, ninf = informative_first(pat, w, codes, datatype)
    rng = np.random.default_rng(1000 + seed if tree_seed is None else tree_seed)
    bn, bs = synth.random_tree_rings(n, rng)
    return dict(n=n, datatype=datatype, chars=pat, codes=codes, weights=w, n_inf=ninf, bn=bn, bs=bs)
