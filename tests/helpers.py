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
