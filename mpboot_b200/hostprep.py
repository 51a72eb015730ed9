"""Host-side input preparation (numpy), standing in for what the MPBoot host does before the
hot path starts: site->pattern compression (alignment.cpp addPattern), character coding
(pllBaseSubstitute) and the informative-patterns-first order that optimizeAlignment's sort
establishes (phyloanalysis.cpp:2800-2816).  Not part of the timed path."""
import numpy as np

from . import encoding, synth

UNDETERMINED = {synth.PLL_BINARY_DATA: 3, synth.PLL_DNA_DATA: 15, synth.PLL_AA_DATA: 22, synth.PLL_GENERIC_32: 32}


def informative_mask(codes, datatype):
    """isInformative (sprparsimony.cpp:2460): >= 2 distinct codes below `undetermined`."""
    und = UNDETERMINED[datatype]
    distinct = np.zeros(codes.shape[1], dtype=np.int32)
    for c in range(und):
        distinct += (codes == c).any(axis=0)
    return distinct >= 2


def prepare(chars, datatype, compress=True):
    """chars: uint8 ASCII [n][sites] -> dict(chars, codes, weights, n_inf) with informative
    patterns first.  compress=False keeps every site as its own pattern (weight 1)."""
    if compress:
        pat, w = synth.compress_patterns(chars)
    else:
        pat, w = chars, np.ones(chars.shape[1], dtype=np.int32)
    codes = encoding.encode(pat, datatype)
    inf = informative_mask(codes, datatype)
    order = np.concatenate([np.nonzero(inf)[0], np.nonzero(~inf)[0]])
    return dict(chars=np.ascontiguousarray(pat[:, order]), codes=np.ascontiguousarray(codes[:, order]),
                weights=np.ascontiguousarray(w[order]), n_inf=int(inf.sum()))
