"""Synthetic inputs for tests and bench: alignments evolved down a random
tree, site->pattern compression, and random unrooted binary trees in PLL "ring table" form.

Ring tables mirror the reference's node rings (pllrepo/src/pll.h:687-702): nodes are numbered
1..n (tips) and n+1..2n-2 (inner); an inner node has three ring slots 0,1,2 (slot s+1 is
`->next` of slot s), a tip has slot 0 only.  `back_node[3*i+s]`, `back_slot[3*i+s]` give the
node number and slot that slot s of node i is hooked to (`->back`); 0 = NULL.
"""
import numpy as np

# PLL data types (pllrepo/src/pll.h:238-245)
PLL_BINARY_DATA = 0
PLL_DNA_DATA = 1
PLL_AA_DATA = 2
PLL_GENERIC_32 = 6

STATES = {PLL_BINARY_DATA: 2, PLL_DNA_DATA: 4, PLL_AA_DATA: 20, PLL_GENERIC_32: 32}
ALPHABET = {
    PLL_BINARY_DATA: b"01",
    PLL_DNA_DATA: b"ACGT",
    PLL_AA_DATA: b"ARNDCQEGHILKMFPSTWYV",
    PLL_GENERIC_32: b"0123456789ABCDEFGHIJKLMNOPQRSTUV",
}
AMBIGUITY = {PLL_DNA_DATA: b"RYN", PLL_AA_DATA: b"BZX"}


def random_tree_rings(n, rng):
    """Random unrooted binary tree on tips 1..n by random stepwise edge insertion.
    Returns (back_node, back_slot), int32 arrays of length 3*(2n-1)."""
    assert n >= 4
    bn = np.zeros(3 * (2 * n - 1), dtype=np.int32)
    bs = np.zeros(3 * (2 * n - 1), dtype=np.int32)

    def hook(a, sa, b, sb):
        bn[3 * a + sa] = b; bs[3 * a + sa] = sb
        bn[3 * b + sb] = a; bs[3 * b + sb] = sa

    order = rng.permutation(n) + 1
    inner = n + 1
    # three-tip star around the first inner node
    for s in range(3):
        hook(inner, s, int(order[s]), 0)
    edges = [(inner, s) for s in range(3)]          # each edge listed once by one of its ends
    inner += 1
    for t in order[3:]:
        t = int(t)
        a, sa = edges[int(rng.integers(len(edges)))]
        b, sb = int(bn[3 * a + sa]), int(bs[3 * a + sa])
        perm = rng.permutation(3)                   # random ring orientation of the new node
        hook(inner, int(perm[0]), a, sa)
        hook(inner, int(perm[1]), b, sb)
        hook(inner, int(perm[2]), t, 0)
        edges.append((inner, int(perm[1])))
        edges.append((inner, int(perm[2])))
        inner += 1
    assert inner == 2 * n - 1
    return bn, bs


def evolve_alignment(n, nsites, datatype, mu, seed, gap=0.01, amb=0.001):
    """Characters (uint8 ASCII) [n][nsites]: root i.i.d. uniform, per-branch probability `mu`
    of switching to a different uniform state, down a random Yule-like tree; then `gap`
    fraction of '-' and `amb` fraction of ambiguity codes."""
    rng = np.random.default_rng(seed)
    S = STATES[datatype]
    alpha = np.frombuffer(ALPHABET[datatype], dtype=np.uint8)
    # random rooted topology by successive splitting of a random current leaf
    seqs = [rng.integers(0, S, size=nsites, dtype=np.uint8)]
    while len(seqs) < n:
        k = int(rng.integers(len(seqs)))
        parent = seqs.pop(k)
        for _ in range(2):
            child = parent.copy()
            hit = rng.random(nsites) < mu
            nh = int(hit.sum())
            if nh:
                child[hit] = (child[hit] + rng.integers(1, S, size=nh, dtype=np.uint8)) % S
            seqs.append(child)
    order = rng.permutation(n)
    chars = np.empty((n, nsites), dtype=np.uint8)
    for i, k in enumerate(order):
        chars[i] = alpha[seqs[int(k)]]
    if gap > 0:
        chars[rng.random((n, nsites)) < gap] = ord("-")
    if amb > 0 and datatype in AMBIGUITY:
        codes = np.frombuffer(AMBIGUITY[datatype], dtype=np.uint8)
        m = rng.random((n, nsites)) < amb
        chars[m] = codes[rng.integers(0, len(codes), size=int(m.sum()))]
    return chars


def evolve_alignment_blocked(n, nsites, datatype, mu, seed, gap=0.01, amb=0.001, block=1 << 17):
    """The same model as evolve_alignment for alignments too large for its n x nsites float temporaries
    (C4: 1000 x 1 000 000): the topology (sequence of leaf splits) is drawn first, then the sites are
    evolved down it block by block.  A different random stream than evolve_alignment -- only the bench's
    large workloads use it."""
    rng = np.random.default_rng(seed)
    S = STATES[datatype]
    alpha = np.frombuffer(ALPHABET[datatype], dtype=np.uint8)
    splits, live = [], 1
    while live < n:
        splits.append(int(rng.integers(live)))
        live += 1
    order = rng.permutation(n)
    codes = np.frombuffer(AMBIGUITY[datatype], dtype=np.uint8) if datatype in AMBIGUITY else None
    chars = np.empty((n, nsites), dtype=np.uint8)
    for s0 in range(0, nsites, block):
        m = min(block, nsites - s0)
        seqs = [rng.integers(0, S, size=m, dtype=np.uint8)]
        for k in splits:
            parent = seqs.pop(k)
            for _ in range(2):
                child = parent.copy()
                hit = rng.random(m, dtype=np.float32) < mu
                nh = int(hit.sum())
                if nh:
                    child[hit] = (child[hit] + rng.integers(1, S, size=nh, dtype=np.uint8)) % S
                seqs.append(child)
        blk = np.empty((n, m), dtype=np.uint8)
        for i, k in enumerate(order):
            blk[i] = alpha[seqs[int(k)]]
        if gap > 0:
            blk[rng.random((n, m), dtype=np.float32) < gap] = ord("-")
        if amb > 0 and codes is not None:
            msk = rng.random((n, m), dtype=np.float32) < amb
            blk[msk] = codes[rng.integers(0, len(codes), size=int(msk.sum()))]
        chars[:, s0:s0 + m] = blk
    return chars


def compress_patterns(chars):
    """Unique columns (in order of first appearance) and their frequencies:
    the Alignment::addPattern step of the host (alignment.cpp), restated with numpy."""
    n, L = chars.shape
    cols = np.ascontiguousarray(chars.T)
    view = cols.view([("", cols.dtype)] * n).ravel()
    _, first, counts = np.unique(view, return_index=True, return_counts=True)
    order = np.argsort(first, kind="stable")
    first = first[order]
    return np.ascontiguousarray(chars[:, first]), counts[order].astype(np.int32)


def bootstrap_weights(weights, B, seed):
    """B multinomial resamplings of the site->pattern frequencies (Alignment::createBootstrapAlignment,
    alignment.cpp:1971-2040, restated with numpy's generator): uint16 [B][P]."""
    rng = np.random.default_rng(seed)
    L = int(weights.sum())
    p = weights.astype(np.float64) / L
    return rng.multinomial(L, p, size=B).astype(np.uint16)
