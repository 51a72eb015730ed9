"""TEST INFRASTRUCTURE -- not product code.  Only tests/ may import this.

Plain-Python restatement of the split bookkeeping behind `-bb` supports in the reference:
  MTree::convertSplits          mtree.cpp:917-939    every edge's split, pushed after the splits of the subtree below it
  Split::shouldInvert / invert  split.cpp:90-107     keep the side with fewer taxa; at a tie the side containing taxon 0
  MTreeSet::convertSplits       mtreeset.cpp:362-440 trees in order, weight-0 trees skipped, weights summed per distinct
                                                     split (SW_COUNT), new splits appended in first-seen order
Pinned to the reference itself through tests/test_gpu_dropin.py: the patched MPBoot builds its split graph from
mpgpu_split_table and must write byte-identical .splits.nex / .contree.
"""
import random


def random_tree(taxa, rng):
    """Random rooted-at-a-leaf binary tree as nested lists; returns (root_taxon, subtree) where subtree hangs off the root leaf."""
    items = list(taxa)
    rng.shuffle(items)
    root = items.pop()
    nodes = items
    while len(nodes) > 1:
        i = rng.randrange(len(nodes)); a = nodes.pop(i)
        j = rng.randrange(len(nodes)); b = nodes.pop(j)
        nodes.append([a, b])
    return root, nodes[0]


def tokens_of(subtree):
    """Reverse-Polish token stream of MTree::convertSplits' traversal below the root leaf: leaf t -> t, inner node -> -k."""
    out = []

    def rec(x):
        if isinstance(x, list):
            for c in x:
                rec(c)
            out.append(-len(x))
        else:
            out.append(int(x))
    rec(subtree)
    return out


def split_table(ntaxa, trees, weights):
    """trees: token lists.  Returns (splits as frozensets in first-seen order, weights, emit rows per token, concatenated)."""
    order, index, wsum, emit = [], {}, [], []
    for toks, wt in zip(trees, weights):
        stack = []
        for t in toks:
            if t >= 0:
                s = frozenset([t])
            else:
                k = -t
                s = frozenset().union(*stack[len(stack) - k:])
                del stack[len(stack) - k:]
            stack.append(s)
            c = len(s)
            inv = 2 * c > ntaxa or (2 * c == ntaxa and 0 not in s)
            key = frozenset(range(ntaxa)) - s if inv else s
            if key not in index:
                index[key] = len(order); order.append(key); wsum.append(0)
            wsum[index[key]] += wt
            emit.append(index[key])
    return order, wsum, emit


def bits_of(split, ntaxa):
    W = (ntaxa + 31) // 32
    words = [0] * W
    for t in split:
        words[t >> 5] |= 1 << (t & 31)
    return words
