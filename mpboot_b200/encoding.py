"""Character -> PLL code tables of the host (pllBaseSubstitute, pllrepo/src/utils.c:98-157 and
:2526), restated; 255 marks characters the reference does not map.  The product library takes
codes, exactly like tr->yVector; this module is for hosts that start from characters (tests,
bench).  Pinned against the reference tables by tests/test_oracle_cpu.py::test_char_maps."""
import numpy as np

from .synth import PLL_AA_DATA, PLL_BINARY_DATA, PLL_DNA_DATA, PLL_GENERIC_32


def _table(pairs, both_cases=True):
    t = np.full(256, 255, dtype=np.uint8)
    for ch, code in pairs:
        t[ord(ch)] = code
        if both_cases and ch.isalpha():
            t[ord(ch.lower())] = code
    return t


_DNA = _table([("A", 1), ("C", 2), ("G", 4), ("T", 8), ("U", 8), ("M", 3), ("R", 5), ("S", 6), ("V", 7),
               ("W", 9), ("Y", 10), ("H", 11), ("K", 12), ("D", 13), ("B", 14),
               ("N", 15), ("O", 15), ("X", 15), ("-", 15), ("?", 15)])
_AA = _table([(c, i) for i, c in enumerate("ARNDCQEGHILKMFPSTWYV")] +
             [("B", 20), ("Z", 21), ("X", 22), ("-", 22), ("?", 22), ("*", 22)])
_BIN = _table([("0", 1), ("1", 2), ("-", 3), ("?", 3)])
# '?' maps to state 22 ('M'), not to "missing" -- reference quirk, SURVEY 8a item 8
_G32 = _table([(c, i) for i, c in enumerate("0123456789ABCDEFGHIJKLMNOPQRSTUV")] +
              [("-", 32), ("*", 32), ("?", 22)])

CHAR_MAP = {PLL_DNA_DATA: _DNA, PLL_AA_DATA: _AA, PLL_BINARY_DATA: _BIN, PLL_GENERIC_32: _G32}


def encode(chars, datatype):
    """uint8 ASCII matrix -> PLL codes (tr->yVector contents)."""
    codes = CHAR_MAP[datatype][np.asarray(chars, dtype=np.uint8)]
    if (codes == 255).any():
        raise ValueError("alignment contains characters the reference does not map for this data type")
    return codes
