"""Known-answer vectors for Philox4x32-10 (Random123 kat_vectors, Salmon et al. SC'11)."""
import ctypes as C

import numpy as np

KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox_known_answers(oracle_lib):
    for ctr, key, want in KAT:
        c = np.array(ctr, np.uint32)
        k = np.array(key, np.uint32)
        o = np.zeros(4, np.uint32)
        oracle_lib.gxo_philox(c.ctypes.data_as(C.c_void_p), k.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p))
        assert tuple(int(v) for v in o) == want
