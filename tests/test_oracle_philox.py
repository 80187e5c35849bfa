"""Pins oracle/philox.hpp against the Random123 known-answer vectors (kat_vectors, philox4x32 10)."""
import numpy as np

import oracle_binding as ob

KAT = [
    ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
    ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
    ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
     [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
]


def test_philox_known_answers():
    for ctr, key, want in KAT:
        got = ob.philox(ctr, key)
        assert [int(v) for v in got] == want


def test_u01_matches_rand09_standard_uniform():
    # rand 0.9 StandardUniform<f32>: 24 high bits scaled by 2^-24, always in [0,1)
    L = ob.lib()
    assert L.okg_u01_f32(0) == 0.0
    assert L.okg_u01_f32(0xFFFFFFFF) == np.float32(1.0) - np.float32(2.0 ** -24)
    assert L.okg_u01_f32(0x80000000) == 0.5
    assert L.okg_u01_f32(0x000000FF) == 0.0  # low 8 bits are discarded
