"""Variable-inclusion wire format (pymc_bart/utils.py:1368-1398): the one exact fixture the
reference holds for this path (tests/test_utils.py:101-113) plus hand-derived known answers."""
import base64

import numpy as np

from pymc_bart_b200.utils import _decode_vi, _encode_vi


def test_reference_round_trip_cases():
    cases = [np.zeros(3, dtype=int), np.ones(10, dtype=int), np.array([4, 0, 1, 0, 2, 0, 3, 0, 0, 0]),
             np.array([100, 50, 0, 1]), np.array([1, 2, 4, 8, 16])]          # tests/test_utils.py:103-109
    for case in cases:
        assert np.array_equal(_decode_vi(_encode_vi(case), len(case)), case)


def test_known_answers():
    # LEB128: values <= 127 are one byte; 300 = 0b1_0010_1100 -> 0xAC 0x02; 16384 -> 0x80 0x80 0x01
    assert _encode_vi([100, 50, 0, 1]) == base64.b64encode(bytes([100, 50, 0, 1])).decode() == "ZDIAAQ=="
    assert _encode_vi([300]) == base64.b64encode(bytes([0xAC, 0x02])).decode() == "rAI="
    assert _encode_vi([127, 128, 16384]) == base64.b64encode(bytes([0x7F, 0x80, 0x01, 0x80, 0x80, 0x01])).decode()
    assert _decode_vi("rAI=", 1) == [300]
    assert _decode_vi(_encode_vi([]), 0) == []
    assert _decode_vi(_encode_vi([2**31 - 1, 0, 2**20]), 3) == [2**31 - 1, 0, 2**20]


def test_decode_stops_at_length():
    assert _decode_vi(_encode_vi([1, 2, 3, 4]), 2) == [1, 2]
