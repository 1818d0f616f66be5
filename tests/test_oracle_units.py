"""Pins the CPU oracle against every unit vector in the reference's own tests.

Each test names the reference test it restates (/root/reference/src/...).
The golden strings of rice_coding.rs / phase_in_coding.rs went through the
reference's BitWriterMock, whose write() is LSB-first per field
(bitwrite_mock.rs:30-41); the oracle's `mock=True` writer reproduces that sink,
`mock=False` is the real MSB-first BitWriter<_, BigEndian>.
"""
import random

import numpy as np
import pytest

from oracle import felics_oracle as fo


# ---- coding/rice_coding.rs ---------------------------------------------------
def test_rice_encoding_mock_strings():  # rice_coding.rs:70-82
    assert fo.rice_bits(4, 7, mock=True) == "01110"
    assert fo.rice_bits(0, 12, mock=True) == "1111111111110"
    assert fo.rice_bits(3, 10, mock=True) == "10010"


def test_rice_encoding_real_bit_order():
    # same codes through the MSB-first writer: unary part unchanged, remainder reversed
    assert fo.rice_bits(4, 7) == "0" + "0111"
    assert fo.rice_bits(0, 12) == "1" * 12 + "0"
    assert fo.rice_bits(3, 10) == "10" + "010"


def test_rice_panic_k32():  # rice_coding.rs:84-88
    with pytest.raises(fo.OracleError):
        fo.rice_bits(32, 1)


def test_rice_decoding():  # rice_coding.rs:90-107
    dec, data = fo.codes_roundtrip([0, 0, 0], [4, 0, 3], [7, 12, 10])
    assert list(dec) == [7, 12, 10]
    # 0 0111 | 1111111111110 | 10 010 -> 23 bits, padded with one zero
    assert data == int("00111" + "1" * 12 + "0" + "10010" + "0", 2).to_bytes(3, "big")


def test_rice_decoding_extensive():  # rice_coding.rs:109-135 (#[ignore] in the reference)
    numbers = list(range(65535 * 2))
    random.Random(1).shuffle(numbers)
    dec, _ = fo.codes_roundtrip([0] * len(numbers), [8] * len(numbers), numbers)
    assert list(dec) == numbers


def test_rice_code_length():  # rice_coding.rs:137-148
    for number in range(0, 3000, 7):
        for k in range(32):
            assert fo.rice_code_length(k, number) == len(fo.rice_bits(k, number)) == (number >> k) + 1 + k


# ---- coding/phase_in_coding.rs -------------------------------------------------
def test_phase_in_panics():  # phase_in_coding.rs:123-133
    with pytest.raises(fo.OracleError):
        fo.phase_in_params(0)
    with pytest.raises(fo.OracleError):
        fo.phase_in_params(1 << 31)


def test_new_coder():  # phase_in_coding.rs:136-161: (m, left_p, right_p)
    assert fo.phase_in_params(1) == (0, 0, 1)
    assert fo.phase_in_params(7) == (2, 3, 1)
    assert fo.phase_in_params(15) == (3, 7, 1)
    assert fo.phase_in_params(32) == (5, 0, 32)


def test_value_outside_range():  # phase_in_coding.rs:163-170
    with pytest.raises(fo.OracleError):
        fo.phase_in_bits(15, 15)


PHASE_IN_MOCK_TABLES = {  # phase_in_coding.rs:186-224
    7: ["011", "110", "111", "00", "100", "101", "010"],
    8: ["000", "100", "010", "110", "001", "101", "011", "111"],
    9: ["1111", "000", "100", "010", "110", "001", "101", "011", "1110"],
    15: ["0011", "1010", "1011", "0110", "0111", "1110", "1111", "000", "1000", "1001", "0100", "0101", "1100", "1101", "0010"],
    16: ["0000", "1000", "0100", "1100", "0010", "1010", "0110", "1110", "0001", "1001", "0101", "1101", "0011", "1011", "0111", "1111"],
    17: ["11111", "0000", "1000", "0100", "1100", "0010", "1010", "0110", "1110", "0001", "1001", "0101", "1101", "0011", "1011", "0111", "11110"],
}


@pytest.mark.parametrize("n", sorted(PHASE_IN_MOCK_TABLES))
def test_phase_in_encoding(n):
    assert [fo.phase_in_bits(n, v, mock=True) for v in range(n)] == PHASE_IN_MOCK_TABLES[n]
    # real writer: the m-bit field is MSB-first, i.e. the mock string with its first m bits reversed
    m, _, _ = fo.phase_in_params(n)
    for v in range(n):
        mock = PHASE_IN_MOCK_TABLES[n][v]
        assert fo.phase_in_bits(n, v) == mock[:m][::-1] + mock[m:]


def test_phase_in_decoding_extensive():  # phase_in_coding.rs:229-252 (#[ignore] in the reference)
    rnd = random.Random(2)
    for n in list(range(1, 600)) + [1023, 1024, 1025, 1999]:
        domain = list(range(n))
        rnd.shuffle(domain)
        dec, _ = fo.codes_roundtrip([1] * n, [n] * n, domain)
        assert list(dec) == domain


def test_phase_in_lengths_are_prefix_free_and_short():
    for n in range(1, 520):
        codes = [fo.phase_in_bits(n, v) for v in range(n)]
        m = n.bit_length() - 1
        assert all(len(c) in (m, m + 1) for c in codes)
        assert len(set(codes)) == n or n == 1
        for a in codes:
            for b in codes:
                assert a == b or not b.startswith(a) or a == ""


# ---- compression/parameter_selection.rs -----------------------------------------
def test_estimator_context_map():  # parameter_selection.rs:95-124
    k_values = [0, 1, 2, 4, 8, 16]
    est = fo.KEstimator(300, k_values, None)
    add = {100: [4, 8, 13, 45, 85], 80: [7, 800, 1000, 1273, 85], 75: [7, 13, 1000, 200, 85],
           255: [1, 4, 142, 563, 1246, 2464], 0: [0, 100, 3]}
    for ctx, vals in add.items():
        for v in vals:
            est.update(ctx, v)
    for ctx, vals in add.items():
        for i, k in enumerate(k_values):
            assert est.entry(ctx, i) == sum(fo.rice_code_length(k, v) for v in vals)


def test_estimator_get_k():  # parameter_selection.rs:126-146
    est = fo.KEstimator(400, [0, 1, 2, 4, 5, 16], None)
    for v in (10, 40, 5):
        est.update(100, v)
    assert est.get_k(100) == 4
    for v in (1000, 200, 1250, 300):
        est.update(255, v)
    assert est.get_k(255) == 16


def test_estimator_no_k_values():  # parameter_selection.rs:148-152
    with pytest.raises(fo.OracleError):
        fo.KEstimator(100, [], None)


def test_estimator_periodic_count_scaling():  # parameter_selection.rs:154-183
    est = fo.KEstimator(120, [0, 1, 2], 1024)
    est.update(43, 400)
    assert [est.entry(43, i) for i in range(3)] == [401, 202, 103]
    est.update(43, 531)
    assert [est.entry(43, i) for i in range(3)] == [933, 469, 238]
    est.update(43, 2000)
    assert [est.entry(43, i) for i in range(3)] == [2934, 1471, 741]  # min 741 <= 1024: no halving
    est.update(43, 1733)
    assert [est.entry(43, i) for i in range(3)] == [2334, 1169, 588]  # 4668/2339/1177 halved


def test_estimator_untouched_context_picks_largest_k():  # DOC.md:354, :430 (ties go to the last index)
    assert fo.KEstimator(510, [0, 1, 2, 3, 4, 5], 1024).get_k(17) == 5
    assert fo.KEstimator(131070, list(range(15)), 1024).get_k(4242) == 14


def test_estimator_context_out_of_range_panics():  # parameter_selection.rs:50, :72
    est = fo.KEstimator(10, [0, 1], None)
    with pytest.raises(fo.OracleError):
        est.update(11, 1)
    with pytest.raises(fo.OracleError):
        est.get_k(11)


# ---- compression/misc.rs -------------------------------------------------------
def test_nearest_neighbours():  # misc.rs:33-69
    w = 23

    def pti(x, y, width=w):
        return y * width + x

    assert fo.nearest_neighbours(pti(5, 8), w) == (pti(4, 8), pti(5, 7))
    assert fo.nearest_neighbours(pti(0, 8), w) == (pti(0, 7), pti(0, 6))
    assert fo.nearest_neighbours(pti(2, 0), w) == (pti(1, 0), pti(0, 0))
    assert fo.nearest_neighbours(pti(1, 1), w) == (pti(0, 1), pti(1, 0))
    assert fo.nearest_neighbours(pti(1, 0), w) is None
    assert fo.nearest_neighbours(pti(0, 1), w) == (pti(0, 0), pti(1, 0))
    assert fo.nearest_neighbours(pti(0, 0, 1), 1) is None
    assert fo.nearest_neighbours(pti(0, 1, 1), 1) is None
    assert fo.nearest_neighbours(pti(0, 2, 1), 1) == (pti(0, 1, 1), pti(0, 0, 1))


# ---- compression/color_transform.rs ----------------------------------------------
def test_color_transform8():  # color_transform.rs:35-73 (exhaustive 2^24)
    rc, (ymin, ymax, comin, comax, cgmin, cgmax) = fo.color_transform8_exhaustive()
    assert rc == 0
    assert ymax - ymin <= 510 and comax - comin <= 510 and cgmax - cgmin <= 510
    assert (ymin, ymax, comin, comax, cgmin, cgmax) == (0, 255, -255, 255, -255, 255)


def test_color_transform16():  # color_transform.rs:75-120
    values = [(0, 65535, 0), (0, 0, 65535), (65535, 0, 0), (65535, 65535, 65535), (65535, 0, 65535),
              (1726, 12640, 26649), (0, 0, 0), (9127, 65535, 3)]
    ys, cos, cgs = [], [], []
    for r, g, b in values:
        y, co, cg = fo.rgb_to_ycocg(r, g, b)
        assert fo.ycocg_to_rgb(y, co, cg) == (r, g, b)
        ys.append(y), cos.append(co), cgs.append(cg)
    for seq in (ys, cos, cgs):
        assert max(seq) - min(seq) <= 131070


def test_color_transform_doc_example():  # DOC.md:465
    assert fo.rgb_to_ycocg(231, 27, 30) == (79, 201, -103)


def test_color_transform_division_truncates_toward_zero():  # color_transform.rs:13,15 use i32 `/`
    # co = -1 -> co/2 = 0 (an arithmetic shift would give -1)
    assert fo.rgb_to_ycocg(0, 0, 1) == (1, -1, -1)
    y, co, cg = fo.rgb_to_ycocg(10, 3, 13)
    assert (co, cg) == (-3, 3 - (13 + (-3 // 2 + 1)))
