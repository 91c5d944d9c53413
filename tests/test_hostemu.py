"""Runs the product's kernel code (charls_b200/csrc/jls_interval.cuh compiled for the host, one interval after the
other) against the oracle.  This is what keeps the device code honest in the GPU-less container; tests/test_gpu_*.py
repeat the comparison through the real kernels."""
import numpy as np
import pytest

from tests.hostemu_lib import HostEmu
from tests.support import s_mixed, s_noise, s_smooth


@pytest.fixture(scope="module")
def hostemu():
    return HostEmu()


def check_scan(oracle, he, plane, bits, nc, near, ilv, xf, ri, pc=None):
    h, w = plane.shape[0], plane.shape[1]
    sp = oracle.params(w, h, bits, nc, near, ilv, xf, pc, ri)
    want = oracle.encode_scan(sp, plane)
    hp = he.params(sp)
    expected = np.zeros_like(plane)
    assert oracle.decode_scan(sp, want + b"\xff\xd9", expected) == len(want)
    for force_general in ((False, True) if he.fast(hp) else (True,)):
        n, got = he.encode(hp, plane, len(want) + 64, force_general)
        assert n == len(want) and got == want, ("encode", plane.shape, bits, nc, near, ilv, xf, ri, pc, force_general)
        out = np.zeros_like(plane)
        n = he.decode(hp, want + b"\xff\xd9", out, force_general)
        assert n == len(want) and np.array_equal(out, expected), ("decode", plane.shape, bits, nc, near, ilv, xf, ri, pc, force_general)


@pytest.mark.parametrize("bits", [2, 4, 8, 10, 12, 16])
def test_scalar_lines(oracle, hostemu, bits):
    for gen in (s_smooth, s_noise, s_mixed):
        for (h, w) in ((9, 33), (1, 9), (13, 1), (12, 130)):
            img = gen(h, w, bits)
            for near in (0, 1, 3):
                if near > ((1 << bits) - 1) // 2:
                    continue
                for ri in (1, 0, 3):
                    check_scan(oracle, hostemu, img, bits, 1, near, 0, 0, ri)


@pytest.mark.parametrize("bits", [5, 8, 16])
def test_multi_component(oracle, hostemu, bits):
    for cc in (2, 3, 4):
        for ilv in (1, 2):
            for gen in (s_smooth, s_mixed):
                img = gen(11, 37, bits, cc, layout="interleaved")
                for near in (0, 2):
                    for ri in (1, 0, 4):
                        check_scan(oracle, hostemu, img, bits, cc, near, ilv, 0, ri)
                if cc == 3 and bits in (8, 16):
                    for xf in (1, 2, 3):
                        for ri in (1, 0):
                            check_scan(oracle, hostemu, img, bits, cc, 0, ilv, xf, ri)


def test_presets_masking_and_long_runs(oracle, hostemu):
    for pc in ((255, 9, 9, 9, 31), (0, 0, 0, 0, 3), (0, 5, 6, 200, 255)):
        img = s_mixed(17, 65, 8)
        for ri in (1, 0):
            check_scan(oracle, hostemu, img, 8, 1, 0, 0, 0, ri, pc)
            check_scan(oracle, hostemu, img, 8, 1, 2, 0, 0, ri, pc)
    # unused high bits are masked away (reference copy_to_line_buffer.hpp:106-116)
    check_scan(oracle, hostemu, s_noise(8, 40, 16), 12, 1, 0, 0, 0, 1)
    check_scan(oracle, hostemu, s_noise(8, 40, 16, 3, layout="interleaved"), 12, 3, 0, 2, 0, 1)
    # very long zero runs exercise run_index 31 and lines wider than 65535
    z = np.zeros((3, 70000), np.uint8)
    z[1, 69999] = 7
    z[2, 100] = 9
    check_scan(oracle, hostemu, z, 8, 1, 0, 0, 0, 1)
    check_scan(oracle, hostemu, z, 8, 1, 0, 0, 0, 0)


def test_stuffed_bytes_on_wide_noise_lines(oracle, hostemu):
    """Long incompressible lines: hundreds of 0xFF bytes per line, so both sides of the bit writer's and reader's
    word paths (plain / stuffed, jls_fast.cuh FastWriter::flush_word, FastReader::refill_once) run back to back."""
    for bits, cc, ilv in ((8, 1, 0), (12, 1, 0), (16, 1, 0), (16, 3, 2), (8, 3, 2)):
        img = s_noise(3, 4099, bits, cc, seed=bits + cc, layout="interleaved")
        sp = oracle.params(4099, 3, bits, cc, 0, ilv, 0, None, 1)
        assert oracle.encode_scan(sp, img).count(b"\xff") > 20
        for near in (0, 1):
            check_scan(oracle, hostemu, img, bits, cc, near, ilv, 0, 1)


def test_golden_vectors(oracle, hostemu, golden):
    """Kernel code == reference-made streams (restart interval 1 stitched from per-row reference encodings)."""
    from tests import jlsio

    for v in golden:
        if v.ilv == 0 and v.image.ndim == 3:
            continue  # multi-scan: covered per plane by the other tests
        s = jlsio.parse(v.ri1)
        sc = s.scans[0]
        h, w = (v.image.shape[0], v.image.shape[1])
        sp = oracle.params(w, h, v.bits, sc.component_count, v.near, v.ilv, v.xf, v.pc, 1)
        hp = hostemu.params(sp)
        want = v.ri1[sc.data_offset : sc.data_end]
        n, got = hostemu.encode(hp, v.image, len(want) + 64)
        assert got == want, v.name
        out = np.zeros_like(v.image)
        assert hostemu.decode(hp, want + b"\xff\xd9", out) == len(want), v.name
        assert np.array_equal(out, v.dec1), v.name


@pytest.mark.parametrize("seed", range(8))
def test_random_parameters(oracle, hostemu, seed):
    """Seeded random walk over the parameter space (bit depth, NEAR up to its limit, presets, component counts,
    interleave modes, colour transforms, restart intervals, odd sizes): kernel code == oracle, byte for byte."""
    import random

    rng = random.Random(4000 + seed)
    for _ in range(40):
        bits = rng.choice([2, 3, 5, 7, 8, 8, 9, 10, 12, 12, 15, 16, 16])
        maxval = (1 << bits) - 1
        cc = rng.choice([1, 1, 1, 2, 3, 3, 4])
        ilv = 0 if cc == 1 else rng.choice([1, 2])
        near = rng.choice([0, 0, 0, 1, 2, 3, min(255, maxval // 2)])
        near = min(near, maxval // 2, 255)
        xf = rng.choice([0, 1, 2, 3]) if (cc == 3 and near == 0 and bits in (8, 16)) else 0
        ri = rng.choice([1, 1, 1, 0, 2, 5])
        w, h = rng.choice([1, 2, 3, 17, 64, 65, 127, 200, 301]), rng.choice([1, 2, 5, 9])
        pc = None
        if rng.random() < 0.35:
            t1 = rng.randint(near + 1, maxval)
            t2 = rng.randint(t1, maxval)
            t3 = rng.randint(t2, maxval)
            pc = (0, t1, t2, t3, rng.choice([3, 4, 31, 64, 255]))
        gen = rng.choice([s_smooth, s_noise, s_mixed])
        img = gen(h, w, bits, cc, seed=rng.randrange(1 << 30), layout="interleaved") if cc > 1 else gen(h, w, bits, seed=rng.randrange(1 << 30))
        check_scan(oracle, hostemu, img, bits, cc, near, ilv, xf, ri, pc)


def test_golomb_parameter_table_form_is_exact(hostemu):
    """The tile kernels compute k from ceil(2^31 / N) (jls_fast.cuh: golomb_parameter_reciprocal).  Exhaustive against
    the definition -- min k with (N << k) >= A, reference src/regular_mode_context.hpp:99-111 -- for every N the kernels
    use it for (N <= RESET <= 64) and every A below the reference's sanity bound 2^24 (:52-54), A = 0 included."""
    import ctypes

    check = hostemu.dll.hostemu_check_golomb_parameter
    check.restype = ctypes.c_uint64
    check.argtypes = [ctypes.c_int32, ctypes.c_int32]
    assert check(64, 1 << 24) == 0


def test_damaged_lines_multi_component(oracle, hostemu):
    """Bit flips and truncation in lines of three-component pixels (the decoder that takes two stream words per top-up,
    FastReaderT<2>::refill_pair): never a crash; whatever the oracle accepts decodes to the same samples.  Which damaged
    lines get *rejected* may differ (profiles / DESIGN.md section 8), so that direction is only counted."""
    rng = np.random.default_rng(11)
    checked = accepted = 0
    for bits, cc in ((16, 3), (8, 3), (12, 2), (16, 4)):
        img = s_mixed(6, 150, bits, cc, seed=bits, layout="interleaved")
        sp = oracle.params(150, 6, bits, cc, 0, 2, 0, None, 1)
        good = oracle.encode_scan(sp, img)
        hp = hostemu.params(sp)
        for trial in range(40):
            data = bytearray(good)
            if trial % 2 == 0:
                for _ in range(1 + trial % 3):
                    i = int(rng.integers(0, len(data)))
                    data[i] ^= 1 << int(rng.integers(0, 8))
                    if data[i] == 0xFF or (i and data[i - 1] == 0xFF and data[i] >= 0x80):
                        data[i] = 0x55  # keep the marker structure: this test is about the entropy-coded bits
            else:
                cut = int(rng.integers(1, len(data)))
                # truncation inside a line: keep every restart marker, shorten one interval
                marker = bytes(data).find(b"\xff", cut)
                if marker > cut:
                    del data[cut:marker]
            data = bytes(data) + b"\xff\xd9"
            want = np.zeros_like(img)
            n_oracle = oracle.decode_scan(sp, data, want)
            got = np.zeros_like(img)
            n_ours = hostemu.decode(hp, data, got, False)
            checked += 1
            if n_oracle >= 0:
                accepted += 1
                assert n_ours == n_oracle and np.array_equal(got, want), (bits, cc, trial)
    assert checked == 160 and accepted > 0


def test_damaged_walk_against_the_reference(oracle, hostemu, reference):
    """1620 damaged scans through the kernels' codec code on the host (tests/damaged_walk.py): everything the reference
    accepts is accepted with identical samples; the set it rejects and we accept has exactly the known size."""
    from tests import damaged_walk

    params = {}

    def ours(stream, data, img, sp):
        hp = params.setdefault(id(sp), hostemu.params(sp))
        out = np.zeros_like(img)
        return out if hostemu.decode(hp, data + b"\xff\xd9", out, False) >= 0 else None

    accepts, extra = damaged_walk.run_walk(oracle, reference, ours)
    assert accepts == damaged_walk.EXPECTED_REFERENCE_ACCEPTS
    assert extra == damaged_walk.EXPECTED_ACCEPTED_THOUGH_REFERENCE_REJECTS
