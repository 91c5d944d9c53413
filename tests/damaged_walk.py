"""The damaged-scan walk shared by tests/test_hostemu.py (kernel code on the host), tests/test_gpu_parity.py (through the
kernels) and tools/parity_hunt.py: 9 parameter sets x 3 image kinds x 60 damaged copies of a valid scan (bit flips,
truncation, spliced bytes) = 1620 streams, each decoded by the unmodified reference (the arbiter) and by us.

The rule is asymmetric.  Whatever the reference accepts we must accept with identical samples: 0 exceptions.  What the
reference rejects we reject too, except for a known set: the reference looks for the restart marker (or the end of the scan)
where its 64-bit read cache happened to stop (src/scan_decoder.hpp:237-243, 335-349), so a few stray bytes in front of a marker
pass or fail depending on its refill schedule, which we do not model.  That count is pinned exactly."""
import numpy as np

CASES = ((8, 1, 0, 0, 0), (8, 1, 0, 0, 1), (12, 1, 0, 2, 3), (16, 3, 2, 0, 0), (8, 3, 1, 0, 2), (16, 1, 0, 0, 1), (5, 4, 2, 1, 0),
         (8, 3, 2, 0, 1), (2, 1, 0, 0, 0))
EXPECTED_REFERENCE_ACCEPTS = 176
EXPECTED_ACCEPTED_THOUGH_REFERENCE_REJECTS = 30


def damaged_streams(oracle):
    """Yields (key, whole stream, scan data, image, scan parameters) for every damaged copy, deterministically."""
    from tests import jlsio
    from tests.support import s_mixed, s_noise, s_smooth

    rng = np.random.default_rng(3)
    for bits, cc, ilv, near, ri in CASES:
        for gen in (s_mixed, s_smooth, s_noise):
            img = gen(7, 90, bits, cc, seed=bits + cc, layout="interleaved") if cc > 1 else gen(7, 90, bits, seed=bits)
            sp = oracle.params(90, 7, bits, cc, near, ilv, 0, None, ri)
            good = oracle.encode_scan(sp, img)
            whole = oracle.encode_image(img, bits, near=near, ilv=ilv, ri=ri)
            sc = jlsio.parse(whole).scans[0]
            assert whole[sc.data_offset : sc.data_end] == good
            for trial in range(60):
                data = bytearray(good)
                kind = trial % 3
                if kind == 0:
                    for _ in range(1 + trial % 4):
                        i = int(rng.integers(0, len(data)))
                        data[i] ^= 1 << int(rng.integers(0, 8))
                elif kind == 1:
                    data = data[: int(rng.integers(1, len(data)))]
                else:
                    i = int(rng.integers(0, len(data)))
                    data[i : i + int(rng.integers(1, 6))] = bytes(rng.integers(0, 256, size=int(rng.integers(0, 5)), dtype=np.uint8))
                data = bytes(data)
                yield (bits, cc, ilv, near, ri, gen.__name__, trial), whole[: sc.data_offset] + data + b"\xff\xd9", data, img, sp


def run_walk(oracle, reference, decode_ours):
    """decode_ours(stream, data, img, sp) -> samples (any shape) or None when rejected.  Returns the two counts."""
    from charls_b200 import codec
    from charls_b200.capi import CharlsError

    reference_accepts = accepted_though_rejected = 0
    for key, stream, data, img, sp in damaged_streams(oracle):
        try:
            want, _, _ = codec.decode(stream, lib=reference)
        except CharlsError:
            want = None
        got = decode_ours(stream, data, img, sp)
        if want is not None:
            reference_accepts += 1
            assert got is not None, ("the reference accepts this stream, we reject it", key)
            assert np.array_equal(np.asarray(got).reshape(want.shape), want), ("accepted with different samples", key)
        elif got is not None:
            accepted_though_rejected += 1
    return reference_accepts, accepted_though_rejected
