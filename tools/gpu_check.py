"""Quick GPU parity sweep through the C ABI (development aid; the real tests live in tests/)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.support import oracle, reference_library, have_reference_build, s_smooth, s_noise, s_mixed
from tests import jlsio
from charls_b200 import codec, capi

lib = capi.default_library()
o = oracle()
ref = reference_library() if have_reference_build() else None
bad = 0
n = 0


def check(img, bits, near=0, ilv=0, xf=0, ri=1, pc=None, tag=""):
    global bad, n
    n += 1
    try:
        got = codec.encode(img, bits, near_lossless=near, interleave_mode=ilv, color_transformation=xf, restart_interval=ri, preset=pc, lib=lib)
    except Exception as e:
        bad += 1
        print("ENC EXC", tag, img.shape, bits, near, ilv, xf, ri, e)
        return
    want = o.encode_image(img, bits, near=near, ilv=ilv, xform=xf, pc=pc, ri=ri)
    sa, sb = jlsio.parse(got), jlsio.parse(want)
    ok = len(sa.scans) == len(sb.scans) and all(got[x.data_offset:x.data_end] == want[y.data_offset:y.data_end] for x, y in zip(sa.scans, sb.scans))
    if not ok:
        bad += 1
        print("ENC MISMATCH", tag, img.shape, bits, near, ilv, xf, ri, len(got), len(want))
        return
    exp, _ = o.decode_image(got)
    try:
        px, fi, _ = codec.decode(got, lib=lib)
    except Exception as e:
        bad += 1
        print("DEC EXC", tag, img.shape, bits, near, ilv, xf, ri, e)
        return
    if not np.array_equal(px, exp):
        bad += 1
        print("DEC MISMATCH", tag, img.shape, bits, near, ilv, xf, ri)
        return
    if ref is not None:
        pr, _, _ = codec.decode(got, lib=ref)
        if not np.array_equal(pr, exp):
            bad += 1
            print("REF DEC MISMATCH", tag)


t0 = time.time()
for bits in (8, 12, 16, 2, 5):
    for gen in (s_smooth, s_noise, s_mixed):
        for (h, w) in ((9, 33), (1, 9), (13, 1), (70, 300)):
            img = gen(h, w, bits)
            for near in (0, 2):
                if near > ((1 << bits) - 1) // 2:
                    continue
                for ri in (1, 0, 3):
                    check(img, bits, near, 0, 0, ri, tag=gen.__name__)
print("mono", n, bad, time.time() - t0)
for bits in (8, 16):
    for cc in (2, 3, 4):
        for ilv in (0, 1, 2):
            img = s_mixed(11, 37, bits, cc, layout="planar" if ilv == 0 else "interleaved")
            for near in (0, 2):
                for ri in (1, 0):
                    check(img, bits, near, ilv, 0, ri, tag="color")
            if cc == 3 and ilv != 0:
                for xf in (1, 2, 3):
                    check(img, bits, 0, ilv, xf, 1, tag="xf")
print("color", n, bad, time.time() - t0)
check(s_smooth(64, 64, 8), 8, 0, 0, 0, 1, (255, 9, 9, 9, 31), tag="pc")
# BASELINE configs at full size
for name, img, bits, near, ilv, xf in (
    ("cfg2", s_smooth(4096, 4096, 8), 8, 0, 0, 0),
    ("cfg3", s_smooth(4096, 4096, 12), 12, 2, 0, 0),
    ("cfg4", s_smooth(2048, 2048, 16, 3, layout="interleaved"), 16, 0, 2, 1),
):
    t = time.time()
    s = codec.encode(img, bits, near_lossless=near, interleave_mode=ilv, color_transformation=xf, lib=lib)
    t1 = time.time()
    px, fi, _ = codec.decode(s, lib=lib)
    t2 = time.time()
    okd = True
    if ref is not None:
        pr, _, _ = codec.decode(s, lib=ref)
        okd = np.array_equal(pr, px)
    lossless_ok = near != 0 or np.array_equal(px, img)
    near_ok = near == 0 or int(np.abs(px.astype(np.int64) - img.astype(np.int64)).max()) <= near
    print(name, "bytes", len(s), "ratio %.3f" % (img.nbytes / len(s)), "enc %.1f ms dec %.1f ms" % ((t1 - t) * 1e3, (t2 - t1) * 1e3),
          "ref-decode-equal", okd, "lossless", lossless_ok, "near", near_ok)
    if not (okd and lossless_ok and near_ok):
        bad += 1
print("TOTAL", n, "BAD", bad)
sys.exit(1 if bad else 0)
