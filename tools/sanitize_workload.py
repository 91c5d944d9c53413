"""A small but varied pass over every kernel family, meant to run under compute-sanitizer (tools/sanitize_gpu.sh): tile kernels
(whole tiles, ragged right and bottom edges), the per-lane fallback (rows that are not 4-byte aligned, line interleave), the
general path (restart interval 0 and 4), side tables of interval offsets, the device-resident batch calls, and damaged streams
(truncated, bytes flipped, markers dropped) whose decode has to fail cleanly instead of reading or writing out of bounds.
Sizes are tiny: a sanitizer slows kernels down by one to two orders of magnitude.  Exits non-zero on a wrong result."""
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from charls_b200 import codec  # noqa: E402
from charls_b200.capi import CharlsError  # noqa: E402


def image(h, w, bits, cc, seed, layout, kind):
    rng = np.random.default_rng(seed)
    mx = (1 << bits) - 1
    if kind == "noise":
        a = rng.integers(0, mx + 1, size=(h, w, cc))
    elif kind == "flat":
        a = np.full((h, w, cc), mx // 3)
        a[h // 2 :, w // 3 :, :] = mx
    else:
        y, x = np.mgrid[0:h, 0:w]
        base = 0.8 * mx * (0.5 + 0.25 * np.sin(x / 9.7) + 0.25 * np.cos(y / 13.1))
        a = np.stack([np.clip(base * (1 - 0.1 * c) + rng.normal(0, 0.01 * mx, (h, w)), 0, mx) for c in range(cc)], axis=-1)
    a = a.astype(np.uint8 if bits <= 8 else np.uint16)
    if cc == 1:
        return a[:, :, 0].copy()
    return a.copy() if layout == "interleaved" else np.ascontiguousarray(a.transpose(2, 0, 1))


def round_trip(img, bits, near, ilv, xf, ri, label):
    stream = codec.encode(img, bits, near_lossless=near, interleave_mode=ilv, color_transformation=xf, restart_interval=ri)
    out, _, _ = codec.decode(stream)
    if near == 0:
        ok = np.array_equal(out, img)
    else:
        ok = int(np.abs(out.astype(np.int64) - img.astype(np.int64)).max()) <= near
    if not ok:
        raise SystemExit(f"round trip failed: {label}")
    return stream


def main():
    count = 0
    streams = []
    # tile kernels and their edges; per-lane fallback for odd row lengths; every component count
    configs = ((8, 1, 0, 0), (12, 1, 0, 0), (16, 1, 0, 0), (8, 3, 2, 1), (16, 3, 2, 1), (8, 3, 1, 0), (8, 3, 0, 0), (8, 4, 2, 0),
               (16, 2, 2, 0), (5, 1, 0, 0))
    key = (0, 2, 4)  # 8-bit mono, 16-bit mono, 16-bit RGB
    for (h, w) in ((70, 131), (33, 64), (5, 4099), (64, 17)):
        for index, (bits, cc, ilv, xf) in enumerate(configs):
            if (h, w) not in ((70, 131), (33, 64)) and index not in key:
                continue
            for near in (0, 2):
                if near and (xf or (h, w) != (70, 131)):
                    continue
                for kind in ("smooth", "noise", "flat"):
                    if kind != "smooth" and ((h, w) != (70, 131) or index not in key or near):
                        continue
                    img = image(h, w, bits, cc, count, "interleaved" if ilv else "planar", kind)
                    s = round_trip(img, bits, near, ilv, xf, 1, f"{h}x{w} {bits}b x{cc} ilv{ilv} xf{xf} near{near} {kind} ri1")
                    streams.append(s)
                    count += 1
    # general path: no restart interval, and intervals of four lines
    for bits, cc, ilv in ((8, 1, 0), (16, 3, 2), (8, 3, 1)):
        for ri in (0, 4):
            for near in (0, 3):
                img = image(37, 53, bits, cc, count, "interleaved" if ilv else "planar", "smooth")
                streams.append(round_trip(img, bits, near, ilv, 0, ri, f"general {bits}b x{cc} ilv{ilv} near{near} ri{ri}"))
                count += 1
    # side table of interval offsets through the single-image calls
    img = image(130, 96, 8, 1, 7, "planar", "smooth")
    with codec.JpegLSEncoder() as enc:
        enc.frame_info(96, 130, 8, 1).offset_table(True)
        dst = np.empty(enc.estimated_destination_size(), dtype=np.uint8)
        enc.destination(dst)
        n = enc.encode(img)
    table_stream = dst[:n].tobytes()
    out, _, _ = codec.decode(table_stream)
    if not np.array_equal(out, img):
        raise SystemExit("offset table round trip failed")
    streams.append(table_stream)
    count += 1

    # damaged streams: every decode must return (pixels or an error), never touch memory it does not own
    rng = np.random.default_rng(99)
    damaged = 0
    for s in streams[:: max(1, len(streams) // 45)] + [table_stream]:
        b = bytearray(s)
        variants = [bytes(b[: len(b) * 2 // 3]), bytes(b[:-3])]
        for _ in range(4):
            c = bytearray(b)
            for _ in range(3):
                c[int(rng.integers(len(c) // 4, len(c)))] = int(rng.integers(0, 256))
            variants.append(bytes(c))
        c = bytearray(b)  # drop a restart marker
        i = c.find(b"\xff\xd1")
        if i > 0:
            del c[i : i + 2]
            variants.append(bytes(c))
        for v in variants:
            try:
                codec.decode(v)
            except CharlsError:
                pass
            damaged += 1

    # device-resident batch calls (with and without side tables)
    import torch
    from charls_b200.batch import BatchCodec

    for table in (False, True):
        for (w, h, bits, cc, ilv, xf) in ((96, 40, 8, 1, 0, 0), (64, 33, 16, 3, 2, 1), (131, 40, 8, 1, 0, 0), (33, 35, 16, 3, 2, 1)):
            frames = np.stack([image(h, w, bits, cc, 500 + i, "interleaved", "smooth") for i in range(3)])
            dev = torch.from_numpy(frames.view(np.int16) if bits > 8 else frames).cuda()
            bc = BatchCodec(w, h, bits, cc, interleave_mode=ilv, color_transformation=xf, offset_table=table)
            st = torch.zeros((3, bc.stream_capacity), dtype=torch.uint8, device="cuda")  # the decoder reads whole words
            sizes = bc.encode(dev, st)
            back = bc.decode_new(st, sizes)
            torch.cuda.synchronize()
            if not torch.equal(back, dev):
                raise SystemExit("batch round trip failed")
            count += 3
    # host-batch calls (staging pipeline with the helper engine) and the two-part single-image calls
    for table in (False, True):
        w, h = (80, 48) if table else (81, 48)  # 81: rows that are not 4-byte aligned get an aligned pitch in staging
        frames = [image(h, w, 8, 1, 700 + i, "planar", "smooth") for i in range(9)]
        bc = BatchCodec(w, h, 8, 1, offset_table=table)
        outs = [np.empty(bc.stream_capacity, dtype=np.uint8) for _ in frames]
        sizes = bc.encode_host(frames, outs)
        back = [np.empty_like(f) for f in frames]
        bc.decode_host(outs, sizes, back)
        if not all(np.array_equal(a, b) for a, b in zip(frames, back)):
            raise SystemExit("host batch round trip failed")
        count += len(frames)
    img = image(40, 72, 16, 3, 900, "interleaved", "smooth")
    with codec.JpegLSEncoder() as enc:
        enc.frame_info(72, 40, 16, 3).interleave_mode(2).color_transformation(1)
        dst = np.empty(enc.estimated_destination_size(), dtype=np.uint8)
        enc.destination(dst)
        enc.encode_begin(img)
        n = enc.encode_end()
    with codec.JpegLSDecoder() as dec:
        dec.source(dst[:n].tobytes()).read_header()
        out = np.empty(dec.destination_size(), dtype=np.uint8)
        dec.decode_begin(out)
        dec.decode_end()
    if not np.array_equal(out.view("<u2").reshape(img.shape), img):
        raise SystemExit("two-part round trip failed")
    count += 1
    print(f"sanitize workload ok: {count} images, {damaged} damaged streams")


if __name__ == "__main__":
    main()
