"""Minimal JPEG-LS container reader/writer in pure Python -- test infrastructure only.

It exists so that tests can drive the scan-level oracle (oracle/jls_oracle.c) with complete streams and take apart the
streams the product writes.  Segment layouts follow ISO/IEC 14495-1 Annex C as the reference writes/reads them
(reference src/jpeg_stream_writer.cpp:20-245, src/jpeg_stream_reader.cpp:87-700).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field


@dataclass
class Scan:
    component_count: int
    component_ids: list
    near_lossless: int
    interleave_mode: int
    restart_interval: int
    data_offset: int  # offset of the first entropy-coded byte
    data_end: int = -1  # offset of the 0xFF of the marker that ends the scan (next SOS/DNL/EOI...)


@dataclass
class Stream:
    width: int = 0
    height: int = 0
    bits_per_sample: int = 0
    component_count: int = 0
    color_transformation: int = 0
    pc: tuple | None = None  # (MAXVAL, T1, T2, T3, RESET) from an LSE type-1 segment
    restart_interval: int = 0
    scans: list = field(default_factory=list)
    comments: list = field(default_factory=list)
    segments: list = field(default_factory=list)  # (marker, offset, payload)


def find_scan_end(data: bytes, pos: int) -> int:
    """Offset of the first marker that is not RSTm / stuffed data at or after `pos`."""
    n = len(data)
    while True:
        i = data.find(b"\xff", pos)
        if i < 0 or i + 1 >= n:
            return n
        nxt = data[i + 1]
        if nxt == 0xFF:  # fill byte
            pos = i + 1
            continue
        if nxt < 0x80 or 0xD0 <= nxt <= 0xD7:
            pos = i + 2
            continue
        return i


def parse(data: bytes) -> Stream:
    s = Stream()
    assert data[:2] == b"\xff\xd8", "SOI missing"
    pos = 2
    restart_interval = 0
    while pos < len(data):
        assert data[pos] == 0xFF, f"marker expected at {pos}"
        while data[pos + 1] == 0xFF:
            pos += 1
        marker = data[pos + 1]
        pos += 2
        if marker == 0xD9:
            break
        (length,) = struct.unpack(">H", data[pos : pos + 2])
        payload = data[pos + 2 : pos + length]
        s.segments.append((marker, pos - 2, payload))
        pos += length
        if marker == 0xF7:  # SOF55
            s.bits_per_sample = payload[0]
            s.height, s.width = struct.unpack(">HH", payload[1:5])
            s.component_count = payload[5]
        elif marker == 0xF8:  # LSE
            if payload[0] == 1:
                s.pc = struct.unpack(">5H", payload[1:11])
            elif payload[0] == 4:
                wxy = payload[1]
                s.height = int.from_bytes(payload[2 : 2 + wxy], "big")
                s.width = int.from_bytes(payload[2 + wxy : 2 + 2 * wxy], "big")
        elif marker == 0xDD:  # DRI (2, 3 or 4 byte payload)
            restart_interval = int.from_bytes(payload, "big")
            s.restart_interval = restart_interval
        elif marker == 0xE8 and len(payload) == 5 and payload[:4] == b"mrfx":
            s.color_transformation = payload[4]
        elif marker == 0xFE:
            s.comments.append(bytes(payload))
        elif marker == 0xDC:  # DNL
            s.height = int.from_bytes(payload, "big")
        elif marker == 0xDA:  # SOS
            ns = payload[0]
            ids = [payload[1 + 2 * i] for i in range(ns)]
            near = payload[1 + 2 * ns]
            ilv = payload[2 + 2 * ns]
            scan = Scan(ns, ids, near, ilv, restart_interval, pos)
            scan.data_end = find_scan_end(data, pos)
            s.scans.append(scan)
            pos = scan.data_end
    return s


def _seg(marker: int, payload: bytes) -> bytes:
    return bytes([0xFF, marker]) + struct.pack(">H", len(payload) + 2) + payload


def write_stream(
    width,
    height,
    bits_per_sample,
    component_count,
    scans,  # list of (component_count_in_scan, near, ilv, entropy_bytes)
    *,
    color_transformation=0,
    pc=None,
    restart_interval=0,
) -> bytes:
    """SOI [APP8 mrfx] SOF55 [LSE] [DRI] (SOS data)* EOI, the layout the reference encoder writes (+DRI)."""
    out = bytearray(b"\xff\xd8")
    if color_transformation:
        out += _seg(0xE8, b"mrfx" + bytes([color_transformation]))
    sof = bytes([bits_per_sample]) + struct.pack(">HH", height, width) + bytes([component_count])
    for c in range(component_count):
        sof += bytes([c + 1, 0x11, 0])
    out += _seg(0xF7, sof)
    if pc is not None:
        out += _seg(0xF8, b"\x01" + struct.pack(">5H", *pc))
    if restart_interval:
        out += _seg(0xDD, struct.pack(">H", restart_interval) if restart_interval < 65536 else struct.pack(">I", restart_interval))
    cid = 1
    for ns, near, ilv, payload in scans:
        sos = bytes([ns])
        for _ in range(ns):
            sos += bytes([cid, 0])
            cid += 1
        sos += bytes([near, ilv, 0])
        out += _seg(0xDA, sos) + payload
    out += b"\xff\xd9"
    return bytes(out)


# ---------------------------------------------------------------------------------------------------------------------
# Side table of interval offsets (extension, charls_b200/csrc/jls_common.h): APP11 "JLS-OFFT" segments in front of SOS
# ---------------------------------------------------------------------------------------------------------------------
OFFSET_TABLE_ID = b"JLS-OFFT"
OFFSET_TABLE_ENTRIES_PER_SEGMENT = (65533 - 22) // 4


def interval_starts(data: bytes) -> list:
    """Offsets of the interval starts of one scan's entropy-coded bytes plus their total length: what the table lists."""
    starts, pos = [0], 0
    while True:
        i = data.find(b"\xff", pos)
        if i < 0 or i + 1 >= len(data):
            break
        if 0xD0 <= data[i + 1] <= 0xD7:
            starts.append(i + 2)
            pos = i + 2
        else:
            pos = i + 1
    return starts + [len(data)]


def offset_table_segments(entries: list) -> bytes:
    out = bytearray()
    total = len(entries)
    for first in range(0, total, OFFSET_TABLE_ENTRIES_PER_SEGMENT):
        part = entries[first : first + OFFSET_TABLE_ENTRIES_PER_SEGMENT]
        payload = OFFSET_TABLE_ID + bytes([1, 0]) + struct.pack(">III", first, len(part), total) + b"".join(struct.pack(">I", e) for e in part)
        out += _seg(0xEB, payload)
    return bytes(out)


def read_offset_tables(stream: bytes) -> list:
    """The tables of a stream, one list of entries per scan that has one (segments in order)."""
    tables, current = [], []
    for marker, _, payload in parse(stream).segments:
        if marker == 0xEB and payload[:8] == OFFSET_TABLE_ID:
            first, count, total = struct.unpack(">III", payload[10:22])
            assert first == len(current) and len(payload) == 22 + 4 * count
            current += list(struct.unpack(f">{count}I", payload[22:]))
            if len(current) == total:
                tables.append(current)
                current = []
    return tables


def with_offset_table(stream: bytes, entries: list | None = None) -> bytes:
    """The single-scan stream with a side table in front of its SOS (entries: default = the true interval starts)."""
    s = parse(stream)
    assert len(s.scans) == 1
    scan = s.scans[0]
    sos = next(off for marker, off, _ in s.segments if marker == 0xDA)
    if entries is None:
        entries = interval_starts(stream[scan.data_offset : scan.data_end])
    return stream[:sos] + offset_table_segments(entries) + stream[sos:]


def without_offset_table(stream: bytes) -> bytes:
    out, pos = bytearray(), 0
    for marker, off, payload in parse(stream).segments:
        if marker == 0xEB and payload[:8] == OFFSET_TABLE_ID:
            out += stream[pos:off]
            pos = off + 4 + len(payload)
    return bytes(out + stream[pos:])
