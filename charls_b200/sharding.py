"""Work split of a batch of independent frames across the GPUs of one box.

A batch of frames shards naturally: frame i is coded by exactly one rank, no sample or stream byte ever moves between
GPUs.  The only exchange is the control plane -- every rank learns every frame's compressed size so that a global
offset table (where would frame i start in one concatenated output?) can be built."""
from __future__ import annotations


def frame_range(total_frames: int, world_size: int, rank: int) -> range:
    """Contiguous block of frames owned by `rank`; the first (total % world) ranks get one extra frame."""
    base, extra = divmod(total_frames, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def gather_sizes(local_sizes, total_frames: int, dist=None, device=None):
    """All-gathers per-frame stream sizes (uneven shards allowed). Returns the list of all sizes in frame order."""
    import torch

    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return list(local_sizes)
    world = dist.get_world_size()
    longest = max(len(frame_range(total_frames, world, r)) for r in range(world))
    padded = torch.zeros(longest, dtype=torch.int64, device=device)
    padded[: len(local_sizes)] = torch.tensor(list(local_sizes), dtype=torch.int64, device=device)
    gathered = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(gathered, padded)
    out = []
    for r in range(world):
        out += gathered[r][: len(frame_range(total_frames, world, r))].tolist()
    return out


def offset_table(sizes):
    """Exclusive prefix sum: byte offset of every frame in the concatenated output, plus the total."""
    offsets, total = [], 0
    for s in sizes:
        offsets.append(total)
        total += int(s)
    return offsets, total


# ---------------------------------------------------------------------------------------------------------------------
# One large frame across several GPUs: with one restart interval per line every line is an
# independent work item, so a frame splits into strips of whole lines.  Rank r codes its strip as an image of its own
# through the ordinary C ABI; the strips' entropy-coded segments, joined by the restart marker that a single encoder
# would have written between them, ARE the single-encoder stream (byte for byte), under one header whose frame height is
# the sum.  Strips start on multiples of eight lines so that the RSTm numbering (m = line mod 8, reference
# src/scan_decoder.hpp:335-349) continues across the joints.  Only compressed bytes move between ranks.
# Planar frames (one scan per component) are joined scan by scan.
# ---------------------------------------------------------------------------------------------------------------------
STRIP_ALIGNMENT = 8


def strip_range(height: int, world_size: int, rank: int) -> range:
    """Lines of the frame owned by `rank`: contiguous, starting on a multiple of eight lines, possibly empty."""
    groups = (height + STRIP_ALIGNMENT - 1) // STRIP_ALIGNMENT
    mine = frame_range(groups, world_size, rank)
    return range(min(mine.start * STRIP_ALIGNMENT, height), min(mine.stop * STRIP_ALIGNMENT, height))


def _segments(stream: bytes):
    """(marker, segment start, payload start) for the marker segments up to and including the first SOS."""
    if stream[:2] != b"\xff\xd8":
        raise ValueError("not a JPEG-LS stream")
    position, out = 2, []
    while True:
        if stream[position] != 0xFF:
            raise ValueError("marker expected")
        marker = stream[position + 1]
        length = int.from_bytes(stream[position + 2 : position + 4], "big")
        out.append((marker, position, position + 2 + length))
        position += 2 + length
        if marker == 0xDA:
            return out


def _scans(stream: bytes):
    """(bytes in front of the first SOS segment, [(bytes from the end of the previous scan to the end of this scan's SOS
    segment, entropy-coded bytes of the scan), ...]) of a complete stream."""
    import numpy as np

    data = np.frombuffer(stream, dtype=np.uint8)
    # inside entropy-coded data 0xFF is followed by a byte below 0x80; FF D0..D7 are restart markers; anything else ends it
    ends = np.flatnonzero((data[:-1] == 0xFF) & (data[1:] >= 0x80) & ((data[1:] & 0xF8) != 0xD0) & (data[1:] != 0xFF))
    segments = _segments(stream)
    first_sos = segments[-1][1]
    scans, position = [], first_sos
    while True:
        if stream[position : position + 2] != b"\xff\xda":
            raise ValueError("SOS expected")
        begin = position + 2 + int.from_bytes(stream[position + 2 : position + 4], "big")
        following = ends[ends >= begin]
        if len(following) == 0:
            raise ValueError("scan without an end")
        end = int(following[0])
        while end > begin and stream[end - 1] == 0xFF:  # fill bytes in front of the marker stay outside the scan
            end -= 1
        scans.append((stream[position:begin], stream[begin:end]))
        position = int(following[0])
        if stream[position + 1] == 0xD9:
            return stream[:first_sos], scans
        if stream[position + 1] != 0xDA:
            raise ValueError("marker segments between scans are not handled")


def _with_height(header: bytes, height: int) -> bytes:
    """The header with the frame height of its SOF55 segment replaced (heights above 65535 need the LSE form: not here)."""
    if height > 0xFFFF:
        raise ValueError("frames taller than 65535 lines are not split")
    for marker, start, _ in _segments(header + b"\xff\xda\x00\x02"):
        if marker == 0xF7:
            return header[: start + 5] + height.to_bytes(2, "big") + header[start + 7 :]
    raise ValueError("no SOF55 segment")


def stitch_strips(strip_streams, strip_heights) -> bytes:
    """Joins the streams of consecutive strips (each a complete restart-interval-1 stream of its lines) into the stream
    of the whole frame, scan by scan (planar frames have one scan per component).  Empty strips (height 0) are skipped."""
    header, scan_headers, pieces, total = None, None, None, 0
    for stream, height in zip(strip_streams, strip_heights):
        if height == 0:
            continue
        strip_header, scans = _scans(stream)
        if header is None:
            header, scan_headers, pieces = strip_header, [sos for sos, _ in scans], [[] for _ in scans]
        if len(scans) != len(pieces):
            raise ValueError("strips with different numbers of scans")
        if total % STRIP_ALIGNMENT != 0:
            raise ValueError("strips must start on a multiple of eight lines")
        for piece, (_, payload) in zip(pieces, scans):
            if piece:
                piece.append(bytes([0xFF, 0xD0 + (total - 1) % 8]))
            piece.append(payload)
        total += height
    if header is None:
        raise ValueError("no strips")
    body = b"".join(sos + b"".join(piece) for sos, piece in zip(scan_headers, pieces))
    return _with_height(header, total) + body + b"\xff\xd9"


def split_stream(stream: bytes, world_size: int):
    """The inverse: cuts a restart-interval-1 stream of a whole frame into one complete stream per rank (None for ranks
    without lines).  The marker positions come from vectorised passes over the bytes on the host."""
    import numpy as np

    header, scans = _scans(stream)
    height = None
    for marker, start, _ in _segments(stream):
        if marker == 0xF7:
            height = int.from_bytes(stream[start + 5 : start + 7], "big")
    cuts = []
    for _, payload in scans:
        data = np.frombuffer(payload, dtype=np.uint8)
        candidates = np.flatnonzero((data[:-1] == 0xFF) & ((data[1:] & 0xF8) == 0xD0))
        if len(candidates) != height - 1:
            raise ValueError("not a stream with one restart interval per line")
        cuts.append(candidates)
    out = []
    for rank in range(world_size):
        lines = strip_range(height, world_size, rank)
        if len(lines) == 0:
            out.append(None)
            continue
        body = b""
        for (sos, payload), candidates in zip(scans, cuts):
            begin = 0 if lines.start == 0 else int(candidates[lines.start - 1]) + 2
            end = len(payload) if lines.stop == height else int(candidates[lines.stop - 1])
            body += sos + payload[begin:end]
        out.append(_with_height(header, len(lines)) + body + b"\xff\xd9")
    return out


def encode_frame_split(image, encode_strip, dist=None, line_axis=0):
    """Every rank passes the whole frame (numpy array; lines along `line_axis`: 0 for [H, W] and [H, W, C], 1 for planar
    [C, H, W]) and `encode_strip(lines) -> bytes`, which codes an array of lines as a restart-interval-1 stream (e.g.
    functools.partial(charls_b200.codec.encode, ..., restart_interval=1)).  Returns the stream of the whole frame on
    every rank."""
    import numpy as np

    height = image.shape[line_axis]
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    lines = strip_range(height, world, rank)
    mine = encode_strip(np.ascontiguousarray(np.take(image, lines, axis=line_axis))) if len(lines) else b""
    if world == 1:
        streams = [mine]
    else:
        streams = [None] * world
        dist.all_gather_object(streams, mine)
    return stitch_strips(streams, [len(strip_range(height, world, r)) for r in range(world)])


def decode_frame_split(stream, decode_strip, dist=None, line_axis=0):
    """Every rank passes the stream of the whole frame and `decode_strip(stream) -> numpy lines`; returns the lines of the
    whole frame on every rank."""
    import numpy as np

    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    mine = split_stream(stream, world)[rank]
    lines = decode_strip(mine) if mine is not None else None
    if world == 1:
        return lines
    strips = [None] * world
    dist.all_gather_object(strips, lines)
    return np.concatenate([s for s in strips if s is not None], axis=line_axis)
