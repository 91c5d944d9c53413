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


# ---------------------------------------------------------------------------------------------------------------------
# The same split with everything on the devices: strips are coded by charlsx_batch_* from / into CUDA tensors, their
# entropy-coded segments travel between the GPUs with NCCL (all_gather over NVLink) and are joined on the device; no sample
# and no stream byte visits the host (only the few header bytes the host has to parse).  Decoding uses the side table of
# interval offsets in the stream's header (APP11 "JLS-OFFT", include/charls_b200.h): a rank reads where its lines start
# and end instead of searching the whole stream for restart markers, and fetches only those bytes.
# ---------------------------------------------------------------------------------------------------------------------
def _header_and_scan_offset(prefix: bytes):
    """(bytes in front of the SOS segment, SOS segment) of a single-scan stream whose first bytes are `prefix`."""
    segments = _segments(prefix)
    sos_start, sos_end = segments[-1][1], segments[-1][2]
    return prefix[:sos_start], prefix[sos_start:sos_end]


def _without_offset_table(header: bytes) -> bytes:
    out, position = bytearray(header[:2]), 2
    for marker, start, end in _segments(header + b"\xff\xda\x00\x02")[:-1]:
        if not (marker == 0xEB and header[start + 4 : start + 12] == b"JLS-OFFT"):
            out += header[start:end]
        position = end
    return bytes(out)


def encode_frame_split_device(frame, make_codec, dist=None):
    """`frame`: CUDA tensor [H, W] or [H, W, C] (the whole frame on every rank; a rank only reads its own lines).
    `make_codec(height) -> charls_b200.batch.BatchCodec` for a strip of that many lines (restart interval 1, no offset table).
    Returns (stream as a CUDA uint8 tensor, size) on every rank: byte for byte what one GPU writes for the whole frame."""
    import torch

    height = frame.shape[0]
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    device = frame.device
    heights = [len(strip_range(height, world, r)) for r in range(world)]
    lines = strip_range(height, world, rank)
    payload = torch.empty(0, dtype=torch.uint8, device=device)
    header = sos = b""
    if len(lines):
        codec = make_codec(len(lines))
        strip = frame[lines.start : lines.stop].unsqueeze(0).contiguous()
        streams = torch.empty((1, codec.stream_capacity), dtype=torch.uint8, device=device)
        (size,) = codec.encode(strip, streams)
        header, sos = _header_and_scan_offset(streams[0, :1024].cpu().numpy().tobytes())
        begin = len(header) + len(sos)
        payload = streams[0, begin : size - 2]  # without EOI
        codec.close()
    sizes = torch.tensor([payload.numel()], dtype=torch.int64, device=device)
    if world > 1:
        all_sizes = torch.empty(world, dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(all_sizes, sizes)
        longest = int(all_sizes.max().item())
        padded = torch.zeros(longest, dtype=torch.uint8, device=device)
        padded[: payload.numel()] = payload
        gathered = torch.empty((world, longest), dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(gathered, padded)  # NCCL: the strips' bytes go GPU to GPU
        all_sizes = all_sizes.tolist()
        headers = [None] * world
        dist.all_gather_object(headers, (header, sos))  # a few dozen bytes of marker segments
        header, sos = next(h for h, n in zip(headers, heights) if n)
        pieces = [gathered[r, : all_sizes[r]] for r in range(world)]
    else:
        pieces = [payload]
    front = torch.frombuffer(bytearray(_with_height(header, height) + sos), dtype=torch.uint8).to(device)
    parts, done = [front], 0
    for r in range(world):
        if heights[r] == 0:
            continue
        if done:
            parts.append(torch.tensor([0xFF, 0xD0 + (done - 1) % 8], dtype=torch.uint8, device=device))
        parts.append(pieces[r])
        done += heights[r]
    parts.append(torch.tensor([0xFF, 0xD9], dtype=torch.uint8, device=device))
    stream = torch.cat(parts)
    return stream, int(stream.numel())


def decode_frame_split_device(stream, size, make_codec, dist=None):
    """`stream`: CUDA uint8 tensor with a single-scan restart-interval-1 stream that carries the side table of interval
    offsets (every rank has it, e.g. from a broadcast; a rank reads only its strip's bytes).  `make_codec(height)` as above.
    Returns the whole frame as a CUDA tensor on every rank ([H, row samples...] as the codec shapes it)."""
    import struct

    import torch

    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    device = stream.device
    # the marker segments in front of the scan: SOF55 (height) and the table; the table is at most ~4 bytes per line
    prefix = stream[: min(size, 1 << 16)].cpu().numpy().tobytes()
    height = None
    for marker, start, end in _segments_lenient(prefix):
        if marker == 0xF7:
            height = int.from_bytes(prefix[start + 5 : start + 7], "big")
    prefix = stream[: min(size, 1024 + 4 * (height + 1) + 64 * ((height + 1) // 16377 + 1))].cpu().numpy().tobytes()
    header, sos = _header_and_scan_offset(prefix)
    entries = []
    for marker, start, end in _segments(prefix)[:-1]:
        if marker == 0xEB and prefix[start + 4 : start + 12] == b"JLS-OFFT":
            first, count, total = struct.unpack(">III", prefix[start + 14 : start + 26])
            entries += list(struct.unpack(f">{count}I", prefix[start + 26 : end]))
    if len(entries) != height + 1:
        raise ValueError("the stream carries no complete side table of interval offsets")
    scan_begin = len(header) + len(sos)
    lines = strip_range(height, world, rank)
    heights = [len(strip_range(height, world, r)) for r in range(world)]
    decoded = None
    if len(lines):
        begin = scan_begin + entries[lines.start]
        end = scan_begin + entries[lines.stop] - (2 if lines.stop != height else 0)  # without the restart marker in front of the next strip
        front = torch.frombuffer(bytearray(_with_height(_without_offset_table(header), len(lines)) + sos), dtype=torch.uint8).to(device)
        strip_stream = torch.cat([front, stream[begin:end], torch.tensor([0xFF, 0xD9], dtype=torch.uint8, device=device)])
        codec = make_codec(len(lines))
        out = codec.decode_new(strip_stream.unsqueeze(0), [int(strip_stream.numel())])
        decoded = out[0]
        codec.close()
    if world == 1:
        return decoded
    longest = max(heights)
    row_shape = None
    shapes = [None] * world
    dist.all_gather_object(shapes, None if decoded is None else (tuple(decoded.shape[1:]), str(decoded.dtype)))
    row_shape, dtype_name = next(s for s in shapes if s is not None)
    dtype = getattr(torch, dtype_name.split(".")[-1])
    padded = torch.zeros((longest,) + row_shape, dtype=dtype, device=device)
    if decoded is not None:
        padded[: decoded.shape[0]] = decoded
    gathered = torch.empty((world, longest) + row_shape, dtype=dtype, device=device)
    dist.all_gather_into_tensor(gathered, padded)  # NCCL: decoded strips GPU to GPU
    return torch.cat([gathered[r, : heights[r]] for r in range(world) if heights[r]])


def _segments_lenient(prefix: bytes):
    """Like _segments, but stops quietly where the prefix ends (the table may reach beyond it)."""
    position, out = 2, []
    while position + 4 <= len(prefix) and prefix[position] == 0xFF:
        marker = prefix[position + 1]
        length = int.from_bytes(prefix[position + 2 : position + 4], "big")
        out.append((marker, position, position + 2 + length))
        position += 2 + length
        if marker == 0xDA:
            break
    return out
