"""Frame-parallel sharding of a PFV stream over the GPUs of one box.

The path has no exchange step: macroblocks of a frame are independent, key frames are independent, and a P frame
needs only the previous reconstructed frame of its own GOP (src/dec.rs:425-432 reads self.framebuffer), so the unit
of independence is the GOP = a key frame plus everything up to the next key frame.  GOP boundaries are found from
the packet headers alone (u8 type + u32 len, src/dec.rs:179-180; no entropy decoding), GOP g goes to rank g mod N,
every rank decodes its GOPs on its own GPU with its own context, and NO collective runs on the data path.
torch.distributed is used only for the control plane: the barrier around timed regions and, in `gather_ordered`,
collecting small per-frame results (checksums, timings) back into stream order.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple

from . import codec


@dataclass(frozen=True)
class Gop:
    index: int                  # GOP number in the stream
    first_frame: int            # display index of its first picture (drop frames count as a display slot)
    packets: Tuple[int, int]    # [begin, end) into the packet list
    nframes: int                # display slots in the GOP (pictures + drop frames)
    byte_range: Tuple[int, int]  # [begin, end) of the GOP's packets in the stream


def split_gops(packets: Sequence[Tuple[int, int, int]]) -> List[Gop]:
    """packets = [(type, len, payload_offset)] from codec.index_packets.  A GOP starts at every key frame
    (type 1, len > 0).  Packets before the first key frame (a stream that starts on a P frame is undecodable
    without its GOP head, SURVEY §8e caveat) are attached to a GOP 0 that starts at packet 0."""
    gops: List[Gop] = []
    start, frame0, nfr = None, 0, 0
    display = 0

    def close(end_idx):
        nonlocal start
        if start is None:
            return
        b = packets[start][2] - 5
        last = packets[end_idx - 1]
        e = last[2] + (last[1] if last[0] != 0 else 0)
        gops.append(Gop(len(gops), frame0, (start, end_idx), nfr, (b, e)))
        start = None

    for i, (t, l, p) in enumerate(packets):
        if t == 0:
            close(i)
            break
        is_key = t == 1 and l > 0
        shows = (t == 1) or (t == 2)                                  # pictures and drop frames take a display slot
        if is_key or start is None:
            close(i)
            start, frame0, nfr = i, display, 0
        if shows:
            nfr += 1
            display += 1
    else:
        close(len(packets))
    return gops


def assign(ngops: int, rank: int, world: int) -> List[int]:
    """GOP g -> rank g mod world (round robin keeps every rank's GOPs spread over the stream)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, ngops, world))


def substream(data: bytes, header_end: int, gop: Gop) -> bytes:
    """A self-contained stream for one GOP: the file header, the GOP's packets, an EOF packet.  Feeding it to a
    Decoder reproduces exactly the pictures the whole-stream decode yields for that GOP (the key frame overwrites
    every macroblock of the framebuffer, src/common.rs:477-496)."""
    b, e = gop.byte_range
    return data[:header_end] + data[b:e] + b"\x00\x00\x00\x00\x00"


def plan(data: bytes, rank: int, world: int):
    """-> (info, all gops, this rank's gops)"""
    info, _ = codec.parse_header(data)
    packets, truncated = codec.index_packets(data, info.first_packet)
    if truncated:
        raise codec.DecodeError(-8, "stream is cut short: cannot shard it safely")
    gops = split_gops(packets)
    return info, gops, [gops[g] for g in assign(len(gops), rank, world)]


def gather_ordered(local: List[Tuple[int, object]], dist=None) -> List[object]:
    """Control plane only: every rank contributes [(display_index, small result)]; returns the results of all ranks
    in display order on every rank.  `dist` is torch.distributed (initialised) or None for a single process."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        merged = list(local)
    else:
        parts = [None] * dist.get_world_size()
        dist.all_gather_object(parts, local)
        merged = [x for part in parts for x in part]
    merged.sort(key=lambda kv: kv[0])
    idx = [k for k, _ in merged]
    if idx != sorted(set(idx)):
        raise RuntimeError("shards overlap: a display index was produced twice")
    return [v for _, v in merged]
