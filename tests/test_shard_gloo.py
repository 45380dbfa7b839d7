"""CPU, world_size 2, gloo: the N>1 host logic (GOP discovery from packet headers, GOP -> rank assignment, order-
preserving gather of per-frame results).  The per-GOP pictures are produced here by the ORACLE decoder (this tier has
no GPU); the property checked is the one the multi-GPU path relies on: decoding each rank's GOP sub-streams
independently and merging by display index gives exactly the whole-stream decode, with no data-path collective."""
import hashlib
import os
import socket

import numpy as np
import pytest

import pfvo
from pretty_fast_video_b200 import codec, shard
from test_codec_cpu import oracle_stream


def _sha(fr):
    return hashlib.sha256(b"".join(p.tobytes() for p in fr)).hexdigest()


def _decode_display(data):
    """[(digest or 'drop')] per display slot with the oracle decoder"""
    dec = pfvo.Decoder(data)
    out = []
    while True:
        more, fr = dec.advance_frame()
        if not more:
            break
        out.append("drop" if fr is None else _sha(fr))
    return out


def _make_stream():
    # 5 GOPs of uneven length, one drop frame inside GOP 1
    return oracle_stream(96, 64, 17, 3, 4, 31, drop_at=(6,))[0]


def test_split_and_assign():
    data = _make_stream()
    info, _ = codec.parse_header(data)
    pk, _ = codec.index_packets(data, info.first_packet)
    gops = shard.split_gops(pk)
    assert [g.first_frame for g in gops] == [0, 4, 8, 12, 16]
    assert [g.nframes for g in gops] == [4, 4, 4, 4, 1]
    assert sum(g.nframes for g in gops) == 17
    # byte ranges tile the packet area exactly, EOF packet excluded
    assert gops[0].byte_range[0] == info.first_packet
    for a, b in zip(gops, gops[1:]):
        assert a.byte_range[1] == b.byte_range[0]
    assert gops[-1].byte_range[1] == len(data) - 5
    assert shard.assign(5, 0, 2) == [0, 2, 4] and shard.assign(5, 1, 2) == [1, 3]
    assert sorted(shard.assign(5, 0, 3) + shard.assign(5, 1, 3) + shard.assign(5, 2, 3)) == list(range(5))
    with pytest.raises(ValueError):
        shard.assign(5, 2, 2)


def test_substreams_reproduce_the_whole_stream_single_process():
    data = _make_stream()
    whole = _decode_display(data)
    info, gops, mine = shard.plan(data, 0, 1)
    got = []
    for g in mine:
        part = _decode_display(shard.substream(data, info.first_packet, g))
        assert len(part) == g.nframes
        got += [(g.first_frame + i, d) for i, d in enumerate(part)]
    assert shard.gather_ordered(got) == whole


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, data, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        info, gops, mine = shard.plan(data, rank, world)
        local = []
        for g in mine:
            part = _decode_display(shard.substream(data, info.first_packet, g))
            local += [(g.first_frame + i, d) for i, d in enumerate(part)]
        dist.barrier()
        merged = shard.gather_ordered(local, dist)
        q.put((rank, [g.index for g in mine], merged))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_sharded_decode_equals_whole_stream():
    import torch.multiprocessing as mp
    data = _make_stream()
    whole = _decode_display(data)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, data, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4] and res[1][1] == [1, 3]           # disjoint, covering
    assert res[0][2] == whole and res[1][2] == whole                 # both ranks see the full, ordered result


def test_overlapping_shards_are_detected():
    with pytest.raises(RuntimeError):
        shard.gather_ordered([(0, "a"), (0, "b")])
