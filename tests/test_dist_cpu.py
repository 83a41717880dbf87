"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: row-range sharding and the ragged all-gather
that replicates the build-side columns."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from csvplus_b200.dist import allgather_packed, allgather_ragged, pack_layout, shard_range


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(rank)
        # a ragged "column": rank r holds r*7+3 values
        lens = rng.integers(0, 9, size=rank * 7 + 3)
        off = np.zeros(len(lens) + 1, np.uint32); off[1:] = np.cumsum(lens)
        data = rng.integers(32, 127, size=int(off[-1]), dtype=np.uint8)
        po = allgather_ragged(torch.from_numpy(off.view(np.uint8).copy()), dist)
        pd = allgather_ragged(torch.from_numpy(data.copy()), dist)
        empty = allgather_ragged(torch.zeros(0 if rank == 0 else 5, dtype=torch.uint8), dist)
        # the same column + an empty/ragged segment through the packed (single-collective) gather
        pk = allgather_packed([torch.from_numpy(off.view(np.uint8).copy()), torch.from_numpy(data.copy()),
                               torch.zeros(0 if rank == 0 else 5, dtype=torch.uint8)], dist)
        q.put((rank, [p.numpy().tobytes() for p in po], [p.numpy().tobytes() for p in pd], [e.numel() for e in empty],
               off.tobytes(), data.tobytes(), [[s.numpy().tobytes() for s in r] for r in pk]))
    finally:
        dist.destroy_process_group()


def test_allgather_ragged_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    offs = [r[4] for r in res]; datas = [r[5] for r in res]
    for rank, po, pd, empty, _, _, pk in res:
        assert po == offs and pd == datas  # every rank sees every rank's columns, in rank order
        assert empty == [0, 5]
        assert [r[0] for r in pk] == offs and [r[1] for r in pk] == datas
        assert [len(r[2]) for r in pk] == [0, 5]


def test_pack_layout_alignment():
    offs, tot = pack_layout([0, 1, 16, 17, 0, 5])
    assert offs == [0, 0, 16, 32, 64, 64] and tot == 80
    assert pack_layout([]) == ([], 0)


def test_shard_range_partitions_rows():
    for total in (0, 1, 7, 100, 10_000_019):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


# ------------------------------------------------------------------ the library's all-gather-v (csrc/comm.cu)
def _column(rank):
    rng = np.random.default_rng(100 + rank)
    lens = rng.integers(0, 9, size=rank * 5 + 4)
    off = np.zeros(len(lens) + 1, np.uint32); off[1:] = np.cumsum(lens)
    data = rng.integers(32, 127, size=int(off[-1]), dtype=np.uint8)
    return off, data


def _layout_worker(rank, world, port, q):
    """what cpb_allgather_table does, with gloo carrying the bytes and numpy standing in for the device buffers: exchange
    the metadata, take the row / byte bases from the library's own layout routine, place every rank's offsets and bytes at
    its base, rebase the offsets by (byte base - first offset).  Rank 1 contributes a row-range VIEW (offsets not from 0)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from csvplus_b200 import _abi
        from csvplus_b200.dist import allgather_layout
        lib = _abi.load()
        off, data = _column(rank)
        lo = 2 if rank == 1 else 0  # a view: rows [lo, n)
        voff = off[lo:]
        meta = torch.tensor([len(voff) - 1, int(voff[0]), int(voff[-1])], dtype=torch.int64)
        allmeta = [torch.zeros_like(meta) for _ in range(world)]
        dist.all_gather(allmeta, meta)
        m = np.stack([a.numpy() for a in allmeta]).astype(np.uint64)
        rb, bb = allgather_layout(lib, m, 1)
        out_off = np.zeros(int(rb[-1]) + 1, np.uint32); out_data = np.zeros(int(bb[0][-1]), np.uint8)
        for r in range(world):  # the grouped broadcasts: root r's buffers land at its bases on every rank
            rows, first, end = (int(x) for x in m[r])
            seg_off = torch.from_numpy(voff[:rows].astype(np.int64)) if r == rank else torch.zeros(rows, dtype=torch.int64)
            seg_dat = torch.from_numpy(data[first:end].copy()) if r == rank else torch.zeros(end - first, dtype=torch.uint8)
            if rows:
                dist.broadcast(seg_off, src=r)
            if end > first:
                dist.broadcast(seg_dat, src=r)
            out_off[int(rb[r]): int(rb[r]) + rows] = (seg_off.numpy() - first + int(bb[0][r])).astype(np.uint32)
            out_data[int(bb[0][r]): int(bb[0][r + 1])] = seg_dat.numpy()
        out_off[-1] = int(bb[0][-1])
        q.put((rank, out_off.tobytes(), out_data.tobytes()))
    finally:
        dist.destroy_process_group()


def test_library_allgather_layout_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_layout_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # expected: rows of rank 0, then rows [2, n) of rank 1
    vals = []
    for r in range(world):
        off, data = _column(r)
        lo = 2 if r == 1 else 0
        vals += [data[off[i]:off[i + 1]].tobytes() for i in range(lo, len(off) - 1)]
    exp_off = np.zeros(len(vals) + 1, np.uint32); exp_off[1:] = np.cumsum([len(v) for v in vals])
    for rank, o, d in res:
        assert o == exp_off.tobytes() and d == b"".join(vals)


def test_allgather_layout_arithmetic():
    from csvplus_b200 import _abi
    from csvplus_b200.dist import allgather_layout
    lib = _abi.load()
    # 3 ranks, 2 columns: rows, (first, end) per column; rank 1 is empty, rank 2 is a view starting at byte 10 / 7
    meta = np.array([[4, 0, 20, 0, 9], [0, 0, 0, 0, 0], [3, 10, 25, 7, 7]], np.uint64)
    rb, bb = allgather_layout(lib, meta, 2)
    assert rb.tolist() == [0, 4, 4, 7]
    assert bb.tolist() == [[0, 20, 20, 35], [0, 9, 9, 9]]
