"""Multi-GPU plumbing: the probe stream shards by row range with no data-path collective; the
only exchange step is the all-gather of the build-side columns so that every rank can build the full Index.
torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests) is plumbing here, not the product.

`allgather_ragged` is backend-agnostic (plain tensors) so that its padding / size-exchange logic is covered by
world_size-2 gloo tests on CPU; `allgather_table` wraps device columns of a cpb_table around it.
"""
from __future__ import annotations

import ctypes as C


def init_comm(ctx, dist, group=None) -> None:
    """gives `ctx` the library's own NCCL communicator (cpb_comm_init_rank): rank 0 draws the unique id and
    torch.distributed — any backend, it only carries 128 bytes — hands it to the other ranks"""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    buf = (C.c_uint8 * 128)()
    if rank == 0:
        st = ctx.lib.cpb_comm_unique_id(buf)
        if st:
            raise RuntimeError(f"cpb_comm_unique_id failed with status {st}")
    box = [bytes(buf)]
    dist.broadcast_object_list(box, src=0, group=group)
    raw = (C.c_uint8 * 128).from_buffer_copy(box[0])
    st = ctx.lib.cpb_comm_init_rank(ctx.h, world, rank, raw)
    if st:
        from .api import _raise
        _raise(st, None, ctx)


def allgather_table_nccl(ctx, table):
    """cpb_allgather_table: the library's all-gather-v of every column over its own communicator, straight into the
    concatenated table (one host sync for the sizes); collective over the ranks of init_comm / cpb_init_multi"""
    from .api import Table, _raise
    h = C.c_void_p()
    st = ctx.lib.cpb_allgather_table(ctx.h, table.h, C.byref(h))
    if st:
        _raise(st, None, ctx)
    return Table(ctx, h)


def allgather_layout(lib, meta, ncols: int):
    """cpb_allgather_layout (pure host arithmetic): meta = per rank [rows, first_0, end_0, first_1, end_1, ...] ->
    (row_base[nranks+1], byte_base[ncols][nranks+1])"""
    import numpy as np
    m = np.ascontiguousarray(meta, dtype=np.uint64)
    nranks = m.shape[0]
    rb = np.zeros(nranks + 1, np.uint64); bb = np.zeros((max(ncols, 1), nranks + 1), np.uint64)
    U = C.POINTER(C.c_uint64)
    st = lib.cpb_allgather_layout(nranks, ncols, m.ctypes.data_as(U), rb.ctypes.data_as(U), bb.ctypes.data_as(U))
    assert st == 0
    return rb, bb[:ncols]


def allgather_u64(ctx, values):
    """cpb_allgather_u64 over the library communicator -> per-rank lists (a ctx without a communicator: one rank)"""
    n = len(values)
    world = ctx.lib.cpb_comm_size(ctx.h)
    inp = (C.c_uint64 * n)(*[int(v) for v in values])
    out = (C.c_uint64 * (n * world))()
    st = ctx.lib.cpb_allgather_u64(ctx.h, inp, n, out)
    if st:
        from .api import _raise
        _raise(st, None, ctx)
    return [[int(out[r * n + i]) for i in range(n)] for r in range(world)]


class ShardedParse:
    """One file parsed by byte-range shards, one object per rank.  The three collective steps are explicit
    so that the same code runs one-process-per-GPU (exchange = allgather_u64 / torch.distributed) and, in the tests,
    as N simulated ranks on one GPU:

        sp = ShardedParse(ctx, rank, world, file_size, read)      # read(lo, hi) -> bytes of the file
        q = sp.step1_parity()                                     # -> exchange -> all_q
        info = sp.step2_parse(all_q, spec=..., pred=...)          # -> exchange -> all_info
        table, error = sp.step3_finish(all_info)                  # rows of this rank; the global DataSourceError

    A record belongs to the shard in whose (lo, hi] its first byte lies; concatenating the tables in rank order gives the
    reference's row order; the error, if any, is the reference's: first failing record, Line counted over the whole file
    (csvplus.go:1102-1137), rows after it dropped on every rank."""

    def __init__(self, ctx, rank: int, world: int, size: int, read, lookahead: int = 1 << 20, head: int = 1 << 20):
        self.ctx, self.rank, self.world, self.size = ctx, rank, world, size
        self.lo, self.hi = rank * size // world, (rank + 1) * size // world
        self.is_last = rank == world - 1
        self.buf = read(self.lo, size if self.is_last else min(size, self.hi + lookahead))
        self.head = read(0, min(size, head)) if rank > 0 else None

    def step1_parity(self) -> list:
        from .api import csv_quote_parity
        import numpy as np
        own = np.frombuffer(self.buf, np.uint8)[: self.hi - self.lo]
        return [csv_quote_parity(self.ctx, own)]

    def step2_parse(self, all_q, *, spec=None, pred=None, delimiter=",", num_fields=0) -> list:
        from .api import parse_csv, parse_csv_shard
        pin = 0
        for r in range(self.rank):
            pin ^= all_q[r][0] & 1
        hdr = True
        base = 2
        if self.rank > 0:
            # the header row lives in shard 0: resolve it from the head of the file (any rank can read it) — a data error
            # of the truncated head is not this rank's business, a header error (row 1) is everybody's
            t, err = parse_csv(self.ctx, self.head, spec=spec, delimiter=delimiter, num_fields=num_fields)
            if err is not None and err.Line <= 1:
                raise err
            fields, nf = t.parsed_from()
            spec = list(fields.items())
            num_fields = num_fields if num_fields != 0 else nf
            hdr = False
        self.table, self.records, self.err = parse_csv_shard(
            self.ctx, self.buf, own_bytes=self.hi - self.lo, shard_index=self.rank, is_last=self.is_last, initial_parity=pin,
            delimiter=delimiter, num_fields=num_fields, header_from_first_row=hdr, spec=spec, pred=pred)
        self.base = base
        return [self.records, 1 if self.err is not None else 0, self.err.Line if self.err is not None else 0]

    def step3_finish(self, all_info):
        from .api import DataSourceError
        first_bad = next((r for r in range(self.world) if all_info[r][1]), None)
        if first_bad is None:
            return self.table, None
        line = self.base + sum(all_info[r][0] for r in range(first_bad)) + all_info[first_bad][2]
        msg = self.err.Err if self.rank == first_bad else None
        if self.rank > first_bad:
            self.table = self.table.slice(0, 0)  # the reference stopped before these rows
        return self.table, (DataSourceError(line, msg) if msg is not None else DataSourceError(line, "(error raised by rank %d)" % first_bad))


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """contiguous row range of `rank` (row-range data parallelism; concatenation in rank order = input order)"""
    return rank * total // world, (rank + 1) * total // world


def allgather_ragged(t, dist, group=None):
    """all-gather-v of a 1-D tensor whose length differs per rank -> list of per-rank tensors (in rank order)"""
    import torch
    world = dist.get_world_size(group)
    n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(max(sizes), 1)
    pad = torch.zeros(mx, dtype=t.dtype, device=t.device)
    pad[: t.numel()] = t
    out = torch.empty(world * mx, dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return [out[r * mx: r * mx + sizes[r]] for r in range(world)]


def pack_layout(seg_bytes, align: int = 16):
    """byte offset of every segment in a packed buffer (each segment starts `align`-aligned) and the total size"""
    offs, pos = [], 0
    for b in seg_bytes:
        offs.append(pos)
        pos += (int(b) + align - 1) // align * align
    return offs, pos


def allgather_packed(segments, dist, group=None):
    """all-gather-v of SEVERAL ragged uint8 tensors with one size exchange and one data collective.

    Every rank passes the same number (>= 1) of 1-D uint8 tensors; segment k of rank r may have any length.  The
    segments are packed 16-byte aligned into one buffer, padded to the longest rank, and gathered with a single
    all_gather_into_tensor.  Returns out[r][k] = view of rank r's segment k (rank order).
    """
    import torch
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    k = len(segments)
    dev = segments[0].device
    meta = torch.tensor([s.numel() for s in segments], dtype=torch.int64, device=dev)
    allmeta = torch.empty(world * k, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allmeta, meta, group=group)
    sizes = allmeta.view(world, k).tolist()  # the one host sync of the exchange
    layouts = [pack_layout(sizes[r]) for r in range(world)]
    mx = max(max(tot for _, tot in layouts), 16)
    pack = torch.empty(mx, dtype=torch.uint8, device=dev)
    for s, o in zip(segments, layouts[rank][0]):
        if s.numel():
            pack[o: o + s.numel()].copy_(s)
    out = torch.empty(world * mx, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(out, pack, group=group)
    return [[out[r * mx + o: r * mx + o + sizes[r][j]] for j, o in enumerate(layouts[r][0])] for r in range(world)]


class _DevArr:
    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (max(nbytes, 1),), "typestr": "|u1", "data": (ptr, False), "version": 2}


def _as_tensor(ptr: int, nbytes: int):
    import torch
    return torch.as_tensor(_DevArr(ptr, nbytes), device="cuda")[:nbytes]


class _PendingGather:
    """an all-gather of a table in flight (allgather_table_async): the collective is enqueued on torch's current
    stream; wait() blocks until it is done and assembles the concatenated table"""

    def __init__(self, ctx, cols, parts, event, world, keep):
        self.ctx, self.cols, self.parts, self.event, self.world, self.keep = ctx, cols, parts, event, world, keep

    def wait(self):
        from .api import Table, _raise, _strs
        ctx, cols, parts = self.ctx, self.cols, self.parts
        self.event.synchronize()
        names, keep = _strs(cols)
        tabs = []
        for r in range(self.world):
            rows = parts[r][0].numel() // 4 - 1
            oa = (C.c_void_p * len(cols))(*[parts[r][2 * k].data_ptr() for k in range(len(cols))])
            da = (C.c_void_p * len(cols))(*[parts[r][2 * k + 1].data_ptr() if parts[r][2 * k + 1].numel() else 0
                                            for k in range(len(cols))])
            h = C.c_void_p()
            st = ctx.lib.cpb_table_from_device(ctx.h, len(cols), names, oa, da, rows, C.byref(h))
            if st:
                _raise(st, None, ctx)
            tabs.append(Table(ctx, h))
        ctx.sync()  # the parts were copied out of the torch buffers
        self.parts = self.keep = None
        return Table.concat(tabs) if self.world > 1 else tabs[0]


def allgather_table_async(ctx, table, dist, group=None):
    """starts the all-gather of `table`'s rows (one size exchange + one collective for all columns) and returns a
    handle; the caller may run other GPU work (e.g. parse its probe shard) before handle.wait().
    (torch.distributed plumbing of round 1, kept as a cross-check of allgather_table_nccl.  It ships whole column buffers:
    a row-range view whose first offset is not 0 — Top/Drop/Find results — is refused by cpb_table_from_device
    ("imported offsets must start at 0"); the library's all-gather, allgather_table_nccl, takes views.)"""
    import torch
    cols = table.columns
    n = len(table)
    ctx.sync()
    segs = []
    nb = C.c_uint64()
    for i, c in enumerate(cols):
        po, pd = table.device_column(c)
        ctx.lib.cpb_table_col_bytes(ctx.h, table.h, i, 0, n, C.byref(nb))
        segs.append(_as_tensor(po, 4 * (n + 1)))
        segs.append(_as_tensor(pd, nb.value))
    parts = allgather_packed(segs, dist, group)
    ev = torch.cuda.Event()
    ev.record()
    return _PendingGather(ctx, cols, parts, ev, dist.get_world_size(group), (table, segs))


def allgather_table(ctx, table, dist, group=None):
    """every rank contributes its rows; returns the concatenation of all ranks' rows (rank order) on every rank"""
    return allgather_table_async(ctx, table, dist, group).wait()
