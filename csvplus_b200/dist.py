"""Multi-GPU plumbing: the probe stream shards by row range with no data-path collective; the
only exchange step is the all-gather of the build-side columns so that every rank can build the full Index.
torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests) is plumbing here, not the product.

`allgather_ragged` is backend-agnostic (plain tensors) so that its padding / size-exchange logic is covered by
world_size-2 gloo tests on CPU; `allgather_table` wraps device columns of a cpb_table around it.
"""
from __future__ import annotations

import ctypes as C


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


class _DevArr:
    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (max(nbytes, 1),), "typestr": "|u1", "data": (ptr, False), "version": 2}


def _as_tensor(ptr: int, nbytes: int):
    import torch
    return torch.as_tensor(_DevArr(ptr, nbytes), device="cuda")[:nbytes]


def allgather_table(ctx, table, dist, group=None):
    """every rank contributes its rows; returns the concatenation of all ranks' rows (rank order) on every rank"""
    import torch

    from . import _abi
    from .api import Table, _raise, _strs
    cols = table.columns
    n = len(table)
    ctx.sync()
    parts_off, parts_data = [], []
    nb = C.c_uint64()
    for i, c in enumerate(cols):
        po, pd = table.device_column(c)
        ctx.lib.cpb_table_col_bytes(ctx.h, table.h, i, 0, n, C.byref(nb))
        parts_off.append(allgather_ragged(_as_tensor(po, 4 * (n + 1)), dist, group))
        parts_data.append(allgather_ragged(_as_tensor(pd, nb.value), dist, group))
    torch.cuda.synchronize()
    world = dist.get_world_size(group)
    names, keep = _strs(cols)
    tabs = []
    for r in range(world):
        rows = parts_off[0][r].numel() // 4 - 1
        oa = (C.c_void_p * len(cols))(*[parts_off[k][r].data_ptr() for k in range(len(cols))])
        da = (C.c_void_p * len(cols))(*[parts_data[k][r].data_ptr() if parts_data[k][r].numel() else 0 for k in range(len(cols))])
        h = C.c_void_p()
        st = ctx.lib.cpb_table_from_device(ctx.h, len(cols), names, oa, da, rows, C.byref(h))
        if st:
            _raise(st, None, ctx)
        tabs.append(Table(ctx, h))
    ctx.sync()  # the parts were copied out of the torch buffers
    return Table.concat(tabs) if world > 1 else tabs[0]
