"""2-rank GPU worker of tests/test_gpu_scale.py::test_two_rank_allgathered_three_way_join (torchrun, NCCL):
each rank parses half of the customers file, the columns are all-gathered, every rank builds the full index and
joins its own shard of the orders; the result of every rank is compared with the oracle's rows of that shard."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import csvplus_b200 as cp
    from csvplus_b200.dist import allgather_table, allgather_table_nccl, init_comm
    from oracle import oracle as orc
    from tests.helpers import assert_table_equals_oracle
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = cp.Context(local)
    SEED = 0xC5B200
    n_orders, n_cust, n_prod = 200_000, 60_001, 5_000  # per-rank orders; an odd customers count: unequal shards
    lo, hi = rank * n_cust // world, (rank + 1) * n_cust // world
    shard = ctx.gen_csv("customers", (lo, hi), seed=SEED, n_cust=n_cust, permute=True, header=True)
    tc, err = cp.parse_csv(ctx, shard, spec=[("id", -1), ("name", -1), ("surname", -1)])
    assert err is None
    init_comm(ctx, dist)
    full = allgather_table_nccl(ctx, tc)          # the library's collective (cpb_allgather_table)
    assert len(full) == n_cust
    sliced = allgather_table_nccl(ctx, tc.slice(5, len(tc)))  # row-range views are accepted (offsets do not start at 0)
    assert len(sliced) == n_cust - 5 * world
    via_torch = allgather_table(ctx, tc, dist)   # the torch.distributed plumbing of round 1 gives the same table
    for c in full.columns:
        a, b = full.column(c), via_torch.column(c)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), c
    cidx = full.index_on("id", unique=True)
    prod = ctx.gen_csv("products", (0, n_prod), seed=SEED, n_prod=n_prod, permute=True)
    tp, _ = cp.parse_csv(ctx, prod, spec=[("prod_id", -1), ("product", -1), ("price", -1)])
    pidx = tp.index_on("prod_id", unique=True)
    orders = ctx.gen_csv("orders", (rank * n_orders, (rank + 1) * n_orders), seed=SEED, n_cust=n_cust, n_prod=n_prod)
    to, _ = cp.parse_csv(ctx, orders, spec=[("cust_id", -1), ("prod_id", -1), ("qty", -1), ("ts", -1)])
    j = to.join(cidx, "cust_id").join(pidx)
    # oracle on this rank's probe shard against the WHOLE build side
    cust_all = ctx.gen_csv("customers", (0, n_cust), seed=SEED, n_cust=n_cust, permute=True).to_host()
    ocidx = orc.reader_rows(cust_all, select=["id", "name", "surname"]).unique_index_on("id")
    opidx = orc.reader_rows(prod.to_host(), select=["prod_id", "product", "price"]).unique_index_on("prod_id")
    oj = orc.reader_rows(orders.to_host(), select=["cust_id", "prod_id", "qty", "ts"]).join(ocidx, "cust_id").join(opidx)
    assert len(j) == n_orders
    assert_table_equals_oracle(j, oj)
    # the gathered table itself is the concatenation of the shards in rank order
    og = orc.reader_rows(cust_all, select=["id", "name", "surname"])
    assert_table_equals_oracle(full, og)
    dist.barrier()
    print(f"RANK{rank}_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
