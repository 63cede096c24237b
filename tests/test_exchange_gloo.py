"""N > 1 host-side logic on CPU: world_size-2 gloo processes drive partition → all-to-all → per-rank operator, and the
PARTIAL → FINAL aggregation hand-off, with the oracle standing in for the GPU kernels (no GPU in this container)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, result_dir: str):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from presto_b200 import (AGG_COUNT_STAR, AGG_SUM_DOUBLE, BIGINT, DOUBLE, STEP_FINAL, STEP_PARTIAL, Aggregate, Block,
                             HashAggregationOperator, Page, PartitionedOutputOperator, to_pages)
    from presto_b200.exchange import all_to_all_pages, gather_pages_to_root
    from presto_b200.hashing import get_hash_block, get_partition
    ob = oracle.backend()
    rng = np.random.default_rng(100 + rank)
    n = 5000
    keys = rng.integers(0, 300, n).astype(np.int64)
    vals = rng.random(n)
    knull = (rng.random(n) < 0.05).astype(np.uint8) if rank == 0 else None     # only rank 0 carries a null map
    page = Page([Block.from_numpy(BIGINT, keys, knull), Block.from_numpy(DOUBLE, vals)])

    # ---- repartition by key: PartitionedOutput → all-to-all ----
    op = PartitionedOutputOperator([BIGINT, DOUBLE], world, [0], backend=ob)
    op.add_input(page)
    out = op.get_output()
    offsets = op.get_offsets()
    mine = all_to_all_pages(out, offsets, [BIGINT, DOUBLE])
    parts = get_partition(get_hash_block([mine.blocks[0]], ob).values, world)
    assert (parts == rank).all(), "every received row must belong to this rank's partition"

    # ---- each rank aggregates its partition completely (keys are co-partitioned → no second exchange) ----
    agg = HashAggregationOperator([BIGINT], [0], [Aggregate(AGG_SUM_DOUBLE, 1), Aggregate(AGG_COUNT_STAR)], backend=ob)
    rows = [r for p in to_pages(agg, [mine]) for r in p.rows()]

    # ---- PARTIAL on the unpartitioned shard → gather → FINAL on rank 0 (the Q1 plan) ----
    partial = HashAggregationOperator([BIGINT], [0], [Aggregate(AGG_SUM_DOUBLE, 1), Aggregate(AGG_COUNT_STAR)], step=STEP_PARTIAL, backend=ob, max_output_rows=1 << 30)
    pp = to_pages(partial, [page])
    gathered = gather_pages_to_root(pp[0] if pp else None)
    final_rows = None
    if rank == 0:
        final = HashAggregationOperator([BIGINT], [0], [Aggregate(AGG_SUM_DOUBLE, 1), Aggregate(AGG_COUNT_STAR, 2)], step=STEP_FINAL, backend=ob, max_output_rows=1 << 30)
        final_rows = [r for p in to_pages(final, gathered) for r in p.rows()]
    np.save(os.path.join(result_dir, f"in_{rank}.npy"), np.stack([np.where(knull == 1, -1, keys) if knull is not None else keys, vals]))
    import pickle
    with open(os.path.join(result_dir, f"out_{rank}.pkl"), "wb") as f:
        pickle.dump({"rows": rows, "final": final_rows, "received": mine.position_count}, f)
    dist.barrier()
    dist.destroy_process_group()


def test_partition_exchange_aggregate_world2(tmp_path):
    import pickle
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    inputs = [np.load(tmp_path / f"in_{r}.npy") for r in range(world)]
    outs = [pickle.load(open(tmp_path / f"out_{r}.pkl", "rb")) for r in range(world)]
    all_keys = np.concatenate([i[0] for i in inputs]).astype(np.int64)
    all_vals = np.concatenate([i[1] for i in inputs])
    assert sum(o["received"] for o in outs) == len(all_keys), "the exchange must neither lose nor duplicate rows"
    expect = {}
    for k, v in zip(all_keys, all_vals):
        key = None if k == -1 else int(k)      # -1 marks the NULL keys of rank 0 (real keys are >= 0)
        s, c = expect.get(key, (0.0, 0))
        expect[key] = (s + v, c + 1)
    # partitioned plan: the union of the per-rank results is the global aggregate, every group on exactly one rank
    seen = {}
    for o in outs:
        for k, s, c in o["rows"]:
            assert k not in seen, "a group must live on exactly one rank"
            seen[k] = (s, c)
    assert set(seen) == set(expect)
    for k, (s, c) in expect.items():
        assert seen[k][1] == c and abs(seen[k][0] - s) <= 1e-9 * max(1.0, abs(s))
    # partial → final plan gives the same aggregate
    final = {k: (s, c) for k, s, c in outs[0]["final"]}
    assert set(final) == set(expect)
    for k, (s, c) in expect.items():
        assert final[k][1] == c and abs(final[k][0] - s) <= 1e-9 * max(1.0, abs(s))
