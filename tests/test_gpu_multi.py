"""-m gpu, needs >= 2 GPUs (skipped otherwise): the repartitioned Q3 plan over NCCL — PartitionedOutput → all-to-all → join →
aggregation on every rank — returns, as the union over ranks, exactly the groups the single-GPU plan returns."""
import os
import pickle
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir, fused):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    from presto_b200 import BIGINT, DATE, DOUBLE, INTEGER, Block, Page
    from presto_b200._lib import load
    from presto_b200.pipelines import run_q3_distributed
    from presto_b200.synthetic import q3_tables_np
    backend = load()
    backend.check(backend.set_device(rank))
    customer, orders, lineitem = q3_tables_np(60_000)
    dev = f"cuda:{rank}"

    def shard_pages(types, arrays, n_pages):
        arrays = [a[rank::world] for a in arrays]
        n = len(arrays[0])
        cuts = np.linspace(0, n, n_pages + 1).astype(int)
        pages = []
        for i in range(n_pages):
            tensors = [torch.from_numpy(np.ascontiguousarray(a[cuts[i]:cuts[i + 1]])).to(dev) for a in arrays]
            pages.append(Page([Block.device(t, x.numel(), x.data_ptr(), keepalive=x) for t, x in zip(types, tensors)]))
        return pages

    stats = {}
    out = run_q3_distributed(shard_pages([BIGINT, INTEGER], customer, 1), shard_pages([BIGINT, BIGINT, DATE, INTEGER], orders, 2),
                             shard_pages([BIGINT, DOUBLE, DOUBLE, DATE], lineitem, 3), backend=backend, stats=stats, device=rank, fused=fused)
    from presto_b200.exchange import page_to_host
    rows = [r for p in out for r in page_to_host(p).rows()]
    with open(os.path.join(out_dir, f"q3_{rank}.pkl"), "wb") as f:
        pickle.dump({"rows": rows, "stats": stats}, f)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("fused", [True, False], ids=["fused_scatter_exchange", "nccl_all_to_all"])
def test_q3_repartitioned_across_two_gpus(gpu_backend, oracle_backend, tmp_path, fused):
    import ctypes as C
    n = C.c_int32()
    gpu_backend.check(gpu_backend.device_count(C.byref(n)))
    if n.value < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from presto_b200 import BIGINT, DATE, DOUBLE, INTEGER
    from presto_b200.hashing import get_partition, hash_long
    from presto_b200.pipelines import run_q3
    from presto_b200.synthetic import q3_tables_np
    from test_gpu_pipelines import pages_of
    from util import rows_of
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), fused), nprocs=world, join=True)
    customer, orders, lineitem = q3_tables_np(60_000)
    want = rows_of(run_q3(pages_of([BIGINT, INTEGER], customer, 1), pages_of([BIGINT, BIGINT, DATE, INTEGER], orders, 2),
                          pages_of([BIGINT, DOUBLE, DOUBLE, DATE], lineitem, 3), backend=oracle_backend))
    got = {}
    for r in range(world):
        d = pickle.load(open(tmp_path / f"q3_{r}.pkl", "rb"))
        for row in d["rows"]:
            key = row[:3]
            assert key not in got, "a group must live on exactly one rank"
            # the group's rank is the partition HashGenerator.getPartition assigns to its orderkey
            assert get_partition(hash_long(np.array([key[0]], dtype=np.int64)), world)[0] == r
            got[key] = row[3]
    assert set(got) == {w[:3] for w in want}
    for w in want:
        assert abs(got[w[:3]] - w[3]) <= 1e-6 * abs(w[3])


def _exchange_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    from presto_b200 import BIGINT, DOUBLE, INTEGER, MEM_DEVICE, Block, Page, PartitionedOutputOperator
    from presto_b200._lib import load
    from presto_b200.exchange import FusedExchange, all_to_all_pages, page_to_host
    backend = load()
    backend.check(backend.set_device(rank))
    types = [BIGINT, INTEGER, DOUBLE]
    fx = FusedExchange(types, 100, with_nulls=True, group=None, backend=backend)       # too small on purpose: must regrow collectively
    op_f = PartitionedOutputOperator(types, world, [0], output_memory=MEM_DEVICE, backend=backend, device=rank)
    op_n = PartitionedOutputOperator(types, world, [0], output_memory=MEM_DEVICE, backend=backend, device=rank)
    same = True
    for it, n in enumerate([5000, 0 if rank == 0 else 3000, 1, 70_001]):
        rng = np.random.default_rng(1000 * it + rank)
        cols = [rng.integers(-50, 50, n), rng.integers(-2**31, 2**31, n).astype(np.int32), rng.random(n)]
        nulls = [rng.random(n) < 0.1, None, rng.random(n) < 0.5]
        keep = []
        blocks = []
        for t, v, m in zip(types, cols, nulls):
            tv = torch.from_numpy(np.ascontiguousarray(v)).to(f"cuda:{rank}")
            tm = None if m is None else torch.from_numpy(m.astype(np.uint8)).to(f"cuda:{rank}")
            keep += [tv, tm]
            blocks.append(Block.device(t, n, tv.data_ptr() if n else 0, tm.data_ptr() if (tm is not None and n) else 0, keepalive=(tv, tm)))
        page = Page(blocks, n)
        got = fx.exchange(op_f, page)
        got_rows = page_to_host(got).rows() if got.position_count else []
        if n:
            op_n.add_input(page)
            slab, offs = op_n.get_output(copy=False), op_n.get_offsets()
        else:
            slab, offs = page, [0] * (world + 1)
        want = all_to_all_pages(slab, offs, types)
        want_rows = page_to_host(want).rows() if want.position_count else []
        same = same and got_rows == want_rows and int(fx.last_counts[:, rank].sum()) == len(want_rows)
    fx.close(); op_f.close(); op_n.close()
    with open(os.path.join(out_dir, f"x_{rank}.pkl"), "wb") as f:
        pickle.dump({"same": same, "capacity": fx.capacity}, f)
    dist.barrier()
    dist.destroy_process_group()


def test_fused_exchange_delivers_what_all_to_all_delivers(gpu_backend, tmp_path):
    """Peer-memory scatter == PartitionedOutput slab + all_to_all_single, row for row and in the same (sender-major) order, with null
    maps, an empty page on one rank, and a receive capacity that has to regrow."""
    import ctypes as C
    n = C.c_int32()
    gpu_backend.check(gpu_backend.device_count(C.byref(n)))
    if n.value < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_exchange_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        d = pickle.load(open(tmp_path / f"x_{r}.pkl", "rb"))
        assert d["same"]
        assert d["capacity"] > 100


def test_count_then_scatter_to_matches_add_input(gpu_backend):
    """One GPU: pg_partitioned_output_count + pg_partitioned_output_scatter_to into P separate destination columns give the
    partitions pg_operator_add_input leaves back to back (OptimizedPartitionedOutputOperator.java:384-824 order: input order
    within a partition)."""
    import ctypes as C
    from presto_b200 import BIGINT, DOUBLE, MEM_HOST, Block, Page, PartitionedOutputOperator
    rng = np.random.default_rng(5)
    P, n = 7, 40_013
    keys, vals = rng.integers(0, 1000, n), rng.random(n)
    nulls = rng.random(n) < 0.2
    page = Page([Block.from_numpy(BIGINT, keys, nulls), Block.from_numpy(DOUBLE, vals)])
    ref = PartitionedOutputOperator([BIGINT, DOUBLE], P, [0], output_memory=MEM_HOST, backend=gpu_backend)
    ref.add_input(page)
    slab, offs = ref.get_output(), ref.get_offsets()
    op = PartitionedOutputOperator([BIGINT, DOUBLE], P, [0], backend=gpu_backend)
    counts = op.count(page)
    assert counts.tolist() == np.diff(offs).tolist()
    lead = 3                      # rows another sender would have put first
    dv = np.zeros((P, 2), dtype=np.uint64); dn = np.zeros((P, 2), dtype=np.uint64)
    allocs = []
    for p in range(P):
        for c, w in enumerate((8, 8)):
            for table, nbytes in ((dv, (counts[p] + lead) * w), (dn, counts[p] + lead)):
                ptr = C.c_void_p()
                gpu_backend.check(gpu_backend.device_alloc(max(int(nbytes), 16), C.byref(ptr)))
                table[p, c] = ptr.value; allocs.append(ptr)
    op.scatter_to(dv, dn, np.full(P, lead, dtype=np.int64))
    for p in range(P):
        k = np.empty(counts[p] + lead, dtype=np.int64); v = np.empty(counts[p] + lead, dtype=np.float64)
        kn = np.empty(counts[p] + lead, dtype=np.uint8); vn = np.empty(counts[p] + lead, dtype=np.uint8)
        if counts[p]:
            for host, table, c in ((k, dv, 0), (v, dv, 1), (kn, dn, 0), (vn, dn, 1)):
                gpu_backend.check(gpu_backend.memcpy_d2h(host.ctypes.data, C.c_void_p(int(table[p, c])), host.nbytes))
        lo, hi = offs[p], offs[p + 1]
        want_null = slab.blocks[0].is_null[lo:hi].astype(bool)
        assert (kn[lead:].astype(bool) == want_null).all()
        assert (k[lead:][~want_null] == slab.blocks[0].values[lo:hi][~want_null]).all()
        assert (v[lead:] == slab.blocks[1].values[lo:hi]).all()
        assert not vn[lead:].any()
    for a in allocs:
        gpu_backend.device_free(a)
    op.close(); ref.close()


def _q1_pages(seed, n):
    from presto_b200 import DATE, DOUBLE, Block, Page
    rng = np.random.default_rng(seed)
    rf = [b"ANR"[i:i + 1] for i in rng.integers(0, 3, n)]; ls = [b"FO"[i:i + 1] for i in rng.integers(0, 2, n)]
    return Page([Block.from_strings(rf), Block.from_strings(ls), Block.from_numpy(DOUBLE, rng.integers(1, 51, n).astype(np.float64)),
                 Block.from_numpy(DOUBLE, np.round(rng.random(n) * 1e5, 2)), Block.from_numpy(DOUBLE, rng.integers(0, 11, n) / 100),
                 Block.from_numpy(DOUBLE, rng.integers(0, 9, n) / 100), Block.from_numpy(DATE, rng.integers(8035, 10592, n).astype(np.int32))])


def _q1_operator(backend, step, **kw):
    from presto_b200 import DATE, DOUBLE, VARCHAR, FusedAggregationOperator
    from presto_b200.pipelines import Q1_AGGREGATES, q1_program_typed
    types = [VARCHAR, VARCHAR, DOUBLE, DOUBLE, DOUBLE, DOUBLE, DATE]
    return FusedAggregationOperator(types, q1_program_typed(VARCHAR), [], [VARCHAR, VARCHAR], [0, 1], Q1_AGGREGATES, step=step, expected_groups=1024,
                                    group_max_lengths=[1, 1], backend=backend, max_output_rows=1 << 30, **kw)


def _state_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    from presto_b200 import STEP_PARTIAL, STEP_SINGLE
    from presto_b200._lib import load
    from presto_b200.exchange import merge_aggregation_states
    backend = load()
    backend.check(backend.set_device(rank))
    rows = {}
    for shared in (False, True):          # the operator on its own stream / on torch's current stream
        kw = dict(cuda_stream=torch.cuda.current_stream().cuda_stream) if shared else {}
        op = _q1_operator(backend, STEP_SINGLE if rank == 0 else STEP_PARTIAL, device=rank, **kw)
        op.add_input(_q1_pages(100 + rank, 40_000 + 10_000 * rank))
        if merge_aggregation_states(op, max_groups=64):
            op.finish()
            rows[shared] = op.get_output().rows()
        torch.cuda.synchronize()
        op.close()
    with open(os.path.join(out_dir, f"s_{rank}.pkl"), "wb") as f:
        pickle.dump(rows, f)
    dist.barrier()
    dist.destroy_process_group()


def test_aggregation_states_merge_across_two_gpus(gpu_backend, oracle_backend, tmp_path):
    """Q1 on two GPUs: each aggregates its shard, rank 1's group table travels as a device-side state (one all-gather) and is merged
    into rank 0's operator — equal to ONE operator over rank 0's page then rank 1's (the oracle, SINGLE step)."""
    import ctypes as C
    from presto_b200 import STEP_SINGLE, to_pages
    from util import rows_close, rows_of
    n = C.c_int32()
    gpu_backend.check(gpu_backend.device_count(C.byref(n)))
    if n.value < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_state_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    want = rows_of(to_pages(_q1_operator(oracle_backend, STEP_SINGLE), [_q1_pages(100, 40_000), _q1_pages(101, 50_000)]))
    got = pickle.load(open(tmp_path / "s_0.pkl", "rb"))
    assert pickle.load(open(tmp_path / "s_1.pkl", "rb")) == {}
    for shared in (False, True):
        rows_close(got[shared], want)
