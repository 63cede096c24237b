"""-m gpu: BASELINE-sized shapes checked through size-independent properties (the oracle would take minutes at these sizes):
row counts that must add up, checksums against plain torch reductions of the same device-resident columns, first-appearance order,
stable order inside partitions.  These are the inputs on which the large-input code paths switch on by themselves — direct-mode
aggregation, the range-partitioned build of the dense join index (>= 4 M build rows), multi-unit scans — so they are exercised here
without the environment overrides the small parity tests use."""
import numpy as np
import pytest

from presto_b200 import (
    AGG_COUNT_STAR, AGG_SUM_DOUBLE, BIGINT, DATE, DOUBLE, INTEGER, JOIN_INNER, MEM_DEVICE, Aggregate, Block, FilterAndProjectOperator,
    HashAggregationOperator, HashBuilderOperator, JoinBridge, LookupJoinOperator, Page, PartitionedOutputOperator, Program,
)
from presto_b200.hashing import get_partition, hash_long

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def device_page(tensors, types):
    torch.cuda.synchronize()          # the operator runs on its own stream: the generated columns must be complete
    n = tensors[0].numel()
    return Page([Block.device(t, n, x.data_ptr(), keepalive=x) for t, x in zip(types, tensors)])


def column(page, i):
    from presto_b200.exchange import block_to_tensor
    return block_to_tensor(page.blocks[i])


@pytest.mark.parametrize("groups", [4, 60, 250, 1000, 100_000, 2_000_000],
                         ids=["tier0", "warp_private_sets", "warp_private_sets_full", "shared_table", "direct_one_pass", "direct_slot_range_passes"])
def test_group_by_sums_counts_and_first_appearance_order(gpu_backend, groups):
    n = 1 << 25
    g = torch.Generator(device="cuda"); g.manual_seed(groups)
    keys = torch.randint(0, groups, (n,), device="cuda", generator=g) * 7919 - 5          # sparse, negative values included
    v0 = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 1e5
    v1 = torch.randint(0, 11, (n,), device="cuda", generator=g).to(torch.float64) / 100
    op = HashAggregationOperator([BIGINT], [0], [Aggregate(AGG_SUM_DOUBLE, 1), Aggregate(AGG_SUM_DOUBLE, 2), Aggregate(AGG_COUNT_STAR)],
                                 expected_groups=groups, backend=gpu_backend, output_memory=MEM_DEVICE, max_output_rows=1 << 30)
    op.add_input(device_page([keys, v0, v1], [BIGINT, DOUBLE, DOUBLE]))
    op.finish()
    out = op.get_output(copy=False)
    got_keys, s0, s1, cnt = [column(out, i).clone() for i in range(4)]
    op.close()
    uniq, inverse = torch.unique(keys, return_inverse=True)
    assert got_keys.numel() == uniq.numel()
    assert int(cnt.sum()) == n                                              # every row counted exactly once
    # per-group values against torch (float64 scatter-add; both sides sum in a data-dependent order: 1e-9 relative is ~1000 ulps)
    order = torch.argsort(got_keys)
    assert torch.equal(got_keys[order], uniq)
    want0 = torch.zeros(uniq.numel(), dtype=torch.float64, device="cuda").index_add_(0, inverse, v0)
    want1 = torch.zeros(uniq.numel(), dtype=torch.float64, device="cuda").index_add_(0, inverse, v1)
    wantc = torch.bincount(inverse, minlength=uniq.numel())
    assert torch.equal(cnt[order], wantc)
    assert torch.allclose(s0[order], want0, rtol=1e-9, atol=0) and torch.allclose(s1[order], want1, rtol=1e-9, atol=0)
    # output order = order of first appearance (BigintGroupByHash.java:255-262): the first row of group i precedes that of group i+1
    first = torch.full((uniq.numel(),), n, dtype=torch.int64, device="cuda").scatter_reduce_(0, inverse, torch.arange(n, device="cuda"), "amin")
    first_in_output_order = first[torch.searchsorted(uniq, got_keys)]
    assert bool((first_in_output_order[1:] > first_in_output_order[:-1]).all())


def test_join_with_the_partitioned_dense_build(gpu_backend):
    """orders x lineitem shape: 6 M unique sparse build keys shuffled (above the 4 M threshold of the range-partitioned build),
    24 M probe rows of which ~1/8 miss; every output row must carry the payload of its key, in probe order."""
    nb, npr = 6_000_000, 24_000_000
    g = torch.Generator(device="cuda"); g.manual_seed(11)
    i = torch.arange(1, nb + 1, device="cuda")
    key_of = ((i >> 3) << 5) + (i & 7)                                    # TPC-H order keys: 8 of every 32 integers
    perm = torch.randperm(nb, device="cuda", generator=g)
    okey = key_of[perm].contiguous()
    odate = torch.randint(8035, 10441, (nb,), dtype=torch.int32, device="cuda", generator=g)
    oprio = (okey % 5).to(torch.int32)
    pick = torch.randint(0, nb, (npr,), device="cuda", generator=g)
    lkey = torch.where(torch.rand(npr, device="cuda", generator=g) < 0.125, key_of[pick] + 8, key_of[pick]).contiguous()   # +8: never a key
    lprice = torch.rand(npr, dtype=torch.float64, device="cuda", generator=g)
    bridge = JoinBridge(backend=gpu_backend)
    builder = HashBuilderOperator(bridge, [BIGINT, DATE, INTEGER], [0], [1, 2], expected_positions=nb, backend=gpu_backend)
    builder.add_input(device_page([okey, odate, oprio], [BIGINT, DATE, INTEGER]))
    builder.finish()
    probe = LookupJoinOperator(bridge, [BIGINT, DOUBLE], [0], [0, 1], join_type=JOIN_INNER, backend=gpu_backend, output_memory=MEM_DEVICE)
    probe.add_input(device_page([lkey, lprice], [BIGINT, DOUBLE]))
    out = probe.get_output(copy=False)
    got = [column(out, c).clone() for c in range(4)]
    probe.close(); builder.close(); bridge.close()
    # reference with torch: position of each probe key among the sorted build keys
    sorted_keys, order = torch.sort(okey)
    pos = torch.searchsorted(sorted_keys, lkey).clamp(max=nb - 1)
    hit = sorted_keys[pos] == lkey
    assert got[0].numel() == int(hit.sum())
    assert torch.equal(got[0], lkey[hit]) and torch.equal(got[1], lprice[hit])          # probe columns, probe order
    assert torch.equal(got[2], odate[order][pos][hit]) and torch.equal(got[3], oprio[order][pos][hit])


def test_partitioned_output_counts_checksums_and_stable_order(gpu_backend):
    n, parts = 1 << 25, 8
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    key = torch.randint(-2**62, 2**62, (n,), device="cuda", generator=g)
    rowid = torch.arange(n, device="cuda")
    val = torch.randint(0, 1 << 30, (n,), dtype=torch.int32, device="cuda", generator=g)
    op = PartitionedOutputOperator([BIGINT, BIGINT, INTEGER], parts, [0], backend=gpu_backend, output_memory=MEM_DEVICE)
    op.add_input(device_page([key, rowid, val], [BIGINT, BIGINT, INTEGER]))
    out, offs = op.get_output(copy=False), op.get_offsets()
    k, r, v = [column(out, c).clone() for c in range(3)]
    op.close()
    assert offs[0] == 0 and offs[-1] == n
    # the partition function on a sample, with the numpy restatement of HashGenerator.getPartition / AbstractLongType.hash
    sample = torch.randint(0, n, (200_000,), device="cuda", generator=g)
    want_part = get_partition(hash_long(key[sample].cpu().numpy()), parts)
    bounds = torch.as_tensor(np.asarray(offs), device="cuda")
    for p in range(parts):
        lo, hi = int(offs[p]), int(offs[p + 1])
        rows = r[lo:hi]
        assert bool((rows[1:] > rows[:-1]).all())                       # input order kept inside a partition
        assert torch.equal(k[lo:hi], key[rows]) and torch.equal(v[lo:hi], val[rows])   # rows travel whole
    where = torch.empty(n, dtype=torch.int64, device="cuda"); where[r] = torch.arange(n, device="cuda")     # output position of every input row
    got_part = (torch.searchsorted(bounds, where[sample], right=True) - 1).cpu().numpy()
    assert (got_part == want_part).all()
    assert int(r.sum()) == n * (n - 1) // 2                                # a permutation: nothing lost, nothing duplicated


def test_filter_project_selection_and_projection(gpu_backend):
    n = 1 << 25
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    ep = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 1e5
    disc = torch.randint(0, 11, (n,), device="cuda", generator=g).to(torch.float64) / 100
    ship = torch.randint(8035, 10600, (n,), dtype=torch.int32, device="cuda", generator=g)
    key = torch.arange(n, device="cuda")
    p = Program([BIGINT, DOUBLE, DOUBLE, DATE])
    k, e, d, s = [p.input(i) for i in range(4)]
    p.set_filter(s > 9204)
    p.project(k, e * (1.0 - d))
    op = FilterAndProjectOperator([BIGINT, DOUBLE, DOUBLE, DATE], p, backend=gpu_backend, output_memory=MEM_DEVICE)
    op.add_input(device_page([key, ep, disc, ship], [BIGINT, DOUBLE, DOUBLE, DATE]))
    out = op.get_output(copy=False)
    got_k, got_rev = column(out, 0).clone(), column(out, 1).clone()
    op.close()
    sel = ship > 9204
    assert torch.equal(got_k, key[sel])                                     # exactly the selected rows, in order
    assert torch.equal(got_rev, (ep * (1.0 - disc))[sel])                   # IEEE double arithmetic, bit for bit (no FMA contraction)
