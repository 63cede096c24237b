"""Seeded synthetic TPC-H-shaped columns, the same values on the host (numpy) and in HBM (torch, used only
as the device allocator / elementwise generator for benchmark INPUT data — never on the operator path).

  splitmix64(seed, i) = mix(seed + (i + 1) * 0x9E3779B97F4A7C15)
  group-by key   = (splitmix64(42, i) >> 11) % G                       (config 2: bigint key, G groups)
  double column c= (splitmix64(43 + c, i) >> 11) * 2^-53 * 1e5          (config 2: 4 double aggregates)
"""
from __future__ import annotations

import numpy as np

_GOLD = 0x9E3779B97F4A7C15
_M1 = 0xBF58476D1CE4E5B9
_M2 = 0x94D049BB133111EB


def _s64(x: int) -> int:
    x &= (1 << 64) - 1
    return x - (1 << 64) if x >= (1 << 63) else x


def splitmix64_np(seed: int, start: int, n: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        i = np.arange(start + 1, start + 1 + n, dtype=np.uint64)
        z = np.uint64(seed) + i * np.uint64(_GOLD)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(_M1)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(_M2)
        return z ^ (z >> np.uint64(31))


def group_key_np(start: int, n: int, groups: int, seed: int = 42) -> np.ndarray:
    return ((splitmix64_np(seed, start, n) >> np.uint64(11)) % np.uint64(groups)).astype(np.int64)


def double_column_np(column: int, start: int, n: int, seed: int = 43) -> np.ndarray:
    return (splitmix64_np(seed + column, start, n) >> np.uint64(11)).astype(np.float64) * (2.0 ** -53 * 1e5)


# ---- torch twins (device) -------------------------------------------------------------------------------------------------
def _lsr(t, k):
    return (t >> k) & ((1 << (64 - k)) - 1)


def splitmix64_torch(seed: int, start: int, n: int, device):
    import torch
    i = torch.arange(start + 1, start + 1 + n, dtype=torch.int64, device=device)
    z = i * _s64(_GOLD) + _s64(seed)
    z = (z ^ _lsr(z, 30)) * _s64(_M1)
    z = (z ^ _lsr(z, 27)) * _s64(_M2)
    return z ^ _lsr(z, 31)


def fill_group_key_torch(out, start: int, groups: int, seed: int = 42, chunk: int = 1 << 25) -> None:
    n = out.numel()
    for b in range(0, n, chunk):
        m = min(chunk, n - b)
        out[b:b + m] = _lsr(splitmix64_torch(seed, start + b, m, out.device), 11) % groups


def fill_double_column_torch(out, column: int, start: int, seed: int = 43, chunk: int = 1 << 25) -> None:
    n = out.numel()
    for b in range(0, n, chunk):
        m = min(chunk, n - b)
        out[b:b + m] = _lsr(splitmix64_torch(seed + column, start + b, m, out.device), 11).to(out.dtype) * (2.0 ** -53 * 1e5)


# ---- join (config 3): orders ⋈ lineitem ------------------------------------------------------------------------------------
def sparse_order_key_np(i: np.ndarray) -> np.ndarray:
    """dbgen mk_sparse: 8 of every 32 integers are order keys (velox/tpch/gen/dbgen/build.cpp:144-156)."""
    i = i.astype(np.int64)
    return ((i >> 3) << 5) + (i & 7)


# ---- Q3 (config 4): customer / orders / lineitem -----------------------------------------------------------------------------
def q3_tables_np(n_orders: int, seed: int = 7):
    """TPC-H-shaped Q3 inputs (numpy, for the parity tests): customer(c_custkey, c_mktsegment code 0..4), orders(o_orderkey sparse,
    o_custkey, o_orderdate U[8035, 10440], o_shippriority 0) in shuffled order, lineitem clustered by orderkey with 1..7 rows per
    order (l_orderkey, l_extendedprice, l_discount, l_shipdate = o_orderdate + U[1, 121])."""
    rng = np.random.default_rng(seed)
    n_cust = max(n_orders // 10, 10)
    c_custkey = np.arange(1, n_cust + 1, dtype=np.int64)
    c_seg = rng.integers(0, 5, n_cust).astype(np.int32)
    i = np.arange(1, n_orders + 1, dtype=np.int64)
    okey = sparse_order_key_np(i)
    odate = rng.integers(8035, 10441, n_orders).astype(np.int32)
    ocust = rng.integers(1, n_cust + 1, n_orders).astype(np.int64)
    lines = rng.integers(1, 8, n_orders)
    idx = np.repeat(np.arange(n_orders), lines)
    n_li = len(idx)
    lineitem = (okey[idx], np.round(rng.random(n_li) * 1e5, 2), rng.integers(0, 11, n_li) / 100.0,
                (odate[idx] + rng.integers(1, 122, n_li)).astype(np.int32))
    perm = rng.permutation(n_orders)
    orders = (okey[perm], ocust[perm], odate[perm], np.zeros(n_orders, dtype=np.int32))
    return (c_custkey, c_seg), orders, lineitem


MKTSEGMENTS = [b"AUTOMOBILE", b"BUILDING", b"FURNITURE", b"HOUSEHOLD", b"MACHINERY"]     # TPC-H c_mktsegment's domain (dbgen dists.dss "msegmnt")


def mktsegment_varchar_np(codes: np.ndarray):
    """c_mktsegment as the VARCHAR column the reference scans (bytes, int32 offsets) from the 0..4 codes of q3_tables_np."""
    lens = np.array([len(s) for s in MKTSEGMENTS], dtype=np.int64)[codes]
    offsets = np.zeros(len(codes) + 1, dtype=np.int32)
    np.cumsum(lens, out=offsets[1:])
    table = np.zeros((5, 10), dtype=np.uint8)
    for i, s in enumerate(MKTSEGMENTS):
        table[i, :len(s)] = np.frombuffer(s, dtype=np.uint8)
    data = table[codes][np.arange(10)[None, :] < lens[:, None]]
    return np.ascontiguousarray(data), offsets


def mktsegment_varchar_torch(codes):
    """The same on the device (benchmark input only): (bytes uint8, offsets int32)."""
    import torch
    dev = codes.device
    lens = torch.tensor([len(s) for s in MKTSEGMENTS], dtype=torch.int64, device=dev)[codes.long()]
    offsets = torch.zeros(codes.numel() + 1, dtype=torch.int32, device=dev)
    offsets[1:] = torch.cumsum(lens, 0).to(torch.int32)
    table = torch.zeros((5, 10), dtype=torch.uint8, device=dev)
    for i, s in enumerate(MKTSEGMENTS):
        table[i, :len(s)] = torch.tensor(list(s), dtype=torch.uint8, device=dev)
    data = table[codes.long()][torch.arange(10, device=dev)[None, :] < lens[:, None]]
    return data.contiguous(), offsets


def q3_tables_torch(n_orders: int, device, seed: int = 7):
    """The same shape generated in HBM with torch (benchmark input only)."""
    import torch
    g = torch.Generator(device=device); g.manual_seed(seed)
    n_cust = max(n_orders // 10, 10)
    c_custkey = torch.arange(1, n_cust + 1, dtype=torch.int64, device=device)
    c_seg = torch.randint(0, 5, (n_cust,), dtype=torch.int32, device=device, generator=g)
    i = torch.arange(1, n_orders + 1, dtype=torch.int64, device=device)
    okey = ((i >> 3) << 5) + (i & 7)
    odate = torch.randint(8035, 10441, (n_orders,), dtype=torch.int32, device=device, generator=g)
    ocust = torch.randint(1, n_cust + 1, (n_orders,), dtype=torch.int64, device=device, generator=g)
    lines = torch.randint(1, 8, (n_orders,), dtype=torch.int64, device=device, generator=g)
    idx = torch.repeat_interleave(torch.arange(n_orders, device=device), lines)
    n_li = idx.numel()
    l_orderkey = okey[idx]
    l_ship = (odate[idx] + torch.randint(1, 122, (n_li,), dtype=torch.int32, device=device, generator=g)).to(torch.int32)
    del idx, lines
    l_ep = torch.rand(n_li, dtype=torch.float64, device=device, generator=g) * 1e5
    l_disc = torch.randint(0, 11, (n_li,), device=device, generator=g).to(torch.float64) / 100
    perm = torch.randperm(n_orders, device=device, generator=g)
    orders = (okey[perm].contiguous(), ocust[perm].contiguous(), odate[perm].contiguous(), torch.zeros(n_orders, dtype=torch.int32, device=device))
    return (c_custkey, c_seg), orders, (l_orderkey, l_ep, l_disc, l_ship)
