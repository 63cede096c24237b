"""Cross-GPU exchange: the one all-to-all that stands in for Presto's PartitionedOutput → HTTP → ExchangeOperator shuffle
inside an 8-GPU box.

The reference moves repartitioned rows as SerializedPages over HTTP long-poll (MB/execution/buffer/PartitionedOutputBuffer.java:174-181,
MB/operator/ExchangeClient.java:71); here `PartitionedOutputOperator` leaves the P partition slabs back to back in HBM and each column
is exchanged with one `all_to_all_single` (NCCL over NVLink 5 / NVSwitch; gloo on CPU for the tests).  torch.distributed is plumbing
only: no operator kernel lives here.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from . import _abi as abi
from .page import Block, Page, TYPE_WIDTH

_TORCH_DTYPE = None


def _torch_dtypes():
    global _TORCH_DTYPE
    if _TORCH_DTYPE is None:
        import torch
        _TORCH_DTYPE = {abi.BIGINT: torch.int64, abi.DOUBLE: torch.float64, abi.INTEGER: torch.int32, abi.DATE: torch.int32, abi.BOOLEAN: torch.int8}
    return _TORCH_DTYPE


class _CudaArray:
    """Zero-copy view of operator-owned device memory for torch (__cuda_array_interface__)."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


_TYPESTR = {abi.BIGINT: "<i8", abi.DOUBLE: "<f8", abi.INTEGER: "<i4", abi.DATE: "<i4", abi.BOOLEAN: "|i1"}


def block_to_tensor(block: Block, device=None):
    """A torch tensor over a block's values: zero-copy for device blocks (the operator keeps owning the memory)."""
    import torch
    if block.on_device:
        if block.position_count == 0:
            return torch.empty(0, dtype=_torch_dtypes()[block.type], device=device)
        return torch.as_tensor(_CudaArray(block.device_ptrs["values"], block.position_count, _TYPESTR[block.type]), device=device)
    return torch.from_numpy(block.values)


def nulls_to_tensor(block: Block, device=None):
    import torch
    if block.on_device:
        p = block.device_ptrs["is_null"]
        return None if not p else torch.as_tensor(_CudaArray(p, block.position_count, "|u1"), device=device)
    return None if block.is_null is None else torch.from_numpy(block.is_null)


def tensor_to_block(type_: int, tensor, nulls=None) -> Block:
    if tensor.is_cuda:
        return Block.device(type_, tensor.numel(), tensor.data_ptr(), nulls.data_ptr() if nulls is not None else 0, keepalive=(tensor, nulls))
    return Block.from_numpy(type_, tensor.numpy(), None if nulls is None else nulls.numpy())


def all_to_all_pages(page: Page, offsets: Sequence[int], types: Sequence[int], group=None) -> Page:
    """Send partition p of `page` (rows offsets[p]:offsets[p+1], the layout PartitionedOutputOperator.get_output returns) to
    rank p; returns the rows this rank received, sender-major (rank 0's rows first) — the order an ExchangeOperator draining
    per-source buffers in order would see.  One count exchange + one all_to_all_single per column."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    assert len(offsets) == world + 1, "one partition per rank"
    send_counts = [int(offsets[p + 1] - offsets[p]) for p in range(world)]
    on_device = page.on_device
    dev = torch.device("cuda", torch.cuda.current_device()) if on_device else torch.device("cpu")
    sc = torch.tensor(send_counts, dtype=torch.int64, device=dev)
    rc = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_to_all_single(rc, sc, group=group)
    recv_counts = [int(x) for x in rc.tolist()]
    total = sum(recv_counts)
    # null maps travel only for columns where SOME rank has one: one small all_reduce decides for all columns at once
    srcs = [block_to_tensor(b, dev) for b in page.blocks]
    src_nulls = [nulls_to_tensor(b, dev) for b in page.blocks]
    has = torch.tensor([1 if x is not None else 0 for x in src_nulls], dtype=torch.int64, device=dev)
    dist.all_reduce(has, op=dist.ReduceOp.MAX, group=group)
    has = [bool(x) for x in has.tolist()]
    blocks: List[Block] = []
    for t, src, nulls, with_nulls in zip(types, srcs, src_nulls, has):
        dst = torch.empty(total, dtype=src.dtype, device=dev)
        dist.all_to_all_single(dst, src, output_split_sizes=recv_counts, input_split_sizes=send_counts, group=group)
        dst_nulls = None
        if with_nulls:
            if nulls is None:
                nulls = torch.zeros(src.numel(), dtype=torch.uint8, device=dev)
            dst_nulls = torch.empty(total, dtype=torch.uint8, device=dev)
            dist.all_to_all_single(dst_nulls, nulls, output_split_sizes=recv_counts, input_split_sizes=send_counts, group=group)
        blocks.append(tensor_to_block(t, dst, dst_nulls))
    return Page(blocks, total)


def _pack_page(page: Optional[Page]) -> np.ndarray:
    """A small host page of fixed-width columns → one uint8 buffer: [n_cols, n_rows, (type, has_nulls) per column, column data...]."""
    if page is None or page.position_count == 0:
        return np.zeros(0, dtype=np.uint8)
    n = page.position_count
    head = [len(page.blocks), n]
    parts = []
    for b in page.blocks:
        assert b.type != abi.VARCHAR and b.encoding == abi.ENC_FLAT and not b.on_device, "partial aggregation states are flat fixed-width host columns"
        head += [b.type, 1 if b.is_null is not None else 0]
        parts.append(np.ascontiguousarray(b.values).view(np.uint8)[: n * TYPE_WIDTH[b.type]])
        if b.is_null is not None:
            parts.append(np.ascontiguousarray(b.is_null, dtype=np.uint8)[:n])
    return np.concatenate([np.asarray(head, dtype=np.int64).view(np.uint8)] + parts)


def _unpack_page(buf: np.ndarray) -> Optional[Page]:
    if buf.size == 0:
        return None
    n_cols, n = [int(x) for x in buf[:16].view(np.int64)]
    meta = buf[16:16 + 16 * n_cols].view(np.int64)
    off = 16 + 16 * n_cols
    blocks = []
    from .page import _NP_DTYPE
    for c in range(n_cols):
        t, has_nulls = int(meta[2 * c]), int(meta[2 * c + 1])
        nbytes = n * TYPE_WIDTH[t]
        vals = buf[off:off + nbytes].copy().view(_NP_DTYPE[t]); off += nbytes
        nulls = None
        if has_nulls:
            nulls = buf[off:off + n].copy(); off += n
        blocks.append(Block(t, n, vals, nulls))
    return Page(blocks, n)


def gather_pages_to_root(page: Optional[Page], group=None, root: int = 0) -> List[Page]:
    """Small host pages (partial aggregation states) from every rank to `root` — the PARTIAL → FINAL hand-off of
    HashAggregationOperator across GPUs (Q1: a few hundred bytes per rank).  Two fixed-shape collectives (sizes, padded bytes)
    instead of a pickled object gather."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    buf = _pack_page(page)
    size = torch.tensor([buf.size], dtype=torch.int64, device=dev)
    sizes = [torch.empty(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, size, group=group)
    sizes = [int(x.item()) for x in sizes]
    cap = max(max(sizes), 1)
    send = torch.zeros(cap, dtype=torch.uint8, device=dev)
    if buf.size:
        send[:buf.size] = torch.from_numpy(buf).to(dev)
    recv = [torch.empty(cap, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(recv, send, group=group)
    if dist.get_rank(group) != root:
        return []
    pages = []
    for r in range(world):
        p = _unpack_page(recv[r][:sizes[r]].cpu().numpy())
        if p is not None:
            pages.append(p)
    return pages
