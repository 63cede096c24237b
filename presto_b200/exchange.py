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
    blocks: List[Block] = []
    for t, b in zip(types, page.blocks):
        src = block_to_tensor(b, dev)
        dst = torch.empty(total, dtype=src.dtype, device=dev)
        dist.all_to_all_single(dst, src, output_split_sizes=recv_counts, input_split_sizes=send_counts, group=group)
        src_nulls = nulls_to_tensor(b, dev)
        # null maps travel only when some rank has one for this column
        has = torch.tensor([1 if src_nulls is not None else 0], dtype=torch.int64, device=dev)
        dist.all_reduce(has, op=dist.ReduceOp.MAX, group=group)
        dst_nulls = None
        if int(has.item()):
            if src_nulls is None:
                src_nulls = torch.zeros(src.numel(), dtype=torch.uint8, device=dev)
            dst_nulls = torch.empty(total, dtype=torch.uint8, device=dev)
            dist.all_to_all_single(dst_nulls, src_nulls, output_split_sizes=recv_counts, input_split_sizes=send_counts, group=group)
        blocks.append(tensor_to_block(t, dst, dst_nulls))
    return Page(blocks, total)


def gather_pages_to_root(page: Optional[Page], group=None, root: int = 0) -> List[Page]:
    """Small host pages (partial aggregation states) from every rank to `root` — the PARTIAL → FINAL hand-off of
    HashAggregationOperator across GPUs (Q1: a few hundred bytes per rank)."""
    import torch.distributed as dist
    payload = None
    if page is not None and page.position_count > 0:
        payload = [(b.type, b.values, b.is_null, b.offsets) for b in page.blocks]
    world = dist.get_world_size(group)
    gathered = [None] * world
    dist.all_gather_object(gathered, payload, group=group)
    if dist.get_rank(group) != root:
        return []
    pages = []
    for item in gathered:
        if item is None:
            continue
        blocks = [Block(t, len(v) if off is None else len(off) - 1, v, nu, off) for (t, v, nu, off) in item]
        pages.append(Page(blocks))
    return pages
