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
        return None if not p or block.position_count == 0 else torch.as_tensor(_CudaArray(p, block.position_count, "|u1"), device=device)
    return None if block.is_null is None else torch.from_numpy(block.is_null)


def tensor_to_block(type_: int, tensor, nulls=None) -> Block:
    if tensor.is_cuda:
        return Block.device(type_, tensor.numel(), tensor.data_ptr(), nulls.data_ptr() if nulls is not None else 0, keepalive=(tensor, nulls))
    return Block.from_numpy(type_, tensor.numpy(), None if nulls is None else nulls.numpy())


def page_to_host(page: Page) -> Page:
    """Copy a device page of flat fixed-width blocks to host numpy blocks (tests and result hand-off)."""
    if not page.on_device:
        return page
    blocks = []
    for b in page.blocks:
        v = block_to_tensor(b).cpu().numpy()
        n = nulls_to_tensor(b)
        blocks.append(Block.from_numpy(b.type, v, None if n is None else n.cpu().numpy()))
    return Page(blocks, page.position_count)


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
    if on_device:
        # the collectives ran on torch's stream; the operators that consume the page run on their own streams
        torch.cuda.current_stream().synchronize()
    return Page(blocks, total)


def serialize_partitions(serde, page: Page, offsets: Sequence[int]) -> List[np.ndarray]:
    """Off-box leg of a repartition: one SerializedPage (wire bytes) per partition of a PartitionedOutputOperator output page
    (partition p = rows offsets[p]:offsets[p+1]) — what OptimizedPartitionedOutputOperator.flush hands to OutputBuffer.enqueue
    (MB/operator/repartition/OptimizedPartitionedOutputOperator.java:774-797).  One native call (pg_pages_serde_serialize_ranges): device
    pages are sliced by pointer arithmetic inside the library; the serde runs on the GPU."""
    return serde.serialize_ranges(page, offsets)


class FusedExchange:
    """Repartition and exchange in ONE kernel: pass 2 of PartitionedOutputOperator stores each row straight into the destination
    GPU's receive column over NVLink (peer memory mapped once through CUDA IPC), so the shuffle costs one scatter instead of a
    local scatter + an all-to-all copy.  What is left for torch.distributed is one all_gather of the P per-destination row counts
    (which also fixes every sender's row offset: sender-major, the order all_to_all_pages delivers) and one barrier.

    Replaces PartitionedOutputBuffer → HTTP → ExchangeClient (MB/execution/buffer/PartitionedOutputBuffer.java:174-181,
    MB/operator/ExchangeClient.java:71) for GPUs of one NVSwitch domain.  The page ``exchange`` returns lives in this object's
    receive columns and is valid until the next ``exchange`` call on ANY rank: consume it (or copy it) first."""

    def __init__(self, types: Sequence[int], capacity_rows: int, *, with_nulls: bool = False, group=None, backend=None):
        import torch
        import torch.distributed as dist
        from ._lib import load
        self._b = backend or load()
        self._group = group
        self.types = list(types)
        self.with_nulls = with_nulls
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self._dev = torch.device("cuda", torch.cuda.current_device())
        self._local, self._imported = [], []
        self.last_counts = None
        self._allocate(int(capacity_rows))

    def _allocate(self, capacity: int) -> None:
        """Collective: receive columns for `capacity` rows on every rank, exported and mapped into every peer."""
        import ctypes as C
        import torch
        import torch.distributed as dist
        self.capacity = capacity
        n_cols = len(self.types)
        self._local = []        # (values, nulls) device addresses of this rank's receive columns
        handles = np.zeros((n_cols, 2, abi.IPC_HANDLE_BYTES), dtype=np.uint8)
        for c, t in enumerate(self.types):
            ptrs = []
            for k, nbytes in enumerate((capacity * TYPE_WIDTH[t], capacity if self.with_nulls else 0)):
                if nbytes == 0:
                    ptrs.append(0)
                    continue
                p = C.c_void_p()
                self._b.check(self._b.device_alloc(max(nbytes, 2 << 20), C.byref(p)))      # >= one 2 MiB mapping granule: never sub-allocated, so the IPC handle maps exactly this column
                self._b.check(self._b.ipc_export(p, handles[c, k].ctypes.data))
                ptrs.append(p.value)
            self._local.append(tuple(ptrs))
        mine = torch.from_numpy(handles.reshape(-1)).to(self._dev)
        everyone = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(everyone, mine, group=self._group)
        # dst[p][c]: where partition p's rows of channel c go
        self._values = np.zeros((self.world, n_cols), dtype=np.uint64)
        self._nulls = np.zeros((self.world, n_cols), dtype=np.uint64) if self.with_nulls else None
        self._imported = []
        for r in range(self.world):
            h = everyone[r].cpu().numpy().reshape(n_cols, 2, abi.IPC_HANDLE_BYTES)
            for c in range(n_cols):
                for k in range(2 if self.with_nulls else 1):
                    if r == self.rank:
                        addr = self._local[c][k]
                    else:
                        p = C.c_void_p()
                        self._b.check(self._b.ipc_import(np.ascontiguousarray(h[c, k]).ctypes.data, C.byref(p)))
                        self._imported.append(p.value)
                        addr = p.value
                    (self._values if k == 0 else self._nulls)[r, c] = addr

    def _release(self) -> None:
        """Collective: unmap the peers' columns, then (after everyone unmapped) free this rank's."""
        import ctypes as C
        import torch.distributed as dist
        for p in self._imported:
            self._b.ipc_close(C.c_void_p(p))
        done = __import__("torch").zeros(1, dtype=__import__("torch").int32, device=self._dev)
        dist.all_reduce(done, group=self._group)
        done.item()
        for v, n in self._local:
            for p in (v, n):
                if p:
                    self._b.device_free(C.c_void_p(p))
        self._imported, self._local = [], []

    def exchange(self, op, page: Page) -> Page:
        """``op`` is a PartitionedOutputOperator with partition_count == world size over ``self.types``."""
        import os
        import time
        import torch
        import torch.distributed as dist
        assert op.partition_count == self.world, "one partition per rank"
        trace = os.environ.get("PG_Q3_TRACE")
        t = [time.perf_counter()] if trace else None
        counts = op.count(page)
        if trace: t.append(time.perf_counter())
        mine = torch.from_numpy(counts).to(self._dev)
        matrix = torch.empty(self.world * self.world, dtype=torch.int64, device=self._dev)
        # every rank has entered this call once the gather returns, so nobody still reads the previous received page
        dist.all_gather_into_tensor(matrix, mine, group=self._group)
        m = matrix.cpu().numpy().reshape(self.world, self.world)          # m[sender][destination]
        if trace: t.append(time.perf_counter())
        received = m[:, self.rank]
        total = int(received.sum())
        need = int(m.sum(axis=0).max())
        if need > self.capacity:          # every rank sees the same matrix, so every rank regrows together
            self._release()
            self._allocate(need + need // 4)
        row_offsets = m[:self.rank, :].sum(axis=0).astype(np.int64)       # rows lower-ranked senders put before mine, per destination
        op.scatter_to(self._values, self._nulls, row_offsets)            # returns after the scatter kernel (and its peer stores) completed
        if trace: t.append(time.perf_counter())
        done = torch.zeros(1, dtype=torch.int32, device=self._dev)
        dist.all_reduce(done, group=self._group)                          # every sender's rows have landed
        done.item()
        if trace:
            t.append(time.perf_counter())
            self.trace_ms = [round((b - a) * 1e3, 3) for a, b in zip(t[:-1], t[1:])]     # count, count all-gather, scatter, barrier
        self.last_counts = m
        blocks = [Block.device(t_, total, v, n if self._nulls is not None else 0, keepalive=self) for t_, (v, n) in zip(self.types, self._local)]
        return Page(blocks, total)

    def close(self):
        if self._b is not None:
            self._release()
            self._b = None


def _pack_page(page: Optional[Page]) -> np.ndarray:
    """A small host page of fixed-width columns → one uint8 buffer: [n_cols, n_rows, (type, has_nulls) per column, column data...]."""
    if page is None or page.position_count == 0:
        return np.zeros(0, dtype=np.uint8)
    n = page.position_count
    head = [len(page.blocks), n]
    parts = []
    for b in page.blocks:
        assert b.encoding == abi.ENC_FLAT and not b.on_device, "partial aggregation states are flat host columns"
        head += [b.type, 1 if b.is_null is not None else 0]
        if b.type == abi.VARCHAR:      # (group keys) offsets rebased to 0, then the byte count, then the bytes
            offs = np.ascontiguousarray(b.offsets[: n + 1], dtype=np.int32)
            lo, hi = int(offs[0]), int(offs[n])
            parts.append((offs - lo).astype(np.int32).view(np.uint8))
            parts.append(np.asarray([hi - lo], dtype=np.int64).view(np.uint8))
            parts.append(np.ascontiguousarray(b.values).view(np.uint8)[lo:hi])
        else:
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
        offsets = None
        if t == abi.VARCHAR:
            offsets = buf[off:off + 4 * (n + 1)].copy().view(np.int32); off += 4 * (n + 1)
            nbytes = int(buf[off:off + 8].copy().view(np.int64)[0]); off += 8
            vals = buf[off:off + nbytes].copy() if nbytes else np.zeros(1, dtype=np.uint8); off += nbytes
        else:
            nbytes = n * TYPE_WIDTH[t]
            vals = buf[off:off + nbytes].copy().view(_NP_DTYPE[t]); off += nbytes
        nulls = None
        if has_nulls:
            nulls = buf[off:off + n].copy(); off += n
        blocks.append(Block(t, n, vals, nulls, offsets))
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


def merge_aggregation_states(op, group=None, root: int = 0, max_groups: int = 4096) -> bool:
    """Device-side PARTIAL → FINAL hand-off of a (fused) hash aggregation across the GPUs of one node: every rank but `root` compacts
    its operator's group table into a fixed-capacity device buffer (pg_hash_aggregation_export_state: one kernel, no host round trip),
    ONE all_gather_into_tensor moves the buffers over NVLink, and `root` merges them into ITS operator's table
    (pg_hash_aggregation_merge_states: one small header read + one kernel per source).  The operators must come from the same
    configuration (the root's step decides what its output holds: STEP_SINGLE for final values).  Returns True on `root`, whose
    operator then finishes / outputs as usual; the other ranks' operators are simply closed.

    Replaces gather_pages_to_root for operators on GPUs: no page is built, serialised, or copied through the host."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = torch.device("cuda", torch.cuda.current_device())
    stride = (op.state_bytes(max_groups) + 255) // 256 * 256
    send = torch.empty(stride, dtype=torch.uint8, device=dev)
    recv = torch.empty(world * stride, dtype=torch.uint8, device=dev)
    shared_stream = op.cuda_stream != 0 and op.cuda_stream == torch.cuda.current_stream().cuda_stream
    if rank == root:
        send[:64].zero_()                 # an all-zero header: "no state in this slot"
    else:
        if not shared_stream:
            torch.cuda.current_stream().synchronize()      # `send` exists before the operator's stream writes it
        op.export_state(send.data_ptr(), stride, max_groups)
        if not shared_stream:
            torch.cuda.synchronize()
    dist.all_gather_into_tensor(recv, send, group=group)
    if rank != root:
        return False
    if not shared_stream:
        torch.cuda.current_stream().synchronize()          # the operator's own stream reads `recv`
    op.merge_states(recv.data_ptr(), world, stride, max_groups, order_base=1)
    op._keep_states = recv                # the merge kernels are asynchronous: keep the buffer until the operator is gone
    return True
