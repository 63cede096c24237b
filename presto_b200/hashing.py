"""Vectorised (numpy) restatement of the reference's hash functions for callers that precompute ``$hashvalue`` channels
the way HashGenerationOptimizer.java:975-987 does (TypeUtils.getHashBlock, MB/type/TypeUtils.java:149-172).

  type hash  : AbstractLongType.hash        PC/type/AbstractLongType.java:137-141 (BIGINT / INTEGER / DATE)
               DoubleType.hash              PC/type/DoubleType.java:92-100 (doubleToLongBits, canonical NaN)
               BooleanType.hash             PC/type/BooleanType.java:121-125
               XxHash64.hash(slice)         PC/block/AbstractVariableWidthBlock.java:102-106 (via the C library)
  combine    : CombineHashFunction.getHash  MB/operator/scalar/CombineHashFunction.java:26-31   (31*h + v, NULL → 0)
  partition  : HashGenerator.getPartition   MB/operator/HashGenerator.java:22-29
"""
from __future__ import annotations

from typing import Sequence

import numpy as np

from . import _abi as abi
from .page import Block

_C1 = np.uint64(0xC2B2AE3D27D4EB4F)
_C2 = np.uint64(0x9E3779B185EBCA87)


def hash_long(values: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        v = values.astype(np.int64).view(np.uint64) * _C1
        v = (v << np.uint64(31)) | (v >> np.uint64(33))
        return (v * _C2).view(np.int64)


def murmur_hash3(values: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = values.astype(np.int64).view(np.uint64).copy()
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xFF51AFD7ED558CCD)
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xC4CEB9FE1A85EC53)
        x ^= x >> np.uint64(33)
        return x.view(np.int64)


def type_hash(block: Block, backend=None) -> np.ndarray:
    """TypeUtils.hashPosition for every position of a flat host block (NULL → 0)."""
    assert block.encoding == abi.ENC_FLAT and not block.on_device
    n = block.position_count
    t = block.type
    if t in (abi.BIGINT, abi.INTEGER, abi.DATE):
        h = hash_long(block.values.astype(np.int64))
    elif t == abi.DOUBLE:
        d = block.values.astype(np.float64)
        bits = d.view(np.int64).copy()
        bits[np.isnan(d)] = 0x7FF8000000000000
        bits[d == 0] = 0          # TypeUtils.doubleHashCode: -0.0 hashes like +0.0
        h = hash_long(bits)
    elif t == abi.BOOLEAN:
        h = np.where(block.values != 0, 1231, 1237).astype(np.int64)
    elif t == abi.VARCHAR:
        if backend is None:
            from ._lib import load_library
            backend = load_library()
        h = np.empty(n, dtype=np.int64)
        base = block.values.ctypes.data
        for i in range(n):
            h[i] = backend.hash_xxhash64(base + int(block.offsets[i]), int(block.offsets[i + 1] - block.offsets[i]))
    else:
        raise ValueError(f"unsupported type {t}")
    if block.is_null is not None:
        h = np.where(block.is_null != 0, 0, h)
    return h


def get_hash_block(blocks: Sequence[Block], backend=None) -> Block:
    """TypeUtils.getHashBlock: h = 0; for each block: h = 31*h + typeHash."""
    n = blocks[0].position_count
    h = np.zeros(n, dtype=np.int64)
    with np.errstate(over="ignore"):
        for b in blocks:
            h = (h.view(np.uint64) * np.uint64(31) + type_hash(b, backend).view(np.uint64)).view(np.int64)
    return Block.from_numpy(abi.BIGINT, h)


def get_partition(raw_hash: np.ndarray, partition_count: int) -> np.ndarray:
    u = raw_hash.astype(np.int64).view(np.uint64)
    folded = (u ^ (u >> np.uint64(32))) & np.uint64(0xFFFFFFFF)
    return ((folded * np.uint64(partition_count)) >> np.uint64(32)).astype(np.int32)


def local_partition(raw_hash: np.ndarray, partition_count: int) -> np.ndarray:
    """LocalPartitionGenerator.getPartition  MB/operator/exchange/LocalPartitionGenerator.java:43-63:
    (int) XxHash64.hash(Long.reverse(rawHash)) & (partitionCount - 1), numpy restatement (XXH64 seed 0 of the 8 LE bytes)."""
    assert partition_count > 0 and partition_count & (partition_count - 1) == 0, "partitionCount must be a power of 2"
    M = np.uint64(0xFFFFFFFFFFFFFFFF)
    x = raw_hash.astype(np.int64).view(np.uint64).copy()
    r = np.zeros_like(x)
    for _ in range(64):                       # Long.reverse
        r = (r << np.uint64(1)) | (x & np.uint64(1))
        x = x >> np.uint64(1)
    P1, P2, P3, P4, P5 = (np.uint64(v) for v in (0x9E3779B185EBCA87, 0xC2B2AE3D27D4EB4F, 0x165667B19E3779F9, 0x85EBCA77C2B2AE63, 0x27D4EB2F165667C5))

    def rotl(v, k):
        return ((v << np.uint64(k)) | (v >> np.uint64(64 - k))) & M
    with np.errstate(over="ignore"):
        h = P5 + np.uint64(8)
        k = rotl(r * P2, 31) * P1
        h = rotl(h ^ k, 27) * P1 + P4
        h ^= h >> np.uint64(33); h = h * P2
        h ^= h >> np.uint64(29); h = h * P3
        h ^= h >> np.uint64(32)
    return (h & np.uint64(partition_count - 1)).astype(np.int32)
