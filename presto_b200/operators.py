"""Host-side mirror of the reference's operator interfaces for the hot path, over the C ABI.

Same names, argument meaning and error behaviour as presto-main-base/src/main/java/com/facebook/presto/operator:
Operator.java:20-102 (needsInput / addInput / getOutput / finish / isFinished / close), HashAggregationOperator,
HashBuilderOperator + LookupJoinOperator (joined by a JoinBridge), PartitionedOutputOperator, FilterAndProjectOperator,
GroupByHash.  Every class takes the ``Backend`` to run on; the default is the CUDA library (no CPU fallback).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _abi as abi
from .page import Block, Page


def _default_backend():
    from ._lib import load
    return load()


def _options(device: int = 0, output_memory: int = abi.MEM_HOST, cuda_stream: int = 0, max_output_rows: int = 0) -> abi.pg_operator_options:
    o = abi.pg_operator_options()
    o.device = device
    o.output_memory = output_memory
    o.cuda_stream = cuda_stream or None
    o.max_output_rows = max_output_rows
    return o


class Operator:
    """MB/operator/Operator.java:20-102."""

    def __init__(self, backend, handle: C.c_void_p):
        self._b = backend
        self._h = handle
        self._closed = False

    def needs_input(self) -> bool:
        out = C.c_int32()
        self._b.check(self._b.operator_needs_input(self._h, C.byref(out)))
        return bool(out.value)

    def is_blocked(self) -> bool:
        """Operator.isBlocked() (Operator.java:32): True while the operator waits for another Driver (a LookupSource still being built)."""
        out = C.c_int32()
        self._b.check(self._b.operator_is_blocked(self._h, C.byref(out)))
        return bool(out.value)

    def add_input(self, page: Page) -> None:
        cpage, keep = page.as_c()
        self._b.check(self._b.operator_add_input(self._h, C.byref(cpage)))
        del keep

    def get_output(self, copy: bool = True) -> Optional[Page]:
        out = abi.pg_page()
        self._b.check(self._b.operator_get_output(self._h, C.byref(out)))
        return Page.from_c(out, self._b, copy=copy)

    def finish(self) -> None:
        self._b.check(self._b.operator_finish(self._h))

    def is_finished(self) -> bool:
        out = C.c_int32()
        self._b.check(self._b.operator_is_finished(self._h, C.byref(out)))
        return bool(out.value)

    def get_stats(self) -> dict:
        s = abi.pg_operator_stats()
        self._b.check(self._b.operator_get_stats(self._h, C.byref(s)))
        return {name: getattr(s, name) for name, _ in abi.pg_operator_stats._fields_}

    def close(self) -> None:
        if not self._closed:
            self._closed = True
            self._b.operator_destroy(self._h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def to_pages(operator: Operator, input_pages: Sequence[Page]) -> List[Page]:
    """OperatorAssertion.toPages (presto-main-base/src/test/.../operator/OperatorAssertion.java:78-130): emulate the Driver
    loop — while needsInput: addInput, drain getOutput; then finish and drain until isFinished."""
    out: List[Page] = []
    it = iter(input_pages)
    pending = next(it, None)
    for _ in range(10_000_000):
        if pending is None:
            break
        if operator.needs_input():
            operator.add_input(pending)
            pending = next(it, None)
        page = operator.get_output()
        if page is not None and page.position_count > 0:
            out.append(page)
    operator.finish()
    for _ in range(10_000_000):
        if operator.is_finished():
            break
        page = operator.get_output()
        if page is not None and page.position_count > 0:
            out.append(page)
    assert operator.is_finished(), "Operator did not finish"
    assert not operator.needs_input(), "Operator still wants input after finish"
    return out


class Aggregate:
    def __init__(self, function: int, input_channel: int = -1, mask_channel: int = -1):
        self.function, self.input_channel, self.mask_channel = function, input_channel, mask_channel


class _AggregationState:
    """PARTIAL -> FINAL hand-off between GPUs without intermediate pages (include/presto_gpu.h pg_hash_aggregation_*_state): the
    operator's group table compacted into a device buffer, merged by an identically configured operator on another GPU."""

    def state_bytes(self, max_groups: int) -> int:
        out = C.c_size_t()
        self._b.check(self._b.hash_aggregation_state_bytes(self._h, max_groups, C.byref(out)))
        return int(out.value)

    def export_state(self, device_pointer: int, capacity: int, max_groups: int) -> None:
        """Asynchronous on the operator's stream."""
        self._b.check(self._b.hash_aggregation_export_state(self._h, C.c_void_p(device_pointer), capacity, max_groups))

    def merge_states(self, device_pointer: int, count: int, stride: int, max_groups: int, order_base: int = 1) -> None:
        self._b.check(self._b.hash_aggregation_merge_states(self._h, C.c_void_p(device_pointer), count, stride, max_groups, order_base))


class HashAggregationOperator(_AggregationState, Operator):
    """HashAggregationOperator.java:56-725 (in-memory builder)."""

    FACTORY_KIND = 1

    @staticmethod
    def _config(group_types: Sequence[int], group_channels: Sequence[int], aggregates: Sequence[Aggregate], *,
                hash_channel: int = -1, step: int = abi.STEP_SINGLE, expected_groups: int = 10_000,
                group_max_lengths: Optional[Sequence[int]] = None, **options):
        cfg = abi.pg_hash_aggregation_config()
        gc, gt = abi.i32_array(group_channels), abi.i32_array(group_types)
        gl = abi.i32_array(group_max_lengths if group_max_lengths is not None else [0] * len(group_types))
        aggs = (abi.pg_aggregate * max(1, len(aggregates)))()
        for i, a in enumerate(aggregates):
            aggs[i].function, aggs[i].input_channel, aggs[i].mask_channel = a.function, a.input_channel, a.mask_channel
        cfg.n_group_channels = len(group_channels)
        cfg.group_channels, cfg.group_types, cfg.group_max_lengths = gc, gt, gl
        cfg.hash_channel = hash_channel
        cfg.n_aggregates = len(aggregates)
        cfg.aggregates = aggs
        cfg.step = step
        cfg.expected_groups = expected_groups
        cfg.options = _options(**options)
        return cfg, (gc, gt, gl, aggs), None

    def __init__(self, group_types: Sequence[int], group_channels: Sequence[int], aggregates: Sequence[Aggregate], *, backend=None, **kwargs):
        backend = backend or _default_backend()
        cfg, keep, _ = self._config(group_types, group_channels, aggregates, **kwargs)
        handle = C.c_void_p()
        backend.check(backend.hash_aggregation_create(C.byref(cfg), C.byref(handle)))
        super().__init__(backend, handle)
        self.cuda_stream = int(kwargs.get("cuda_stream", 0) or 0)


class GroupByHash:
    """GroupByHash.java:61-89 — createGroupByHash picks Bigint iff exactly one BIGINT channel (:46-59)."""

    def __init__(self, types: Sequence[int], channels: Sequence[int], *, hash_channel: int = -1, expected_size: int = 100,
                 max_lengths: Optional[Sequence[int]] = None, backend=None, **options):
        self._b = backend or _default_backend()
        cfg = abi.pg_group_by_hash_config()
        ch, ty = abi.i32_array(channels), abi.i32_array(types)
        ml = abi.i32_array(max_lengths if max_lengths is not None else [0] * len(types))
        cfg.n_channels = len(channels)
        cfg.channels, cfg.types, cfg.max_lengths = ch, ty, ml
        cfg.hash_channel = hash_channel
        cfg.expected_size = expected_size
        cfg.options = _options(**options)
        self._h = C.c_void_p()
        self._b.check(self._b.group_by_hash_create(C.byref(cfg), C.byref(self._h)))
        self._closed = False

    def add_page(self, page: Page) -> None:
        cpage, keep = page.as_c()
        self._b.check(self._b.group_by_hash_add_page(self._h, C.byref(cpage)))

    def get_group_ids(self, page: Page) -> np.ndarray:
        assert not page.on_device
        cpage, keep = page.as_c()
        out = np.empty(page.position_count, dtype=np.int64)
        self._b.check(self._b.group_by_hash_get_group_ids(self._h, C.byref(cpage), out.ctypes.data))
        return out

    def get_group_count(self) -> int:
        out = C.c_int64()
        self._b.check(self._b.group_by_hash_get_group_count(self._h, C.byref(out)))
        return out.value

    def contains(self, page: Page) -> np.ndarray:
        cpage, keep = page.as_c()
        out = np.empty(page.position_count, dtype=np.uint8)
        self._b.check(self._b.group_by_hash_contains(self._h, C.byref(cpage), out.ctypes.data))
        return out.astype(bool)

    def get_keys(self) -> Optional[Page]:
        out = abi.pg_page()
        self._b.check(self._b.group_by_hash_get_keys(self._h, C.byref(out)))
        return Page.from_c(out, self._b)

    def close(self) -> None:
        if not self._closed:
            self._closed = True
            self._b.group_by_hash_destroy(self._h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class JoinBridge:
    """JoinBridgeManager / LookupSourceFactory: the hand-off of the finished LookupSource from build to probe."""

    def __init__(self, device: int = 0, backend=None, outer: bool = False):
        """``outer``: the build side is an outer side (RIGHT / FULL OUTER): the LookupSource tracks the build positions the probes emit
        (OuterLookupSource.java:100-105) for a LookupOuterOperator."""
        self._b = backend or _default_backend()
        self._h = C.c_void_p()
        self._b.check(self._b.join_bridge_create(device, C.byref(self._h)))
        self._closed = False
        if outer:
            self._b.check(self._b.join_bridge_set_outer(self._h))

    def no_more_probe_operators(self) -> None:
        """The probe OperatorFactory is closed (JoinBridgeManager.probeOperatorFactoryClosed): once the live probe operators have
        finished, the LookupOuterOperator is unblocked."""
        self._b.check(self._b.join_bridge_no_more_probe_operators(self._h))

    def get_position_count(self) -> int:
        out = C.c_int64()
        self._b.check(self._b.join_bridge_get_position_count(self._h, C.byref(out)))
        return out.value

    def close(self) -> None:
        if not self._closed:
            self._closed = True
            self._b.join_bridge_destroy(self._h)


class HashBuilderOperator(Operator):
    """HashBuilderOperator.java:55-692."""

    FACTORY_KIND = 3

    @staticmethod
    def _config(bridge: JoinBridge, types: Sequence[int], hash_channels: Sequence[int], output_channels: Sequence[int], *,
                precomputed_hash_channel: int = -1, expected_positions: int = 10_000, filter_program=None, **options):
        """``filter_program``: the JoinFilterFunction — an expressions.Program over the build channels followed by the probe channels whose
        filter is the non-equi part of the join condition (JoinFilterFunctionCompiler; it belongs to the build side like in the reference)."""
        cfg = abi.pg_hash_builder_config()
        ty, hc, oc = abi.i32_array(types), abi.i32_array(hash_channels), abi.i32_array(output_channels)
        cfg.n_channels, cfg.types = len(types), ty
        cfg.n_hash_channels, cfg.hash_channels = len(hash_channels), hc
        cfg.n_output_channels, cfg.output_channels = len(output_channels), oc
        cfg.precomputed_hash_channel = precomputed_hash_channel
        cfg.expected_positions = expected_positions
        cfg.options = _options(**options)
        fkeep = None
        if filter_program is not None:
            assert filter_program.filter >= 0, "the join filter program needs set_filter(...)"
            fcfg = abi.pg_filter_project_config()
            it, nodes = abi.i32_array(filter_program.input_types), filter_program.c_nodes()
            fcfg.n_input_channels, fcfg.input_types = len(filter_program.input_types), it
            fcfg.n_nodes, fcfg.nodes = len(filter_program.nodes), nodes
            fcfg.filter_node = filter_program.filter
            fcfg.n_projections = 0
            cfg.filter = C.cast(C.pointer(fcfg), C.c_void_p)
            fkeep = (fcfg, it, nodes)
        return cfg, (ty, hc, oc, fkeep), bridge

    def __init__(self, bridge: JoinBridge, types: Sequence[int], hash_channels: Sequence[int], output_channels: Sequence[int], *, backend=None, **kwargs):
        backend = backend or bridge._b
        cfg, keep, _ = self._config(bridge, types, hash_channels, output_channels, **kwargs)
        handle = C.c_void_p()
        backend.check(backend.hash_builder_create(C.byref(cfg), bridge._h, C.byref(handle)))
        super().__init__(backend, handle)
        self._bridge = bridge


class LookupJoinOperator(Operator):
    """LookupJoinOperator.java:63-760 (INNER and PROBE_OUTER, no filter function)."""

    FACTORY_KIND = 4

    @staticmethod
    def _config(bridge: JoinBridge, types: Sequence[int], hash_channels: Sequence[int], output_channels: Sequence[int], *,
                join_type: int = abi.JOIN_INNER, precomputed_hash_channel: int = -1, **options):
        cfg = abi.pg_lookup_join_config()
        ty, hc, oc = abi.i32_array(types), abi.i32_array(hash_channels), abi.i32_array(output_channels)
        cfg.join_type = join_type
        cfg.n_channels, cfg.types = len(types), ty
        cfg.n_hash_channels, cfg.hash_channels = len(hash_channels), hc
        cfg.n_output_channels, cfg.output_channels = len(output_channels), oc
        cfg.precomputed_hash_channel = precomputed_hash_channel
        cfg.options = _options(**options)
        return cfg, (ty, hc, oc), bridge

    def __init__(self, bridge: JoinBridge, types: Sequence[int], hash_channels: Sequence[int], output_channels: Sequence[int], *, backend=None, **kwargs):
        backend = backend or bridge._b
        cfg, keep, _ = self._config(bridge, types, hash_channels, output_channels, **kwargs)
        handle = C.c_void_p()
        backend.check(backend.lookup_join_create(C.byref(cfg), bridge._h, C.byref(handle)))
        super().__init__(backend, handle)
        self._bridge = bridge


class LookupOuterOperator(Operator):
    """LookupOuterOperator.java:109-225: source operator emitting the build positions no probe visited (NULL probe channels + build
    output channels), unblocked once every probe operator of the (outer) bridge is done."""

    def __init__(self, bridge: JoinBridge, probe_output_types: Sequence[int], *, backend=None, **options):
        backend = backend or bridge._b
        cfg = abi.pg_lookup_outer_config()
        ty = abi.i32_array(probe_output_types)
        cfg.n_probe_output_channels, cfg.probe_output_types = len(probe_output_types), ty
        cfg.options = _options(**options)
        handle = C.c_void_p()
        backend.check(backend.lookup_outer_create(C.byref(cfg), bridge._h, C.byref(handle)))
        super().__init__(backend, handle)
        self._bridge = bridge


class HashSemiJoinOperator(Operator):
    """HashSemiJoinOperator.java:142-192: every input channel + a BOOLEAN "probe value is in the set" channel (SQL three-valued).
    The set (SetBuilderOperator / ChannelSet) is a JoinBridge filled by ``HashBuilderOperator(bridge, [type], [0], [])``."""

    def __init__(self, bridge: "JoinBridge", types: Sequence[int], probe_join_channel: int, *, backend=None, **options):
        backend = backend or _default_backend()
        cfg = abi.pg_semi_join_config()
        ty = abi.i32_array(types)
        cfg.n_channels, cfg.types = len(types), ty
        cfg.probe_join_channel = probe_join_channel
        cfg.precomputed_hash_channel = -1
        cfg.options = _options(**options)
        handle = C.c_void_p()
        backend.check(backend.hash_semi_join_create(C.byref(cfg), bridge._h, C.byref(handle)))
        super().__init__(backend, handle)


class PartitionedOutputOperator(Operator):
    """OptimizedPartitionedOutputOperator.java:167-179,384-824 — the partition + copy-by-position step.  get_output returns
    one page per input page with all partitions back to back; ``get_offsets`` gives the partition boundaries."""

    def __init__(self, types: Sequence[int], partition_count: int, partition_channels: Sequence[int], *,
                 precomputed_hash_channel: int = -1, replicates_any_row: bool = False, null_channel: int = -1,
                 partition_function: int = abi.PARTITION_HASH_GENERATOR, bucket_to_partition: Optional[Sequence[int]] = None, backend=None, **options):
        backend = backend or _default_backend()
        cfg = abi.pg_partitioned_output_config()
        cfg.partition_function = partition_function
        if bucket_to_partition is not None:       # BucketPartitionFunction: the function yields a bucket, the table names its partition
            table = abi.i32_array(bucket_to_partition)
            cfg.bucket_count, cfg.bucket_to_partition = len(bucket_to_partition), table
        ty, pc = abi.i32_array(types), abi.i32_array(partition_channels)
        cfg.n_channels, cfg.types = len(types), ty
        cfg.partition_count = partition_count
        cfg.n_partition_channels, cfg.partition_channels = len(partition_channels), pc
        cfg.precomputed_hash_channel = precomputed_hash_channel
        cfg.replicates_any_row = int(replicates_any_row)
        cfg.null_channel = null_channel
        cfg.options = _options(**options)
        handle = C.c_void_p()
        backend.check(backend.partitioned_output_create(C.byref(cfg), C.byref(handle)))
        super().__init__(backend, handle)
        self.partition_count = partition_count

    def get_offsets(self) -> np.ndarray:
        out = np.empty(self.partition_count + 1, dtype=np.int64)
        self._b.check(self._b.partitioned_output_get_offsets(self._h, out.ctypes.data))
        return out

    # ---- fused repartition + exchange (exchange.FusedExchange drives these) ----
    def count(self, page: Page) -> np.ndarray:
        """Pass 1 only: rows of ``page`` per partition.  The page stays staged in the operator until ``scatter_to``."""
        out = np.empty(self.partition_count, dtype=np.int64)
        cpage, keep = page.as_c()
        self._b.check(self._b.partitioned_output_count(self._h, C.byref(cpage), out.ctypes.data))
        self._count_keepalive = (page, keep)
        return out

    def scatter_to(self, dst_values: np.ndarray, dst_nulls: Optional[np.ndarray], dst_row_offsets: np.ndarray) -> None:
        """Pass 2: partition p's rows go to ``dst_values[p, c]`` (device addresses, possibly peer-GPU memory) starting at row
        ``dst_row_offsets[p]``, keeping their input order."""
        dv = np.ascontiguousarray(dst_values, dtype=np.uint64)
        dn = None if dst_nulls is None else np.ascontiguousarray(dst_nulls, dtype=np.uint64)
        ro = np.ascontiguousarray(dst_row_offsets, dtype=np.int64)
        self._b.check(self._b.partitioned_output_scatter_to(self._h, dv.ctypes.data, None if dn is None else dn.ctypes.data, ro.ctypes.data))
        self._count_keepalive = None


class FilterAndProjectOperator(Operator):
    """FilterAndProjectOperator.java:88-99 over a PageProcessor (PageProcessor.java:112-148)."""

    FACTORY_KIND = 2

    @staticmethod
    def _config(input_types: Sequence[int], program, **options):
        cfg = abi.pg_filter_project_config()
        it = abi.i32_array(input_types)
        nodes = program.c_nodes()
        proj = abi.i32_array(program.projections)
        cfg.n_input_channels, cfg.input_types = len(input_types), it
        cfg.n_nodes, cfg.nodes = len(program.nodes), nodes
        cfg.filter_node = program.filter
        cfg.n_projections, cfg.projection_nodes = len(program.projections), proj
        cfg.options = _options(**options)
        return cfg, (it, nodes, proj), None

    def __init__(self, input_types: Sequence[int], program, *, backend=None, **options):
        """``program`` is an expressions.Program (flattened RowExpression DAG + filter + projections)."""
        backend = backend or _default_backend()
        cfg, keep, _ = self._config(input_types, program, **options)
        handle = C.c_void_p()
        backend.check(backend.filter_project_create(C.byref(cfg), C.byref(handle)))
        super().__init__(backend, handle)

    @staticmethod
    def explain(input_types: Sequence[int], program, nullable_inputs: Sequence[int] = (), backend=None) -> str:
        """The PTX generated for ``program`` when the channels in ``nullable_inputs`` carry null maps (no GPU needed)."""
        if backend is None:
            from ._lib import load_library
            backend = load_library()
        cfg = abi.pg_filter_project_config()
        it = abi.i32_array(input_types)
        nodes = program.c_nodes()
        proj = abi.i32_array(program.projections)
        cfg.n_input_channels, cfg.input_types = len(input_types), it
        cfg.n_nodes, cfg.nodes = len(program.nodes), nodes
        cfg.filter_node = program.filter
        cfg.n_projections, cfg.projection_nodes = len(program.projections), proj
        mask = 0
        for ch in nullable_inputs:
            mask |= 1 << ch
        need = C.c_size_t()
        backend.check(backend.filter_project_explain(C.byref(cfg), mask, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        backend.check(backend.filter_project_explain(C.byref(cfg), mask, buf, need.value, C.byref(need)))
        return buf.value.decode()

    def get_last_selection(self):
        is_list, size = C.c_int32(), C.c_int64()
        self._b.check(self._b.filter_project_get_last_selection(self._h, C.byref(is_list), C.byref(size)))
        return bool(is_list.value), size.value


class FusedJoin:
    """One INNER LookupJoinOperator of a fused fragment: probes ``bridge`` with channel ``hash_channels`` of the incoming layout."""

    def __init__(self, bridge: "JoinBridge", types: Sequence[int], hash_channels: Sequence[int], output_channels: Sequence[int]):
        self.bridge, self.types, self.hash_channels, self.output_channels = bridge, list(types), list(hash_channels), list(output_channels)


class FusedAggregationOperator(_AggregationState, Operator):
    """FilterAndProjectOperator → LookupJoinOperator* → HashAggregationOperator of one Driver as ONE operator
    (include/presto_gpu.h pg_fused_aggregation_create): on the CUDA backend the fragment is generated as a single kernel; the tests'
    oracle chains the three reference operators the way Driver.processInternal does — the results are the same by definition."""

    def __init__(self, input_types: Sequence[int], program, joins: Sequence[FusedJoin], group_types: Sequence[int], group_channels: Sequence[int],
                 aggregates: Sequence[Aggregate], *, step: int = abi.STEP_SINGLE, expected_groups: int = 10_000,
                 group_max_lengths: Optional[Sequence[int]] = None, hash_channel: int = -1, backend=None, **options):
        backend = backend or _default_backend()
        cfg, self._keep = self._config(input_types, program, joins, group_types, group_channels, aggregates, step, expected_groups,
                                       group_max_lengths, hash_channel, options)
        handle = C.c_void_p()
        backend.check(backend.fused_aggregation_create(C.byref(cfg), C.byref(handle)))
        super().__init__(backend, handle)
        self.cuda_stream = int(options.get("cuda_stream", 0) or 0)

    @staticmethod
    def _config(input_types, program, joins, group_types, group_channels, aggregates, step, expected_groups, group_max_lengths, hash_channel, options):
        cfg = abi.pg_fused_aggregation_config()
        keep = []
        it, nodes, proj = abi.i32_array(input_types), program.c_nodes(), abi.i32_array(program.projections)
        f = cfg.page_processor
        f.n_input_channels, f.input_types = len(input_types), it
        f.n_nodes, f.nodes = len(program.nodes), nodes
        f.filter_node = program.filter
        f.n_projections, f.projection_nodes = len(program.projections), proj
        keep += [it, nodes, proj]
        jarr = (abi.pg_lookup_join_config * max(1, len(joins)))()
        barr = (C.c_void_p * max(1, len(joins)))()
        for i, j in enumerate(joins):
            ty, hc, oc = abi.i32_array(j.types), abi.i32_array(j.hash_channels), abi.i32_array(j.output_channels)
            jarr[i].join_type = abi.JOIN_INNER
            jarr[i].n_channels, jarr[i].types = len(j.types), ty
            jarr[i].n_hash_channels, jarr[i].hash_channels = len(j.hash_channels), hc
            jarr[i].n_output_channels, jarr[i].output_channels = len(j.output_channels), oc
            jarr[i].precomputed_hash_channel = -1
            barr[i] = j.bridge._h if j.bridge is not None else None
            keep += [ty, hc, oc]
        cfg.n_joins, cfg.joins, cfg.bridges = len(joins), jarr, barr
        keep += [jarr, barr]
        a = cfg.aggregation
        gc, gt = abi.i32_array(group_channels), abi.i32_array(group_types)
        gl = abi.i32_array(group_max_lengths if group_max_lengths is not None else [0] * len(group_types))
        aggs = (abi.pg_aggregate * max(1, len(aggregates)))()
        for i, g in enumerate(aggregates):
            aggs[i].function, aggs[i].input_channel, aggs[i].mask_channel = g.function, g.input_channel, g.mask_channel
        a.n_group_channels = len(group_channels)
        a.group_channels, a.group_types, a.group_max_lengths = gc, gt, gl
        a.hash_channel = hash_channel
        a.n_aggregates, a.aggregates = len(aggregates), aggs
        a.step, a.expected_groups = step, expected_groups
        a.options = _options(**options)
        keep += [gc, gt, gl, aggs]
        return cfg, keep

    @staticmethod
    def explain(input_types, program, joins, build_types, group_types, group_channels, aggregates, *, nullable_inputs=(), step: int = abi.STEP_SINGLE,
                group_max_lengths=None, backend=None) -> str:
        """The CUDA C++ generated for the fragment (no GPU needed); ``build_types[j]`` = types of join j's build output channels."""
        if backend is None:
            from ._lib import load_library
            backend = load_library()
        cfg, keep = FusedAggregationOperator._config(input_types, program, joins, group_types, group_channels, aggregates, step, 1000,
                                                     group_max_lengths, -1, {})
        bt = [abi.i32_array(t) for t in build_types]
        bptr = (C.POINTER(C.c_int32) * max(1, len(bt)))(*[C.cast(b, C.POINTER(C.c_int32)) for b in bt])
        nb = abi.i32_array([len(t) for t in build_types])
        mask = 0
        for ch in nullable_inputs:
            mask |= 1 << ch
        need = C.c_size_t()
        backend.check(backend.fused_aggregation_explain(C.byref(cfg), mask, bptr, nb, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        backend.check(backend.fused_aggregation_explain(C.byref(cfg), mask, bptr, nb, buf, need.value, C.byref(need)))
        return buf.value.decode()

    def get_info(self) -> dict:
        if not self._b.has_device:
            return {"fused_pages": 0, "unfused_pages": 0, "registers": 0, "reason": "reference operators chained"}
        f, u, r = C.c_int64(), C.c_int64(), C.c_int32()
        buf = C.create_string_buffer(2048)
        self._b.check(self._b.fused_aggregation_get_info(self._h, C.byref(f), C.byref(u), C.byref(r), buf, 2048))
        return {"fused_pages": f.value, "unfused_pages": u.value, "registers": r.value, "reason": buf.value.decode()}


class FusedPageOperator(Operator):
    """FilterAndProjectOperator → LookupJoinOperator* of one Driver as ONE operator (include/presto_gpu.h pg_fused_page_create); its
    output pages are the last operator's.  The oracle chains the reference operators."""

    def __init__(self, input_types: Sequence[int], program, joins: Sequence[FusedJoin], *, backend=None, **options):
        backend = backend or _default_backend()
        cfg, self._keep = self._config(input_types, program, joins, options)
        handle = C.c_void_p()
        backend.check(backend.fused_page_create(C.byref(cfg), C.byref(handle)))
        super().__init__(backend, handle)

    @staticmethod
    def explain(input_types, program, joins, build_types, *, nullable_inputs=(), backend=None) -> str:
        """The CUDA C++ of the selection and emit passes (no GPU needed); ``build_types[j]`` = types of join j's build output channels."""
        if backend is None:
            from ._lib import load_library
            backend = load_library()
        cfg, keep = FusedPageOperator._config(input_types, program, joins, {})
        bt = [abi.i32_array(t) for t in build_types]
        bptr = (C.POINTER(C.c_int32) * max(1, len(bt)))(*[C.cast(b, C.POINTER(C.c_int32)) for b in bt])
        nb = abi.i32_array([len(t) for t in build_types])
        mask = 0
        for ch in nullable_inputs:
            mask |= 1 << ch
        need = C.c_size_t()
        backend.check(backend.fused_page_explain(C.byref(cfg), mask, bptr, nb, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        backend.check(backend.fused_page_explain(C.byref(cfg), mask, bptr, nb, buf, need.value, C.byref(need)))
        return buf.value.decode()

    @staticmethod
    def _config(input_types, program, joins, options):
        cfg = abi.pg_fused_page_config()
        self = type("K", (), {})()
        it, nodes, proj = abi.i32_array(input_types), program.c_nodes(), abi.i32_array(program.projections)
        f = cfg.page_processor
        f.n_input_channels, f.input_types = len(input_types), it
        f.n_nodes, f.nodes = len(program.nodes), nodes
        f.filter_node = program.filter
        f.n_projections, f.projection_nodes = len(program.projections), proj
        self._keep = [it, nodes, proj]
        jarr = (abi.pg_lookup_join_config * max(1, len(joins)))()
        barr = (C.c_void_p * max(1, len(joins)))()
        for i, j in enumerate(joins):
            ty, hc, oc = abi.i32_array(j.types), abi.i32_array(j.hash_channels), abi.i32_array(j.output_channels)
            jarr[i].join_type = abi.JOIN_INNER
            jarr[i].n_channels, jarr[i].types = len(j.types), ty
            jarr[i].n_hash_channels, jarr[i].hash_channels = len(j.hash_channels), hc
            jarr[i].n_output_channels, jarr[i].output_channels = len(j.output_channels), oc
            jarr[i].precomputed_hash_channel = -1
            barr[i] = j.bridge._h if j.bridge is not None else None
            self._keep += [ty, hc, oc]
        cfg.n_joins, cfg.joins, cfg.bridges = len(joins), jarr, barr
        self._keep += [jarr, barr]
        cfg.options = _options(**options)
        return cfg, self._keep

    def get_info(self) -> dict:
        if not self._b.has_device:
            return {"fused_pages": 0, "unfused_pages": 0, "reason": "reference operators chained"}
        f, u = C.c_int64(), C.c_int64()
        buf = C.create_string_buffer(2048)
        self._b.check(self._b.fused_page_get_info(self._h, C.byref(f), C.byref(u), buf, 2048))
        return {"fused_pages": f.value, "unfused_pages": u.value, "reason": buf.value.decode()}


class OperatorFactory:
    """OperatorFactory.java:19-51 over the C ABI's pg_operator_factory_*: ``OperatorFactory(HashAggregationOperator, group_types, ...)``
    takes the operator class and the arguments its constructor takes (the plan node's part; the factory keeps a deep copy), and
    ``create_operator(**driver_options)`` hands out one operator per Driver (device / cuda_stream / output_memory / max_output_rows)."""

    def __init__(self, operator_class, *args, backend=None, _handle=None, **kwargs):
        self._cls = operator_class
        self._closed = False
        if _handle is not None:
            self._b, self._h, self._bridge = backend, _handle, kwargs.get("_bridge")
            return
        cfg, keep, bridge = operator_class._config(*args, **kwargs)
        self._b = backend or (bridge._b if bridge is not None else _default_backend())
        self._bridge = bridge
        self._h = C.c_void_p()
        self._b.check(self._b.operator_factory_create(operator_class.FACTORY_KIND, C.byref(cfg), bridge._h if bridge is not None else None, C.byref(self._h)))
        del keep          # the factory copied every array

    def create_operator(self, **options) -> Operator:
        handle = C.c_void_p()
        opts = _options(**options) if options else None
        self._b.check(self._b.operator_factory_create_operator(self._h, C.byref(opts) if opts is not None else None, C.byref(handle)))
        op = self._cls.__new__(self._cls)
        Operator.__init__(op, self._b, handle)
        op.cuda_stream = int(options.get("cuda_stream", 0) or 0)
        op._bridge = self._bridge
        return op

    def no_more_operators(self) -> None:
        self._b.check(self._b.operator_factory_no_more_operators(self._h))

    def duplicate(self) -> "OperatorFactory":
        handle = C.c_void_p()
        self._b.check(self._b.operator_factory_duplicate(self._h, C.byref(handle)))
        return OperatorFactory(self._cls, backend=self._b, _handle=handle, _bridge=self._bridge)

    @property
    def created_count(self) -> int:
        out = C.c_int64()
        self._b.check(self._b.operator_factory_get_created_count(self._h, C.byref(out)))
        return int(out.value)

    def close(self) -> None:
        if not self._closed:
            self._b.operator_factory_destroy(self._h)
            self._closed = True


class PagesSerde:
    """PagesSerde (presto-spi/.../spi/page/PagesSerde.java:67-156) + PagesSerdeUtil.writeSerializedPage / readSerializedPage
    (:64-88): pages to and from the SerializedPage wire bytes (21-byte header + raw page)."""

    def __init__(self, *, checksum_enabled: bool = False, compression: int = abi.COMPRESSION_NONE, backend=None, **options):
        self._b = backend or _default_backend()
        cfg = abi.pg_pages_serde_config()
        cfg.checksum_enabled = int(checksum_enabled)
        cfg.compression = compression        # COMPRESSION_LZ4: pages that shrink to <= 0.8 travel as one LZ4 block (PagesSerde.java:118-143)
        cfg.options = _options(**options)
        self._memory = cfg.options.output_memory
        self._h = C.c_void_p()
        self._b.check(self._b.pages_serde_create(C.byref(cfg), C.byref(self._h)))

    def serialize(self, page: Page, copy: bool = True):
        """The wire bytes of ``page``: a numpy uint8 array (copy=True) or (address, size) of serde-owned memory in the serde's
        output memory (valid until the next call)."""
        cpage, keep = page.as_c()
        ptr, size = C.c_void_p(), C.c_int64()
        self._b.check(self._b.pages_serde_serialize(self._h, C.byref(cpage), C.byref(ptr), C.byref(size)))
        del keep
        if not copy:
            return ptr.value, size.value
        out = np.empty(size.value, dtype=np.uint8)
        if self._memory == abi.MEM_DEVICE:
            self._b.check(self._b.memcpy_d2h(out.ctypes.data, ptr, size.value))
        else:
            C.memmove(out.ctypes.data, ptr, size.value)
        return out

    def serialize_ranges(self, page: Page, row_offsets: Sequence[int]) -> List[np.ndarray]:
        """One SerializedPage per row range [row_offsets[r], row_offsets[r + 1]) — a PartitionedOutputOperator's output page and its
        partition offsets give the pages OptimizedPartitionedOutputOperator.flush enqueues (one native call, the serde runs on the GPU)."""
        cpage, keep = page.as_c()
        n = len(row_offsets) - 1
        offs = (C.c_int64 * (n + 1))(*[int(x) for x in row_offsets])
        byte_offsets = (C.c_int64 * (n + 1))()
        ptr = C.c_void_p()
        self._b.check(self._b.pages_serde_serialize_ranges(self._h, C.byref(cpage), offs, n, C.byref(ptr), byte_offsets))
        del keep
        total = int(byte_offsets[n])
        buf = np.empty(total, dtype=np.uint8)
        if total:
            if self._memory == abi.MEM_DEVICE:
                self._b.check(self._b.memcpy_d2h(buf.ctypes.data, ptr, total))
            else:
                C.memmove(buf.ctypes.data, ptr, total)
        return [buf[int(byte_offsets[r]):int(byte_offsets[r + 1])] for r in range(n)]

    def deserialize(self, data, types: Sequence[int], *, size: Optional[int] = None, memory: int = abi.MEM_HOST, copy: bool = True) -> Page:
        """``data``: numpy uint8 array (host) or an address (with ``size`` and ``memory``)."""
        if isinstance(data, np.ndarray):
            data = np.ascontiguousarray(data, dtype=np.uint8)
            ptr, size, memory = data.ctypes.data, data.size, abi.MEM_HOST
        else:
            ptr = int(data)
        ty = abi.i32_array(types)
        out = abi.pg_page()
        self._b.check(self._b.pages_serde_deserialize(self._h, C.c_void_p(ptr), size, memory, ty, len(types), C.byref(out)))
        return Page.from_c(out, self._b, copy=copy)

    def close(self) -> None:
        if self._h:
            self._b.pages_serde_destroy(self._h)
            self._h = C.c_void_p()
