"""ctypes mirror of include/presto_gpu.h (the C ABI of libprestogpu.so).

Every Structure below is field-for-field the struct of the same name in the header.  `Backend` binds the entry points
of a shared library by prefix: the product uses prefix ``pg_`` on ``presto_b200/csrc/libprestogpu.so``; the test-only
oracle (``oracle/``) exposes the same shapes with prefix ``or_`` and constructs its own Backend — this module never
imports or loads it.
"""
from __future__ import annotations

import ctypes as C

# status codes (presto_gpu.h)
PG_OK = 0
PG_ERR_INVALID_ARGUMENT = 1
PG_ERR_INVALID_STATE = 2
PG_ERR_UNSUPPORTED = 3
PG_ERR_OUT_OF_MEMORY = 4
PG_ERR_INSUFFICIENT_RESOURCES = 5
PG_ERR_NUMERIC_OVERFLOW = 6
PG_ERR_CUDA = 7
PG_ERR_NO_DEVICE = 8

_STATUS_NAMES = {
    1: "INVALID_ARGUMENT", 2: "INVALID_STATE", 3: "UNSUPPORTED", 4: "OUT_OF_MEMORY", 5: "GENERIC_INSUFFICIENT_RESOURCES",
    6: "NUMERIC_VALUE_OUT_OF_RANGE", 7: "CUDA_ERROR", 8: "NO_DEVICE",
}

# pg_type
BIGINT, INTEGER, DOUBLE, DATE, BOOLEAN, VARCHAR, LONG_DECIMAL = 1, 2, 3, 4, 5, 6, 7
IPC_HANDLE_BYTES = 64
# pg_encoding
ENC_FLAT, ENC_DICTIONARY, ENC_RLE = 0, 1, 2
# pg_memory
MEM_HOST, MEM_DEVICE = 0, 1
# pg_aggregate_function
AGG_COUNT_STAR, AGG_COUNT_COLUMN, AGG_SUM_DOUBLE, AGG_SUM_BIGINT, AGG_AVG_DOUBLE, AGG_AVG_BIGINT = 0, 1, 2, 3, 4, 5
AGG_MIN_DOUBLE, AGG_MAX_DOUBLE, AGG_MIN_BIGINT, AGG_MAX_BIGINT = 6, 7, 8, 9
AGG_SUM_SHORT_DECIMAL, AGG_SUM_LONG_DECIMAL = 10, 11
# pg_aggregation_step
STEP_SINGLE, STEP_PARTIAL, STEP_FINAL = 0, 1, 2
# pg_join_type
JOIN_INNER, JOIN_PROBE_OUTER, JOIN_LOOKUP_OUTER, JOIN_FULL_OUTER = 0, 1, 2, 3
# pg_expr_op
EX_INPUT, EX_CONST_BIGINT, EX_CONST_DOUBLE, EX_CONST_BOOLEAN, EX_CONST_NULL = 0, 1, 2, 3, 4
EX_CONST_VARCHAR = 5
EX_ADD, EX_SUB, EX_MUL, EX_DIV, EX_NEG = 10, 11, 12, 13, 14
EX_EQ, EX_NE, EX_LT, EX_LE, EX_GT, EX_GE = 20, 21, 22, 23, 24, 25
EX_AND, EX_OR, EX_NOT, EX_IS_NULL = 30, 31, 32, 33
EX_CAST_DOUBLE, EX_CAST_BIGINT = 40, 41
EX_BETWEEN = 50
EX_HASH, EX_COMBINE_HASH = 60, 61


class PrestoException(RuntimeError):
    """Python twin of the reference's PrestoException(StandardErrorCode, message)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"{_STATUS_NAMES.get(code, code)}: {message}")
        self.code = code
        self.message = message


class pg_block(C.Structure):
    _fields_ = [
        ("type", C.c_int32), ("encoding", C.c_int32), ("position_count", C.c_int64),
        ("values", C.c_void_p), ("is_null", C.c_void_p), ("offsets", C.c_void_p), ("ids", C.c_void_p),
        ("dictionary_size", C.c_int64),
    ]


class pg_page(C.Structure):
    _fields_ = [("position_count", C.c_int64), ("channel_count", C.c_int32), ("memory", C.c_int32), ("blocks", C.POINTER(pg_block))]


class pg_operator_stats(C.Structure):
    _fields_ = [
        ("input_positions", C.c_int64), ("input_bytes", C.c_int64), ("output_positions", C.c_int64), ("output_bytes", C.c_int64),
        ("device_bytes", C.c_int64), ("kernel_launches", C.c_int64), ("group_count", C.c_int64), ("h2d_bytes", C.c_int64),
        ("d2h_bytes", C.c_int64), ("gpu_ms", C.c_double), ("kernel_paths", C.c_int64),
    ]


PATH_AGG_FAST, PATH_AGG_WARP, PATH_AGG_MID, PATH_AGG_DIRECT, PATH_AGG_GENERIC, PATH_FUSED_TIERS, PATH_FUSED_DIRECT, PATH_UNFUSED_CHAIN = 1, 2, 4, 8, 16, 32, 64, 128
PATH_FUSED_TMA = 256
PATH_INTERNED_KEYS = 512
PATH_DICTIONARY_KEYS = 1024
PATH_FP_JIT, PATH_FP_INTERPRETER = 2048, 4096


class pg_operator_options(C.Structure):
    _fields_ = [("device", C.c_int32), ("output_memory", C.c_int32), ("cuda_stream", C.c_void_p), ("max_output_rows", C.c_int64)]


class pg_aggregate(C.Structure):
    _fields_ = [("function", C.c_int32), ("input_channel", C.c_int32), ("mask_channel", C.c_int32), ("reserved", C.c_int32)]


_I32P = C.POINTER(C.c_int32)


class pg_hash_aggregation_config(C.Structure):
    _fields_ = [
        ("n_group_channels", C.c_int32), ("group_channels", _I32P), ("group_types", _I32P), ("group_max_lengths", _I32P),
        ("hash_channel", C.c_int32), ("n_aggregates", C.c_int32), ("aggregates", C.POINTER(pg_aggregate)),
        ("step", C.c_int32), ("expected_groups", C.c_int32), ("options", pg_operator_options),
    ]


class pg_group_by_hash_config(C.Structure):
    _fields_ = [
        ("n_channels", C.c_int32), ("channels", _I32P), ("types", _I32P), ("max_lengths", _I32P),
        ("hash_channel", C.c_int32), ("expected_size", C.c_int32), ("options", pg_operator_options),
    ]


class pg_hash_builder_config(C.Structure):
    _fields_ = [
        ("n_channels", C.c_int32), ("types", _I32P), ("n_hash_channels", C.c_int32), ("hash_channels", _I32P),
        ("n_output_channels", C.c_int32), ("output_channels", _I32P), ("precomputed_hash_channel", C.c_int32),
        ("expected_positions", C.c_int64), ("options", pg_operator_options),
        ("filter", C.c_void_p),      # const pg_filter_project_config* (the JoinFilterFunction) or NULL
    ]


class pg_lookup_join_config(C.Structure):
    _fields_ = [
        ("join_type", C.c_int32), ("n_channels", C.c_int32), ("types", _I32P), ("n_hash_channels", C.c_int32), ("hash_channels", _I32P),
        ("n_output_channels", C.c_int32), ("output_channels", _I32P), ("precomputed_hash_channel", C.c_int32),
        ("options", pg_operator_options),
    ]


class pg_lookup_outer_config(C.Structure):
    _fields_ = [("n_probe_output_channels", C.c_int32), ("probe_output_types", _I32P), ("options", pg_operator_options)]


class pg_partitioned_output_config(C.Structure):
    _fields_ = [
        ("n_channels", C.c_int32), ("types", _I32P), ("partition_count", C.c_int32), ("n_partition_channels", C.c_int32),
        ("partition_channels", _I32P), ("precomputed_hash_channel", C.c_int32), ("replicates_any_row", C.c_int32),
        ("null_channel", C.c_int32), ("partition_function", C.c_int32), ("bucket_count", C.c_int32), ("bucket_to_partition", _I32P),
        ("options", pg_operator_options),
    ]


PARTITION_HASH_GENERATOR, PARTITION_LOCAL, PARTITION_ROUND_ROBIN = 0, 1, 2


class pg_expr_node(C.Structure):
    _fields_ = [
        ("op", C.c_int32), ("type", C.c_int32), ("a", C.c_int32), ("b", C.c_int32), ("c", C.c_int32), ("reserved", C.c_int32),
        ("imm_i", C.c_int64), ("imm_d", C.c_double),
    ]


class pg_filter_project_config(C.Structure):
    _fields_ = [
        ("n_input_channels", C.c_int32), ("input_types", _I32P), ("n_nodes", C.c_int32), ("nodes", C.POINTER(pg_expr_node)),
        ("filter_node", C.c_int32), ("n_projections", C.c_int32), ("projection_nodes", _I32P), ("options", pg_operator_options),
    ]


class pg_fused_aggregation_config(C.Structure):
    _fields_ = [("page_processor", pg_filter_project_config), ("n_joins", C.c_int32), ("joins", C.POINTER(pg_lookup_join_config)),
                ("bridges", C.POINTER(C.c_void_p)), ("aggregation", pg_hash_aggregation_config)]


class pg_fused_page_config(C.Structure):
    _fields_ = [("page_processor", pg_filter_project_config), ("n_joins", C.c_int32), ("joins", C.POINTER(pg_lookup_join_config)),
                ("bridges", C.POINTER(C.c_void_p)), ("options", pg_operator_options)]


def i32_array(values):
    values = list(values)
    return (C.c_int32 * max(1, len(values)))(*values)


class pg_semi_join_config(C.Structure):
    _fields_ = [("n_channels", C.c_int32), ("types", _I32P), ("probe_join_channel", C.c_int32), ("precomputed_hash_channel", C.c_int32),
                ("options", pg_operator_options)]


class pg_pages_serde_config(C.Structure):
    _fields_ = [("checksum_enabled", C.c_int32), ("compression", C.c_int32), ("options", pg_operator_options)]


SERIALIZED_PAGE_HEADER_BYTES = 21
COMPRESSION_NONE, COMPRESSION_LZ4 = 0, 1
PAGE_MARKER_COMPRESSED, PAGE_MARKER_ENCRYPTED, PAGE_MARKER_CHECKSUMMED = 1, 2, 4


class Backend:
    """Entry points of one shared library, bound by prefix (``pg_`` for the CUDA product)."""

    # name -> (restype, argtypes); operator handles travel as void*
    _SIGNATURES = {
        "version": (C.c_char_p, []),
        "last_error": (C.c_char_p, []),
        "hash_bigint": (C.c_int64, [C.c_int64]),
        "hash_combine": (C.c_int64, [C.c_int64, C.c_int64]),
        "hash_murmur3_fmix64": (C.c_int64, [C.c_int64]),
        "hash_partition": (C.c_int32, [C.c_int64, C.c_int32]),
        "hash_xxhash64": (C.c_int64, [C.c_void_p, C.c_size_t]),
        "hash_double": (C.c_int64, [C.c_double]),
        "hash_local_partition": (C.c_int32, [C.c_int64, C.c_int32]),
        "operator_needs_input": (C.c_int32, [C.c_void_p, C.POINTER(C.c_int32)]),
        "operator_is_blocked": (C.c_int32, [C.c_void_p, C.POINTER(C.c_int32)]),
        "operator_add_input": (C.c_int32, [C.c_void_p, C.POINTER(pg_page)]),
        "operator_get_output": (C.c_int32, [C.c_void_p, C.POINTER(pg_page)]),
        "operator_finish": (C.c_int32, [C.c_void_p]),
        "operator_is_finished": (C.c_int32, [C.c_void_p, C.POINTER(C.c_int32)]),
        "operator_get_stats": (C.c_int32, [C.c_void_p, C.POINTER(pg_operator_stats)]),
        "operator_destroy": (C.c_int32, [C.c_void_p]),
        "hash_aggregation_create": (C.c_int32, [C.POINTER(pg_hash_aggregation_config), C.POINTER(C.c_void_p)]),
        "group_by_hash_create": (C.c_int32, [C.POINTER(pg_group_by_hash_config), C.POINTER(C.c_void_p)]),
        "group_by_hash_add_page": (C.c_int32, [C.c_void_p, C.POINTER(pg_page)]),
        "group_by_hash_get_group_ids": (C.c_int32, [C.c_void_p, C.POINTER(pg_page), C.c_void_p]),
        "group_by_hash_get_group_count": (C.c_int32, [C.c_void_p, C.POINTER(C.c_int64)]),
        "group_by_hash_contains": (C.c_int32, [C.c_void_p, C.POINTER(pg_page), C.c_void_p]),
        "group_by_hash_get_keys": (C.c_int32, [C.c_void_p, C.POINTER(pg_page)]),
        "group_by_hash_destroy": (C.c_int32, [C.c_void_p]),
        "join_bridge_create": (C.c_int32, [C.c_int32, C.POINTER(C.c_void_p)]),
        "join_bridge_destroy": (C.c_int32, [C.c_void_p]),
        "join_bridge_get_position_count": (C.c_int32, [C.c_void_p, C.POINTER(C.c_int64)]),
        "hash_builder_create": (C.c_int32, [C.POINTER(pg_hash_builder_config), C.c_void_p, C.POINTER(C.c_void_p)]),
        "lookup_join_create": (C.c_int32, [C.POINTER(pg_lookup_join_config), C.c_void_p, C.POINTER(C.c_void_p)]),
        "join_bridge_set_outer": (C.c_int32, [C.c_void_p]),
        "join_bridge_no_more_probe_operators": (C.c_int32, [C.c_void_p]),
        "lookup_outer_create": (C.c_int32, [C.POINTER(pg_lookup_outer_config), C.c_void_p, C.POINTER(C.c_void_p)]),
        "hash_semi_join_create": (C.c_int32, [C.POINTER(pg_semi_join_config), C.c_void_p, C.POINTER(C.c_void_p)]),
        "partitioned_output_create": (C.c_int32, [C.POINTER(pg_partitioned_output_config), C.POINTER(C.c_void_p)]),
        "partitioned_output_get_offsets": (C.c_int32, [C.c_void_p, C.c_void_p]),
        "filter_project_create": (C.c_int32, [C.POINTER(pg_filter_project_config), C.POINTER(C.c_void_p)]),
        "filter_project_get_last_selection": (C.c_int32, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
        "fused_aggregation_create": (C.c_int32, [C.POINTER(pg_fused_aggregation_config), C.POINTER(C.c_void_p)]),
        "fused_page_create": (C.c_int32, [C.POINTER(pg_fused_page_config), C.POINTER(C.c_void_p)]),
        "pages_serde_create": (C.c_int32, [C.POINTER(pg_pages_serde_config), C.POINTER(C.c_void_p)]),
        "pages_serde_serialize": (C.c_int32, [C.c_void_p, C.POINTER(pg_page), C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
        "pages_serde_serialize_ranges": (C.c_int32, [C.c_void_p, C.POINTER(pg_page), C.POINTER(C.c_int64), C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
        "pages_serde_deserialize": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.POINTER(pg_page)]),
        "pages_serde_destroy": (C.c_int32, [C.c_void_p]),
    }
    # only the CUDA product exports these
    _DEVICE_SIGNATURES = {
        "device_count": (C.c_int32, [C.POINTER(C.c_int32)]),
        "set_device": (C.c_int32, [C.c_int32]),
        "host_alloc": (C.c_int32, [C.c_size_t, C.POINTER(C.c_void_p)]),
        "host_free": (C.c_int32, [C.c_void_p]),
        "device_alloc": (C.c_int32, [C.c_size_t, C.POINTER(C.c_void_p)]),
        "device_free": (C.c_int32, [C.c_void_p]),
        "memcpy_h2d": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t]),
        "memcpy_d2h": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t]),
        "device_synchronize": (C.c_int32, []),
        "release_cached_memory": (C.c_int32, []),
        "hash_crc32_chunked": (C.c_int64, [C.c_void_p, C.c_size_t, C.c_size_t]),
        "lz4_compress_bound": (C.c_int64, [C.c_int64]),
        "lz4_compress_host": (C.c_int64, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]),
        "lz4_decompress_host": (C.c_int64, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]),
        "partitioned_output_count": (C.c_int32, [C.c_void_p, C.POINTER(pg_page), C.c_void_p]),
        "partitioned_output_scatter_to": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
        "ipc_export": (C.c_int32, [C.c_void_p, C.c_void_p]),
        "ipc_import": (C.c_int32, [C.c_void_p, C.POINTER(C.c_void_p)]),
        "ipc_close": (C.c_int32, [C.c_void_p]),
        "filter_project_explain": (C.c_int32, [C.POINTER(pg_filter_project_config), C.c_uint64, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
        "hash_aggregation_state_bytes": (C.c_int32, [C.c_void_p, C.c_int64, C.POINTER(C.c_size_t)]),
        "hash_aggregation_export_state": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int64]),
        "hash_aggregation_merge_states": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_size_t, C.c_int64, C.c_int32]),
        "fused_aggregation_get_info": (C.c_int32, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.c_char_p, C.c_size_t]),
        "fused_aggregation_explain": (C.c_int32, [C.POINTER(pg_fused_aggregation_config), C.c_uint64, C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.c_int32),
                                                  C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
        "fused_page_get_info": (C.c_int32, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_char_p, C.c_size_t]),
        "fused_page_explain": (C.c_int32, [C.POINTER(pg_fused_page_config), C.c_uint64, C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.c_int32),
                                           C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
        "operator_factory_create": (C.c_int32, [C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
        "operator_factory_create_operator": (C.c_int32, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
        "operator_factory_no_more_operators": (C.c_int32, [C.c_void_p]),
        "operator_factory_duplicate": (C.c_int32, [C.c_void_p, C.POINTER(C.c_void_p)]),
        "operator_factory_get_created_count": (C.c_int32, [C.c_void_p, C.POINTER(C.c_int64)]),
        "operator_factory_destroy": (C.c_int32, [C.c_void_p]),
        "jit_compile_check": (C.c_int32, [C.c_char_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_int64)]),
    }

    def __init__(self, lib: C.CDLL, prefix: str, has_device: bool):
        self.lib = lib
        self.prefix = prefix
        self.has_device = has_device
        sigs = dict(self._SIGNATURES)
        if has_device:
            sigs.update(self._DEVICE_SIGNATURES)
        for name, (restype, argtypes) in sigs.items():
            fn = getattr(lib, prefix + name)   # AttributeError if the library does not export a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
            setattr(self, name, fn)

    def check(self, status: int) -> None:
        if status != PG_OK:
            msg = self.last_error()
            raise PrestoException(status, msg.decode() if msg else "")
