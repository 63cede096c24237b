/*
 * presto_gpu.h — C ABI of libprestogpu.so: the B200 (sm_100a) operator pipeline that sits behind
 * prestodb/presto's Operator / OperatorFactory and Page / Block interfaces.
 *
 * Every entry point is `extern "C"`, takes plain pointers and sizes, and returns an int32 status
 * (PG_OK == 0).  No torch / C++ types cross this boundary.  A JNI shim binds these
 * 1:1; the standalone drivers in this repo (pytest, bench.py) bind them with ctypes.
 *
 * Reference interfaces each group of functions replaces (paths relative to the presto checkout):
 *   MB = presto-main-base/src/main/java/com/facebook/presto/
 *   PC = presto-common/src/main/java/com/facebook/presto/common/
 *
 *   pg_block / pg_page ............ PC/block/Block.java:47-403, PC/Page.java:45-498,
 *                                   LongArrayBlock.java:37-52, IntArrayBlock.java:43-47,
 *                                   ByteArrayBlock.java:50-54, VariableWidthBlock.java:53-58,
 *                                   DictionaryBlock.java:59-64, RunLengthEncodedBlock.java:48-49
 *   pg_operator_* ................. MB/operator/Operator.java:20-102 (needsInput/addInput/getOutput/
 *                                   finish/isFinished/close), MB/operator/OperatorFactory.java:18-50
 *   pg_hash_aggregation_create .... MB/operator/HashAggregationOperator.java:56-725 (+ GroupByHash.java:46-59,
 *                                   BigintGroupByHash.java:44-437, MultiChannelGroupByHash.java:56-872,
 *                                   aggregation/builder/InMemoryHashAggregationBuilder.java:204-352)
 *   pg_group_by_hash_* ............ MB/operator/GroupByHash.java:61-89 (addPage/getGroupIds/getGroupCount/
 *                                   appendValuesTo/contains)
 *   pg_join_bridge_*, pg_hash_builder_create, pg_lookup_join_create
 *                                   MB/operator/HashBuilderOperator.java:55-692, PagesIndex.java:202-226,
 *                                   PagesHash.java:55-181, ArrayPositionLinks.java:38-100, JoinHash.java:81-124,
 *                                   JoinProbe.java:74-119, LookupJoinOperator.java:212-343,481-616,
 *                                   JoinBridgeManager / LookupSourceFactory hand-off (LookupJoinOperator.java:127)
 *   pg_partitioned_output_create .. MB/operator/repartition/OptimizedPartitionedOutputOperator.java:167-179,384-824,
 *                                   PartitionedOutputOperator.java:394-500, MB/operator/HashGenerator.java:22-29,
 *                                   MB/operator/InterpretedHashGenerator.java:70-79, PrecomputedHashGenerator.java:31-35
 *   pg_filter_project_create ...... MB/operator/FilterAndProjectOperator.java:88-99,
 *                                   MB/operator/project/PageProcessor.java:112-148,180-343, PageFilter.java:27-50,
 *                                   SelectedPositions.java:23-86
 *   pg_hash_* (scalar helpers) .... PC/type/AbstractLongType.java:137-141, MB/operator/scalar/CombineHashFunction.java:26-31,
 *                                   MB/operator/PagesHash.java:236-252
 */
#ifndef PRESTO_GPU_H
#define PRESTO_GPU_H

#ifndef __CUDACC_RTC__   /* NVRTC (fused pipeline kernels compile this header for its enums): the including file provides the integer types */
#include <stddef.h>
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------
 * Status codes.  Mirror the PrestoException(StandardErrorCode) / IllegalStateException split of the
 * reference (SURVEY §8b "Errors"); the JNI shim maps them back.
 * ---------------------------------------------------------------------------------------------- */
enum {
    PG_OK = 0,
    PG_ERR_INVALID_ARGUMENT = 1,      /* checkArgument(...) failures */
    PG_ERR_INVALID_STATE = 2,         /* checkState(...): addInput when !needsInput, etc. */
    PG_ERR_UNSUPPORTED = 3,           /* type / encoding / shape not on the GPU path (no CPU fallback) */
    PG_ERR_OUT_OF_MEMORY = 4,         /* cudaMalloc failure → EXCEEDED_LOCAL_MEMORY_LIMIT */
    PG_ERR_INSUFFICIENT_RESOURCES = 5,/* "Size of hash table cannot exceed 1 billion entries" BigintGroupByHash.java:273-276 */
    PG_ERR_NUMERIC_OVERFLOW = 6,      /* NUMERIC_VALUE_OUT_OF_RANGE: bigint addition overflow (BigintOperators.java:68-78) */
    PG_ERR_CUDA = 7,                  /* CUDA runtime error; message in pg_last_error */
    PG_ERR_NO_DEVICE = 8              /* no CUDA device: the product path fails loudly, there is no CPU fallback */
};

/* ------------------------------------------------------------------------------------------------
 * Types and block encodings (the subset of PC/type + PC/block on the hot path).
 * ---------------------------------------------------------------------------------------------- */
typedef enum {
    PG_TYPE_BIGINT = 1,   /* LongArrayBlock, 8 B */
    PG_TYPE_INTEGER = 2,  /* IntArrayBlock, 4 B */
    PG_TYPE_DOUBLE = 3,   /* LongArrayBlock holding doubleToLongBits, 8 B */
    PG_TYPE_DATE = 4,     /* IntArrayBlock (days since epoch), 4 B */
    PG_TYPE_BOOLEAN = 5,  /* ByteArrayBlock, 1 B */
    PG_TYPE_VARCHAR = 6,  /* VariableWidthBlock: bytes + int32 offsets[n+1] (VARBINARY travels in the same block: the intermediate state of
                             sum(decimal)) */
    PG_TYPE_LONG_DECIMAL = 7 /* Int128ArrayBlock (PC/block/Int128ArrayBlock.java), 16 B per position: decimal(19..38) unscaled values in
                                UnscaledDecimal128Arithmetic's sign-magnitude form — long 0 = low 64 bits of the magnitude, long 1 = high 63
                                bits | sign << 63 (UnscaledDecimal128Arithmetic.java:46-58,172-180).  Short decimals (precision <= 18) are
                                plain longs: PG_TYPE_BIGINT. */
} pg_type;

typedef enum {
    PG_ENC_FLAT = 0,
    PG_ENC_DICTIONARY = 1, /* DictionaryBlock: ids[n] into a flat dictionary */
    PG_ENC_RLE = 2         /* RunLengthEncodedBlock: one value repeated n times */
} pg_encoding;

typedef enum {
    PG_MEM_HOST = 0,   /* pointers are host addresses (JVM heap via GetPrimitiveArrayCritical, or pinned buffers) */
    PG_MEM_DEVICE = 1  /* pointers are device addresses on the operator's GPU: chained GPU operators never round-trip */
} pg_memory;

/* One column.  Equivalent of a presto Block after arrayOffset has been applied to the pointers. */
typedef struct pg_block {
    int32_t type;             /* pg_type */
    int32_t encoding;         /* pg_encoding */
    int64_t position_count;   /* rows in this block */
    const void* values;       /* FLAT: position_count values (VARCHAR: the byte slice).
                                 DICTIONARY/RLE: the dictionary's / the single value's storage */
    const uint8_t* is_null;   /* boolean[] valueIsNull, one byte per position (of the dictionary for DICTIONARY/RLE); NULL = no nulls */
    const int32_t* offsets;   /* VARCHAR only: position_count+1 (or dictionary_size+1) offsets into values */
    const int32_t* ids;       /* DICTIONARY only: position_count dictionary indices */
    int64_t dictionary_size;  /* DICTIONARY: entries in the dictionary; RLE: 1; FLAT: 0 */
} pg_block;

/* A batch of rows.  Equivalent of presto Page. */
typedef struct pg_page {
    int64_t position_count;
    int32_t channel_count;
    int32_t memory;           /* pg_memory: where every pointer of every block lives */
    pg_block* blocks;         /* channel_count blocks (the array itself is always host memory) */
} pg_page;

/* ------------------------------------------------------------------------------------------------
 * Library / device
 * ---------------------------------------------------------------------------------------------- */
/* Library version string "presto_b200 <semver> sm_100a". */
const char* pg_version(void);
/* Number of visible CUDA devices; returns PG_ERR_NO_DEVICE (and *count = 0) when there is none. */
int32_t pg_device_count(int32_t* count);
/* Bind the calling thread's subsequent pg_* calls to a device (cudaSetDevice). */
int32_t pg_set_device(int32_t device);
/* Thread-local message of the last failing call on this thread. */
const char* pg_last_error(void);
/* Pinned host buffers for the JNI side (direct ByteBuffers) and the benchmarks. */
int32_t pg_host_alloc(size_t bytes, void** out);
int32_t pg_host_free(void* p);
/* Device buffers (used by tests / bench to build device-resident pages without torch). */
int32_t pg_device_alloc(size_t bytes, void** out);
int32_t pg_device_free(void* p);
int32_t pg_memcpy_h2d(void* dst_device, const void* src_host, size_t bytes);
int32_t pg_memcpy_d2h(void* dst_host, const void* src_device, size_t bytes);
int32_t pg_device_synchronize(void);
/* Give the device / pinned blocks the library keeps cached for reuse back to the driver (operators allocate through a caching
 * allocator; nothing an operator currently owns is touched). */
int32_t pg_release_cached_memory(void);

/* ------------------------------------------------------------------------------------------------
 * Scalar hash helpers (host side; bit-exact with the reference; used by tests and by callers that
 * precompute $hashvalue channels the way HashGenerationOptimizer.java:975-987 would).
 * ---------------------------------------------------------------------------------------------- */
int64_t pg_hash_bigint(int64_t value);                 /* AbstractLongType.hash  PC/type/AbstractLongType.java:137-141 */
int64_t pg_hash_combine(int64_t previous, int64_t h);  /* CombineHashFunction.getHash  :26-31 */
int64_t pg_hash_murmur3_fmix64(int64_t x);             /* fastutil HashCommon.murmurHash3(long) = PagesHash.java:236-252 */
int32_t pg_hash_partition(int64_t raw_hash, int32_t partition_count); /* HashGenerator.getPartition  :22-29 */
int64_t pg_hash_xxhash64(const void* data, size_t len);/* airlift XxHash64.hash(Slice), seed 0 */
int64_t pg_hash_double(double value);                  /* DoubleType.hash → TypeUtils.doubleHashCode  PC/type/TypeUtils.java:232-238 (-0.0 == +0.0, one NaN) */
/* LocalPartitionGenerator.getPartition(rawHash)  MB/operator/exchange/LocalPartitionGenerator.java:43-63:
 * (int) XxHash64.hash(Long.reverse(rawHash)) & (partitionCount - 1); partitionCount must be a power of two (-1 otherwise). */
int32_t pg_hash_local_partition(int64_t raw_hash, int32_t partition_count);

/* ------------------------------------------------------------------------------------------------
 * Operator protocol — MB/operator/Operator.java:20-102.
 * Every pg_operator* is single-threaded by contract (Driver.java:357-359,404), owns one CUDA stream.
 * ---------------------------------------------------------------------------------------------- */
typedef struct pg_operator pg_operator;

typedef struct pg_operator_stats {   /* MB/operator/OperatorStats.java:40 subset */
    int64_t input_positions;
    int64_t input_bytes;
    int64_t output_positions;
    int64_t output_bytes;
    int64_t device_bytes;            /* current device memory held → LocalMemoryContext.setBytes */
    int64_t kernel_launches;         /* CUDA kernels launched by this operator so far */
    int64_t group_count;             /* aggregation: groups; join build: positions indexed */
    int64_t h2d_bytes;
    int64_t d2h_bytes;
    double  gpu_ms;                  /* device time accumulated by event-bracketed sections (if enabled) */
    int64_t kernel_paths;            /* aggregation: bit mask of pg_kernel_path — which kernels this operator's rows went through */
} pg_operator_stats;

typedef enum {                       /* pg_operator_stats.kernel_paths */
    PG_PATH_AGG_FAST = 1,            /* hashAggFastKernel tiers 0/1/2 (few groups) */
    PG_PATH_AGG_WARP = 2,            /* per-warp accumulator sets (11..40 groups) */
    PG_PATH_AGG_MID = 4,             /* one shared table per SM (.. ~3000 groups) */
    PG_PATH_AGG_DIRECT = 8,          /* straight to the HBM table, L2-resident slot ranges */
    PG_PATH_AGG_GENERIC = 16,        /* hashAggKernel<W, MAXV>: NULLs, masks, bigint sums, wide keys */
    PG_PATH_FUSED_TIERS = 32,        /* generated fused pipeline kernel, on-chip tiers */
    PG_PATH_FUSED_DIRECT = 64,       /* generated fused pipeline kernel, HBM table */
    PG_PATH_UNFUSED_CHAIN = 128,     /* fused operator that ran its fragment operator by operator (fragment not fusable for that page) */
    PG_PATH_FUSED_TMA = 256,         /* (with FUSED_TIERS / FUSED_DIRECT) eager columns staged by a cp.async.bulk (TMA) ring */
    PG_PATH_INTERNED_KEYS = 512,     /* VARCHAR group keys without a declared length of <= 7 bytes: replaced by value ids (VarcharInterner) first */
    PG_PATH_DICTIONARY_KEYS = 1024,  /* (with INTERNED_KEYS) a DictionaryBlock / RLE key channel: its dictionary was interned, rows only mapped */
    PG_PATH_FP_JIT = 2048,           /* FilterAndProject: generated PTX (driver JIT) evaluated a page */
    PG_PATH_FP_INTERPRETER = 4096    /* FilterAndProject: the interpreter kernels evaluated a page (program outside the PTX generator's set) */
} pg_kernel_path;

/* Options common to all operators. */
typedef struct pg_operator_options {
    int32_t device;            /* CUDA device ordinal */
    int32_t output_memory;     /* pg_memory of pages returned by get_output */
    void*   cuda_stream;       /* cudaStream_t to run on; NULL = operator creates its own non-blocking stream */
    int64_t max_output_rows;   /* rows per output page; <=0 = default (1<<20) */
} pg_operator_options;

int32_t pg_operator_needs_input(pg_operator* op, int32_t* out);
/* Operator.isBlocked() (MB/operator/Operator.java:32; default NOT_BLOCKED): *out = 1 while the operator waits for something another
 * Driver produces — a LookupJoin / semi join / fused fragment whose LookupSource is still being built (LookupJoinOperator.java:198-210:
 * the lookupSourceProviderFuture) — 0 otherwise.  The JNI shim turns this into the ListenableFuture the Driver polls. */
int32_t pg_operator_is_blocked(pg_operator* op, int32_t* out);
/* Borrow `page` for the duration of the call (host pages are staged to HBM before return;
 * device pages are consumed on the operator's stream and must stay valid until the next
 * pg_operator_* call that synchronises: get_output / finish / stats). */
int32_t pg_operator_add_input(pg_operator* op, const pg_page* page);
/* Returns a page with position_count == 0 and channel_count == 0 where Java would return null.
 * Buffers are owned by the operator and valid until the next get_output / destroy. */
int32_t pg_operator_get_output(pg_operator* op, pg_page* out);
int32_t pg_operator_finish(pg_operator* op);
int32_t pg_operator_is_finished(pg_operator* op, int32_t* out);
int32_t pg_operator_get_stats(pg_operator* op, pg_operator_stats* out);
int32_t pg_operator_destroy(pg_operator* op);   /* Operator.close() */

/* ------------------------------------------------------------------------------------------------
 * HashAggregationOperator
 * ---------------------------------------------------------------------------------------------- */
typedef enum {
    PG_AGG_COUNT_STAR = 0,   /* CountAggregation.java:34-50 */
    PG_AGG_COUNT_COLUMN = 1, /* CountColumn: count of non-null inputs */
    PG_AGG_SUM_DOUBLE = 2,   /* DoubleSumAggregation.java:33-61 (NULL when no non-null input) */
    PG_AGG_SUM_BIGINT = 3,   /* LongSumAggregation.java:33-56 (Math.addExact overflow → PG_ERR_NUMERIC_OVERFLOW) */
    PG_AGG_AVG_DOUBLE = 4,   /* AverageAggregations.java:41-66 */
    PG_AGG_AVG_BIGINT = 5,   /* AverageAggregations.java:34-39 */
    PG_AGG_MIN_DOUBLE = 6, PG_AGG_MAX_DOUBLE = 7,
    PG_AGG_MIN_BIGINT = 8, PG_AGG_MAX_BIGINT = 9,
    /* DecimalSumAggregation.java:60-185: sum(decimal(p,s)) -> decimal(38,s).  SHORT: the input channel holds unscaled longs (PG_TYPE_BIGINT
     * blocks, inputShortDecimal :145-153); LONG: PG_TYPE_LONG_DECIMAL blocks (inputLongDecimal :155-163).  Output PG_TYPE_LONG_DECIMAL, NULL
     * without a non-null input; PG_ERR_NUMERIC_OVERFLOW when the sum leaves decimal(38) (outputLongDecimal :179-185).  Intermediate state
     * (PARTIAL output / FINAL input) = one VARBINARY of 3 longs {overflow, low, high} per group
     * (LongDecimalWithOverflowStateSerializer.java:37-49), NULL for a group without input. */
    PG_AGG_SUM_SHORT_DECIMAL = 10, PG_AGG_SUM_LONG_DECIMAL = 11
} pg_aggregate_function;

typedef enum {              /* AggregationNode.Step */
    PG_STEP_SINGLE = 0,
    PG_STEP_PARTIAL = 1,    /* output = intermediate states (sum, count columns) */
    PG_STEP_FINAL = 2       /* input = intermediate states */
} pg_aggregation_step;

typedef struct pg_aggregate {
    int32_t function;       /* pg_aggregate_function */
    int32_t input_channel;  /* -1 for count(*) */
    int32_t mask_channel;   /* BOOLEAN channel or -1 (AccumulatorCompiler mask handling) */
    int32_t reserved;
} pg_aggregate;

typedef struct pg_hash_aggregation_config {
    int32_t n_group_channels;
    const int32_t* group_channels;      /* input channel of each group-by key */
    const int32_t* group_types;         /* pg_type of each key */
    const int32_t* group_max_lengths;   /* VARCHAR(n) group channels: n in 1..7 packs the value into the key word; 0 (unbounded VARCHAR) or n > 7: the
                                           channel's slices are interned into value ids on the device first (PG_PATH_INTERNED_KEYS).  Other types: 0 */
    int32_t hash_channel;               /* precomputed $hashvalue BIGINT channel or -1 */
    int32_t n_aggregates;
    const pg_aggregate* aggregates;
    int32_t step;                       /* pg_aggregation_step */
    int32_t expected_groups;            /* expectedSize of GroupByHash.createGroupByHash */
    pg_operator_options options;
} pg_hash_aggregation_config;

/* Output page layout = HashAggregationOperator's: group keys [, $hashvalue] , one column per aggregate,
 * rows in ascending group id = first-appearance order (InMemoryHashAggregationBuilder.java:326-352,400-403). */
int32_t pg_hash_aggregation_create(const pg_hash_aggregation_config* config, pg_operator** out);

/* PARTIAL -> FINAL hand-off between GPUs without intermediate pages.  The reference ships a PARTIAL operator's groups as pages of
 * intermediate states through an exchange to the FINAL operator (HashAggregationOperator.java:451-501 getOutput -> ... -> addInput
 * of the FINAL step: GroupedAccumulator.addIntermediate, InMemoryHashAggregationBuilder.java:204-260).  Here the state of a source
 * operator (hash aggregation or fused aggregation) is compacted into a fixed-capacity buffer in ITS GPU's memory, the buffers travel
 * in one collective (all-gather over NVLink), and the destination — an operator created from the same configuration, step aside —
 * merges them into its own group table: packed keys + physical accumulators, i.e. exactly the intermediate states.
 *   state_bytes   size of one state buffer able to hold max_groups groups
 *   export_state  asynchronous on the operator's stream; the operator itself is unchanged (it may go on aggregating or produce output);
 *                 more than max_groups groups -> the state is marked overflowed and merge_states rejects it (PG_ERR_INSUFFICIENT_RESOURCES)
 *   merge_states  `count` states `stride` bytes apart (state j at device_buffers + j * stride), written on this operator's stream or
 *                 on work it has been ordered after; all-zero states are skipped (an unused slot of the gather).  Source j's groups
 *                 order after this operator's own input and after sources < j (first-appearance order = the order in which the
 *                 FINAL operator would have received the sources' pages); order_base + j tags the source (order_base >= 1 when the
 *                 operator has input of its own).  DOUBLE sums are added source by source: reproducible.
 *                 PG_ERR_INVALID_ARGUMENT for a state of a differently planned operator. */
int32_t pg_hash_aggregation_state_bytes(pg_operator* op, int64_t max_groups, size_t* bytes);
int32_t pg_hash_aggregation_export_state(pg_operator* op, void* device_buffer, size_t capacity, int64_t max_groups);
int32_t pg_hash_aggregation_merge_states(pg_operator* op, const void* device_buffers, int32_t count, size_t stride, int64_t max_groups, int32_t order_base);

/* ------------------------------------------------------------------------------------------------
 * GroupByHash (standalone) — MB/operator/GroupByHash.java:61-89
 * ---------------------------------------------------------------------------------------------- */
typedef struct pg_group_by_hash pg_group_by_hash;

typedef struct pg_group_by_hash_config {
    int32_t n_channels;
    const int32_t* channels;
    const int32_t* types;
    const int32_t* max_lengths;
    int32_t hash_channel;      /* inputHashChannel or -1 */
    int32_t expected_size;
    pg_operator_options options;
} pg_group_by_hash_config;

int32_t pg_group_by_hash_create(const pg_group_by_hash_config* config, pg_group_by_hash** out);
/* addPage: insert the page's keys. */
int32_t pg_group_by_hash_add_page(pg_group_by_hash* h, const pg_page* page);
/* getGroupIds: insert + return the dense group id of every row (ids in first-appearance order).
 * out_group_ids: position_count int64 in page->memory's space (GroupByIdBlock is a BIGINT block). */
int32_t pg_group_by_hash_get_group_ids(pg_group_by_hash* h, const pg_page* page, int64_t* out_group_ids);
int32_t pg_group_by_hash_get_group_count(pg_group_by_hash* h, int64_t* out);
/* contains(position, page): out_flags[position] = 1/0 for every row of the page (host or device per page->memory). */
int32_t pg_group_by_hash_contains(pg_group_by_hash* h, const pg_page* page, uint8_t* out_flags);
/* appendValuesTo for groups [0, groupCount): returns a page of the key columns [+ raw hash] in group-id order. */
int32_t pg_group_by_hash_get_keys(pg_group_by_hash* h, pg_page* out);
int32_t pg_group_by_hash_destroy(pg_group_by_hash* h);

/* ------------------------------------------------------------------------------------------------
 * Hash join: HashBuilderOperator → (JoinBridge) → LookupJoinOperator
 * ---------------------------------------------------------------------------------------------- */
typedef struct pg_join_bridge pg_join_bridge;   /* JoinBridgeManager / LookupSourceFactory */

typedef enum {       /* LookupJoinOperators.JoinType (MB/operator/LookupJoinOperators.java:30-36) */
    PG_JOIN_INNER = 0,
    PG_JOIN_PROBE_OUTER = 1,
    PG_JOIN_LOOKUP_OUTER = 2,  /* RIGHT join: the probe behaves like INNER; the bridge tracks visited build positions */
    PG_JOIN_FULL_OUTER = 3     /* the probe behaves like PROBE_OUTER; the bridge tracks visited build positions */
} pg_join_type;

int32_t pg_join_bridge_create(int32_t device, pg_join_bridge** out);
int32_t pg_join_bridge_destroy(pg_join_bridge* b);
/* LookupSource statistics once the build finished (PagesHash.getPositionCount / getHashCollisions ... ). */
int32_t pg_join_bridge_get_position_count(pg_join_bridge* b, int64_t* out);

typedef struct pg_hash_builder_config {
    int32_t n_channels;                 /* channels of the build pages */
    const int32_t* types;               /* pg_type per channel */
    int32_t n_hash_channels;
    const int32_t* hash_channels;       /* join key channels (BIGINT / INTEGER / DATE / VARCHAR: slices are interned into value ids when the build finishes,
                                           probes look theirs up) */
    int32_t n_output_channels;
    const int32_t* output_channels;     /* build channels appended to join output (any pg_type but LONG_DECIMAL; a VARCHAR channel keeps the
                                           position-ordered hash-table LookupSource) */
    int32_t precomputed_hash_channel;   /* $hashvalue channel or -1 */
    int64_t expected_positions;
    pg_operator_options options;
    /* JoinFilterFunction (MB/operator/JoinFilterFunction.java; compiled by MB/sql/gen/JoinFilterFunctionCompiler.java:97-250 and handed to
     * the BUILD side: HashBuilderOperator's filterFunctionFactory -> JoinHash.isJoinPositionEligible :113-118): the non-equi part of the
     * join condition, a BOOLEAN expression whose inputs are the build pages' channels [0, n_channels) followed by the probe page's
     * channels (LocalExecutionPlanner.createJoinSourcesLayout); only TRUE keeps a candidate (LookupJoinOperator.joinCurrentPosition
     * :567-586; a PROBE_OUTER row without an eligible candidate is emitted once with NULL build columns).  Only nodes / filter_node /
     * n_input_channels / input_types of the filter's pg_filter_project_config, declared further down, are read.  NULL = no filter function. */
    const struct pg_filter_project_config* filter;
} pg_hash_builder_config;

int32_t pg_hash_builder_create(const pg_hash_builder_config* config, pg_join_bridge* bridge, pg_operator** out);

typedef struct pg_lookup_join_config {
    int32_t join_type;                  /* pg_join_type */
    int32_t n_channels;                 /* channels of the probe pages */
    const int32_t* types;
    int32_t n_hash_channels;
    const int32_t* hash_channels;       /* probe join-key channels */
    int32_t n_output_channels;
    const int32_t* output_channels;     /* probe channels emitted (probeOutputChannels; VARCHAR slices are copied like LookupJoinPageBuilder does) */
    int32_t precomputed_hash_channel;
    pg_operator_options options;
} pg_lookup_join_config;

/* Output page = probe output channels followed by build output channels, rows in probe order, matches of one
 * probe row in PositionLinks order (descending build position; PagesHash.java:108-123, ArrayPositionLinks.java:45-50). */
int32_t pg_lookup_join_create(const pg_lookup_join_config* config, pg_join_bridge* bridge, pg_operator** out);

/* Build-outer joins (RIGHT / FULL OUTER).  OuterLookupSource.java:36-203: every build position a probe appends is marked visited
 * (appendTo :100-105); LookupOuterOperator.java:109-225 is the SOURCE operator that, once every probe operator is done, emits the
 * positions nobody visited — ascending build position — as NULL probe channels + the build output channels.
 *   pg_join_bridge_set_outer                 before the build finishes (PartitionedLookupSourceFactory(outer = true)): the LookupSource keeps
 *                                            its position-ordered columns and a visited map; LOOKUP_OUTER / FULL_OUTER probes need it
 *   pg_join_bridge_no_more_probe_operators   the probe OperatorFactory is closed (JoinBridgeManager.probeOperatorFactoryClosed); together
 *                                            with every probe operator having finished / been destroyed this completes the
 *                                            outerPositionsFuture: the outer operator stops being blocked
 *   pg_lookup_outer_create                   needs_input is always false; get_output returns nothing while blocked, then pages of at most
 *                                            max_output_rows rows; finish() terminates the output (LookupOuterOperator.finish :155-160). */
typedef struct pg_lookup_outer_config {
    int32_t n_probe_output_channels;
    const int32_t* probe_output_types;   /* pg_type of the probe's output channels (all NULL in the output) */
    pg_operator_options options;
} pg_lookup_outer_config;
int32_t pg_join_bridge_set_outer(pg_join_bridge* bridge);
int32_t pg_join_bridge_no_more_probe_operators(pg_join_bridge* bridge);
int32_t pg_lookup_outer_create(const pg_lookup_outer_config* config, pg_join_bridge* bridge, pg_operator** out);

/* HashSemiJoinOperator (MB/operator/HashSemiJoinOperator.java:142-192) over the ChannelSet a SetBuilderOperator builds
 * (MB/operator/SetBuilderOperator.java, ChannelSet.java:33-77).  The set is a pg_join_bridge fed by a
 * pg_hash_builder whose only hash channel is the set channel and which has no output channels.  Output page = every input channel
 * + one BOOLEAN channel: NULL probe value → NULL (false when the set has no rows at all); otherwise contained → true, not
 * contained → NULL when the set holds a NULL, else false.  Input channels: any pg_type but LONG_DECIMAL (VARCHAR channels pass through as
 * copied slices; a VARCHAR join channel is looked up in the set's interned values). */
typedef struct pg_semi_join_config {
    int32_t n_channels;
    const int32_t* types;
    int32_t probe_join_channel;
    int32_t precomputed_hash_channel;   /* accepted and ignored (the slot function is recomputed from the key) */
    pg_operator_options options;
} pg_semi_join_config;
int32_t pg_hash_semi_join_create(const pg_semi_join_config* config, pg_join_bridge* set, pg_operator** out);

/* ------------------------------------------------------------------------------------------------
 * PartitionedOutputOperator (row repartition)
 * ---------------------------------------------------------------------------------------------- */
typedef enum {
    PG_PARTITION_HASH_GENERATOR = 0, /* HashGenerator.getPartition (MB/operator/HashGenerator.java:22-29): between tasks / GPUs */
    PG_PARTITION_LOCAL = 1,          /* LocalPartitionGenerator (MB/operator/exchange/LocalPartitionGenerator.java:43-63): LocalExchange and
                                        PartitionedLookupSource inside one task; partition_count must be a power of two */
    PG_PARTITION_ROUND_ROBIN = 2     /* RoundRobinBucketFunction (MB/sql/planner/SystemPartitioningHandle.java:282-305): row i of the operator's
                                        input goes to bucket i mod bucket_count (the counter runs on across pages); no partition channels */
} pg_partition_function;

typedef struct pg_partitioned_output_config {
    int32_t n_channels;
    const int32_t* types;
    int32_t partition_count;
    int32_t n_partition_channels;
    const int32_t* partition_channels;  /* hashed with InterpretedHashGenerator unless hash channel given */
    int32_t precomputed_hash_channel;   /* PrecomputedHashGenerator channel or -1 */
    int32_t replicates_any_row;         /* OptimizedPartitionedOutputOperator.java:504-512 */
    int32_t null_channel;               /* replicate rows whose null_channel is null to every partition; -1 = none */
    int32_t partition_function;         /* pg_partition_function */
    /* BucketPartitionFunction (MB/operator/BucketPartitionFunction.java:25-54): the function above yields a BUCKET in [0, bucket_count) and
     * bucket_to_partition[bucket] names the partition (several buckets may share one).  NULL table: bucket_count is ignored, the bucket is
     * the partition (the identity table of SystemPartitioningHandle's hash partitioning with one bucket per partition). */
    int32_t bucket_count;
    const int32_t* bucket_to_partition; /* [bucket_count], values in [0, partition_count) */
    pg_operator_options options;
} pg_partitioned_output_config;

/* get_output returns, per input page, ONE page holding all partitions back to back (partition p occupies rows
 * [offsets[p], offsets[p+1]) ), rows inside a partition in input order — what PartitionBuffer.appendData would have
 * appended.  pg_partitioned_output_get_offsets gives the partition_count+1 row offsets of the page returned last. */
int32_t pg_partitioned_output_create(const pg_partitioned_output_config* config, pg_operator** out);
int32_t pg_partitioned_output_get_offsets(pg_operator* op, int64_t* out_offsets /* partition_count+1, host */);

/* Fused repartition + exchange across the GPUs of one box (replaces PartitionedOutputBuffer → HTTP → ExchangeClient,
 * MB/execution/buffer/PartitionedOutputBuffer.java:174-181, MB/operator/ExchangeClient.java:71):
 *   pg_partitioned_output_count      pass 1 over `page` (partition ids + histograms); returns the rows per partition.  The page must
 *                                    stay valid until the matching scatter_to call.
 *   pg_partitioned_output_scatter_to pass 2: partition p's rows are written, in input order, to row offset dst_row_offsets[p] of the
 *                                    destination columns dst_values[p * n_channels + c] (device addresses that may live on a PEER GPU,
 *                                    e.g. opened with pg_ipc_import) — the scatter IS the transfer, one NVLink store per value.
 *                                    dst_nulls entries may be NULL for channels without a null map on any rank.
 * The caller exchanges the counts (a P x P int64 all-gather) between the two calls and synchronises the ranks afterwards. */
int32_t pg_partitioned_output_count(pg_operator* op, const pg_page* page, int64_t* out_counts /* partition_count, host */);
int32_t pg_partitioned_output_scatter_to(pg_operator* op, void* const* dst_values, uint8_t* const* dst_nulls, const int64_t* dst_row_offsets);

/* CUDA IPC plumbing for the above: export a device allocation made with pg_device_alloc, map a peer's allocation. */
#define PG_IPC_HANDLE_BYTES 64
int32_t pg_ipc_export(const void* device_ptr, uint8_t* handle /* PG_IPC_HANDLE_BYTES */);
int32_t pg_ipc_import(const uint8_t* handle, void** out);
int32_t pg_ipc_close(void* imported);

/* ------------------------------------------------------------------------------------------------
 * FilterAndProjectOperator / PageProcessor
 * A RowExpression tree is flattened to a register program (one instruction per node).
 * ---------------------------------------------------------------------------------------------- */
typedef enum {
    PG_EX_INPUT = 0,      /* a = channel */
    PG_EX_CONST_BIGINT = 1, PG_EX_CONST_DOUBLE = 2, PG_EX_CONST_BOOLEAN = 3, PG_EX_CONST_NULL = 4, /* imm */
    PG_EX_CONST_VARCHAR = 5, /* a VARCHAR literal: imm_i = address of its bytes (host memory, copied when the operator is created),
                                reserved = its length.  VARCHAR-typed nodes are INPUT or CONST_VARCHAR; they may feed the comparisons
                                (VarcharOperators.java:54-105: Slice.equals / Slice.compareTo = unsigned bytes, then length), BETWEEN,
                                IS_NULL and HASH (XxHash64 of the slice); an INPUT may also be a projection (pass-through) */
    PG_EX_ADD = 10, PG_EX_SUB = 11, PG_EX_MUL = 12, PG_EX_DIV = 13, PG_EX_NEG = 14,
    PG_EX_EQ = 20, PG_EX_NE = 21, PG_EX_LT = 22, PG_EX_LE = 23, PG_EX_GT = 24, PG_EX_GE = 25,
    PG_EX_AND = 30, PG_EX_OR = 31, PG_EX_NOT = 32, PG_EX_IS_NULL = 33,
    PG_EX_CAST_DOUBLE = 40, PG_EX_CAST_BIGINT = 41,
    PG_EX_BETWEEN = 50,   /* a BETWEEN b AND c */
    PG_EX_HASH = 60,      /* type hash of a (TypeUtils.hashPosition; NULL → 0) */
    PG_EX_COMBINE_HASH = 61 /* 31*a + b  (CombineHashFunction) */
} pg_expr_op;

typedef struct pg_expr_node {
    int32_t op;        /* pg_expr_op */
    int32_t type;      /* pg_type of the result */
    int32_t a, b, c;   /* operand node indices (or channel for INPUT); -1 unused */
    int32_t reserved;
    int64_t imm_i;     /* CONST_BIGINT / CONST_BOOLEAN */
    double  imm_d;     /* CONST_DOUBLE */
} pg_expr_node;

typedef struct pg_filter_project_config {
    int32_t n_input_channels;
    const int32_t* input_types;
    int32_t n_nodes;
    const pg_expr_node* nodes;       /* topologically ordered: operands precede users */
    int32_t filter_node;             /* node index of the BOOLEAN filter, or -1 = select all */
    int32_t n_projections;
    const int32_t* projection_nodes; /* node index per output channel */
    pg_operator_options options;
} pg_filter_project_config;

int32_t pg_filter_project_create(const pg_filter_project_config* config, pg_operator** out);
/* SelectedPositions of the last page filtered (PageFilter.positionsArrayToSelectedPositions, PageFilter.java:27-50):
 * is_list = 0 → positionsRange(0, size) (none or all rows selected), 1 → positionsList of `size` positions. */
int32_t pg_filter_project_get_last_selection(pg_operator* op, int32_t* is_list, int64_t* size);
/* The PTX the operator would hand to the driver JIT for this program when the input channels in `nullable_inputs` (bit i =
 * channel i) carry a null map — the counterpart of dumping the bytecode PageFunctionCompiler generates
 * (MB/sql/gen/PageFunctionCompiler.java:195-861).  Needs no GPU.  *needed = bytes including the terminating NUL; the text is
 * copied when `capacity` suffices.  PG_ERR_UNSUPPORTED when the program runs on the interpreter kernels instead. */
int32_t pg_filter_project_explain(const pg_filter_project_config* config, uint64_t nullable_inputs, char* buffer, size_t capacity, size_t* needed);

/* ------------------------------------------------------------------------------------------------
 * Fused pipeline fragment: FilterAndProjectOperator → LookupJoinOperator* → HashAggregationOperator as ONE operator.
 * The reference chains these operators inside one Driver and hands <= 1 MB pages between them in the CPU cache
 * (MB/operator/Driver.java:402-536; planner sites LocalExecutionPlanner.java:1633-1666, 2650-2686, 3544-3640; scan + filter + project
 * are already a single operator there: ScanFilterAndProjectOperator).  On the GPU the intermediate pages would cross HBM, so the
 * Gpu*OperatorFactory for such a chain creates this operator instead: the PageProcessor's expressions, the probes of the joins' dense
 * LookupSources and the aggregation's key packing are generated as one row program, compiled at run time (NVRTC, like the reference
 * compiles them to bytecode) and run in a single pass over the raw input page.  Results are those of the three operators chained:
 * when the fragment cannot be fused for a page (LookupSource with duplicate keys, nullable aggregate inputs, bigint sums, masks, ...)
 * the operator runs the chain operator by operator on the GPU instead (pg_operator_stats.kernel_paths tells which).
 *   page_processor   over the raw input page; its projections are the layout the first join / the aggregation sees
 *   joins[k]         INNER lookup joins, each over the previous layout; output layout = probe output channels, then the
 *                    bridge's build output channels (LookupJoinOperator's page layout)
 *   aggregation      over the last layout; step SINGLE or PARTIAL
 * needs_input is false until every bridge's build has finished (LookupJoinOperator.java:198-210).
 * ---------------------------------------------------------------------------------------------- */
typedef struct pg_fused_aggregation_config {
    pg_filter_project_config page_processor;
    int32_t n_joins;                          /* 0..2 */
    const pg_lookup_join_config* joins;
    pg_join_bridge* const* bridges;
    pg_hash_aggregation_config aggregation;   /* its options are the operator's */
} pg_fused_aggregation_config;
int32_t pg_fused_aggregation_create(const pg_fused_aggregation_config* config, pg_operator** out);
/* Pages that ran fused / operator by operator so far, the registers per thread of the generated kernel (0 before the first fused
 * page) and, when a page was not fused, why (copied into reason when capacity allows). */
int32_t pg_fused_aggregation_get_info(pg_operator* op, int64_t* fused_pages, int64_t* unfused_pages, int32_t* registers, char* reason, size_t capacity);
/* The CUDA C++ the operator would hand to NVRTC for this fragment when the raw input channels in `nullable_inputs` (bit i = channel
 * i) carry a null map — the counterpart of dumping generated bytecode.  Needs no GPU; join build sides are described by
 * `build_types` / `build_nullable` (per join: the bridge's output channel types, since no LookupSource exists yet): payload rows are
 * laid out as HashBuilderOperator would.  PG_ERR_UNSUPPORTED (reason in pg_last_error) when the fragment is not fusable. */
int32_t pg_fused_aggregation_explain(const pg_fused_aggregation_config* config, uint64_t nullable_inputs, const int32_t* const* build_types,
                                     const int32_t* n_build_types, char* buffer, size_t capacity, size_t* needed);
/* The same fragment WITHOUT the aggregation: FilterAndProjectOperator → LookupJoinOperator* (INNER) as one operator whose output pages
 * are the last operator's (rows in input order; layout = last join's probe output channels then its bridge's build output channels, or
 * the projections when there is no join).  One generated row program, launched twice (selection pass, emit pass); fragments outside the
 * fused set (hash-table LookupSource, VARCHAR channels) run operator by operator.  get_output returns one page per input page. */
typedef struct pg_fused_page_config {
    pg_filter_project_config page_processor;
    int32_t n_joins;                          /* 0..2 */
    const pg_lookup_join_config* joins;
    pg_join_bridge* const* bridges;
    pg_operator_options options;
} pg_fused_page_config;
int32_t pg_fused_page_create(const pg_fused_page_config* config, pg_operator** out);
int32_t pg_fused_page_get_info(pg_operator* op, int64_t* fused_pages, int64_t* unfused_pages, char* reason, size_t capacity);
int32_t pg_fused_page_explain(const pg_fused_page_config* config, uint64_t nullable_inputs, const int32_t* const* build_types, const int32_t* n_build_types,
                              char* buffer, size_t capacity, size_t* needed);

/* NVRTC compile check of a CUDA C++ source against the embedded device headers (no GPU needed): cubin size or the compiler log. */
int32_t pg_jit_compile_check(const char* source, char* log_buffer, size_t capacity, int64_t* cubin_bytes);

/* ------------------------------------------------------------------------------------------------
 * OperatorFactory — MB/operator/OperatorFactory.java:19-51.  A factory owns a deep copy of one plan node's configuration (the
 * caller's structs may be freed after create); every Driver asks it for its own operator.
 *   create_operator      OperatorFactory.createOperator(DriverContext) :21 — `options` is the Driver's part (device, stream, output
 *                        memory, page size), NULL = the options of the configuration.  PG_ERR_INVALID_STATE "Factory is already
 *                        closed" after no_more_operators (HashAggregationOperator.java:222).
 *   no_more_operators    :33 (HashAggregationOperator.java:249-252)
 *   duplicate            :50 — a new OPEN factory over the same plan (HashAggregationOperator.java:255-275); join factories share
 *                        the bridge (JoinBridgeManager), they do not copy it
 * `bridge` is required for PG_FACTORY_HASH_BUILDER / LOOKUP_JOIN / HASH_SEMI_JOIN and ignored otherwise (fused configurations carry
 * their bridges).
 * ---------------------------------------------------------------------------------------------- */
typedef struct pg_operator_factory pg_operator_factory;
typedef enum {
    PG_FACTORY_HASH_AGGREGATION = 1,   /* config: pg_hash_aggregation_config */
    PG_FACTORY_FILTER_PROJECT = 2,     /* pg_filter_project_config */
    PG_FACTORY_HASH_BUILDER = 3,       /* pg_hash_builder_config */
    PG_FACTORY_LOOKUP_JOIN = 4,        /* pg_lookup_join_config */
    PG_FACTORY_HASH_SEMI_JOIN = 5,     /* pg_semi_join_config */
    PG_FACTORY_PARTITIONED_OUTPUT = 6, /* pg_partitioned_output_config */
    PG_FACTORY_FUSED_AGGREGATION = 7,  /* pg_fused_aggregation_config */
    PG_FACTORY_FUSED_PAGE = 8          /* pg_fused_page_config */
} pg_operator_kind;
int32_t pg_operator_factory_create(int32_t kind, const void* config, pg_join_bridge* bridge, pg_operator_factory** out);
int32_t pg_operator_factory_create_operator(pg_operator_factory* factory, const pg_operator_options* options, pg_operator** out);
int32_t pg_operator_factory_no_more_operators(pg_operator_factory* factory);
int32_t pg_operator_factory_duplicate(pg_operator_factory* factory, pg_operator_factory** out);
int32_t pg_operator_factory_get_created_count(pg_operator_factory* factory, int64_t* count);
int32_t pg_operator_factory_destroy(pg_operator_factory* factory);

/* ------------------------------------------------------------------------------------------------
 * SerializedPage wire format: what PartitionedOutput hands to an OutputBuffer and an
 * ExchangeClient receives when pages leave the box.
 *   PagesSerde.serialize / deserialize            presto-spi/.../spi/page/PagesSerde.java:67-156
 *   PagesSerdeUtil.writeSerializedPage / read...   presto-spi/.../spi/page/PagesSerdeUtil.java:45-120
 *   block encodings                                presto-common/.../common/block/{LongArray,IntArray,ByteArray,VariableWidth,
 *                                                  RunLength,Dictionary}BlockEncoding.java, EncoderUtil.java:31-100
 * Wire bytes: [positionCount i32][markers u8][uncompressedSize i32][size i32][checksum i64] (21 bytes, little endian) then the raw
 * page: [channelCount i32] and per channel [name length i32][encoding name][positionCount i32][mayHaveNull u8][null bits, MSB first]
 * [non-null values].  With compression = PG_COMPRESSION_LZ4 the raw page travels as ONE LZ4 block (the PageCompressor PagesSerdeFactory.java:75
 * creates: airlift Lz4Compressor, block format) when that makes it at most 0.8 of its size (PagesSerde.java:118-143, MINIMUM_COMPRESSION_RATIO),
 * marker COMPRESSED; any conforming LZ4 block is read (the bytes this encoder produces are valid LZ4, not aircompressor's byte for
 * byte).  Encrypted pages are rejected with PG_ERR_UNSUPPORTED.
 * ---------------------------------------------------------------------------------------------- */
typedef struct pg_pages_serde pg_pages_serde;
typedef struct pg_pages_serde_config {
    int32_t checksum_enabled;      /* PagesSerde.checksumEnabled: CHECKSUMMED marker + CRC32 over data, markers, positionCount, uncompressedSize */
    int32_t compression;           /* PG_COMPRESSION_NONE / PG_COMPRESSION_LZ4: the serde's compressor (pages of either kind are always READ) */
    pg_operator_options options;   /* device; output_memory = where serialized bytes AND deserialized pages are returned; cuda_stream */
} pg_pages_serde_config;
#define PG_COMPRESSION_NONE 0
#define PG_COMPRESSION_LZ4 1
#define PG_SERIALIZED_PAGE_HEADER_BYTES 21
#define PG_PAGE_MARKER_COMPRESSED 1
#define PG_PAGE_MARKER_ENCRYPTED 2
#define PG_PAGE_MARKER_CHECKSUMMED 4

/* The host half of the page checksum: CRC-32 (java.util.zip.CRC32) of a host buffer computed as the serde does — one CRC per
 * `chunk` bytes (0 = 4096), combined with the zero-extension operator.  Needs no GPU; exists so that the combine logic is testable. */
int64_t pg_hash_crc32_chunked(const void* data, size_t len, size_t chunk);
/* The LZ4 block codec of the serde run on the host (the kernels compile the same functions, csrc/lz4_codec.cuh); needs no GPU.
 * compress: returns the compressed size, -1 when `capacity` < pg_lz4_compress_bound(n).  decompress: the decoded size, -1 for a malformed
 * block or one that decodes to more than `capacity` bytes. */
int64_t pg_lz4_compress_bound(int64_t n);
int64_t pg_lz4_compress_host(const void* src, int64_t n, void* dst, int64_t capacity);
int64_t pg_lz4_decompress_host(const void* src, int64_t n, void* dst, int64_t capacity);
int32_t pg_pages_serde_create(const pg_pages_serde_config* config, pg_pages_serde** out);
/* serialize(Page) + writeSerializedPage.  FLAT blocks → LONG_ARRAY / INT_ARRAY / BYTE_ARRAY / VARIABLE_WIDTH, RLE blocks → RLE,
 * DICTIONARY blocks -> DICTIONARY of the compacted dictionary (DictionaryBlockEncoding.writeBlock :34-54 over DictionaryBlock.compact
 * :584-628: referenced entries in order of first reference, remapped ids) + a per-dictionary instance id.
 * *bytes (in options.output_memory) stays valid until the next call on this serde. */
int32_t pg_pages_serde_serialize(pg_pages_serde* serde, const pg_page* page, const uint8_t** bytes, int64_t* size);
/* One SerializedPage per row range [row_offsets[r], row_offsets[r + 1]) of `page`, back to back in one buffer (byte_offsets: n_ranges + 1
 * entries): what OptimizedPartitionedOutputOperator.flush :774-797 enqueues, one page per partition, when `page` / `row_offsets` are a
 * pg_partitioned_output operator's output page and pg_partitioned_output_get_offsets.  Empty ranges give empty (0-position) pages. */
int32_t pg_pages_serde_serialize_ranges(pg_pages_serde* serde, const pg_page* page, const int64_t* row_offsets, int32_t n_ranges,
                                        const uint8_t** bytes, int64_t* byte_offsets);
/* readSerializedPage + deserialize.  `memory` = where `bytes` live; `types` gives the SQL type of every channel (the wire only
 * carries encodings).  RLE / DICTIONARY channels come back as blocks of those encodings.  A CHECKSUMMED page whose CRC32 does not
 * match fails with PG_ERR_INVALID_ARGUMENT.  out->blocks (in options.output_memory) stay valid until the next call. */
int32_t pg_pages_serde_deserialize(pg_pages_serde* serde, const uint8_t* bytes, int64_t size, int32_t memory,
                                   const int32_t* types, int32_t n_types, pg_page* out);
int32_t pg_pages_serde_destroy(pg_pages_serde* serde);

#ifdef __cplusplus
}
#endif
#endif /* PRESTO_GPU_H */
