"""presto_b200 — B200-native (sm_100a) operator pipeline behind prestodb/presto's Operator / Page interfaces.

Host-side mirror (python) of the reference's operator API over the C ABI in include/presto_gpu.h; the compute is the
hand-written CUDA in presto_b200/csrc (libprestogpu.so).  Nothing here imports the test oracle.
"""
from ._abi import (  # noqa: F401
    BIGINT, INTEGER, DOUBLE, DATE, BOOLEAN, VARCHAR, MEM_HOST, MEM_DEVICE, STEP_SINGLE, STEP_PARTIAL, STEP_FINAL,
    JOIN_INNER, JOIN_PROBE_OUTER, JOIN_LOOKUP_OUTER, JOIN_FULL_OUTER, AGG_COUNT_STAR, AGG_COUNT_COLUMN, AGG_SUM_DOUBLE, AGG_SUM_BIGINT, AGG_AVG_DOUBLE,
    AGG_AVG_BIGINT, AGG_SUM_SHORT_DECIMAL, AGG_SUM_LONG_DECIMAL, LONG_DECIMAL, AGG_MIN_DOUBLE, AGG_MAX_DOUBLE, AGG_MIN_BIGINT, AGG_MAX_BIGINT, PrestoException,
)
from .page import Block, Page  # noqa: F401
from .expressions import Program  # noqa: F401
from .operators import (  # noqa: F401
    Aggregate, FilterAndProjectOperator, FusedAggregationOperator, FusedJoin, FusedPageOperator, GroupByHash, HashAggregationOperator, HashBuilderOperator, HashSemiJoinOperator, JoinBridge,
    LookupJoinOperator, LookupOuterOperator, Operator, OperatorFactory, PagesSerde, PartitionedOutputOperator, to_pages,
)

__version__ = "0.1.0"
