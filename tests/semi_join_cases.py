"""TestHashSemiJoinOperator (presto-main-base/src/test/java/com/facebook/presto/operator/TestHashSemiJoinOperator.java) re-encoded:
(set rows, probe rows as (key, other...) tuples, expected output rows)."""
REFERENCE_CASES = {
    # testSemiJoin :92-153
    "testSemiJoin": ([10, 30, 30, 35, 36, 37, 50], [(30 + i, i) for i in range(10)],
                     [(30, 0, True), (31, 1, False), (32, 2, False), (33, 3, False), (34, 4, False), (35, 5, True), (36, 6, True), (37, 7, True),
                      (38, 8, False), (39, 9, False)]),
    # testBuildSideNulls :250-305: a NULL in the set turns "absent" into NULL
    "testBuildSideNulls": ([0, 1, 2, 2, 3, None], [(1,), (2,), (3,), (4,)], [(1, True), (2, True), (3, True), (4, None)]),
    # testProbeSideNulls :308-363
    "testProbeSideNulls": ([0, 1, 3], [(0,), (None,), (1,), (2,)], [(0, True), (None, None), (1, True), (2, False)]),
    # testProbeAndBuildNulls :366-422
    "testProbeAndBuildNulls": ([0, 1, None, 3], [(0,), (None,), (1,), (2,)], [(0, True), (None, None), (1, True), (2, None)]),
}


def varchar_key(v):
    """The same cases over a VARCHAR join channel: key k becomes the slice b"key-<k>" (NULL stays NULL)."""
    return None if v is None else b"key-%d" % v


def run_semi_join(backend, set_values, probe_rows, set_pages=1, varchar=False):
    import numpy as np
    from presto_b200 import BIGINT, VARCHAR, Block, HashBuilderOperator, HashSemiJoinOperator, JoinBridge, Page, to_pages
    key_type = VARCHAR if varchar else BIGINT
    def key_block(values):
        return Block.from_strings([varchar_key(v) for v in values]) if varchar else Block.from_values(BIGINT, values, [v is None for v in values])
    bridge = JoinBridge(backend=backend)
    builder = HashBuilderOperator(bridge, [key_type], [0], [], backend=backend)
    cuts = np.linspace(0, len(set_values), set_pages + 1).astype(int)
    pages = [Page([key_block(set_values[a:b])], b - a) for a, b in zip(cuts[:-1], cuts[1:])]
    assert to_pages(builder, pages) == []
    width = len(probe_rows[0]) if probe_rows else 1
    op = HashSemiJoinOperator(bridge, [key_type] + [BIGINT] * (width - 1), 0, backend=backend)
    cols = [[r[c] for r in probe_rows] for c in range(width)]
    probe = Page([key_block(cols[0])] + [Block.from_values(BIGINT, col, [v is None for v in col]) for col in cols[1:]], len(probe_rows))
    out = [r for p in to_pages(op, [probe]) for r in p.rows()]
    op.close(); builder.close(); bridge.close()
    return out
