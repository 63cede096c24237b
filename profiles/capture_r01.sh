#!/bin/bash
# Round-1 profile capture (run on the GPU box through gpurun; outputs land in gpurun_out/, summaries are made from them on the CPU box
# with profiles/ncu_summary.py and committed under profiles/).  Nothing printed under ncu is a bench value.
set -x
mkdir -p gpurun_out
SMALL="--e2e-rows 1048576 --cpu-rows 1048576 --steps 1 --warmup 0"
# 1. launch list of the default bench command (fewer steps)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r01_launches_bench_1b.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r01_bench_under_ncu.log 2>&1
# 2. one full capture per hot kernel
timeout 600 ncu --set full --import-source on --clock-control none -k regex:hashAggFastKernel -c 1 -o gpurun_out/r01_hashagg_fast_1b -f python bench.py --extra 0 $SMALL > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:hashAggDirectKernel --launch-skip 1 -c 1 -o gpurun_out/r01_hashagg_direct_g10000 -f python bench.py --rows 268435456 --groups 10000 --extra 0 $SMALL > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:dense -c 8 -o gpurun_out/r01_join -f python bench.py --rows 16777216 --extra-scale 1.0 --only hash_join_orders_lineitem $SMALL > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:partition -c 2 -o gpurun_out/r01_partitioned_output -f python bench.py --rows 16777216 --extra-scale 1.0 --only partitioned_output_8_bigint_cols $SMALL > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fp_ -c 2 -o gpurun_out/r01_page_processor_q1 -f python bench.py --rows 16777216 --extra-scale 1.0 --only page_processor_q1 $SMALL > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
