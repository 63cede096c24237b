#!/bin/bash
# Round-2 profile capture (run on the GPU box through gpurun; outputs land in gpurun_out/, summaries are made from them on the CPU box with
# profiles/ncu_summary.py / launch_summary.py and committed under profiles/).  Nothing printed under ncu is a bench value.
set -x
mkdir -p gpurun_out
# 1. launch lists of one step of the headline (Q1 SF100, VARCHAR keys) and of the fused Q3 SF100 pipeline
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_q1_headline.csv python profiles/run_pipeline.py q1h --steps 1 > gpurun_out/r02_q1h_under_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_q3_pipeline.csv python profiles/run_pipeline.py q3 --steps 1 > gpurun_out/r02_q3_under_ncu.log 2>&1
# 2. full captures: the generated Q3 lineitem fragment kernel, the generic aggregation kernel (unfused Q3: (bigint, date, integer) keys),
#    the join's build + probe kernels (configs[2] at 1/8 scale), the state hand-off kernels
timeout 600 ncu --set full --import-source on --clock-control none -k regex:fused_kernel -c 2 -o gpurun_out/r02_q3_fused_kernel -f python profiles/run_pipeline.py q3 --steps 1 --warmup 0 > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:hashAggKernel -c 1 -o gpurun_out/r02_hashagg_generic_q3 -f python profiles/run_pipeline.py q3 --unfused --steps 1 --warmup 0 > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:dense -c 8 -o gpurun_out/r02_join -f python profiles/run_pipeline.py join --scale 0.125 --steps 1 --warmup 0 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
