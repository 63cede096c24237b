#!/bin/bash
# Group-by cardinality sweep of the aggregation modes (PG_AGG_MODE: tiered = shared-memory tiers / hash-subset passes, direct = straight
# to the HBM table with L2-sized slot-range passes).  256 M resident rows, 4 sums.  usage: bash profiles/agg_mode_sweep.sh "G..." "MODES" "BUDGETS_MB"
GS=${1:-"10000 100000 1000000 3000000"}; MODES=${2:-"direct"}; BUDGETS=${3:-"48"}
for G in $GS; do for MODE in $MODES; do for B in $BUDGETS; do
  echo "G=$G MODE=$MODE L2_BUDGET_MB=$B: $(PG_AGG_MODE=$MODE PG_AGG_L2_BUDGET_MB=$B timeout 300 python bench.py --rows 268435456 --groups $G --extra 0 --steps 2 --warmup 1 --e2e-rows 1048576 --cpu-rows 1048576 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,2),'G rows/s;', round(d['ms_per_step'],2),'ms/step; add_input', round(d['roofline']['kernel_ms'],2), 'ms; launches',d['gpu_launches'],'groups',d['groups'])")"
done; done; done
