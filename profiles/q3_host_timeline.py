#!/usr/bin/env python
"""Wall-clock per operator call of the single-GPU Q3 pipeline (device-resident SF100-shaped tables): where the time between
kernels goes.  usage: python profiles/q3_host_timeline.py [n_orders]"""
import os, sys, time, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from presto_b200 import BIGINT, DATE, DOUBLE, INTEGER, operators
from presto_b200._lib import load
from presto_b200.pipelines import run_q3
from presto_b200.synthetic import q3_tables_torch
from bench import device_page

n_orders = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 150_000_000
backend = load()
customer, orders, lineitem = q3_tables_torch(n_orders, "cuda:0")
torch.cuda.synchronize()
cp = [device_page(torch, list(customer), [BIGINT, INTEGER])]
op = [device_page(torch, list(orders), [BIGINT, BIGINT, DATE, INTEGER])]
lp = [device_page(torch, list(lineitem), [BIGINT, DOUBLE, DOUBLE, DATE])]
log = []
def wrap(cls, name):
    f = getattr(cls, name)
    def g(self, *a, **k):
        t0 = time.perf_counter()
        r = f(self, *a, **k)
        log.append((type(self).__name__ + "." + name, (time.perf_counter() - t0) * 1e3))
        return r
    setattr(cls, name, g)
for name in ("add_input", "get_output", "finish", "close", "__init__"):
    wrap(operators.Operator, name) if name != "__init__" else None
for cls in (operators.HashBuilderOperator, operators.LookupJoinOperator, operators.FilterAndProjectOperator, operators.HashAggregationOperator,
            operators.FusedAggregationOperator, operators.FusedPageOperator):
    wrap(cls, "__init__")
stream = torch.cuda.current_stream()
for it in range(3):
    log.clear()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run_q3(cp, op, lp, backend=backend, device_resident=True, device=0, cuda_stream=stream.cuda_stream, fused="--unfused" not in sys.argv)
    torch.cuda.synchronize()
    total = (time.perf_counter() - t0) * 1e3
print(f"total {total:.2f} ms; inside operator calls {sum(t for _, t in log):.2f} ms")
for name, t in log:
    if t > 0.02:
        print(f"  {name:45s} {t:8.3f} ms")
