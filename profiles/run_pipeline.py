#!/usr/bin/env python
"""Run ONE pipeline of bench.py on cuda:0 (for ncu launch lists / captures): run_pipeline.py q1h|q1|q3|join [--unfused] [--scale S] [--steps K]."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("what")
ap.add_argument("--unfused", action="store_true")
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--warmup", type=int, default=1)
a = ap.parse_args()
import torch  # noqa: E402
from presto_b200._lib import load  # noqa: E402
torch.cuda.set_device(0)
backend = load()
backend.check(backend.set_device(0))
if a.what == "q1h":      # the bench headline: Q1 over SF100 lineitem with VARCHAR(1) keys through the fused operator (TMA ring kernel)
    cols = bench.gen_q1_lineitem(torch, "cuda:0", int(bench.Q1_SF100_ROWS * a.scale), 1, "varchar")
    torch.cuda.synchronize()
    rr = bench.run_q1_headline(torch, backend, cols, "varchar", a.steps, a.warmup, 0, None)
    r = {"ms": rr["ms"], "kernel_ms": rr["kernel_ms"], "launches": rr["launches"], "info": rr["info"]}
elif a.what == "q1":
    r = bench.run_q1_pipeline(torch, backend, 0, int(600_037_902 * a.scale), a.steps, a.warmup, fused=not a.unfused)
elif a.what == "q3":
    r = bench.run_q3(torch, backend, 0, int(150_000_000 * a.scale), a.steps, a.warmup, fused=not a.unfused)
elif a.what == "join":
    r = bench.run_join(torch, backend, 0, int(150_000_000 * a.scale), int(600_037_902 * a.scale), a.steps, a.warmup)
else:
    raise SystemExit("unknown pipeline")
print(json.dumps(r))
