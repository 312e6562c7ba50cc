#!/usr/bin/env python
"""Streaming search (40 query rows x 1M bank rows) call and kernel time: bench.py's sim.stream leg alone."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
peaks = bench.load_peaks() if hasattr(bench, "load_peaks") else {"hbm_gbs": 6550.0, "source": "fallback"}
r = bench.bench_sim_stream(peaks)
print(json.dumps({"ms_per_call": r["ms_per_call"], "kernel_ms": r["roofline"]["kernel_ms"], "kernel_frac": r["roofline"]["frac"],
                  "whole_call_frac": r["roofline"]["whole_call_frac"], "exact": r["topk_equal_torch_fp32"], "peak": r["roofline"]["peak"]}))
