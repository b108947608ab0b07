#!/bin/bash
# round 2, call O: host-buffer pipeline with the growing first sub-waves: gpu tier + the headline bench (e2e)
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q ) > gpurun_out/o_pytest_gpu.log 2>&1; tail -4 gpurun_out/o_pytest_gpu.log
( time timeout 600 python bench.py ) > gpurun_out/o_bench_cfg3.log 2>&1; grep "^{" gpurun_out/o_bench_cfg3.log | cut -c1-250
grep -o '"e2e": {[^}]*' gpurun_out/o_bench_cfg3.log | cut -c1-300
