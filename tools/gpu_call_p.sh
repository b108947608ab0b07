#!/bin/bash
# round 2, call P: per-unit frame slots (long units: up to 64 frames per launch round) - the new long-unit tests incl. the
# reference's 65 535-block folders, the gpu tier, the headline bench
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_w_long_units_gpu.py -m gpu -q -s ) > gpurun_out/p_pytest_long.log 2>&1; tail -8 gpurun_out/p_pytest_long.log
( time timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_w_long_units_gpu.py ) > gpurun_out/p_pytest_gpu.log 2>&1; tail -4 gpurun_out/p_pytest_gpu.log
( time timeout 600 python bench.py ) > gpurun_out/p_bench_cfg3.log 2>&1; grep "^{" gpurun_out/p_bench_cfg3.log | cut -c1-250
grep -o '"e2e": {[^}]*' gpurun_out/p_bench_cfg3.log | cut -c1-300
