#!/bin/bash
# round 2, call Q: host-buffer pipeline with a stream per codec (Quantum chains on streams of their own, their output queued
# last; MSZIP output copied before the straggler check): the host-path tests, config 5 (mixed) and config 2 (MSZIP) end to end
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "host or mixed or multi_device or digest" ) > gpurun_out/q_pytest_host.log 2>&1; tail -4 gpurun_out/q_pytest_host.log
( timeout 900 python bench.py --config 5 --steps 3 --e2e-inflight 1 ) > gpurun_out/q_bench_cfg5.log 2>&1; grep "^{" gpurun_out/q_bench_cfg5.log | cut -c1-200; grep -o '"e2e": {[^}]*' gpurun_out/q_bench_cfg5.log | cut -c1-300
( timeout 600 python bench.py --config 2 --steps 5 --e2e-inflight 1 ) > gpurun_out/q_bench_cfg2.log 2>&1; grep "^{" gpurun_out/q_bench_cfg2.log | cut -c1-200; grep -o '"e2e": {[^}]*' gpurun_out/q_bench_cfg2.log | cut -c1-300
