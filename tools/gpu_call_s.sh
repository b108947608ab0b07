#!/bin/bash
# round 2, call S: Quantum chains on streams of their own also in Quantum-only host batches: host-path tests, config 6 end to end
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "host or mixed or quantum" ) > gpurun_out/s_pytest_host.log 2>&1; tail -4 gpurun_out/s_pytest_host.log
( timeout 900 python bench.py --config 6 --steps 3 --e2e-inflight 1 ) > gpurun_out/s_bench_cfg6.log 2>&1; grep "^{" gpurun_out/s_bench_cfg6.log | cut -c1-200; grep -o '"e2e": {[^}]*' gpurun_out/s_bench_cfg6.log | cut -c1-300
