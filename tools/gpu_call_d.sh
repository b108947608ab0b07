#!/bin/bash
# round 2, call D: parity tier after the bulk-copy fix and lane parking; config 4 / 3 / 2 lines; ncu of Quantum P1 and of P2
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q ) > gpurun_out/d_pytest_gpu.log 2>&1; tail -5 gpurun_out/d_pytest_gpu.log
for c in 4 3 2; do
  ( timeout 600 python bench.py --config $c --steps 5 --e2e-inflight 1 ) > gpurun_out/d_bench_cfg$c.log 2>&1; grep "^{" gpurun_out/d_bench_cfg$c.log | cut -c1-200; grep -o '"kernel_ms_per_step": [0-9.]*, "p2_resolve_ms_per_step": [0-9.]*' gpurun_out/d_bench_cfg$c.log
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_p1_qtm" -c 1 -f -o gpurun_out/d_prof_p1qtm python tools/quickbench.py 2 16384 1 > gpurun_out/d_ncu_qtm.log 2>&1; tail -1 gpurun_out/d_ncu_qtm.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_p2_resolve" -c 1 -f -o gpurun_out/d_prof_p2 python tools/quickbench.py 3 65536 1 > gpurun_out/d_ncu_p2.log 2>&1; tail -1 gpurun_out/d_ncu_p2.log
