#!/bin/bash
# round 2, call F: gpu tier (cabinet sets, drop-in read sizes), headline + config 4 with the copy-engine window, launch list + full ncu of P1 / P2
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/f_pytest_gpu.log 2>&1; tail -4 gpurun_out/f_pytest_gpu.log
( timeout 600 python bench.py --steps 10 ) > gpurun_out/f_bench_cfg3.log 2>&1; grep "^{" gpurun_out/f_bench_cfg3.log | cut -c1-200
( timeout 600 python bench.py --config 4 --steps 5 ) > gpurun_out/f_bench_cfg4.log 2>&1; grep "^{" gpurun_out/f_bench_cfg4.log | cut -c1-200; grep -o '"kernel_ms_per_step": [0-9.]*, "p2_resolve_ms_per_step": [0-9.]*' gpurun_out/f_bench_cfg4.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches_bench_lzx65536.csv python bench.py --steps 2 --warmup 1 --cpu-sample 256 --e2e-inflight 1 > gpurun_out/f_launches.log 2>&1; tail -2 gpurun_out/f_launches_bench_lzx65536.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_p1_lzx" -c 1 -f -o gpurun_out/f_prof_p1lzx python tools/quickbench.py 3 65536 1 > gpurun_out/f_ncu_p1.log 2>&1; tail -1 gpurun_out/f_ncu_p1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_p2_resolve" -c 1 -f -o gpurun_out/f_prof_p2 python tools/quickbench.py 3 65536 1 > gpurun_out/f_ncu_p2.log 2>&1; tail -1 gpurun_out/f_ncu_p2.log
