#!/bin/bash
# round 2, final call: the whole gpu tier (incl. the 65 535-block folders), smoke, the headline bench, its ncu launch list, and
# ncu --set full captures of the two kernels of the headline step (roofline.traffic)
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/z_pytest_gpu.log 2>&1; tail -4 gpurun_out/z_pytest_gpu.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/z_smoke.log 2>&1; tail -3 gpurun_out/z_smoke.log
( time timeout 600 python bench.py ) > gpurun_out/z_bench_cfg3.log 2>&1; grep "^{" gpurun_out/z_bench_cfg3.log | cut -c1-250; grep -o '"e2e": {[^}]*' gpurun_out/z_bench_cfg3.log | cut -c1-300
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/z_launches_bench_cfg3.csv python bench.py --steps 2 --warmup 1 --e2e-inflight 1 > gpurun_out/z_bench_under_ncu.log 2>&1; tail -c 300 gpurun_out/z_launches_bench_cfg3.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_p1_lzx|k_p2_resolve" -c 2 -f -o gpurun_out/z_prof_step python tools/quickbench.py 3 65536 1 > gpurun_out/z_ncu_step.log 2>&1; tail -2 gpurun_out/z_ncu_step.log
