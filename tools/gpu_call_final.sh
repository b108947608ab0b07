#!/bin/bash
# round 2, final call: gpu tier, smoke, bench.py over the configs on one GPU (JSON lines -> profiles/r2_final_*.json), the reference arm,
# the launch list of the headline bench and full ncu captures of the three kernels that carry it
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q ) > gpurun_out/n_pytest_gpu.log 2>&1; tail -4 gpurun_out/n_pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/n_smoke.log 2>&1; tail -1 gpurun_out/n_smoke.log
( time timeout 600 python bench.py ) > gpurun_out/n_bench_cfg3.log 2>&1; grep "^{" gpurun_out/n_bench_cfg3.log | cut -c1-250
( time timeout 300 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/n_bench_reference.log 2>&1; grep "^{" gpurun_out/n_bench_reference.log | cut -c1-200
for c in 2 6 4; do
  ( time timeout 600 python bench.py --config $c --steps 5 --e2e-inflight 1 ) > gpurun_out/n_bench_cfg$c.log 2>&1; grep "^{" gpurun_out/n_bench_cfg$c.log | cut -c1-250
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/n_launches_bench_lzx65536.csv python bench.py --steps 2 --warmup 1 --cpu-sample 256 --e2e-inflight 1 > gpurun_out/n_launches.log 2>&1; tail -2 gpurun_out/n_launches_bench_lzx65536.csv | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_p1_lzx" -c 1 -f -o gpurun_out/n_prof_p1lzx python tools/quickbench.py 3 65536 1 > gpurun_out/n_ncu_p1.log 2>&1; tail -1 gpurun_out/n_ncu_p1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_p2_resolve" -c 1 -f -o gpurun_out/n_prof_p2 python tools/quickbench.py 3 65536 1 > gpurun_out/n_ncu_p2.log 2>&1; tail -1 gpurun_out/n_ncu_p2.log
