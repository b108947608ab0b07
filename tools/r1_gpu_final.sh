#!/bin/bash
# Final round-1 GPU pass: the whole gpu test tier, smoke, bench both arms, the launch list of bench.py under ncu, one full ncu capture of P1.
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_final.log; tail -6 gpurun_out/pytest_gpu_final.log
( timeout 100 python __graft_entry__.py --smoke ) > gpurun_out/smoke_final.log 2>&1; tail -1 gpurun_out/smoke_final.log
( timeout 200 python bench.py ) > gpurun_out/bench_final.log 2>&1; grep "^{" gpurun_out/bench_final.log | cut -c1-400
( timeout 150 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_reference_final.log 2>&1; grep "^{" gpurun_out/bench_reference_final.log | cut -c1-200
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --cpu-sample 256 > gpurun_out/launches_final.log 2>&1; tail -2 gpurun_out/launches_final.csv
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_p1_lzx" -c 1 -f -o gpurun_out/prof_p1lzx_v30 python tools/quickbench.py 3 65536 1 > gpurun_out/ncu_v30.log 2>&1; tail -1 gpurun_out/ncu_v30.log
