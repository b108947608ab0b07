#!/bin/bash
# Second GPU call of the next round: confirm the shapes picked from tools/r2_first_call.sh's measurements BEFORE they become the
# compiled-in defaults.  The shapes come from the environment (they are read when a context is created), e.g.
#   gpurun --timeout 1200 -- 'MSGPU_LZX_VARIANT=52 MSGPU_P2_VARIANT=1 MSGPU_ZIP_VARIANT=17 MSGPU_QTM_VARIANT=3 bash tools/r2_second_call.sh'
# 1. the whole gpu test tier with those shapes in force; 2. smoke; 3. both bench arms; 4. the launch list of bench.py under ncu;
# 5. one full ncu capture each of P1 (LZX) and P2 for profiles/ (tools/ncu_summary.sh turns them into text).
mkdir -p gpurun_out
echo "shapes: LZX=${MSGPU_LZX_VARIANT:-30} ZIP=${MSGPU_ZIP_VARIANT:-14} QTM=${MSGPU_QTM_VARIANT:-0} P2=${MSGPU_P2_VARIANT:-0}" | tee gpurun_out/r2b_shapes.txt
( time timeout 400 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2b_pytest_gpu.log; tail -6 gpurun_out/r2b_pytest_gpu.log
( timeout 100 python __graft_entry__.py --smoke ) > gpurun_out/r2b_smoke.log 2>&1; tail -1 gpurun_out/r2b_smoke.log
( timeout 240 python bench.py --e2e-inflight 2 ) > gpurun_out/r2b_bench.log 2>&1; grep "^{" gpurun_out/r2b_bench.log | cut -c1-400
( timeout 150 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r2b_bench_reference.log 2>&1; grep "^{" gpurun_out/r2b_bench_reference.log | cut -c1-200
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 2 --warmup 1 --cpu-sample 256 > gpurun_out/r2b_launches.log 2>&1; tail -2 gpurun_out/r2b_launches.csv
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_p1_lzx" -c 1 -f -o gpurun_out/r2b_prof_p1lzx python tools/quickbench.py 3 65536 1 > gpurun_out/r2b_ncu_p1.log 2>&1; tail -1 gpurun_out/r2b_ncu_p1.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_p2_resolve" -c 1 -f -o gpurun_out/r2b_prof_p2 python tools/quickbench.py 3 65536 1 > gpurun_out/r2b_ncu_p2.log 2>&1; tail -1 gpurun_out/r2b_ncu_p2.log
