#!/bin/bash
# First GPU call of the next round (everything below was prepared, built and CPU/emulation-tested after round 1's GPU budget ran out):
#   1. the whole gpu test tier - tests/test_z_kwaj.py (KWAJ framing, MSZIP repair mode) has never run on a B200;
#   2. the LZX P1 experiments 31-53 against the default 30, the MSZIP ones (15-19) against 14 (tools/variant_bench.py: one batch, stage timing, verified);
#   3. the other BASELINE configs at bench size (MSZIP, reset intervals, mixed, Quantum) and a per-GPU share of configs[3];
#   4. the bench line.
# usage: gpurun --timeout 2400 -- 'bash tools/r2_first_call.sh'      (about 20-25 minutes of box time; every shape is measured in a
#        process of its own, so one that faults or hangs costs its own line only)
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests -m gpu -q --durations=10 ) > gpurun_out/r2_pytest_gpu.log 2>&1; tail -8 gpurun_out/r2_pytest_gpu.log
( MSGPU_TEST_EXPERIMENTAL=1 timeout 420 python -m pytest tests/test_experimental_gpu.py -m gpu -q ) > gpurun_out/r2_pytest_experimental.log 2>&1; tail -5 gpurun_out/r2_pytest_experimental.log
timeout 700 python tools/variant_bench.py 65536 30 31 32 33 34 35 36 37 38 39 40 41 42 43 44 45 46 47 48 49 50 51 52 53 > gpurun_out/r2_variants.log 2>&1; cat gpurun_out/r2_variants.log
VB_P2=0,1 timeout 150 python tools/variant_bench.py 65536 30 > gpurun_out/r2_variants_p2.log 2>&1; cat gpurun_out/r2_variants_p2.log      # the byte-parallel pass A of P2
VB_P2=0,1 VB_CODEC=1 timeout 360 python tools/variant_bench.py 32768 14 15 16 17 18 19 > gpurun_out/r2_variants_p2_zip.log 2>&1; cat gpurun_out/r2_variants_p2_zip.log
VB_CODEC=2 timeout 240 python tools/variant_bench.py 16384 0 1 2 3 4 7 > gpurun_out/r2_variants_qtm.log 2>&1; cat gpurun_out/r2_variants_qtm.log      # Quantum: two-level model scan (1), loop-free renormalisation (2), both (3), reciprocal divisions (4), all (7)
timeout 300 python tools/configs_bench.py 32768 > gpurun_out/r2_configs.log 2>&1; cat gpurun_out/r2_configs.log
timeout 200 python tools/config4_bench.py 65536 > gpurun_out/r2_config4.log 2>&1; tail -1 gpurun_out/r2_config4.log
timeout 240 python bench.py --e2e-inflight 2 > gpurun_out/r2_bench.log 2>&1; grep "^{" gpurun_out/r2_bench.log | cut -c1-300; grep -o '"e2e": {.*' gpurun_out/r2_bench.log | cut -c1-700
python tools/rank_variants.py gpurun_out/r2_variants*.log > gpurun_out/r2_ranking.txt 2>&1; cat gpurun_out/r2_ranking.txt
