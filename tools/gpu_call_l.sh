#!/bin/bash
# round 2, call L: gpu tier on the pipelined resolve (all instantiations); P1 LZX with the LENGTH-symbol load deferred (alternative build, MSGPU_LIB) - parity + A/B; headline line
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/l_pytest_gpu.log 2>&1; tail -4 gpurun_out/l_pytest_gpu.log
( MSGPU_LIB=$PWD/tools/libmsgpu_alt.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_widening_gpu.py -m gpu -q ) > gpurun_out/l_pytest_gpu_alt.log 2>&1; tail -2 gpurun_out/l_pytest_gpu_alt.log
for r in 1 2; do
( QB_STAGE=1 timeout 300 python tools/quickbench.py 3 65536 4 ) > gpurun_out/l_qb_lzx_base$r.log 2>&1; echo "lzx base: $(grep 'stage_timing=True' gpurun_out/l_qb_lzx_base$r.log | tail -1) $(grep -o 'roundtrip_ok=[A-Za-z]*' gpurun_out/l_qb_lzx_base$r.log)"
( MSGPU_LIB=$PWD/tools/libmsgpu_alt.so QB_STAGE=1 timeout 300 python tools/quickbench.py 3 65536 4 ) > gpurun_out/l_qb_lzx_alt$r.log 2>&1; echo "lzx alt:  $(grep 'stage_timing=True' gpurun_out/l_qb_lzx_alt$r.log | tail -1) $(grep -o 'roundtrip_ok=[A-Za-z]*' gpurun_out/l_qb_lzx_alt$r.log)"
done
( timeout 600 python bench.py --steps 10 ) > gpurun_out/l_bench_cfg3.log 2>&1; grep "^{" gpurun_out/l_bench_cfg3.log | cut -c1-200
