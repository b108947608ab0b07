#!/bin/bash
# round 2, call W: LZX P1 with byte-wise literal stores + the two-round shared-memory length search: parity (LZX, MSZIP, CHM,
# cabinets, the reference suites) and the headline batch with stage timing; MSZIP batch for the length search
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_w_long_units_gpu.py::test_large_files_cab_65535_block_folders ) > gpurun_out/w_pytest_gpu.log 2>&1; tail -4 gpurun_out/w_pytest_gpu.log
( QB_STAGE=1 timeout 300 python tools/quickbench.py 3 65536 4 ) > gpurun_out/w_qb_lzx.log 2>&1; echo "lzx: $(grep 'stage_timing=True' gpurun_out/w_qb_lzx.log | tail -1) $(grep -o 'roundtrip_ok=[A-Za-z]*' gpurun_out/w_qb_lzx.log) $(grep -o 'best [0-9.]* ms = [0-9.]* GB/s' gpurun_out/w_qb_lzx.log)"
( QB_STAGE=1 timeout 300 python tools/quickbench.py 1 65536 4 ) > gpurun_out/w_qb_zip.log 2>&1; echo "zip: $(grep 'stage_timing=True' gpurun_out/w_qb_zip.log | tail -1) $(grep -o 'roundtrip_ok=[A-Za-z]*' gpurun_out/w_qb_zip.log) $(grep -o 'best [0-9.]* ms = [0-9.]* GB/s' gpurun_out/w_qb_zip.log)"
