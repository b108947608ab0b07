#!/bin/bash
# round 2, call E: gpu tier (drop-in fixes, digest sinks, E8 epilogue), P2 copy-engine A/B, headline bench with the MD5 sink line
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/e_pytest_gpu.log 2>&1; tail -4 gpurun_out/e_pytest_gpu.log
timeout 200 python tools/variant_bench.py 65536 - MSGPU_P2_BULK=1 - MSGPU_P2_BULK=1 > gpurun_out/e_p2_bulk_ab.log 2>&1; cat gpurun_out/e_p2_bulk_ab.log
VB_CODEC=1 timeout 200 python tools/variant_bench.py 65536 - MSGPU_P2_BULK=1 > gpurun_out/e_p2_bulk_ab_zip.log 2>&1; cat gpurun_out/e_p2_bulk_ab_zip.log
( timeout 600 python bench.py --steps 10 ) > gpurun_out/e_bench_cfg3.log 2>&1; grep "^{" gpurun_out/e_bench_cfg3.log | cut -c1-200; grep -o '"sink_md5": {[^}]*}' gpurun_out/e_bench_cfg3.log
