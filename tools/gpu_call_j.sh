#!/bin/bash
# round 2, call J (call I ran the same forms with a register allocation that halved the occupancy - discarded): forms of the resolve kernel (owner pass A / word loads in pass B / five CTAs per SM) on the LZX and MSZIP headline batches
mkdir -p gpurun_out
for m in 4 5; do for o in 0 1 2 3; do
  ( MSGPU_P2_MINB=$m MSGPU_P2_OWNER=$o QB_STAGE=1 timeout 300 python tools/quickbench.py 3 65536 4 ) > gpurun_out/j_qb_lzx_m${m}_o$o.log 2>&1; echo "lzx minb $m form $o: $(grep 'stage_timing=True' gpurun_out/j_qb_lzx_m${m}_o$o.log | tail -1) $(grep -o 'roundtrip_ok=[A-Za-z]*' gpurun_out/j_qb_lzx_m${m}_o$o.log)"
done; done
for m in 4 5; do for o in 0 3; do
  ( MSGPU_P2_MINB=$m MSGPU_P2_OWNER=$o QB_STAGE=1 timeout 300 python tools/quickbench.py 1 65536 4 ) > gpurun_out/j_qb_zip_m${m}_o$o.log 2>&1; echo "zip minb $m form $o: $(grep 'stage_timing=True' gpurun_out/j_qb_zip_m${m}_o$o.log | tail -1) $(grep -o 'roundtrip_ok=[A-Za-z]*' gpurun_out/j_qb_zip_m${m}_o$o.log)"
done; done
