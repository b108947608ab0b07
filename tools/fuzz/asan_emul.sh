#!/bin/bash
# The device logic compiled for the host (tests/emul/emul.cpp) under AddressSanitizer + UBSan: the emulation tests and short runs of
# the fuzzers (alignment checks included: a misaligned word access is a fault on the GPU).  Intermediate arrays (match records, frame info, aux tables, the shared-memory structs) are sized exactly as on the
# device, so an overrun of any of them is reported here.  CPU only; run from the repository root.
set -e
SO=tests/emul/libmsgpu_emul.so
cp $SO /tmp/libmsgpu_emul_keep.so
trap 'cp /tmp/libmsgpu_emul_keep.so $SO; touch $SO' EXIT
g++ -O1 -g -std=c++17 -fPIC -shared -fsanitize=address,undefined -fno-omit-frame-pointer -Wno-unknown-pragmas -o $SO tests/emul/emul.cpp
touch $SO
export LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0:halt_on_error=1
python -m pytest tests/test_emulation.py -x -q -s 2>&1 | grep -a "runtime error\|ERROR: Addr\|passed\|failed" | sort | uniq -c
python tools/fuzz/fuzz_repair.py 11 20
python tools/fuzz/fuzz_kwaj.py 11 300
python tools/fuzz/fuzz_units.py 2>&1 | tail -1
