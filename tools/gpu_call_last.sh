#!/bin/bash
# round 2, last call: the reference's own suites, the cabinet front end and the widening tests on the final build
mkdir -p gpurun_out
( time timeout 140 python -m pytest tests/test_y_reference_suites_gpu.py tests/test_x_oab.py tests/test_z_kwaj.py tests/test_cab_frontend.py tests/test_widening_gpu.py -m gpu -q -x ) > gpurun_out/last_pytest.log 2>&1; tail -4 gpurun_out/last_pytest.log
