#!/bin/bash
# round 2, call Y5: K5a at its new default (two resident blocks): a few of the ensemble tests
mkdir -p gpurun_out
( timeout 80 python -m pytest tests/test_gpu_lattice_ensemble.py -m gpu -q -k "1024-rk5_i or 300-rk5_a or 300-rk78_a or 37-bsd" ) > gpurun_out/r02_y5_pytest.log 2>&1; tail -3 gpurun_out/r02_y5_pytest.log
