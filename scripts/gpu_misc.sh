#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lattice.py -m gpu -x -q -k "fixed_ends_emulated" > gpurun_out/pytest_misc.log 2>&1; echo "pytest rc=$?"; grep "AssertionError\|assert bad\|passed\|failed" gpurun_out/pytest_misc.log | tail -5
ODEINTR_NO_TILED=1 timeout 900 python -m pytest tests/test_gpu_lattice.py -m gpu -x -q -k "fixed_ends_emulated" > gpurun_out/pytest_misc2.log 2>&1; echo "pytest (staged) rc=$?"; grep "AssertionError\|assert bad\|passed\|failed" gpurun_out/pytest_misc2.log | tail -5
