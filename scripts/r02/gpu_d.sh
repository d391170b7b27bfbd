#!/bin/bash
# round 2, call D: lattice tests (margin fix, two steps per pass), K4w comparison and sweep
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_lattice.py tests/test_gpu_lattice_large.py tests/test_gpu_lattice_ensemble.py -m gpu -q --durations=5 ) > gpurun_out/r02_d_pytest.log 2>&1
tail -12 gpurun_out/r02_d_pytest.log
( timeout 600 python scripts/r02/bench_warp.py 50 ) > gpurun_out/r02_d_bench_warp.log 2>&1; cat gpurun_out/r02_d_bench_warp.log
bash scripts/r02/sweep_warp.sh > gpurun_out/r02_d_sweep_warp.log 2>&1; cat gpurun_out/r02_d_sweep_warp.log
# adaptive dopri5 on the 2^26-cell lattice: warp kernel (K4dw) against the block-tiled one (K4d); same final-state hash
python scripts/bench_dopri5_lattice.py > gpurun_out/r02_d_dopri5_warp.log 2>&1; ODEINTR_NO_WARP=1 python scripts/bench_dopri5_lattice.py > gpurun_out/r02_d_dopri5_block.log 2>&1
cat gpurun_out/r02_d_dopri5_warp.log gpurun_out/r02_d_dopri5_block.log
