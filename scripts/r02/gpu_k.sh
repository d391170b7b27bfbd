#!/bin/bash
# round 2, call K: bs / bsd on large systems; fused step pairs in the sharded rk4 run (tests, one-rank timing A/B)
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_lattice.py -m gpu -q -k "bulirsch or general_snippets or shard" ) > gpurun_out/r02_k_pytest.log 2>&1
tail -15 gpurun_out/r02_k_pytest.log
( timeout 300 python scripts/run_sharded_lattice.py 27 100 ) > gpurun_out/r02_k_sharded1_pairs.log 2>&1
( ODEINTR_FUSE=0 timeout 300 python scripts/run_sharded_lattice.py 27 100 ) > gpurun_out/r02_k_sharded1_single.log 2>&1
( ODEINTR_NO_OVERLAP=1 timeout 300 python scripts/run_sharded_lattice.py 27 100 ) > gpurun_out/r02_k_sharded1_pairs_serial.log 2>&1
grep -h "parity\|metric" gpurun_out/r02_k_sharded1_*.log
