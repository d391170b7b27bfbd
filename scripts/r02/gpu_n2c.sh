#!/bin/bash
# round 2, third 2-GPU call: sharded rk4 with the pairs' edges on the second stream (parity + timing at 2^25 / 2^27 per GPU)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
( timeout 300 $TR scripts/run_sharded_lattice.py 25 200 ) > gpurun_out/r02_n2c_sharded_25.log 2>&1
( timeout 300 $TR scripts/run_sharded_lattice.py 27 200 ) > gpurun_out/r02_n2c_sharded_27.log 2>&1
( SPE=3 timeout 300 $TR scripts/run_sharded_lattice.py 25 100 ) > gpurun_out/r02_n2c_sharded_25_spe3.log 2>&1
grep -h "parity\|metric\|Error\|error" gpurun_out/r02_n2c_sharded_*.log | cut -c1-330
