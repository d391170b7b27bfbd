#!/bin/bash
# round 2, second 2-GPU call: the driver's bench at N = 2 with fused step pairs in the sharded lattice secondary,
# sharded adaptive dopri5 parity, rk4 sharded script pairs vs single steps
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
( timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/r02_n2b_bench.json 2> gpurun_out/r02_n2b_bench.err
echo "bench rc=$?"; tail -c 2500 gpurun_out/r02_n2b_bench.json; tail -5 gpurun_out/r02_n2b_bench.err
( timeout 600 $TR scripts/r02/run_sharded_adaptive.py 25 ) > gpurun_out/r02_n2b_sharded_adaptive.log 2>&1
echo "adaptive rc=$?"; grep -v "^\[" gpurun_out/r02_n2b_sharded_adaptive.log | tail -4
( timeout 600 $TR scripts/run_sharded_lattice.py 27 100 ) > gpurun_out/r02_n2b_sharded_pairs.log 2>&1
( ODEINTR_FUSE=0 timeout 600 $TR scripts/run_sharded_lattice.py 27 100 ) > gpurun_out/r02_n2b_sharded_single.log 2>&1
grep -h "parity\|metric" gpurun_out/r02_n2b_sharded_pairs.log gpurun_out/r02_n2b_sharded_single.log
