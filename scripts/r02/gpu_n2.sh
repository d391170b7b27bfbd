#!/bin/bash
# round 2, 2-GPU call: the driver's bench at N = 2 (sharded lattice secondary with the NCCL ring exchange, overlapped),
# sharded adaptive dopri5 parity + timing, rk4 sharded script with and without overlap
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
( timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/r02_n2_bench.json 2> gpurun_out/r02_n2_bench.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/r02_n2_bench.json; tail -5 gpurun_out/r02_n2_bench.err
( timeout 600 $TR scripts/r02/run_sharded_adaptive.py 25 ) > gpurun_out/r02_n2_sharded_adaptive.log 2>&1
echo "adaptive rc=$?"; grep -v "^\[" gpurun_out/r02_n2_sharded_adaptive.log | tail -8
( timeout 600 $TR scripts/run_sharded_lattice.py 27 100 ) > gpurun_out/r02_n2_sharded_overlap.log 2>&1
( ODEINTR_NO_OVERLAP=1 timeout 600 $TR scripts/run_sharded_lattice.py 27 100 ) > gpurun_out/r02_n2_sharded_serial.log 2>&1
( timeout 600 $TR scripts/run_sharded_lattice.py 25 200 ) > gpurun_out/r02_n2_sharded_overlap_25.log 2>&1
( ODEINTR_NO_OVERLAP=1 timeout 600 $TR scripts/run_sharded_lattice.py 25 200 ) > gpurun_out/r02_n2_sharded_serial_25.log 2>&1
grep -h "parity\|metric" gpurun_out/r02_n2_sharded_*.log
