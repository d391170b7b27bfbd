#!/bin/bash
# round 2, 8-GPU call: the driver's bench at N = 8 (headline, e2e against the box's PCIe ceiling, sharded lattice with
# fused step pairs and the overlapped NCCL ring exchange), sharded adaptive dopri5 parity at 8 ranks
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
( timeout 600 $TR bench.py --gpus 8 --steps 5 --warmup 3 ) > gpurun_out/r02_n8_bench.json 2> gpurun_out/r02_n8_bench.err
echo "bench rc=$?"; tail -c 2500 gpurun_out/r02_n8_bench.json; tail -3 gpurun_out/r02_n8_bench.err
( timeout 300 $TR scripts/r02/run_sharded_adaptive.py 25 ) > gpurun_out/r02_n8_sharded_adaptive.log 2>&1
echo "adaptive rc=$?"; grep -v "^\[" gpurun_out/r02_n8_sharded_adaptive.log | grep "parity\|metric" | tail -4
