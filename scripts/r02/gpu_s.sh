#!/bin/bash
# round 2, call S: K5a other steppers + plain wide ensembles (tests), sharded run with the pairs' edges on the second stream
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_lattice_ensemble.py tests/test_gpu_lattice.py -m gpu -q -k "ensemble or shard" ) > gpurun_out/r02_s_pytest.log 2>&1
tail -12 gpurun_out/r02_s_pytest.log
for lg in 25 27; do
  ( timeout 300 python scripts/run_sharded_lattice.py $lg 200 ) > gpurun_out/r02_s_sharded1_${lg}.log 2>&1
  ( ODEINTR_NO_OVERLAP=1 timeout 300 python scripts/run_sharded_lattice.py $lg 200 ) > gpurun_out/r02_s_sharded1_${lg}_serial.log 2>&1
done
grep -h "parity\|metric" gpurun_out/r02_s_sharded1_*.log
