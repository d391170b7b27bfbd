#!/bin/bash
# round 2, call P: K2 / K3 with the step attempt outlined (code size / instruction-cache A/B)
mkdir -p gpurun_out
( timeout 900 python scripts/r02/bench_outline.py ) > gpurun_out/r02_p_outline.log 2>&1
cut -c1-330 gpurun_out/r02_p_outline.log
