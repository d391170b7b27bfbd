#!/bin/bash
# round 2, last call: smoke() on the final tree
mkdir -p gpurun_out
( time timeout 300 python __graft_entry__.py smoke ) > gpurun_out/r02_last_smoke.log 2>&1; echo "smoke rc=$?"; tail -14 gpurun_out/r02_last_smoke.log
