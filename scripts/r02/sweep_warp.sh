#!/bin/bash
# K4w parameter sweep: cells per lane, resident blocks per SM, one or two steps per pass
for fuse in 0 1; do for cpt in 9 11 13; do for mb in 1 2; do
  echo "== bistable fuse=$fuse cpt=$cpt min_blocks=$mb"; ODEINTR_FUSE=$fuse ODEINTR_WARP_CPT=$cpt ODEINTR_WARP_MIN_BLOCKS=$mb timeout 120 python scripts/r02/prof_warp.py bistable 2>&1 | tail -1
done; done; done
for fuse in 0 1; do for cpt in 5 7 9; do for mb in 1 2; do
  echo "== chain fuse=$fuse cpt=$cpt min_blocks=$mb"; ODEINTR_FUSE=$fuse ODEINTR_WARP_CPT=$cpt ODEINTR_WARP_MIN_BLOCKS=$mb timeout 120 python scripts/r02/prof_warp.py chain 2>&1 | tail -1
done; done; done
