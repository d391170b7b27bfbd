#!/bin/bash
# K4w parameter sweep: cells per lane and resident blocks per SM (bistable = 1 field, chain = 2 fields)
for cpt in 5 7 9 11; do for mb in 2 3 4; do
  echo "== bistable cpt=$cpt min_blocks=$mb"; ODEINTR_WARP_CPT=$cpt ODEINTR_WARP_MIN_BLOCKS=$mb timeout 120 python scripts/r02/prof_warp.py bistable 2>&1 | tail -1
done; done
for cpt in 3 5 7; do for mb in 1 2 3; do
  echo "== chain cpt=$cpt min_blocks=$mb"; ODEINTR_WARP_CPT=$cpt ODEINTR_WARP_MIN_BLOCKS=$mb timeout 120 python scripts/r02/prof_warp.py chain 2>&1 | tail -1
done; done
