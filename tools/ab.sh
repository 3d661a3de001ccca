#!/bin/bash
# A/B of library variants in ONE gpurun call (same box): tools/ab.sh out_prefix libA.so libB.so ...   ("" = the in-tree library)
P=$1; shift
for round in 1 2; do
for L in "$@"; do
  if [ -n "$L" ]; then export MANUS_B200_LIB=$L; else unset MANUS_B200_LIB; fi
  echo "== ${L:-default} (round $round)" >> $P
  python tools/vif_sweep.py 1 4 >> $P 2>&1
done
done
