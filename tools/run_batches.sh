#!/bin/bash
# Cameras per splat launch (P3D_MAX_BATCH) against throughput: tools/run_batches.sh N K H "b1 b2 ..."
for b in $4; do
  echo "== batch $b: $(P3D_MAX_BATCH=$b python tools/probe_sweep.py $1 $2 $3 all 2>&1 | grep '^sweep' | tail -1)"
done
