#!/bin/bash
# Probe every library variant under build/variants/: tools/run_variants.sh [N K H]
for f in build/variants/lib_*.so; do
  n=$(basename $f .so)
  echo "== $n: $(P3D_LIB=$PWD/$f python tools/probe_sweep.py ${1:-512} ${2:-1024} ${3:-1024} all 2>&1 | grep '^sweep' | tail -1)"
done
