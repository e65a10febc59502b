#!/bin/bash
# Bench every library variant under build/variants/ (built with build_native.build(defines=..., out=...)).
for f in build/variants/lib_*.so; do
  n=$(basename $f .so)
  P3D_LIB=$PWD/$f python bench.py --no-cpu-baseline --no-carve --no-extra --steps ${STEPS:-4} --warmup 3 2> gpurun_out/$n.err | python -c "import json,sys; d=json.load(sys.stdin); print('$n', d['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['best'])"
done
