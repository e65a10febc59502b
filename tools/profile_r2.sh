#!/bin/bash
# Round-2 ncu captures (run under gpurun; summaries are made afterwards with tools/ncu_summary.py)
set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --no-carve --no-sweep-total"
$B > gpurun_out/r2_plain_bench.json 2> gpurun_out/r2_plain_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:splat_seg -s 40 -c 1 -o gpurun_out/r2_splat_seg $B > /dev/null 2>&1
ncu --set full --clock-control none -k regex:score_kernel -s 40 -c 1 -o gpurun_out/r2_score $B > /dev/null 2>&1
P="python tools/probe_sweep.py 1024 64 2048 all"
$P > gpurun_out/r2_plain_1024.log 2>&1 || exit 1
ncu --set full --clock-control none -k regex:splat_seg -s 4 -c 1 -o gpurun_out/r2_splat_seg_1024 $P > /dev/null 2>&1
ncu --set full --clock-control none -k regex:score_kernel -s 4 -c 1 -o gpurun_out/r2_score_1024 $P > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
