#!/bin/bash
# Final profile of a round (one gpurun call): bench without ncu, then the launch list of the same command, then one
# `--set full` capture of the render kernel for each of the two bench workloads.  Outputs under gpurun_out/.
CMD="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
T=${1:-s2}
$CMD > gpurun_out/plain_$T.log 2> gpurun_out/plain_$T.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$T.csv $CMD > gpurun_out/ncu_l_$T.log 2>&1
$CMD > gpurun_out/plain2_$T.log 2> gpurun_out/plain2_$T.err &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:rt_fast::render_kernel<.int.128, .int.6, .bool.0, .bool.1, .bool.1>" -s 6 -c 1 \
    -f -o gpurun_out/prof_${T}_car_only $CMD > gpurun_out/ncu_f_$T.log 2>&1
$CMD > gpurun_out/plain3_$T.log 2> gpurun_out/plain3_$T.err &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:rt_fast::render_kernel<.int.128, .int.8, .bool.0, .bool.1, .bool.0>" -s 4 -c 1 \
    -f -o gpurun_out/prof_${T}_car_boxed_4k $CMD > gpurun_out/ncu_g_$T.log 2>&1
ls -la gpurun_out/*$T*
