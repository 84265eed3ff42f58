#!/bin/bash
# Round-2 profile set (one gpurun call): bench without ncu, the launch list of the same command, one `--set full` capture of
# the render kernel of each bench workload (scripts/few_frames.py: the same library calls, default parameters), and one of the
# 50.1 M-triangle scene.
CMD="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ref-gpu"
T=${1:-r02}
python bench.py --steps 30 --warmup 5 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
$CMD > gpurun_out/${T}_plain.log 2> gpurun_out/${T}_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_ncu_l.log 2>&1
python scripts/few_frames.py car_only 1920 1080 8 > gpurun_out/${T}_plain2.log 2> gpurun_out/${T}_plain2.err &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:rt_fast::render_kernel" -s 6 -c 1 \
    -f -o gpurun_out/${T}_prof_car_only python scripts/few_frames.py car_only 1920 1080 8 > gpurun_out/${T}_ncu_f.log 2>&1
python scripts/few_frames.py car_boxed 3840 2160 6 > gpurun_out/${T}_plain3.log 2> gpurun_out/${T}_plain3.err &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:rt_fast::render_kernel" -s 4 -c 1 \
    -f -o gpurun_out/${T}_prof_car_boxed_4k python scripts/few_frames.py car_boxed 3840 2160 6 > gpurun_out/${T}_ncu_g.log 2>&1
python scripts/config5_one.py > gpurun_out/${T}_config5_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:rt_fast::render_kernel" -s 3 -c 1 \
    -f -o gpurun_out/${T}_prof_config5 python scripts/config5_one.py > gpurun_out/${T}_ncu_c5.log 2>&1
ls -la gpurun_out/${T}_*
