#!/bin/bash
# Round-2 final evidence run (one B200): bench lines, sweeps, launch list, ncu captures, timelines, sanitizer.  Outputs under gpurun_out/r02_final2/.
O=gpurun_out/r02_final2; mkdir -p $O
python bench.py > $O/r02_bench_line_train.json 2> $O/bench_train.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_line_reference_arm.json 2> $O/bench_ref.err
python bench.py --workload inference --lean > $O/r02_bench_line_inference.json 2> $O/bench_inf.err
python bench.py --workload sun_train --lean > $O/r02_bench_line_sun_train.json 2> $O/bench_sun.err
python bench.py --workload sweep --math tf32 > $O/r02_sweep_tf32.json 2> $O/sweep_tf32.err
python bench.py --workload sweep --math 3xtf32 > $O/r02_sweep_3xtf32.json 2> $O/sweep_3x.err
python bench.py --lean --height 64 --width 256 --steps 10 --warmup 3 > $O/r02_scale_config5_64x256_1gpu.json 2> $O/c5.err
python bench.py --lean --trace-out $O/r02_step_timeline_3xtf32.json > /dev/null 2>&1
python tools/step_phases.py > $O/r02_step_phases_3xtf32.txt 2>&1
N="ncu --set full --clock-control none --import-source on"
$N -k regex:strip_wgrad -s 2 -c 1 -o $O/strip_wgrad_trunk_bench python tools/run_da_wgrad.py 32 8 32 128 128 3 4 > /dev/null 2>&1
$N -k regex:strip_wgrad -s 1 -c 1 -o $O/strip_wgrad_trunk_64x256 python tools/run_da_wgrad.py 64 64 256 128 128 3 3 > /dev/null 2>&1
$N -k regex:strip_wgrad -s 2 -c 1 -o $O/strip_wgrad_k7 python tools/run_da_wgrad.py 32 32 128 32 32 7 4 > /dev/null 2>&1
$N -k regex:strip_conv -s 2 -c 1 -o $O/strip_fwd_trunk_bench_3xtf32 python tools/run_da_layer.py 32 8 32 128 128 3 3xtf32 4 > /dev/null 2>&1
$N -k regex:strip_conv -s 1 -c 1 -o $O/strip_fwd_trunk_64x256_tf32 python tools/run_da_layer.py 64 64 256 128 128 3 tf32 3 > /dev/null 2>&1
for r in strip_wgrad_trunk_bench strip_wgrad_trunk_64x256 strip_wgrad_k7 strip_fwd_trunk_bench_3xtf32 strip_fwd_trunk_64x256_tf32; do
  ncu -i $O/$r.ncu-rep --page raw --csv > $O/$r.raw.csv 2>/dev/null
done
python tools/trace_wgrad.py 64 64 256 128 128 3 > $O/r02_wgrad_timeline_64x256.txt 2>&1
python tools/trace_wgrad.py 32 8 32 128 128 3 > $O/r02_wgrad_timeline_bench.txt 2>&1
( python tools/run_da_wgrad.py 32 8 32 128 128 3 5; python tools/run_da_wgrad.py 32 32 128 32 32 7 5; python tools/run_da_wgrad.py 64 32 128 128 128 3 4; python tools/run_da_wgrad.py 64 64 256 128 128 3 4; python tools/run_da_wgrad.py 64 128 512 128 128 3 3 ) > $O/r02_wgrad_layer_timings.txt 2>&1
compute-sanitizer --tool memcheck python tools/run_da_wgrad.py 4 8 32 128 128 3 1 > $O/r02_compute_sanitizer_memcheck_wgrad.log 2>&1
compute-sanitizer --tool memcheck python tools/run_da_wgrad.py 4 16 64 32 32 7 1 >> $O/r02_compute_sanitizer_memcheck_wgrad.log 2>&1
compute-sanitizer --tool racecheck python tools/run_da_wgrad.py 4 8 32 128 128 3 1 > $O/r02_compute_sanitizer_racecheck_wgrad.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/r02_launch_list_bench_train.csv python bench.py --lean --steps 2 --warmup 1 > $O/ncu_bench.log 2>&1
ls -la $O | head -60
