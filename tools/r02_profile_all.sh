#!/bin/bash
# Round-2 evidence run (one B200): bench lines, sweeps, launch list, ncu captures, timeline, sanitizer.  Outputs under gpurun_out/r02_final/.
O=gpurun_out/r02_final; mkdir -p $O
python bench.py > $O/r02_bench_line_train.json 2> $O/bench_train.err
python bench.py --impl reference --steps 5 --warmup 1 > $O/r02_bench_line_reference_arm.json 2> $O/bench_ref.err
python bench.py --workload inference --lean > $O/r02_bench_line_inference.json 2> $O/bench_inf.err
python bench.py --workload sun_train --lean > $O/r02_bench_line_sun_train.json 2> $O/bench_sun.err
python bench.py --workload sweep --math tf32 > $O/r02_sweep_tf32.json 2> $O/sweep_tf32.err
python bench.py --workload sweep --math 3xtf32 > $O/r02_sweep_3xtf32.json 2> $O/sweep_3x.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r02_launch_list_bench_train.csv python bench.py --lean --steps 2 --warmup 1 > $O/ncu_bench.log 2>&1
N="ncu --set full --clock-control none --import-source on"
$N -k regex:strip_conv -s 2 -c 1 -o $O/strip_fwd_trunk_bench_tf32 python tools/run_da_layer.py 32 8 32 128 128 3 tf32 4 > /dev/null 2>&1
$N -k regex:strip_conv -s 2 -c 1 -o $O/strip_fwd_trunk_bench_3xtf32 python tools/run_da_layer.py 32 8 32 128 128 3 3xtf32 4 > /dev/null 2>&1
$N -k regex:strip_conv -s 1 -c 1 -o $O/strip_fwd_trunk_64x256_tf32 python tools/run_da_layer.py 64 64 256 128 128 3 tf32 3 > /dev/null 2>&1
$N -k regex:strip_conv -s 2 -c 1 -o $O/strip_fwd_k7_tf32 python tools/run_da_layer.py 32 32 128 32 32 7 tf32 4 > /dev/null 2>&1
$N -k regex:strip_conv_kernel -s 1 -c 1 -o $O/strip_dgrad_trunk_bench python tools/run_da_dgrad.py 32 8 32 128 128 3 3 > /dev/null 2>&1
$N -k regex:wgrad -s 2 -c 1 -o $O/wgrad_trunk_bench python tools/run_da_wgrad.py 32 8 32 128 128 3 4 > /dev/null 2>&1
python tools/trace_strip.py 32 8 32 128 128 3 tf32 > $O/r02_strip_timeline_trunk_tf32.txt 2>&1
python tools/run_da_layer.py 32 8 32 128 128 3 tf32 6 > $O/layer_timings.txt 2>&1
python tools/run_da_layer.py 32 8 32 128 128 3 3xtf32 6 >> $O/layer_timings.txt 2>&1
python tools/run_da_layer.py 32 32 128 32 32 7 tf32 6 >> $O/layer_timings.txt 2>&1
python tools/run_da_layer.py 32 32 128 32 32 7 3xtf32 6 >> $O/layer_timings.txt 2>&1
python tools/run_da_layer.py 64 64 256 128 128 3 tf32 4 >> $O/layer_timings.txt 2>&1
python tools/run_da_layer.py 64 64 256 128 128 3 tf32 4 band >> $O/layer_timings.txt 2>&1
python tools/run_da_layer.py 32 8 32 128 128 3 tf32 6 band >> $O/layer_timings.txt 2>&1
python tools/run_da_dgrad.py 32 8 32 128 128 3 5 >> $O/layer_timings.txt 2>&1
python tools/run_da_dgrad.py 32 32 128 32 32 7 5 >> $O/layer_timings.txt 2>&1
python tools/run_da_dgrad.py 64 64 256 128 128 3 4 >> $O/layer_timings.txt 2>&1
python tools/run_da_wgrad.py 32 8 32 128 128 3 5 >> $O/layer_timings.txt 2>&1
compute-sanitizer --tool memcheck python tools/run_da_layer.py 4 8 32 128 128 3 3xtf32 1 > $O/r02_compute_sanitizer_memcheck_strip.log 2>&1
compute-sanitizer --tool memcheck python tools/run_da_dgrad.py 4 8 32 64 64 3 2 >> $O/r02_compute_sanitizer_memcheck_strip.log 2>&1
compute-sanitizer --tool racecheck python tools/run_da_layer.py 4 8 32 64 64 3 tf32 1 > $O/r02_compute_sanitizer_racecheck_strip.log 2>&1
ls -la $O | head -50
