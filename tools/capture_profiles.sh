#!/bin/bash
# Round-end evidence for profiles/rNN (run on the GPU box through gpurun; everything lands in gpurun_out/$1):
#   launches.csv          ncu launch list of `bench.py --steps 2 --warmup 1` (per-launch durations, cold and serialised)
#   ncu_summary.json      ncu --set full of ONE step's kernels (third step of tools/one_step.py), summarised on the box
#   traffic.json          DRAM bytes per kernel and per step from the same capture, keyed by the hash of csrc/
#   role_waits.log        cycles each warp role of the backward kernels spent waiting (RNNTB200_PROF=1)
#   accuracy.json         16-bit path against the fp32 path at C3
#   bench_c3.json, bench_c3_reference_arm.json
# usage: gpurun -- 'tools/capture_profiles.sh r2f'
set -u
O=gpurun_out/$1
mkdir -p $O
K='regex:^(alpha_beta|bwd_d|convert_w|finish_costs|joint_tc|row_scale|sum_p|tile_compact)'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-op-path > $O/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none -k "$K" -s 22 -c 11 -o $O/step python tools/one_step.py 3 > $O/ncu_step.log 2>&1
python tools/ncu_summary.py $O/step.ncu-rep 1 $O/traffic.json > $O/ncu_summary.json 2> $O/ncu_summary.err
rm -f $O/step.ncu-rep        # > 64 MiB: the summary travels instead
timeout 300 python tools/role_profile.py > $O/role_waits.log 2>&1
timeout 600 python tools/accuracy_c3.py > $O/accuracy.json 2> $O/accuracy.err
timeout 600 python bench.py > $O/bench_c3.json 2> $O/bench_c3.err
timeout 900 python bench.py --impl reference > $O/bench_c3_reference_arm.json 2> $O/bench_ref.err
ls -la $O
