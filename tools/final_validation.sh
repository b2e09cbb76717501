#!/bin/bash
# Round-end evidence run on one B200 (gpurun): GPU tests, both bench arms, ncu launch list + ncu --set full of the convs, sweep.
T=${1:-final2}
timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; echo bench rc=$?
timeout 200 python bench.py --impl reference > gpurun_out/bench_${T}_ref.json 2> gpurun_out/bench_${T}_ref.err; echo ref rc=$?
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/bench_launches_$T.csv python bench.py --steps 10 --warmup 3 --train-steps 0 --no-cpu-baseline > gpurun_out/ncu_bench_$T.log 2>&1; echo ncu1 rc=$?
timeout 300 ncu --set full --clock-control none -k regex:"conv_first|conv_halo|conv_tc" -s 22 -c 22 -o /tmp/prof_$T python tools/profile_forward.py 64 2 > gpurun_out/ncu_fwd_full_$T.log 2>&1; echo ncu2 rc=$?
ncu -i /tmp/prof_$T.ncu-rep --page raw --csv > gpurun_out/prof_${T}_raw.csv 2>/dev/null; ls -la /tmp/prof_$T.ncu-rep gpurun_out/prof_${T}_raw.csv
timeout 200 python tools/sweep.py > gpurun_out/sweep_$T.jsonl 2> gpurun_out/sweep_$T.err; echo sweep rc=$?; tail -3 gpurun_out/sweep_$T.jsonl
