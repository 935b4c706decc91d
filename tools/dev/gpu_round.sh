#!/bin/bash
# Development helper, run ON the GPU box (gpurun -- 'bash tools/dev/gpu_round.sh <tag>'): the measurement set behind profiles/*_<tag>.*
# GPU tests, the default bench line, the reference arm, the phase-clock tables, the ncu launch list and one `ncu --set full` capture per hot
# kernel at the benchmarked batch (exported as raw CSV: the .ncu-rep files exceed the 64 MiB return limit), the secondary bench configurations.
t=${1:-dev}
o=gpurun_out
mkdir -p $o
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > $o/${t}_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > $o/bench_${t}.json 2> $o/${t}_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > $o/bench_${t}_reference.json 2> $o/${t}_ref.err
timeout 300 python tools/phase_clock.py --node 37 --batch 256 > $o/phase_clock_${t}_node37.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $o/launches_${t}.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $o/${t}_ll.log 2>&1
# -s: launches of that kernel to skip.  One solve launches every LQ / QP kernel once and rollout_kernel 14 times (the whole back-tracking
# ladder): skip the warm-up solve (+ the upload solve) so that the capture is a steady-state launch; rollout: the FIRST trial of a solve.
for k in ricb:riccati_bwd:2 ricf:riccati_fwd:2 lqa:lq_dyn_kernel:2 lqp:lq_projdyn_kernel:2 lqb:lq_proj_kernel:2 lu:lu_kernel:2 ro:rollout_kernel:28; do
  IFS=: read s n skip <<< "$k"
  timeout 400 ncu --set full --clock-control none -k regex:$n -s $skip -c 1 -f -o /tmp/prof_${s} python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $o/${t}_ncu_$s.log 2>&1
  ncu -i /tmp/prof_${s}.ncu-rep --page raw --csv > $o/ncuraw_${s}_${t}.csv 2>/dev/null
done
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --batch 1 > $o/bench_${t}_batch1.json 2>/dev/null
timeout 300 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --sqp-iteration 10 > $o/bench_${t}_sqp10.json 2>/dev/null
timeout 300 python tools/bench_receding_device.py --cycles 120 > $o/bench_${t}_receding_device.json 2>/dev/null
tail -n 3 $o/${t}_tests.log
