#!/bin/bash
# Round-end measurement on one B200 (run under gpurun): GPU tests, smoke, bench lines of every workload, the reference
# arm, the ncu launch list of the bench command and one `ncu --set full` capture per shape.  Everything lands in
# gpurun_out/<tag>_*; copy what should be judged into profiles/.   usage: bash tools/final_measure.sh <tag>
tag=${1:-final}
out=gpurun_out
mkdir -p $out
python -m pytest tests -x -q -m gpu > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_gpu.log
python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> $out/${tag}_smoke.log
t0=$(date +%s)
python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_cfg2.json 2> $out/${tag}_bench_cfg2.err
echo "bench.py --steps 20 --warmup 5: wall $(( $(date +%s) - t0 )) s" > $out/${tag}_bench_cfg2.time
for w in cfg1 cfg3 cfg4 cfg5; do
  python bench.py --steps 20 --warmup 5 --workload $w --no-strong --no-train > $out/${tag}_bench_$w.json 2> $out/${tag}_bench_$w.err
done
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_reference_arm.err
# profiler passes come last: nothing above ran under ncu
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_ncu_launches.csv \
    python bench.py --steps 2 --warmup 1 --lean > $out/${tag}_ncu_launches.log 2>&1
for w in cfg2 cfg3 cfg5; do
  n=5; [ $w = cfg2 ] && n=3
  ncu --set full --clock-control none --import-source on -k regex:lct_kernel -s $((2 * n)) -c $n -f -o $out/${tag}_$w \
      python tools/profile_cases.py $w > $out/${tag}_ncu_$w.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:lct_kernel -s 6 -c 3 -f -o $out/${tag}_cfg2_fp \
    python tools/profile_cases.py cfg2 --fp > $out/${tag}_ncu_cfg2_fp.log 2>&1
tail -2 $out/${tag}_pytest_gpu.log; tail -3 $out/${tag}_smoke.log; python tools/bench_summary.py $out/${tag}_bench_cfg*.json 2>/dev/null
ls -la $out | grep $tag
