timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -1
for v in 1 0; do
  export LCT_L2_PREFETCH=$v
  for c in cfg4 cfg3 cfg5; do timeout 300 python bench.py --workload $c --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/pf_${v}_$c.json 2>/dev/null; echo "== prefetch $v $c"; python tools/show.py gpurun_out/pf_${v}_$c.json; done
done
