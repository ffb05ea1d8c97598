timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for lib in "" build/lib_a.so build/lib_b.so; do
  echo "== lib '$lib'"
  HIDDENPOSE_LCT_LIB=$lib timeout 300 python tools/kbench.py cfg2 cfg3 cfg4 cfg5 cfg1 custom:32,256,64 --reps 30 2>&1 | grep "^\["
  HIDDENPOSE_LCT_LIB=$lib timeout 300 python tools/kbench.py cfg2 cfg4 --fp --reps 30 2>&1 | grep "^\["
done
