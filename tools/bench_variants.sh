#!/bin/bash
# Times the pair kernels (and the whole step) for every tuning build under build/.
shopt -s nullglob
for lib in default build/libtcsfm_*.so; do
  if [ "$lib" != default ]; then export TCSFM_B200_LIB=$PWD/$lib; else unset TCSFM_B200_LIB; fi
  python bench.py --steps 60 --warmup 10 --no-cpu-baseline > gpurun_out/bench_var.json 2> gpurun_out/bench_var.err || tail -3 gpurun_out/bench_var.err
  python - "$lib" <<'PY'
import json, sys
d = json.load(open("gpurun_out/bench_var.json"))
k = d["roofline"]["kernels"]
print("%-28s value %6d  step %.4f ms  fwd %.4f  bwd %.4f" % (sys.argv[1], d["value"], d["ms_per_step"], k["pair_loss_fwd"]["avg_ms"], k["pair_loss_bwd"]["avg_ms"]))
PY
done
