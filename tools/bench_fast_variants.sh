#!/bin/bash
# Builds tuning variants of the library (same sources, -D macros) and times the pair kernels + the step of each
# on the GPU.  Usage (on the GPU box): bash tools/bench_fast_variants.sh "NAME:-DX=1,-DY=2" ...
mkdir -p build gpurun_out
for spec in "$@"; do
  name=${spec%%:*}; defs=${spec#*:}
  python - "$name" "$defs" <<'PY'
import sys
sys.path.insert(0, ".")
from tcsfm_b200 import build
name, defs = sys.argv[1], sys.argv[2]
build.build_variant("build/libtcsfm_%s.so" % name, [d[2:] for d in defs.split(",") if d.startswith("-D")])
PY
done
for lib in default $(ls build/libtcsfm_*.so 2>/dev/null); do
  if [ "$lib" != default ]; then export TCSFM_B200_LIB=$PWD/$lib; else unset TCSFM_B200_LIB; fi
  for arith in ${ARITHS:-fast}; do
    python bench.py --steps 60 --warmup 10 --only loss --no-cpu-baseline --arith $arith > gpurun_out/bench_var.json 2> gpurun_out/bench_var.err || tail -3 gpurun_out/bench_var.err
    python - "$lib" "$arith" <<'PY'
import json, sys
d = json.load(open("gpurun_out/bench_var.json"))
k = d["roofline"]["kernels"]
print("%-32s %-5s value %6d  step %.4f ms  fwd %.4f  bwd %.4f  e2e %6d" % (sys.argv[1], sys.argv[2], d["value"], d["ms_per_step"], k["pair_loss_fwd"]["avg_ms"], k["pair_loss_bwd"]["avg_ms"], d["e2e"]["value"]))
PY
  done
done
