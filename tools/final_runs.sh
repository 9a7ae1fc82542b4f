#!/bin/bash
# Round-end measurement set on one B200: parity suite, smoke, the default bench line (both arithmetic flavours, PFT,
# sequence, training step, eager-CUDA baseline), the other workloads, the reference arm, per-operator timings, the ncu
# launch lists and full ncu captures of the pair kernels (each ncu pass only after the plain run exited 0).
# Usage (on the GPU box): bash tools/final_runs.sh [outdir]
set -u
out=${1:-gpurun_out/final}; mkdir -p $out
python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee $out/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee $out/smoke.txt
( time python bench.py > $out/bench_default.json 2> $out/bench_default.err ) 2> $out/bench_default.time || tail -5 $out/bench_default.err
python bench.py --arith exact --only loss --no-cpu-baseline --steps 200 > $out/bench_exact.json 2> $out/bench_exact.err || tail -5 $out/bench_exact.err
python bench.py --profile train --only loss --no-cpu-baseline --steps 200 > $out/bench_profile_train.json 2> $out/bench_train.err || tail -5 $out/bench_train.err
for w in scannet scannet448 kitti376x4; do
  python bench.py --workload $w --only loss --no-cpu-baseline --steps 200 > $out/bench_$w.json 2> $out/bench_$w.err || tail -5 $out/bench_$w.err
done
python bench.py --impl reference --steps 5 --warmup 1 > $out/bench_reference_arm.json 2> $out/bench_reference.err || tail -5 $out/bench_reference.err
python tools/bench_ops.py > $out/bench_ops.json 2> $out/bench_ops.err || tail -5 $out/bench_ops.err
# launch list of the bench command itself (loss step only: warm-up, graph capture, replays, the eager per-kernel pass, e2e)
python bench.py --only loss --no-cpu-baseline --steps 3 --warmup 3 > $out/bench_short.json 2> $out/bench_short.err && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/launches_bench.csv \
      python bench.py --only loss --no-cpu-baseline --steps 3 --warmup 3 > $out/ncu_bench.log 2>&1
for arith in fast exact; do
  TCSFM_ARITH=$arith python tools/profile_step.py 3 2 > $out/prof_plain_$arith.log 2>&1 && \
    TCSFM_ARITH=$arith ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$arith.csv \
      python tools/profile_step.py 3 2 > $out/ncu_list_$arith.log 2>&1
  TCSFM_ARITH=$arith python tools/profile_step.py 2 1 > $out/prof_plain2_$arith.log 2>&1 && \
    TCSFM_ARITH=$arith ncu --set full --clock-control none --import-source on -k regex:pair_ -c 2 -s 2 -o $out/prof_$arith -f \
      python tools/profile_step.py 2 1 > $out/ncu_full_$arith.log 2>&1
  ncu -i $out/prof_$arith.ncu-rep --page raw --csv > $out/prof_${arith}_raw.csv 2>/dev/null
  ncu -i $out/prof_$arith.ncu-rep --page source --csv --print-source cuda,sass > $out/prof_${arith}_src.csv 2>/dev/null
done
# the PFT hot path proper (leaf disparities, no depth network): launch list of three epochs
python tools/profile_pft_hotpath.py 3 > $out/pft_hot_plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $out/launches_pft_hotpath.csv \
      python tools/profile_pft_hotpath.py 3 > $out/ncu_pft_hot.log 2>&1
for f in $out/bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    r = d.get("roofline") or {}
    print("%-40s value %10.1f %s  ms/step %.4f  e2e %.1f  frac %s (%s)" % (sys.argv[1].split("/")[-1], d["value"], d["unit"], d["ms_per_step"], (d.get("e2e") or {}).get("value", 0), round(r.get("frac", 0), 4), r.get("kernel")))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
cat $out/bench_default.time
