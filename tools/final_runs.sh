#!/bin/bash
# Round-end measurement set on one B200: parity suite, smoke, every bench workload, the ncu launch list
# and one full ncu capture of the pair kernels (each ncu pass only after the plain run exited 0).
set -u
out=gpurun_out/final; mkdir -p $out
python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee $out/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee $out/smoke.txt
python bench.py > $out/bench_default.json 2> $out/bench_default.err || tail -5 $out/bench_default.err
python bench.py --profile train --no-cpu-baseline --steps 1000 > $out/bench_profile_train.json 2> $out/bench_train.err || tail -5 $out/bench_train.err
for w in scannet scannet448 kitti376x4; do
  python bench.py --workload $w --no-cpu-baseline --steps 500 > $out/bench_$w.json 2> $out/bench_$w.err || tail -5 $out/bench_$w.err
done
python bench.py --workload pft --no-cpu-baseline > $out/bench_pft.json 2> $out/bench_pft.err || tail -5 $out/bench_pft.err
python tools/bench_ops.py > $out/bench_ops.json 2> $out/bench_ops.err || tail -5 $out/bench_ops.err
python tools/profile_step.py 3 2 > $out/prof_plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python tools/profile_step.py 3 2 > $out/ncu_list.log 2>&1
python tools/profile_step.py 2 1 > $out/prof_plain2.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:pair_ -c 2 -s 2 -o $out/prof_final -f python tools/profile_step.py 2 1 > $out/ncu_full.log 2>&1
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
