"""The training-step sub-benchmark of bench.py alone (config 5), N times: python tools/bench_train.py [repeats]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

R = bench.Ranks()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    out = bench.train_measure(R, 10, 3)
    print(json.dumps({k: out[k] for k in ("value", "ms_per_step")}), json.dumps({k: out["eager"][k] for k in ("value", "ms_per_step")}))
R.close()
