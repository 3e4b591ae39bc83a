#!/bin/bash
# quick iteration: parity tests + small bench (+ optional env sweeps)
mkdir -p gpurun_out
if [ "${TESTS:-1}" = "1" ]; then
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
fi
for mode in ${MODES:-default}; do
  echo "== bench small mode=$mode"
  if [ "$mode" != "default" ]; then export HRT_SCATTER_MODE=$mode; else unset HRT_SCATTER_MODE; fi
  HRT_BENCH_RAYS=${RAYS:-8e6} HRT_REF_PATHS=100 timeout 600 python bench.py --steps 2 --warmup 1 > gpurun_out/bench_$mode.json 2> gpurun_out/bench_$mode.err; echo "rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$mode.json")); r=d["roofline"]
    print("value %.4g rb/s  ms/step %.1f  e2e %.4g  frac %.3f  achieved %.2f TF  box/q %.1f tri/q %.1f  share %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r["frac"], r["achieved"], r["box_tests_per_shadow_query"], r["tri_tests_per_shadow_query"], r["kernel_share_of_step"]))
except Exception as e: print("parse fail", e); print(open("gpurun_out/bench_$mode.err").read()[-2000:])
PY
done
