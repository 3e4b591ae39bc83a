#!/bin/bash
# env sweeps of the small bench: each line "VAR=val VAR=val" in $SWEEP (semicolon separated)
mkdir -p gpurun_out
IFS=';' read -ra CASES <<< "${SWEEP:-default}"
i=0
for c in "${CASES[@]}"; do
  i=$((i+1))
  echo "== case $i: $c"
  ( if [ "$c" != "default" ]; then export $c; fi
    HRT_BENCH_RAYS=${RAYS:-2e7} HRT_REF_PATHS=100 timeout 600 python bench.py --steps 2 --warmup 1 > gpurun_out/sweep_$i.json 2> gpurun_out/sweep_$i.err; echo "rc=$?" )
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sweep_$i.json")); r=d["roofline"]
    print("value %.4g rb/s  ms/step %.1f  frac %.3f  achieved %.2f TF  box/q %.1f tri/q %.1f  share %.3f" % (d["value"], d["ms_per_step"], r["frac"], r["achieved"], r["box_tests_per_shadow_query"], r["tri_tests_per_shadow_query"], r["kernel_share_of_step"]))
except Exception as e: print("parse fail", e); print(open("gpurun_out/sweep_$i.err").read()[-1500:])
PY
done
