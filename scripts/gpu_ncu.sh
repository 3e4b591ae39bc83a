#!/bin/bash
# one full ncu capture of the dominant kernel (after a plain run of the same command)
mkdir -p gpurun_out
export HRT_BENCH_RAYS=${RAYS:-2e6} HRT_REF_PATHS=100
python bench.py --steps 1 --warmup 0 > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-k_scatter} -s ${SKIP:-1} -c ${COUNT:-1} -f -o gpurun_out/prof \
    python bench.py --steps 1 --warmup 0 > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_full.log
