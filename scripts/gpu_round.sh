#!/bin/bash
# One GPU session: parity tests, small + full bench, ncu launch list and one
# full capture of the dominant kernel.  Everything is logged to gpurun_out/.
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1

echo "== pytest -m gpu"; 
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -15 $OUT/pytest_gpu.log

echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
tail -3 $OUT/smoke.log

echo "== bench small (4e6 rays)"
HRT_BENCH_RAYS=4e6 HRT_REF_PATHS=200 HRT_BENCH_SKIP_DENSE=1 HRT_BENCH_SKIP_C5=1 timeout 600 python bench.py --steps 2 --warmup 1 > $OUT/bench_small.json 2> $OUT/bench_small.err; echo "rc=$?"
cat $OUT/bench_small.json | cut -c1-1500; tail -5 $OUT/bench_small.err

if [ "${FULL:-1}" = "1" ]; then
echo "== bench full (1e8 rays)"
timeout 1200 python bench.py --steps ${STEPS:-3} --warmup 3 > $OUT/bench_full.json 2> $OUT/bench_full.err; echo "rc=$?"
cat $OUT/bench_full.json | cut -c1-3000; tail -5 $OUT/bench_full.err
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "rc=$?"
cat $OUT/bench_ref.json | cut -c1-400
fi

echo "== dense configs + C5 shard"
python scripts/run_dense.py > $OUT/dense.json 2> $OUT/dense.err; echo "rc=$?"
python scripts/run_c5.py 6.25e7 1024 256 65536 > $OUT/c5_shard.json 2> $OUT/c5.err; echo "rc=$?"; cut -c1-600 $OUT/c5_shard.json

if [ "${NCU:-1}" = "1" ]; then
echo "== ncu launch list"
export HRT_BENCH_RAYS=8e6 HRT_REF_PATHS=100 HRT_BENCH_SKIP_DENSE=1 HRT_BENCH_SKIP_C5=1 HRT_BENCH_SKIP_INLIB=1 HRT_BENCH_SKIP_BVH_MODE=1
python bench.py --steps 1 --warmup 0 > $OUT/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 0 > $OUT/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
# SURVEY 8(d) cross-check: executed fp32 operations of k_scatter (ffma counts 2 flops), receiver maps and BVH mode
ncu --metrics smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,gpu__time_duration.sum \
    --clock-control none -k regex:k_scatter -c 5 --csv --log-file $OUT/fp_ops.csv python bench.py --steps 1 --warmup 0 > $OUT/ncu_fp.log 2>&1
HRT_RXMAP=0 ncu --metrics smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,gpu__time_duration.sum \
    --clock-control none -k regex:k_scatter -c 5 --csv --log-file $OUT/fp_ops_bvh.csv python bench.py --steps 1 --warmup 0 > $OUT/ncu_fp_bvh.log 2>&1
echo "ncu fp ops rc=$?"
python bench.py --steps 1 --warmup 0 > $OUT/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_scatter -s 1 -c 2 -f -o $OUT/prof_scatter \
    python bench.py --steps 1 --warmup 0 > $OUT/ncu_full.log 2>&1
echo "ncu full rc=$?"
python scripts/run_c5.py 6.25e7 1024 256 65536 > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_scatter -s 1 -c 1 -f -o $OUT/prof_c5 \
    python scripts/run_c5.py 6.25e7 1024 256 65536 > $OUT/ncu_c5.log 2>&1
echo "ncu c5 rc=$?"
ls -la $OUT
fi
