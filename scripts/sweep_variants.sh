for v in mb1 b384x2 b256x3 b256x4; do
  export HRT_LIB=hermespy-rt_b200/build/var_$v/libhermespy_rt.so
  echo "== $v"
  HRT_BENCH_RAYS=2e7 HRT_REF_PATHS=100 timeout 600 python bench.py --steps 2 --warmup 1 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('C4 value %.4g ms %.1f' % (d['value'], d['ms_per_step']))"
  python scripts/run_c5.py 6.25e7 1024 256 65536 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('C5 q/s %.4g ms_scatter %.1f' % (d['closest_hit_queries_per_s'], d['ms_scatter']))"
done
