#!/bin/bash
# tools/variant_bench.sh -- time every kernel-variant build in variants/ on the default bench workload (device-resident only).
for f in rapidquadcoptercollisiondetection_b200/librqcd.so variants/librqcd_*.so; do
  [ -f "$f" ] || continue
  RQCD_LIB=$PWD/$f python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e ${BENCH_ARGS} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('$f', 'checks/s %.3e' % d['value'], 'ms %.3f' % d['ms_per_step'], 'mismatch', d['parity_mismatches_vs_oracle_sample'], 'frac %.3f' % d['roofline']['frac'], d['clocks']['sm_mhz'])
"
done
