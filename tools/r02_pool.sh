#!/bin/bash
# tools/r02_pool.sh TAG -- under gpurun: GPU tests, then device-timed lines of the k_pool workloads (c3, c4, c5)
TAG=${1:-r02p}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/${TAG}_gpu_tests.log 2>&1; tail -3 $O/${TAG}_gpu_tests.log
for w in ${WLS:-c3 c4 c5}; do
  timeout 600 python bench.py --workload $w --steps ${STEPS:-10} --no-extras --no-cpu-baseline --no-e2e > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err
  tail -2 $O/${TAG}_bench_$w.err
  python -c "import json;d=json.load(open('$O/${TAG}_bench_$w.json'));print('$w','%.4g'%d['value'],d['ms_per_step'],d['parity_mismatches_vs_oracle_sample'],d['summary']['pair_counts'])"
done
