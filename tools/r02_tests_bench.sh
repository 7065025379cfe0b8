#!/bin/bash
# tools/r02_tests_bench.sh TAG [workloads...] -- under gpurun: all GPU tests (no -x), then the default bench line and the
# listed extra workloads. Outputs: gpurun_out/TAG_*.
TAG=${1:-r02h}; shift
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/${TAG}_gpu_tests.log 2>&1; tail -25 $O/${TAG}_gpu_tests.log
timeout 900 python bench.py > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err; cut -c1-600 $O/${TAG}_bench_c3.json; tail -5 $O/${TAG}_bench_c3.err
for w in "$@"; do
  timeout 600 python bench.py --workload $w --steps 10 --no-extras > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err
  tail -3 $O/${TAG}_bench_$w.err
  python -c "import json;d=json.load(open('$O/${TAG}_bench_$w.json'));print('$w','%.4g'%d['value'],'e2e %.4g'%d['e2e']['value'],d['roofline']['frac'],d['roofline']['kernel'],d['parity_mismatches_vs_oracle_sample'],d['summary'])"
done
