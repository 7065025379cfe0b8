#!/bin/bash
# tools/r02_quick.sh TAG -- parity tests of the table kernels, c3 at a few trip thresholds, one ncu capture (under gpurun)
TAG=${1:-r02d}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/${TAG}_gpu_tests.log 2>&1; tail -3 $O/${TAG}_gpu_tests.log
for th in ${THS:-"32 24" "32 16" "28 28"}; do
  set -- $th
  RQCD_TH_SOLVE=$1 RQCD_TH_PLANE=$2 timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-e2e > $O/${TAG}_c3_$1_$2.json 2> $O/${TAG}_c3_$1_$2.err
  python -c "import json;d=json.load(open('$O/${TAG}_c3_$1_$2.json'));print('c3 th',$1,$2,'%.4g'%d['value'],d['ms_per_step'],d['parity_mismatches_vs_oracle_sample'],d['summary']['pair_counts'])"
done
for w in ${WLS:-c5 c2 c4}; do
  timeout 600 python bench.py --workload $w --steps 5 --no-cpu-baseline --no-e2e > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err
  python -c "import json;d=json.load(open('$O/${TAG}_bench_$w.json'));print('$w','%.4g'%d['value'],d['roofline']['frac'],d['parity_mismatches_vs_oracle_sample'],d['summary'])"
done
bash tools/r02_ncu.sh $TAG c3
