#!/bin/bash
# tools/r02_feas.sh TAG -- under gpurun: the feasibility parity tests, k_feas alone at the three occupancies, c1 / c2 lines
TAG=${1:-r02f}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -k "feasib or filter or performance_eval or forest" > $O/${TAG}_feas_tests.log 2>&1; tail -3 $O/${TAG}_feas_tests.log
for occ in 4 5 6; do
  echo "occ $occ: $(RQCD_FEAS_OCC=$occ timeout 300 python tools/feas_bench.py 22 10 2>&1 | tail -1)"
done
for w in ${WLS:-c1 c2}; do
  timeout 600 python bench.py --workload $w --steps 10 --no-extras --no-cpu-baseline > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err
  tail -3 $O/${TAG}_bench_$w.err
  python -c "import json;d=json.load(open('$O/${TAG}_bench_$w.json'));print('$w','%.4g'%d['value'],d['ms_per_step'],'e2e %.4g'%d['e2e']['value'],d['parity_mismatches_vs_oracle_sample'],d.get('known_answer',{}).get('match'))"
done
