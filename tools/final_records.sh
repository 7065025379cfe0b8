#!/bin/bash
# tools/final_records.sh TAG -- everything the round's records come from, on one B200 (run under gpurun):
# GPU tests, bench lines of every workload + the reference arm, the ncu launch list and one full capture of the
# default workload, and the division/sqrt bit-identity check. Outputs go to gpurun_out/TAG_*.
TAG=${1:-final}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/${TAG}_gpu_tests.log 2>&1; tail -2 $O/${TAG}_gpu_tests.log
tools/divcheck 64 > $O/${TAG}_divcheck.txt 2>&1; cat $O/${TAG}_divcheck.txt
python bench.py > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err; cut -c1-400 $O/${TAG}_bench_c3.json
for w in c1 c2 c4 c5; do
  python bench.py --workload $w > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err
  python -c "import json;d=json.load(open('$O/${TAG}_bench_$w.json'));print('$w',d['value'],d['e2e']['value'],d['roofline']['frac'],d['parity_mismatches_vs_oracle_sample'])"
done
python bench.py --impl reference > $O/${TAG}_bench_reference_c3.json 2> $O/${TAG}_bench_reference_c3.err; cut -c1-300 $O/${TAG}_bench_reference_c3.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_c3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/${TAG}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_check --launch-skip 1 --launch-count 1 -f -o $O/${TAG}_prof_c3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/${TAG}_ncu_full.log 2>&1; tail -1 $O/${TAG}_ncu_full.log
python tools/big_parity.py 24 256 > $O/${TAG}_big_parity.log 2>&1; tail -1 $O/${TAG}_big_parity.log
