#!/bin/bash
# tools/r02_final.sh TAG -- the round's records on one B200 (run under gpurun): GPU tests, the bench line of every workload and
# of the reference arm, the ncu launch list of the default command, one `ncu --set full` capture of the check kernel of every
# workload and of k_feas, the big parity runs. Outputs: gpurun_out/TAG_*. Numbers printed under ncu are never bench values.
TAG=${1:-r02z}
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/${TAG}_gpu_tests.log 2>&1; tail -3 $O/${TAG}_gpu_tests.log
timeout 900 python bench.py > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err; cut -c1-300 $O/${TAG}_bench_c3.json; tail -3 $O/${TAG}_bench_c3.err
for w in c1 c2 c4 c5; do
  timeout 900 python bench.py --workload $w --no-extras > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err; tail -3 $O/${TAG}_bench_$w.err
  python -c "import json;d=json.load(open('$O/${TAG}_bench_$w.json'));print('$w','%.4g'%d['value'],'e2e %.4g'%d['e2e']['value'],'%.3f'%d['roofline']['frac'],d['roofline']['kernel'],d['parity_mismatches_vs_oracle_sample'],d.get('known_answer',{}).get('match'))"
done
for w in c3 c1 c2; do
  timeout 1500 python bench.py --impl reference --workload $w --steps 3 --warmup 3 > $O/${TAG}_bench_reference_$w.json 2> $O/${TAG}_bench_reference_$w.err; cut -c1-200 $O/${TAG}_bench_reference_$w.json
done
# launch list of the default command (every kernel launch with its device time; cold-cache, serialised)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches_c3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/${TAG}_ncu_list.log 2>&1
# full captures: the check kernel of every workload (second launch), k_feas
for w in c3 c5 c4 c2 c1; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^(k_pool|k_check)$' --launch-skip 1 --launch-count 1 -f -o $O/${TAG}_prof_$w python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $O/${TAG}_ncu_full_$w.log 2>&1; tail -1 $O/${TAG}_ncu_full_$w.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_feas --launch-skip 1 --launch-count 1 -f -o $O/${TAG}_prof_feas python tools/feas_bench.py 22 2 > $O/${TAG}_ncu_full_feas.log 2>&1; tail -1 $O/${TAG}_ncu_full_feas.log
timeout 900 python tools/big_parity.py 23 64 > $O/${TAG}_big_parity_m64.log 2>&1; tail -1 $O/${TAG}_big_parity_m64.log
timeout 900 python tools/big_parity.py 22 256 > $O/${TAG}_big_parity_m256.log 2>&1; tail -1 $O/${TAG}_big_parity_m256.log
timeout 900 python tools/big_parity.py 24 32 moving > $O/${TAG}_big_parity_moving.log 2>&1; tail -1 $O/${TAG}_big_parity_moving.log
timeout 900 python tools/big_parity.py 25 5 forest > $O/${TAG}_big_parity_forest.log 2>&1; tail -1 $O/${TAG}_big_parity_forest.log
# gpurun brings back at most 64 MiB of gpurun_out/: summarise the captures here (profiles/ of THIS copy), hand the summaries
# back in gpurun_out/collected/, and keep only the c3 report itself
python tools/r02_collect.py $TAG | tail -8
mkdir -p $O/collected
cp profiles/${TAG}_ncu_*.md profiles/ncu_traffic.json profiles/${TAG}_sass_kernels.txt $O/collected/
rm -f $O/${TAG}_prof_c[1245].ncu-rep $O/${TAG}_prof_feas.ncu-rep
du -sh $O
