#!/bin/bash
# tools/r02_final2.sh TAG -- the round's records once more, sized for what gpurun brings back (64 MiB): bench lines of every
# workload, the ncu launch list, one `ncu --set full` capture of the check kernel of every workload and of k_feas, summarised
# ON the box (tools/r02_collect.py) into gpurun_out/collected/; only the c3 report itself travels. (GPU tests, reference arms
# and the big parity runs: tools/r02_final.sh; their output of the same build is profiles/TAG_final_run_stdout.log.)
TAG=${1:-r02z}
O=gpurun_out
mkdir -p $O
timeout 300 python bench.py > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err; cut -c1-200 $O/${TAG}_bench_c3.json; tail -2 $O/${TAG}_bench_c3.err
for w in c1 c2 c4 c5; do
  timeout 240 python bench.py --workload $w --no-extras > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err; tail -2 $O/${TAG}_bench_$w.err
  python -c "import json;d=json.load(open('$O/${TAG}_bench_$w.json'));print('$w','%.4g'%d['value'],'e2e %.4g'%d['e2e']['value'],'%.3f'%d['roofline']['frac'],d['roofline']['kernel'],d['parity_mismatches_vs_oracle_sample'],d.get('known_answer',{}).get('match'))"
done
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches_c3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/${TAG}_ncu_list.log 2>&1
for w in c3 c4 c5 c1 c2; do
  timeout 150 ncu --set full --clock-control none --import-source on -k regex:'^(k_pool|k_check)$' --launch-skip 1 --launch-count 1 -f -o $O/${TAG}_prof_$w python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $O/${TAG}_ncu_full_$w.log 2>&1; tail -1 $O/${TAG}_ncu_full_$w.log
done
timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_feas --launch-skip 1 --launch-count 1 -f -o $O/${TAG}_prof_feas python tools/feas_bench.py 22 2 > $O/${TAG}_ncu_full_feas.log 2>&1; tail -1 $O/${TAG}_ncu_full_feas.log
python tools/r02_collect.py $TAG | tail -8
mkdir -p $O/collected
cp profiles/${TAG}_ncu_*.md profiles/ncu_traffic.json profiles/${TAG}_sass_kernels.txt $O/collected/
rm -f $O/${TAG}_prof_c[1245].ncu-rep $O/${TAG}_prof_feas.ncu-rep
du -sh $O
