#!/bin/bash
# tools/r02_baseline.sh TAG -- profile the kernel that is benched (run under gpurun): GPU tests, c3 bench line, the ncu
# launch list of the same command and one `ncu --set full` capture of the dominant kernel of c3. Outputs: gpurun_out/TAG_*.
TAG=${1:-r02a}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/${TAG}_gpu_tests.log 2>&1; tail -2 $O/${TAG}_gpu_tests.log
python bench.py > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err; cut -c1-300 $O/${TAG}_bench_c3.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_c3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/${TAG}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_check --launch-skip 1 --launch-count 1 -f -o $O/${TAG}_prof_c3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/${TAG}_ncu_full.log 2>&1; tail -1 $O/${TAG}_ncu_full.log
