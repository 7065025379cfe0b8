#!/bin/bash
# tools/r02_full.sh TAG -- under gpurun: GPU tests, the default bench line, the ncu launch list of the same command and one
# `ncu --set full` capture of the check kernel of c3. Outputs: gpurun_out/TAG_*.
TAG=${1:-r02g}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_gpu_tests.log 2>&1; tail -15 $O/${TAG}_gpu_tests.log
timeout 600 python bench.py > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err; cut -c1-400 $O/${TAG}_bench_c3.json; tail -3 $O/${TAG}_bench_c3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_c3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/${TAG}_ncu_list.log 2>&1
bash tools/r02_ncu.sh $TAG c3
