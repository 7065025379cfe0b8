#!/bin/bash
# tools/r02_ncu.sh TAG [workload] -- one `ncu --set full` capture of the check kernel of a workload (run under gpurun)
TAG=${1:-r02c}; WL=${2:-c3}
O=gpurun_out
mkdir -p $O
ncu --set full --clock-control none --import-source on -k regex:'^(k_pool|k_check)$' --launch-skip 1 --launch-count 1 -f -o $O/${TAG}_prof_$WL python bench.py --workload $WL --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/${TAG}_ncu_full_$WL.log 2>&1; tail -1 $O/${TAG}_ncu_full_$WL.log
