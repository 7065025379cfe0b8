#!/bin/bash
# tools/r02_sweep2.sh TAG -- under gpurun: c3 (and c5) at a few trip thresholds and CTA sizes of k_pool
TAG=${1:-r02w}
O=gpurun_out
mkdir -p $O
run() {  # name, workload, env...
  local name=$1 w=$2; shift 2
  env "$@" timeout 300 python bench.py --workload $w --steps ${STEPS:-10} --no-extras --no-cpu-baseline --no-e2e > $O/${TAG}_$name.json 2> $O/${TAG}_$name.err
  python -c "import json;d=json.load(open('$O/${TAG}_$name.json'));print('$name','%.4g'%d['value'],d['ms_per_step'],d['parity_mismatches_vs_oracle_sample'])"
}
run c3_default c3 A=1
run c3_32_16 c3 RQCD_TH_SOLVE=32 RQCD_TH_PLANE=16
run c3_28_24 c3 RQCD_TH_SOLVE=28 RQCD_TH_PLANE=24
run c3_24_20 c3 RQCD_TH_SOLVE=24 RQCD_TH_PLANE=20
run c3_32_32 c3 RQCD_TH_SOLVE=32 RQCD_TH_PLANE=32
run c3_t384 c3 RQCD_POOL_THREADS=384
run c3_t512 c3 RQCD_POOL_THREADS=512
STEPS=3 run c5_28_24 c5 RQCD_TH_SOLVE=28 RQCD_TH_PLANE=24
STEPS=3 run c5_32_16 c5 RQCD_TH_SOLVE=32 RQCD_TH_PLANE=16
run c4_32_16 c4 RQCD_TH_SOLVE=32 RQCD_TH_PLANE=16
run c4_28_24 c4 RQCD_TH_SOLVE=28 RQCD_TH_PLANE=24
