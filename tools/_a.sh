for pool in 0 1; do for w in c2 c4; do
RQCD_POOL=$pool timeout 300 python bench.py --workload $w --steps 10 --no-extras --no-cpu-baseline --no-e2e > gpurun_out/r02k_${w}_pool$pool.json 2> gpurun_out/r02k_${w}_pool$pool.err
python -c "import json;d=json.load(open('gpurun_out/r02k_${w}_pool$pool.json'));print('$w pool$pool','%.4g'%d['value'],d['roofline']['frac'],d['roofline']['kernel'])"
done; done
