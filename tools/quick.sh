#!/bin/bash
# one GPU round trip while iterating on a kernel: parity subset, per-launch ncu numbers of one kernel, bench value
#   tools/quick.sh <kernel regex> [launch index to capture]
K=${1:-syncw_kernel}; S=${2:-5}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size \
    --clock-control none -k regex:$K -s $S -c 1 python tools/run_once.py 91 1 2>&1 | grep -E "duration|inst_executed|issue_active|warps_active|registers|grid_size|frames"
python bench.py --steps 6 --no-cpu --no-config64 2>gpurun_out/q.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.1f GS/s  e2e %.2f GS/s  ms/step %.3f  kernels %s' % (d['value']/1e3, d['e2e']['value']/1e3, d['ms_per_step'], {k:round(v,3) for k,v in d['kernels_ms_per_step'].items() if k!='note'}))"
tail -3 gpurun_out/q.err
