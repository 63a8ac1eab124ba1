#!/bin/bash
# bench value under a list of environment settings: tools/sweep.sh "A=1 B=2" "A=3" ...
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg python bench.py --steps 6 --no-cpu --no-config64 2>gpurun_out/sweep.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.1f GS/s  ms/step %.3f  kernels %s chunks %d' % (d['value']/1e3, d['ms_per_step'], {k:round(v,3) for k,v in d['kernels_ms_per_step'].items() if k!='note'}, d['config']['pipeline_chunks_per_step']))" || tail -3 gpurun_out/sweep.err
done
