#!/bin/bash
# ncu evidence of round 2 (one B200): launch list of a short bench run + one --set full capture per receive kernel
# (a full-size pipeline chunk: 65536 blocks = 33,554,432 wideband samples); reports land in gpurun_out/
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-config64 --no-tx --seconds 0 > gpurun_out/r02_launches_bench.log 2>&1
for k in analyzer8_kernel syncw_kernel packet_plain_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 5 -c 1 -f -o gpurun_out/r02_$k python tools/run_once.py 91 1 > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
