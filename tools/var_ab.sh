#!/bin/bash
# A/B of library variants built into liquid-usrp_b200/build/var/lib_<name>.so (scratch copies on the GPU box):
#   gpurun -- 'bash tools/var_ab.sh "python tools/c3_rate.py 5" name1 name2 ...'
cmd="$1"; shift
cd "$(dirname "$0")/.."
cp liquid-usrp_b200/libb200ofdm.so /tmp/lib_main.so
echo "== main"; $cmd 2>&1 | tail -1
for v in "$@"; do
  cp liquid-usrp_b200/build/var/lib_$v.so liquid-usrp_b200/libb200ofdm.so
  echo "== $v"; $cmd 2>&1 | tail -1
done
cp /tmp/lib_main.so liquid-usrp_b200/libb200ofdm.so
