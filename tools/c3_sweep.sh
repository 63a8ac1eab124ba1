#!/bin/bash
# config64 receive rate over pipeline chunk sizes and Viterbi region counts: bash tools/c3_sweep.sh
cd "$(dirname "$0")/.."
for v in 48 72 96; do for c in 262144 131072; do
  echo -n "vit_ctas=$v chunk=$c: "; B2_VIT_CTAS=$v B2_CHUNK_BLOCKS=$c timeout 100 python tools/c3_rate.py 5 2>&1 | tail -1 | sed -E "s/.*'ms_per_step': ([0-9.]+).*'sync_kernel': np.float64\(([0-9.]+)\).*'packet_decode_kernel': np.float64\(([0-9.]+)\).*'call': np.float64\(([0-9.]+)\).*/step \1 sync \2 dec \3 call \4/"
done; done
