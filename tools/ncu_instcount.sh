#!/bin/bash
# executed warp instructions per kernel launch for the given libraries (development A/B): tools/ncu_instcount.sh out.txt lib...
OUT=$1; shift
mkdir -p gpurun_out; : > gpurun_out/$OUT
for v in "$@"; do
  echo "== $v" >> gpurun_out/$OUT
  ALPB200_LIB=$v ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:'decode_kernel|decode_sum_kernel|encode_kernel' -c 40 --csv python tools/probe_dec.py ${LG:-26} 2>/dev/null | python -c "
import csv,sys
rows=[r for r in csv.reader(sys.stdin) if len(r)>10]
seen={}
for r in rows[1:]:
    k=(r[4][:60], r[-3]); seen.setdefault(k,[]).append(r[-1])
for k,v in seen.items(): print(k[0], k[1], v[-1])
" >> gpurun_out/$OUT
done
cat gpurun_out/$OUT
