#!/bin/bash
# Capture one kernel with ncu --set full and keep only TEXT summaries under gpurun_out/ (details page, stall totals +
# hottest SASS lines + opcode mix, compact per-instruction counts for tools/sass_lines.py); the .ncu-rep and the raw source CSVs are deleted
# (gpurun_out/ is capped at 64 MiB).  usage: tools/ncu_one.sh <out-stem> <kernel-regex> <skip> <probe args...>
set -u
STEM=$1; KRE=$2; SKIP=$3; shift 3
mkdir -p gpurun_out
TMP=$(mktemp -d)
ncu --set full --import-source on --clock-control none -k regex:$KRE -s $SKIP -c 1 -f -o $TMP/rep python tools/ncu_probe.py "$@" > gpurun_out/${STEM}_run.log 2>&1
ncu -i $TMP/rep.ncu-rep --page details > gpurun_out/${STEM}_details.txt 2>&1
ncu -i $TMP/rep.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr, units, vals = rows[0], rows[1], rows[2]
for h, u, v in zip(hdr, units, vals):
    if h.startswith(('dram__bytes', 'gpu__time_duration', 'lts__t_bytes', 'sm__inst_executed.sum', 'smsp__inst_executed.sum', 'launch__grid_size', 'launch__registers', 'sm__warps_active.avg.pct')):
        print('%-60s %-14s %s' % (h, u, v))
" > gpurun_out/${STEM}_raw.txt
ncu -i $TMP/rep.ncu-rep --page source --csv > $TMP/sass.csv 2>/dev/null
N=$(( (1 << ${@: -1}) / 1024 ))
python tools/ncu_hot.py $TMP/sass.csv 30 > gpurun_out/${STEM}_hot.txt 2>&1
python tools/ncu_mix.py $TMP/sass.csv $N >> gpurun_out/${STEM}_hot.txt 2>&1
python tools/ncu_compact.py $TMP/sass.csv gpurun_out/${STEM}_counts.csv
echo "== $STEM"
grep -E "Duration|DRAM Throughput|Executed Ipc Active|Issue Slots Busy|Registers Per|Achieved Occupancy|Theoretical Occupancy|Eligible Warps|Active Warps Per Sch" gpurun_out/${STEM}_details.txt | head -12
head -12 gpurun_out/${STEM}_hot.txt
rm -rf $TMP
