#!/bin/bash
# One ncu --set full capture of one kernel + text summaries. Usage: scripts/ncu_one.sh <tag> <kernel-regex> <mangled-substring> <units> -- <bench args...>
# (the same command is run once without ncu first and must exit 0)
set -u
TAG=$1; RE=$2; MANGLED=$3; UNITS=$4; shift 5
O=gpurun_out; mkdir -p $O
python bench.py "$@" > $O/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 $O/${TAG}_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -f -k regex:"$RE" -c 1 -o $O/$TAG python bench.py "$@" > $O/${TAG}_ncu.log 2>&1
python tools/ncu_summary.py $O/$TAG.ncu-rep > $O/r2_ncu_$TAG.txt 2>&1
python tools/sass_by_line.py $O/$TAG.ncu-rep ofdm_b200/libofdm_b200.so "$MANGLED" $UNITS > $O/r2_ncu_${TAG}_by_line.txt 2>&1
ls -la $O/$TAG.ncu-rep; rm -f $O/$TAG.ncu-rep
head -60 $O/r2_ncu_$TAG.txt; head -45 $O/r2_ncu_${TAG}_by_line.txt
