#!/bin/bash
# ncu --set full of one kernel of one profiles/bench_configs.py line (run under gpurun); exports the raw and source
# pages as csv on the box and drops the report (gpurun copies at most 64 MiB back):
#   profiles/profile_op.sh <tag> <kernel regex> <scale> <config name>
tag=$1; regex=$2; scale=$3; cfg=$4
CMD="python profiles/bench_configs.py $scale $cfg"
$CMD > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$regex -c 1 -f -o /tmp/${tag}_prof $CMD > gpurun_out/${tag}_ncu.log 2>&1
rc=$?
ncu -i /tmp/${tag}_prof.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ncu -i /tmp/${tag}_prof.ncu-rep --page source --csv > gpurun_out/${tag}_src.csv 2>/dev/null
echo "$tag rc=$rc"
