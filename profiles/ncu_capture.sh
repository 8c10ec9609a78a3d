#!/bin/bash
# usage (on the GPU box, from the repo root): profiles/ncu_capture.sh <tag> <main|c3|radial> [skip]
# One `ncu --set full` capture of the last MC launch of profiles/prof_run.py <what>; exports the raw metric page and
# the per-source-line page as CSV into gpurun_out/ and drops the .ncu-rep (gpurun brings back at most 64 MiB).
tag=$1; what=$2; skip=${3:-2}
python profiles/prof_run.py $what > gpurun_out/${tag}_prof_$what.log 2>&1 || { tail -5 gpurun_out/${tag}_prof_$what.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:mc_run -s $skip -c 1 -o /tmp/${tag}_$what -f \
    python profiles/prof_run.py $what > gpurun_out/${tag}_ncu_$what.log 2>&1
ncu -i /tmp/${tag}_$what.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_${what}_raw.csv 2>/dev/null
ncu -i /tmp/${tag}_$what.ncu-rep --page source --csv --print-source cuda 2>/dev/null | gzip > gpurun_out/${tag}_ncu_${what}_source_cuda.csv.gz
ncu -i /tmp/${tag}_$what.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > gpurun_out/${tag}_ncu_${what}_source_sass.csv.gz
ls -la gpurun_out/${tag}_ncu_${what}_*
