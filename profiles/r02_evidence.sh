#!/bin/bash
# Round-2 evidence on ONE B200 (run from the repo root on the GPU box: gpurun -- 'bash profiles/r02_evidence.sh [what]').
# Everything lands in gpurun_out/ (scratch); the summaries that are cited are copied into profiles/ afterwards.
#   tests   full GPU suite                      -> r02_pytest_gpu.log
#   bench   both arms of the driver's bench     -> r02_bench.json, r02_bench_reference.json
#   ncu     launch list of bench.py + one `--set full` capture per kernel family
what=${1:-all}
O=gpurun_out
if [ "$what" = all ] || [ "$what" = tests ]; then
  python -m pytest tests -m gpu -q --durations=8 > $O/r02_pytest_gpu.log 2>&1; tail -3 $O/r02_pytest_gpu.log
fi
if [ "$what" = all ] || [ "$what" = bench ]; then
  python bench.py --impl reference > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err; tail -c 600 $O/r02_bench_reference.json
  python bench.py > $O/r02_bench.json 2> $O/r02_bench.err || tail -5 $O/r02_bench.err
  python - <<'PY'
import json
for l in open("gpurun_out/r02_bench.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print(d["value"], d["roofline"]["frac"], "e2e", d["e2e"]["value"], "strong", d["strong"]["value"], d["clocks"])
        print({k: v["value"] for k, v in d["secondary"].items()})
PY
fi
if [ "$what" = all ] || [ "$what" = ncu ]; then
  # launch list of the bench command (cold-cache, serialised: shares, not absolutes)
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_bench.csv \
      python bench.py --steps 2 --warmup 3 > $O/r02_ncu_bench.log 2>&1
  grep -c mcd $O/r02_launches_bench.csv
  # the bench launch itself (1024 chains x 200 MC steps): DRAM traffic per launch
  ncu --set full --clock-control none --import-source on -k regex:mc_run -s 1 -c 1 -o /tmp/r02_main -f \
      python profiles/prof_bench_launch.py > $O/r02_ncu_main.log 2>&1
  ncu -i /tmp/r02_main.ncu-rep --page raw --csv > $O/r02_ncu_full_mc_run_sq.csv 2>/dev/null
  bash profiles/ncu_capture.sh r02f c3
  bash profiles/ncu_capture.sh r02f radial
  for k in digitize count; do
    ncu --set full --clock-control none --import-source on -k regex:traj_$k -s 1 -c 1 -o /tmp/r02_traj_$k -f \
        python profiles/prof_traj.py once > $O/r02_ncu_traj_$k.log 2>&1
    ncu -i /tmp/r02_traj_$k.ncu-rep --page raw --csv > $O/r02_ncu_full_traj_$k.csv 2>/dev/null
  done
  ls -la $O | grep r02 | tail -20
fi
