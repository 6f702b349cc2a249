#!/bin/bash
# Developer: ncu launch list (durations only) of the bench command -> gpurun_out/<tag>_launches_bench.{csv,txt}
tag=${1:-r02}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches_bench.csv $CMD > gpurun_out/${tag}_ncu_launch.log 2>&1
python tools/summarize_launches.py gpurun_out/${tag}_launches_bench.csv gpurun_out/${tag}_launches_bench.txt | tail -5
