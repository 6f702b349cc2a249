#!/bin/bash
# Developer: the round's ncu evidence on a GPU box (one GPU): launch list of the bench command and one `--set full` capture of the
# largest-channel instance of each hot kernel family.  usage: tools/ncu_round.sh <tag>   -> gpurun_out/<tag>_*
tag=${1:-r02}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches_bench.csv $CMD > gpurun_out/${tag}_ncu_launch.log 2>&1
python tools/summarize_launches.py gpurun_out/${tag}_launches_bench.csv gpurun_out/${tag}_launches_bench.txt | tail -40
# (kernel regex, launches of that family to skip so that the C = 4 / width-48 instance is captured)
for spec in "level_bwd:3" "mlp_bwd:1" "level_fwd:2" "mlp_fwd:1" "radial_bwd:0" "radial_fwd_multi:0" "reduce_segs:0"; do
  k=${spec%%:*}; s=${spec##*:}
  ncu --set full --clock-control none --import-source on -k regex:${k}_kernel -s $s -c 1 -f -o gpurun_out/${tag}_ncu_${k} $CMD > gpurun_out/${tag}_ncu_${k}.log 2>&1
  python tools/ncu_summary.py gpurun_out/${tag}_ncu_${k}.ncu-rep > gpurun_out/${tag}_ncu_${k}.txt 2>&1
  echo "== $k"; grep -E "Kernel Name|gpu__time_duration|pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed|tensor_src_fp64|warps_active|registers_per_thread|dram__bytes" gpurun_out/${tag}_ncu_${k}.txt
done
