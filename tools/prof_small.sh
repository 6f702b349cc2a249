CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
for spec in "latent_bridge_bwd:0"; do
  k=${spec%%:*}; s=${spec##*:}
  timeout 300 ncu --set full --cache-control none --clock-control none --import-source on -k regex:${k} -s $s -c 1 -f -o gpurun_out/small2_${k} $CMD > gpurun_out/small2_${k}.log 2>&1
  python tools/ncu_summary.py gpurun_out/small2_${k}.ncu-rep 2>&1 | grep -E "Kernel Name|gpu__time_duration|cycles_elapsed.max|smsp__cycles_active.avg|inst_executed"
done
