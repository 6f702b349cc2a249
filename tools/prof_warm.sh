set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
timeout 600 python bench.py --impl reference --ref-device cuda --steps 3 --warmup 1 > gpurun_out/ref_cuda_noise.log 2>&1; tail -2 gpurun_out/ref_cuda_noise.log
for spec in "level_fwd:2" "level_bwd:3"; do
  k=${spec%%:*}; s=${spec##*:}
  timeout 600 ncu --set full --cache-control none --clock-control none --import-source on -k regex:${k}_kernel -s $s -c 1 -f -o gpurun_out/warm_${k} $CMD > gpurun_out/warm_${k}.log 2>&1
  python tools/ncu_summary.py gpurun_out/warm_${k}.ncu-rep 2>&1 | grep -E "Kernel Name|Block Size|gpu__time_duration|pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed|warps_active|registers_per_thread|dram__bytes"
done
