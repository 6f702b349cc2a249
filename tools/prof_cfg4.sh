# Developer: ncu --set full of the generic-path (cfg-4) kernels, one instance each
CMD="python tools/cfg4_bench.py"
for spec in "cg_agg_multi_fwd:1" "cg_agg_multi_bwd_edge:1" "cg_agg_multi_bwd_node:1" "cg_pt_bwd:40" "cg_pt_fwd:40" "rad_bwd_kernel:1"; do
  k=${spec%%:*}; s=${spec##*:}
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:${k} -s $s -c 1 -f -o gpurun_out/cfg4_${k} $CMD > gpurun_out/cfg4_${k}.log 2>&1
  python tools/ncu_summary.py gpurun_out/cfg4_${k}.ncu-rep 2>&1 | grep -E "Kernel Name|Grid Size|Block Size|gpu__time_duration|occupancy_limit|registers_per_thread|shared_mem_per_block_dynamic|warps_active|dram__bytes|pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed|dram_throughput"
done
