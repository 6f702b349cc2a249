for k in lin_bwd_w_mma_kernel lin_mma_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o gpurun_out/prof_$k python tools/linear_bench.py > gpurun_out/prof_$k.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_$k.ncu-rep 2>&1 | grep -E "Kernel Name|Grid Size|gpu__time_duration|tensor_src_fp64|warps_active|registers_per_thread|dram__bytes|issue_active"
done
