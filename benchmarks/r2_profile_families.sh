set -x
( time python benchmarks/profile_driver_r2.py > gpurun_out/r2_units.json 2> gpurun_out/r2_units.err ) 2> gpurun_out/r2_units.time
MET=gpu__time_duration.sum,launch__registers_per_thread,launch__grid_size,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fmaheavy.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum
( time timeout 900 ncu --metrics $MET --clock-control none -k regex:'k_pair|k_check2|k_scalar_mul|k_fixed_mul|k_gt_|k_miller_lines|k_final_exp|k_multi_pair|k_subset|k_msm|k_hash|k_wvm|k_vm' --csv --log-file gpurun_out/r2_families_metrics.csv python benchmarks/profile_driver_r2.py > gpurun_out/r2_families_ncu.log 2>&1 ) 2> gpurun_out/r2_families.time
tail -3 gpurun_out/r2_units.time gpurun_out/r2_families.time
