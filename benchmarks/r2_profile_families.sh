# Round-2 profile collection (one GPU): GPU tests, the family metrics pass, a full-set capture of k_pair, the launch list
# of the bench command.  Every ncu command runs only after the same program has exited 0 without ncu.
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests5.log
timeout 120 python benchmarks/profile_driver_r2.py > gpurun_out/r2_units.json 2> gpurun_out/r2_units.err || exit 1
MET=gpu__time_duration.sum,launch__registers_per_thread,launch__grid_size,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fmaheavy.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $MET --clock-control none --csv --log-file gpurun_out/r2_families_metrics.csv python benchmarks/profile_driver_r2.py > gpurun_out/r2_families_ncu.log 2>&1; echo "families ncu rc=$?"
python bench.py --log2-batch 18 --steps 2 --warmup 3 --light --rows none > gpurun_out/r2_bench_light18.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pair -c 1 -o gpurun_out/r2_k_pair_full -f python bench.py --log2-batch 18 --steps 2 --warmup 3 --light --rows none > gpurun_out/r2_k_pair_ncu.log 2>&1; echo "k_pair ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_log2_18.csv python bench.py --log2-batch 18 --steps 2 --warmup 3 --light --rows none > gpurun_out/r2_launches_ncu.log 2>&1; echo "launch list rc=$?"
