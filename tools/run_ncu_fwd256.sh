ARGS="--workload config4 --steps 2 --warmup 3 --repeats 1 --advance 100 --no-cpu-baseline"
python bench.py $ARGS > gpurun_out/r02_ncu256_plain.json 2> gpurun_out/r02_ncu256_plain.err &&
ncu --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.sum,sm__inst_executed_pipe_tensor.sum,sm__cycles_elapsed.avg,sm__cycles_elapsed.avg.per_second,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none --profile-from-start off -k regex:"mlp256_fwd" -c 3 --csv --log-file gpurun_out/r02_ncu_fwd256_tensor_metrics.csv python bench.py $ARGS > gpurun_out/r02_ncu256_metrics.log 2>&1
tail -3 gpurun_out/r02_ncu256_metrics.log
grep -c mlp256 gpurun_out/r02_ncu_fwd256_tensor_metrics.csv
