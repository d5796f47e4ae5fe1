#!/bin/bash
# tools/run_profiles_n1.sh -- 1-GPU evidence: bench lines of configs 3/4/5, A/B of the forward kernels, pipeline trace, and the ncu
# launch list + full capture of one steady-regime window of config 3 (bench.py brackets the window with cudaProfilerStart/Stop).
timeout 600 python bench.py > gpurun_out/r02_bench_config3_n1.json 2> gpurun_out/r02_bench_config3_n1.err; echo "rc=$?" >> gpurun_out/r02_bench_config3_n1.err
timeout 900 python bench.py --workload config4 --advance 100 --repeats 3 --steps 10 > gpurun_out/r02_bench_config4_n1.json 2> gpurun_out/r02_bench_config4_n1.err; echo "rc=$?" >> gpurun_out/r02_bench_config4_n1.err
timeout 600 python bench.py --workload config5 --steps 6 --repeats 3 > gpurun_out/r02_bench_config5_n1.json 2> gpurun_out/r02_bench_config5_n1.err; echo "rc=$?" >> gpurun_out/r02_bench_config5_n1.err
timeout 300 python tools/time_fwd.py > gpurun_out/r02_fwd_two_vs_three_slots.log 2>&1
timeout 200 python tools/time_fwd256.py > gpurun_out/r02_fwd256_isolated.log 2>&1
timeout 120 python tools/trace_fwd.py > gpurun_out/r02_fwd3_pipeline_trace.txt 2>&1
ARGS="--steps 2 --warmup 3 --repeats 1 --advance 300 --no-cpu-baseline"
python bench.py $ARGS > gpurun_out/r02_ncu_plain.json 2> gpurun_out/r02_ncu_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_ncu_launch_list.csv python bench.py $ARGS > gpurun_out/r02_ncu_list.log 2>&1
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"mlp_fwd3|mlp_wgrad|mlp_dgrad|mlp_fwd_tc|visibility_mask|march_count|march_write|march_head|composite_mse|compact_head|sample_candidates" -c 14 -o gpurun_out/r02_full python bench.py $ARGS > gpurun_out/r02_ncu_full.log 2>&1
ncu --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.sum,sm__inst_executed_pipe_tensor.sum,sm__cycles_elapsed.avg,sm__cycles_elapsed.avg.per_second,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -k regex:"mlp_fwd3" -c 2 --csv --log-file gpurun_out/r02_ncu_fwd3_tensor_metrics.csv python bench.py $ARGS > gpurun_out/r02_ncu_metrics.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r02_smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests_n1.log 2>&1; echo "rc=$?" >> gpurun_out/r02_tests_n1.log
timeout 600 python tools/psnr_check.py --workload config2 --rays 5625 --iters 3000 > gpurun_out/r02_psnr_config2_bf16_vs_fp32.log 2>&1
tail -n 3 gpurun_out/r02_smoke.log gpurun_out/r02_tests_n1.log gpurun_out/r02_psnr_config2_bf16_vs_fp32.log
tail -n 2 gpurun_out/r02_bench_config3_n1.err gpurun_out/r02_bench_config4_n1.err gpurun_out/r02_bench_config5_n1.err gpurun_out/r02_fwd_two_vs_three_slots.log
ls -la gpurun_out/r02_full.ncu-rep gpurun_out/r02_ncu_launch_list.csv
