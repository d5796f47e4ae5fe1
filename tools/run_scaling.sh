#!/bin/bash
# tools/run_scaling.sh N  -- the multi-GPU evidence of one box: data-parallel equivalence (both gradient-exchange paths) and the
# bench lines of configs 3, 4 and 5 at N GPUs.  Outputs land in gpurun_out/r02_*_n$N.*
# tools/run_scaling.sh N train   -- only the training lines (configs 3 and 4), e.g. after a kernel change that cannot affect the rest
N=$1
ONLY=${2:-all}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for p in 1 0; do
  [ "$ONLY" = "train" ] && continue
  ANGIO_P2P=$p timeout 300 $TR --master-port 2951$p tools/check_dp_equivalence.py > gpurun_out/r02_dp_equivalence_n${N}_p2p$p.log 2>&1
  echo "rc=$?" >> gpurun_out/r02_dp_equivalence_n${N}_p2p$p.log
done
timeout 600 $TR --master-port 29533 bench.py --gpus $N --no-cpu-baseline > gpurun_out/r02_bench_config3_n$N.json 2> gpurun_out/r02_bench_config3_n$N.err
echo "rc=$?" >> gpurun_out/r02_bench_config3_n$N.err
timeout 900 $TR --master-port 29534 bench.py --gpus $N --workload config4 --advance 100 --repeats 3 --steps 10 --no-cpu-baseline > gpurun_out/r02_bench_config4_n$N.json 2> gpurun_out/r02_bench_config4_n$N.err
echo "rc=$?" >> gpurun_out/r02_bench_config4_n$N.err
if [ "$ONLY" != "train" ]; then
timeout 600 $TR --master-port 29535 bench.py --gpus $N --workload config5 --steps 6 --repeats 3 > gpurun_out/r02_bench_config5_n$N.json 2> gpurun_out/r02_bench_config5_n$N.err
echo "rc=$?" >> gpurun_out/r02_bench_config5_n$N.err
fi
[ "$ONLY" = "train" ] && { tail -n 2 gpurun_out/r02_bench_config3_n$N.err gpurun_out/r02_bench_config4_n$N.err; exit 0; }
tail -n 2 gpurun_out/r02_dp_equivalence_n${N}_p2p1.log gpurun_out/r02_dp_equivalence_n${N}_p2p0.log gpurun_out/r02_bench_config3_n$N.err gpurun_out/r02_bench_config4_n$N.err gpurun_out/r02_bench_config5_n$N.err
