#!/bin/bash
# One GPU visit: tests, bench, ncu launch list of the bench's timed steps, full ncu capture of our kernels.
mkdir -p gpurun_out
TAG=${1:-r1}
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/pytest_gpu_$TAG.log
tail -2 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
head -c 600 gpurun_out/bench_$TAG.json; echo
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> /dev/null
head -c 300 gpurun_out/bench_ref_$TAG.json; echo
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_$TAG.log 2>&1
wc -l gpurun_out/launches_$TAG.csv
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:'knn|graph_conv|surface_conv|orl_|bn_|upsample|residual|gather_max|chamfer|kf_|dir_reduce|sqnorm' -c 160 -f -o gpurun_out/prof_$TAG python tools/ncu_step.py 128 bf16 > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log; ls -la gpurun_out/ | grep $TAG
