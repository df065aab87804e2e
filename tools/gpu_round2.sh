#!/bin/bash
# One GPU visit (round 2): tests, bench lines, launch list of the bench's timed steps, full ncu capture of our kernels.
mkdir -p gpurun_out
TAG=${1:-r2}
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -v "UserWarning\|run_backward" | tail -6) > gpurun_out/pytest_gpu_$TAG.log
tail -2 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
head -c 300 gpurun_out/bench_$TAG.json; echo
timeout 300 python bench.py --steps 20 --warmup 5 --optimizer ranger --no-cpu-baseline > gpurun_out/bench_${TAG}_ranger.json 2>/dev/null
timeout 300 python bench.py --steps 20 --warmup 5 --precision fp32 --no-cpu-baseline > gpurun_out/bench_${TAG}_fp32.json 2>/dev/null
timeout 300 python bench.py --steps 20 --warmup 5 --batch 64 --no-cpu-baseline > gpurun_out/bench_${TAG}_b64.json 2>/dev/null
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference_arm.json 2> /dev/null
timeout 300 python bench.py --impl reference-gpu --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference_gpu_arm.json 2> /dev/null
for f in ranger fp32 b64 reference_arm reference_gpu_arm; do head -c 200 gpurun_out/bench_${TAG}_$f.json; echo; done
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_$TAG.log 2>&1
wc -l gpurun_out/launches_$TAG.csv
python tools/summarize_launches.py gpurun_out/launches_$TAG.csv 2 > gpurun_out/${TAG}_launches_summary.txt 2>/dev/null
gzip -f -k gpurun_out/launches_$TAG.csv
# NOTE: ~420 captures of the B = 128 step with --set full took 25 GPU-minutes in round 2: cap the count when the budget is short
timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:'gemm_tc|losses_|optim_|augment|split_bf16|knn|graph_conv|surface_conv|orl_|bn_|upsample|residual|gather_max|chamfer|kf_|dir_reduce|sqnorm|colmax|pair_dirs|absmax|normalize_cols' -c 420 -f -o gpurun_out/prof_$TAG python tools/ncu_step.py 128 bf16 > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
python tools/ncu_summary.py gpurun_out/prof_$TAG.ncu-rep gpurun_out/${TAG}_ncu_full_summary.md gpurun_out/${TAG}_kernel_traffic.json > /dev/null 2>&1
# the GEMM kernel's source-level stall reasons (needs -lineinfo): keep only a compact extract
ncu -i gpurun_out/prof_$TAG.ncu-rep --page details --csv -k regex:gemm_tc 2>/dev/null | head -400 > gpurun_out/${TAG}_ncu_gemm_details.csv
ls -la gpurun_out/prof_$TAG.ncu-rep; rm -f gpurun_out/prof_$TAG.ncu-rep
ls gpurun_out | grep $TAG
