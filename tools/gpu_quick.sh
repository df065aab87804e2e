#!/bin/bash
# Short GPU visit: selected tests, the GEMM shape sweep, one bench line and the launch list of two steps.
mkdir -p gpurun_out
TAG=${1:-q}
SEL=${2:-tests/test_kernels_gpu.py}
(timeout 900 python -m pytest $SEL -m gpu -q -x 2>&1 | grep -v "UserWarning\|run_backward" | tail -8)
DBG=0 CTAS=2 TILES=256 timeout 200 python tools/gemm_sweep.py wgrad 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
head -c 300 gpurun_out/bench_$TAG.json; echo
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
   --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_$TAG.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_$TAG.csv 2 > gpurun_out/${TAG}_launches_summary.txt 2>/dev/null
head -12 gpurun_out/${TAG}_launches_summary.txt
