#!/bin/bash
# K6 bring-up: every case in its own process under a timeout (a hung kernel must not take the box).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { timeout 60 python tools/gemm_check.py "$@" 2>&1 | tail -3; rc=${PIPESTATUS[0]}; [ $rc -ne 0 ] && echo "  -> rc=$rc for: $*"; }
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== 1-CTA K-major/K-major"
run 256 128 128 0 0 0 1 128 1
run 300 200 328 0 0 0 1 64 1
run 1028 1024 1296 0 0 0 1 256 1
run 1028 1024 1296 0 0 1 1 256 1
echo "== 1-CTA MN-major B / A"
run 512 256 256 0 1 0 1 128 1
run 1028 1024 128 0 1 0 1 256 1
run 512 256 512 1 0 1 1 128 1
run 1024 1296 4112 1 1 1 1 256 1
run 1024 1296 4112 1 1 1 3 256 1
echo "== 2-CTA"
run 512 256 256 0 0 0 1 128 2
run 1028 1024 1296 0 0 0 1 256 2
run 1000 520 1296 0 1 0 1 256 2
run 1024 1296 4112 1 1 1 3 256 2
run 1028 100 304 0 0 0 1 64 2
echo "== timing (B=128 shapes)"
run 131584 1024 1296 0 0 0 1 256 2 t
run 131584 1024 1296 0 0 0 1 256 1 t
run 131584 1024 1296 0 0 0 1 128 2 t
run 131584 3584 1296 0 0 0 1 256 2 t
run 131584 1024 128 0 1 0 1 256 2 t
run 131584 1296 3584 0 1 0 1 256 2 t
run 3584 1296 131584 1 1 1 2 256 2 t
run 131584 256 1024 0 0 0 1 256 2 t
} | tee gpurun_out/gemm_round.log
