#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
H=flashattention.c_b200/harness
mkdir -p gpurun_out
L=gpurun_out/dbg.log
: > $L
run() { echo "== $*" >> $L; timeout 120 $H/fa_check "$@" >> $L 2>&1 || echo "  (exit $?)" >> $L; }
run f32 32 128 1024 1 0 30 0
run f32 32 64 1024 1 0 5 0
run f32 32 40 512 1 0 5 0
run f32 32 80 256 1 0 5 0
run bf16 64 128 1024 1 0 5 0
run f32 32 128 1024 0 0 30 0
run f32 64 16 8192 0 0 20 0
run bf16 128 128 8192 0 0 10 0
run f32 64 16 1024 0 0 30 0
cut -c1-60,150-400 $L
for i in 1 2 3; do timeout 120 $H/fa_check f32 32 128 1024 1 0 50 0 | cut -c1-60,150-400; timeout 120 $H/fa_check bf16 64 128 1024 1 0 50 0 | cut -c1-60,150-400; done
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
