#!/bin/bash
# round 2, call C (1 GPU): smoke, ncu launch list of the bench command, ncu --set full for C1 / C3 / C2 / C4 / C2-precise
# (summarised ON the box with scripts/ncu_summary.py: the raw reports together exceed what gpurun copies back),
# reference-kernel table (scripts/ref_table.py, identical inputs for both kernels)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
H=flashattention.c_b200/harness
mkdir -p gpurun_out /tmp/prof
L=gpurun_out/r2c.log
S=gpurun_out/r2_ncu_summary.md
echo "== smoke" > $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1
echo "== ncu launch list of the bench command" >> $L
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_launches.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > gpurun_out/r2c_bench_under_ncu.json 2>> $L
NCU="ncu --set full --clock-control none --import-source on -k regex:fa_fwd_sm100 -s 2 -c 1 -f"
echo "# Round 2 — ncu --set full summaries (B200, --clock-control none), kernel v10" > $S
prof() {  # name, title, env, fa_check args...
  name=$1; title=$2; envs=$3; shift 3
  env $envs timeout 300 $NCU -o /tmp/prof/$name $H/fa_check "$@" >> $L 2>&1
  python scripts/ncu_summary.py /tmp/prof/$name.ncu-rep "$title" >> $S 2>> $L
}
prof c1 "C1: fp32->tf32 d=64 B*H=16 N=1024 non-causal (128 split-KV items on 148 SMs)" A=1 f32 64 16 1024 0 0 2 0
prof c3 "C3: fp32->tf32 d=32 B*H=128 N=1024 non-causal" A=1 f32 32 128 1024 0 0 2 0
prof c2 "C2: fp32->tf32 d=64 B*H=16 N=8192 non-causal (headline)" A=1 f32 64 16 8192 0 0 2 0
prof c4 "C4: bf16 d=128 B*H=128 N=8192 non-causal" A=1 bf16 128 128 8192 0 0 2 0
prof c2p "C2 with FA_FLAG_PRECISE (3xTF32, one-slot instance, slot B's warps write the lo copies)" FA_B200_PRECISE=1 f32 64 16 8192 0 0 2 0
cp /tmp/prof/c3.ncu-rep gpurun_out/r2_prof_c3.ncu-rep
echo "== reference kernel table" >> $L
timeout 900 python scripts/ref_table.py > gpurun_out/r2c_ref_table.jsonl 2>> $L
cat gpurun_out/r2c_ref_table.jsonl >> $L
grep -v "^==PROF==\|^==WARNING==\|^$" $L | cut -c1-250 | tail -n 14
wc -l $S
