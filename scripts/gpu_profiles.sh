#!/bin/bash
# Round-end evidence: launch list of one bench-shaped step + ncu --set full captures of the named kernels.
mkdir -p gpurun_out
rm -f gpurun_out/prof_*_r01.ncu-rep
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? $name"; tail -n 2 gpurun_out/$name.log | cut -c1-200; }
export PYTHONPATH=$PWD
NB="--kernel-name-base demangled"
FULL="--set full --clock-control none $NB -f"
run launches ncu --nvtx --nvtx-include "step/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv python scripts/profile_step.py 2
run ncu_gemm ncu $FULL --import-source on -k 'regex:gemm_tc_kernel<\(int\)128, \(int\)3, \(bool\)0, \(int\)0, \(bool\)0>' -s 30 -c 4 -o gpurun_out/prof_gemm_enc_r01 python scripts/profile_step.py 1
run ncu_attn_win ncu $FULL -k 'regex:vit_attention_t' -s 0 -c 1 -o gpurun_out/prof_attn_win_r01 python scripts/profile_step.py 1
run ncu_attn_glob ncu $FULL -k 'regex:vit_attention_t' -s 5 -c 1 -o gpurun_out/prof_attn_glob_r01 python scripts/profile_step.py 1
run ncu_attn_dino ncu $FULL --import-source on -k 'regex:vit_attention_t' -s 30 -c 2 -o gpurun_out/prof_attn_dino_r01 python scripts/profile_step.py 1
run ncu_i2t ncu $FULL --import-source on -k 'regex:dec_i2t_layer_kernel' -c 2 -o gpurun_out/prof_dec_i2t_r01 python scripts/profile_step.py 1
run ncu_t2i ncu $FULL --import-source on -k 'regex:dec_t2i_kernel' -c 3 -o gpurun_out/prof_dec_t2i_r01 python scripts/profile_step.py 1
run ncu_up ncu $FULL -k 'regex:gemm_tc_kernel<\(int\)(256|128), \(int\)3, \(bool\)0, \(int\)(2|3)' -c 2 -o gpurun_out/prof_gemm_up_r01 python scripts/profile_step.py 1
run ncu_post ncu $FULL -k 'regex:post_' -c 4 -o gpurun_out/prof_post_r01 python scripts/profile_step.py 1
run ncu_postfull ncu $FULL -k 'regex:post_write_quad' -c 2 -o gpurun_out/prof_post_p1024_r01 python scripts/bench_post.py
# keep the box -> container transfer small: summaries + per-kernel detail pages instead of the raw reports
python scripts/ncu_summary.py gpurun_out/prof_*_r01.ncu-rep > gpurun_out/ncu_summary_r01.csv
for f in gpurun_out/prof_*_r01.ncu-rep; do ncu -i $f --page details > ${f%.ncu-rep}.details.txt 2>/dev/null; done
ls -la gpurun_out/*.ncu-rep
rm -f gpurun_out/prof_*_r01.ncu-rep
du -sh gpurun_out
