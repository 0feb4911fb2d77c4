#!/bin/bash
# Evidence for the final state of round 2 (tagged r04): launch list of one bench-shaped step + ncu --set full captures of
# the named kernels ON THE BENCH INPUTS.  Same recipe as gpu_profiles_r03.sh; kernels are selected by template arguments
# because the two encoder streams interleave their launches.
mkdir -p gpurun_out
rm -f gpurun_out/prof_*_r04.ncu-rep
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? $name"; tail -n 2 gpurun_out/$name.log | cut -c1-200; }
export PYTHONPATH=$PWD
NB="--kernel-name-base demangled"
FULL="--set full --metrics lts__t_bytes.sum,lts__t_sectors_op_read.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__cycles_active.avg --clock-control none $NB -f"
run launches ncu --nvtx --nvtx-include "step/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r04.csv python scripts/profile_step.py 2
run ncu_gemm ncu $FULL -k 'regex:gemm_pair_kernel' -s 30 -c 8 -o gpurun_out/prof_gemm_enc_r04 python scripts/profile_step.py 1
run ncu_attn_dino ncu $FULL -k 'regex:vit_attention_ts_kernel<\(int\)3, \(int\)0, \(int\)2' -s 5 -c 2 -o gpurun_out/prof_attn_dino_r04 python scripts/profile_step.py 1
run ncu_attn_win ncu $FULL -k 'regex:vit_attention_ts_kernel<\(int\)3, \(int\)1,' -s 2 -c 1 -o gpurun_out/prof_attn_win_r04 python scripts/profile_step.py 1
run ncu_attn_glob ncu $FULL -k 'regex:vit_attention_ts_kernel<\(int\)3, \(int\)2,' -s 1 -c 1 -o gpurun_out/prof_attn_glob_r04 python scripts/profile_step.py 1
run ncu_i2t ncu $FULL -k 'regex:dec_i2t_layer_kernel' -c 2 -o gpurun_out/prof_dec_i2t_r04 python scripts/profile_step.py 1
run ncu_t2i ncu $FULL -k 'regex:dec_t2i_kernel' -c 3 -o gpurun_out/prof_dec_t2i_r04 python scripts/profile_step.py 1
run ncu_up ncu $FULL -k 'regex:gemm_tc_kernel<\(int\)(256|128), \(int\)3, \(bool\)0, \(int\)(2|3)' -c 2 -o gpurun_out/prof_gemm_up_r04 python scripts/profile_step.py 1
run ncu_postfull ncu $FULL -k 'regex:post_(stats|write)_quad' -s 4 -c 2 -o gpurun_out/prof_post_p1024_r04 python scripts/bench_post.py 1
python scripts/launch_summary.py gpurun_out/launches_r04.csv > gpurun_out/launches_r04_summary.csv
python scripts/ncu_summary.py gpurun_out/prof_*_r04.ncu-rep > gpurun_out/ncu_summary_r04.csv
for f in gpurun_out/prof_gemm_enc_r04.ncu-rep gpurun_out/prof_attn_dino_r04.ncu-rep gpurun_out/prof_attn_win_r04.ncu-rep gpurun_out/prof_attn_glob_r04.ncu-rep; do ncu -i $f --page details > ${f%.ncu-rep}.details.txt 2>/dev/null; done
ls -la gpurun_out/*.ncu-rep
rm -f gpurun_out/prof_*_r04.ncu-rep
head -24 gpurun_out/launches_r04_summary.csv
cat gpurun_out/ncu_summary_r04.csv | cut -c1-330
