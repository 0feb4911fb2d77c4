timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "decoder_fused" 2>&1 | tail -3
for pf in 0 1; do echo "t2i pf=$pf"; CSAM_T2I_PF=$pf timeout 100 python scripts/prof_t2i.py 1024 2>&1 | tail -1; done
export CSAM_LIB_PATH=$PWD/crowdsam_b200/_C_trace/libcsam_sm100.so
timeout 100 python scripts/trace_dec.py t2i 296 > gpurun_out/trace_t2i.txt 2>&1
