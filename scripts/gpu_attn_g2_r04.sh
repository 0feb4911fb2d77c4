#!/bin/bash
# Global-attention variant G2 (column bias terms in shared memory, 24 KB stages, two CTAs per SM): full GPU suite with it
# as the default, parity of the old path (CSAM_ATTN_G2=0), alternating bench A/B, one ncu capture.
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests -q -m gpu --timeout 900 -x 2>&1 | tail -3
echo "=== G2=0 parity"; CSAM_ATTN_G2=0 timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -m gpu -k "vit_attention_relpos" 2>&1 | tail -1
CSAM_ATTN_G2=0 timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
for rep in 1 2; do for v in "CSAM_ATTN_G2=0" "CSAM_ATTN_G2=1"; do
  env $v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/exp_$v.json 2> gpurun_out/exp.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/exp_$v.json').read().strip().splitlines()[-1])
    print('$v', round(d['ms_per_step'], 2), 'ms', round(d['value'], 2), 'img/s  e2e', round(d['e2e']['value'], 2), 'clk', d['clocks']['sm_mhz'], 'attn', round(d['kernel_ms_per_step']['vit_attention'], 2))
except Exception as e:
    print('$v unparsed', e); print(open('gpurun_out/exp.err').read()[-800:])
PY
done; done
NB="--kernel-name-base demangled"
FULL="--set full --metrics lts__t_bytes.sum,lts__t_sectors_op_read.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__cycles_active.avg --clock-control none $NB -f"
timeout 600 ncu $FULL -k 'regex:vit_attention_ts_kernel<\(int\)3, \(int\)2,' -s 1 -c 1 -o gpurun_out/prof_attn_glob_g2_r04 python scripts/profile_step.py 1 > gpurun_out/ncu_g2.log 2>&1
python scripts/ncu_summary.py gpurun_out/prof_attn_glob_g2_r04.ncu-rep | cut -c1-330
ncu -i gpurun_out/prof_attn_glob_g2_r04.ncu-rep --page details > gpurun_out/prof_attn_glob_g2_r04.details.txt 2>/dev/null
rm -f gpurun_out/prof_attn_glob_g2_r04.ncu-rep
