export PYTHONPATH=$PWD
NB="--kernel-name-base demangled"
FULL="--set full --metrics lts__t_bytes.sum,lts__t_sectors_op_read.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__cycles_active.avg --clock-control none $NB -f"
timeout 900 ncu $FULL --import-source on -k 'regex:dec_i2t_layer_kernel' -c 2 -o gpurun_out/prof_dec_i2t_r03 python scripts/profile_step.py 1 > gpurun_out/ncu_i2t.log 2>&1
timeout 900 ncu $FULL --import-source on -k 'regex:dec_t2i_kernel' -c 3 -o gpurun_out/prof_dec_t2i_r03 python scripts/profile_step.py 1 > gpurun_out/ncu_t2i.log 2>&1
python scripts/ncu_summary.py gpurun_out/prof_dec_*_r03.ncu-rep > gpurun_out/ncu_summary_r03_dec.csv
for f in gpurun_out/prof_dec_*_r03.ncu-rep; do ncu -i $f --page details > ${f%.ncu-rep}.details.txt 2>/dev/null; done
for t in dec_i2t dec_t2i; do ncu -i gpurun_out/prof_${t}_r03.ncu-rep --page source --csv > gpurun_out/src_${t}_r03.csv 2>/dev/null; done
rm -f gpurun_out/prof_*_r03.ncu-rep
cut -d, -f1-9 gpurun_out/ncu_summary_r03_dec.csv | cut -c1-260
