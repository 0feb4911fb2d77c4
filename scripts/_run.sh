for pf in 0 2 0 2; do echo "pf=$pf"; CSAM_T2I_PF=$pf timeout 100 python scripts/prof_t2i.py 1024 2>&1 | tail -1; done
