"""profiles/ncu_summary_rNN.csv -> profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch of
the kernel behind every roofline entry of bench.py (mean over the captured launches of that class).
usage: python scripts/make_traffic.py profiles/ncu_summary_r01.csv > profiles/traffic.json"""
import csv, json, sys

CLASSES = {   # bench.py roofline name -> (report file prefix, kernel-name substring)
    "gemm_tensor": ("prof_gemm_enc", "gemm_tc_kernel<128, 3, 0, 0, 0>"),
    "gemm_hbm": ("prof_gemm_up", "gemm_tc_kernel"),
    "vit_attention": ("prof_attn_dino", "vit_attention_t"),
    "dec_i2t_layer": ("prof_dec_i2t", "dec_i2t_layer_kernel"),
    "dec_t2i": ("prof_dec_t2i", "dec_t2i_kernel"),
    "mask_post_write": ("prof_post_r", "post_write_quad_kernel"),
    "mask_post_stats": ("prof_post_r", "post_stats_quad_kernel"),
    "mask_post_write_p1024": ("prof_post_p1024", "post_write_quad_kernel"),
}
rows = list(csv.DictReader(open(sys.argv[1])))
rd = next(k for k in rows[0] if k.startswith("dram__bytes_read.sum"))
wr = next(k for k in rows[0] if k.startswith("dram__bytes_write.sum"))
out = {"_note": "bytes per launch (dram read + write), mean over the ncu --set full captures listed in " + sys.argv[1]}
for name, (rep, kern) in CLASSES.items():
    v = [float(r[rd]) + float(r[wr]) for r in rows if r["report"].startswith(rep) and kern in r["Kernel Name"]]
    if v:
        out[name] = sum(v) / len(v)
print(json.dumps(out, indent=1))
