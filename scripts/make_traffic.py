"""profiles/ncu_summary_rNN.csv -> profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch of
the kernel behind every roofline entry of bench.py (mean over the captured launches of that class).
usage: python scripts/make_traffic.py profiles/ncu_summary_r01.csv > profiles/traffic.json"""
import csv, json, sys

CLASSES = {   # bench.py roofline name -> (report file prefix, kernel-name substring)
    # encoder / DINOv2 GEMMs: cta_group::2 kernel with 256x128 pair tiles since r04 (single-CTA kernel before)
    "gemm_tensor": ("prof_gemm_enc", "gemm_"),
    "gemm_pair": ("prof_gemm_pair", "gemm_pair_kernel"),
    "mask_post_p1024": ("prof_post_p1024", "post_"),
    "gemm_hbm": ("prof_gemm_up", "gemm_tc_kernel"),
    "vit_attention": ("prof_attn_dino", "vit_attention_t"),
    "dec_i2t_layer": ("prof_dec_i2t", "dec_i2t_layer_kernel"),
    "dec_t2i": ("prof_dec_t2i", "dec_t2i_kernel"),
    # (the in-step K-POST launches depend on how many prompts survive on the captured image -- 336 on image 0 against
    #  ~900 on the bench images -- so only the P = 1024 all-survive pair, run on bench.post_fixture, is kept as evidence)
}
rows = list(csv.DictReader(open(sys.argv[1])))
rd = next(k for k in rows[0] if k.startswith("dram__bytes_read.sum"))
wr = next(k for k in rows[0] if k.startswith("dram__bytes_write.sum"))
out = {"_note": "bytes per launch (dram read + write), mean over the ncu --set full captures listed in " + sys.argv[1]}
for name, (rep, kern) in CLASSES.items():
    v = [float(r[rd]) + float(r[wr]) for r in rows if r["report"].startswith(rep) and kern in r["Kernel Name"]]
    if name in ("dec_i2t_layer", "dec_t2i") and len(v) > 1:
        # the first launch of a step works on the keys all prompts share (layer 0): its own roofline entry
        out[name + "_shared"] = v[0]
        v = v[1:]
    if v:
        # the combined K-POST entry is a PAIR of launches (stats + write): sum, not mean
        out[name] = sum(v) if name == "mask_post_p1024" else sum(v) / len(v)
print(json.dumps(out, indent=1))
