#!/usr/bin/env python
"""bench.py — images/sec of the Crowd-SAM hot path on B200 (BASELINE.json metric).

Workload (BASELINE.json configs[1]): SAM ViT-L random-init ("recipe v1" synthetic weights),
DINOv2 ViT-L/14, one synthetic 1024x1024 image per step, 32x32 point grid = 1024 prompts through the
batched mask decoder, PWD-Net scoring, stability / IoU filters, K-POST masks, box NMS.
A step = one image.  Reported on one JSON line:
  value   device-resident throughput (uint8 image already in HBM, no mask RLE / D2H)
  e2e     the same metric through the public API `CrowdSAM.generate(np.ndarray)`: pinned-host image ->
          H2D every step, full result dict (boxes, scores, COCO RLEs) read back to the host
  roofline      dominant kernel class, CUDA-event timed inside the timed region
  cpu_baseline  ONE full, un-extrapolated image of the same workload through the reference's CPU path
`--impl reference` times the reference's own CPU path on full images: the real /root/reference code when that tree
(or baseline/_ref) is present, else the oracle port (oracle/restate.py); as many whole images as fit the time budget,
never an extrapolated sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec (1024px, 32x32 prompt grid)"
UNIT = "images/s"
WORKLOAD = "vit_l_1024px_grid32_pwd_nms"
GRID = 32
ARCH = "vit_l"
DINO = "dinov2_vitl14"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons, pw = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
                try:
                    pw.append(float(r[2]))
                except ValueError:
                    pass                                   # power.draw can read "[N/A]"; the clocks / reasons still count
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "power_w": statistics.median(pw) if pw else None}


def test_cfg(points_per_batch: int):
    """Reference config overrides of SURVEY.md §8d: every grid cell becomes a prompt, EPS never prunes,
    the OpenCV small-region pass (not on the north_star path) is off; filters keep their YAML defaults."""
    from crowdsam_b200.synthetic import DEFAULT_TEST_CFG

    c = dict(DEFAULT_TEST_CFG)
    c.update(grid_size=GRID, pos_sim_thresh=-1, max_prompts=GRID * GRID, points_per_batch=points_per_batch,
             filter_thresh=1e9, min_mask_region_area=0, apply_box_offsets=False, fuse_simmap=False, output_rles=True)
    return c


# ------------------------------------------------------------------------------------------------
# The reference's CPU path on whole images (cpu_baseline leg and --impl reference)
# ------------------------------------------------------------------------------------------------
CPU_POINTS_PER_BATCH = 64      # 1024 prompts at once would materialise 2 x 17 GB of fp32 masks on the host


def make_cpu_model():
    """-> (generate(image) -> result, kind, description).  Probes baseline/_ref, then $CROWDSAM_REFERENCE or
    /root/reference, for the real reference (crowdsam/model.py on CPU with the three import shims of SURVEY §8c);
    falls back to the oracle port, which is pinned to the real reference by tests/golden."""
    import torch
    from crowdsam_b200 import synthetic as weights

    torch.set_num_threads(os.cpu_count() or 1)
    cfg = test_cfg(CPU_POINTS_PER_BATCH)
    sam_sd, dino_sd = weights.make_sam_state(ARCH), weights.make_dino_state(DINO)
    for cand in (os.path.join(ROOT, "baseline", "_ref"), os.environ.get("CROWDSAM_REFERENCE", "/root/reference")):
        if os.path.isdir(os.path.join(cand, "segment_anything_cs")) and os.path.isdir(os.path.join(cand, "crowdsam")):
            os.environ["CROWDSAM_REFERENCE"] = cand
            from oracle import ref_import

            ref_import.REF_ROOT = cand
            sam, dino = ref_import.build_sam(sam_sd, ARCH), ref_import.build_dino(dino_sd, DINO)
            m = ref_import.build_crowdsam(sam, dino, cfg)
            return (lambda img: m.generate(img)), "reference", f"real reference at {cand} (crowdsam.model.CrowdSAM.generate, CPU fp32)"
    from oracle import restate

    _, depth, heads, glob = weights.SAM_ARCHS[ARCH]
    _, ddepth, dheads = weights.DINO_ARCHS[DINO]
    m = restate.OracleCrowdSAM(sam_sd, dino_sd, (depth, heads, glob), (ddepth, dheads), cfg)
    return (lambda img: m.generate(img)), "port", ("oracle port of crowdsam.model.CrowdSAM.generate (oracle/restate.py, CPU fp32; "
                                                   "no reference tree at baseline/_ref or /root/reference on this host)")


def time_cpu_images(gen, first_index: int, max_images: int, budget_s: float):
    """Whole images through the CPU path until `max_images` or the budget is reached (at least one)."""
    import torch
    from crowdsam_b200 import synthetic as weights

    secs = []
    with torch.no_grad():
        while len(secs) < max_images:
            img = weights.synthetic_image(first_index + len(secs))
            np.random.seed(42)
            t0 = time.perf_counter()
            res = gen(img)
            secs.append(time.perf_counter() - t0)
            n_det = len(res["boxes"])
            if sum(secs) + max(secs) > budget_s:
                break
    return secs, n_det


def run_reference(args, rank, world):
    if rank != 0:
        return
    import torch

    gen, kind, what = make_cpu_model()
    budget = float(os.environ.get("CSAM_REF_BUDGET_S", "200"))
    secs, n_det = time_cpu_images(gen, args.warmup, max(1, args.steps), budget)
    per_img = sum(secs) / len(secs)
    value = 1.0 / per_img
    sample = (f"{what}; {len(secs)} whole image(s) timed end to end, un-extrapolated ({', '.join(f'{x:.1f}' for x in secs)} s; "
              f"requested {args.steps}, time budget {budget:.0f} s), {CPU_POINTS_PER_BATCH} prompts per decoder batch, no warm-up image")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(secs),
            "steps_requested": args.steps, "warmup": 0, "ms_per_step": per_img * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(ARCH, max(1, args.gpus)),
            "run": {"points_per_batch": CPU_POINTS_PER_BATCH, "masks_into_nms": None, "detections": n_det},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def bench_config(arch, world):
    """The workload, identical on both arms (the driver compares `config`); how each arm batched the prompts and
    what it found goes under `run`."""
    return {"workload": WORKLOAD, "arch": arch, "dino": DINO, "grid": GRID, "prompts": GRID * GRID,
            "filters": "pred_iou 0.1, stability 0.8 @ offset 1, box NMS 0.65 (YAML defaults), EPS pruning off, small-region pass off",
            "l2": "working set (weights 2.5 GB + per-batch activations > 10 GB) far exceeds the 126 MB L2",
            "parallelism": f"images sharded one per rank x{world}, all-gather of detections"}


def post_fixture(P: int, dev):
    """Instance-like low-res logits [P,4,256,256] on the device (plateau +6 inside an ellipse, -6 outside, edge width
    0.4-3 px, N(0,0.5) pixel noise: the shape of oracle/fixtures.py injected_decoder_outputs) + a selected plane."""
    import torch

    g = torch.Generator(device=dev).manual_seed(0)
    yy, xx = torch.meshgrid(torch.arange(256, device=dev, dtype=torch.float32),
                            torch.arange(256, device=dev, dtype=torch.float32), indexing="ij")
    cx, cy = (torch.rand(P, 1, 1, 1, device=dev, generator=g) * 255 for _ in range(2))
    r = 4 + 18 * torch.rand(P, 1, 1, 1, device=dev, generator=g)
    tau = 0.4 + 2.6 * torch.rand(P, 1, 1, 1, device=dev, generator=g)
    asp = 0.6 + 1.2 * torch.rand(P, 1, 1, 1, device=dev, generator=g)
    grow = torch.tensor([0.6, 0.9, 1.2, 1.5], device=dev).view(1, 4, 1, 1)
    low = torch.empty((P, 4, 256, 256), device=dev)
    for s0 in range(0, P, 128):
        sl = slice(s0, min(s0 + 128, P))
        d = torch.sqrt((xx - cx[sl]) ** 2 + ((yy - cy[sl]) / asp[sl]) ** 2)
        low[sl] = 6.0 * torch.tanh((r[sl] * grow - d) / tau[sl])
    low += 0.5 * torch.randn(low.shape, device=dev, generator=g)
    sel = torch.randint(0, 4, (P,), device=dev, generator=g).to(torch.int32)
    return low.contiguous(), sel


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--points-per-batch", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--arch", default=ARCH, help="vit_l (headline) | vit_h (BASELINE configs[3]) | vit_b")
    ap.add_argument("--grid", type=int, default=GRID, help="prompt grid side: 32 (headline), 64 = BASELINE configs[2]")
    ap.add_argument("--gemm-shapes", action="store_true", help="stderr: per-shape GEMM time table")
    args = ap.parse_args()
    if args.grid != GRID or args.arch != ARCH:
        # a non-headline workload: name it, and let test_cfg / bench_config follow
        globals()["GRID"] = args.grid
        globals()["WORKLOAD"] = f"{args.arch}_1024px_grid{args.grid}_pwd_nms"
        if args.points_per_batch == 1024:
            args.points_per_batch = args.grid * args.grid
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist

    from crowdsam_b200 import graphs, lib, ops, parallel
    from crowdsam_b200.build import _build_sam
    from crowdsam_b200.modules import DinoVisionTransformer
    from crowdsam_b200.pipeline import CrowdSAM
    from crowdsam_b200.predictor import SamPredictor
    from crowdsam_b200 import synthetic as weights     # synthetic weight recipe + images (pure data; no oracle import)

    if args.warmup < 3:
        print(f"[bench] --warmup {args.warmup} raised to 3 (timing rules: at least 3 warm-up steps)", file=sys.stderr)
        args.warmup = 3
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib.load()
    arch = args.arch
    D, depth, heads, glob = weights.SAM_ARCHS[arch]
    sam_sd, dino_sd = weights.make_sam_state(arch), weights.make_dino_state(DINO)
    sam = _build_sam(D, depth, heads, 1, glob)
    sam.load_state_dict(sam_sd, strict=True)
    dD, ddepth, dheads = weights.DINO_ARCHS[DINO]
    dino = DinoVisionTransformer(dD, ddepth, dheads)
    dino.load_state_dict(dino_sd, strict=True)
    pred = SamPredictor(sam.to(dev), dino.to(dev))
    cfg = {"environ": {"device": str(dev)}, "model": {"trainfree": False}, "test": test_cfg(args.points_per_batch)}
    model = CrowdSAM(cfg, None, predictor=pred)

    n_total = args.warmup + args.steps
    # per-rank images: rank r owns global images r, r+world, ... (independent units)
    imgs_np = [weights.synthetic_image(rank + world * i) for i in range(n_total)]
    pinned = [torch.as_tensor(x).pin_memory() for x in imgs_np]
    resident = [torch.as_tensor(x).permute(2, 0, 1).contiguous().to(dev) for x in imgs_np]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather_dets(dets):
        """The one exchange step: all-gather of padded [K, Nmax, 6] detections + counts."""
        if world > 1:
            parallel.gather_detections(dets, device=dev)

    # ---------------- device-resident leg (value) ----------------
    # One-time setup outside the W warm-up steps: the CUDA graphs of set_image / decode are captured on the second /
    # third call with a given shape, and torch.cuda.graph() empties the allocator cache on entry, so the first step
    # after a capture pays cudaMalloc for ~2 GB of result buffers.  Prime until no further capture happens.
    def prime(fn):
        n = 0
        for _ in range(6):
            c0 = graphs.captures
            np.random.seed(42)
            fn()
            n += 1
            if graphs.captures == c0 and n >= 2:
                break
        return n

    priming = prime(lambda: model.run_resident(resident[0]))
    for i in range(args.warmup):
        np.random.seed(42)
        model.run_resident(resident[i])
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = lib.launch_count() + graphs.replayed_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dets, nk = [], (0, 0)
    ev0.record()
    for i in range(args.steps):
        np.random.seed(42)
        d = model.run_resident(resident[args.warmup + i])
        nk = model.last_counts
        dets.append(None if d is None else {"boxes": d["boxes"], "scores": d["scores"], "categories": d["categories"]})
    gather_dets(dets)
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    print(f"[bench rank {rank}] resident leg: {ms_total / args.steps:.1f} ms/step, masks into NMS {nk[0]}, kept {nk[1]}",
          file=sys.stderr)
    launches = lib.launch_count() + graphs.replayed_launches - l0      # eager launches + kernels inside graph replays
    clocks = sampler.stop()
    # Per-kernel-class CUDA-event timing (the roofline numbers) runs as a SECOND pass over the same K images:
    # two event records around each of ~460 launches cost ~5 ms per step, which must not sit inside `value`.
    # Same kernels, same inputs, same stream, launched one by one (no graph replay: events cannot be timed inside a
    # graph); each class's share is taken against this pass's own step time.
    ops.PROFILER = ops.Profiler(detail=args.gemm_shapes)
    pv0, pv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pv0.record()
    for i in range(args.steps):
        np.random.seed(42)
        model.run_resident(resident[args.warmup + i])
    pv1.record()
    torch.cuda.synchronize()
    prof_ms_total = pv0.elapsed_time(pv1)
    prof = ops.PROFILER.summary()
    ops.PROFILER = None
    if args.gemm_shapes and rank == 0:
        rows = sorted(((k, v) for k, v in prof.items() if k.startswith("gemm ")), key=lambda kv: -kv[1]["total_ms"])
        for k, v in rows:
            tf = v["work"] / (v["total_ms"] * 1e-3) / 1e12
            print(f"[gemm] {k:28s} n/step={v['launches'] / args.steps:6.1f} ms/step={v['total_ms'] / args.steps:8.3f} "
                  f"avg_us={1e3 * v['total_ms'] / v['launches']:8.1f} TFLOP/s={tf:7.1f}", file=sys.stderr)
        prof = {k: v for k, v in prof.items() if not k.startswith("gemm ")}
    t = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * args.steps / (ms_total / 1e3)

    # ---------------- K-POST at full size: the "all-survive" run of SURVEY §8d (N = P = 1024 masks) ----------
    # stats pass + write pass TOGETHER on instance-like logits (plateau blobs with ragged edges, as the injected
    # parity fixture): per prompt the two passes read the selected 256x256 fp32 plane once each and write one
    # 1024x1024 bool mask = 1,572,864 B (SURVEY §8d counts all four planes once: 2,097,216 B; also reported).
    post_roof = None
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath))
    if rank == 0:
        low, sel = post_fixture(1024, dev)
        for _ in range(3):
            ops.mask_post_stats(low, sel, (1024, 1024), (1024, 1024), 0.0, 1.0)
            ops.mask_post_write(low, sel, None, (1024, 1024), (1024, 1024), 0.0)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        reps, ms_s, ms_w = 5, 0.0, 0.0
        for _ in range(reps):
            evs[0].record()
            ops.mask_post_stats(low, sel, (1024, 1024), (1024, 1024), 0.0, 1.0)
            evs[1].record()
            mk, _ = ops.mask_post_write(low, sel, None, (1024, 1024), (1024, 1024), 0.0)
            evs[2].record()
            torch.cuda.synchronize()
            ms_s += evs[0].elapsed_time(evs[1]) / reps
            ms_w += evs[1].elapsed_time(evs[2]) / reps
        nbytes = 1024 * (2 * 262144.0 + 1048576.0)
        pk = load_peaks()
        ach = nbytes / ((ms_s + ms_w) * 1e-3) / 1e9
        post_roof = {"kernel": "mask_post stats+write @P=1024 (all-survive, instance-like logits, 1.6 GB / pass pair > L2)",
                     "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                     "traffic": traffic.get("mask_post_p1024"), "avg_launch_ms": ms_s + ms_w, "stats_ms": ms_s,
                     "write_ms": ms_w, "bytes_per_prompt": 1572864, "frac_survey_bytes": ach / pk["hbm_gbs"] * 2097216 / 1572864,
                     "write_only_frac": 1024 * (262144.0 + 1048576.0) / (ms_w * 1e-3) / 1e9 / pk["hbm_gbs"],
                     "peak_source": pk["source"]}
        del low, sel, mk
        torch.cuda.empty_cache()

    # ---------------- end-to-end leg through the public API ----------------
    priming += prime(lambda: model.generate(imgs_np[0]))
    for i in range(min(args.warmup, 3)):
        np.random.seed(42)
        model.generate(imgs_np[i])
    barrier()
    d2h = 0
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        np.random.seed(42)
        res = model.generate(pinned[args.warmup + i].numpy())
        d2h = sum(np.asarray(v).nbytes for k, v in res.items() if isinstance(v, np.ndarray)) + \
            sum(len(r["counts"]) for r in res["rles"])
    e1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    # host work after the last kernel (RLE string of the last image) is part of the call a user makes: the step ends
    # when generate() returns, so the larger of the device-side and the host-side span is the end-to-end time
    e2e_ms = max(e0.elapsed_time(e1), wall_ms)
    t = torch.tensor([e2e_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * args.steps / (float(t.item()) / 1e3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()

    def roof(name, bound):
        r = prof.get(name)
        if not r or r["total_ms"] <= 0:
            return None
        per_launch_ms = r["total_ms"] / r["launches"]
        if bound == "tensor":
            ach = r["work"] / (r["total_ms"] * 1e-3) / 1e12
            peak, unit = peaks["tf_sustained"], "TFLOP/s"
        else:
            ach = r["work"] / (r["total_ms"] * 1e-3) / 1e9
            peak, unit = peaks["hbm_gbs"], "GB/s"
        return {"kernel": name, "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
                "traffic": None, "launches_per_step": r["launches"] / args.steps, "avg_launch_ms": per_launch_ms,
                "share_of_step": r["total_ms"] / prof_ms_total, "peak_source": peaks["source"]}

    roofs = [x for x in (roof("gemm_tensor", "tensor"), roof("gemm_hbm", "hbm"), roof("vit_attention", "tensor"),
                         roof("dec_i2t_layer", "hbm"), roof("dec_i2t_layer_shared", "hbm"),
                         roof("dec_t2i", "hbm"), roof("dec_t2i_shared", "tensor"),
                         roof("mask_post_write", "hbm"), roof("mask_post_stats", "hbm")) if x]
    split_mode = os.environ.get("CSAM_PRECISION", "x3") != "x1"
    for r in roofs:
        if r["bound"] == "tensor":
            # achieved counts ALGORITHMIC flops; in the hi/lo split mode the tensor cores execute 3 MMAs per
            # algorithmic GEMM MMA (hi*hi + lo*hi + hi*lo) and 2.5 per attention MMA (QK^T x3, P*V x2: the
            # probabilities are a single fp16), so the pipe is that much busier than `frac` says
            from crowdsam_b200 import ops as _ops
            attn_mult = {1: 3.0, 0: 2.5}.get(_ops.ATTN_PSPLIT, 2.0)     # P*V as 3 / 2 / 1 MMAs (ops.ATTN_PSPLIT)
            r["mma_multiplier"] = (attn_mult if r["kernel"] == "vit_attention" else 3) if split_mode else 1
            r["frac_executed"] = r["frac"] * r["mma_multiplier"]
        t = traffic.get(r["kernel"])
        if t:
            r["traffic"] = t
    dominant = max(roofs, key=lambda r: r["share_of_step"]) if roofs else None
    if post_roof:
        roofs.append(post_roof)

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        gen, kind, what = make_cpu_model()
        secs, _ = time_cpu_images(gen, args.warmup, 1, 0.0)
        cpu = {"value": 1.0 / secs[0], "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
               "sample": f"{what}; ONE whole image of this workload end to end, un-extrapolated ({secs[0]:.1f} s), "
                         f"{CPU_POINTS_PER_BATCH} prompts per decoder batch"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16x3(fp32-accurate hi/lo split), fp32 accumulate" if os.environ.get("CSAM_PRECISION", "x3") != "x1" else "f16, fp32 accumulate",
            "data": "synthetic",
            "config": bench_config(arch, world),
            "run": {"points_per_batch": args.points_per_batch, "masks_into_nms": nk[0], "detections": nk[1],
                    # timed legs: SAM encoder and DINOv2 on two streams (engine.two_streams_enabled); the per-class
                    # roofline pass below times every launch serialised on one stream
                    "encoder_streams": 2 if os.environ.get("CSAM_TWO_STREAMS", "1") != "0" and not graphs.usable() else 1,
                    "attn_pv_mmas": {1: 3, 0: 2}.get(ops.ATTN_PSPLIT, 1),
                    # encoder GEMMs: 2 = cta_group::2 kernel with 256x128 pair tiles (library default), 0 = single-CTA
                    "gemm_pair_mode": int(os.environ.get("CSAM_GEMM_PAIR", "2"))},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(imgs_np[0].nbytes),
                    "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches), "graph_captures": int(graphs.captures), "setup_steps_before_warmup": int(priming), "clocks": clocks, "roofline": dominant, "rooflines": roofs,
            "kernel_ms_per_step": {k: v["total_ms"] / args.steps for k, v in prof.items()},
            "profiled_pass_ms_per_step": prof_ms_total / args.steps,
            "step_minus_profiled_kernels_ms": ms_total / args.steps - sum(v["total_ms"] for k, v in prof.items()
                                                                  if k not in ("gemm_tensor", "gemm_hbm")) / args.steps,
            "cpu_baseline": cpu}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
