#!/usr/bin/env python
"""bench.py -- the measurement contract of contextgs_b200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): rendered frames/s @1080p (+ anchor Mbit/s scored, reported beside it).
Workload at every N: BASELINE.json configs[2] -- the mipnerf360/bicycle-shaped synthetic scene,
1.5 M anchors, 1920x1080, decoded model (the reference's published-FPS path, train.py:409-414:
prefilter_voxel + generate_neural_gaussians + rasterize per frame).  One STEP = one frame of one
camera.  With N > 1 every rank holds a replica and renders its own cameras (weak scaling, no
data-path collective -- SURVEY.md 8e).

  value : frames/s with the camera matrices already resident on the device.
  e2e   : the same through the public `render()` call with the camera in HOST memory and every
          rendered image copied back to pinned host memory inside the timed region (second stream,
          double buffered, all copies complete before the closing event).
  roofline : dominant kernel (by summed device time in a separate stage-timed pass with CUDA
          events on the launching stream), algorithmic bytes per SURVEY.md 8d / DESIGN.md.
  cpu_baseline / --impl reference : the CPU oracle (oracle/, OpenMP + torch CPU threads) on the
          same frame workload, time-bounded.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

N_ANCHORS = 1_500_000
SCENE_KIND = "bicycle"
GAUSSIAN_SCALE = 3.0     # base Gaussian size multiplier of the synthetic scene (instances / Gaussian ~ real scenes)
N_CAMERAS = 16
METRIC = "rendered frames/sec @1080p"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--anchors", type=int, default=N_ANCHORS)
    ap.add_argument("--config", type=int, default=2, choices=[2, 4],
                    help="BASELINE.json configs index: 2 = the headline workload (1.5 M anchors), 4 = the data-parallel "
                         "training shape (2 M anchors, one camera per rank); sets --anchors unless given explicitly")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the full-size GPU-vs-oracle comparison")
    ap.add_argument("--no-reference-gpu", action="store_true", help="skip the reference-Python-on-CUDA baseline extras")
    ap.add_argument("--cpu-budget-s", type=float, default=150.0, help="time budget of the reference arm")
    args = ap.parse_args()
    if args.config == 4 and args.anchors == N_ANCHORS:
        args.anchors = 2_000_000
    return args


# ------------------------------------------------------------------------------------ helpers

def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        if self.index is None:   # ranks other than 0: eight pollers at 50 Hz would only add host and driver load
            self.n_before = 0
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        t0 = time.time()
        while self.proc is not None and not self.lines and time.time() - t0 < 3.0:
            time.sleep(0.01)  # nvidia-smi needs ~0.5 s to start: do not begin timing before it samples
        self.n_before = len(self.lines)
        return self

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines[getattr(self, "n_before", 0):]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def make_inputs(n_anchors):
    """Synthetic scene + decoded values + cameras (CPU tensors; identical bits for both arms)."""
    from contextgs_b200 import synthetic
    scene = synthetic.make_scene(SCENE_KIND, n_anchors, seed=0, gaussian_scale=GAUSSIAN_SCALE)
    dec = synthetic.decoded_scene(scene)
    cams = synthetic.make_cameras(SCENE_KIND, N_CAMERAS, device="cpu")
    return scene, dec, cams


def make_model(scene, device):
    """Random-init weights of the reference architecture, seed 6 (identical on every rank and arm:
    nn.Linear initialises on the CPU generator before the move to `device`)."""
    from contextgs_b200.gaussian_model import GaussianModel
    torch.manual_seed(6)
    m = GaussianModel.from_tensors(scene, device=device)
    with torch.no_grad():  # break the symmetric EntropyBottleneck init a little (as training would)
        g = torch.Generator().manual_seed(9)
        for plist in (m.latent_codec.matrices, m.latent_codec.biases, m.latent_codec.factors):
            for p in plist:
                p.add_((torch.randn(p.shape, generator=g) * 0.3).to(p.device))
    return m


def cam_to(cam, device):
    from types import SimpleNamespace
    d = dict(vars(cam))
    for k in ("world_view_transform", "full_proj_transform", "camera_center"):
        d[k] = d[k].to(device).contiguous()
    return SimpleNamespace(**d)


ALGO_BYTES = {
    # SURVEY.md 8d / DESIGN.md section 4: compulsory fp32 traffic of an ideal fused implementation
    "visible_filter": lambda c: 37 * c["N"] + 4 * c["Nv"],   # fused prefilter: anchor 12 + scaling row 24 + mask 1; index list
    "compact_indices": lambda c: 1 * c["N"] + 4 * c["Nv"],
    "neural_gaussians_fwd": lambda c: 396 * c["Nv"] + 50 * c["Nv"] + 56 * c["P"],
    "preprocess": lambda c: 124 * c["P"],                    # 56 in + 60 out + 8 (packed tile rectangle)
    "depth_sort": lambda c: c["P"] * (4 + 4 * 16),
    # binning (csrc/raster_binning.cu): S = (super-tile, Gaussian) pairs, ~1.4 per Gaussian
    "scan_emit_pairs": lambda c: 12 * c["P"] + 8 * c["S"],
    "pair_sort": lambda c: c["S"] * (4 + 16 * max(1, -(-max(1, (c["supertiles"] - 1).bit_length()) // 8))),
    "bin_expand": lambda c: 2 * 16 * c["S"] + 4 * c["R"] + 12 * c["tiles"],
    "render_fwd": lambda c: 40 * c["R"] + 20 * c["W"] * c["H"],
    "render_bwd": lambda c: 40 * c["R"] + 20 * c["W"] * c["H"] + 44 * c["P"],
    "preprocess_bwd": lambda c: 140 * c["P"],
}


# ------------------------------------------------------------------------------------ CPU oracle arm

def oracle_model(scene, dec, m_cpu):
    """Oracle-side decoded model sharing the product model's weights."""
    from oracle import entropy_ref
    pc = entropy_ref.make_model(scene)
    W = lambda seq: [seq[0].weight.detach(), seq[0].bias.detach(), seq[2].weight.detach(), seq[2].bias.detach()]
    pc.mlps = {"opacity": W(m_cpu.mlp_opacity), "cov": W(m_cpu.mlp_cov), "color": W(m_cpu.mlp_color),
               "grid": [W(s) for s in m_cpu.mlp_grid]}
    eb = pc.latent_codec
    eb.matrices = [p.detach() for p in m_cpu.latent_codec.matrices]
    eb.biases = [p.detach() for p in m_cpu.latent_codec.biases]
    eb.factors = [p.detach() for p in m_cpu.latent_codec.factors]
    eb.quantiles = m_cpu.latent_codec.quantiles.detach()
    pc.dec = dec
    return pc


def oracle_frame(pc, cam):
    """prefilter_voxel + generate_neural_gaussians + rasterize forward on the CPU oracle."""
    from oracle import entropy_ref, raster_ref
    d = pc.dec
    st = raster_ref.make_settings(cam.image_width, cam.image_height, math.tan(cam.FoVx * 0.5),
                                  math.tan(cam.FoVy * 0.5), (0, 0, 0), 1.0, cam.world_view_transform.numpy(),
                                  cam.full_proj_transform.numpy())
    N = d["anchor"].shape[0]
    ident = np.zeros((N, 4), np.float32)
    ident[:, 0] = 1
    radii = raster_ref.preprocess(st, d["anchor"].numpy(), d["scaling"][:, :3].numpy(), ident, filter_only=True)
    vis = torch.from_numpy(radii > 0)
    g = entropy_ref.generate_neural_gaussians(pc, cam.camera_center, d["anchor"][vis], d["feat"][vis],
                                              d["offsets"][vis], d["scaling"][vis], d["masks"][vis])
    out = raster_ref.forward(st, g["xyz"].numpy(), g["color"].numpy(), g["opacity"].numpy(), g["scaling"].numpy(),
                             g["rot"].numpy())
    return out


def run_reference(args, rank, world):
    """`--impl reference`: the reference path's CPU restatement (oracle/) on the host cores."""
    if rank != 0:
        return
    from contextgs_b200.gaussian_model import GaussianModel  # parameter container only (CPU tensors, no kernels)
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)  # torchrun exports OMP_NUM_THREADS=1; the oracle is loaded below
    torch.set_num_threads(cores)
    scene, dec, cams = make_inputs(args.anchors)
    m_cpu = make_model(scene, "cpu")
    pc = oracle_model(scene, dec, m_cpu)
    t_budget = args.cpu_budget_s
    t_start = time.perf_counter()
    done_w = 0
    for i in range(min(args.warmup, 1)):  # CPU code has no clocks / caches to warm beyond one frame
        oracle_frame(pc, cams[i % len(cams)])
        done_w += 1
    per_frame = time.perf_counter() - t_start if done_w else None
    k_max = args.steps
    if per_frame is not None:
        k_max = max(1, min(args.steps, int((t_budget - per_frame) / max(per_frame, 1e-3))))
    t0 = time.perf_counter()
    k = 0
    while k < k_max:
        oracle_frame(pc, cams[(done_w + k) % len(cams)])
        k += 1
        if time.perf_counter() - t0 > t_budget:
            break
    dt = time.perf_counter() - t0
    fps = k / dt
    sample = (f"{k} full frames (all {args.anchors} anchors, 1920x1080) of the requested {args.steps}; "
              f"time-bounded to {t_budget:.0f} s of CPU work")
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": k,
        "warmup": done_w, "ms_per_step": 1e3 * dt / k, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference rasterizer / compressai / torchac sources are not in /root/reference; this arm runs the "
                "CPU oracle restatement (oracle/raster_ref.c with OpenMP + oracle/entropy_ref.py on torch CPU threads)",
    }
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {"workload": f"BASELINE configs[{args.config}]: mipnerf360/bicycle-shaped synthetic scene, {args.anchors} anchors x 10 "
                        "offsets, 1920x1080, decoded model, per frame prefilter_voxel + generate_neural_gaussians + "
                        "rasterize forward; camera batch sharded over ranks",
            "anchors": args.anchors, "image": [1920, 1080], "cameras": N_CAMERAS, "gaussian_scale": GAUSSIAN_SCALE,
            "l2_policy": "per-frame inputs (anchor attributes, 464 B x anchors = 696 MB at 1.5 M) exceed the 126 MB L2"}


# ------------------------------------------------------------------------------------ product arm

def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from contextgs_b200 import _lib
    from contextgs_b200.renderer import prefilter_voxel, render

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: contextgs_b200 has no CPU path (use --impl reference for "
                         "the CPU oracle arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()

    scene, dec, cams_cpu = make_inputs(args.anchors)
    pc = make_model(scene, dev)
    pc_train = pc  # non-decoded replica for the entropy / training extras
    pc = make_model(scene, dev).replace_with_decoded(**{k: v.to(dev) for k, v in dec.items()})
    pc.eval()
    cams_dev = [cam_to(c, dev) for c in cams_cpu]
    pipe = type("Pipe", (), {"debug": False})()
    bg = torch.zeros(3, device=dev)
    W, H = cams_cpu[0].image_width, cams_cpu[0].image_height
    my_cam = lambda i: (i * world + rank) % N_CAMERAS  # rank r renders cameras r, r+world, ...

    stats = {}

    def frame(cam):
        with torch.no_grad():
            vis = prefilter_voxel(cam, pc, pipe, bg)
            out = render(cam, pc, pipe, bg, visible_mask=vis)
        stats["P"] = out["radii"].shape[0]
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, after_warmup=None, before_end=None):
        for i in range(warmup):
            fn(i)
        barrier()
        if after_warmup is not None:
            after_warmup()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        if before_end is not None:
            before_end()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def timed_median(fn, steps, warmup):
        """Per-iteration CUDA-event times, median over the iterations x `steps` (max over ranks): the side figures in
        `extras` should not move with a one-off allocator / clock hiccup in a 3-10 iteration sample."""
        for i in range(warmup):
            fn(i)
        barrier()
        times = []
        for i in range(steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(warmup + i)
            e1.record()
            e1.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = float(np.median(times))
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms * steps

    # ---- device-resident pass (value) ------------------------------------------------------
    _lib.launch_counts(reset=True)
    clk = ClockSampler(local_rank if rank == 0 else None)   # the line is rank 0's: its GPU is the one sampled
    clk.__enter__()
    ms = timed(lambda i: frame(cams_dev[my_cam(i)]), args.steps, args.warmup,
               after_warmup=lambda: _lib.launch_counts(reset=True))
    launches_total = sum(_lib.launch_counts().values())  # kernels of libcontextgs_b200.so inside the timed region
    fps = world * args.steps / (ms * 1e-3)

    # ---- end-to-end pass: camera in host memory, image to pinned host memory -----------------
    # Every frame's image is copied to pinned host memory on a second stream, double buffered: the copy of
    # frame i (24.9 MB over PCIe) overlaps the kernels of frame i+1; the timed region ends only after the
    # last copy has landed (the main stream waits for the copy stream before the closing event).
    host_imgs = [torch.empty((3, H, W), dtype=torch.float32).pin_memory() for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    copy_done = [None, None]
    cam_bytes = 4 * (16 + 16 + 3)

    main_stream = torch.cuda.current_stream()

    def e2e_step(i):
        out = frame(cams_cpu[my_cam(i)])              # matrices read on the host, passed by value to the kernels;
        b = i & 1                                     # render() returns after its status read-back: pixels complete
        if copy_done[b] is not None:
            copy_done[b].synchronize()                # the caller has consumed host buffer b (two frames ago)
        torch.cuda.set_stream(copy_stream)            # (cheaper than the context manager on a per-frame host path)
        host_imgs[b].copy_(out["render"], non_blocking=True)
        if copy_done[b] is None:
            copy_done[b] = torch.cuda.Event()         # two events, reused for the whole run
        copy_done[b].record(copy_stream)
        torch.cuda.set_stream(main_stream)
        out["render"].record_stream(copy_stream)
    ms_e2e = timed(e2e_step, args.steps, max(args.warmup, 3),
                   before_end=lambda: torch.cuda.current_stream().wait_stream(copy_stream))
    fps_e2e = world * args.steps / (ms_e2e * 1e-3)
    clk.__exit__()
    clocks = clk.summary()  # sampled over the device-resident and the end-to-end timed regions

    # ---- stage-timed pass (separate, so that event records do not perturb `value`) -------------
    _lib.stage_timing(True)
    n_prof = max(4, min(args.steps, 16))
    for i in range(n_prof):
        out = frame(cams_dev[my_cam(i)])
    torch.cuda.synchronize()
    stage_ms, stage_n = _lib.stage_timing_read()
    _lib.stage_timing(False)
    saved = None
    # instance counts of one representative frame (camera 0 of this rank) for the byte model
    with torch.no_grad():
        vis = prefilter_voxel(cams_dev[my_cam(0)], pc, pipe, bg)
        out = render(cams_dev[my_cam(0)], pc, pipe, bg, visible_mask=vis)
    counts = {"N": args.anchors, "Nv": int(vis.sum()), "P": int(out["radii"].shape[0]), "W": W, "H": H,
              "tiles": ((W + 15) // 16) * ((H + 15) // 16)}
    from contextgs_b200 import rasterizer as _r
    st = _r._state(dev)
    counts["R"] = int(getattr(st, "last_num_rendered", 0) or 0)
    counts["S"] = int(getattr(st, "last_num_pairs", 0) or 0)
    counts["supertiles"] = ((W + 127) // 128) * ((H + 63) // 64)
    per_frame = {k: stage_ms[k] / n_prof for k in stage_ms if stage_n[k] > 0}
    top = max(per_frame, key=per_frame.get)
    peak, peak_src = peaks()
    launches_per_frame_top = stage_n[top] / n_prof
    top_ms = per_frame[top] / launches_per_frame_top
    algo = ALGO_BYTES.get(top, lambda c: 0)(counts) / launches_per_frame_top
    achieved = algo / (top_ms * 1e-3) / 1e9 if top_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):  # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` launch (committed)
        with open(tpath) as f:
            t = json.load(f).get(top)
        if t:
            traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
    roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo, "ms_per_launch": top_ms}
    # Instruction-issue roofline of the same kernel: the blend loop is bound by FP32 / SFU instruction issue per
    # (pixel, instance), not by HBM.  Executed warp instructions come from the tracked ncu capture, scaled
    # by the instance count; peak = SMs x 4 schedulers x the SM clock sampled during the timed region.
    if os.path.exists(tpath):
        with open(tpath) as f:
            t = json.load(f).get(top) or {}
        if t.get("inst_executed") and counts.get("R"):
            sm_hz = 1e6 * float(clocks.get("sm_mhz") or 1965.0)
            n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
            inst = t["inst_executed"] * counts["R"] / float(t["inst_at_R"])
            a_issue, p_issue = inst / (top_ms * 1e-3) / 1e9, n_sm * 4 * sm_hz / 1e9
            roofline["issue"] = {"achieved": a_issue, "peak": p_issue, "unit": "G warp-instructions/s",
                                 "frac": a_issue / p_issue, "warp_instructions_per_launch": inst,
                                 "source": t.get("inst_source")}
            roofline["true_limiter"] = "issue" if a_issue / p_issue > achieved / peak else "hbm"
    roofline["note"] = ("`bound`/`frac` are the HBM roofline the contract asks for; render_fwd / render_bwd are limited by "
                        "instruction issue (see `issue`), every other frame stage by HBM or latency (see roofline_stages)")
    # per-stage table: algorithmic bytes of SURVEY.md 8d / DESIGN.md section 4 over the stage's measured time
    stage_table = {}
    for k, v in per_frame.items():
        if k in ALGO_BYTES and v > 0:
            ab = ALGO_BYTES[k](counts)
            stage_table[k] = {"ms": round(v, 4), "algorithmic_mb": round(ab / 1e6, 2), "gb_per_s": round(ab / (v * 1e-3) / 1e9, 1),
                              "frac_of_hbm_peak": round(ab / (v * 1e-3) / 1e9 / peak, 4)}

    line = {
        "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args),
        "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": cam_bytes,
                "d2h_bytes_per_step": 3 * H * W * 4, "ms_per_step": ms_e2e / args.steps,
                "host_numa_node_rank0": numa_node},
        "gpu_launches": int(launches_total), "clocks": clocks, "roofline": roofline,
        "stage_ms_per_frame": {k: round(v, 4) for k, v in sorted(per_frame.items(), key=lambda kv: -kv[1])},
        "roofline_stages": stage_table, "counts": counts,
    }

    # ---- extras: training-side rasterizer fwd+bwd and the entropy scoring pass -----------------
    if not args.no_extras:
        line["extras"] = extras(args, pc, pc_train, cams_dev, my_cam, pipe, bg, timed_median, world, dev)
        if world == 1 and not args.no_reference_gpu:
            try:
                line["extras"]["reference_gpu"] = reference_gpu_baseline(args, scene, pc, pc_train, cams_dev, pipe, bg, dev,
                                                                         timed_median)
            except Exception as e:
                line["extras"]["reference_gpu"] = {"error": f"{type(e).__name__}: {e}"}

    # ---- full-size parity against the CPU oracle (rank 0, N = 1 only) ----------------------------
    if world == 1 and rank == 0 and not args.no_parity:
        try:
            line["parity"] = parity_block(args, scene, dec, cams_cpu, pc, pc_train, cams_dev[0], pipe, bg, dev)
        except Exception as e:  # the comparison must never take the measurement down with it
            line["parity"] = {"ok": False, "error": f"{type(e).__name__}: {e}"}

    # ---- CPU baseline beside it (rank 0, N = 1 only) --------------------------------------------
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, scene, dec, cams_cpu)
    elif "cpu_baseline" not in line:
        line["cpu_baseline"] = None

    # the data-parallel training figures go LAST in the line (a tail of the output keeps them)
    if world > 1 and "extras" in line:
        line["dp_training"] = {k: line["extras"][k] for k in list(line["extras"]) if "dp_allreduce" in k or k == "grad_bucket_mb"}
        line["dp_training"]["note"] = ("iterations/s summed over the ranks: forward + backward of one camera per rank + ONE NCCL "
                                       "all-reduce of the flat fp32 gradient bucket per step (BASELINE configs[4] shape of work "
                                       f"at {args.anchors} anchors; `--config 4` selects that config's 2 M)")
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def extras(args, pc, pc_train, cams_dev, my_cam, pipe, bg, timed, world, dev):
    from contextgs_b200.renderer import prefilter_voxel
    ex = {}
    H, W = cams_dev[0].image_height, cams_dev[0].image_width
    gt = torch.rand(3, H, W, device=dev)
    k = max(4, args.steps // 4)

    from contextgs_b200.renderer import render
    from contextgs_b200.loss_utils import l1_ssim

    def train_step(i, step):
        """One training iteration of train.py:158-211 without the optimizer: prefilter, render (G1 + rasterizer,
        for step > 10000 also the whole-scene context model), L1 + lambda * bit_per_param, backward."""
        cam = cams_dev[my_cam(i)]
        with torch.no_grad():
            vis = prefilter_voxel(cam, pc_train, pipe, bg)
        out = render(cam, pc_train, pipe, bg, visible_mask=vis, retain_grad=False, step=step)
        Ll1, ssim_v = l1_ssim(out["render"], gt)                       # train.py:200-204, lambda_dssim = 0.2
        loss = 0.8 * Ll1 + 0.2 * (1.0 - ssim_v) + 0.01 * out["scaling"].prod(dim=1).mean()
        if out["bit_per_param"] is not None:
            loss = loss + 0.004 * out["bit_per_param"]
        loss.backward()
        for p in pc_train.parameters():
            p.grad = None
        for p in (pc_train._anchor, pc_train._anchor_feat, pc_train._offset, pc_train._scaling, pc_train._mask,
                  pc_train._hyper_latent):
            p.grad = None

    pc_train.train()
    ms = timed(lambda i: train_step(i, 100), k, 3)
    ex["train_iter_per_s_render_only"] = world * k / (ms * 1e-3)
    ms = timed(lambda i: train_step(i, 20000), k, 3)
    ex["train_iter_per_s_with_context_model"] = world * k / (ms * 1e-3)
    # stage-timed pass of the full training iteration (library-side CUDA events around every stage) with the HBM
    # roofline of the two tcgen05 backward stages: algorithmic bytes per SURVEY.md 8d (backward = 2 x forward)
    from contextgs_b200 import _lib as _L
    _L.stage_timing(True)
    for i in range(3):
        train_step(i, 20000)
    torch.cuda.synchronize()
    st_ms, st_n = _L.stage_timing_read()
    _L.stage_timing(False)
    ex["train_stage_ms_per_iter"] = {kk: round(v / 3, 4) for kk, v in sorted(st_ms.items(), key=lambda kv: -kv[1]) if st_n[kk] > 0}
    with torch.no_grad():
        vis0 = prefilter_voxel(cams_dev[my_cam(0)], pc_train, pipe, bg)
    nv, n_all = int(vis0.sum()), int(pc_train._anchor.shape[0])
    hbm_peak = peaks()[0]
    bwd_roof = {}
    for key, bytes_ in (("neural_gaussians_bwd", 2.0 * (446.0 * nv + 56.0 * 4.7 * nv)), ("context_level_bwd", 2.0 * 1024.0 * n_all)):
        ms_k = ex["train_stage_ms_per_iter"].get(key)
        if ms_k:
            bwd_roof[key] = {"ms": ms_k, "algorithmic_mb": round(bytes_ / 1e6, 1), "gb_per_s": round(bytes_ / (ms_k * 1e-3) / 1e9, 1),
                             "frac_of_hbm_peak": round(bytes_ / (ms_k * 1e-3) / 1e9 / hbm_peak, 4)}
    ex["train_backward_roofline"] = bwd_roof
    if world > 1:
        # data-parallel training step (BASELINE configs[4]): every rank renders its own camera, the
        # gradients of all parameters land in ONE flat fp32 bucket that is all-reduced (NCCL, sum / world) per step
        from contextgs_b200.distributed import GradientBucket
        params = [pp for pp in list(pc_train.parameters()) + [pc_train._anchor_feat, pc_train._offset, pc_train._scaling,
                                                               pc_train._mask, pc_train._hyper_latent]
                  if pp.requires_grad]
        seen, uniq = set(), []
        for pp in params:
            if id(pp) not in seen:
                seen.add(id(pp))
                uniq.append(pp)
        bucket = GradientBucket(uniq).attach()

        def train_step_dp(i, step):
            bucket.zero()
            cam = cams_dev[my_cam(i)]
            with torch.no_grad():
                vis = prefilter_voxel(cam, pc_train, pipe, bg)
            out = render(cam, pc_train, pipe, bg, visible_mask=vis, retain_grad=False, step=step)
            Ll1, ssim_v = l1_ssim(out["render"], gt)
            loss = 0.8 * Ll1 + 0.2 * (1.0 - ssim_v) + 0.01 * out["scaling"].prod(dim=1).mean()
            if out["bit_per_param"] is not None:
                loss = loss + 0.004 * out["bit_per_param"]
            loss.backward()
            bucket.all_reduce()
        ms = timed(lambda i: train_step_dp(i, 100), k, 3)
        ex["train_iter_per_s_render_only_dp_allreduce"] = world * k / (ms * 1e-3)
        ms = timed(lambda i: train_step_dp(i, 20000), k, 3)
        ex["train_iter_per_s_with_context_model_dp_allreduce"] = world * k / (ms * 1e-3)
        ex["grad_bucket_mb"] = bucket.flat.numel() * 4 / 1e6
        for pp in uniq:
            pp.grad = None
    # the photometric loss alone: fused kernels vs the reference's expression (5 grouped 11x11 convolutions + autograd)
    img = torch.rand(3, H, W, device=dev)

    def loss_fused(i):
        a = img.detach().requires_grad_(True)
        l1, sv = l1_ssim(a, gt)
        (0.8 * l1 + 0.2 * (1.0 - sv)).backward()

    def loss_torch(i):
        import torch.nn.functional as F
        a = img.detach().requires_grad_(True)
        g1 = torch.tensor([math.exp(-(x - 5) ** 2 / 4.5) for x in range(11)], device=dev)
        g1 = (g1 / g1.sum()).unsqueeze(1)
        w = g1.mm(g1.t()).expand(3, 1, 11, 11).contiguous()
        c = lambda t: F.conv2d(t, w, padding=5, groups=3)
        mu1, mu2 = c(a), c(gt)
        s11, s22, s12 = c(a * a) - mu1 * mu1, c(gt * gt) - mu2 * mu2, c(a * gt) - mu1 * mu2
        sm = ((2 * mu1 * mu2 + 1e-4) * (2 * s12 + 9e-4)) / ((mu1 * mu1 + mu2 * mu2 + 1e-4) * (s11 + s22 + 9e-4))
        (0.8 * (a - gt).abs().mean() + 0.2 * (1.0 - sm.mean())).backward()
    ex["loss_fwd_bwd_ms_fused"] = timed(loss_fused, 20, 3) / 20
    ex["loss_fwd_bwd_ms_torch_expression"] = timed(loss_torch, 20, 3) / 20
    ex["train_note"] = ("forward + backward of one camera per rank on the NON-decoded model (train.py:158-211 without "
                        "optimizer.step; loss = 0.8 L1 + 0.2 (1 - SSIM) + 0.01 scaling reg (+ lambda bit_per_param), fused "
                        "L1/SSIM kernels): render_only = step <= 3000 regime; with_context_model = step > 10000 regime "
                        "(3-level context model over all anchors, forward and backward, every iteration)")

    # entropy scoring: estimate_final_bits = 3-level context model + likelihoods over ALL anchors
    pc_train.eval()
    res = {}

    def score(i, drop_plan):
        if drop_plan and hasattr(pc_train, "_cgs_level_plan"):
            del pc_train._cgs_level_plan
        res["sums"] = pc_train.estimate_final_bits(return_values=True)
    score(0, True)  # find_divide_scale once (cached in pc.level_scale like the reference, gaussian_model.py:1559)
    ke = 5
    ms = timed(lambda i: score(i, True), ke, 2)
    bits = float(sum(res["sums"][1:5]))
    ex["anchor_mbits_per_s"] = world * bits * ke / (ms * 1e-3) / 1e6
    ex["entropy_pass_ms"] = ms / ke
    ms = timed(lambda i: score(i, False), ke, 2)
    ex["anchor_mbits_per_s_cached_level_plan"] = world * bits * ke / (ms * 1e-3) / 1e6
    ex["entropy_pass_ms_cached_level_plan"] = ms / ke
    ex["scored_mbits"] = bits / 1e6
    from contextgs_b200 import _lib
    _lib.stage_timing(True)
    for i in range(3):
        score(i, False)
    torch.cuda.synchronize()
    st_ms, st_n = _lib.stage_timing_read()
    _lib.stage_timing(False)
    ex["entropy_stage_ms_per_pass"] = {k: round(v / 3, 4) for k, v in st_ms.items() if st_n[k] > 0}
    # the real bitstream: conduct_encoding's work (context model + range coding + stream packing) and its inverse
    from contextgs_b200 import codec
    enc_box = {}

    def encode_est(i):
        enc_box["enc"] = codec.encode_model(pc_train)
    kc = 5
    ex["encode_with_bit_estimate_ms"] = timed(encode_est, kc, 2) / kc

    def encode(i):   # what conduct_encoding runs: the reference's encoder reports stream sizes, not entropy estimates
        enc_box["enc"] = codec.encode_model(pc_train, estimate_bits=False)
    ms = timed(encode, kc, 2)
    enc = enc_box["enc"]
    real_bits = codec.encoded_bits(enc)
    ex["anchor_mbits_per_s_encoded"] = world * real_bits["total"] * kc / (ms * 1e-3) / 1e6
    ex["encode_ms"] = ms / kc
    ex["encoded_mbits"] = {k: round(v / 1e6, 3) for k, v in real_bits.items()}

    def decode(i):
        from contextgs_b200.gaussian_model import GaussianModel
        dec = enc_box.get("dec")
        if dec is None:
            dec = enc_box["dec"] = GaussianModel(device=dev)
            dec.load_state_dict({k: v for k, v in pc_train.state_dict().items() if not k.startswith("_")}, strict=False)
        enc_box["out"] = codec.decode_model(dec, enc.meta, enc.anchor_q, enc.mask_bytes, enc.mask_lens, enc.hyper_bytes,
                                            enc.hyper_lens, enc.levels)
    ms = timed(decode, kc, 2)
    ex["anchor_mbits_per_s_decoded"] = world * real_bits["total"] * kc / (ms * 1e-3) / 1e6
    ex["decode_ms"] = ms / kc
    ex["codec_round_trip_exact"] = bool(torch.equal(enc_box["out"]["feat"], enc.quantised["feat"]) and
                                        torch.equal(enc_box["out"]["scaling"], enc.quantised["scaling"]) and
                                        torch.equal(enc_box["out"]["hyper"], enc.quantised["hyper"]))
    if world > 1:
        # BASELINE configs[3]: ONE model, anchors sharded over the ranks by dependency root; no data-path collective while
        # coding, one scalar all-reduce for the size, one all-reduce (sum) to assemble the decoded attributes
        box = {}

        def enc_sh(i):
            box["enc"], box["bits"] = codec.encode_model_sharded(pc_train)
        ms = timed(enc_sh, kc, 2)
        ex["anchor_mbits_per_s_encoded_sharded"] = box["bits"] * kc / (ms * 1e-3) / 1e6
        ex["encode_sharded_ms"] = ms / kc
        from contextgs_b200.gaussian_model import GaussianModel
        dec_sh_model = GaussianModel(device=dev)
        dec_sh_model.load_state_dict({k: v for k, v in pc_train.state_dict().items() if not k.startswith("_")}, strict=False)

        def dec_sh(i):
            box["out"] = codec.decode_model_sharded(dec_sh_model, box["enc"])
        ms = timed(dec_sh, kc, 2)
        ex["anchor_mbits_per_s_decoded_sharded"] = box["bits"] * kc / (ms * 1e-3) / 1e6
        ex["decode_sharded_ms"] = ms / kc
        ex["codec_sharded_round_trip_exact"] = bool(torch.equal(box["out"]["feat"], enc.quantised["feat"]) and
                                                    torch.equal(box["out"]["scaling"], enc.quantised["scaling"]))
    ex["codec_note"] = ("encoded = bytes actually produced by the GPU range coder (anchors 16 bit raw + masks + hyper + "
                        "feat / scaling / offsets streams + per-chunk side info) / time of the whole encode_model call "
                        "(level division cached in the model, context model, coding, packing; streams stay in HBM)")
    ex["entropy_note"] = ("anchor Mbit/s = estimated bits (hyper+feat+scaling+masked offsets of every valid anchor, "
                          "what estimate_final_bits sums, gaussian_model.py:1685) / wall time of the full 3-level "
                          "scoring pass; first figure rebuilds the level division every call like the reference")
    return ex


def bind_to_gpu_numa_node(local_rank):
    """One process per GPU: run this process (and, by first touch, place its pinned host buffers) on the NUMA node the GPU
    hangs off -- the end-to-end pass moves 25 MB per frame per rank to host memory, and a buffer on the other socket costs
    inter-socket bandwidth that eight ranks share.  Returns the node, or None when the topology cannot be read."""
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


def parity_block(args, scene, dec, cams_cpu, pc, pc_train, cam_dev, pipe, bg, dev):
    """Full-size parity: the frame `value` is quoted on and the 3-level scoring pass, GPU vs the
    CPU oracle on the SAME inputs at the bench configuration.  The oracle is the checker here, never the thing timed.
      * whole frame (prefilter + G1 + rasterize), oracle chain vs GPU chain: image rel-L2, Gaussian / instance counts
        (a selection `tanh(x) * mask > 0` may flip where |x| is at fp32 rounding level: such a Gaussian has opacity ~ 0
        and cannot change a pixel);
      * rasterizer alone on the GPU's OWN Gaussians: R equal, point_list / ranges bit-exact, image rel-L2;
      * scoring pass: the six bit sums of estimate_final_bits vs oracle/entropy_ref (same level scales)."""
    from contextgs_b200 import rasterizer as _r
    from contextgs_b200.renderer import _scratch, prefilter_voxel, render
    from oracle import entropy_ref, raster_ref
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)
    torch.set_num_threads(cores)
    rel = lambda a, b: float(np.linalg.norm(a.astype(np.float64) - b.astype(np.float64)) /
                             (np.linalg.norm(b.astype(np.float64)) + 1e-30))
    res = {}
    cam = cams_cpu[0]
    m_cpu = make_model(scene, "cpu")
    pc_o = oracle_model(scene, dec, m_cpu)
    t0 = time.perf_counter()
    ref = oracle_frame(pc_o, cam)
    with torch.no_grad():
        vis = prefilter_voxel(cam_dev, pc, pipe, bg)
        out = render(cam_dev, pc, pipe, bg, visible_mask=vis)
    torch.cuda.synchronize()
    sc = _scratch[(dev.index, torch.cuda.current_stream(dev).cuda_stream)]
    st = _r._state(dev)
    P, R = int(out["radii"].shape[0]), int(st.last_num_rendered)
    img = out["render"].cpu().numpy()
    res["frame"] = {"P": P, "P_oracle": int(ref["radii"].shape[0]), "R": R, "R_oracle": int(ref["R"]),
                    "visible_anchors": int(vis.sum()), "image_rel_l2": rel(img, ref["color"]),
                    "final_T_rel_l2": rel(sc.final_T.cpu().numpy(), ref["final_T"])}
    # the rasterizer alone, on the Gaussians the GPU generated
    g = {k: getattr(sc, k)[:P].cpu().numpy() for k in ("xyz", "color", "opacity", "scaling", "rot")}
    stt = raster_ref.make_settings(cam.image_width, cam.image_height, math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5),
                                   (0, 0, 0), 1.0, cam.world_view_transform.numpy(), cam.full_proj_transform.numpy())
    ref2 = raster_ref.forward(stt, g["xyz"], g["color"], g["opacity"], g["scaling"], g["rot"])
    pl = sc.point_list[:R].cpu().numpy().view(np.uint32)
    rg = sc.ranges.cpu().numpy().view(np.uint32)
    res["rasterizer_same_gaussians"] = {
        "R_equal": bool(R == int(ref2["R"])),
        "radii_bit_exact": bool(np.array_equal(out["radii"].cpu().numpy(), ref2["radii"])),
        "point_list_bit_exact": bool(R == int(ref2["R"]) and np.array_equal(pl, ref2["point_list"])),
        "ranges_bit_exact": bool(np.array_equal(rg, ref2["ranges"])),
        "image_rel_l2": rel(img, ref2["color"]),
        "n_contrib_mismatch_frac": float((sc.n_contrib.cpu().numpy().view(np.uint32) != ref2["n_contrib"]).mean())}
    res["frame_seconds_oracle"] = round(time.perf_counter() - t0, 2)
    # scoring pass (3-level context model over every valid anchor), same level scales as the GPU search found
    if not args.no_extras:
        t0 = time.perf_counter()
        pc_train.eval()
        got = pc_train.estimate_final_bits(return_values=True)
        pc_o.level_scale = list(pc_train.level_scale)
        sel = pc_o.get_mask_anchor
        with torch.no_grad():
            want = entropy_ref.multi_scale_generating(
                pc_o, pc_o.get_anchor[sel], pc_o._hyper_latent[sel], pc_o._anchor_feat[sel], pc_o._offset[sel],
                pc_o.get_scaling[sel], pc_o.get_mask[sel], predict_bpp=True, return_sum_bits=True)
        names = ["anchor", "hyper", "feat", "scaling", "offsets", "masks"]
        res["scoring_pass"] = {
            "bit_sums_rel_err": {n: abs(float(a) - float(b)) / max(abs(float(b)), 1e-30) for n, a, b in zip(names, got, want)},
            "level_scale": [float(v) for v in pc_train.level_scale], "seconds_oracle": round(time.perf_counter() - t0, 2)}
        res["scoring_pass"]["max_rel_err"] = max(res["scoring_pass"]["bit_sums_rel_err"].values())
    fr, rs = res["frame"], res["rasterizer_same_gaussians"]
    res["ok"] = bool(fr["image_rel_l2"] < 1e-4 and rs["R_equal"] and rs["point_list_bit_exact"] and rs["ranges_bit_exact"]
                     and rs["image_rel_l2"] < 1e-4 and res.get("scoring_pass", {}).get("max_rel_err", 0.0) < 2e-4)
    res["tolerances"] = "image rel-L2 < 1e-4, tile / sort indices bit-exact, bit sums rtol 2e-4 (north_star)"
    return res


def reference_gpu_baseline(args, scene, pc, pc_train, cams_dev, pipe, bg, dev, timed):
    """Labelled baseline, NOT the product and not `value`: the reference's OWN Python (oracle/_ref/*.refbin, byte-compiled
    from /root/reference in the build container; stand-ins for absent third-party modules in oracle/ref_loader.py)
    with its tensors on CUDA -- its real deployment mode (SURVEY.md 8d "reference GPU").
      * entropy path: scene/gaussian_model.py:1541-1707 `multi_scale_generating(predict_bpp=True, return_sum_bits=True)`
        over every valid anchor = the work of `estimate_final_bits` (:981), against contextgs_b200's scoring pass
        on the same model; the six bit sums are compared as well (parity with the reference's own code at full size).
        Shims: factory calls default to CUDA (`torch.arange(N)[cuda_mask]`, :1571,1590, no longer works in torch 2.x),
        torch.save is a no-op while timing (the reference dumps two debug files per call, :1681-1682), level scales
        are taken from the GPU search (the reference caches them too, :1559);
      * anchor -> Gaussian generation: gaussian_renderer/__init__.py:25-150 (torch / cuBLAS fp32 MLPs, boolean-mask
        compaction) on the decoded model, against the fused tcgen05 kernel, same camera, same visible anchors.
    The rasterizer has no reference GPU baseline: its source is not in the reference tree."""
    from oracle import ref_loader
    if not ref_loader.available():
        return {"unavailable": "oracle/_ref/*.refbin not built (python -m oracle.build_ref needs /root/reference)"}
    from contextgs_b200.neural_gaussians import generate_neural_gaussians
    from contextgs_b200.renderer import prefilter_voxel
    ref = ref_loader.load()
    out = {}

    def share_weights(theirs, ours):
        for name in ("mlp_opacity", "mlp_cov", "mlp_color", "mlp_grid"):
            getattr(theirs, name).load_state_dict(getattr(ours, name).state_dict())
        theirs.latent_codec.load_ref(ours.latent_codec)

    # ---- entropy path ----------------------------------------------------------------------------------------
    theirs = ref_loader.reference_model(scene)
    share_weights(theirs, pc_train)
    theirs.eval()
    theirs.x_bound_min, theirs.x_bound_max = pc_train.x_bound_min.clone(), pc_train.x_bound_max.clone()
    pc_train.eval()
    ours_sums = pc_train.estimate_final_bits(return_values=True)
    theirs.level_scale = list(pc_train.level_scale)
    box = {}
    real_save = torch.save

    def ref_score(i):
        with torch.no_grad(), torch.device(dev):
            sel = theirs.get_mask_anchor
            box["sums"] = ref.gaussian_model.multi_scale_generating(
                theirs, theirs.get_anchor[sel], theirs._hyper_latent[sel], theirs._anchor_feat[sel], theirs._offset[sel],
                theirs.get_scaling[sel], theirs.get_mask[sel], predict_bpp=True, return_sum_bits=True)
    torch.save = lambda *a, **k: None
    try:
        ms = timed(ref_score, 3, 1)
    finally:
        torch.save = real_save
    bits = float(sum(box["sums"][1:5]))
    names = ["anchor", "hyper", "feat", "scaling", "offsets", "masks"]
    out["entropy_pass_ms_reference_gpu"] = ms / 3
    out["anchor_mbits_per_s_reference_gpu"] = bits * 3 / (ms * 1e-3) / 1e6
    out["bit_sums_rel_err_vs_reference_gpu"] = {n: abs(float(a) - float(b)) / max(abs(float(b)), 1e-30)
                                                for n, a, b in zip(names, ours_sums, box["sums"])}
    del theirs
    torch.cuda.empty_cache()

    # ---- anchor -> Gaussian generation -------------------------------------------------------------------------
    theirs = ref_loader.reference_model(scene)
    share_weights(theirs, pc)
    P = torch.nn.Parameter
    theirs._hyper_latent, theirs._anchor_feat, theirs._offset = P(pc._hyper_latent.detach().clone()), \
        P(pc._anchor_feat.detach().clone()), P(pc._offset.detach().clone())
    theirs.decoded_version = True
    theirs._anchor, theirs._scaling, theirs._mask = P(pc._anchor.detach().clone()), P(pc._scaling.detach().clone()), \
        P(pc._mask.detach().clone())
    theirs.eval()
    cam = cams_dev[0]
    with torch.no_grad():
        vis = prefilter_voxel(cam, pc, pipe, bg)
        vis_plain = vis.clone()   # without the attached index list: the reference glue indexes with the bool mask

    def ref_g1(i):
        with torch.no_grad():
            box["g"] = ref.gaussian_renderer.generate_neural_gaussians(cam, theirs, vis_plain, is_training=False)

    def our_g1(i):
        with torch.no_grad():
            box["o"] = generate_neural_gaussians(cam, pc, vis, is_training=False)
    ms_ref = timed(ref_g1, 5, 2)
    ms_our = timed(our_g1, 5, 2)
    out["neural_gaussians_ms_reference_gpu"] = ms_ref / 5
    out["neural_gaussians_ms_ours_same_call"] = ms_our / 5
    g, o = box["g"], box["o"]
    out["neural_gaussians_P"] = [int(o[0].shape[0]), int(g[0].shape[0])]
    if o[0].shape == g[0].shape:
        rel = lambda a, b: float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))
        out["neural_gaussians_rel_l2_vs_reference_gpu"] = {k: rel(o[i], g[i]) for i, k in
                                                           enumerate(("xyz", "color", "opacity", "scaling", "rot"))}
    return out


def cpu_baseline(args, scene, dec, cams_cpu):
    from contextgs_b200.gaussian_model import GaussianModel  # noqa: F401  (parameter container on the CPU)
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)
    torch.set_num_threads(cores)
    m_cpu = make_model(scene, "cpu")
    pc = oracle_model(scene, dec, m_cpu)
    t0 = time.perf_counter()
    k = 0
    while True:
        oracle_frame(pc, cams_cpu[k % len(cams_cpu)])
        k += 1
        if time.perf_counter() - t0 > 12.0 or k >= 8:
            break
    dt = time.perf_counter() - t0
    return {"value": k / dt, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{k} full frame(s) of the same workload ({args.anchors} anchors, 1920x1080) on the CPU oracle "
                      f"(oracle/raster_ref.c OpenMP + oracle/entropy_ref.py torch CPU), {dt:.1f} s"}


if __name__ == "__main__":
    main()
