#!/usr/bin/env python
"""bench.py — ToonCrafter_512 hot path on B200: frames/sec at 320x512x16 frames, DDIM-50, CFG 7.5.

One "step" = one clip: DDIMSampler.sample (50 DDIM steps, cond + uncond UNet evaluation per step) followed by the
two dual-reference VAE decodes of scripts/evaluation/inference.py:262-270 (T = 16, then T = 14).  Weights are
seeded synthetic (no checkpoint / network here), inputs synthetic of the reference's shapes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--ddim-steps 50] [--impl reference]

N > 1 is launched by torchrun (one process per GPU, shared-nothing clips, one NCCL weight broadcast at init).
Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

UNET_TF = 12.603                      # TFLOP per UNet forward per sample (matmul/conv, 2 flops/MAC)
DEC_TF = {16: 37.875, 14: 33.148}     # TFLOP per decode pass
H, W, T = 40, 64, 16                  # latent geometry of 320x512, 16 frames


def clip_tflop(S):
    return 2 * S * UNET_TF + DEC_TF[16] + DEC_TF[14]


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(tflops=float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0))), hbm=float(d["hbm_gbs"]),
                    src="measured (MEASURED_PEAKS.json, sustained cuBLAS bf16)")
    return dict(tflops=1400.0, hbm=6650.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ---------------------------------------------------------------------------------------------------- model
def full_model_config():
    from tiny_config import FULL_DDCONFIG, FULL_UNET, model_config
    return model_config(FULL_UNET, FULL_DDCONFIG)


def build_model(dev, rank, world):
    """Full-size LatentVisualDiffusion with seeded synthetic weights; rank 0 generates, NCCL broadcasts."""
    from tooncrafter_b200 import diffusion, synthetic
    cfg = full_model_config()
    with torch.device("meta"):
        skeleton = diffusion.instantiate_from_config(cfg)
    m = skeleton.to_empty(device=dev)
    m.reset_schedule_buffers()
    with torch.no_grad():
        if rank == 0:
            for k, prm in m.named_parameters():
                prm.copy_(synthetic.synthetic_tensor(k, tuple(prm.shape), 0).to(dev))
        if world > 1:
            from tooncrafter_b200.distributed import broadcast_parameters
            # ONE broadcast of the weights at init; no collective on the data path afterwards
            broadcast_parameters(m, src=0)
    m.perframe_ae = True
    return m.eval()


def host_inputs(seed):
    """Pinned host buffers of one clip's inputs (what scripts/evaluation/inference.py hands to the sampler)."""
    from tiny_config import FULL_DDCONFIG
    from tooncrafter_b200 import synthetic
    x_T, cond, uncond = synthetic.synthetic_inputs(1, T, H, W, 1024, seed=seed)
    ref = synthetic.synthetic_ref_context(FULL_DDCONFIG["ch"], FULL_DDCONFIG["ch_mult"], 8 * H, 8 * W, seed=seed)
    pin = lambda t: t.contiguous().pin_memory()
    return dict(x_T=pin(x_T), ctx_c=pin(cond["c_crossattn"][0]), ctx_u=pin(uncond["c_crossattn"][0]),
                c_concat=pin(cond["c_concat"][0]), ref=[pin(r.half()) for r in ref])


def to_device(hi, dev):
    cc = hi["c_concat"].to(dev, non_blocking=True)
    d = dict(x_T=hi["x_T"].to(dev, non_blocking=True),
             cond={"c_crossattn": [hi["ctx_c"].to(dev, non_blocking=True)], "c_concat": [cc]},
             uncond={"c_crossattn": [hi["ctx_u"].to(dev, non_blocking=True)], "c_concat": [cc]},
             ref=[r.to(dev, non_blocking=True) for r in hi["ref"]])
    return d


def h2d_bytes(hi):
    n = sum(hi[k].numel() * hi[k].element_size() for k in ("x_T", "ctx_c", "ctx_u", "c_concat"))
    return n + sum(r.numel() * r.element_size() for r in hi["ref"])


def run_clip(model, sampler, di, S, fs):
    """The hot path through the reference-facing API: sample() + the two decode_first_stage() calls."""
    samples, _ = sampler.sample(S=S, batch_size=1, shape=[4, T, H, W], conditioning=di["cond"],
                                unconditional_conditioning=di["uncond"], eta=1.0, unconditional_guidance_scale=7.5,
                                x_T=di["x_T"], fs=fs, timestep_spacing="uniform_trailing", guidance_rescale=0.7,
                                verbose=False)
    trimmed = torch.cat([samples[:, :, :1], samples[:, :, 2:-2], samples[:, :, -1:]], dim=2)   # drop frames 1 and 14
    group = getattr(sampler, "latency_group", None)
    if group is None:
        video = model.decode_first_stage(samples, ref_context=di["ref"])
        video2 = model.decode_first_stage(trimmed, ref_context=di["ref"])
        video[:, :, 7:9] = video2[:, :, 6:8]                                                    # inference.py:264-270
        return video
    # latency mode: both ranks hold identical samples; pair-rank 0 decodes the 16 frames, pair-rank 1 the 14-frame
    # variant at the same time and ships its two middle frames over
    import torch.distributed as dist
    r = dist.get_rank(group)
    if r == 0:
        video = model.decode_first_stage(samples, ref_context=di["ref"])
        patch = torch.empty_like(video[:, :, 7:9]).contiguous()
    else:
        video = model.decode_first_stage(trimmed, ref_context=di["ref"])
        patch = video[:, :, 6:8].contiguous()
    dist.broadcast(patch, src=dist.get_global_rank(group, 1), group=group)
    if r == 0:
        video[:, :, 7:9] = patch
    return video


def workload_config(S):
    """The `config` object of the JSON line (both arms name the same workload)."""
    return {"workload": "ToonCrafter_512 320x512x16f DDIM-%d fp16, CFG 7.5 (cond+uncond batched), eta 1.0, "
                        "uniform_trailing, guidance_rescale 0.7, sample() + decode T=16 + decode T=14; "
                        "1 clip per GPU per step; random-init weights, synthetic inputs" % S,
            "l2": "working set per clip (2.9 GB fp16 weights + activations) exceeds the 126 MB L2",
            "baseline_note": "vs_baseline = value / 0.667 frames/s (README.md:222: ~24 s/clip on A100)"}


# ---------------------------------------------------------------------------------------------------- roofline
def gemm_roofline(model, dev):
    """Live per-launch CUDA-event timing of every tc_gemm_kernel launch of one eager UNet forward (B = 2)."""
    from tooncrafter_b200 import ops
    eng = model.model.diffusion_model._engine
    plan = eng.plan_for(2, T, H, W, 77 + 16 * T)
    stream = torch.cuda.current_stream()
    recs = []
    for fn, a, kw in plan.main.calls:
        if fn is ops.conv_gemm:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn(*a, **kw)
            e1.record(stream)
            a_dims, taps, out_dims, n_cols = a[1], a[4], a[6], a[7]
            fl = 2.0 * out_dims[0] * out_dims[1] * out_dims[2] * n_cols * len(taps) * a_dims[3]
            recs.append((e0, e1, fl))
        else:
            fn(*a, **kw)
    torch.cuda.synchronize()
    tot_ms = sum(e0.elapsed_time(e1) for e0, e1, _ in recs)
    tot_fl = sum(fl for _, _, fl in recs)
    return dict(launches=len(recs), avg_launch_us=1e3 * tot_ms / max(len(recs), 1), tflops=tot_fl / tot_ms / 1e9,
                flops_per_launch=tot_fl / max(len(recs), 1), gemm_ms_per_forward=tot_ms)


# ---------------------------------------------------------------------------------------------------- CPU arm
def gemm_traffic():
    """(dram bytes per tc_gemm_kernel launch, provenance) from the committed ncu launch list, or (None, reason)."""
    p = Path(__file__).resolve().parent / "profiles" / "r01_unet_b2_launches_final.gemm_traffic.json"
    try:
        d = json.loads(p.read_text())
        return float(d["dram_bytes_per_launch"]), f"profiles/{p.name}: {d['launches']} launches, {d['source']}"
    except Exception as e:      # noqa: BLE001 - the file is optional evidence, not a dependency of the measurement
        return None, f"no ncu capture committed ({type(e).__name__})"


def cpu_threads():
    """Host threads for the CPU arm: all cores up to 32 (beyond that the fp32 conv/GEMM mix of this UNet stops
    scaling and oversubscribed boxes get slower: 128 threads measured 185 s/forward vs 35 s on 8 dedicated cores)."""
    return max(1, min(os.cpu_count() or 1, 32))


def cpu_unet_forward_seconds(threads):
    """Reference algorithm (oracle port, fp32) on the host cores: ONE full-size UNet forward, B = 1."""
    from oracle import unet_oracle
    from tiny_config import FULL_UNET
    from tooncrafter_b200 import layout, modules, synthetic
    torch.set_num_threads(threads)
    lay = layout.unet_layout(FULL_UNET)
    with torch.device("meta"):
        sk = modules.UNetModel(**FULL_UNET)
    sd = {}
    for k, p in sk.named_parameters():
        # full synthetic init costs minutes of host randn; the timing does not depend on the values
        sd["model.diffusion_model." + k] = torch.empty(p.shape).normal_(0, 0.02) if p.dim() > 1 else torch.ones(p.shape)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 8, T, H, W, generator=g)
    ctx = torch.randn(1, 77 + 16 * T, 1024, generator=g)
    t0 = time.perf_counter()
    unet_oracle.unet_forward(sd, lay, x, torch.tensor([500]), ctx, torch.tensor([10]), "model.diffusion_model.")
    return time.perf_counter() - t0


def reference_arm(args):
    """--impl reference: the reference algorithm on the box's host cores (oracle port; the reference itself is
    pure PyTorch and cannot travel to the GPU box).  Each step = one bounded sample = ONE UNet forward."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = cpu_threads()
    S = args.ddim_steps
    times = []
    for i in range(args.warmup + args.steps):
        dt = cpu_unet_forward_seconds(cores)
        if i >= args.warmup:
            times.append(dt)
    t_fwd = sum(times) / len(times)
    sec_per_clip = t_fwd * clip_tflop(S) / UNET_TF           # extrapolated by algorithmic flops
    fps = 16.0 / sec_per_clip
    sample = (f"{len(times)} x one full-size UNet forward (B=1, fp32, {UNET_TF} of the {clip_tflop(S):.0f} TFLOP of "
              f"a clip); sec/clip extrapolated by flops")
    out = {"impl": "reference", "metric": "frames/sec (320x512x16f, DDIM-%d)" % S, "value": fps, "unit": "frames/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_fwd,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": dict(workload_config(S), reference_arm="same workload, reference algorithm (fp32 oracle port) on the host "
                                                             "cores; each step is a bounded sample (one UNet forward)"),
           "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--ddim-steps", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--latency-mode", action="store_true",
                    help="pairs of GPUs share one clip: each evaluates one classifier-free-guidance branch per step "
                         "(one NCCL all-gather per step) and one of the two decodes; needs an even --gpus >= 2")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from tooncrafter_b200 import ops
    from tooncrafter_b200.sampler import DDIMSampler
    S = args.ddim_steps
    model = build_model(dev, rank, world)
    sampler = DDIMSampler(model)
    fs = torch.tensor([10], device=dev)
    from tooncrafter_b200.distributed import clip_seed, shard_clips
    pair_group, in_pair = None, 0
    if args.latency_mode:
        if world < 2:
            raise SystemExit("--latency-mode needs torchrun with an even number of GPUs >= 2")
        from tooncrafter_b200.distributed import latency_pairs
        pair_group, my_clip, in_pair = latency_pairs()     # one clip per GPU PAIR per step
        sampler.latency_group = pair_group
    else:
        my_clip = shard_clips(world, rank, world)[0]       # one clip per GPU per step (weak scaling)
    hi = host_inputs(seed=clip_seed(123, my_clip))          # per-clip seed: results independent of the world size
    di = to_device(hi, dev)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up (builds plans, captures graphs) ------------------------------------------------------------
    for _ in range(max(args.warmup, 1)):
        run_clip(model, sampler, di, S, fs)
    barrier()
    n0 = ops.launch_count()

    # ---- timed: device-resident inputs ----------------------------------------------------------------------
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        run_clip(model, sampler, di, S, fs)
    e1.record()
    barrier()
    ms_dev = max_over_ranks(e0.elapsed_time(e1))
    clk = clocks.stop() if rank == 0 else None

    # ---- timed: end to end through the public API with HOST buffers -----------------------------------------
    out_host = torch.empty(1, 3, T, 8 * H, 8 * W, dtype=torch.float16).pin_memory()
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        dj = to_device(hi, dev)                            # H2D of this step's inputs from pinned memory
        video = run_clip(model, sampler, dj, S, fs)
        out_host[:, :, :video.shape[2]].copy_(video, non_blocking=True)   # D2H of the step's result (latency mode:
                                                                          # pair-rank 1 holds the 14-frame variant)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks = load_peaks()
    unet = model.model.diffusion_model
    plan = unet._engine.plan_for(2, T, H, W, 77 + 16 * T)
    dec = model.first_stage_model._dec_engine
    # launches inside captured graphs are replayed, not re-issued through the C ABI: count them from the programs
    kernels_unet = getattr(plan.main, "kernels", None)
    roof = gemm_roofline(model, dev)
    n_gemm_clip = 0
    if kernels_unet is None:
        c0 = ops.launch_count()
        plan.main.run()
        torch.cuda.synchronize()
        kernels_unet = ops.launch_count() - c0
    kernels_dec = 0
    for key in ((16, H, W), (14, H, W)):
        pl = dec.plan_for(*key)
        c0 = ops.launch_count()
        pl.main.run()
        torch.cuda.synchronize()
        kernels_dec += ops.launch_count() - c0
    launches_per_clip = S * (kernels_unet + 2) + kernels_dec
    clips = args.steps * (world // 2 if args.latency_mode else world)
    fps_dev = 16.0 * clips / (ms_dev / 1e3)
    fps_e2e = 16.0 * clips / (ms_e2e / 1e3)
    out = {
        "metric": "frames/sec (320x512x16f, DDIM-%d)" % S, "value": fps_dev, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "strong" if args.latency_mode else "weak",
        "vs_baseline": fps_dev / (16.0 / 24.0) if S == 50 else None, "dtype": "f16",
        "data": "synthetic",
        "config": workload_config(S),
        "sec_per_clip": ms_dev / args.steps / 1e3,
        "mode": "latency (one clip per GPU pair: CFG branches split, all-gather per step)" if args.latency_mode
                else "throughput (one clip per GPU)",
        "tflops_per_gpu": clip_tflop(S) / (ms_dev / args.steps / 1e3) / (2 if args.latency_mode else 1),
        "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d_bytes(hi),
                "d2h_bytes_per_step": out_host.numel() * out_host.element_size(),
                "sec_per_clip": ms_e2e / args.steps / 1e3},
        "gpu_launches": launches_per_clip * args.steps,
        "clocks": clk,
        "roofline": {"bound": "tensor", "kernel": "tc_gemm_kernel (implicit-GEMM conv / linear, tcgen05)",
                     "achieved": roof["tflops"], "peak": peaks["tflops"], "unit": "TFLOP/s",
                     "frac": roof["tflops"] / peaks["tflops"], "traffic": gemm_traffic()[0],
                     "traffic_unit": "bytes of DRAM read+write per launch (ncu, average over the GEMM launches of one UNet forward)",
                     "traffic_source": gemm_traffic()[1], "peak_source": peaks["src"],
                     "launches_per_unet_forward": roof["launches"], "avg_launch_us": roof["avg_launch_us"],
                     "flops_per_launch": roof["flops_per_launch"],
                     "share_of_unet_forward": roof["gemm_ms_per_forward"]},
    }
    if not args.no_cpu_baseline:
        cores = cpu_threads()
        t_fwd = cpu_unet_forward_seconds(cores)
        out["cpu_baseline"] = {"value": 16.0 / (t_fwd * clip_tflop(S) / UNET_TF), "unit": "frames/s", "cores": cores,
                               "kind": "port",
                               "sample": f"one full-size UNet forward (B=1, fp32 oracle) = {t_fwd:.1f} s on {cores} "
                                         f"threads; sec/clip extrapolated by flops ({UNET_TF} of {clip_tflop(S):.0f} TFLOP)"}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
