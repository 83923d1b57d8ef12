"""Time the full-size UNet forward (inference_512_v1.0, B = 2 = cond + uncond, T = 16, latent 40x64) on one B200.

    python scripts/bench_unet.py [--iters 10] [--no-graph] [--B 2]
"""
import argparse
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from tiny_config import FULL_UNET  # noqa: E402
from tooncrafter_b200 import modules, ops, synthetic  # noqa: E402
from tooncrafter_b200.engine import UNetEngine  # noqa: E402

UNET_TFLOP_PER_SAMPLE = 12.603   # SURVEY §8d (matmul/conv flops, 2 per MAC)


def build_full_unet(dev="cuda", seed=0):
    with torch.device("meta"):
        sk = modules.UNetModel(**FULL_UNET)
    m = sk.to_empty(device=dev)
    with torch.no_grad():
        for k, p in m.named_parameters():
            p.copy_(synthetic.synthetic_tensor("model.diffusion_model." + k, tuple(p.shape), seed).to(dev))
    return m.eval()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--B", type=int, default=2)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--out", default="gpurun_out/unet_bench.json")
    args = ap.parse_args()
    t0 = time.time()
    m = build_full_unet()
    print(f"weights ready in {time.time() - t0:.1f}s", flush=True)
    eng = UNetEngine(m, use_graph=not args.no_graph)
    B = args.B
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, 8, 16, 40, 64, generator=g).cuda()
    ctx = torch.randn(B, 77 + 256, 1024, generator=g).cuda()
    t = torch.full((B,), 500, device="cuda")
    fs = torch.full((B,), 10, device="cuda")
    n0 = ops.launch_count()
    y = eng.forward(x, t, ctx, fs)
    torch.cuda.synchronize()
    plan = eng.plan_for(B, 16, 40, 64, ctx.shape[1])
    print(f"first forward done; launches per forward {len(plan.main)}, ctx launches {len(plan.ctx)}, "
          f"arena high water {plan.arena.high_water / 2**20:.0f} MiB, finite={bool(torch.isfinite(y).all())}, "
          f"|y|max={y.float().abs().max().item():.3f}", flush=True)
    for _ in range(args.warmup):
        eng.forward(x, t, ctx, fs)
    torch.cuda.synchronize()
    ts = []
    for _ in range(args.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.forward(x, t, ctx, fs)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    if not ts:
        return
    ts.sort()
    med = ts[len(ts) // 2]
    res = dict(B=B, ms_per_forward=med, ms_min=ts[0], ms_max=ts[-1], tflops=UNET_TFLOP_PER_SAMPLE * B / med * 1e3,
               launches=len(plan.main), graph=not args.no_graph)
    print(json.dumps(res), flush=True)
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text(json.dumps(res))


if __name__ == "__main__":
    main()
