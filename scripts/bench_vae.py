"""Time the full-size dual-reference decoder (BASELINE config #4: 320x512x16 latents, 1 GPU).

    python scripts/bench_vae.py [--T 16] [--iters 5] [--no-graph]
"""
import argparse, json, sys, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from tiny_config import FULL_DDCONFIG
from tooncrafter_b200 import diffusion, synthetic
from tooncrafter_b200.vae_engine import DecoderEngine

DEC_TF = {16: 37.875, 14: 33.148}

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--T", type=int, default=16)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--out", default="gpurun_out/vae_bench.json")
    a = ap.parse_args()
    dev = torch.device("cuda")
    with torch.device("meta"):
        sk = diffusion.AutoencoderKL_Dualref(ddconfig=FULL_DDCONFIG, embed_dim=4)
    ae = sk.to_empty(device=dev)
    with torch.no_grad():
        for k, p in ae.named_parameters():
            p.copy_(synthetic.synthetic_tensor("first_stage_model." + k, tuple(p.shape), 0).to(dev))
    ae.eval()
    eng = DecoderEngine(ae.decoder, use_graph=not a.no_graph)
    ref = [r.to(dev) for r in synthetic.synthetic_ref_context(128, [1, 2, 4, 4], 320, 512)]
    z = torch.randn(a.T, 4, 40, 64, device=dev) * 3
    y = eng.decode(z, ref)
    torch.cuda.synchronize()
    plan = eng.plan_for(a.T, 40, 64)
    print(f"first decode ok: launches {len(plan.main)}, ctx {len(plan.ctx)}, arena high water "
          f"{plan.arena.high_water / 2**30:.2f} GiB, finite={bool(torch.isfinite(y).all())}, |y|max={y.float().abs().max().item():.2f}", flush=True)
    for _ in range(a.warmup):
        eng.decode(z, ref)
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.decode(z, ref); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    if ts:
        ts.sort(); med = ts[len(ts) // 2]
        res = dict(T=a.T, ms=med, tflops=DEC_TF.get(a.T, 0) / med * 1e3, launches=len(plan.main))
        print(json.dumps(res), flush=True)
        Path(a.out).parent.mkdir(parents=True, exist_ok=True); Path(a.out).write_text(json.dumps(res))

if __name__ == "__main__":
    main()
