"""Per-op CUDA-event timing of one eager UNet forward (B = 2, warm L2, back-to-back launches): the in-situ
complement of the cold-cache ncu launch list.  Aggregates by op shape like scripts/join_launches.py.

    python scripts/profile_unet_events.py [--out gpurun_out/unet_events.txt]
"""
import argparse, collections, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "scripts")); sys.path.insert(0, str(ROOT / "tests"))
from bench_unet import build_full_unet
from join_launches import describe
from tooncrafter_b200.engine import UNetEngine

def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--out", default="gpurun_out/unet_events.txt"); ap.add_argument("--B", type=int, default=2)
    a = ap.parse_args()
    m = build_full_unet()
    eng = UNetEngine(m, use_graph=False)
    B = a.B
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, 8, 16, 40, 64, generator=g).cuda(); ctx = torch.randn(B, 333, 1024, generator=g).cuda()
    t = torch.full((B,), 500, device="cuda"); fs = torch.full((B,), 10, device="cuda")
    for _ in range(3):
        eng.forward(x, t, ctx, fs)
    torch.cuda.synchronize()
    plan = eng.plan_for(B, 16, 40, 64, 333)
    reps = 3
    acc = collections.defaultdict(lambda: [0.0, 0, 0.0, 0.0])
    for _ in range(reps):
        evs = []
        for fn, args, kw in plan.main.calls:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(*args, **kw); e1.record()
            evs.append((e0, e1, describe(fn, args, kw)))
        torch.cuda.synchronize()
        for e0, e1, d in evs:
            key = d["op"]
            if d["op"] == "gemm": key = f"gemm M={d['M']} N={d['N']} K={d['K']}"
            elif d["op"] == "groupnorm": key = f"groupnorm C={d['C']} elems={d['elems']} fps={d['fps']}"
            elif d["op"] == "attention": key = f"attention Lq={d['Lq']} Lk={d['Lk']} heads={d['heads']}"
            elif d["op"] == "layernorm": key = f"rowstats C={d['C']} elems={d['elems']}"
            elif d["op"] == "temporal_attention": key = f"temporal_attention elems={d['elems']}"
            v = acc[key]; v[0] += e0.elapsed_time(e1) / reps; v[1] += 1; v[2] += d["flops"] / reps; v[3] += d["bytes"] / reps
    tot = sum(v[0] for v in acc.values())
    lines = [f"total {tot:.3f} ms per eager forward (B={B}), CUDA events per op, warm L2"]
    cat = collections.defaultdict(float)
    for k, (ms, n, fl, by) in sorted(acc.items(), key=lambda kv: -kv[1][0]):
        lines.append(f"{ms:8.3f} ms {100*ms/tot:5.1f}% n={n//reps:3d} {fl/ms/1e9 if ms else 0:8.1f} TF/s {by/ms/1e6 if ms else 0:8.1f} GB/s  {k}")
        cat[k.split()[0]] += ms
    lines.insert(1, "by kind: " + ", ".join(f"{k} {v:.2f}" for k, v in sorted(cat.items(), key=lambda kv: -kv[1])))
    Path(a.out).parent.mkdir(parents=True, exist_ok=True); Path(a.out).write_text("\n".join(lines) + "\n")
    print("\n".join(lines[:45]))

if __name__ == "__main__":
    main()
