"""Micro-benchmarks of the hot kernels at the UNet/VAE shapes (CUDA-event timed, L2 flushed between iterations).

Usage: python scripts/bench_kernels.py [--out gpurun_out/kernels.json]
"""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from tooncrafter_b200 import ops  # noqa: E402

DEV = "cuda"
_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)
    _flush.zero_()


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush_l2()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def bench_conv(N, H, W, Cin, Cout, taps="3x3"):
    x = torch.randn(N, H, W, Cin, device=DEV).half()
    nt = 9 if taps == "3x3" else 1
    w = (torch.randn(Cout, nt * Cin, device=DEV) * (nt * Cin) ** -0.5).half()
    b = torch.zeros(Cout, device=DEV)
    out = torch.empty(N, H, W, Cout, device=DEV, dtype=torch.float16)
    tp = ops.TAPS_3x3 if taps == "3x3" else ops.TAPS_1x1
    fn = lambda: ops.conv_gemm(x, (N, H, W, Cin), (H * W * Cin, W * Cin, Cin), w, tp, out, (N, H, W), Cout, bias=b)
    ms = timeit(fn)
    fl = 2.0 * N * H * W * Cout * nt * Cin
    return dict(kind=f"conv{taps}", shape=[N, H, W, Cin, Cout], ms=ms, tflops=fl / ms / 1e9)


def bench_linear(rows, K, Nc, geglu=False):
    x = torch.randn(rows, K, device=DEV).half()
    w = (torch.randn(Nc, K, device=DEV) * K ** -0.5).half()
    b = torch.zeros(Nc, device=DEV)
    out = torch.empty(rows, Nc // 2 if geglu else Nc, device=DEV, dtype=torch.float16)
    fn = lambda: ops.linear(x, w, out, rows=rows, K=K, n_cols=Nc, bias=b, geglu=geglu, block_n=256 if geglu else 0)
    ms = timeit(fn)
    return dict(kind="geglu" if geglu else "linear", shape=[rows, K, Nc], ms=ms, tflops=2.0 * rows * K * Nc / ms / 1e9)


def bench_attn(B, L, heads, Lk=None):
    Lk = Lk or L
    C = heads * 64
    q = torch.randn(B, L, C, device=DEV).half()
    k = torch.randn(B, Lk, C, device=DEV).half()
    v = torch.randn(B, Lk, C, device=DEV).half()
    out = torch.empty_like(q)
    fn = lambda: ops.attention(q, [dict(k=k, v=v, ldk=C, ldv=C, Lk=Lk)], out, q_batches=B, Lq=L, heads=heads,
                               scale=0.125, ldq=C, ldo=C)
    ms = timeit(fn)
    return dict(kind="attention", shape=[B, L, Lk, heads], ms=ms, tflops=4.0 * B * heads * L * Lk * 64 / ms / 1e9)


def bench_attn_wide(N, L, D):
    """fused single-head attention of the VAE mid block (head dim D)"""
    qkv = torch.randn(N, L, 3 * D, device=DEV).half()
    out = torch.empty(N, L, D, device=DEV, dtype=torch.float16)
    fn = lambda: ops.attention_wide(qkv, out, batches=N, L=L, D=D, scale=D ** -0.5, ld=3 * D, ldo=D, k_offset=D, v_offset=2 * D)
    ms = timeit(fn)
    return dict(kind="attention_wide", shape=[N, L, D], ms=ms, tflops=4.0 * N * L * L * D / ms / 1e9)


def bench_gn(frames, fps, hw, C):
    x = torch.randn(frames, hw, C, device=DEV).half()
    y = torch.empty_like(x)
    g = torch.ones(C, device=DEV)
    b = torch.zeros(C, device=DEV)
    fn = lambda: ops.groupnorm(x, y, g, b, frames=frames, frames_per_stat=fps, hw=hw, C=C, silu=True)
    ms = timeit(fn)
    by = 2.0 * x.numel() * 2  # algorithmic: read once + write once
    return dict(kind="groupnorm_silu", shape=[frames, fps, hw, C], ms=ms, gbps=by / ms / 1e6)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/kernels.json")
    args = ap.parse_args()
    res = []
    F = 32  # B=2 (cond+uncond) x 16 frames
    for a in [(F, 40, 64, 320, 320), (F, 20, 32, 640, 640), (F, 10, 16, 1280, 1280), (F, 5, 8, 1280, 1280),
              (16, 80, 128, 512, 512), (16, 320, 512, 128, 128)]:
        res.append(bench_conv(*a))
        print(res[-1], flush=True)
    for a in [(F * 2560, 320, 320), (F * 640, 640, 640), (F * 160, 1280, 1280), (F * 2560, 1280, 320)]:
        res.append(bench_linear(*a))
        print(res[-1], flush=True)
    for a in [(F * 2560, 320, 2560), (F * 640, 640, 5120), (F * 160, 1280, 10240)]:
        res.append(bench_linear(*a, geglu=True))
        print(res[-1], flush=True)
    for a in [(F, 2560, 5), (F, 640, 10), (F, 160, 20), (16, 10240, 8, 20480)]:
        res.append(bench_attn(*a))
        print(res[-1], flush=True)
    for a in [(F, 1, 2560, 320), (F, 16, 2560, 320), (16, 1, 163840, 128)]:
        res.append(bench_gn(*a))
        print(res[-1], flush=True)
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
