"""A/B of the attention kernels on the B200: third generation with its options (exponential-phase token, start stagger;
ATTN_AB_ALL=1 adds the second-generation kernel and the polynomial fractions) with
0..4 of every 8 exponential pairs on the FMA pipe (TC_ATTN_POLY).  For each variant: parity against torch fp32
(max error, fraction outside rtol 1e-3 / atol 1e-4) on small / ragged shapes incl. data with large score ranges (forces
the lazy-rescale path), then CUDA-event timings at the UNet level-0 and VAE fusion shapes.

    python scripts/attn_ab.py [--quick]
"""
import argparse
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from tooncrafter_b200 import ops  # noqa: E402

DEV = "cuda"


def ref_attn(q, k, v, heads, kv_div=1):
    B, Lq, _ = q.shape
    k = k.repeat_interleave(kv_div, 0)[:B]
    v = v.repeat_interleave(kv_div, 0)[:B]
    sp = lambda t: t.float().reshape(t.shape[0], t.shape[1], heads, 64).transpose(1, 2)
    s = (sp(q) @ sp(k).transpose(-1, -2)) * 64 ** -0.5
    return (s.softmax(-1) @ sp(v)).transpose(1, 2).reshape(B, Lq, heads * 64)


def run(q, k, v, heads, kv_div=1):
    B, Lq, C = q.shape
    out = torch.zeros_like(q)
    ops.attention(q, [dict(k=k, v=v, ldk=C, ldv=C, Lk=k.shape[1], kv_div=kv_div)], out, q_batches=B, Lq=Lq, heads=heads,
                  scale=64 ** -0.5, ldq=C, ldo=C)
    return out


def variants():
    if os.environ.get("ATTN_AB_ALL") == "1":
        yield "v2", dict(TC_ATTN_IMPL="v2")
        for p in (1, 2, 3, 4):
            yield f"v3 poly={p}/8", dict(TC_ATTN_IMPL="v3", TC_ATTN_POLY=str(p))
    yield "v3 p0", dict(TC_ATTN_IMPL="v3", TC_ATTN_POLY="0")
    yield "v3 single", dict(TC_ATTN_IMPL="v3", TC_ATTN_POLY="0", TC_ATTN_SINGLE="1")
    if os.environ.get("ATTN_AB_ALL") == "1":
        yield "v3 p0 token", dict(TC_ATTN_IMPL="v3", TC_ATTN_POLY="0", TC_ATTN_PP="1")
        yield "v3 p0 stagger", dict(TC_ATTN_IMPL="v3", TC_ATTN_POLY="0", TC_ATTN_STAGGER="1200")


def set_env(e):
    for k in ("TC_ATTN_IMPL", "TC_ATTN_POLY", "TC_ATTN_PP", "TC_ATTN_STAGGER", "TC_ATTN_SINGLE"):
        os.environ.pop(k, None)
    os.environ.update(e)


def parity():
    g = torch.Generator().manual_seed(0)
    cases = []
    for (B, Lq, Lk, heads, kv_div, spread) in [(2, 300, 300, 2, 1, 1.0), (3, 128, 128, 1, 1, 1.0), (2, 2560, 2560, 5, 1, 1.0),
                                               (4, 40, 40, 3, 1, 1.0), (1, 640, 640, 10, 1, 1.0), (2, 160, 160, 20, 1, 1.0),
                                               (2, 500, 1000, 2, 1, 6.0), (4, 1024, 2048, 8, 4, 1.0), (2, 257, 129, 1, 1, 12.0),
                                               (1, 256, 4096, 2, 1, 25.0)]:
        C = heads * 64
        q = (torch.randn(B, Lq, C, generator=g) * spread).half().to(DEV)
        kb = (B + kv_div - 1) // kv_div
        k = torch.randn(kb, Lk, C, generator=g).half().to(DEV)
        # rows whose maximum keeps growing along the key axis: exercises the O rescale
        k = k * torch.linspace(0.5, 2.0, Lk, device=DEV).half()[None, :, None]
        v = torch.randn(kb, Lk, C, generator=g).half().to(DEV)
        cases.append((f"B{B} Lq{Lq} Lk{Lk} h{heads} div{kv_div} x{spread}", q, k, v, heads, kv_div))
    ok = True
    for name, env in variants():
        set_env(env)
        worst, worst_v = 0.0, 0.0
        for cname, q, k, v, heads, kv_div in cases:
            out = run(q, k, v, heads, kv_div).float()
            torch.cuda.synchronize()
            ref = ref_attn(q, k, v, heads, kv_div)
            d = (out - ref).abs()
            viol = (d > 1e-4 + 1e-3 * ref.abs()).float().mean().item()
            bad = not torch.isfinite(out).all().item() or d.max().item() > 3e-3 * ref.abs().max().item() + 1e-3
            if bad:
                ok = False
                print(f"  !! {name}: {cname}: max err {d.max().item():.3e} (ref max {ref.abs().max().item():.2f}) viol {viol:.2e}")
            worst, worst_v = max(worst, d.max().item()), max(worst_v, viol)
        print(f"parity {name:14s}: worst max err {worst:.3e}, worst fraction outside rtol 1e-3/atol 1e-4 {worst_v:.2e}", flush=True)
    return ok


def timing(quick):
    shapes = [("unet L0 self  (32 x 5 heads, L 2560)", 32, 2560, 2560, 5, 1),
              ("unet L1 self  (32 x 10 heads, L 640)", 32, 640, 640, 10, 1),
              ("vae fusion    (16 x 8 heads, 10240 x 20480)", 16, 10240, 20480, 8, 16)]
    if quick:
        shapes = shapes[:2]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for sname, B, Lq, Lk, heads, kv_div in shapes:
        C = heads * 64
        q = torch.randn(B, Lq, C, device=DEV).half()
        kb = (B + kv_div - 1) // kv_div
        k = torch.randn(kb, Lk, C, device=DEV).half()
        v = torch.randn(kb, Lk, C, device=DEV).half()
        fl = 4.0 * B * heads * Lq * Lk * 64
        for name, env in variants():
            set_env(env)
            for _ in range(3):
                run(q, k, v, heads, kv_div)
            ts = []
            for _ in range(7):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                run(q, k, v, heads, kv_div)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            ms = ts[len(ts) // 2]
            print(f"time {sname:46s} {name:14s}: {ms * 1e3:9.1f} us  {fl / ms / 1e9:7.1f} TF/s", flush=True)


def cross_timing():
    """text (77 keys, shared by the 16 frames of a sample) + image (16 keys per frame) cross attention at the UNet levels"""
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for name, N, L, heads in (("unet L0 cross", 32, 2560, 5), ("unet L1 cross", 32, 640, 10), ("unet L2 cross", 32, 160, 20)):
        C = heads * 64
        q = torch.randn(N, L, C, device=DEV).half()
        kt, vt = torch.randn(2, 77, C, device=DEV).half(), torch.randn(2, 77, C, device=DEV).half()
        ki, vi = torch.randn(N, 16, C, device=DEV).half(), torch.randn(N, 16, C, device=DEV).half()
        out = torch.zeros_like(q)
        segs = [dict(k=kt, v=vt, ldk=C, ldv=C, Lk=77, kv_div=16), dict(k=ki, v=vi, ldk=C, ldv=C, Lk=16)]
        fn = lambda: ops.attention(q, segs, out, q_batches=N, Lq=L, heads=heads, scale=64 ** -0.5, ldq=C, ldo=C)
        fl = 4.0 * N * heads * L * 93 * 64
        for vname, env in (("v2", dict(TC_ATTN_IMPL="v2")), ("resident kv", dict(TC_ATTN_IMPL="v3"))):
            set_env(env)
            for _ in range(3):
                fn()
            ts = []
            for _ in range(7):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            print(f"time {name:46s} {vname:14s}: {ts[3] * 1e3:9.1f} us  {fl / ts[3] / 1e9:7.1f} TF/s  "
                  f"{(4.0 * N * L * C) / ts[3] / 1e6:7.0f} GB/s (q + out)", flush=True)


def wide():
    """fused single-head attention, head dim 512 (VAE mid block): parity + time against the unfused GEMM / softmax path"""
    for N, L, D in ((2, 300, 512), (2, 256, 256), (16, 2560, 512)):
        qkv = torch.randn(N, L, 3 * D, device=DEV).half()
        out = torch.zeros(N, L, D, dtype=torch.float16, device=DEV)
        fn = lambda: ops.attention_wide(qkv, out, batches=N, L=L, D=D, scale=D ** -0.5, ld=3 * D, ldo=D, k_offset=D, v_offset=2 * D)
        fn()
        torch.cuda.synchronize()
        q, k, v = (qkv[..., i * D:(i + 1) * D].float() for i in range(3))
        ref = ((q @ k.transpose(-1, -2)) * D ** -0.5).softmax(-1) @ v
        d = (out.float() - ref).abs()
        viol = (d > 1e-4 + 1e-3 * ref.abs()).float().mean().item()
        ts = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        fl = 4.0 * N * L * L * D
        print(f"wide N={N} L={L} D={D}: max err {d.max().item():.3e} viol {viol:.2e} finite {torch.isfinite(out).all().item()} "
              f"{ts[3] * 1e3:8.1f} us {fl / ts[3] / 1e9:7.1f} TF/s", flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--cross-only", action="store_true")
    ap.add_argument("--wide-only", action="store_true")
    a = ap.parse_args()
    ap2 = a
    if getattr(a, "cross_only", False):
        cross_timing()
        sys.exit(0)
    if a.wide_only:
        wide()
        sys.exit(0)
    good = parity()
    timing(a.quick)
    print("PARITY_OK" if good else "PARITY_FAILED")
