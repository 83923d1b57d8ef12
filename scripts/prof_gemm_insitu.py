"""The K = 320 GEMM flavours exactly as the UNet engine launches them at level 0 (B = 2: 81920 rows), for ncu:
   qkv    : LN-folded Linear 320 -> 960 consuming producer-side row statistics
   geglu  : LN-folded GEGLU Linear 320 -> 2560 (output 1280) consuming producer-side row statistics
   proj   : Linear 320 -> 320 + bias + residual (tensor-core identity blocks) producing row statistics for the next LN
   ff2    : Linear 1280 -> 320 + bias + residual
Usage: python scripts/prof_gemm_insitu.py [case ...]   (prints CUDA-event times; under ncu use -k regex:tc_gemm -s 3 -c 1)
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from tooncrafter_b200 import engine, ops  # noqa: E402

DEV = "cuda"
ROWS = 81920


def setup():
    g = torch.Generator().manual_seed(0)
    r = lambda *s, k=1.0: (torch.randn(*s, generator=g) * k).to(DEV)
    C = 320
    x = r(ROWS, C).half()
    ln = torch.nn.LayerNorm(C).to(DEV)
    slots = 2
    part = torch.zeros(ROWS, slots, 2, device=DEV)
    w0, b0, res = r(C, C, k=C ** -0.5).half(), r(C).float(), r(ROWS, C).half()
    mid = torch.zeros(ROWS, C, dtype=torch.float16, device=DEV)
    proj = lambda: ops.linear(x, w0, mid, rows=ROWS, K=C, n_cols=C, bias=b0, res=res, block_n=160, row_stats=part,
                              row_stats_slots=slots)
    proj()
    fq = engine.fold_layernorm(r(960, C, k=C ** -0.5), None, ln, torch.device(DEV))
    oq = torch.zeros(ROWS, 960, dtype=torch.float16, device=DEV)
    qkv = lambda: ops.linear(mid, fq.w, oq, rows=ROWS, K=C, n_cols=960, bias=fq.c, ln_stats=part, ln_u=fq.u, ln_nslots=slots,
                             ln_eps=1e-5)
    fg = engine.fold_layernorm(r(2560, C, k=C ** -0.5), r(2560), ln, torch.device(DEV), perm=engine.geglu_perm(2560).to(DEV))
    og = torch.zeros(ROWS, 1280, dtype=torch.float16, device=DEV)
    geglu = lambda: ops.linear(mid, fg.w, og, rows=ROWS, K=C, n_cols=2560, bias=fg.c, ln_stats=part, ln_u=fg.u,
                               ln_nslots=slots, ln_eps=1e-5, geglu=True, block_n=256)
    w2, b2 = r(C, 1280, k=1280 ** -0.5).half(), r(C).float()
    o2 = torch.zeros(ROWS, C, dtype=torch.float16, device=DEV)
    ff2 = lambda: ops.linear(og, w2, o2, rows=ROWS, K=1280, n_cols=C, bias=b2, res=mid)
    return dict(proj=(proj, 2.0 * ROWS * C * C), qkv=(qkv, 2.0 * ROWS * C * 960), geglu=(geglu, 2.0 * ROWS * C * 2560),
                ff2=(ff2, 2.0 * ROWS * 1280 * C))


if __name__ == "__main__":
    cases = setup()
    modes = [0]
    args = sys.argv[1:]
    if args and args[0].startswith("--modes="):
        # tc_debug_set_gemm_mode bits: 1 = no global / TMA stores, 2 = no epilogue body (mainloop + handshakes only),
        # 8 = no proxy fence, 16 = no staging stores
        modes = [int(m) for m in args.pop(0).split("=")[1].split(",")]
    names = args or list(cases)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    from tooncrafter_b200 import _lib
    lib = _lib.load()
    for n in [(nm, md) for nm in names for md in modes]:
        n, mode = n
        _lib.check(lib.tc_debug_set_gemm_mode(mode))
        torch.cuda.synchronize()
        fn, fl = cases[n]
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
        print(f"{n:6s} mode {mode:2d} {ts[3] * 1e3:8.1f} us  {fl / ts[3] / 1e9:7.1f} TF/s", flush=True)
    _lib.check(lib.tc_debug_set_gemm_mode(0))
