"""Split-K A/B on the B200: the UNet's few-tile / long-K GEMMs (levels 1-3 of the 320x512x16 clip, both guidance branches)
timed with the k-slice cap forced to 1..4 (tc_debug_set_gemm_mode bits 8..11) and with the heuristic, next to cuDNN /
cuBLAS fp16 on the same operands.  L2 is flushed between timed launches.

    python scripts/ksplit_ab.py
"""
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from tooncrafter_b200 import _lib, ops  # noqa: E402

DEV = "cuda"
lib = _lib.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def timeit(fn, reps=9):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def conv_case(N, H, W, Cin, Cout, res):
    x = torch.randn(N, H, W, Cin, device=DEV).half()
    w = (torch.randn(Cout, Cin, 3, 3, device=DEV) * (9 * Cin) ** -0.5)
    wp = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous().half()
    bias = torch.randn(Cout, device=DEV)
    r = torch.randn(N, H, W, Cout, device=DEV).half() if res else None
    out = torch.zeros(N, H, W, Cout, dtype=torch.float16, device=DEV)
    ours = lambda: ops.conv_gemm(x, (N, H, W, Cin), (H * W * Cin, W * Cin, Cin), wp, ops.TAPS_3x3, out, (N, H, W), Cout,
                                 bias=bias, res=r)
    xc = x.permute(0, 3, 1, 2)                                   # channels-last view
    wc = w.half().contiguous(memory_format=torch.channels_last)
    libfn = lambda: F.conv2d(xc, wc, bias.half(), padding=1)
    return ours, libfn, 2.0 * N * H * W * Cout * 9 * Cin


def tconv_case(B, T, HW, C):
    x = torch.randn(B, T, HW, C, device=DEV).half()
    wp = (torch.randn(C, 3 * C, device=DEV) * (3 * C) ** -0.5).half()
    bias = torch.randn(C, device=DEV)
    r = torch.randn(B, T, HW, C, device=DEV).half()
    out = torch.zeros_like(x)
    ours = lambda: ops.conv_gemm(x, (B, T, HW, C), (T * HW * C, HW * C, C), wp, ops.TAPS_T3, out, (B, T, HW), C, bias=bias, res=r)
    return ours, None, 2.0 * B * T * HW * C * 3 * C


def lin_case(rows, K, N, res):
    x = torch.randn(rows, K, device=DEV).half()
    w = (torch.randn(N, K, device=DEV) * K ** -0.5).half()
    bias = torch.randn(N, device=DEV)
    r = torch.randn(rows, N, device=DEV).half() if res else None
    out = torch.zeros(rows, N, dtype=torch.float16, device=DEV)
    ours = lambda: ops.linear(x, w, out, rows=rows, K=K, n_cols=N, bias=bias, res=r)
    bh = bias.half()
    libfn = (lambda: torch.addmm(r, x, w.t())) if res else (lambda: F.linear(x, w, bh))
    return ours, libfn, 2.0 * rows * K * N


CASES = [
    ("L1 conv3x3 640->640   (32x20x32)", lambda: conv_case(32, 20, 32, 640, 640, False)),
    ("L1 conv3x3 1280->640  (32x20x32)", lambda: conv_case(32, 20, 32, 1280, 640, False)),
    ("L1 conv3x3 1920->640  (32x20x32)", lambda: conv_case(32, 20, 32, 1920, 640, False)),
    ("L2 conv3x3 640->1280  (32x10x16)", lambda: conv_case(32, 10, 16, 640, 1280, False)),
    ("L2 conv3x3 1280->1280 (32x10x16)", lambda: conv_case(32, 10, 16, 1280, 1280, True)),
    ("L2 conv3x3 2560->1280 (32x10x16)", lambda: conv_case(32, 10, 16, 2560, 1280, False)),
    ("L3 conv3x3 1280->1280 (32x5x8)", lambda: conv_case(32, 5, 8, 1280, 1280, True)),
    ("L3 conv3x3 2560->1280 (32x5x8)", lambda: conv_case(32, 5, 8, 2560, 1280, False)),
    ("L2 temporal conv 1280 (2x16x160)", lambda: tconv_case(2, 16, 160, 1280)),
    ("L3 temporal conv 1280 (2x16x40)", lambda: tconv_case(2, 16, 40, 1280)),
    ("L1 temporal conv 640  (2x16x640)", lambda: tconv_case(2, 16, 640, 640)),
    ("L2 linear 1280->1280 +res (5120)", lambda: lin_case(5120, 1280, 1280, True)),
    ("L2 linear 5120->1280 +res (5120)", lambda: lin_case(5120, 5120, 1280, True)),
    ("L3 linear 1280->1280 +res (1280)", lambda: lin_case(1280, 1280, 1280, True)),
    ("L3 linear 5120->1280 +res (1280)", lambda: lin_case(1280, 5120, 1280, True)),
    ("L1 linear 2560->640 +res (20480)", lambda: lin_case(20480, 2560, 640, True)),
    ("L1 linear 640->640 +res (20480)", lambda: lin_case(20480, 640, 640, True)),
]

if __name__ == "__main__":
    torch.manual_seed(0)
    print(f"{'case':36s} {'cap1':>22s} {'cap2':>22s} {'cap3':>22s} {'cap4':>22s} {'auto':>22s} {'library':>9s}")
    for name, mk in CASES:
        ours, libfn, fl = mk()
        cols = []
        for cap in (1, 2, 3, 4, 0):
            _lib.check(lib.tc_debug_set_gemm_mode(cap << 8))
            us = timeit(ours)
            c = ops.last_gemm_config()
            cols.append(f"{us:7.1f}us bn{c['block_n']}{'p' if c['pair'] else 's'}k{c['ksplit']}")
        _lib.check(lib.tc_debug_set_gemm_mode(0))
        lt = f"{timeit(libfn):7.1f}us" if libfn else "        -"
        print(f"{name:36s} " + " ".join(f"{c:>22s}" for c in cols) + f" {lt}  ({fl / 1e9:.1f} GF)", flush=True)
