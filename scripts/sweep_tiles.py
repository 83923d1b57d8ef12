"""Sweep block_n x {single, pair} for a few GEMM shapes (pair forced through TC_GEMM_PAIR, so run once per setting)."""
import os, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent))
import bench_kernels as bk
from tooncrafter_b200 import ops
def conv(N, H, W, Cin, Cout, bn):
    x = torch.randn(N, H, W, Cin, device="cuda").half(); w = (torch.randn(Cout, 9 * Cin, device="cuda") * (9 * Cin) ** -0.5).half()
    b = torch.zeros(Cout, device="cuda"); out = torch.empty(N, H, W, Cout, device="cuda", dtype=torch.float16)
    fn = lambda: ops.conv_gemm(x, (N, H, W, Cin), (H * W * Cin, W * Cin, Cin), w, ops.TAPS_3x3, out, (N, H, W), Cout, bias=b, block_n=bn)
    t = bk.timeit(fn); fl = 2 * N * H * W * Cout * 9 * Cin
    return f"{t*1e3:7.1f}us/{fl/t/1e9:6.0f}TF"
print("pair env:", os.environ.get("TC_GEMM_PAIR"))
for shp in [(32, 40, 64, 320, 320), (32, 20, 32, 640, 640), (32, 10, 16, 1280, 1280), (32, 5, 8, 1280, 1280), (32, 10, 16, 2560, 1280), (32, 20, 32, 1280, 640)]:
    line = f"conv {shp}: "
    for bn in (0, 64, 128, 160, 256):
        if bn and shp[4] % bn: continue
        line += f" bn{bn}:{conv(*shp, bn)}"
    print(line, flush=True)
