"""A/B timing of a few long-K convolutions under the TC_GEMM_* environment toggles (L2 flushed between iterations)."""
import os, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent))
import bench_kernels as bk
from tooncrafter_b200 import ops
def conv(N, H, W, Cin, Cout, res):
    x = torch.randn(N, H, W, Cin, device="cuda").half(); w = (torch.randn(Cout, 9 * Cin, device="cuda") * (9 * Cin) ** -0.5).half()
    b = torch.zeros(Cout, device="cuda"); r = torch.randn(N, H, W, Cout, device="cuda").half() if res else None
    out = torch.empty(N, H, W, Cout, device="cuda", dtype=torch.float16)
    fn = lambda: ops.conv_gemm(x, (N, H, W, Cin), (H * W * Cin, W * Cin, Cin), w, ops.TAPS_3x3, out, (N, H, W), Cout, bias=b, res=r)
    t = bk.timeit(fn); fl = 2 * N * H * W * Cout * 9 * Cin
    print(f"conv3x3 {N}x{H}x{W} {Cin}->{Cout} res={res}: {t*1e3:8.1f} us  {fl/t/1e9:7.1f} TF/s", flush=True)
print({k: v for k, v in os.environ.items() if k.startswith("TC_GEMM")})
for a in [(32, 10, 16, 1280, 1280, False), (32, 10, 16, 1280, 1280, True), (32, 20, 32, 640, 640, False), (32, 20, 32, 640, 640, True),
          (32, 40, 64, 320, 320, False), (32, 40, 64, 320, 320, True), (32, 10, 16, 2560, 1280, False), (32, 5, 8, 1280, 1280, True)]:
    conv(*a)
