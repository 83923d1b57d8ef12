"""GroupNorm(+SiLU) timings at the UNet's shapes, cold (L2 flushed) and warm (back to back)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent))
import bench_kernels as bk
from tooncrafter_b200 import ops
def warm(fn, iters=20):
    """GPU time per call, back to back inside one CUDA graph (no host launch gaps)."""
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for frames, fps, hw, C in [(32, 16, 40, 1280), (32, 1, 40, 1280), (32, 16, 160, 1280), (32, 1, 160, 1280), (32, 16, 640, 640), (32, 1, 640, 640),
                           (32, 16, 2560, 320), (32, 1, 2560, 320), (32, 16, 160, 2560), (16, 1, 320 * 512 // 4, 256)]:
    x = torch.randn(frames * hw, C, device="cuda").half(); y = torch.empty_like(x)
    g = torch.ones(C, device="cuda"); b = torch.zeros(C, device="cuda")
    fn = lambda: ops.groupnorm(x, y, g, b, frames=frames, frames_per_stat=fps, hw=hw, C=C, silu=True)
    tc, tw = bk.timeit(fn) * 1e3, warm(fn) * 1e3
    mb = x.numel() * 2 / 1e6
    print(f"gn frames={frames} fps={fps} hw={hw} C={C} ({mb:6.1f} MB): cold {tc:7.1f} us  warm {tw:7.1f} us  (3-pass floor @6.5TB/s {3*mb/6.5:5.1f} us)", flush=True)
# per-launch floor inside a graph: a tiny copy, and a chain of two tiny copies
a = torch.randn(64, 64, device="cuda").half(); b2 = torch.empty_like(a); c2 = torch.empty_like(a)
print(f"floor: 1 tiny copy {warm(lambda: ops.copy2d(a, b2, rows=64, cols=64, lds=64, ldd=64)) * 1e3:.2f} us", flush=True)
def two():
    ops.copy2d(a, b2, rows=64, cols=64, lds=64, ldd=64); ops.copy2d(b2, c2, rows=64, cols=64, lds=64, ldd=64)
print(f"floor: 2 chained tiny copies {warm(two) * 1e3:.2f} us", flush=True)
x = torch.randn(1280, 1280, device="cuda").half(); y = torch.empty_like(x)
print(f"copy 3.3 MB: {warm(lambda: ops.copy2d(x, y, rows=1280, cols=1280, lds=1280, ldd=1280)) * 1e3:.2f} us", flush=True)
