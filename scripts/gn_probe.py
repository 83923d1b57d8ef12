"""ncu target: GroupNorm at three UNet sizes (3 launches each) + plain torch copy / reduction of the same tensors."""
import sys, torch
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from tooncrafter_b200 import ops
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def t(fn, n=7):
    for _ in range(2): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]
for frames, fps, hw, C in [(32, 16, 160, 1280), (32, 16, 640, 640), (32, 16, 2560, 320)]:
    x = torch.randn(frames * hw, C, device="cuda").half(); y = torch.empty_like(x)
    g = torch.ones(C, device="cuda"); b = torch.zeros(C, device="cuda")
    gn = lambda: ops.groupnorm(x, y, g, b, frames=frames, frames_per_stat=fps, hw=hw, C=C, silu=True)
    print(f"{x.numel() * 2 / 1e6:6.1f} MB: groupnorm {t(gn):6.1f} us | torch copy_ {t(lambda: y.copy_(x)):6.1f} us | "
          f"torch sum {t(lambda: x.sum(dtype=torch.float32)):6.1f} us | torch silu {t(lambda: torch.nn.functional.silu(x)):6.1f} us", flush=True)
