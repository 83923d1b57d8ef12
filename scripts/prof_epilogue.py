"""Where does a skinny GEMM spend its time?  Times the same launches with (0) the normal epilogue, (1) global stores
removed, (2) the epilogue body removed (mainloop + handshakes only)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent))
import bench_kernels as bk
from tooncrafter_b200 import _lib, ops
lib = _lib.load()
MODES = [int(m) for m in sys.argv[1].split(',')] if len(sys.argv) > 1 else [0, 1, 2]
def run(rows, K, N, res):
    x = torch.randn(rows, K, device="cuda").half(); w = (torch.randn(N, K, device="cuda") * K ** -0.5).half()
    b = torch.zeros(N, device="cuda"); r = torch.randn(rows, N, device="cuda").half() if res else None
    out = torch.empty(rows, N, device="cuda", dtype=torch.float16)
    fn = lambda: ops.linear(x, w, out, rows=rows, K=K, n_cols=N, bias=b, res=r)
    line = f"linear {rows}x{K}x{N} res={res}:"
    for mode in MODES:
        _lib.check(lib.tc_debug_set_gemm_mode(mode)); torch.cuda.synchronize()
        line += f"  mode{mode} {bk.timeit(fn) * 1e3:.1f} us"
    _lib.check(lib.tc_debug_set_gemm_mode(0))
    print(line, flush=True)
for args in [(81920, 320, 320, False), (81920, 320, 320, True), (81920, 320, 960, False), (20480, 640, 640, True), (81920, 1280, 320, True), (5120, 1280, 1280, True)]:
    run(*args)
