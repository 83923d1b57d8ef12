"""In-kernel timeline of tc_gemm_kernel: per CTA / tile clock64() stamps of the TMA producer, the MMA issuer and one
epilogue warp (tc_debug_set_gemm_mode(4)).  Prints, for a few CTAs, cycles relative to the CTA's first stamp.

    python scripts/trace_gemm.py ROWS K N [res]
"""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from tooncrafter_b200 import _lib, ops
lib = _lib.load()
if sys.argv[1] in ("proj", "qkv", "geglu", "ff2"):
    # the in-situ flavours of scripts/prof_gemm_insitu.py (LayerNorm folds, row statistics, residual, GEGLU)
    sys.path.insert(0, str(Path(__file__).resolve().parent))
    import prof_gemm_insitu
    fn = prof_gemm_insitu.setup()[sys.argv[1]][0]
    rows, K, N, res = sys.argv[1], "", "", ""
else:
    rows, K, N = (int(a) for a in sys.argv[1:4]); res = len(sys.argv) > 4
    x = torch.randn(rows, K, device="cuda").half(); w = (torch.randn(N, K, device="cuda") * K ** -0.5).half()
    b = torch.zeros(N, device="cuda"); r = torch.randn(rows, N, device="cuda").half() if res else None
    out = torch.empty(rows, N, device="cuda", dtype=torch.float16)
    fn = lambda: ops.linear(x, w, out, rows=rows, K=K, n_cols=N, bias=b, res=r)
for _ in range(3): fn()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda"); flush.zero_(); torch.cuda.synchronize()
_lib.check(lib.tc_debug_set_gemm_mode(4)); fn(); torch.cuda.synchronize(); _lib.check(lib.tc_debug_set_gemm_mode(0))
buf = np.zeros(160 * 32 * 16, dtype=np.uint64)
_lib.check(lib.tc_debug_read_gemm_trace(buf.ctypes.data, buf.size))
t = buf.reshape(160, 32, 16).astype(np.int64)
names = ["prod_start", "prod_issued", "mma_start", "mma_1st_full", "mma_commit", "epi_wait", "epi_got", "epi_done",
         "c0_begin", "c0_math", "c0_boxfree", "c0_staged", "c0_synced", "c0_stored", "c0_1st_ld", "epi_bar"]
print(f"linear {rows}x{K}x{N} res={res}   columns: " + " ".join(names))
for cta in (0, 77):
    t0 = t[cta][t[cta] > 0].min() if (t[cta] > 0).any() else 0
    print(f"-- CTA {cta}")
    for ti in range(12):
        if t[cta, ti].max() == 0: break
        print(f"  tile {ti:2d}: " + " ".join(f"{(v - t0) if v else -1:8d}" for v in t[cta, ti]))
    last = max(i for i in range(32) if t[cta, i].max() > 0)
    per = (t[cta, last, 7] - t[cta, 0, 7]) / max(last, 1) if t[cta, last, 7] and t[cta, 0, 7] else 0
    print(f"  tiles traced {last + 1}, epilogue-done to epilogue-done {per:.0f} cycles/tile")
