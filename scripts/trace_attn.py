"""In-kernel timeline of tc_attn3_kernel (trace build: TC_BUILD_TRACE=1 / -DTC_ATTN_TRACE=1, loaded through TC_LIB_PATH):
per key block, cycles since the first stamp, for one softmax thread of each query tile and the MMA issuer of CTA (0,0,0).

    TC_LIB_PATH=tooncrafter_b200/libtc_trace.so python scripts/trace_attn.py
"""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from tooncrafter_b200 import _lib, ops
lib = _lib.load()
B, L, heads = 32, 2560, 5
C = heads * 64
q, k, v = (torch.randn(B, L, C, device="cuda").half() for _ in range(3))
out = torch.zeros_like(q)
fn = lambda: ops.attention(q, [dict(k=k, v=v, ldk=C, ldv=C, Lk=L)], out, q_batches=B, Lq=L, heads=heads, scale=64 ** -0.5, ldq=C, ldo=C)
for _ in range(3): fn()
torch.cuda.synchronize()
buf = np.zeros(24 * 16, dtype=np.uint64)
_lib.check(lib.tc_debug_read_attn_trace(buf.ctypes.data, buf.size))
t = buf.reshape(24, 16).astype(np.int64)
t0 = t[t > 0].min()
names = ["t0:S_ok", "t0:S_regs", "t0:max", "t0:PV_ok", "t0:exp", "t0:P_st", "t1:S_ok", "t1:S_regs", "t1:max", "t1:PV_ok", "t1:exp", "t1:P_st",
         "mma:S0", "mma:S1", "mma:PV0", "mma:PV1"]
print("block " + " ".join(f"{n:>9s}" for n in names))
for g in range(20):
    print(f"{g:5d} " + " ".join(f"{(x - t0) if x else -1:9d}" for x in t[g]))
d = np.diff(t[2:19, 5])
print("tile0 P_st to P_st per block:", d.tolist(), "mean", d.mean())
