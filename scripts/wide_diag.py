"""Diagnostic matrix for tc_attention_wide: error by (D, L) and its structure (per 64-channel block, per 32-row block)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from tooncrafter_b200 import ops

DEV = "cuda"
torch.manual_seed(0)


def run(N, L, D, mode="rand"):
    qkv = torch.randn(N, L, 3 * D, device=DEV)
    if mode == "q0":
        qkv[..., :D] = 0
    qkv = qkv.half()
    out = torch.zeros(N, L, D, dtype=torch.float16, device=DEV)
    ops.attention_wide(qkv, out, batches=N, L=L, D=D, scale=D ** -0.5, ld=3 * D, ldo=D, k_offset=D, v_offset=2 * D)
    torch.cuda.synchronize()
    q, k, v = (qkv[..., i * D:(i + 1) * D].float() for i in range(3))
    ref = ((q @ k.transpose(-1, -2)) * D ** -0.5).softmax(-1) @ v
    d = (out.float() - ref).abs()
    cb = d.reshape(N, L, D // 64, 64).amax((0, 1, 3)).tolist()
    rb = d[0].amax(1)
    rb = [rb[i:i + 32].max().item() for i in range(0, L, 32)]
    print(f"N={N} L={L} D={D} {mode}: max err {d.max().item():.3e} ref max {ref.abs().max().item():.2f} | per 64-ch block "
          + " ".join(f"{x:.1e}" for x in cb) + " | per 32 rows " + " ".join(f"{x:.1e}" for x in rb[:12]), flush=True)


for D in (64, 128, 256, 512):
    for L in (128, 256, 384):
        run(1, L, D)
run(1, 128, 256, "q0")
run(1, 256, 512, "q0")
run(2, 128, 128)
