import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from tooncrafter_b200 import ops
N, L, D = (int(a) for a in sys.argv[1:4])
torch.manual_seed(0)
qkv = torch.randn(N, L, 3 * D, device="cuda").half()
out = torch.zeros(N, L, D, dtype=torch.float16, device="cuda")
for it in range(int(sys.argv[4]) if len(sys.argv) > 4 else 3):
    ops.attention_wide(qkv, out, batches=N, L=L, D=D, scale=D ** -0.5, ld=3 * D, ldo=D, k_offset=D, v_offset=2 * D)
    torch.cuda.synchronize()
q, k, v = (qkv[..., i * D:(i + 1) * D].float() for i in range(3))
ref = ((q @ k.transpose(-1, -2)) * D ** -0.5).softmax(-1) @ v
print(f"N={N} L={L} D={D}: max err {(out.float() - ref).abs().max().item():.3e}", flush=True)
