"""Run ONE kernel configuration a few times (target for `ncu --set full -k regex:<kernel> -s 2 -c 1`).

    python scripts/prof_one.py conv320 | conv640 | conv512 | lin320 | geglu320 | attn2560 | attnfusion | attnwide | gn320 | gn1280
"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))
import bench_kernels as bk

CASES = {
    "conv320": lambda: bk.bench_conv(32, 40, 64, 320, 320),
    "conv640": lambda: bk.bench_conv(32, 20, 32, 640, 640),
    "conv1280": lambda: bk.bench_conv(32, 10, 16, 1280, 1280),
    "conv512": lambda: bk.bench_conv(16, 80, 128, 512, 512),
    "lin320": lambda: bk.bench_linear(81920, 320, 320),
    "lin1280": lambda: bk.bench_linear(5120, 1280, 1280),
    "geglu320": lambda: bk.bench_linear(81920, 320, 2560, geglu=True),
    "attn2560": lambda: bk.bench_attn(32, 2560, 5),
    "attn640": lambda: bk.bench_attn(32, 640, 10),
    "attnfusion": lambda: bk.bench_attn(16, 10240, 8, 20480),     # VAE level-2 dual-reference fusion attention
    "attnwide": lambda: bk.bench_attn_wide(16, 2560, 512),         # VAE mid-block AttnBlock, D = 512
    "gn1280": lambda: bk.bench_gn(32, 16, 160, 1280),              # single-pass GroupNorm (cluster of 3 CTAs per unit)
    "gn320": lambda: bk.bench_gn(32, 1, 2560, 320),
    "gn320t": lambda: bk.bench_gn(32, 16, 2560, 320),
}

if __name__ == "__main__":
    for name in sys.argv[1:]:
        print(name, CASES[name](), flush=True)
