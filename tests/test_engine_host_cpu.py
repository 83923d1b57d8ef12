"""CPU host-logic tests: the UNet engine's recorded program (weight packing, slices, strides, tap tables, arena
reuse, program order) is interpreted by tests/ops_emulator.py in plain PyTorch and compared with the reference
golden.  This validates everything EXCEPT the CUDA kernels themselves (tests/test_kernels_gpu.py, -m gpu)."""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE / "golden"))

import ops_emulator  # noqa: E402
from make_golden import SEED, golden_inputs  # noqa: E402
from tiny_config import TINY_T, TINY_UNET  # noqa: E402

from tooncrafter_b200 import engine, modules, synthetic  # noqa: E402
from tooncrafter_b200.runtime import Arena  # noqa: E402

GOLD = np.load(HERE / "golden" / "tiny_reference_outputs.npz")


def test_arena_alloc_free_coalesce():
    a = Arena(1 << 20, "cpu")
    offs = [a.alloc(1000) for _ in range(8)]
    assert len(set(offs)) == 8 and all(o % Arena.ALIGN == 0 for o in offs)
    for o in offs[1:7]:
        a.free(o)
    big = a.alloc(6 * 1024)              # the six freed 1 KiB blocks coalesced into one hole
    assert big == offs[1]
    a.free(big); a.free(offs[0]); a.free(offs[7])
    assert a.free_blocks == [(0, 1 << 20)] and a.in_use == 0


def test_unet_program_interpreted_on_cpu_matches_reference_golden():
    m = modules.UNetModel(**TINY_UNET)
    synthetic.fill_module_(m, seed=SEED, prefix="model.diffusion_model.")
    eng = engine.UNetEngine(m.eval(), device="cpu", plan_only=True)
    gi = golden_inputs()["unet"]
    y = eng.forward(gi["x"], gi["t"], gi["ctx"], gi["fs"], executor=ops_emulator.executor)
    ref = torch.from_numpy(GOLD["unet_y"])
    err = (y.float() - ref).abs().max().item()
    assert err < 2e-2 * ref.abs().max().item(), err       # fp16 activation storage in the interpreter
    plan = eng.plan_for(2, TINY_T, 16, 16, gi["ctx"].shape[1])
    assert len(plan.main) > 500 and len(plan.ctx) == 32    # 16 spatial transformers x (text, image) K/V GEMMs
    assert plan.arena.in_use <= plan.arena.high_water <= plan.arena.buf.numel()
    # every block boundary matches its geometry
    for _, name, act in plan.marks:
        assert act.rows * act.C > 0 and act.ld >= act.C
