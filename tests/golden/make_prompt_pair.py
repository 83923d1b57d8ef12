"""Fixture for BASELINE config #1: the reference's first interpolation pair, pre-processed exactly like
scripts/evaluation/inference.py:65-69 (Resize(min(320, 512)) -> CenterCrop((320, 512)) -> ToTensor), stored as
uint8 [2, 320, 512, 3] (the Normalize(0.5, 0.5) step is x / 127.5 - 1 at load time).  Run in the authoring container:

    python tests/golden/make_prompt_pair.py

/root/reference does not exist on the GPU box, so bench.py --config pair10 reads this derived fixture instead.
"""
from pathlib import Path

import numpy as np
from PIL import Image
from torchvision import transforms

HERE = Path(__file__).resolve().parent
SRC = Path("/root/reference/prompts/512_interp")
NAMES = ("74906_1462_frame1.png", "74906_1462_frame3.png")


def main():
    tf = transforms.Compose([transforms.Resize(min((320, 512))), transforms.CenterCrop((320, 512))])
    frames = np.stack([np.asarray(tf(Image.open(SRC / n).convert("RGB")), dtype=np.uint8) for n in NAMES])
    assert frames.shape == (2, 320, 512, 3)
    np.savez_compressed(HERE / "prompt_pair_74906.npz", frames=frames)
    print("wrote prompt_pair_74906.npz", frames.shape, frames.mean())


if __name__ == "__main__":
    main()
