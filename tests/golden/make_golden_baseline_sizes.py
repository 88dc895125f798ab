#!/usr/bin/env python
"""BASELINE-size fixture from the REAL reference (run in the build container only, like make_golden.py):

    cd /tmp && python /root/repo/tests/golden/make_golden_baseline_sizes.py

BASELINE.json configs[1] as written: 16 cases x 2 stains = 32 bags with N_i ~ randint(200, 4001) patches (generator seed 1234, 68 063
tokens), symmetric InfoNCE at the scripts' temperature 0.001, forward + backward, fp32 on the CPU.  The reference cannot batch
unequal bags, so it is driven the way its own bs = 1 paths drive it: ``wsi_embedders`` + ``projector`` bag by bag
(Model.py:97-107), the 32 slide embeddings stacked, ``InfoNCE`` (loss.py:58-133) on HE vs IHC, ``backward()``.  Stored: the bag
lengths, the slide embeddings, the loss and per-parameter gradient digests — tests/test_gpu_baseline_parity.py regenerates the
inputs from the same seeds and compares the CUDA path with these numbers directly (no oracle in between)."""
import os
import sys
import time
from argparse import Namespace

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path = [p for p in sys.path if os.path.abspath(p or os.getcwd()) != REPO]
sys.path.insert(0, "/root/reference")
sys.path.insert(1, HERE)

import torch  # noqa: E402

import madeleine  # noqa: E402

assert madeleine.__file__.startswith("/root/reference"), madeleine.__file__
from madeleine.models.Model import MADELEINE  # noqa: E402
from madeleine.utils.loss import InfoNCE  # noqa: E402
from weights import make_state_dict  # noqa: E402

torch.set_num_threads(os.cpu_count() or 8)
torch.backends.cuda.matmul.allow_tf32 = False
TAU = 0.001


def main():
    mods = ["HE", "IHC"]
    cfg = Namespace(MODALITIES=mods, wsi_encoder="abmil", patch_embedding_dim=512, wsi_encoder_hidden_dim=512, activation="softmax",
                    n_heads=4)
    model = MADELEINE(cfg, stain_encoding=False)
    model.load_state_dict(make_state_dict(0, n_mod=2), strict=True)
    model.eval()
    g = torch.Generator().manual_seed(1234)
    lens = torch.randint(200, 4001, (32,), generator=g).tolist()
    cu = [0]
    for n in lens:
        cu.append(cu[-1] + n)
    x = torch.randn(cu[-1], 512, generator=g)                 # the same draw as tests/test_gpu_baseline_parity.py
    t0 = time.time()
    slides = []
    for r in range(32):
        bag = x[cu[r]:cu[r + 1]].unsqueeze(0)                 # [1, N_r, 512]
        emb = model.wsi_embedders(bag)                        # [1, 1, 512, 4] (ABMILEmbedder.forward, Model.py:383-451)
        slides.append(model.projector(emb.reshape(1, -1)))    # Model.py:105-106
    slide = torch.cat(slides)                                 # bags 0..15 = HE, 16..31 = IHC (the GPU test's convention)
    loss = InfoNCE(temperature=TAU)(query=slide[:16], positive_key=slide[16:], symmetric=True)
    model.zero_grad()
    loss.backward()
    grads = {}
    for name, p in model.named_parameters():
        if p.grad is None:
            continue
        flat = p.grad.detach().flatten()
        gen = torch.Generator().manual_seed(flat.numel())
        idx = torch.randint(0, flat.numel(), (64,), generator=gen)
        grads[name] = {"norm": flat.double().norm(), "idx": idx, "samples": flat[idx].clone()}
    out = {"meta": {"torch": torch.__version__, "device": "cpu", "dtype": "float32", "reference": "mahmoodlab/MADELEINE@419287dc",
                    "seconds": time.time() - t0},
           "lens": lens, "tau": TAU, "x_checksum": float(x.double().abs().sum()), "slide": slide.detach().clone(), "loss": loss.detach().clone(),
           "grads": grads}
    path = os.path.join(HERE, "baseline_sizes.pt")
    torch.save(out, path)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB); loss {float(loss):.6f}; {out['meta']['seconds']:.1f} s")


if __name__ == "__main__":
    main()
