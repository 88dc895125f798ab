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

torch.Tensor.cuda = lambda self, *a, **k: self  # GOT hard-codes .cuda() (quirk Q7)

import madeleine  # noqa: E402

assert madeleine.__file__.startswith("/root/reference"), madeleine.__file__
from madeleine.models.Model import MADELEINE  # noqa: E402
from madeleine.utils.loss import InfoNCE, GOT  # noqa: E402
from madeleine.utils.trainer import calculate_losses  # noqa: E402
from weights import make_state_dict, make_feats  # noqa: E402

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
    # For context (reported next to this repo's bf16 mode, not a parity target): the same evaluation under torch.autocast(bfloat16), the
    # precision the reference's scripts train in (trainer.py:108 wraps forward AND loss).  The in-batch logits are then a bf16 matmul of
    # +-1000-sized values (tau = 0.001): the loss and the gradients come out far from the fp32 ones.
    with torch.autocast("cpu", dtype=torch.bfloat16):
        ac = torch.cat([model.projector(model.wsi_embedders(x[cu[r]:cu[r + 1]].unsqueeze(0)).reshape(1, -1)) for r in range(32)])
        ac_loss = InfoNCE(temperature=TAU)(query=ac[:16], positive_key=ac[16:], symmetric=True)
    model.zero_grad()
    ac_loss.backward()
    total = sum(float(d["norm"]) ** 2 for d in grads.values()) ** 0.5
    worst = max(abs(float(p.grad.double().norm()) - float(grads[n]["norm"])) / float(grads[n]["norm"])
                for n, p in model.named_parameters() if n in grads and float(grads[n]["norm"]) >= 1e-6 * total)
    out["autocast_bf16_cpu"] = {"slide": ac.detach().float().clone(), "loss": ac_loss.detach().float().clone(),
                                "emb_max_abs_err_vs_fp32": float((ac.detach().float() - slide.detach()).abs().max()),
                                "grad_norm_rel_err_max_vs_fp32": worst}
    path = os.path.join(HERE, "baseline_sizes.pt")
    torch.save(out, path)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB); loss {float(loss):.6f}; {out['meta']['seconds']:.1f} s")


def digest(model):
    grads = {}
    for name, p in model.named_parameters():
        if p.grad is None:
            continue
        flat = p.grad.detach().flatten()
        gen = torch.Generator().manual_seed(flat.numel())
        idx = torch.randint(0, flat.numel(), (64,), generator=gen)
        grads[name] = {"norm": flat.double().norm(), "idx": idx, "samples": flat[idx].clone()}
    return grads


def config3(T=512):
    """BASELINE configs[2] (batch 32, 5 stains, stain encodings, ACROBAT availability rates, InfoNCE tau = 0.001 + GOT) through the
    reference's own forward(train=True) + calculate_losses + backward.  The reference materialises every activation of the whole
    batch for autograd (~120 KB per token in fp32: 39 GB at the configuration's 2048 tokens per bag, more than this container can
    give), so the fixture uses 512 tokens per bag: 81 920 tokens, every code path of the full configuration."""
    mods = ["HE", "HER2", "PGR", "KI67", "ER"]
    cfg = Namespace(MODALITIES=mods, wsi_encoder="abmil", patch_embedding_dim=512, wsi_encoder_hidden_dim=512, activation="softmax",
                    n_heads=4)
    model = MADELEINE(cfg, stain_encoding=True)
    model.load_state_dict(make_state_dict(3, n_mod=5, stain_encoding=True), strict=True)
    model.eval()
    bs = 32
    g = torch.Generator().manual_seed(0)
    labels = (torch.rand(bs, 5, generator=g) < torch.tensor([1.0, 0.46, 0.73, 0.73, 0.73])).float()
    labels[:, 0] = 1
    feats = torch.randn(bs, 5, T, 512, generator=g) * labels[:, :, None, None]
    args = Namespace(global_loss="info-nce", symmetric_cl=True, local_loss_weight=1.0)
    t0 = time.time()
    embs, toks = model({"feats": feats}, "cpu", train=True, n_views=1)
    torch.manual_seed(11)
    loss, flag = calculate_losses(mods[1:], InfoNCE(temperature=TAU), GOT, None, embs, toks, labels[:, 1:], args)
    model.zero_grad()
    loss.backward()
    out = {"meta": {"torch": torch.__version__, "device": "cpu", "dtype": "float32", "reference": "mahmoodlab/MADELEINE@419287dc",
                    "seconds": time.time() - t0},
           "T": T, "bs": bs, "tau": TAU, "loss_seed": 11, "labels": labels, "x_checksum": float(feats.double().abs().sum()),
           "embs": {m: embs[m].detach().clone() for m in mods}, "tok_sum": {m: float(toks[m].detach().double().sum()) for m in mods},
           "tok_head": {m: toks[m].detach()[:, :2].clone() for m in mods}, "loss": loss.detach().clone(), "flag": flag, "grads": digest(model)}
    path = os.path.join(HERE, "baseline_config3_t512.pt")
    torch.save(out, path)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB); loss {float(loss.detach()):.6f}; {out['meta']['seconds']:.1f} s")


def inference():
    """BASELINE configs[4] shape: slides of 4000 x 512 through the reference's inference entry points, H&E model without stain
    encodings.  `encode_he` (Model.py:95-106, what bin/extract_slide_embeddings.py calls) for six slides, and
    `forward(train=False, return_attention=True)` (Model.py:205-216) for two of them: the slide embedding plus the raw attention
    logits [1, 4000, 1, 4] whose ranking is the "attention indices" of the north star."""
    cfg = Namespace(MODALITIES=["HE"], wsi_encoder="abmil", patch_embedding_dim=512, wsi_encoder_hidden_dim=512, activation="softmax", n_heads=4)
    model = MADELEINE(cfg, stain_encoding=False)
    model.load_state_dict(make_state_dict(0), strict=True)
    model.eval()
    out = {"meta": {"torch": torch.__version__, "device": "cpu", "dtype": "float32", "reference": "mahmoodlab/MADELEINE@419287dc"},
           "seed_w": 0, "n_tokens": 4000, "seeds": list(range(100, 106)), "encode_he": [], "attention": {}}
    with torch.no_grad():
        for seed in out["seeds"]:
            x = make_feats(seed, 1, 4000, 512)
            out["encode_he"].append(model.encode_he(x, "cpu").clone())
        for seed in out["seeds"][:2]:
            x = make_feats(seed, 1, 1, 4000, 512)
            emb, raw = model({"feats": x}, "cpu", train=False, return_attention=True)
            out["attention"][seed] = {"emb": emb.clone(), "raw": raw.clone()}
    out["encode_he"] = torch.cat(out["encode_he"])
    path = os.path.join(HERE, "baseline_inference.pt")
    torch.save(out, path)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


def config4_global_batch():
    """BASELINE configs[3]: batch = 64 cases (x 2 stains x 2000 patches = 256 000 tokens), the batch that 8 ranks x 8 cases shard,
    evaluated by the reference as ONE global batch: `forward(train=True)`'s encoder + projector per case (the reference's own
    modules; a monolithic call would need ~31 GB of autograd tape), `InfoNCE` (tau = 0.001, symmetric) over the 64 HE / IHC pairs
    through `calculate_losses`, and backward in two stages (loss -> slide embeddings, then case by case into the parameters: the
    same gradients as one tape, by the chain rule)."""
    mods = ["HE", "IHC"]
    cfg = Namespace(MODALITIES=mods, wsi_encoder="abmil", patch_embedding_dim=512, wsi_encoder_hidden_dim=512, activation="softmax", n_heads=4)
    model = MADELEINE(cfg, stain_encoding=False)
    model.load_state_dict(make_state_dict(0, n_mod=2), strict=True)
    model.eval()
    bs, T = 64, 2000
    feats = make_feats(64, bs, 2, T, 512)
    args = Namespace(global_loss="info-nce", symmetric_cl=True, local_loss_weight=1.0)
    t0 = time.time()
    with torch.no_grad():
        parts = [model({"feats": feats[c:c + 1]}, "cpu", train=True, n_views=1)[0] for c in range(bs)]
    embs = {m: torch.cat([p[m] for p in parts]).requires_grad_(True) for m in mods}      # HE [64, 1, 512, 1], IHC [64, 1, 512]
    loss, flag = calculate_losses(mods[1:], InfoNCE(temperature=TAU), None, None, embs, None, torch.ones(bs, 1), args)
    loss.backward()
    model.zero_grad()
    for c in range(bs):
        e, _ = model({"feats": feats[c:c + 1]}, "cpu", train=True, n_views=1)
        torch.autograd.backward([e[m] for m in mods], [embs[m].grad[c:c + 1] for m in mods])
    out = {"meta": {"torch": torch.__version__, "device": "cpu", "dtype": "float32", "reference": "mahmoodlab/MADELEINE@419287dc",
                    "seconds": time.time() - t0},
           "bs": bs, "T": T, "seed_x": 64, "tau": TAU, "x_checksum": float(feats.double().abs().sum()),
           "embs": {m: embs[m].detach().clone() for m in mods}, "loss": loss.detach().clone(), "flag": flag, "grads": digest(model)}
    path = os.path.join(HERE, "baseline_config4_global.pt")
    torch.save(out, path)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB); loss {float(loss.detach()):.6f}; {out['meta']['seconds']:.1f} s")


def got_shipped_batch():
    """GOT (loss.py:278-301) at the reference's shipped batch size and at the boundary of the shared-memory kernel: 65 and 96 cases
    with the stain (problems of 65 / 96 tokens, quirk Q3), token embeddings [m, m + 16, 128].  Loss, gradient norms, the full gradient
    of the first two cases and 512 sampled entries."""
    out = {"meta": {"torch": torch.__version__, "device": "cpu", "dtype": "float32", "reference": "mahmoodlab/MADELEINE@419287dc"}, "cases": []}
    for m in (65, 96):
        v = make_feats(700 + m, m, m + 16, 128).requires_grad_(True)
        q = (v.detach() + 0.5 * make_feats(800 + m, m, m + 16, 128)).requires_grad_(True)
        torch.manual_seed(2000 + m)
        loss = GOT(v, q, subsample=256)
        loss.backward()
        case = {"m": m, "torch_seed": 2000 + m, "seed_v": 700 + m, "seed_q": 800 + m, "loss": loss.detach().clone()}
        for nm, t in (("dv", v.grad), ("dq", q.grad)):
            flat = t.flatten()
            idx = torch.randint(0, flat.numel(), (512,), generator=torch.Generator().manual_seed(m))
            case[nm] = {"norm": flat.double().norm(), "idx": idx, "samples": flat[idx].clone(), "first2": t[:2].clone(),
                        "beyond_m_abs_max": float(t[:, m:].abs().max())}
        out["cases"].append(case)
        print(f"m = {m}: loss {float(loss.detach()):.6f}")
    path = os.path.join(HERE, "got_shipped_batch.pt")
    torch.save(out, path)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


def n_views3_backward():
    """forward(train=True, n_views=3) (Model.py:419-440: whole view + two random half views, numpy's global RNG) + calculate_losses with
    the global AND the intra-modality InfoNCE (trainer.py:52-66), backward: 6 cases x 2 stains x 300 patches, tau = 0.1, for the
    four attention activations.  Embeddings of all three views, loss, gradient digests."""
    import numpy as np
    out = {"meta": {"torch": torch.__version__, "numpy": np.__version__, "device": "cpu", "dtype": "float32",
                    "reference": "mahmoodlab/MADELEINE@419287dc"}, "np_seed": 77, "seed_x": 5, "seed_w": 21, "shape": (6, 2, 300, 512), "tau": 0.1,
           "activations": {}}
    mods = ["HE", "IHC"]
    args = Namespace(global_loss="info-nce", symmetric_cl=True, local_loss_weight=1.0)
    for activation in ("softmax", "relu", "leaky_relu", "sigmoid"):
        cfg = Namespace(MODALITIES=mods, wsi_encoder="abmil", patch_embedding_dim=512, wsi_encoder_hidden_dim=512, activation=activation, n_heads=4)
        model = MADELEINE(cfg, stain_encoding=False)
        model.load_state_dict(make_state_dict(21, n_mod=2), strict=True)
        model.eval()
        feats = make_feats(5, 6, 2, 300, 512)
        fn = InfoNCE(temperature=0.1)
        np.random.seed(77)
        embs, toks = model({"feats": feats}, "cpu", train=True, n_views=3)
        loss, flag = calculate_losses(mods[1:], fn, None, fn, embs, toks, torch.ones(6, 1), args)
        model.zero_grad()
        loss.backward()
        out["activations"][activation] = {"embs": {m: embs[m].detach().clone() for m in mods}, "loss": loss.detach().clone(), "grads": digest(model)}
        print(activation, float(loss.detach()))
    path = os.path.join(HERE, "n_views3_backward.pt")
    torch.save(out, path)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


if __name__ == "__main__":
    if "--nviews3" in sys.argv:
        n_views3_backward()
    elif "--got65" in sys.argv:
        got_shipped_batch()
    elif "--config4" in sys.argv:
        config4_global_batch()
    elif "--config3" in sys.argv:
        config3()
    elif "--inference" in sys.argv:
        inference()
    else:
        main()
