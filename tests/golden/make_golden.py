#!/usr/bin/env python
"""Generate golden fixtures by RUNNING THE REAL REFERENCE (mahmoodlab/MADELEINE) on CPU, fp32.

Run in the build container only (needs /root/reference):

    cd /tmp && python /root/repo/tests/golden/make_golden.py

The reference has no tests or golden vectors of its own, so these files are what pins the
oracle (oracle/madeleine_oracle.py) and, through it, the CUDA path.  Fixtures hold seeds + small inputs +
reference outputs; the 20 MB checkpoints are re-derived from seeds by tests/golden/weights.py and guarded
by a checksum.  Nothing at test/bench time reads /root/reference.
"""
import os
import sys
from argparse import Namespace

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
# The repo ships drop-in packages named `madeleine`/`core`; make sure the REAL reference wins here.
sys.path = [p for p in sys.path if os.path.abspath(p or os.getcwd()) != REPO]
sys.path.insert(0, "/root/reference")
sys.path.insert(1, HERE)

import numpy as np  # noqa: E402
import torch  # noqa: E402

torch.Tensor.cuda = lambda self, *a, **k: self  # GOT hard-codes .cuda() (quirk Q7)

import madeleine  # noqa: E402

assert madeleine.__file__.startswith("/root/reference"), madeleine.__file__
from madeleine.models.Model import MADELEINE  # noqa: E402
from madeleine.utils.loss import (InfoNCE, GOT, cost_matrix_batch_torch, cos_batch_torch,  # noqa: E402
                                  IPOT_torch_batch_uniform, GW_distance_uniform, IPOT_distance_torch_batch_uniform)
from madeleine.utils.trainer import calculate_losses  # noqa: E402
from weights import make_state_dict, make_feats, make_ragged_lengths, checksum  # noqa: E402

torch.set_num_threads(8)
torch.backends.cuda.matmul.allow_tf32 = False


def cfg(modalities):
    return Namespace(MODALITIES=modalities, wsi_encoder="abmil", patch_embedding_dim=512,
                     wsi_encoder_hidden_dim=512, activation="softmax", n_heads=4)


def build(modalities, stain_encoding, seed):
    model = MADELEINE(cfg(modalities), stain_encoding=stain_encoding)
    sd = make_state_dict(seed, n_mod=len(modalities), stain_encoding=stain_encoding)
    model.load_state_dict(sd, strict=True)
    return model.eval(), sd


def grad_digest(model, n_samples=256):
    """Full grads for small tensors; L2 norm + fixed sampled entries for the big ones."""
    out = {}
    for name, p in model.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        flat = g.detach().flatten()
        if flat.numel() <= 4096:
            out[name] = {"full": flat.clone()}
        else:
            gen = torch.Generator().manual_seed(flat.numel())
            idx = torch.randint(0, flat.numel(), (n_samples,), generator=gen)
            out[name] = {"idx": idx, "samples": flat[idx].clone(), "norm": flat.double().norm().float()}
    return out


def save(name, obj):
    path = os.path.join(HERE, name + ".pt")
    torch.save(obj, path)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


META = {"torch": torch.__version__, "device": "cpu", "dtype": "float32", "allow_tf32": False,
        "reference": "mahmoodlab/MADELEINE@419287dc"}


def gen_encoder():
    out = {"meta": META}
    # --- BASELINE.json configs[0]: single H&E slide, 256 x 512, encode_he on CPU
    model, sd = build(["HE"], False, seed=0)
    feats = make_feats(0, 1, 256, 512)
    with torch.no_grad():
        out["cfg1"] = {"seed_w": 0, "seed_x": 0, "shape": (1, 256, 512), "w_checksum": checksum(sd),
                       "x_checksum": checksum(feats), "encode_he": model.encode_he(feats, "cpu")}
        # embedder internals on a small batch
        x = make_feats(1, 2, 64, 512)
        slide, tok = model.wsi_embedders(x, return_preattn_feats=True)
        slide2, raw = model.wsi_embedders(x, return_attention=True)
        assert torch.equal(slide, slide2)
        out["embedder"] = {"seed_w": 0, "seed_x": 1, "shape": (2, 64, 512), "slide": slide, "raw_attention": raw,
                           "tokens_head": tok[:, :4].clone(), "tokens_checksum": checksum(tok)}
        # n_views = 3 (numpy global RNG, quirk Q10)
        np.random.seed(123)
        sv = model.wsi_embedders(x, n_views=3)
        out["n_views3"] = {"seed_w": 0, "seed_x": 1, "np_seed": 123, "shape": (2, 64, 512), "slide": sv}
        # return_attention through MADELEINE.forward + top-k "attention indices"
        x4 = make_feats(2, 2, 1, 200, 512)
        emb, raw4 = model({"feats": x4}, "cpu", train=False, return_attention=True)
        out["attention"] = {"seed_w": 0, "seed_x": 2, "shape": (2, 1, 200, 512), "emb": emb, "raw_attention": raw4,
                            "argsort": raw4.squeeze(2).transpose(1, 2).argsort(dim=-1, descending=True)}
        ev = model({"feats": x4}, "cpu", train=False)
        out["eval"] = {"seed_w": 0, "seed_x": 2, "shape": (2, 1, 200, 512), "emb": ev["HE"]}
        # ragged bags: the reference cannot batch unequal bags → bs=1 loop (cfg 2 shape, scaled down)
        lens = make_ragged_lengths(1234, 8, 20, 400)
        packed = make_feats(3, sum(lens), 512)
        outs, o = [], 0
        for n in lens:
            outs.append(model.encode_he(packed[o:o + n].unsqueeze(0), "cpu"))
            o += n
        out["ragged"] = {"seed_w": 0, "seed_x": 3, "lens": lens, "encode_he": torch.cat(outs, 0)}
    save("encoder", out)


def gen_forward_train():
    out = {"meta": META}
    mods = ["HE", "HER2", "PGR"]
    for se in (False, True):
        model, sd = build(mods, se, seed=1)
        feats = make_feats(4, 3, 3, 48, 512)
        with torch.no_grad():
            embs, toks = model({"feats": feats}, "cpu", train=True, n_views=1)
            entry = {"seed_w": 1, "seed_x": 4, "shape": (3, 3, 48, 512), "modalities": mods,
                     "w_checksum": checksum(sd),
                     "embs": {k: v.clone() for k, v in embs.items()},
                     "toks": {k: v.clone() for k, v in toks.items()}}
            if se:
                # NB: the reference's eval + stain-encoding path only works for bs == 1 (Model.py:186-189 builds a
                # [1, bs*T, 32] encoding); use one case.
                ev = model({"feats": feats[:1, 1:2]}, "cpu", train=False, custom_stain_idx=1)
                entry["eval_custom_stain1"] = {k: v.clone() for k, v in ev.items()}
        out["stain_enc" if se else "plain"] = entry
    save("forward_train", out)


def gen_infonce():
    out = {"meta": META, "cases": []}
    g = torch.Generator().manual_seed(99)
    for m, d, tau, sym, scale in [(2, 512, 0.001, True, 1.0), (16, 512, 0.001, True, 1.0), (16, 512, 0.1, False, 1.0),
                                  (65, 512, 0.001, True, 3.0), (33, 512, 0.07, True, 0.01), (7, 128, 0.001, False, 1.0)]:
        q = (torch.randn(m, d, generator=g) * scale).requires_grad_(True)
        k = (torch.randn(m, d, generator=g) * scale)
        # make positives correlated so the loss is not saturated everywhere
        k = (0.7 * q.detach() + 0.3 * k).requires_grad_(True)
        loss = InfoNCE(temperature=tau)(query=q, positive_key=k, symmetric=sym)
        loss.backward()
        out["cases"].append({"q": q.detach().clone(), "k": k.detach().clone(), "tau": tau, "symmetric": sym,
                             "loss": loss.detach().clone(), "dq": q.grad.clone(), "dk": k.grad.clone()})
    save("infonce", out)


def gen_got():
    out = {"meta": META, "cases": [], "internals": {}}
    g = torch.Generator().manual_seed(7)
    for m, n_tok in [(2, 16), (5, 24), (16, 32), (33, 40)]:
        base = torch.randn(m, n_tok, 128, generator=g)
        v = (base + 0.5 * torch.randn(m, n_tok, 128, generator=g)).requires_grad_(True)
        q = (base + 0.5 * torch.randn(m, n_tok, 128, generator=g)).requires_grad_(True)
        torch.manual_seed(1000 + m)
        perm = torch.randperm(m)  # what GOT will draw (quirk Q3)
        torch.manual_seed(1000 + m)
        loss = GOT(v, q, subsample=256)
        loss.backward()
        out["cases"].append({"v": v.detach().clone(), "q": q.detach().clone(), "torch_seed": 1000 + m,
                             "perm": perm, "loss": loss.detach().clone(),
                             "dv": v.grad.clone(), "dq": q.grad.clone()})
    # internals on one small problem
    with torch.no_grad():
        x = torch.randn(3, 128, 12, generator=g)
        y = torch.randn(3, 128, 12, generator=g)
        C = cost_matrix_batch_torch(x, y)
        out["internals"] = {"x": x, "y": y, "cost": C, "cos_thresh": cos_batch_torch(x, x),
                            "ipot_T": IPOT_torch_batch_uniform(C, 3, 12, 12, beta=0.5, iteration=30),
                            "ipot_dist": IPOT_distance_torch_batch_uniform(C, 3, 12, 12, 30),
                            "gw": GW_distance_uniform(x, y)}
    save("got", out)


def gen_losses_and_grads():
    """configs[2]-shaped (scaled down): 5 stains, stain encodings, availability mask, global + local loss, grads."""
    out = {"meta": META}
    mods = ["HE", "HER2", "PGR", "KI67", "ER"]
    stains = mods[1:]
    for tag, use_local, se, bs, T, seed in [("global_only", False, False, 6, 40, 2), ("global_local_se", True, True, 6, 40, 3)]:
        model, sd = build(mods, se, seed=seed)
        feats = make_feats(10 + seed, bs, len(mods), T, 512)
        gm = torch.Generator().manual_seed(seed)
        labels = (torch.rand(bs, len(mods), generator=gm) < torch.tensor([1.0, 0.46, 0.73, 0.73, 0.73])).float()
        labels[:, 0] = 1
        labels[:2, 1:] = 1  # make sure ≥ 2 cases for each stain
        feats = feats * labels[:, :, None, None]  # missing stain ⇒ all-zero bag (quirk Q8)
        args = Namespace(global_loss="info-nce", symmetric_cl=True, local_loss_weight=1.0)
        embs, toks = model({"feats": feats}, "cpu", train=True, n_views=1)
        torch.manual_seed(4242)
        loss, flag = calculate_losses(stains, InfoNCE(temperature=0.001), GOT if use_local else None, None,
                                      embs, toks, labels[:, 1:], args)
        model.zero_grad()
        loss.backward()
        out[tag] = {"seed_w": seed, "seed_x": 10 + seed, "shape": tuple(feats.shape), "modalities": mods,
                    "labels": labels, "stain_encoding": se, "torch_seed": 4242, "loss": loss.detach().clone(),
                    "flag": flag, "grads": grad_digest(model), "w_checksum": checksum(sd)}
    # ragged cfg-2-style fwd+bwd through encode path is covered by the oracle itself (ragged is not a reference feature)
    save("losses_grads", out)


if __name__ == "__main__":
    gen_encoder()
    gen_forward_train()
    gen_infonce()
    gen_got()
    gen_losses_and_grads()
