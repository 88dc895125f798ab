#!/usr/bin/env python
"""More fixtures from the REAL reference (run in the build container only, like make_golden.py):

    cd /tmp && python /root/repo/tests/golden/make_golden_variants.py

They pin the oracle on the settings the later GPU tests check against it rather than against a fixture: other
patch-embedding widths (config.patch_embedding_dim = 1024 / 768 / 384, with and without stain encodings) through
forward(train=True) + calculate_losses (InfoNCE + GOT), and the non-softmax attention activations of ABMILEmbedder."""
import os
import sys
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

torch.set_num_threads(8)
MODS = ["HE", "ER", "PR"]


def cfg(mods, d_in=512, activation="softmax"):
    return Namespace(MODALITIES=mods, wsi_encoder="abmil", patch_embedding_dim=d_in, wsi_encoder_hidden_dim=512,
                     activation=activation, n_heads=4)


def main():
    out = {"torch": torch.__version__, "widths": {}, "activations": {}}
    args = Namespace(global_loss="info-nce", symmetric_cl=True, local_loss_weight=1.0)
    for d_in, se in [(1024, False), (768, True), (384, False), (1536, True), (128, True)]:
        bs, T = 3, 40
        sd = make_state_dict(21, n_mod=3, stain_encoding=se, d_in=d_in)
        model = MADELEINE(cfg(MODS, d_in), stain_encoding=se)
        model.load_state_dict(sd, strict=True)
        model.eval()
        feats = make_feats(d_in, bs, 3, T, d_in)
        labels = torch.ones(bs, 3)
        embs, toks = model({"feats": feats}, "cpu", train=True, n_views=1)
        torch.manual_seed(5)
        loss, _ = calculate_losses(MODS[1:], InfoNCE(temperature=0.1), GOT, None, embs, toks, labels[:, 1:], args)
        model.zero_grad()
        loss.backward()
        out["widths"][(d_in, se)] = {
            "seed_w": 21, "seed_x": d_in, "shape": (bs, 3, T, d_in), "torch_seed": 5, "tau": 0.1,
            "embs": {m: embs[m].detach().clone() for m in MODS}, "tok_sum": {m: toks[m].detach().double().sum() for m in MODS},
            "tok_head": {m: toks[m].detach()[:, :2].clone() for m in MODS}, "loss": loss.detach().clone(),
            "grad_norms": {n: p.grad.double().norm() for n, p in model.named_parameters() if p.grad is not None}}
    for act in ["leaky_relu", "relu", "sigmoid"]:
        sd = make_state_dict(31, n_mod=1)
        model = MADELEINE(cfg(["HE"], 512, act), stain_encoding=False)
        model.load_state_dict(sd, strict=True)
        model.eval()
        x = make_feats(77, 2, 37, 512)
        with torch.no_grad():
            slide = model.wsi_embedders(x)
        out["activations"][act] = {"seed_w": 31, "seed_x": 77, "shape": (2, 37, 512), "slide": slide.clone()}
    # GOT problems of more than 96 tokens (batches of more than 96 cases): the token generator of tests/test_gpu_got.py::_tokens
    out["got_large"] = {}
    for m, n in [(3, 130), (2, 256)]:
        g = torch.Generator().manual_seed(n)
        v = torch.randn(m, n, 128, generator=g)
        q = v + 0.5 * torch.randn(m, n, 128, generator=g)
        vr, qr = v.clone().requires_grad_(True), q.clone().requires_grad_(True)
        loss = GOT(vr, qr)                       # subsample=None: the tokens are used as given
        loss.backward()
        out["got_large"][(m, n)] = {"seed": n, "loss": loss.detach().clone(), "dv": vr.grad.clone(), "dq": qr.grad.clone()}
    torch.save(out, os.path.join(HERE, "variants.pt"))
    print("wrote variants.pt", os.path.getsize(os.path.join(HERE, "variants.pt")), "bytes")


if __name__ == "__main__":
    main()
