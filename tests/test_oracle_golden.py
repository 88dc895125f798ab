"""Pin the CPU oracle against fixtures produced by the real reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

import oracle
from weights import make_state_dict, make_feats, checksum

RTOL, ATOL = 1e-4, 1e-5  # oracle vs reference, both fp32 on CPU: only summation-order noise


def close(a, b, rtol=RTOL, atol=ATOL):
    torch.testing.assert_close(a, b, rtol=rtol, atol=atol)


def test_cfg1_encode_he(golden):
    g = golden("encoder")["cfg1"]
    sd = make_state_dict(g["seed_w"])
    assert checksum(sd) == pytest.approx(g["w_checksum"], rel=1e-12), "seeded weights drifted (torch RNG changed?)"
    x = make_feats(g["seed_x"], *g["shape"])
    assert checksum(x) == pytest.approx(g["x_checksum"], rel=1e-12)
    close(oracle.encode_he(sd, x), g["encode_he"])


def test_embedder_internals(golden):
    g = golden("encoder")["embedder"]
    sd = make_state_dict(g["seed_w"])
    x = make_feats(g["seed_x"], *g["shape"])
    slide, tok, raw = oracle.abmil_embedder(sd, x)
    close(slide, g["slide"])
    close(raw, g["raw_attention"])
    close(tok[:, :4], g["tokens_head"])
    assert checksum(tok) == pytest.approx(g["tokens_checksum"], rel=1e-5)


def test_n_views3(golden):
    g = golden("encoder")["n_views3"]
    sd = make_state_dict(g["seed_w"])
    x = make_feats(g["seed_x"], *g["shape"])
    np.random.seed(g["np_seed"])
    slide, _, _ = oracle.abmil_embedder(sd, x, n_views=3)
    close(slide, g["slide"])


def test_attention_and_eval(golden):
    enc = golden("encoder")
    g = enc["attention"]
    sd = make_state_dict(g["seed_w"])
    x = make_feats(g["seed_x"], *g["shape"])
    emb, raw = oracle.madeleine_forward_attention(sd, x)
    close(emb, g["emb"])
    close(raw, g["raw_attention"])
    # "attention indices": same ranking up to near-ties in the reference's own logits
    order = raw.squeeze(2).transpose(1, 2).argsort(dim=-1, descending=True)
    ref_logits = g["raw_attention"].squeeze(2).transpose(1, 2)
    ref_sorted = torch.gather(ref_logits, -1, g["argsort"])
    assert float((torch.gather(ref_logits, -1, order) - ref_sorted).abs().max()) <= 1e-5
    assert torch.equal(oracle.topk_attention_indices(raw, 8), g["argsort"][..., :8])
    ev = oracle.madeleine_forward_eval(sd, x, ["HE"])
    close(ev["HE"], enc["eval"]["emb"])


def test_ragged(golden):
    g = golden("encoder")["ragged"]
    sd = make_state_dict(g["seed_w"])
    lens = g["lens"]
    x = make_feats(g["seed_x"], sum(lens), 512)
    cu = np.concatenate([[0], np.cumsum(lens)]).tolist()
    close(oracle.encode_packed(sd, x, cu), g["encode_he"])


@pytest.mark.parametrize("tag", ["plain", "stain_enc"])
def test_forward_train(golden, tag):
    g = golden("forward_train")[tag]
    se = tag == "stain_enc"
    sd = make_state_dict(g["seed_w"], n_mod=3, stain_encoding=se)
    assert checksum(sd) == pytest.approx(g["w_checksum"], rel=1e-12)
    x = make_feats(g["seed_x"], *g["shape"])
    embs, toks = oracle.madeleine_forward_train(sd, x, g["modalities"], stain_encoding=se)
    for m in g["modalities"]:
        assert embs[m].shape == g["embs"][m].shape and toks[m].shape == g["toks"][m].shape
        close(embs[m], g["embs"][m])
        close(toks[m], g["toks"][m])
    if se:
        ev = oracle.madeleine_forward_eval(sd, x[:1, 1:2], g["modalities"], stain_encoding=True, custom_stain_idx=1)
        for k, v in g["eval_custom_stain1"].items():
            close(ev[k], v)


def test_infonce(golden):
    for c in golden("infonce")["cases"]:
        q = c["q"].clone().requires_grad_(True)
        k = c["k"].clone().requires_grad_(True)
        loss = oracle.info_nce(q, k, temperature=c["tau"], symmetric=c["symmetric"])
        loss.backward()
        close(loss, c["loss"], rtol=1e-4, atol=1e-4)
        close(q.grad, c["dq"], rtol=1e-3, atol=1e-2 * float(c["dq"].abs().max()))  # saturated softmax ⇒ cancellation noise
        close(k.grad, c["dk"], rtol=1e-3, atol=1e-2 * float(c["dk"].abs().max()))


def test_infonce_errors():
    with pytest.raises(ValueError):
        oracle.info_nce(torch.zeros(2, 3, 4), torch.zeros(2, 4))
    with pytest.raises(ValueError):
        oracle.info_nce(torch.zeros(2, 4), torch.zeros(3, 4))
    with pytest.raises(ValueError):
        oracle.info_nce(torch.zeros(2, 4), torch.zeros(2, 5))


def test_got_internals(golden):
    g = golden("got")["internals"]
    C = oracle.cosine_cost(g["x"], g["y"])
    close(C, g["cost"])
    close(oracle.thresholded_cosine_cost(g["x"], g["x"]), g["cos_thresh"])
    close(oracle.ipot_plan(C, beta=0.5, iteration=30), g["ipot_T"])
    close(oracle.ipot_distance(C, iteration=30), g["ipot_dist"])
    close(oracle.gromov_wasserstein(g["x"], g["y"]), g["gw"], rtol=1e-3, atol=1e-5)


def test_got(golden):
    for c in golden("got")["cases"]:
        v = c["v"].clone().requires_grad_(True)
        q = c["q"].clone().requires_grad_(True)
        torch.manual_seed(c["torch_seed"])
        loss = oracle.got(v, q, subsample=256)
        loss.backward()
        close(loss, c["loss"], rtol=1e-3, atol=1e-4)
        for got_g, ref_g in ((v.grad, c["dv"]), (q.grad, c["dq"])):
            close(got_g, ref_g, rtol=1e-2, atol=1e-3 * float(ref_g.abs().max()))
        # explicit-permutation entry point gives the same answer
        loss2 = oracle.got(c["v"], c["q"], subsample=256, perm=c["perm"])
        close(loss2, loss.detach(), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("tag", ["global_only", "global_local_se"])
def test_losses_and_grads(golden, tag):
    g = golden("losses_grads")[tag]
    mods = g["modalities"]
    sd = make_state_dict(g["seed_w"], n_mod=len(mods), stain_encoding=g["stain_encoding"])
    assert checksum(sd) == pytest.approx(g["w_checksum"], rel=1e-12)
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    x = make_feats(g["seed_x"], *g["shape"]) * g["labels"][:, :, None, None]
    embs, toks = oracle.madeleine_forward_train(sd, x, mods, stain_encoding=g["stain_encoding"])
    torch.manual_seed(g["torch_seed"])
    loss, flag = oracle.calculate_losses(mods[1:], embs, toks, g["labels"][:, 1:], temperature=0.001, symmetric=True,
                                         use_local=(tag == "global_local_se"))
    assert flag == g["flag"]
    close(loss, g["loss"], rtol=1e-3, atol=1e-3)
    loss.backward()
    for name, d in g["grads"].items():
        gr = sd[name].grad
        gr = torch.zeros_like(sd[name]) if gr is None else gr
        flat = gr.flatten()
        if "full" in d:
            ref = d["full"]
            close(flat, ref, rtol=2e-2, atol=2e-3 * float(ref.abs().max()) + 2e-6)  # attention_c.bias grad is analytically 0 (softmax shift invariance): pure noise
        else:
            ref = d["samples"]
            close(flat[d["idx"]], ref, rtol=2e-2, atol=2e-3 * float(ref.abs().max()) + 2e-6)  # attention_c.bias grad is analytically 0 (softmax shift invariance): pure noise
            assert float(flat.double().norm()) == pytest.approx(float(d["norm"]), rel=2e-2)


# ---------------------------------------------------------------------------------------------- variants (round 1, late)
def test_oracle_other_widths_against_reference(golden):
    """Fixtures from the real reference for patch_embedding_dim = 1024 / 768 / 384 / 1536 / 128 (with and without stain
    encodings): forward(train=True) + InfoNCE + GOT.  Pins the oracle where the GPU tests use it as the yardstick."""
    import oracle
    from weights import make_state_dict, make_feats
    mods = ["HE", "ER", "PR"]
    for (d_in, se), g in golden("variants")["widths"].items():
        sd = {k: v.clone().requires_grad_(True) for k, v in make_state_dict(g["seed_w"], n_mod=3, stain_encoding=se, d_in=d_in).items()}
        feats = make_feats(g["seed_x"], *g["shape"])
        embs, toks = oracle.madeleine_forward_train(sd, feats, mods, stain_encoding=se)
        torch.manual_seed(g["torch_seed"])
        loss, _ = oracle.calculate_losses(mods[1:], embs, toks, torch.ones(g["shape"][0], 2), temperature=g["tau"], symmetric=True,
                                          use_local=True)
        loss.backward()
        for m in mods:
            torch.testing.assert_close(embs[m].detach(), g["embs"][m], rtol=1e-4, atol=1e-5)
            torch.testing.assert_close(toks[m].detach()[:, :2], g["tok_head"][m], rtol=1e-4, atol=1e-5)
            assert float(toks[m].detach().double().sum()) == pytest.approx(float(g["tok_sum"][m]), rel=1e-4, abs=1e-3)
        torch.testing.assert_close(loss.detach(), g["loss"], rtol=1e-4, atol=1e-4)
        for name, ref_norm in g["grad_norms"].items():
            if float(ref_norm) > 1e-6:
                assert float(sd[name].grad.double().norm()) == pytest.approx(float(ref_norm), rel=2e-2), (d_in, se, name)


def test_oracle_activation_variants_against_reference(golden):
    import oracle
    from weights import make_state_dict, make_feats
    for act, g in golden("variants")["activations"].items():
        sd = make_state_dict(g["seed_w"], n_mod=1)
        slide, _, _ = oracle.abmil_embedder(sd, make_feats(g["seed_x"], *g["shape"]), activation=act)
        torch.testing.assert_close(slide, g["slide"], rtol=1e-4, atol=1e-5)


def test_oracle_got_large_problems_against_reference(golden):
    """The reference's GOT on problems of 130 and 256 tokens (what batches of more than 96 cases produce)."""
    import oracle
    for (m, n), g in golden("variants")["got_large"].items():
        gen = torch.Generator().manual_seed(g["seed"])
        v = torch.randn(m, n, 128, generator=gen)
        q = v + 0.5 * torch.randn(m, n, 128, generator=gen)
        vr, qr = v.clone().requires_grad_(True), q.clone().requires_grad_(True)
        loss = oracle.got(vr, qr)
        loss.backward()
        torch.testing.assert_close(loss.detach(), g["loss"], rtol=1e-4, atol=1e-5)
        for got, ref in ((vr.grad, g["dv"]), (qr.grad, g["dq"])):
            assert float((got - ref).norm() / ref.norm()) < 1e-3


def test_oracle_at_baseline_size_against_the_real_reference(golden):
    """BASELINE configs[1] at its full size (32 ragged bags, 68 063 tokens, tau = 0.001): the oracle the GPU tests use as yardstick
    there (tests/parity_utils.py, here in fp32 on the CPU) against the real reference's slide embeddings, loss and gradient norms
    (tests/golden/baseline_sizes.pt, make_golden_baseline_sizes.py).  tau = 0.001 amplifies rounding differences of the embeddings a
    thousandfold in the logits, hence the wider loss / gradient tolerances than for the small fixtures."""
    import parity_utils as pu
    g = golden("baseline_sizes")
    gen = torch.Generator().manual_seed(1234)
    lens = torch.randint(200, 4001, (32,), generator=gen).tolist()
    assert lens == g["lens"]
    cu = [0]
    for n in lens:
        cu.append(cu[-1] + n)
    x = torch.randn(cu[-1], 512, generator=gen)
    assert float(x.double().abs().sum()) == pytest.approx(g["x_checksum"], rel=1e-12)
    sd = pu.to_oracle_sd(make_state_dict(0, n_mod=2), torch.device("cpu"), torch.float32)
    loss, emb = pu.oracle_packed_infonce_step(sd, x, cu, g["tau"])
    close(emb, g["slide"], rtol=1e-4, atol=1e-5)
    assert float(loss) == pytest.approx(float(g["loss"]), rel=1e-4)
    total = sum(float(d["norm"]) ** 2 for d in g["grads"].values()) ** 0.5
    for name, d in g["grads"].items():
        ref = float(d["norm"])
        if ref < 1e-6 * total:      # attention_c.bias: exactly 0 under a softmax over the tokens, ~1e-7 of rounding noise in any fp32 evaluation
            assert float(sd[name].grad.double().norm()) < 1e-6 * total, name
            continue
        assert float(sd[name].grad.double().norm()) == pytest.approx(ref, rel=1e-3), name


def test_oracle_config3_shape_against_the_real_reference(golden):
    """BASELINE configs[2] shape (32 cases x 5 stains, stain encodings, ACROBAT availability rates, InfoNCE tau = 0.001 + GOT) at 512
    tokens per bag — the largest size the reference's monolithic autograd tape allows in the build container — against the real
    reference's forward(train=True) + calculate_losses + backward (tests/golden/baseline_config3_t512.pt)."""
    g = golden("baseline_config3_t512")
    mods = ["HE", "HER2", "PGR", "KI67", "ER"]
    gen = torch.Generator().manual_seed(0)
    labels = (torch.rand(g["bs"], 5, generator=gen) < torch.tensor([1.0, 0.46, 0.73, 0.73, 0.73])).float()
    labels[:, 0] = 1
    assert torch.equal(labels, g["labels"])
    feats = torch.randn(g["bs"], 5, g["T"], 512, generator=gen) * labels[:, :, None, None]
    assert float(feats.double().abs().sum()) == pytest.approx(g["x_checksum"], rel=1e-12)
    sd = {k: v.clone().requires_grad_(True) for k, v in make_state_dict(3, n_mod=5, stain_encoding=True).items()}
    embs, toks = oracle.madeleine_forward_train(sd, feats, mods, stain_encoding=True)
    torch.manual_seed(g["loss_seed"])
    loss, flag = oracle.calculate_losses(mods[1:], embs, toks, labels[:, 1:], temperature=g["tau"], symmetric=True, use_local=True)
    loss.backward()
    assert flag == g["flag"]
    for m in mods:
        close(embs[m].detach(), g["embs"][m], rtol=1e-4, atol=1e-5)
        close(toks[m].detach()[:, :2], g["tok_head"][m], rtol=1e-4, atol=1e-5)
    assert float(loss.detach()) == pytest.approx(float(g["loss"]), rel=1e-4)
    total = sum(float(d["norm"]) ** 2 for d in g["grads"].values()) ** 0.5
    for name, d in g["grads"].items():
        ref = float(d["norm"])
        if ref < 1e-6 * total:
            continue
        assert float(sd[name].grad.double().norm()) == pytest.approx(ref, rel=2e-2), name


def test_oracle_inference_at_4000_tokens_against_the_real_reference(golden):
    """BASELINE configs[4] shape (tests/golden/baseline_inference.pt): `encode_he` of six 4000-patch slides and the raw attention
    logits of two of them from the real reference."""
    g = golden("baseline_inference")
    sd = make_state_dict(g["seed_w"])
    for k, seed in enumerate(g["seeds"]):
        x = make_feats(seed, 1, g["n_tokens"], 512)
        close(oracle.encode_he(sd, x), g["encode_he"][k:k + 1])
    for seed, att in g["attention"].items():
        x = make_feats(seed, 1, 1, g["n_tokens"], 512)
        _, _, raw = oracle.abmil_embedder(sd, x[:, 0])
        close(raw.reshape(att["raw"].shape), att["raw"], rtol=1e-4, atol=2e-5)


def test_oracle_got_at_the_shipped_batch_size_against_the_real_reference(golden):
    """GOT on 65 cases (the reference's shipped batch: problems of 65 tokens) and on 96 (the largest problem of the shared-memory
    kernel), tests/golden/got_shipped_batch.pt."""
    for c in golden("got_shipped_batch")["cases"]:
        m = c["m"]
        v = make_feats(c["seed_v"], m, m + 16, 128).requires_grad_(True)
        q = (v.detach() + 0.5 * make_feats(c["seed_q"], m, m + 16, 128)).requires_grad_(True)
        torch.manual_seed(c["torch_seed"])
        loss = oracle.got(v, q, subsample=256)
        loss.backward()
        close(loss.detach(), c["loss"], rtol=1e-4, atol=1e-5)
        for t, d in ((v.grad, c["dv"]), (q.grad, c["dq"])):
            assert float(t.double().norm()) == pytest.approx(float(d["norm"]), rel=1e-3)
            close(t[:2], d["first2"], rtol=1e-2, atol=1e-3 * float(d["first2"].abs().max()))
            assert float(t[:, m:].abs().max()) == d["beyond_m_abs_max"] == 0.0


@pytest.mark.parametrize("activation", ["softmax", "relu", "leaky_relu", "sigmoid"])
def test_oracle_n_views3_forward_backward_against_the_real_reference(golden, activation):
    """forward(train=True, n_views=3) + global and intra-modality InfoNCE + backward (tests/golden/n_views3_backward.pt)."""
    g = golden("n_views3_backward")
    ga = g["activations"][activation]
    mods = ["HE", "IHC"]
    sd = {k: v.clone().requires_grad_(True) for k, v in make_state_dict(g["seed_w"], n_mod=2).items()}
    feats = make_feats(g["seed_x"], *g["shape"])
    np.random.seed(g["np_seed"])
    embs, toks = oracle.madeleine_forward_train(sd, feats, mods, n_views=3, activation=activation)
    loss, _ = oracle.calculate_losses(mods[1:], embs, toks, torch.ones(g["shape"][0], 1), temperature=g["tau"], symmetric=True, use_intra=True)
    loss.backward()
    for m in mods:
        scale = max(1.0, float(ga["embs"][m].abs().max()))
        close(embs[m].detach(), ga["embs"][m], rtol=1e-4, atol=1e-5 * scale)
    assert float(loss.detach()) == pytest.approx(float(ga["loss"]), rel=1e-5)
    total = sum(float(d["norm"]) ** 2 for d in ga["grads"].values()) ** 0.5
    for name, d in ga["grads"].items():
        if float(d["norm"]) < 1e-6 * total:
            continue
        assert float(sd[name].grad.double().norm()) == pytest.approx(float(d["norm"]), rel=1e-3), name
