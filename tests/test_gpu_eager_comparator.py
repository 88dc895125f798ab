"""Same-box comparator: the reference's algorithm as plain PyTorch-eager ops ON THE B200 (the oracle's torch
restatement with its tensors moved to the GPU — what running the reference itself on this GPU executes: cuBLAS GEMMs plus
dozens of elementwise kernels per layer) against the CUDA path, on the bench step (16 cases x 2 stains x 2000 x 512, fwd+bwd,
symmetric InfoNCE).  The oracle is used here as a yardstick inside tests/, never by the product.  Prints one JSON line
(captured into profiles/ per round) and checks that both paths compute the same loss."""
import json
from argparse import Namespace

import pytest
import torch

pytestmark = pytest.mark.gpu

from weights import make_state_dict  # noqa: E402

DEV = torch.device("cuda")
MODS = ["HE", "IHC"]


def _time(fn, n=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


def test_bench_step_against_pytorch_eager_on_the_same_gpu(capsys):
    import oracle
    from madeleine.models.Model import MADELEINE
    from madeleine.utils.loss import InfoNCE
    from madeleine.utils.trainer import calculate_losses
    sd_cpu = make_state_dict(0, n_mod=2)
    feats = torch.randn(16, 2, 2000, 512, generator=torch.Generator().manual_seed(3)).to(DEV)
    labels = torch.ones(16, 2)
    tau = 0.1
    sd = {k: v.to(DEV).requires_grad_(True) for k, v in sd_cpu.items()}

    def eager_step():
        for v in sd.values():
            v.grad = None
        embs, toks = oracle.madeleine_forward_train(sd, feats, MODS)
        loss, _ = oracle.calculate_losses(MODS[1:], embs, toks, labels[:, 1:], temperature=tau, symmetric=True)
        loss.backward()
        return loss.detach()

    torch.backends.cuda.matmul.allow_tf32 = False
    ms_eager_fp32, loss_eager = _time(eager_step)

    def eager_bf16_step():
        with torch.amp.autocast("cuda", dtype=torch.bfloat16):
            return eager_step()
    ms_eager_bf16, _ = _time(eager_bf16_step)

    res = {}
    losses = {}
    for prec in ("fp32", "bf16"):
        cfg = Namespace(MODALITIES=MODS, wsi_encoder="abmil", patch_embedding_dim=512, wsi_encoder_hidden_dim=512,
                        activation="softmax", n_heads=4, b200_precision=prec)
        model = MADELEINE(cfg, stain_encoding=False)
        model.load_state_dict(sd_cpu, strict=True)
        model.to(DEV).eval()                   # dropout off on both sides so that the losses are comparable
        args = Namespace(global_loss="info-nce", symmetric_cl=True, local_loss_weight=1.0)
        loss_fn = InfoNCE(temperature=tau)

        def step():
            model.zero_grad(set_to_none=True)
            embs, toks = model({"feats": feats}, device=DEV, n_views=1)
            loss, _ = calculate_losses(MODS[1:], loss_fn, None, None, embs, toks, labels[:, 1:], args)
            loss.backward()
            return loss.detach()
        res[prec], losses[prec] = _time(step, n=10)

    torch.testing.assert_close(losses["fp32"], loss_eager, rtol=1e-3, atol=1e-4)
    line = {"workload": "bench step: 16 cases x 2 stains x 2000 x 512, fwd+bwd, symmetric InfoNCE (eval mode, tau 0.1)",
            "pytorch_eager_fp32_ms": round(ms_eager_fp32, 2), "pytorch_eager_bf16_autocast_ms": round(ms_eager_bf16, 2),
            "madeleine_b200_fp32_grade_ms": round(res["fp32"], 2), "madeleine_b200_bf16_ms": round(res["bf16"], 2),
            "speedup_fp32": round(ms_eager_fp32 / res["fp32"], 1), "speedup_bf16": round(ms_eager_bf16 / res["bf16"], 1),
            "slides_per_s": {"eager_fp32": round(32e3 / ms_eager_fp32), "eager_bf16": round(32e3 / ms_eager_bf16),
                             "ours_fp32": round(32e3 / res["fp32"]), "ours_bf16": round(32e3 / res["bf16"])}}
    with capsys.disabled():
        print("\nEAGER_COMPARATOR " + json.dumps(line))
    assert res["fp32"] < ms_eager_fp32 and res["bf16"] < ms_eager_bf16


def test_configs2_step_against_pytorch_eager_on_the_same_gpu(capsys):
    """BASELINE configs[2]: 32 cases x 5 stains x 2048 tokens, stain encodings, ACROBAT availability, InfoNCE + Graph-OT —
    the regime where eager PyTorch is launch-bound (the reference's GOT alone is ~10^3 kernels per stain plus their autograd)."""
    import oracle
    from madeleine.models.Model import MADELEINE
    from madeleine.utils.loss import InfoNCE, GOT
    from madeleine.utils.trainer import calculate_losses
    mods = ["HE", "HER2", "PGR", "KI67", "ER"]
    bs, T = 32, 2048
    sd_cpu = make_state_dict(3, n_mod=5, stain_encoding=True)
    g = torch.Generator().manual_seed(0)
    labels = (torch.rand(bs, 5, generator=g) < torch.tensor([1.0, 0.46, 0.73, 0.73, 0.73])).float()
    labels[:, 0] = 1
    feats = torch.randn(bs, 5, T, 512, generator=g).to(DEV) * labels.to(DEV)[:, :, None, None]
    tau = 0.1
    sd = {k: v.to(DEV).requires_grad_(True) for k, v in sd_cpu.items()}

    def eager_step():
        for v in sd.values():
            v.grad = None
        embs, toks = oracle.madeleine_forward_train(sd, feats, mods, stain_encoding=True)
        torch.manual_seed(11)
        loss, _ = oracle.calculate_losses(mods[1:], embs, toks, labels[:, 1:], temperature=tau, symmetric=True, use_local=True)
        loss.backward()
        return loss.detach()

    torch.backends.cuda.matmul.allow_tf32 = False
    ms_eager, loss_eager = _time(eager_step, n=2)
    peak_eager = torch.cuda.max_memory_allocated() / 1e9
    del sd
    torch.cuda.empty_cache()

    res, losses = {}, {}
    for window in ("off", "batch"):
        cfg = Namespace(MODALITIES=mods, wsi_encoder="abmil", patch_embedding_dim=512, wsi_encoder_hidden_dim=512,
                        activation="softmax", n_heads=4, b200_precision="fp32", b200_token_window=window)
        model = MADELEINE(cfg, stain_encoding=True)
        model.load_state_dict(sd_cpu, strict=True)
        model.to(DEV).eval()
        args = Namespace(global_loss="info-nce", symmetric_cl=True, local_loss_weight=1.0)
        loss_fn = InfoNCE(temperature=tau)

        def step():
            model.zero_grad(set_to_none=True)
            embs, toks = model({"feats": feats, "modality_labels": labels}, device=DEV, n_views=1)
            torch.manual_seed(11)
            loss, _ = calculate_losses(mods[1:], loss_fn, GOT, None, embs, toks, labels[:, 1:], args)
            loss.backward()
            return loss.detach()
        res[window], losses[window] = _time(step, n=5)
    for window in res:
        torch.testing.assert_close(losses[window], loss_eager, rtol=1e-3, atol=1e-3)
    line = {"workload": "configs[2]: 32 cases x 5 stains x 2048 x 512, stain encodings, InfoNCE + GOT, fwd+bwd (eval mode, tau 0.1)",
            "pytorch_eager_fp32_ms": round(ms_eager, 1), "pytorch_eager_peak_mem_gb": round(peak_eager, 1),
            "madeleine_b200_fp32_grade_ms": round(res["off"], 1), "madeleine_b200_fp32_grade_token_window_ms": round(res["batch"], 1),
            "speedup": round(ms_eager / res["off"], 1), "speedup_token_window": round(ms_eager / res["batch"], 1),
            "cases_per_s": {"eager": round(bs * 1e3 / ms_eager, 1), "ours": round(bs * 1e3 / res["off"], 1),
                            "ours_token_window": round(bs * 1e3 / res["batch"], 1)}}
    with capsys.disabled():
        print("\nEAGER_COMPARATOR " + json.dumps(line))
    assert res["off"] < ms_eager
