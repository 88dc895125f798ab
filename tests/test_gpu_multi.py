"""Multi-GPU equivalence (needs >= 2 CUDA devices; skipped on a single-GPU box): sharded InfoNCE + GOT == single process."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_losses_match_single_process():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(REPO, "tests", "multi_gpu", "check_sharded_losses.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("[PASS]") == 4


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_nn_dataparallel_wrapping_like_the_reference_script():
    """madeleine/utils/setup_components.py:185-187 wraps the model in nn.DataParallel when several GPUs are visible and calls
    it with device=torch.device("cuda") (trainer.py:14,111).  The drop-in must survive that: replicas on both GPUs, outputs
    gathered on GPU 0, loss and parameter gradients equal to the single-GPU run."""
    import sys
    from argparse import Namespace
    sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
    from madeleine.models.Model import MADELEINE
    from madeleine.utils.loss import InfoNCE, GOT
    from madeleine.utils.trainer import calculate_losses
    from weights import make_state_dict, make_feats
    mods = ["HE", "ER", "PR"]
    cfg = Namespace(MODALITIES=mods, wsi_encoder="abmil", patch_embedding_dim=512, wsi_encoder_hidden_dim=512,
                    activation="softmax", n_heads=4, b200_precision="fp32")
    sd = make_state_dict(6, n_mod=3)
    feats = make_feats(2, 8, 3, 40, 512)
    labels = torch.ones(8, 3)
    args = Namespace(global_loss="info-nce", symmetric_cl=True, local_loss_weight=1.0)
    dev = torch.device("cuda")

    def run(wrap):
        model = MADELEINE(cfg, stain_encoding=False)
        model.load_state_dict(sd, strict=True)
        model.to(dev).eval()
        net = torch.nn.DataParallel(model) if wrap else model
        embs, toks = net({"feats": feats, "modality_labels": labels}, device=dev, n_views=1)
        torch.manual_seed(4)
        loss, _ = calculate_losses(mods[1:], InfoNCE(temperature=0.1), GOT, None, embs, toks, labels[:, 1:], args)
        loss.backward()
        return loss.detach().cpu(), {n: p.grad.detach().cpu() for n, p in model.named_parameters()}, embs

    l1, g1, e1 = run(False)
    l2, g2, e2 = run(True)
    assert e2["ER"].device.index == 0 and e2["ER"].shape == e1["ER"].shape
    torch.testing.assert_close(l2, l1, rtol=1e-5, atol=1e-5)
    for n in g1:
        if float(g1[n].norm()) > 1e-5:
            assert float((g2[n] - g1[n]).norm() / g1[n].norm()) < 1e-3, n
