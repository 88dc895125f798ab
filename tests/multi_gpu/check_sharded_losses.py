"""2-GPU check (run under torchrun on a multi-GPU box; `gpurun --gpus 2`):  cases sharded over ranks, one all-gather of
slide embeddings, GOT with all-reduced extrema / threshold sums  ==  the single-process loss and gradients on the full
batch.  Prints PASS/FAIL lines; exit code 1 on failure."""
import os
import sys
from argparse import Namespace

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
from madeleine.models.Model import MADELEINE  # noqa: E402
from madeleine.utils.loss import InfoNCE, GOT  # noqa: E402
from madeleine.utils.trainer import calculate_losses  # noqa: E402
from madeleine_b200 import parallel  # noqa: E402
from weights import make_state_dict, make_feats  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    mods = ["HE", "HER2", "PGR"]
    cfg = Namespace(MODALITIES=mods, wsi_encoder="abmil", patch_embedding_dim=512, wsi_encoder_hidden_dim=512,
                    activation="softmax", n_heads=4, b200_precision="fp32", b200_skip_missing_bags=False)
    B, T = 4 * world, 64
    feats = make_feats(21, B, 3, T, 512)
    g = torch.Generator().manual_seed(5)
    labels = (torch.rand(B, 3, generator=g) < 0.8).float()
    labels[:, 0] = 1
    labels[:2, :] = 1
    feats = feats * labels[:, :, None, None]
    args = Namespace(global_loss="info-nce", symmetric_cl=True, local_loss_weight=1.0)
    sd = make_state_dict(4, n_mod=3, stain_encoding=False)

    def fresh(token_window="off"):
        c = Namespace(**vars(cfg))
        c.b200_token_window = token_window
        m = MADELEINE(c, stain_encoding=False)
        m.load_state_dict(sd)
        return m.to(dev).eval()

    # ---- single-process reference on the full batch (every rank computes it on its own GPU)
    ref = fresh()
    parallel.enable_gradient_sync(False)
    embs, toks = ref({"feats": feats}, dev, train=True)
    torch.manual_seed(99)
    loss_ref, _ = calculate_losses(mods[1:], InfoNCE(temperature=0.1), GOT, None, embs, toks, labels[:, 1:], args)
    loss_ref.backward()
    g_ref = {n: p.grad.clone() for n, p in ref.named_parameters()}

    ok = True
    # the token window 'batch' spans the GLOBAL batch (local batch x world size): GOT's permutation is over all cases
    for token_window in ("off", "batch"):
        # ---- sharded: rank r owns cases [r*4, r*4+4)
        model = fresh(token_window)
        parallel.enable_gradient_sync(True)
        bl = B // world
        local = feats[rank * bl:(rank + 1) * bl]
        embs_l, toks_l = model({"feats": local}, dev, train=True)
        embs_g, _ = parallel.gather_slide_embeddings(embs_l, labels[rank * bl:(rank + 1) * bl].to(dev), global_labels_host=labels)
        torch.manual_seed(99)
        loss_sh, flag, parts = parallel.calculate_losses_sharded(mods[1:], InfoNCE(temperature=0.1), True, embs_g, toks_l, labels[:, 1:],
                                                                 args, rank, bl)
        loss_sh.backward()                      # encoder backward all-reduces (SUM) the flat gradient buffer
        local = parts["local"].detach().clone()
        dist.all_reduce(local)
        total = parts["global"].detach() + local
        err = abs(float(total) - float(loss_ref)) / max(1.0, abs(float(loss_ref)))
        ok &= err < 1e-4
        if rank == 0:
            print(f"[{'PASS' if err < 1e-4 else 'FAIL'}] (token window {token_window}) loss sharded(sum over ranks)={float(total):.6f} single={float(loss_ref):.6f}")
        worst = 0.0
        for n, p in model.named_parameters():
            if p.grad is None:
                continue
            d = float((p.grad - g_ref[n]).norm() / (g_ref[n].norm() + 1e-12))
            if float(g_ref[n].norm()) > 1e-5:
                worst = max(worst, d)
        ok &= worst < 2e-2
        if rank == 0:
            print(f"[{'PASS' if worst < 2e-2 else 'FAIL'}] (token window {token_window}) worst relative gradient difference over parameters = {worst:.3e}")
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
