"""Host-side cost of one training step: run the full public-API step on a tiny problem (GPU time ~ 0) and on the bench
problem, timing wall-clock per step of the launch phase (no syncs inside the loop)."""
import os, sys, time, cProfile, pstats, io
from argparse import Namespace
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import torch
from madeleine.models.Model import MADELEINE
from madeleine.utils.loss import InfoNCE
from madeleine.utils.trainer import calculate_losses
from weights import make_state_dict

dev = torch.device("cuda")
MODS = ["HE", "IHC"]
cfg = Namespace(MODALITIES=MODS, wsi_encoder="abmil", patch_embedding_dim=512, wsi_encoder_hidden_dim=512, activation="softmax", n_heads=4, b200_precision="fp32")
model = MADELEINE(cfg, stain_encoding=False)
model.load_state_dict(make_state_dict(0, n_mod=2))
model.to(dev).train()
loss_fn = InfoNCE(temperature=0.001)
largs = Namespace(global_loss="info-nce", symmetric_cl=True, local_loss_weight=1.0)
labels = torch.ones(16, 2)


def step(feats):
    model.zero_grad(set_to_none=True)
    embs, toks = model({"feats": feats}, device=dev, n_views=1)
    loss, ok = calculate_losses(MODS[1:], loss_fn, None, None, embs, toks, labels[:, 1:], largs)
    loss.backward()
    return loss


for T in (8, 2000):
    feats = torch.randn(16, 2, T, 512, device=dev)
    for _ in range(5):
        step(feats)
    torch.cuda.synchronize()
    n = 30
    t0 = time.perf_counter()
    for _ in range(n):
        step(feats)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"T={T}: host launch phase {1e3*(t1-t0)/n:.3f} ms/step, total incl. drain {1e3*(t2-t0)/n:.3f} ms/step")

feats = torch.randn(16, 2, 8, 512, device=dev)
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step(feats)
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
print(s.getvalue()[:6000])
