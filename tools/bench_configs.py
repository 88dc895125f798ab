#!/usr/bin/env python
"""Timings of the other BASELINE.json configurations on one B200 (JSON lines; copied to profiles/ per round).

  configs[2]: batch=32 cases, 5 stains (ACROBAT shape), T=2048, stain encodings, global InfoNCE + local GOT, fwd+bwd
  configs[4]: inference, synthetic slides of N=4000 x 512 through the extraction driver (host -> device -> host)
"""
import argparse
import json
import os
import sys
import time
from argparse import Namespace

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests", "golden"))

from madeleine.models.Model import MADELEINE  # noqa: E402
from madeleine.utils.loss import InfoNCE, GOT  # noqa: E402
from madeleine.utils.trainer import calculate_losses  # noqa: E402
from madeleine_b200 import _lib  # noqa: E402
from madeleine_b200.utils.inference import extract_slide_embeddings  # noqa: E402
from weights import make_state_dict  # noqa: E402


def cfg(mods, precision, token_window="off"):
    return Namespace(MODALITIES=mods, wsi_encoder="abmil", patch_embedding_dim=512, wsi_encoder_hidden_dim=512,
                     activation="softmax", n_heads=4, b200_precision=precision, b200_token_window=token_window)


def config3(precision, steps, warmup, skip=True, bs=32, tag="BASELINE configs[2]: batch=32", breakdown=False, token_window="off"):
    dev = torch.device("cuda")
    mods = ["HE", "HER2", "PGR", "KI67", "ER"]
    model = MADELEINE(cfg(mods, precision, token_window), stain_encoding=True)
    model.load_state_dict(make_state_dict(3, n_mod=5, stain_encoding=True))
    model.to(dev).train()
    T = 2048
    g = torch.Generator().manual_seed(0)
    labels = (torch.rand(bs, 5, generator=g) < torch.tensor([1.0, 0.46, 0.73, 0.73, 0.73])).float()
    labels[:, 0] = 1
    feats = torch.randn(bs, 5, T, 512, device=dev) * labels.to(dev)[:, :, None, None]
    args = Namespace(global_loss="info-nce", symmetric_cl=True, local_loss_weight=1.0)
    loss_fn = InfoNCE(temperature=0.001)

    def step():
        model.zero_grad(set_to_none=True)
        # the reference's collate delivers the availability mask with every batch (data['modality_labels'])
        embs, toks = model({"feats": feats, "modality_labels": labels} if skip else {"feats": feats}, device=dev, n_views=1)
        loss, ok = calculate_losses(mods[1:], loss_fn, GOT, None, embs, toks, labels[:, 1:], args)
        loss.backward()
        return loss

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    _lib.start_timing("all" if breakdown else {"mdl_got_fwd_bwd", "mdl_got_extrema", "mdl_infonce_fwd", "mdl_infonce_bwd"})
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    kt = {k: round(sum(v) / steps, 4) for k, v in _lib.stop_timing().items()}
    bags = bs * 5
    print(json.dumps({"config": tag + ", 5 stains, T=2048, stain encodings, InfoNCE + GOT, fwd+bwd, train mode",
                      "missing_bags_encoded_from_one_token": skip, "token_window": token_window, "missing_bag_fraction": float(1 - labels.mean()),
                      "precision": precision, "ms_per_step": ms, "slides_per_s": bags / (ms * 1e-3), "cases_per_s": bs / (ms * 1e-3),
                      "loss": float(loss), "cases_per_stain": labels[:, 1:].sum(0).tolist(), "loss_kernel_ms_per_step": kt,
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}))


def config5(precision, n_slides, n_tokens):
    """BASELINE configs[4].  Under torchrun every rank is a replica that extracts its stride of the slides (no collective on
    the data path); the job's throughput is n_slides over the slowest rank's time."""
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    model = MADELEINE(cfg(["HE"], precision), stain_encoding=False)
    model.load_state_dict(make_state_dict(0))
    model.to(dev).eval()
    g = torch.Generator().manual_seed(1)
    base = torch.randn(n_tokens + 64, 512, generator=g)
    bags = [base[(i % 64):(i % 64) + n_tokens] for i in range(n_slides)]   # distinct views, no 8 GB of host RAM

    def timed(inputs):
        extract_slide_embeddings(model, inputs[:96], dev)                  # warm-up: three batches, both staging slots
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        emb, idx = extract_slide_embeddings(model, inputs, dev, rank=rank, world=world)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return float(dt), emb

    dt, emb = timed(bags)
    # the same slides delivered in pinned memory (DataLoader(pin_memory=True)): no host-side staging copy
    pbase = base.pin_memory()
    pbags = [pbase[(i % 64):(i % 64) + n_tokens] for i in range(n_slides)]
    dt_pinned, emb_p = timed(pbags)
    assert abs(float(abs(emb_p).sum()) - float(abs(emb).sum())) < 1e-3 * float(abs(emb).sum())
    # device-resident forward only (no H2D), same packing
    x = torch.randn(32 * n_tokens, 512, device=dev)
    cu = torch.arange(0, 33 * n_tokens, n_tokens, dtype=torch.int32, device=dev)
    with torch.no_grad():
        for _ in range(3):
            model.encode_packed(x, cu)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            model.encode_packed(x, cu)
        e1.record()
        torch.cuda.synchronize()
    dev_ms = torch.tensor([e0.elapsed_time(e1) / 10], device=dev)
    if world > 1:
        dist.all_reduce(dev_ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"config": f"BASELINE configs[4]: inference, {n_slides} slides x {n_tokens} x 512, extraction driver (host->device->host)",
                          "n_gpus": world, "replicas": "rank-strided slides, no data-path collective", "precision": precision,
                          "e2e_slides_per_s": n_slides / dt, "e2e_seconds": dt,
                          "e2e_slides_per_s_pinned_inputs": n_slides / dt_pinned,
                          "device_resident_slides_per_s": world * 32 / (float(dev_ms) * 1e-3), "h2d_bytes_per_slide": n_tokens * 512 * 4,
                          "embedding_checksum_rank0": float(abs(emb).sum())}))
    if world > 1:
        dist.destroy_process_group()


def resident_loader(precision, steps, bs=65, n_cases=130, token_window="batch"):
    """The reference's canonical step fed three ways: features already in HBM (static batch), the HBM-resident store with
    on-device resampling (a NEW sample of every bag every step), and the reference's way — a pinned host batch copied in
    every step (prefetched one batch ahead)."""
    from madeleine.datasets.wsi_dataset import ResidentSlideStore, ResidentLoader
    from madeleine_b200.utils.prefetch import DevicePrefetcher
    dev = torch.device("cuda")
    mods = ["HE", "HER2", "PGR", "KI67", "ER"]
    T = 2048
    g = torch.Generator().manual_seed(0)
    avail = torch.rand(n_cases, 5, generator=g) < torch.tensor([1.0, 0.46, 0.73, 0.73, 0.73])
    avail[:, 0] = True
    base = torch.randn(9000, 512, generator=g)
    cases = []
    for c in range(n_cases):
        row = []
        for k in range(5):
            if not avail[c, k]:
                row.append(None)
                continue
            n = int(torch.randint(1500, 8000, (1,), generator=g))          # slides have a few thousand patches
            o = int(torch.randint(0, 9000 - n, (1,), generator=g))
            row.append(base[o:o + n])
        cases.append(row)
    store = ResidentSlideStore(cases, mods, device=dev)
    model = MADELEINE(cfg(mods, precision, token_window), stain_encoding=True)
    model.load_state_dict(make_state_dict(3, n_mod=5, stain_encoding=True))
    model.to(dev).train()
    args = Namespace(global_loss="info-nce", symmetric_cl=True, local_loss_weight=1.0)
    loss_fn = InfoNCE(temperature=0.001)

    def step(batch):
        model.zero_grad(set_to_none=True)
        embs, toks = model(batch, device=dev, n_views=1)
        loss, ok = calculate_losses(mods[1:], loss_fn, GOT, None, embs, toks, batch["modality_labels"][:, 1:], args)
        loss.backward()
        return loss

    def timed(batches, n):
        it = iter(batches)
        for _ in range(2):
            step(next(it))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            step(next(it))
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    def cycle(make):
        while True:
            yield from make()

    static = store.sample_batch(list(range(bs)), T, seed=1)
    ms_static = timed(cycle(lambda: [static]), steps)
    counter = [0]

    def resampled():                      # the SAME cases as the static batch (same work), a new sample of every bag every step
        counter[0] += 1
        return [store.sample_batch(list(range(bs)), T, seed=1000 + counter[0])]
    ms_resident = timed(cycle(resampled), steps)
    loader = ResidentLoader(store, batch_size=bs, sample=T, shuffle=True, drop_last=True, seed=0)
    ms_loader = timed(cycle(lambda: loader), steps)
    host = [{"feats": static["feats"].cpu().pin_memory(), "modality_labels": static["modality_labels"]} for _ in range(2)]
    ms_host = timed(DevicePrefetcher(cycle(lambda: host), dev), steps)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(5):
        store.sample_batch(list(range(bs)), T, seed=100 + i)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"config": f"reference canonical config fed from an HBM-resident store: {n_cases} cases x 5 stains, "
                                f"{store.nbytes / 1e9:.1f} GB of features resident, batch {bs}, sample {T}, token window {token_window}",
                      "precision": precision, "ms_per_step_static_device_batch": round(ms_static, 3),
                      "ms_per_step_resident_store_resampled_every_step": round(ms_resident, 3),
                      "ms_per_step_resident_loader_shuffled_cases": round(ms_loader, 3),
                      "ms_per_step_pinned_host_batch_prefetched": round(ms_host, 3),
                      "sample_gather_kernel_ms": round(e0.elapsed_time(e1) / 5, 3),
                      "h2d_bytes_per_step_host_path": static["feats"].numel() * 4, "h2d_bytes_per_step_resident": 0,
                      "cases_per_s_resident": round(bs / (ms_resident * 1e-3), 1)}))


def config2_ragged(precision, steps):
    """BASELINE configs[1] as written: 16 cases x 2 stains, N_i ~ U[200, 4000] per bag (generator seed 1234), symmetric
    InfoNCE.  The reference can only batch equal-length bags (it loops bs = 1 or pads); here the 32 bags are packed
    back to back ([sum N_i, 512] + cu_seqlens) and go through forward_packed in one pass, no padding."""
    dev = torch.device("cuda")
    mods = ["HE", "IHC"]
    model = MADELEINE(cfg(mods, precision), stain_encoding=False)
    model.load_state_dict(make_state_dict(0, n_mod=2))
    model.to(dev).train()
    g = torch.Generator().manual_seed(1234)
    lens = torch.randint(200, 4001, (32,), generator=g)
    cu = torch.zeros(33, dtype=torch.int32)
    cu[1:] = lens.cumsum(0)
    total = int(cu[-1])
    x = torch.randn(total, 512, device=dev)
    cu_dev = cu.to(dev)
    loss_fn = InfoNCE(temperature=0.001)

    def step():
        model.zero_grad(set_to_none=True)
        slide, _ = model.forward_packed(x, cu_dev, want_tokens=False)      # bags 2c / 2c+1 = HE / IHC slide of case c
        loss = loss_fn(query=slide[0::2], positive_key=slide[1::2], symmetric=True)
        loss.backward()
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({"config": "BASELINE configs[1], ragged: 16 cases x 2 stains, N_i ~ U[200,4000] (seed 1234), packed, symmetric InfoNCE, fwd+bwd",
                      "precision": precision, "tokens": total, "max_len": int(lens.max()), "ms_per_step": round(ms, 3),
                      "slides_per_s": round(32 / (ms * 1e-3), 1), "tokens_per_s": round(total / (ms * 1e-3)),
                      "padded_tokens_if_batched_dense": int(lens.max()) * 32, "loss": float(loss.detach())}))


def got_sizes():
    """Graph-OT loss alone (forward + token gradients) for a range of problem counts / sizes, incl. the n > 96 path."""
    from madeleine_b200 import ops
    dev = torch.device("cuda")
    for m, n in ((53, 53), (65, 65), (96, 96), (128, 128), (192, 192), (256, 256)):
        g = torch.Generator().manual_seed(n)
        v = torch.randn(m, n, 128, generator=g).to(dev)
        q = (v + 0.5 * torch.randn(m, n, 128, device=dev))
        for _ in range(2):
            ops.got_loss(v, q)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            loss = ops.got_loss(v, q)
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"config": "GOT alone (one stain): m problems of n tokens, forward + gradients w.r.t. the tokens",
                          "m": m, "n": n, "ms": round(e0.elapsed_time(e1) / 3, 3), "loss": float(loss),
                          "kernels": "got.cu (shared memory)" if n <= 96 else "got_big.cu (global memory)"}))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="3,5")
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--slides", type=int, default=500)
    ap.add_argument("--token-window", default="off", help="'off' | 'batch' | int (config.b200_token_window)")
    ap.add_argument("--breakdown", action="store_true", help="time every C-ABI entry point (events around each call)")
    a = ap.parse_args()
    if "3" in a.which:
        config3(a.precision, a.steps, 2, skip=True, token_window=a.token_window, breakdown=a.breakdown)
        config3(a.precision, a.steps, 2, skip=False, token_window=a.token_window)
    if "canon" in a.which:
        # the reference's shipped pre-training configuration (scripts/launch_pretrain_withStainEncodings.sh): batch 65
        config3(a.precision, a.steps, 2, skip=True, bs=65, tag="reference canonical config: batch=65", breakdown=a.breakdown,
                token_window=a.token_window)
    if "ragged" in a.which:
        config2_ragged(a.precision, a.steps)
    if "got" in a.which:
        got_sizes()
    if "resident" in a.which:
        resident_loader(a.precision, a.steps, token_window=a.token_window if a.token_window != "off" else "batch")
    if "5" in a.which:
        config5(a.precision, a.slides, 4000)
