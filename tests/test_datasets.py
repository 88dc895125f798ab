"""Input side: the drop-in CPU datasets, the sampler oracle's properties (CPU) and the on-device
resampling kernel against that oracle, bit for bit (GPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import sampler_oracle as so


# ------------------------------------------------------------------------------------------------ CPU: oracle properties
@pytest.mark.parametrize("n", [1, 2, 3, 5, 16, 17, 100, 1000, 2049])
def test_feistel_is_a_permutation(n):
    key = so.bag_key(1234, n)
    image = sorted(so.feistel_perm(i, n, key) for i in range(n))
    assert image == list(range(n))


def test_sampler_rule_matches_sample_n_semantics():
    """wsi_dataset.py:42-50: N >= S -> S distinct rows; N < S -> S rows in range (repeats allowed); missing -> zero bag."""
    lens = [0, 1, 7, 64, 65, 300]
    S = 64
    idx = so.sample_indices(lens, S, seed=99)
    assert (idx[0] == -1).all()
    assert (idx[1] == 0).all()
    assert idx[2].min() >= 0 and idx[2].max() < 7
    for b in (3, 4, 5):
        assert len(set(idx[b].tolist())) == S and idx[b].min() >= 0 and idx[b].max() < lens[b]
    assert not np.array_equal(idx, so.sample_indices(lens, S, seed=100))          # a new seed draws a new sample
    assert not np.array_equal(idx[4], so.sample_indices([65, 65], S, seed=99)[1])    # the key depends on the bag's position


def test_sampler_is_close_to_uniform():
    """Every row of a bag should be picked about equally often over many seeds (randperm(N)[:S] picks each w.p. S/N)."""
    N, S, trials = 40, 10, 600
    counts = np.zeros(N)
    for t in range(trials):
        counts[so.sample_indices([N], S, seed=t * 7919 + 1)[0]] += 1
    expected = trials * S / N
    assert abs(counts.mean() - expected) < 1e-9
    assert counts.min() > 0.7 * expected and counts.max() < 1.3 * expected


# ------------------------------------------------------------------------------------------------ CPU: drop-in datasets
def _write_dataset(tmp_path, n_cases=5, mods=("HE", "ER", "PR"), D=16, seed=0):
    import pandas as pd
    g = torch.Generator().manual_seed(seed)
    rows, bags = [], {}
    for c in range(n_cases):
        row = {"slide_id": f"case{c}", "split": "train"}
        for k, m in enumerate(mods):
            has = 1 if k == 0 or (c + k) % 3 else 0
            row[m] = has
            if has:
                n = int(torch.randint(3, 40, (1,), generator=g))
                x = torch.randn(n, D, generator=g)
                bags[(c, m)] = x
                torch.save(x, tmp_path / f"case{c}_{m}.pt")
        rows.append(row)
    pd.DataFrame(rows).to_csv(tmp_path / "cases.csv", index=False)
    return rows, bags


def test_slide_dataset_and_collate_structure(tmp_path):
    from madeleine.datasets.wsi_dataset import SlideDataset, collate, SimpleDataset, load_features
    mods = ["HE", "ER", "PR"]
    rows, bags = _write_dataset(tmp_path, mods=tuple(mods))
    ds = SlideDataset("toy", str(tmp_path / "cases.csv"), str(tmp_path), mods, embedding_size=16, sample=8, train=True)
    assert len(ds) == 5
    torch.manual_seed(0)
    batch = collate([ds[i] for i in range(len(ds))])
    assert batch["feats"].shape == (5, 3, 8, 16) and batch["modality_labels"].shape == (5, 3)
    assert batch["slide_ids"] == [r["slide_id"] for r in rows]
    for c, r in enumerate(rows):
        for k, m in enumerate(mods):
            got = batch["feats"][c, k]
            if r[m] == 0:
                assert float(got.abs().max()) == 0.0 and batch["modality_labels"][c, k] == 0
            else:
                src = bags[(c, m)]
                # every sampled row is a row of the slide; without replacement when the slide has >= 8 rows
                match = (got[:, None, :] == src[None, :, :]).all(-1)
                assert bool(match.any(1).all())
                if src.shape[0] >= 8:
                    assert len(set(match.float().argmax(1).tolist())) == 8
    simple = SimpleDataset(str(tmp_path))
    feats, sid = simple[0]
    assert torch.equal(feats, load_features(str(tmp_path / (sid + ".pt"))))


def test_resident_store_needs_cuda():
    from madeleine.datasets.wsi_dataset import ResidentSlideStore
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ResidentSlideStore([[torch.zeros(3, 8)]], ["HE"], device="cpu")


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_sample_gather_kernel_bit_exact_against_oracle():
    from madeleine_b200._lib import call, stream_ptr
    dev = torch.device("cuda")
    D, S = 64, 48
    lens = [0, 1, 5, 48, 49, 333, 0, 2000]
    offs, o = [], 0
    for n in lens:
        offs.append(o)
        o += n
    store = torch.randn(o, D, device=dev)
    out = torch.full((len(lens), S, D), 7.0, device=dev)
    idx = torch.empty(len(lens), S, dtype=torch.int32, device=dev)
    seed = (5 << 32) ^ 12345
    call("mdl_sample_gather_f32", store, torch.tensor(offs, dtype=torch.int64, device=dev),
         torch.tensor(lens, dtype=torch.int32, device=dev), len(lens), S, D, seed, out, idx, stream_ptr(dev))
    want = so.sample_indices(lens, S, seed)
    assert np.array_equal(idx.cpu().numpy(), want)                       # index work: bit-exact
    for b, n in enumerate(lens):
        if n == 0:
            assert float(out[b].abs().max()) == 0.0
        else:
            rows = torch.from_numpy(want[b]).long().to(dev) + offs[b]
            assert torch.equal(out[b], store[rows])


@pytest.mark.gpu
def test_resident_loader_feeds_a_training_step(tmp_path):
    """Batches built on the device have the reference collate's structure and drive forward + losses + backward."""
    from argparse import Namespace
    from madeleine.datasets.wsi_dataset import SlideDataset, ResidentSlideStore, ResidentLoader
    from madeleine.models.Model import MADELEINE
    from madeleine.utils.loss import InfoNCE, GOT
    from madeleine.utils.trainer import calculate_losses
    mods = ["HE", "ER", "PR"]
    rows, bags = _write_dataset(tmp_path, n_cases=6, mods=tuple(mods), D=512, seed=3)
    ds = SlideDataset("toy", str(tmp_path / "cases.csv"), str(tmp_path), mods, embedding_size=512, sample=32, train=True)
    store = ResidentSlideStore.from_dataset(ds, device="cuda")
    assert len(store) == 6 and store.nbytes == sum(b.numel() for b in bags.values()) * 4
    loader = ResidentLoader(store, batch_size=4, sample=32, shuffle=True, seed=1)
    batches = list(loader)
    assert [b["feats"].shape[0] for b in batches] == [4, 2]
    b0 = store.sample_batch([0, 1, 2, 3], 32, seed=9, return_indices=True)
    assert b0["feats"].shape == (4, 3, 32, 512) and b0["feats"].is_cuda and not b0["modality_labels"].is_cuda
    for c in range(4):
        for k, m in enumerate(mods):
            if rows[c][m] == 0:
                assert float(b0["feats"][c, k].abs().max()) == 0.0 and b0["modality_labels"][c, k] == 0
            else:
                src = bags[(c, m)].cuda()
                assert torch.equal(b0["feats"][c, k], src[b0["indices"][c, k].long()])
    cfg = Namespace(MODALITIES=mods, wsi_encoder="abmil", patch_embedding_dim=512, wsi_encoder_hidden_dim=512,
                    activation="softmax", n_heads=4, b200_precision="fp32", b200_token_window="batch")
    torch.manual_seed(0)
    model = MADELEINE(cfg, stain_encoding=True).cuda().train()
    args = Namespace(global_loss="info-nce", symmetric_cl=True, local_loss_weight=1.0)
    batch = batches[0]
    embs, toks = model(batch, device=torch.device("cuda"), n_views=1)
    loss, flag = calculate_losses(mods[1:], InfoNCE(temperature=0.1), GOT, None, embs, toks, batch["modality_labels"][:, 1:], args)
    assert flag
    loss.backward()
    assert all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)


def test_load_features_reads_the_reference_hdf5_layout(tmp_path):
    """The reference stores patch embeddings as an HDF5 dataset ``features`` [1, N, D] or [N, D] fp32
    (madeleine/preprocessing/conch_patch_embedder.py:127-131, read by wsi_dataset.py:14-19).  Needs h5py, which this image
    does not ship: skipped there, exercised wherever the reference's own loader can run."""
    h5py = pytest.importorskip("h5py")
    import numpy as np
    from madeleine_b200.datasets.wsi_dataset import load_features
    arr = np.random.default_rng(0).standard_normal((1, 37, 512)).astype(np.float32)
    path = tmp_path / "slide.h5"
    with h5py.File(path, "w") as f:
        f.create_dataset("features", data=arr)
        f.create_dataset("coords", data=np.zeros((37, 2), dtype=np.int64))
    out = load_features(str(path))
    assert out.dtype == torch.float32 and tuple(out.shape) == (37, 512)
    assert torch.equal(out, torch.from_numpy(arr[0]))
