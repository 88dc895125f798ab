"""Drop-in for madeleine/datasets/wsi_dataset.py plus the B200-native alternative to it.

Reference behaviour kept (wsi_dataset.py:14-115): ``load_features`` (HDF5 dataset ``features``, squeezed, fp32 tensor),
``SlideDataset`` (one case per item: per modality either the slide's features or a 2-row zero bag when the stain is
missing, each resampled to ``sample`` tokens), ``collate`` (stack to ``feats [bs, n_mod, sample, D]``,
``modality_labels [bs, n_mod]``, ``slide_ids``) and ``SimpleDataset`` (one slide per item for extraction).

``ResidentSlideStore`` / ``ResidentLoader`` replace the per-step HDF5 read + CPU resampling + 4 MB-per-bag H2D copy:
every slide's features are uploaded ONCE (a pre-training set of a few thousand slides is tens of GB; one B200 has 180 GB),
and each step one kernel (``mdl_sample_gather_f32``) draws the sample and gathers the rows straight into the
``[bs, n_mod, sample, D]`` batch on the device.  The batches have the reference's structure, so ``MADELEINE.forward`` and
``calculate_losses`` consume them unchanged.
"""
from __future__ import annotations

import os
from typing import Callable, Dict, Iterator, List, Optional, Sequence

import numpy as np
import torch
from torch.utils.data import Dataset

from .._lib import call, stream_ptr


def load_features(path: str) -> torch.Tensor:
    """wsi_dataset.py:14-19.  ``.h5``: dataset ``features`` (needs h5py); also ``.pt`` / ``.npy`` for hosts without HDF5."""
    ext = os.path.splitext(path)[1].lower()
    if ext == ".pt":
        feats = torch.load(path, weights_only=True)
    elif ext == ".npy":
        feats = np.load(path)
    else:
        try:
            import h5py
        except ImportError as e:  # pragma: no cover - depends on the host image
            raise ImportError(f"reading {path} needs h5py (the reference's feature format); "
                              "convert to .pt/.npy or install h5py") from e
        with h5py.File(path, "r") as f:
            feats = f["features"][:]
    if isinstance(feats, np.ndarray):
        feats = torch.from_numpy(np.ascontiguousarray(feats))
    feats = feats.squeeze().to(torch.float32)
    return feats.reshape(1, -1) if feats.dim() == 1 else feats


def _feature_file(features_path: str, stem: str) -> str:
    for ext in (".h5", ".pt", ".npy"):
        p = os.path.join(features_path, stem + ext)
        if os.path.exists(p):
            return p
    return os.path.join(features_path, stem + ".h5")


class SlideDataset(Dataset):
    """wsi_dataset.py:21-83.  ``csv_path`` needs columns ``slide_id``, one 0/1 column per modality and (train) ``split``."""

    def __init__(self, dataset_name, csv_path, features_path, modalities, embedding_size=None, sample=-1, train=True):
        import pandas as pd
        self.dataset_name = dataset_name
        self.dataframe = pd.read_csv(csv_path)
        self.features_path = features_path
        self.modalities = modalities
        self.sample = sample
        self.train = train
        self.embedding_size = embedding_size

    def __len__(self):
        return len(self.dataframe)

    def sample_n(self, feats):
        """wsi_dataset.py:42-50: with replacement when the bag is shorter than ``sample``, a random subset otherwise."""
        if self.sample > -1:
            n = feats.shape[0]
            if n < self.sample:
                feats = feats[torch.randint(0, n, (self.sample,))]
            else:
                feats = feats[torch.randperm(n)[: self.sample]]
        return feats

    def case_files(self, index):
        """(slide_id, availability list, feature file per modality) of one case — also used by ResidentSlideStore."""
        row = self.dataframe.iloc[index]
        slide_id = row["slide_id"]
        if not self.train:
            return slide_id, [1], [_feature_file(self.features_path, f"{slide_id}")]
        labels = [int(row[m]) for m in self.modalities]
        tag = "" if row["split"] == "train" else f"_{row['split']}"
        files = [_feature_file(self.features_path, f"{slide_id}_{m}{tag}") for m in self.modalities]
        return slide_id, labels, files

    def __getitem__(self, index):
        slide_id, labels, files = self.case_files(index)
        if self.train:
            feats = [self.sample_n(load_features(f) if lab == 1 else torch.zeros([2, self.embedding_size]))
                     for lab, f in zip(labels, files)]
        else:
            feats = [load_features(files[0])]
        return {"feats": feats, "modality_labels": labels, "slide_id": slide_id}


def collate(batch):
    """wsi_dataset.py:85-99."""
    return {"feats": torch.stack([torch.stack(item["feats"]) for item in batch]),
            "modality_labels": torch.stack([torch.Tensor(item["modality_labels"]) for item in batch]),
            "slide_ids": [item["slide_id"] for item in batch]}


class SimpleDataset(Dataset):
    """wsi_dataset.py:102-115: every feature file of a directory, ``(features, slide_id)`` per item."""

    def __init__(self, features_path):
        self.features_path = features_path
        self.fnames = sorted(fn for fn in os.listdir(features_path) if fn.endswith((".h5", ".pt", ".npy")))

    def __len__(self):
        return len(self.fnames)

    def __getitem__(self, index):
        name = self.fnames[index]
        return load_features(os.path.join(self.features_path, name)), os.path.splitext(name)[0]


def simple_collate(batch):
    """wsi_dataset.py:122-125: (features, slide_id) items → ([bs, N, D] stacked features, list of ids)."""
    feats, ids = zip(*batch)
    return torch.stack(feats), list(ids)


# ----------------------------------------------------------------------------------------------------------------------
# HBM-resident store + on-device resampling
# ----------------------------------------------------------------------------------------------------------------------
class ResidentSlideStore:
    """All (case, modality) bags of a dataset packed into one device tensor ``[sum N, D]`` fp32.

    ``cases``: sequence of per-case lists, one entry per modality: a ``[N, D]`` tensor / array, or ``None`` when the case
    has no slide of that stain.  ``sample_batch`` returns what the reference's DataLoader + ``collate`` would deliver for the
    chosen cases — but built on the device by one kernel."""

    def __init__(self, cases: Sequence[Sequence[Optional[torch.Tensor]]], modalities: Sequence[str], device="cuda",
                 slide_ids: Optional[Sequence] = None):
        self.modalities = list(modalities)
        self.n_mod = len(self.modalities)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ResidentSlideStore keeps the features in GPU memory; there is no CPU fallback")
        self.slide_ids = list(slide_ids) if slide_ids is not None else list(range(len(cases)))
        lens, chunks, D = [], [], None
        for c, case in enumerate(cases):
            if len(case) != self.n_mod:
                raise ValueError(f"case {c} has {len(case)} entries for {self.n_mod} modalities")
            for bag in case:
                if bag is None:
                    lens.append(0)
                    continue
                bag = torch.as_tensor(bag, dtype=torch.float32)
                if bag.dim() != 2 or bag.shape[0] == 0:
                    raise ValueError(f"case {c}: a bag must be a non-empty [N, D] matrix (got {tuple(bag.shape)})")
                D = bag.shape[1] if D is None else D
                if bag.shape[1] != D:
                    raise ValueError(f"case {c}: feature width {bag.shape[1]} != {D}")
                lens.append(bag.shape[0])
                chunks.append(bag)
        if D is None:
            raise ValueError("the store needs at least one bag")
        self.D = D
        self.lens = torch.tensor(lens, dtype=torch.int32).view(len(cases), self.n_mod)           # 0 = stain missing
        offs = torch.zeros(len(lens) + 1, dtype=torch.int64)
        offs[1:] = torch.tensor(lens, dtype=torch.int64).cumsum(0)
        self.offsets = offs[:-1].view(len(cases), self.n_mod)
        total = int(offs[-1])
        self.features = torch.empty(total, D, dtype=torch.float32, device=self.device)
        o = 0
        for bag in chunks:                                   # one upload per bag, once, at construction
            self.features[o:o + bag.shape[0]].copy_(bag, non_blocking=True)
            o += bag.shape[0]
        self.modality_labels = (self.lens > 0).to(torch.float32)
        self._lens_dev = self.lens.to(self.device)
        self._offs_dev = self.offsets.to(self.device)

    @classmethod
    def from_dataset(cls, dataset: "SlideDataset", device="cuda", loader: Callable[[str], torch.Tensor] = load_features):
        """Read every feature file a training ``SlideDataset`` refers to, once."""
        cases, ids = [], []
        for i in range(len(dataset)):
            slide_id, labels, files = dataset.case_files(i)
            cases.append([loader(f) if lab == 1 else None for lab, f in zip(labels, files)])
            ids.append(slide_id)
        return cls(cases, dataset.modalities, device=device, slide_ids=ids)

    def __len__(self):
        return self.lens.shape[0]

    @property
    def nbytes(self) -> int:
        return self.features.numel() * 4

    def sample_batch(self, case_indices: Sequence[int], sample: int, seed: int, return_indices: bool = False) -> Dict:
        """``feats [bs, n_mod, sample, D]`` (device), ``modality_labels [bs, n_mod]`` (CPU, like the reference's collate),
        ``slide_ids``.  Sampling rule = wsi_dataset.py:42-50 (see csrc/sampler.cu); deterministic in (seed, case)."""
        idx = torch.as_tensor(case_indices, dtype=torch.int64)
        bs = idx.numel()
        idx_dev = idx.to(self.device, non_blocking=True)
        lens = self._lens_dev.index_select(0, idx_dev).reshape(-1).contiguous()
        offs = self._offs_dev.index_select(0, idx_dev).reshape(-1).contiguous()
        R = bs * self.n_mod
        out = torch.empty(bs, self.n_mod, sample, self.D, dtype=torch.float32, device=self.device)
        picked = torch.empty(R, sample, dtype=torch.int32, device=self.device) if return_indices else None
        # a bag's draw is keyed by (seed, position of the bag in this batch); the loader changes the seed every step
        with torch.cuda.device(self.device):
            call("mdl_sample_gather_f32", self.features, offs, lens, R, sample, self.D, int(seed) & 0xFFFFFFFFFFFFFFFF, out, picked,
                 stream_ptr(self.device))
        batch = {"feats": out, "modality_labels": self.modality_labels.index_select(0, idx),
                 "slide_ids": [self.slide_ids[i] for i in idx.tolist()]}
        if return_indices:
            batch["indices"] = picked.view(bs, self.n_mod, sample)
        return batch


class ResidentLoader:
    """Iterates a ``ResidentSlideStore`` like ``DataLoader(SlideDataset(..., sample=S), batch_size, shuffle, collate_fn=collate)``:
    a new case order every epoch (torch CPU generator), a new sample of every bag every step."""

    def __init__(self, store: ResidentSlideStore, batch_size: int, sample: int, shuffle: bool = True, drop_last: bool = False,
                 seed: int = 0):
        self.store, self.batch_size, self.sample = store, batch_size, sample
        self.shuffle, self.drop_last = shuffle, drop_last
        self.generator = torch.Generator().manual_seed(seed)
        self.seed = seed
        self.step = 0

    def __len__(self):
        n = len(self.store)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self) -> Iterator[Dict]:
        n = len(self.store)
        order = torch.randperm(n, generator=self.generator) if self.shuffle else torch.arange(n)
        for b in range(len(self)):
            cases = order[b * self.batch_size:(b + 1) * self.batch_size]
            self.step += 1
            yield self.store.sample_batch(cases, self.sample, seed=(self.seed << 32) ^ self.step)
