"""Drop-in for madeleine/utils/loss.py: InfoNCE (global loss) and GOT (local Graph-OT loss) on fused sm_100a kernels."""
import torch
from torch import nn

from .. import ops

__all__ = ["InfoNCE", "info_nce", "GOT", "init_intra_wsi_loss_function"]


class InfoNCE(nn.Module):
    """loss.py:10-133. In-batch-negative InfoNCE (optionally symmetric) computed by mdl_infonce_fwd/bwd."""

    def __init__(self, temperature=0.1, reduction="mean", negative_mode="unpaired"):
        super().__init__()
        self.temperature = temperature
        self.reduction = reduction
        self.negative_mode = negative_mode

    def forward(self, query, positive_key, negative_keys=None, symmetric=False):
        return self.info_nce(query, positive_key, negative_keys, temperature=self.temperature, reduction=self.reduction,
                             negative_mode=self.negative_mode, symmetric=symmetric)

    def info_nce(self, query, positive_key, negative_keys=None, temperature=0.1, reduction="mean", negative_mode="unpaired",
                 symmetric=False):
        if query.dim() != 2:
            raise ValueError("<query> must have 2 dimensions.")
        if positive_key.dim() != 2:
            raise ValueError("<positive_key> must have 2 dimensions.")
        if negative_keys is not None:
            if negative_mode == "unpaired" and negative_keys.dim() != 2:
                raise ValueError("<negative_keys> must have 2 dimensions if <negative_mode> == 'unpaired'.")
            if negative_mode == "paired" and negative_keys.dim() != 3:
                raise ValueError("<negative_keys> must have 3 dimensions if <negative_mode> == 'paired'.")
        if len(query) != len(positive_key):
            raise ValueError("<query> and <positive_key> must must have the same number of samples.")
        if negative_keys is not None:
            if negative_mode == "paired" and len(query) != len(negative_keys):
                raise ValueError("If negative_mode == 'paired', then <negative_keys> must have the same number of samples as <query>.")
        if query.shape[-1] != positive_key.shape[-1]:
            raise ValueError("Vectors of <query> and <positive_key> should have the same number of components.")
        if negative_keys is not None:
            if query.shape[-1] != negative_keys.shape[-1]:
                raise ValueError("Vectors of <query> and <negative_keys> should have the same number of components.")
            # loss.py:93-110 builds logits/labels for explicit negatives but never computes a loss: the reference
            # returns None on this branch; no caller uses it.
            return None
        if reduction not in ("none", "mean", "sum"):
            raise ValueError(f"{reduction} is not a valid value for reduction")
        return ops.info_nce(query, positive_key, temperature=temperature, reduction=reduction, symmetric=symmetric)


def info_nce(query, positive_key, negative_keys=None, temperature=0.1, reduction="mean", negative_mode="unpaired", symmetric=False):
    return InfoNCE(temperature, reduction, negative_mode)(query, positive_key, negative_keys, symmetric=symmetric)


def init_intra_wsi_loss_function(config):
    """loss.py:138-156."""
    if config["intra_modality_mode_wsi"] in ("reconstruct_avg_emb", "reconstruct_masked_emb"):
        return nn.MSELoss()
    return InfoNCE(temperature=config["temperature"])


def GOT(v_, q_, subsample=None, _slot=-1):
    """loss.py:278-301. v_, q_ [m, N, 128] token embeddings of the m cases that have this stain → scalar wd + gwd.

    Quirk Q3 is reproduced: the permutation is drawn over the *batch* size with torch's global CPU generator and
    indexes the token axis, so n = min(m, subsample) of the first m tokens are used."""
    if subsample is not None:
        patch_indices = torch.randperm(v_.shape[0])[:subsample]
        if patch_indices.numel() and int(patch_indices.max()) >= v_.shape[1]:
            # same failure as the reference (index out of range), but with the reason: e.g. a token window smaller than the batch
            raise IndexError(f"GOT indexes the token axis with a permutation of the {v_.shape[0]} cases (loss.py:281-284) but only "
                             f"{v_.shape[1]} tokens per bag are available (b200_token_window must be >= the batch size)")
        patch_indices = patch_indices.to(v_.device)
        v_ = v_[:, patch_indices, :]
        q_ = q_[:, patch_indices, :]
    # _slot >= 0 (used by calculate_losses): run on a side stream; the caller joins with ops.got_join before using the value
    return ops.got_loss(v_, q_, _slot)
