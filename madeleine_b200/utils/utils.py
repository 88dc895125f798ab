"""Small host-side helpers the hot path's callers need (drop-ins for madeleine/utils/utils.py)."""
from collections import OrderedDict
import os

import numpy as np
import torch

DEVICE = torch.device("cuda" if torch.cuda.is_available() else "cpu")
HE_POSITION = 0


def set_model_precision(precision):
    """utils.py:124-140 equivalent: flag string → torch dtype used for torch.amp.autocast."""
    table = {"float64": torch.float64, "float32": torch.float32, "bfloat16": torch.bfloat16, "float16": torch.float16}
    if precision not in table:
        raise ValueError(f"unknown precision {precision}")
    return table[precision]


def smooth_rank_measure(embedding_matrix, eps=1e-7):
    """exp(entropy of the normalised singular values) — the model-selection metric (utils.py:180-201)."""
    _, s, _ = torch.svd(embedding_matrix.float())
    p = s / torch.sum(s, dim=0) + eps
    p = p[: min(embedding_matrix.shape)]
    return float(torch.exp(-torch.sum(p * torch.log(p))))


def run_inference(ssl_model, val_dataloader, config=None, torch_precision=None):
    """utils.py:27-66: encode_he per batch under autocast → {"embeds": [n,512] fp32, "slide_ids": [...]}, rank."""
    ssl_model.eval()
    if torch_precision is None:
        torch_precision = set_model_precision(config.precision)
    all_embeds, all_slide_ids = [], []
    with torch.no_grad():
        for feats, slide_ids in val_dataloader:
            with torch.amp.autocast(device_type="cuda", dtype=torch_precision,
                                    enabled=torch_precision in (torch.bfloat16, torch.float16)):
                wsi_embed = ssl_model.encode_he(feats, device=DEVICE)
            all_embeds.extend(wsi_embed.to(torch.float32).detach().cpu().numpy())
            all_slide_ids.append(slide_ids[0])
    all_embeds = np.array(all_embeds)
    rank = smooth_rank_measure(torch.Tensor(all_embeds))
    return {"embeds": all_embeds, "slide_ids": all_slide_ids}, rank


def load_checkpoint(args, ssl_model, path_to_checkpoint=None):
    """utils.py:92-122: strict load, retrying with the DataParallel ``module.`` prefix stripped."""
    path = path_to_checkpoint if path_to_checkpoint is not None else os.path.join(args.RESULS_SAVE_PATH, "model.pt")
    state_dict = torch.load(path, weights_only=False)
    try:
        ssl_model.load_state_dict(state_dict)
    except RuntimeError:
        ssl_model.load_state_dict(OrderedDict((k[7:], v) for k, v in state_dict.items()))
        print("Model loaded by removing module in state dict...")
    return ssl_model
