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
    """exp(entropy of the normalised singular values), rounded to two decimals — the model-selection metric
    (utils.py:180-201).

    A CUDA matrix stays on the device: its singular values are the square roots of the eigenvalues of the [D, D] Gram
    matrix, accumulated in fp64 by ``mdl_gram_f64`` (one kernel over the [n, D] embeddings) and diagonalised in fp64; a
    single scalar comes back.  fp64 keeps singular values down to 1e-8 of the largest, far below what moves the entropy.
    CPU input takes the reference's ``torch.svd`` route."""
    if embedding_matrix.is_cuda:
        from .._lib import call, stream_ptr
        e = embedding_matrix.detach().float().contiguous()
        n, d = e.shape
        gram = torch.empty(d, d, dtype=torch.float64, device=e.device)
        with torch.cuda.device(e.device):
            call("mdl_gram_f64", e, n, d, gram, stream_ptr(e.device))
        s = torch.linalg.eigvalsh(gram).clamp_min(0).sqrt().flip(0)[: min(n, d)].float()
    else:
        _, s, _ = torch.svd(embedding_matrix.float())
    p = s / torch.sum(s, dim=0) + eps
    p = p[: embedding_matrix.shape[1]]
    return round(float(torch.exp(-torch.sum(p * torch.log(p)))), 2)


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


def extract_slide_level_embeddings(args, val_dataloaders, ssl_model):
    """utils.py:68-90: run_inference on every downstream loader and pickle ``{"embeds", "slide_ids"}`` to
    ``<RESULS_SAVE_PATH>/<dataset>.pkl`` (the rank goes to wandb when ``args.log_ml``)."""
    from .file_utils import save_pkl
    for dataset_name, loader in val_dataloaders.items():
        print(f"\n* Extracting slide-level embeddings of {dataset_name}")
        results, rank = run_inference(ssl_model, loader, config=args)
        print(f"Rank for {dataset_name} = {rank}")
        print("\033[92mDone \033[0m")
        if getattr(args, "log_ml", False):
            import wandb
            wandb.run.summary[f"{dataset_name}_rank"] = rank
        save_pkl(os.path.join(args.RESULS_SAVE_PATH, f"{dataset_name}.pkl"), results)


def set_deterministic_mode(SEED, disable_cudnn=False):
    """utils.py:147-178: seed torch (CPU + every GPU), python and numpy.  The B200 kernels draw their dropout masks from
    torch's seed (ops.py: ``torch.initial_seed()``), so this makes a run repeatable end to end; cuDNN plays no part in the
    hot path, the flags are set only for whatever else the caller runs."""
    import random
    torch.manual_seed(SEED)
    random.seed(SEED)
    np.random.seed(SEED)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(SEED)
    if not disable_cudnn:
        torch.backends.cudnn.benchmark = False
        torch.backends.cudnn.deterministic = True
    else:
        torch.backends.cudnn.enabled = False


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
