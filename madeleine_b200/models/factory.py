"""Drop-in for madeleine/models/factory.py:16-39 (checkpoint bundle = model.pt + model_config.json)."""
import json
import os
from argparse import Namespace

from .Model import create_model
from ..utils.utils import set_model_precision


def create_model_from_pretrained(local_dir: str, download: bool = True):
    """Load a MADELEINE bundle from ``local_dir``.  The reference always calls huggingface_hub.snapshot_download
    (factory.py:23); that needs network access, so it is attempted only when the bundle is not already on disk."""
    os.makedirs(local_dir, exist_ok=True)
    cfg_path = os.path.join(local_dir, "model_config.json")
    ckpt_path = os.path.join(local_dir, "model.pt")
    if download and not (os.path.exists(cfg_path) and os.path.exists(ckpt_path)):
        from huggingface_hub import snapshot_download
        print(f"* Downloading model at {local_dir}")
        snapshot_download(repo_id="MahmoodLab/madeleine", local_dir=local_dir)
    model_cfg = Namespace(**json.load(open(cfg_path)))
    model = create_model(model_cfg, device="cuda", checkpoint_path=ckpt_path)
    precision = set_model_precision(model_cfg.precision)
    return model, precision
