"""Drop-in for madeleine/utils/file_utils.py.

``save_pkl`` / ``load_pkl`` carry the extraction result ``{"embeds": [n, 512] fp32, "slide_ids": [...]}``
(bin/extract_slide_embeddings.py, utils.py:64); ``print_network`` leaves the ``model_config.txt`` summary that
``setup_model`` writes into the results directory (setup_components.py:189)."""
import pickle
from pathlib import Path


def save_pkl(filename, save_object):
    Path(filename).write_bytes(pickle.dumps(save_object, protocol=pickle.HIGHEST_PROTOCOL))


def load_pkl(filename):
    return pickle.loads(Path(filename).read_bytes())


def print_network(net, results_dir=None):
    """Parameter counts of ``net``; with ``results_dir`` also the module tree and the counts in ``model_config.txt``."""
    counts = [(p.numel(), p.requires_grad) for p in net.parameters()]
    total = sum(n for n, _ in counts)
    trainable = sum(n for n, req in counts if req)
    if results_dir is not None:
        summary = "\n".join([str(net), f"Total number of parameters: {total} ", f"Total number of trainable parameters: {trainable} ", ""])
        (Path(results_dir) / "model_config.txt").write_text(summary)
    return total, trainable
