"""Drop-in for madeleine/utils/file_utils.py: the pickle format the extraction script writes
(``{"embeds": [n, 512] fp32, "slide_ids": [...]}``, bin/extract_slide_embeddings.py + utils.py:64) and the
model summary file ``setup_model`` leaves in the results directory."""
import os
import pickle


def save_pkl(filename, save_object):
    with open(filename, "wb") as f:
        pickle.dump(save_object, f, protocol=pickle.HIGHEST_PROTOCOL)


def load_pkl(filename):
    with open(filename, "rb") as f:
        return pickle.load(f)


def print_network(net, results_dir=None):
    """Writes ``model_config.txt`` (module tree + parameter counts) into ``results_dir`` when given."""
    total = sum(p.numel() for p in net.parameters())
    trainable = sum(p.numel() for p in net.parameters() if p.requires_grad)
    if results_dir is not None:
        with open(os.path.join(results_dir, "model_config.txt"), "w") as f:
            f.write(f"{net}\n")
            f.write(f"Total number of parameters: {total} \n")
            f.write(f"Total number of trainable parameters: {trainable} \n")
    return total, trainable
