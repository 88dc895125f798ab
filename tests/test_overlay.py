"""The package overlays the hot path on a checkout of the reference: `core.*` / `madeleine.*` names that are not rebuilt here
(setup_components, process_args, datasets.modalities, preprocessing) resolve to the reference's own files, and those files'
imports of the hot path resolve back to the B200 modules — which is what lets bin/pretrain.py run unchanged."""
import os
import subprocess
import sys
import textwrap

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(code, ref_root, cwd=None):
    env = dict(os.environ, PYTHONPATH=REPO, MADELEINE_REFERENCE_ROOT=str(ref_root))
    return subprocess.run([sys.executable, "-c", textwrap.dedent(code)], env=env, cwd=cwd, capture_output=True, text=True, timeout=300)


def test_fall_through_to_a_reference_checkout(tmp_path):
    ref = tmp_path / "ref"
    (ref / "madeleine" / "utils").mkdir(parents=True)
    (ref / "madeleine" / "datasets").mkdir()
    (ref / "madeleine" / "preprocessing").mkdir()
    (ref / "madeleine" / "__init__.py").write_text("")
    (ref / "madeleine" / "utils" / "setup_components.py").write_text(
        "from madeleine.models.Model import MADELEINE\nfrom madeleine.utils.loss import InfoNCE, GOT\n"
        "from madeleine.utils.file_utils import print_network\nfrom madeleine.datasets.wsi_dataset import SlideDataset, collate\n"
        "from madeleine.datasets.modalities import modality_dicts\nMARK = 42\n")
    (ref / "madeleine" / "utils" / "loss.py").write_text("raise RuntimeError('the reference loss must be shadowed')\n")
    (ref / "madeleine" / "datasets" / "modalities.py").write_text("modality_dicts = {'TOY': ['HE', 'ER']}\n")
    (ref / "madeleine" / "preprocessing" / "__init__.py").write_text("WHO = 'reference'\n")
    out = _run("""
        from core.utils.setup_components import MARK, MADELEINE, InfoNCE, GOT, SlideDataset, modality_dicts
        from core.utils.trainer import train_loop
        from core.utils.utils import extract_slide_level_embeddings, load_checkpoint, set_deterministic_mode, run_inference
        from core.models.factory import create_model_from_pretrained
        from core.utils.file_utils import save_pkl
        from core.datasets.wsi_dataset import SimpleDataset, simple_collate
        import madeleine.preprocessing, madeleine_b200
        import madeleine_b200.models.Model as M, madeleine_b200.utils.loss as L, madeleine_b200.datasets.wsi_dataset as D
        assert MARK == 42 and modality_dicts == {'TOY': ['HE', 'ER']} and madeleine.preprocessing.WHO == 'reference'
        assert MADELEINE is M.MADELEINE and InfoNCE is L.InfoNCE and GOT is L.GOT and SlideDataset is D.SlideDataset
        print('OK')
    """, ref)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr


def test_scripts_style_sys_path_discovery(tmp_path):
    """bin/*.py do `sys.path.append('../')` from the checkout's bin/ directory and then import core.*: no env var needed."""
    ref = tmp_path / "ref"
    (ref / "madeleine" / "utils").mkdir(parents=True)
    (ref / "bin").mkdir()
    (ref / "madeleine" / "__init__.py").write_text("")
    (ref / "madeleine" / "utils" / "setup_components.py").write_text("from madeleine.models.Model import MADELEINE\nMARK = 7\n")
    env = dict(os.environ, PYTHONPATH=REPO)
    env.pop("MADELEINE_REFERENCE_ROOT", None)
    code = "import sys; sys.path.append('../')\nfrom core.utils.setup_components import MARK, MADELEINE\nprint(MARK, MADELEINE.__module__)"
    out = subprocess.run([sys.executable, "-c", code], env=env, cwd=str(ref / "bin"), capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.split() == ["7", "madeleine_b200.models.Model"], out.stdout + out.stderr


@pytest.mark.skipif(not os.path.isdir("/root/reference/madeleine"), reason="the reference checkout exists only in the build container")
def test_real_reference_setup_components_binds_the_b200_modules():
    out = _run("""
        import core.utils.setup_components as sc
        from core.utils.process_args import get_args
        import madeleine_b200.models.Model as M, madeleine_b200.utils.loss as L
        assert sc.__file__.startswith('/root/reference/'), sc.__file__
        assert sc.MADELEINE is M.MADELEINE and sc.InfoNCE is L.InfoNCE and sc.GOT is L.GOT
        assert callable(sc.setup_model) and callable(sc.setup_optim) and callable(sc.setup_losses) and callable(get_args)
        print('OK')
    """, "/root/reference")
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr
