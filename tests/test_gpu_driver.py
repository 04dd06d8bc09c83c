"""End-to-end run of the generation driver on the GPU (random-init SD-1.5 weights, 2 DDIM steps): the files it writes follow
the reference's naming / index contract (txt2img_diffusers_stages_from_txt.py:262-263) and resume skips finished work."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_driver_writes_reference_index_layout(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from divergen_b200.generate import main
    argv = ["--from_file", os.path.join(HERE, "fixtures", "prompts"), "--outdir", str(tmp_path), "--n_samples", "3",
            "--max_batch_size", "2", "--random_init", "--num_inference_steps", "2", "--offset", "10", "--disable_overwrite"]
    assert main(argv) == 0
    got = sorted(os.listdir(tmp_path / "samples" / "sd"))
    want = sorted([f"1_{10 + j:07d}.latent.pt" for j in range(3)] +
                  [f"7_{10 + k * 3 + j:07d}.latent.pt" for k in range(2) for j in range(3)])
    assert got == want
    lat = torch.load(tmp_path / "samples" / "sd" / got[0])
    assert lat.shape == (4, 64, 64) and lat.dtype == torch.float16 and torch.isfinite(lat.float()).all()
    mtimes = {f: os.path.getmtime(tmp_path / "samples" / "sd" / f) for f in got}
    assert main(argv) == 0                                            # second run: everything exists -> skipped
    assert mtimes == {f: os.path.getmtime(tmp_path / "samples" / "sd" / f) for f in got}


def test_driver_writes_pngs_with_vae(tmp_path):
    """With a VAE attached the driver writes `<category_id>_<index:07d>.png`, 512x512 RGB (reference :262-267)."""
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from PIL import Image
    from divergen_b200.generate import main
    argv = ["--from_file", os.path.join(HERE, "fixtures", "prompts", "1.txt"), "--outdir", str(tmp_path), "--n_samples", "2",
            "--max_batch_size", "2", "--random_init", "--decode", "--num_inference_steps", "2", "--offset", "0"]
    assert main(argv) == 0
    got = sorted(os.listdir(tmp_path / "samples" / "sd"))
    assert got == ["1_0000000.png", "1_0000001.png"]
    im = Image.open(tmp_path / "samples" / "sd" / got[0])
    assert im.size == (512, 512) and im.mode == "RGB"


def test_driver_writes_category_layout_directly(tmp_path):
    """With --in_lvis_json_path the files land in `<outdir>/<stage>/<category_name>/` (what convert_dir_structure.py builds
    by copying), same file names; resume looks there too."""
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import json
    from divergen_b200.generate import main
    cats = tmp_path / "cats.json"
    cats.write_text(json.dumps([{"id": 1, "name": "aerosol_can"}, {"id": 7, "name": "alligator"}]))
    argv = ["--from_file", os.path.join(HERE, "fixtures", "prompts", "1.txt"), "--outdir", str(tmp_path / "out"), "--n_samples", "2",
            "--max_batch_size", "2", "--random_init", "--num_inference_steps", "2", "--offset", "0", "--disable_overwrite",
            "--in_lvis_json_path", str(cats)]
    assert main(argv) == 0
    d = tmp_path / "out" / "sd" / "aerosol_can"
    assert sorted(os.listdir(d)) == ["1_0000000.latent.pt", "1_0000001.latent.pt"]
    assert not (tmp_path / "out" / "samples").exists()
    m = os.path.getmtime(d / "1_0000000.latent.pt")
    assert main(argv) == 0 and os.path.getmtime(d / "1_0000000.latent.pt") == m
