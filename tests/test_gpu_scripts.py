"""the reference's own `train.py`, `train_gan.py` and `test.py`, unchanged, on top of the drop-in package.

The scripts and the caller modules they import are the reference's (from /root/reference where present, else the
byte-compiled copies under oracle/_ref); hydra / omegaconf / kornia / piqa are the minimal shims under shims/; the data
group the reference does not ship comes from conf/train/data/default.yaml (synthetic clips).  Two command-line overrides
are needed because of the environment, not because of the drop-in: `train.model.pretrained_flow=false` (the weight blob
is not in the reference repository) and `~train.scheduler.verbose` (torch 2.11's CosineAnnealingLR no longer accepts it)."""
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tools"))

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def project(tmp_path_factory):
    import run_reference_script as R
    try:
        R.reference_tree()
    except FileNotFoundError as e:
        pytest.skip(str(e))
    return tmp_path_factory.mktemp("vsrlab_project")


TRAIN_OVERRIDES = [
    "+experiment=basic", "train.model.pretrained_flow=false", "~train.scheduler.verbose",
    "train.max_epochs=1", "train.data.batch_size=4", "train.num_grad_acc=2", "train.data.num_workers=1",
    "train.data.datasets.train.length=12", "train.data.datasets.val.length=4", "train.data.datasets.train.seq=5",
    "train.data.datasets.train.lr_size=[32,32]", "train.model.cleaning_blocks=1", "train.model.res_blocks=1",
    "core.run_id=dropin",
]


def test_reference_train_py_runs_unchanged(project):
    import run_reference_script as R
    r = R.run("train", TRAIN_OVERRIDES, project, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-3000:]
    assert "Start Training" in r.stdout and "Starting Evaluation" in r.stdout and "Epoch 0 - Elapsed time" in r.stdout
    run_dir = project / "storage" / "video-super-resolution" / "dropin"
    assert (run_dir / "config.yaml").exists()
    ckpt = torch.load(run_dir / "checkpoint.tar", map_location="cpu")
    assert ckpt["epoch"] == 0 and "optimizer_state_dict" in ckpt and "scheduler_state_dict" in ckpt
    from vsrlab.vsr.models.RealBasicVSR.realbasicvsr import RealBasicVSR
    torch.manual_seed(0)
    fresh = RealBasicVSR(cleaning_blocks=1, mid_channels=64, upscale=4, res_blocks=1, pretrained_flow=False, train_flow=True)
    sd = ckpt["model_state_dict"]
    assert list(sd.keys()) == list(fresh.state_dict().keys())
    assert all(torch.isfinite(v).all() for v in sd.values() if v.is_floating_point())
    # the optimizer stepped: trained weights differ from any fresh initialisation's statistics only slightly, but the
    # Adam state holds one step for every trainable tensor
    # 12 clips in loader batches of 4 // 2 (core/utils.py:205), accumulated by 2: three optimizer steps; from the fourth
    # micro-batch on the model's forward / backward replay from CUDA graphs inside the reference's own loop (DDP, world 1)
    steps = {int(s["step"]) for s in ckpt["optimizer_state_dict"]["state"].values()}
    assert steps == {3}


def test_reference_train_gan_py_runs_unchanged(project):
    """`train_gan.py +experiment=basic_gan` with the drop-in generator AND discriminator.  Environment
    overrides: the experiment's `train.restore` names a checkpoint on the author's disk (null here), the VGG weights of
    PerceptualLoss need a download (`train.perceptual_loss=null` takes the script's own dummy_loss branch, train_gan.py:105)."""
    import run_reference_script as R
    ov = ["+experiment=basic_gan", "train.restore=null", "train.perceptual_loss=null", "train.model.pretrained_flow=false",
          "~train.scheduler.generator.verbose", "~train.scheduler.discriminator.verbose",
          "train.max_epochs=1", "train.data.batch_size=2", "train.num_grad_acc=2", "train.data.num_workers=1",
          "train.data.datasets.train.length=4", "train.data.datasets.val.length=2", "train.data.datasets.train.seq=5",
          "train.data.datasets.train.lr_size=[32,32]", "train.model.cleaning_blocks=1", "train.model.res_blocks=1",
          "core.run_id=dropin_gan"]
    r = R.run("train_gan", ov, project, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-3000:]
    assert "build discriminator" in r.stdout and "Start Training" in r.stdout and "Epoch 0 - Elapsed time" in r.stdout
    run_dir = project / "storage" / "video-super-resolution" / "dropin_gan"
    ckpt = torch.load(run_dir / "checkpoint.tar", map_location="cpu")
    assert all(torch.isfinite(v).all() for v in ckpt["model_state_dict"].values() if v.is_floating_point())
    # 4 clips in loader batches of batch_size // num_grad_acc = 1 (core/utils.py:205), accumulated by 2: two generator
    # steps, neither skipped by the GradScaler (finite gradients through the discriminator under fp16 autocast)
    assert {int(s["step"]) for s in ckpt["optimizer_state_dict"]["state"].values()} == {2}


def test_reference_test_py_runs_unchanged(project):
    """test.py's own directory protocol (test.py:94-141): lr_dir/fps=F_crf=C/{frames,video}/<name>, hr_dir/fps=F_crf=5/..."""
    import run_reference_script as R
    from torchvision.utils import save_image
    run_dir = project / "storage" / "video-super-resolution" / "dropin"
    if not (run_dir / "checkpoint.tar").exists():
        pytest.skip("needs the checkpoint written by the train.py test")
    # README.md:25-26 / test.py:81-82 expect a bare state_dict in <cfg_dir>/last.ckpt
    torch.save(torch.load(run_dir / "checkpoint.tar", map_location="cpu")["model_state_dict"], run_dir / "last.ckpt")
    lr_dir, hr_dir, out_dir = project / "lr", project / "hr", project / "out"
    g = torch.Generator().manual_seed(0)
    for fps in (6, 8, 10):
        hr = torch.rand(5, 3, 96, 128, generator=g)
        d = hr_dir / f"fps={fps}_crf=5"
        (d / "frames" / "clip0").mkdir(parents=True, exist_ok=True)
        (d / "video").mkdir(parents=True, exist_ok=True)
        (d / "video" / "clip0").write_bytes(b"x" * 1000)
        for i, f in enumerate(hr):
            save_image(f, str(d / "frames" / "clip0" / f"img{i:05d}.png"))
        lr = torch.nn.functional.interpolate(hr, size=(24, 32), mode="bilinear")
        for crf in (30, 32, 34):
            d = lr_dir / f"fps={fps}_crf={crf}"
            (d / "frames" / "clip0").mkdir(parents=True, exist_ok=True)
            (d / "video").mkdir(parents=True, exist_ok=True)
            (d / "video" / "clip0").write_bytes(b"x" * 100)
            for i, f in enumerate(lr):
                save_image(f, str(d / "frames" / "clip0" / f"img{i:05d}.png"))
    ov = ["+experiment=test", f"cfg_dir={run_dir}", f"lr_dir={lr_dir}", f"hr_dir={hr_dir}", f"out_dir={out_dir}", "window_size=3",
          "num_workers=2"]
    r = R.run("test", ov, project, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-3000:]
    res = out_dir / "dropin"
    csv = (res / "dropin.csv").read_text().splitlines()
    assert len(csv) == 1 + 9 and "PSNR" in csv[0] and "SSIM" in csv[0]
    pngs = sorted((res / "fps=6_crf=30" / "clip0").glob("*.png"))
    assert len(pngs) == 5                                   # 5 frames in windows of 3 + 2 (the ragged last window)
    from PIL import Image
    assert Image.open(pngs[0]).size == (128, 96)


def test_png_folder_pipeline_matches_the_reference_io_path(tmp_path):
    """vsrlab_b200.io.upscale_folder (uint8 in, uint8 out, overlapped copies, threaded PNG encode) writes byte-identical
    pixels to what the reference's test.py loop produces for the same frames: get_video -> fp32 windows -> model -> save_image
    (test.py:112-141; core/utils.py:282-288)."""
    import numpy as np
    from PIL import Image
    from torchvision.transforms.functional import to_tensor
    from torchvision.utils import save_image
    from vsrlab.vsr.models.RealBasicVSR.realbasicvsr import RealBasicVSR
    from vsrlab_b200 import functional as VF
    from vsrlab_b200.io import upscale_folder
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    model = RealBasicVSR(cleaning_blocks=1, mid_channels=64, upscale=4, res_blocks=1, pretrained_flow=False, train_flow=False).to(dev).eval()
    g = torch.Generator().manual_seed(5)
    frames = torch.rand(7, 3, 36, 52, generator=g)
    src = tmp_path / "lr"
    src.mkdir()
    for i, f in enumerate(frames):
        save_image(f, str(src / f"img{i:05d}.png"))
    # the reference's path, on the same decoded frames, in windows of 3 (3 + 3 + 1)
    lr = torch.stack([to_tensor(Image.open(p)) for p in sorted(src.glob("*.png"))]).unsqueeze(0)
    ref_dir = tmp_path / "ref"
    ref_dir.mkdir()
    with torch.no_grad(), VF.precision("bf16"):
        outs = [model(lr[:, i:i + 3].to(dev).contiguous())[0] for i in range(0, 7, 3)]
    sr = torch.cat(outs, dim=1)
    for i, f in enumerate(sr[0]):
        save_image(f, str(ref_dir / f"img{i:05d}.png"))
    res = upscale_folder(model, src, tmp_path / "out", window_size=3, workers=4, precision="bf16", keep_output=True)
    assert res["frames"] == 7 and res["windows"] == 3
    for i in range(7):
        a = np.asarray(Image.open(ref_dir / f"img{i:05d}.png"))
        b = np.asarray(Image.open(tmp_path / "out" / f"img{i:05d}.png"))
        assert a.shape == b.shape == (144, 208, 3) and np.array_equal(a, b), i
    assert np.array_equal(res["output"][0].permute(1, 2, 0).numpy(), np.asarray(Image.open(ref_dir / "img00000.png")))
