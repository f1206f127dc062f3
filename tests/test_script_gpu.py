"""End-to-end run of the launcher (script.py) on one GPU: config.yaml -> spawn -> DDP(ResNet) ->
training_loop (whole-step CUDA graph) -> checkpoints -> resume -> evaluation_loop."""
import os
import subprocess
import sys

import pytest
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

CONFIG = {
    "backend": "nccl", "world_size": 1, "master_addr": "127.0.0.1", "master_port": "12377",
    "dataset_cls_name": "SyntheticCIFAR10", "synthetic_train_size": 512, "synthetic_test_size": 128,
    # the shipped CIFAR augmentation spec: runs through the on-device input pipeline
    "data_aug_train": {"ToTensorTransform": {}, "StandardizeWhiteningTransform": {}, "FlipTransform": {"p": 0.5},
                       "PaddingTransform": {"pad_size": 4, "pad_type": "mirror"},
                       "RandomCropTransform": {"crop_size": 32}},
    "data_aug_test": {"ToTensorTransform": {}, "StandardizeWhiteningTransform": {}},
    "architecture_spec": "c3,16,3,1,1 n a r3 r3 r3 ap8,1,0 fc64,10", "preact": False, "use_proj": False,
    "dropout_prob": 0.0, "max_steps": 12, "batch_size": 64, "num_microbatches": 1, "cuda_graph": True,
    "optimizer_cls_name": "SGD",
    "optimizer_args": {"lr": 0.05, "momentum": 0.9, "dampening": 0.0, "nesterov": False, "weight_decay": 1e-4},
    "scheduler_cls_name": "MultiStepLR", "scheduler_step_unit": "epoch",
    "scheduler_args": {"milestones": [100], "gamma": 0.1},
    "checkpoint_strategy_cls_name": "FrequencyCheckpointStrategy",
    "checkpoint_strategy_args": {"unit": "batch", "frequency": 5},
}


def _run(args, cwd):
    env = dict(os.environ, PYTHONPATH=ROOT)
    return subprocess.run([sys.executable, os.path.join(ROOT, "script.py")] + args, cwd=cwd, env=env,
                          capture_output=True, text=True, timeout=600)


def test_script_train_resume_eval(tmp_path):
    run_dir = tmp_path / "models_dir" / "tiny"
    run_dir.mkdir(parents=True)
    (run_dir / "config.yaml").write_text(yaml.safe_dump(CONFIG, sort_keys=False))
    common = ["--models_dir", str(tmp_path / "models_dir"), "--run_name", "tiny", "--data_dir", str(tmp_path)]
    r = _run(["--mode", "train"] + common, tmp_path)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    losses = [float(l.split("loss:")[1]) for l in r.stdout.splitlines() if l.startswith("global step")]
    # (12 steps on randomly cropped / flipped batches: the per-step loss is noisy, training dynamics are covered by
    #  test_loss_curve_matches_bf16_oracle_200_steps; here: every step logged, finite, sane)
    assert len(losses) == 12 and all(l == l and l < 10.0 for l in losses) and min(losses) < 2.6
    ckpts = sorted(os.listdir(run_dir / "checkpoints"))
    assert "classifier_1.pth" in ckpts and "classifier_6.pth" in ckpts and "classifier_11.pth" in ckpts
    assert "optimizer_11.pth" in ckpts and "checkpoint_strategy_11.pth" in ckpts
    assert "standardizewhiteningtransform_1.pth" in ckpts   # fitted whitening, reference file name
    # resume: the newest aligned step is picked up and nothing is left to train
    r2 = _run(["--mode", "train"] + common, tmp_path)
    assert r2.returncode == 0, r2.stderr[-2000:]
    assert "Loaded classifier checkpoint" in r2.stdout
    r3 = _run(["--mode", "eval"] + common, tmp_path)
    assert r3.returncode == 0, r3.stderr[-2000:]
    assert "Test metrics:" in r3.stdout
