"""CPU checks of the input pipeline: the numpy oracle against golden vectors made from the UNMODIFIED
reference transforms (tests/golden/make_transform_golden.py), and the host-side logic of
utils/transform_util.py / utils/data_util.py (spec parsing, fitting, checkpoint format, loaders)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import transform_oracle as TO  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "transforms_tiny.npz")
CASES = "abcde"


def case_kwargs(z, c):
    wh = str(z[f"{c}/whitening"])
    has = lambda k: f"{c}/{k}" in z.files  # noqa: E731
    return dict(mean=z["fit_mean"] if wh != "none" else None,
                stddev=z["fit_stddev"] if wh == "standardize" else None,
                flips=z[f"{c}/flips"] if has("flips") else None,
                pad_size=int(z[f"{c}/pad_size"]), pad_type=str(z[f"{c}/pad_type"]),
                tops=z[f"{c}/tops"] if has("tops") else None, lefts=z[f"{c}/lefts"] if has("lefts") else None,
                crop_size=int(z[f"{c}/crop_size"]) or None)


@pytest.mark.parametrize("c", CASES)
def test_oracle_pipeline_bit_exact_vs_reference(c):
    z = np.load(GOLDEN)
    y = TO.pipeline(z["data"], z[f"{c}/index"], **case_kwargs(z, c))
    assert y.dtype == np.float32 and np.array_equal(y, z[f"{c}/out"])


def test_oracle_fit_bit_exact_vs_reference():
    z = np.load(GOLDEN)
    xs = np.stack([TO.to_tensor(im) for im in z["data"]])
    mean, std = TO.fit_whitening(xs)
    assert np.array_equal(mean, z["fit_mean"]) and np.array_equal(std, z["fit_stddev"])


def test_pipeline_spec_parsing_and_fit():
    from pytorch_ddp_resnet_b200.utils.transform_util import DeviceTransformPipeline as P
    z = np.load(GOLDEN)
    spec = {"ToTensorTransform": {}, "StandardizeWhiteningTransform": {}, "FlipTransform": {"p": 0.5},
            "PaddingTransform": {"pad_size": 4, "pad_type": "mirror"}, "RandomCropTransform": {"crop_size": 9}}
    p = P(z["data"].shape[1:], spec)
    assert (p.whitening, p.flip_p, p.pad_size, p.pad_type, p.crop_size) == \
        ("StandardizeWhiteningTransform", 0.5, 4, "mirror", 9)
    assert p.output_hw == (9, 9) and not bool(p._fitted)
    with pytest.raises(RuntimeError):
        p(torch.zeros(1, 10, 12, 3, dtype=torch.uint8), torch.zeros(1, dtype=torch.int64))
    p.fit(torch.from_numpy(z["data"]))
    # fp64 batch statistics vs the reference's fp32 streaming formulas: same numbers to fp32 round-off
    assert np.allclose(p._image_mean.numpy(), z["fit_mean"], rtol=1e-5, atol=1e-6)
    assert np.allclose(p._image_stddev.numpy(), z["fit_stddev"], rtol=1e-4, atol=1e-6)
    sd = p.fitted_state()   # keys of the reference's fitted-transform checkpoint
    assert set(sd) == {"_image_mean", "_image_stddev", "_fitted"} and sd["_image_mean"].shape == (3, 10, 12)
    q = P(z["data"].shape[1:], {"ToTensorTransform": {}, "StandardizeWhiteningTransform": {}})
    q.load_fitted_state(sd)
    assert torch.equal(q._image_stddev, p._image_stddev) and q.output_hw == (10, 12)
    assert P((32, 32, 3), {}).output_hw == (32, 32)
    for bad in ({"ToTensorTransform": {}, "ZCAWhiteningTransform": {}},
                {"ToTensorTransform": {}, "RandomCropTransform": {"crop_size": 8}, "FlipTransform": {"p": 0.5}},
                {"FlipTransform": {"p": 0.5}}):
        with pytest.raises(NotImplementedError):
            P((32, 32, 3), bad)
    with pytest.raises(ValueError):
        P((32, 32, 3), {"ToTensorTransform": {}, "NoSuchTransform": {}})


def test_get_datasets_with_augmentation_spec(tmp_path):
    """Synthetic data + the shipped CIFAR augmentation spec: device datasets, fitted statistics saved in the
    reference's `<transform>_1.pth` format and re-used on the next start and by the test split."""
    import yaml
    from pytorch_ddp_resnet_b200.utils import data_util as D
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = yaml.safe_load(open(os.path.join(root, "models_dir", "wrn-28-10-dropout_cifar10", "config.yaml")))
    kw = dict(dataset_cls_name="SyntheticCIFAR10", data_dir=str(tmp_path), data_aug_train=cfg["data_aug_train"],
              data_aug_test=cfg["data_aug_test"], checkpoint_dir=str(tmp_path / "ck"), synthetic_train_size=64,
              synthetic_test_size=32)
    ds = D.get_datasets(**kw)
    tr, te = ds["dataset_train"], ds["dataset_test"]
    assert isinstance(tr, D.DeviceDataset) and tr.data.dtype == torch.uint8 and tuple(tr.data.shape) == (64, 32, 32, 3)
    assert len(te) == 32 and bool(te.pipeline._fitted) and te.pipeline.flip_p is None
    assert torch.equal(te.pipeline._image_mean, tr.pipeline._image_mean)
    path = tmp_path / "ck" / "standardizewhiteningtransform_1.pth"
    assert path.exists() and set(torch.load(path)) == {"_image_mean", "_image_stddev", "_fitted"}
    ds2 = D.get_datasets(**kw)   # second start: loaded, not refitted
    assert torch.equal(ds2["dataset_train"].pipeline._image_stddev, tr.pipeline._image_stddev)
    sm = D.get_samplers(0, 2, tr, te)
    dl = D.get_dataloaders(32, 1, 2, tr, te, **sm)
    assert isinstance(dl["dl_train"], D.DeviceDataLoader) and len(dl["dl_train"]) == 2 and len(dl["dl_test"]) == 1
    # empty specs keep the plain float datasets
    plain = D.get_datasets("SyntheticCIFAR10", str(tmp_path), {}, {}, synthetic_train_size=8, synthetic_test_size=8)
    assert isinstance(plain["dataset_train"], torch.utils.data.TensorDataset)
