"""
Generates tests/golden/transforms_tiny.npz by running the UNMODIFIED reference transform classes
(/root/reference/resnet/utils/transform_util.py) on a small random uint8 dataset, with their per-sample random
draws injected (FlipTransform draws from tc.distributions.Categorical, RandomCropTransform from tc.randint).
Run it in the build container (the reference is not available on the GPU box):

    python tests/golden/make_transform_golden.py
"""
import os
import sys

import numpy as np
import PIL.Image
import torch

sys.path.insert(0, "/root/reference")
from resnet.utils import transform_util as T  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
H, W, C, M = 10, 12, 3, 12


class _Draws:
    """Replaces the two random sources of transform_util with a scripted sequence."""

    def __init__(self):
        self.flip, self.ints = [], []

    def install(self):
        draws = self

        class _Cat:
            def __init__(self, probs):
                pass

            def sample(self):
                return torch.tensor(int(draws.flip.pop(0)))

        self._cat, self._randint = T.tc.distributions.Categorical, T.tc.randint
        T.tc.distributions.Categorical = _Cat
        T.tc.randint = lambda low, high, size: torch.tensor([int(draws.ints.pop(0))])

    def remove(self):
        T.tc.distributions.Categorical, T.tc.randint = self._cat, self._randint


def run_case(data, index, whitening, flips, pad_size, pad_type, tops, lefts, crop_size, fitted):
    """Builds the reference transform chain as data_util._get_transforms does and applies it per sample."""
    shape = data[0].shape                       # (H, W, C) like torchvision's dataset.data[0].shape
    chain = [T.ToTensorTransform(shape)]
    shape = chain[-1].output_shape
    if whitening:
        cls = T.StandardizeWhiteningTransform if whitening == "standardize" else T.ZeroMeanWhiteningTransform
        t = cls(shape)
        with torch.no_grad():
            t._image_mean.copy_(fitted["mean"])
            if whitening == "standardize":
                t._image_stddev.copy_(fitted["stddev"])
        t.register_buffer("_fitted", torch.tensor(True))
        chain.append(t)
    if flips is not None:
        chain.append(T.FlipTransform(shape, p=0.5))
    if pad_size:
        chain.append(T.PaddingTransform(shape, pad_size=pad_size, pad_type=pad_type))
        shape = chain[-1].output_shape
    if crop_size:
        chain.append(T.RandomCropTransform(shape, crop_size=crop_size))
    d = _Draws()
    d.install()
    try:
        out = []
        for b, i in enumerate(index):
            if flips is not None:
                d.flip = [flips[b]]
            if crop_size:
                d.ints = [tops[b], lefts[b]]
            x = PIL.Image.fromarray(data[i])
            with torch.no_grad():
                for t in chain:
                    x = t(x)
            out.append(x.numpy().copy())
    finally:
        d.remove()
    return np.stack(out).astype(np.float32)


def main():
    rng = np.random.default_rng(0)
    data = rng.integers(0, 256, size=(M, H, W, C), dtype=np.uint8)
    # the reference's fit (streaming mean / variance over the ToTensor'ed training set)
    to_t = T.ToTensorTransform((H, W, C))
    ds = [(to_t(PIL.Image.fromarray(im)), 0) for im in data]
    st = T.StandardizeWhiteningTransform(to_t.output_shape)
    with torch.no_grad():
        st.fit(ds)
    zm = T.ZeroMeanWhiteningTransform(to_t.output_shape)
    with torch.no_grad():
        zm.fit(ds)
    fitted = {"mean": st._image_mean.detach().clone(), "stddev": st._image_stddev.detach().clone()}
    assert torch.equal(zm._image_mean, st._image_mean)
    out = {"data": data, "fit_mean": fitted["mean"].numpy(), "fit_stddev": fitted["stddev"].numpy()}
    B = 9
    index = rng.integers(0, M, size=B)
    flips = rng.integers(0, 2, size=B)
    cases = {
        "a": dict(whitening="standardize", flips=flips, pad_size=2, pad_type="mirror", crop_size=9),
        "b": dict(whitening="zeromean", flips=flips[::-1].copy(), pad_size=3, pad_type="zero", crop_size=8),
        "c": dict(whitening=None, flips=None, pad_size=0, pad_type="zero", crop_size=None),
        "d": dict(whitening="standardize", flips=None, pad_size=0, pad_type="zero", crop_size=None),
        "e": dict(whitening=None, flips=flips, pad_size=1, pad_type="zero", crop_size=None),
    }
    for name, c in cases.items():
        tops = lefts = None
        if c["crop_size"]:
            tops = rng.integers(0, H + 2 * c["pad_size"] - c["crop_size"] + 1, size=B)
            lefts = rng.integers(0, W + 2 * c["pad_size"] - c["crop_size"] + 1, size=B)
        y = run_case(data, index, c["whitening"], c["flips"], c["pad_size"], c["pad_type"], tops, lefts,
                     c["crop_size"], fitted)
        out[f"{name}/index"] = index.astype(np.int64)
        out[f"{name}/out"] = y
        out[f"{name}/whitening"] = np.array(c["whitening"] or "none")
        out[f"{name}/pad_size"] = np.array(c["pad_size"])
        out[f"{name}/pad_type"] = np.array(c["pad_type"])
        out[f"{name}/crop_size"] = np.array(c["crop_size"] or 0)
        if c["flips"] is not None:
            out[f"{name}/flips"] = c["flips"].astype(np.uint8)
        if tops is not None:
            out[f"{name}/tops"] = tops.astype(np.int32)
            out[f"{name}/lefts"] = lefts.astype(np.int32)
    path = os.path.join(HERE, "transforms_tiny.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if k.endswith("/out")})


if __name__ == "__main__":
    main()
