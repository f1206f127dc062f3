"""
Input transforms (reference: resnet/utils/transform_util.py:16-205, composed by
resnet/utils/data_util.py:48-113 from the `data_aug_train` / `data_aug_test` config dicts).

The reference applies its transforms per sample, in Python, on the host, inside a `num_workers=0` DataLoader,
and copies the batch to the GPU synchronously (data_util.py:218-227, training.py:94): at the ~24 k img/s of one
B200 that pipeline would starve the kernels more than tenfold. Here the SAME config dict is compiled into a
DeviceTransformPipeline: the dataset lives on the device as uint8 [M,H,W,C], the per-sample random draws are
made on the device, and ONE kernel (b200_augment_batch) builds the whole batch, directly in the bf16 NHWC
layout the stem convolution consumes. Given the same draws the fp32 result is bit-identical to the reference
(tests/test_transform_gpu.py).

Supported, in the order the reference's shipped configs use (any prefix / subset that keeps the order):
    ToTensorTransform, ZeroMeanWhiteningTransform | StandardizeWhiteningTransform, FlipTransform,
    PaddingTransform (zero | mirror), RandomCropTransform
ZCAWhiteningTransform, RandomScaleTransform and ColorTransform are not built (none of the shipped run configs
uses them): asking for one raises NotImplementedError instead of training on un-augmented data.
"""
from collections import OrderedDict
from typing import Dict, Optional

import torch as tc

from pytorch_ddp_resnet_b200 import ops

_ORDER = ["ToTensorTransform", "Whitening", "FlipTransform", "PaddingTransform", "RandomCropTransform"]
_WHITENING = ("ZeroMeanWhiteningTransform", "StandardizeWhiteningTransform")
_UNSUPPORTED = ("ZCAWhiteningTransform", "RandomScaleTransform", "ColorTransform")


class DeviceTransformPipeline(tc.nn.Module):
    """The transform chain of one data_aug spec. Fitted statistics are buffers named like the reference's
    transform attributes, so `fitted_state(name)` has the keys of its `{transform}_1.pth` checkpoints."""

    def __init__(self, data_shape, data_aug: Optional[Dict[str, Dict]]):
        super().__init__()
        H, W, C = (int(v) for v in data_shape)
        self.data_shape = (H, W, C)
        spec = OrderedDict(data_aug or {})
        stage = -1
        self.to_tensor = False
        self.whitening: Optional[str] = None
        self.flip_p: Optional[float] = None
        self.pad_size, self.pad_type = 0, "zero"
        self.crop_size: Optional[int] = None
        for name, kwargs in spec.items():
            kwargs = kwargs or {}
            if name in _UNSUPPORTED:
                raise NotImplementedError(
                    f"{name} is not built in pytorch_ddp_resnet_b200 (supported: ToTensor, ZeroMean / Standardize "
                    "whitening, Flip, Padding, RandomCrop); refusing to train on silently un-augmented data")
            key = "Whitening" if name in _WHITENING else name
            if key not in _ORDER:
                raise ValueError(f"unknown transform {name!r} in the data augmentation spec")
            if _ORDER.index(key) <= stage:
                raise NotImplementedError(
                    f"transform order {list(spec)} is not supported by the fused device pipeline "
                    f"(expected the order {_ORDER})")
            stage = _ORDER.index(key)
            if name == "ToTensorTransform":
                self.to_tensor = True
            elif name in _WHITENING:
                self.whitening = name
            elif name == "FlipTransform":
                self.flip_p = float(kwargs["p"])
            elif name == "PaddingTransform":
                self.pad_size, self.pad_type = int(kwargs["pad_size"]), str(kwargs["pad_type"])
                if self.pad_type not in ("zero", "mirror"):
                    raise ValueError("pad_type must be 'zero' or 'mirror'")
            elif name == "RandomCropTransform":
                self.crop_size = int(kwargs["crop_size"])
        if spec and not self.to_tensor:
            raise NotImplementedError("the device pipeline starts from uint8 images: ToTensorTransform must come first")
        self.register_buffer("_image_mean", tc.zeros(C, H, W))
        self.register_buffer("_image_stddev", tc.ones(C, H, W))
        self.register_buffer("_fitted", tc.tensor(self.whitening is None))

    # ---- shapes -------------------------------------------------------------------------------------
    @property
    def output_hw(self):
        H, W, _ = self.data_shape
        if self.crop_size is not None:
            return self.crop_size, self.crop_size
        return H + 2 * self.pad_size, W + 2 * self.pad_size

    # ---- fitting (reference: FittableTransform.fit, transform_util.py:57-62, 84-99) -------------------
    @tc.no_grad()
    def fit(self, data_u8: tc.Tensor, chunk: int = 8192) -> None:
        """Per-pixel-and-channel mean / standard deviation of the ToTensor'ed training set (population
        variance, like the reference's streaming formulas; accumulated in fp64 on the data's device)."""
        if self.whitening is None:
            return
        M, H, W, C = data_u8.shape
        s = tc.zeros(H, W, C, dtype=tc.float64, device=data_u8.device)
        ss = tc.zeros_like(s)
        for i in range(0, M, chunk):
            x = data_u8[i:i + chunk].to(tc.float64) / 255.0
            s += x.sum(0)
            ss += (x * x).sum(0)
        mean = s / M
        var = (ss / M - mean * mean).clamp_min(0.0)
        self._image_mean.copy_(mean.permute(2, 0, 1).to(tc.float32))
        if self.whitening == "StandardizeWhiteningTransform":
            self._image_stddev.copy_(var.sqrt().permute(2, 0, 1).to(tc.float32))
        self._fitted.fill_(True)

    def fitted_state(self) -> Dict[str, tc.Tensor]:
        """state_dict of the reference's fitted whitening transform (what it saves as `<name>_1.pth`)."""
        sd = {"_image_mean": self._image_mean.detach().cpu().clone(), "_fitted": tc.tensor(True)}
        if self.whitening == "StandardizeWhiteningTransform":
            sd["_image_stddev"] = self._image_stddev.detach().cpu().clone()
        return sd

    def load_fitted_state(self, sd: Dict[str, tc.Tensor]) -> None:
        self._image_mean.copy_(sd["_image_mean"].to(self._image_mean))
        if "_image_stddev" in sd:
            self._image_stddev.copy_(sd["_image_stddev"].to(self._image_stddev))
        self._fitted.fill_(bool(sd.get("_fitted", True)))

    # ---- batch construction ---------------------------------------------------------------------------
    def draw(self, batch: int, device, generator: Optional[tc.Generator] = None):
        """Per-sample random draws, made ON the device (no host round trip): (flip uint8 | None,
        top int32 | None, left int32 | None) with the reference's distributions (Bernoulli(p);
        uniform integers over the valid crop offsets, transform_util.py:161-163, 200-204)."""
        H, W, _ = self.data_shape
        flip = top = left = None
        if self.flip_p is not None:
            flip = (tc.rand(batch, device=device, generator=generator) < self.flip_p).to(tc.uint8)
        if self.crop_size is not None:
            ph, pw = H + 2 * self.pad_size, W + 2 * self.pad_size
            top = tc.randint(0, ph - self.crop_size + 1, (batch,), device=device, generator=generator,
                             dtype=tc.int32)
            left = tc.randint(0, pw - self.crop_size + 1, (batch,), device=device, generator=generator,
                              dtype=tc.int32)
        return flip, top, left

    def forward(self, data_u8: tc.Tensor, index: tc.Tensor, draws=None, generator=None, want_f32: bool = False):
        """Batch for the dataset rows `index`: bf16 logical-NCHW tensor in channels_last memory (what
        ResNet.forward takes without a copy); with want_f32 the reference's fp32 NCHW tensor instead."""
        if not bool(self._fitted):
            raise RuntimeError("whitening transform used before fit() / load_fitted_state()")
        flip, top, left = draws if draws is not None else self.draw(index.numel(), data_u8.device, generator)
        mean = self._image_mean if self.whitening is not None else None
        std = self._image_stddev if self.whitening == "StandardizeWhiteningTransform" else None
        of, ob = ops.augment_batch(data_u8, index, flip=flip, top=top, left=left, mean=mean, stddev=std,
                                   pad=self.pad_size, pad_mirror=self.pad_type == "mirror", out_hw=self.output_hw,
                                   to_tensor=self.to_tensor, want_f32=want_f32, want_bf16=not want_f32)
        return of if want_f32 else ob.permute(0, 3, 1, 2)
