"""
Datasets / samplers / loaders (reference: resnet/utils/data_util.py:21-232), same call surface for script.py.

The reference keeps PIL images on the host, transforms them one at a time in Python inside a `num_workers=0`
DataLoader and copies every batch to the device synchronously (data_util.py:218-227, training.py:94). Here a
dataset is a device-resident uint8 tensor [M,H,W,C] + labels (CIFAR-10: 150 MB of 180 GB), the `data_aug_*`
config dicts are compiled into a DeviceTransformPipeline (utils/transform_util.py) and a DeviceDataLoader
builds each batch with ONE kernel launch from the sampler's indices, already in the layout the stem
convolution consumes. Fitted whitening statistics are stored in the reference's checkpoint format
(`standardizewhiteningtransform_1.pth` ...) so that either implementation can reuse the other's.

  * dataset_cls_name 'SyntheticCIFAR10' / 'SyntheticCIFAR100' / 'SyntheticImageNet': seeded in-memory data (a
    learnable task: class pattern + noise), no files, no network. With empty data_aug specs they are float
    tensors served by a plain DataLoader (as in round 1); with a data_aug spec they are uint8 images that go
    through the device pipeline like a real dataset;
  * any torchvision dataset class exposing `.data` (uint8 [M,H,W,C]) and `.targets`, read from `data_dir`
    WITHOUT downloading (there is no network), through the device pipeline.
The batch-size rule is the reference's: config batch_size is GLOBAL; a loader yields
batch_size // (num_microbatches * world_size) samples (data_util.py:216).
"""
import os
from typing import Any, Dict, Optional

import torch as tc

from pytorch_ddp_resnet_b200.utils.transform_util import DeviceTransformPipeline

_SYNTHETIC = {"SyntheticCIFAR10": (10, 32), "SyntheticCIFAR100": (100, 32), "SyntheticImageNet": (1000, 224)}


def _synthetic(name: str, train: bool, size: int):
    classes, hw = _SYNTHETIC[name]
    g = tc.Generator().manual_seed(0)
    pattern = tc.randn(classes, 3, hw, hw, generator=g)          # shared by train and test
    g = tc.Generator().manual_seed(1 if train else 2)
    y = tc.randint(0, classes, (size,), generator=g)
    x = 0.3 * pattern[y] + tc.randn(size, 3, hw, hw, generator=g)
    return x, y


class DeviceDataset:
    """uint8 images [M,H,W,C] and int64 labels [M], resident on `device`, with the transform pipeline of
    their split. Indexable like a map-style dataset (samplers only need len())."""

    def __init__(self, data_u8: tc.Tensor, labels: tc.Tensor, pipeline: DeviceTransformPipeline, device):
        assert data_u8.dtype == tc.uint8 and data_u8.dim() == 4 and labels.shape[0] == data_u8.shape[0]
        self.data = data_u8.contiguous().to(device)
        self.labels = labels.to(tc.int64).to(device)
        self.pipeline = pipeline.to(device)
        self.device = tc.device(device)

    def __len__(self) -> int:
        return self.data.shape[0]

    def batch(self, index: tc.Tensor, generator=None):
        return self.pipeline(self.data, index, generator=generator), self.labels[index]


class DeviceDataLoader:
    """Iterates the sampler's indices in batches; every batch is one b200_augment_batch launch + one label
    gather on the device. The per-sample random draws come from a device generator seeded per loader."""

    def __init__(self, dataset: DeviceDataset, batch_size: int, sampler, seed: int = 0):
        self.dataset, self.batch_size, self.sampler = dataset, int(batch_size), sampler
        self.generator = None
        if dataset.device.type == "cuda":
            self.generator = tc.Generator(device=dataset.device)
            self.generator.manual_seed(seed)

    def __len__(self) -> int:
        return (len(self.sampler) + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        order = tc.tensor(list(iter(self.sampler)), dtype=tc.int64).to(self.dataset.device, non_blocking=True)
        for i in range(0, order.numel(), self.batch_size):
            yield self.dataset.batch(order[i:i + self.batch_size], self.generator)


def _raw_split(dataset_cls_name: str, data_dir: str, train: bool, kwargs):
    """(uint8 [M,H,W,C], int64 [M]) of one split."""
    if dataset_cls_name in _SYNTHETIC:
        n = int(kwargs.get("synthetic_train_size" if train else "synthetic_test_size", 2048 if train else 512))
        x, y = _synthetic(dataset_cls_name, train, n)
        u8 = (x * 40.0 + 128.0).round().clamp(0, 255).to(tc.uint8).permute(0, 2, 3, 1).contiguous()
        return u8, y
    import torchvision
    cls = getattr(torchvision.datasets, dataset_cls_name)
    ds = cls(root=data_dir, train=train, download=False, transform=None)
    data = tc.as_tensor(ds.data)
    if data.dtype != tc.uint8 or data.dim() != 4:
        raise NotImplementedError(f"{dataset_cls_name}: the device pipeline needs uint8 [M,H,W,C] images in .data")
    return data, tc.as_tensor(ds.targets)


def get_datasets(dataset_cls_name: str, data_dir: str, data_aug_train: Optional[Dict] = None,
                 data_aug_test: Optional[Dict] = None, checkpoint_dir: Optional[str] = None,
                 **kwargs: Dict[str, Any]):
    data_aug_train, data_aug_test = data_aug_train or {}, data_aug_test or {}
    if dataset_cls_name in _SYNTHETIC and not data_aug_train and not data_aug_test:
        n_train = int(kwargs.get("synthetic_train_size", 2048))
        n_test = int(kwargs.get("synthetic_test_size", 512))
        return {"dataset_train": tc.utils.data.TensorDataset(*_synthetic(dataset_cls_name, True, n_train)),
                "dataset_test": tc.utils.data.TensorDataset(*_synthetic(dataset_cls_name, False, n_test))}
    device = tc.device("cuda", tc.cuda.current_device()) if tc.cuda.is_available() else tc.device("cpu")
    x_train, y_train = _raw_split(dataset_cls_name, data_dir, True, kwargs)
    x_test, y_test = _raw_split(dataset_cls_name, data_dir, False, kwargs)
    shape = tuple(x_train.shape[1:])
    pipe_train = DeviceTransformPipeline(shape, data_aug_train)
    pipe_test = DeviceTransformPipeline(shape, data_aug_test)
    ds_train = DeviceDataset(x_train, y_train, pipe_train, device)
    if pipe_train.whitening is not None:
        # fitted once per run directory, stored like the reference's fitted transforms (data_util.py:78-91)
        kind = pipe_train.whitening.lower()
        path = os.path.join(checkpoint_dir, f"{kind}_1.pth") if checkpoint_dir else None
        if path and os.path.exists(path):
            pipe_train.load_fitted_state(tc.load(path, map_location="cpu"))
        else:
            pipe_train.fit(ds_train.data)
            if path:
                os.makedirs(checkpoint_dir, exist_ok=True)
                tc.save(pipe_train.fitted_state(), path)
    if pipe_test.whitening is not None:
        # a fittable test transform must reuse the training one (data_util.py:93-101)
        if pipe_test.whitening != pipe_train.whitening:
            raise ValueError("Fittable test transform not in reusable_transforms.")
        pipe_test.load_fitted_state(pipe_train.fitted_state())
    return {"dataset_train": ds_train, "dataset_test": DeviceDataset(x_test, y_test, pipe_test, device)}


def get_samplers(rank: int, world_size: int, dataset_train, dataset_test, **kwargs: Dict[str, Any]):
    mk = lambda ds: tc.utils.data.distributed.DistributedSampler(  # noqa: E731
        ds, num_replicas=world_size, rank=rank, shuffle=True, seed=0, drop_last=False)
    return {"sampler_train": mk(dataset_train), "sampler_test": mk(dataset_test)}


def get_dataloaders(batch_size: int, num_microbatches: int, world_size: int, dataset_train, dataset_test,
                    sampler_train, sampler_test, **kwargs: Dict[str, Any]):
    per_rank = batch_size // (num_microbatches * world_size)

    def mk(ds, sm, seed):
        if isinstance(ds, DeviceDataset):
            return DeviceDataLoader(ds, per_rank, sm, seed=seed + 1000 * getattr(sm, "rank", 0))
        return tc.utils.data.DataLoader(ds, batch_size=per_rank, sampler=sm, num_workers=0,
                                        pin_memory=tc.cuda.is_available(), drop_last=False)

    return {"dl_train": mk(dataset_train, sampler_train, 1), "dl_test": mk(dataset_test, sampler_test, 2)}
