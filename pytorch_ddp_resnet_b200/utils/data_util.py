"""
Datasets / samplers / loaders (reference: resnet/utils/data_util.py:21-232). The reference's CPU
augmentation pipeline (transform_util.py) is outside the hot path and is not rebuilt; this module
keeps the call surface script.py needs:

  * dataset_cls_name 'SyntheticCIFAR10' / 'SyntheticCIFAR100' / 'SyntheticImageNet': seeded in-memory
    tensors (a learnable task: class pattern + noise), no files, no network;
  * any torchvision dataset class name, read from `data_dir` WITHOUT downloading, ToTensor only.
The batch-size rule is the reference's: config batch_size is GLOBAL; a loader yields
batch_size // (num_microbatches * world_size) samples (data_util.py:216).
"""
from typing import Any, Dict

import torch as tc

_SYNTHETIC = {"SyntheticCIFAR10": (10, 32), "SyntheticCIFAR100": (100, 32), "SyntheticImageNet": (1000, 224)}


def _synthetic(name: str, train: bool, size: int):
    classes, hw = _SYNTHETIC[name]
    g = tc.Generator().manual_seed(0)
    pattern = tc.randn(classes, 3, hw, hw, generator=g)          # shared by train and test
    g = tc.Generator().manual_seed(1 if train else 2)
    y = tc.randint(0, classes, (size,), generator=g)
    x = 0.3 * pattern[y] + tc.randn(size, 3, hw, hw, generator=g)
    return tc.utils.data.TensorDataset(x, y)


def get_datasets(dataset_cls_name: str, data_dir: str, **kwargs: Dict[str, Any]):
    if dataset_cls_name in _SYNTHETIC:
        n_train = int(kwargs.get("synthetic_train_size", 2048))
        n_test = int(kwargs.get("synthetic_test_size", 512))
        return {"dataset_train": _synthetic(dataset_cls_name, True, n_train),
                "dataset_test": _synthetic(dataset_cls_name, False, n_test)}
    import torchvision
    cls = getattr(torchvision.datasets, dataset_cls_name)
    tf = torchvision.transforms.ToTensor()
    return {"dataset_train": cls(root=data_dir, train=True, download=False, transform=tf),
            "dataset_test": cls(root=data_dir, train=False, download=False, transform=tf)}


def get_samplers(rank: int, world_size: int, dataset_train, dataset_test, **kwargs: Dict[str, Any]):
    mk = lambda ds: tc.utils.data.distributed.DistributedSampler(  # noqa: E731
        ds, num_replicas=world_size, rank=rank, shuffle=True, seed=0, drop_last=False)
    return {"sampler_train": mk(dataset_train), "sampler_test": mk(dataset_test)}


def get_dataloaders(batch_size: int, num_microbatches: int, world_size: int, dataset_train, dataset_test,
                    sampler_train, sampler_test, **kwargs: Dict[str, Any]):
    per_rank = batch_size // (num_microbatches * world_size)
    mk = lambda ds, sm: tc.utils.data.DataLoader(  # noqa: E731
        ds, batch_size=per_rank, sampler=sm, num_workers=0, pin_memory=tc.cuda.is_available(), drop_last=False)
    return {"dl_train": mk(dataset_train, sampler_train), "dl_test": mk(dataset_test, sampler_test)}
