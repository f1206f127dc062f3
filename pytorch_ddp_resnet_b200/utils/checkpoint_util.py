"""
Checkpointing with the reference's on-disk format (resnet/utils/checkpoint_util.py:16-220): one file
per checkpointable kind named `{kind}_{steps}.pth` holding its state_dict, the newest five kept per
kind, resume from the newest step that all kinds share. Host-side glue only (not on the hot path);
file names, state_dict keys (`_batch_step`, `_epoch_step`, `_lowest_loss`) and step numbering are
kept so checkpoints written by either implementation load in the other.
"""
import os
import re
from typing import Any, Dict, List, Optional

import torch as tc

KEEP = 5
_NAME = re.compile(r"^(\w+)_(\d+)\.([a-z]+)$")


def _list(checkpoint_dir: str, kind: str) -> List[int]:
    """Sorted step numbers of the files of `kind` in the directory."""
    steps = set()
    for fname in os.listdir(checkpoint_dir):
        m = _NAME.match(fname)
        if m and m.group(1) == kind:
            steps.add(int(m.group(2)))
    return sorted(steps)


def _path(checkpoint_dir: str, kind: str, steps: int) -> str:
    return os.path.join(checkpoint_dir, f"{kind}_{steps}.pth")


def save_checkpoint(checkpoint_dir: str, kind_name: str, checkpointable, steps: int) -> None:
    os.makedirs(checkpoint_dir, exist_ok=True)
    tc.save(checkpointable.state_dict(), _path(checkpoint_dir, kind_name, steps))
    for old in _list(checkpoint_dir, kind_name)[:-KEEP]:
        os.remove(_path(checkpoint_dir, kind_name, old))


def maybe_load_checkpoint(checkpoint_dir: str, kind_name: str, checkpointable, map_location,
                          steps: Optional[int]) -> int:
    os.makedirs(checkpoint_dir, exist_ok=True)
    if steps is None:
        have = _list(checkpoint_dir, kind_name)
        steps = have[-1] if have else None
    path = _path(checkpoint_dir, kind_name, steps) if steps is not None else None
    if path is None or not os.path.exists(path):
        print(f"Bad {kind_name} checkpoint or none at {checkpoint_dir} with step {steps}.")
        print("Running from scratch.")
        return 0
    checkpointable.load_state_dict(tc.load(path, map_location=map_location))
    print(f"Loaded {kind_name} checkpoint from {checkpoint_dir}, with step {steps}.")
    print("Continuing from checkpoint.")
    return steps


def maybe_load_checkpoints(checkpoint_dir: str, checkpointables: Dict[str, Optional[Any]], map_location,
                           steps: Optional[int]) -> int:
    """Loads every non-None checkpointable; all kinds must agree on the step. Returns it (0 = fresh)."""
    found = [maybe_load_checkpoint(checkpoint_dir, kind, obj, map_location, steps)
             for kind, obj in checkpointables.items() if obj is not None]
    if len(set(found)) != 1:
        raise RuntimeError("Checkpoint steps not aligned.")
    return found[0]


def save_checkpoints(checkpoint_dir: str, checkpointables: Dict[str, Optional[Any]], steps: int) -> None:
    for kind, obj in checkpointables.items():
        if obj is not None:
            save_checkpoint(checkpoint_dir, kind, obj, steps)


class CheckpointStrategy(tc.nn.Module):
    """Counts batches / epochs seen (as buffers, so the counters are checkpointed too)."""

    def __init__(self, unit: str):
        if unit not in ("batch", "epoch"):
            raise ValueError("unit must be 'batch' or 'epoch'")
        super().__init__()
        self._unit = unit
        self.register_buffer("_batch_step", tc.tensor(0))
        self.register_buffer("_epoch_step", tc.tensor(0))

    @property
    def unit(self) -> str:
        return self._unit

    @property
    def batch_step(self) -> int:
        return int(self._batch_step.item())

    @property
    def epoch_step(self) -> int:
        return int(self._epoch_step.item())

    def step(self, unit: str) -> None:
        counter = {"batch": self._batch_step, "epoch": self._epoch_step}[unit]
        counter += 1

    def observe(self, **kwargs) -> bool:
        raise NotImplementedError


class FrequencyCheckpointStrategy(CheckpointStrategy):
    def __init__(self, unit, frequency, **kwargs):
        super().__init__(unit)
        self._frequency = frequency

    def observe(self, unit, **kwargs) -> bool:
        due = getattr(self, f"{unit}_step") % self._frequency == 0
        self.step(unit)
        return bool(due) if unit == self.unit else False


class PerformanceCheckpointStrategy(CheckpointStrategy):
    def __init__(self, unit, **kwargs):
        super().__init__(unit)
        self.register_buffer("_lowest_loss", tc.tensor(float("inf")))

    @property
    def lowest_loss(self) -> float:
        return float(self._lowest_loss.item())

    def observe(self, unit, loss, **kwargs) -> bool:
        better = loss < self.lowest_loss
        self.step(unit)
        if unit != self.unit:
            return False
        if better:
            self._lowest_loss.fill_(float(loss))
        return bool(better)


_STRATEGIES = {c.__name__: c for c in (FrequencyCheckpointStrategy, PerformanceCheckpointStrategy)}


def get_checkpoint_strategy(checkpoint_strategy_cls_name: str,
                            checkpoint_strategy_args: Optional[Dict[str, Any]]) -> CheckpointStrategy:
    return _STRATEGIES[checkpoint_strategy_cls_name](**(checkpoint_strategy_args or {}))
