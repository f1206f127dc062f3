"""
Training loop (reference: resnet/algos/training.py:31-171), same signature and observable behaviour:
epochs over the train loader, gradient accumulation over `num_microbatches` (summed, as in the
reference), optimizer / scheduler stepping, rank-0 printing + TensorBoard scalars, checkpoint
triggers, evaluation after every epoch.

What differs is below the API: the model computes in bf16 on the sm_100a kernels on its own (no
torch autocast context is needed; a GradScaler is still honoured if one is passed), and the three
logged metrics travel to the host in one packed all-reduce + one read per microbatch instead of three.
With `lagged_metrics: True` (config key; implied by `cuda_graph: True`) that read does not block either:
the metrics of step i are copied to pinned host memory asynchronously and printed / written to TensorBoard
while step i+1 runs (the reference syncs the device three times per step, metrics.py:32-41). Anything that
needs the loss of the CURRENT step on the host - a ReduceLROnPlateau stepped per batch, a
PerformanceCheckpointStrategy per batch - switches the lag off.
"""
from collections import Counter
from typing import Any, Dict, Optional, Union

import torch as tc

from pytorch_ddp_resnet_b200.algos.evaluation import evaluation_loop
from pytorch_ddp_resnet_b200.algos.metrics import LaggedMetrics, compute_losses_and_metrics, global_means
from pytorch_ddp_resnet_b200.utils.checkpoint_util import FrequencyCheckpointStrategy, save_checkpoints


def requires_loss(scheduler) -> bool:
    return isinstance(scheduler, tc.optim.lr_scheduler.ReduceLROnPlateau)


def step_scheduler(scheduler, loss: Union[tc.Tensor, float]) -> None:
    if requires_loss(scheduler):
        scheduler.step(loss)
    else:
        scheduler.step()


def _make_writer(log_dir: Optional[str]):
    if not log_dir:
        return None
    try:
        from torch.utils.tensorboard import SummaryWriter
    except Exception:  # tensorboard missing: logging is not on the hot path
        return None
    return SummaryWriter(log_dir)


def training_loop(
        rank: int,
        world_size: int,
        device,
        sampler_train,
        sampler_test,
        dl_train,
        dl_test,
        classifier,
        optimizer,
        scaler,
        scheduler,
        scheduler_step_unit: str,
        checkpoint_strategy,
        checkpoint_dir: str,
        num_microbatches: int,
        global_step: int,
        max_steps: int,
        log_dir: str,
        **kwargs: Dict[str, Any]
) -> None:
    """Runs a training loop on a given process; see the reference docstring for the arguments."""
    writer = _make_writer(log_dir) if rank == 0 else None
    # cuda_graph=True (config key, optional): capture the whole step once and replay it
    use_graph = bool(kwargs.get("cuda_graph", False)) and scaler is None and num_microbatches == 1
    graphed = None
    checkpointables = {
        'checkpoint_strategy': checkpoint_strategy,
        'classifier': classifier,
        'optimizer': optimizer,
        'scheduler': scheduler,
        'scaler': scaler
    }
    epoch = int(checkpoint_strategy.epoch_step)
    # metrics read one step late (no host sync in the step): only when nothing consumes the current loss
    loss_needed_now = (scheduler is not None and scheduler_step_unit == 'batch' and requires_loss(scheduler)) or \
        (checkpoint_strategy.unit == 'batch' and not isinstance(checkpoint_strategy, FrequencyCheckpointStrategy))
    lagged = bool(kwargs.get("lagged_metrics", use_graph)) and num_microbatches == 1 and not loss_needed_now
    lag = LaggedMetrics(world_size) if lagged else None

    def report(step: int, logged) -> None:
        if rank == 0:
            print(f"global step: {step}... loss: {logged.get('loss')}")
            if writer:
                for name, value in logged.items():
                    writer.add_scalar(tag=f"train/{name}", scalar_value=value, global_step=step)

    def drain(everything: bool) -> None:
        while lag is not None and len(lag) and (everything or lag.ready()):
            report(*lag.pop())

    while global_step < max_steps:
        if hasattr(sampler_train, "set_epoch"):
            sampler_train.set_epoch(epoch)
        classifier.train()
        acc = Counter()
        for microbatch_id, (x, y) in enumerate(dl_train, 1):
            x, y = x.to(device, non_blocking=True), y.to(device, non_blocking=True)
            if use_graph:
                if graphed is None:
                    from pytorch_ddp_resnet_b200.utils.graph_util import GraphedTrainStep
                    graphed = GraphedTrainStep(classifier, optimizer, x, y)
                metrics = graphed(x, y)
                if lagged:
                    lag.push(global_step, metrics)
                else:
                    acc += global_means(metrics, world_size)
            else:
                metrics = compute_losses_and_metrics(logits=classifier(x), labels=y)
                loss = metrics['loss']
                (scaler.scale(loss) if scaler else loss).backward()
                if lagged:
                    lag.push(global_step, metrics)
                else:
                    acc += global_means(metrics, world_size)

                if microbatch_id % num_microbatches != 0:
                    continue
                if scaler:
                    scaler.step(optimizer)
                    scaler.update()
                else:
                    optimizer.step()
                optimizer.zero_grad(set_to_none=True)

            global_loss = None
            if lagged:
                drain(everything=False)     # prints step i-1 while step i runs on the device
            else:
                logged = {k: v / num_microbatches for k, v in acc.items()}
                global_loss = logged.get('loss')
                report(global_step, logged)
            if scheduler and scheduler_step_unit == 'batch':
                step_scheduler(scheduler, global_loss)
            if rank == 0 and checkpoint_strategy.observe(unit='batch', loss=global_loss):
                save_checkpoints(checkpoint_dir=checkpoint_dir, checkpointables=checkpointables,
                                 steps=global_step + 1)
            acc = Counter()
            global_step += 1
            if global_step >= max_steps:
                break
        drain(everything=True)

        val = evaluation_loop(world_size, device, dl_test, classifier)
        val_loss = val.get('loss')
        if scheduler and scheduler_step_unit == 'epoch':
            step_scheduler(scheduler, val_loss)
        if rank == 0:
            print(f"epoch: {epoch}... validation loss: {val_loss}")
            if writer:
                for name, value in val.items():
                    writer.add_scalar(tag=f"val/{name}", scalar_value=value, global_step=epoch)
            if checkpoint_strategy.observe(unit='epoch', loss=val_loss):
                save_checkpoints(checkpoint_dir=checkpoint_dir, checkpointables=checkpointables,
                                 steps=global_step + 1)
        # every rank advances its epoch (the reference advances it on rank 0 only, so ranks > 0 keep
        # reshuffling with epoch 0: SURVEY.md Q10)
        epoch += 1
