"""Evaluation loop (reference: resnet/algos/evaluation.py:14-42): eval-mode forward over the test
loader (BN from running statistics, dropout off), per-batch metric means, global mean over ranks.
Metrics stay on the device until the single host read at the end.

Precision: the reference evaluates WITHOUT autocast (evaluation.py:32-39), i.e. fp32 tensors whose convolutions
run on TF32 tensor cores. `eval_precision: tf32` (the default) does the same here: fp32 activations, kind::tf32
convolutions on the fp32 master weights, fp32 BN / head / loss. `eval_precision: bf16` reuses the training
kernels (about twice as fast, logits within bf16 rounding of the fp32 ones).
`fold_bn: True` (default): every batch norm whose only input is a convolution is folded into that convolution
and its ReLU runs in the conv epilogue (utils/fold_util.py): one launch instead of two per such pair."""
from typing import Any, Dict

import torch as tc

from pytorch_ddp_resnet_b200 import ops
from pytorch_ddp_resnet_b200.algos.metrics import compute_losses_and_metrics, global_means


@tc.no_grad()
def evaluation_loop(world_size: int, device, dl_test, classifier, **kwargs: Dict[str, Any]) -> Dict[str, float]:
    """
    Evaluates classifier on the validation/test set.

    :param world_size: World size.
    :param device: Device.
    :param dl_test: Test/val dataloader.
    :param classifier: Classifier.
    :return: Dictionary of global metric floats, keyed by name.
    """
    classifier.eval()
    sums, num_batch = None, 0
    mode = kwargs.get("eval_precision", "tf32") if tc.device(device).type == "cuda" else "bf16"
    for x, y in dl_test:
        x, y = x.to(device, non_blocking=True), y.to(device, non_blocking=True)
        with ops.precision(mode), ops.fold_bn(bool(kwargs.get("fold_bn", True))):
            m = compute_losses_and_metrics(logits=classifier(x), labels=y)
        vec = tc.stack([m["loss"].float(), m["top1_err"].float(), m["top5_err"].float()])
        sums = vec if sums is None else sums + vec
        num_batch += 1
    if num_batch == 0:
        return {}
    sums = sums / num_batch
    return global_means({"loss": sums[0], "top1_err": sums[1], "top5_err": sums[2]}, world_size)
