"""
DistributedDataParallel wiring (reference: script.py:53-71 wraps the model with all DDP defaults).

wrap_ddp() keeps that behaviour (bucketed NCCL gradient all-reduce overlapped with backward, rank-0
buffers broadcast before each forward, `module.` prefix in the state_dict) and adds what whole-step
CUDA-graph capture needs: the wrapper is constructed in a side-stream context, so that nothing DDP
enqueues later lands on the legacy default stream while a capture is running.
"""
import os

import torch


def prepare_env_for_graphs() -> None:
    """Call before init_process_group when the training step will be captured in a CUDA graph."""
    os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
    os.environ.setdefault("NCCL_ASYNC_ERROR_HANDLING", "0")


def wrap_ddp(module: torch.nn.Module, device: torch.device, **ddp_kwargs):
    """DistributedDataParallel(module) built on a side stream (safe for later graph capture)."""
    kwargs = dict(gradient_as_bucket_view=True)
    kwargs.update(ddp_kwargs)
    if device.type != "cuda":
        return torch.nn.parallel.DistributedDataParallel(module, **kwargs)
    side = torch.cuda.Stream(device=device)
    side.wait_stream(torch.cuda.current_stream(device))
    with torch.cuda.stream(side):
        ddp = torch.nn.parallel.DistributedDataParallel(module, device_ids=[device.index], **kwargs)
    torch.cuda.current_stream(device).wait_stream(side)
    return ddp
