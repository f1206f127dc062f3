"""
pytorch_ddp_resnet_b200 — the ResNet / Wide-ResNet data-parallel training step of
lucaslingle/pytorch_ddp_resnet rebuilt for NVIDIA B200 (sm_100a): hand-written tcgen05/TMA conv
kernels and fused HBM-bound kernels behind a C ABI (include/b200resnet.h), driven by thin PyTorch
host code that keeps the reference's public API.
"""
__version__ = "0.1.0"
