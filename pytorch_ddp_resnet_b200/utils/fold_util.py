"""
Eval-mode batch-norm folding (reference: resnet/algos/evaluation.py:14-42 evaluates the same
modules it trains, so every `conv -> BatchNorm2d(eval) [-> ReLU]` pair costs it a conv and two more passes).

In eval mode a batch norm is the per-channel affine map y = s * x + t with s = gamma / sqrt(running_var + eps)
and t = beta - running_mean * s. When the ONLY consumer of a convolution's output is such a batch norm, the pair
equals one convolution with filter s[k] * W[k] and bias t[k] (+ s[k] * b[k] if the conv has a bias), and the
ReLU that follows runs in the conv kernel's epilogue: the `bn_act_fwd` launch (a read and a write of the whole
activation) disappears.

Foldable pairs in this model family (the blocks decide, architectures/residual_block.py):
  * post-activation blocks (v1): every conv_j -> norm_j (the last one adds the shortcut before its ReLU, which the
    epilogue does as well);
  * pre-activation blocks: conv_j -> norm_{j+1} inside a block (norm_1 reads the block input, which the shortcut
    also reads, and the last conv's output is summed with the shortcut: neither can fold);
  * a top-level `c...` conv directly followed by `n [a]`.
Switched on by `with ops.fold_bn(True)` (the default of evaluation_loop); numerically it differs from the unfolded
evaluation only by where the bf16 (or TF32) rounding of the filter happens.
"""
import torch

from pytorch_ddp_resnet_b200.architectures import layers


def folded_filter(conv, bn, f32: bool):
    """(filter [K,R,S,C] in the precision of the activation path, fp32 bias [K]) of conv followed by eval-mode bn."""
    key = (conv.weight.data_ptr(), conv.weight._version, None if conv.bias is None else conv.bias._version,
           bn.weight._version, bn.bias._version, bn.running_mean._version, bn.running_var._version,
           layers._weight_generation[0], f32)
    cache = getattr(conv, "_fold_cache", None)
    if cache is not None and cache[0] == key and cache[1] is bn:
        return cache[2], cache[3]
    with torch.no_grad():
        s = bn.weight.float() * torch.rsqrt(bn.running_var.float() + bn.eps)
        t = bn.bias.float() - bn.running_mean.float() * s
        if conv.bias is not None:
            t = t + conv.bias.float() * s
        w = conv.krsc().float() * s[:, None, None, None]
        w = w.contiguous() if f32 else w.to(torch.bfloat16).contiguous()
        t = t.contiguous()
    conv._fold_cache = (key, bn, w, t)
    return w, t
