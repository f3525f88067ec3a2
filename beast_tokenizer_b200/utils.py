"""Tensor helpers with the reference's names and arithmetic (beast/utils.py:4-44).

The tokenizer's own encode/decode never call these — the same arithmetic is fused into the
CUDA kernels (csrc/common.cuh quantize_one / dequantize_one).  They remain for callers that
import them directly and for the small normalise / de-normalise steps of the continuous API.
"""
import torch


def continuous_to_discrete(tensor, min_val=None, max_val=None, num_bins=256):
    if min_val is None:
        min_val = tensor.min()
    if max_val is None:
        max_val = tensor.max()
    scale = torch.clamp(max_val - min_val, min=1e-8)
    normalized = torch.clamp((tensor - min_val) / scale, 0, 1)
    return torch.round(normalized * (num_bins - 1)).to(torch.long)


def discrete_to_continuous(discrete_tensor, min_val=0, max_val=1, num_bins=256):
    normalized = discrete_tensor.float() / (num_bins - 1)
    return torch.clamp(normalized * (max_val - min_val) + min_val, min_val, max_val)


def normalize_tensor(tensor, w_min, w_max, norm_min=-1.0, norm_max=1.0):
    clipped = torch.clamp(tensor, w_min, w_max)
    normalized = (clipped - w_min) / torch.clamp(w_max - w_min, min=1e-8)
    return normalized * (norm_max - norm_min) + norm_min


def denormalize_tensor(normalized_tensor, w_min, w_max, norm_min=-1.0, norm_max=1.0):
    # upstream clamps the Python float (norm_max - norm_min) with torch.clamp, which raises
    # TypeError (beast/utils.py:42); max() is the working equivalent.
    clipped = torch.clamp(normalized_tensor, norm_min, norm_max)
    denormalized = (clipped - norm_min) / max(norm_max - norm_min, 1e-8)
    return denormalized * (w_max - w_min) + w_min
