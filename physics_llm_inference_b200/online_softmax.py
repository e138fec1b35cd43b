"""ch06's online-softmax functions and attention accounting on the B200 path.

  online_softmax(x)                   ch06/online_softmax.py:13-25   same signature, same result (softmax over dim -1)
  online_softmax_with_output(x, v)    ch06/online_softmax.py:28-53   returns (o, d) like the reference
  standard_softmax(x, dim=-1)         ch06/online_softmax.py:4-10    two-pass form, kept for the reference's own tests
  attention_memory_bytes(...)         ch06/attention_memory.py:36-61 -> AttentionMemoryStats (:6-16)
  attention_arithmetic_intensity(...) ch06/attention_memory.py:78-86

The two online functions run as CUDA row kernels behind the C ABI (pli_online_softmax*, csrc/online_softmax.cu): the
same (max, sum, weighted-sum) recurrence the prefill and decode kernels carry per row, exposed element by element.
No CPU path: CPU tensors raise.  The accounting functions are integer arithmetic on the host, as in the reference.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import _lib


@dataclass
class AttentionMemoryStats:
    batch_size: int
    num_heads: int
    seq_len: int
    head_dim: int
    qk_bytes: int
    softmax_bytes: int
    output_bytes: int
    total_bytes: int
    total_mb: float


def attention_memory_bytes(batch_size: int, num_heads: int, seq_len: int, head_dim: int,
                           dtype_bytes: int = 2) -> AttentionMemoryStats:
    """HBM bytes of materialised attention: the N x N scores, the N x N probabilities and the output."""
    square = batch_size * num_heads * seq_len * seq_len * dtype_bytes
    output_bytes = batch_size * num_heads * seq_len * head_dim * dtype_bytes
    total = 2 * square + output_bytes
    return AttentionMemoryStats(batch_size=batch_size, num_heads=num_heads, seq_len=seq_len, head_dim=head_dim,
                                qk_bytes=square, softmax_bytes=square, output_bytes=output_bytes, total_bytes=total,
                                total_mb=total / 1024 / 1024)


def attention_arithmetic_intensity(seq_len: int, head_dim: int) -> float:
    """FLOPs per byte of materialised fp16 attention for one head (two GEMMs + 5 N^2 softmax FLOPs over the score
    matrix written and read once plus Q, K, V)."""
    flops = 4 * seq_len * seq_len * head_dim + 5 * seq_len * seq_len
    bytes_rw = 2 * seq_len * seq_len * 2 + seq_len * head_dim * 2 * 3
    return flops / bytes_rw


def _require_cuda(name: str, x: torch.Tensor) -> None:
    if not isinstance(x, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not x.is_cuda:
        raise RuntimeError(f"{name} is on {x.device}: the B200 path has no CPU fallback "
                           "(the CPU restatement lives in oracle/ and is test infrastructure only)")


def standard_softmax(x: torch.Tensor, dim: int = -1) -> torch.Tensor:
    """Two-pass softmax (max, then exp / sum) on the tensor's own device: the baseline the reference compares with."""
    x_max = x.max(dim=dim, keepdim=True).values
    exp_x = torch.exp(x - x_max)
    return exp_x / exp_x.sum(dim=dim, keepdim=True)


def online_softmax(x: torch.Tensor) -> torch.Tensor:
    """Softmax over the last dimension by the one-pass running (max, sum) recurrence."""
    _require_cuda("x", x)
    if x.dim() < 1 or x.shape[-1] < 1 or x.numel() == 0:
        raise RuntimeError(f"online_softmax needs a non-empty last dimension; got {tuple(x.shape)}")
    n = x.shape[-1]
    x2 = x.reshape(-1, n)
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    out = torch.empty((x2.shape[0], n), dtype=x.dtype, device=x.device)
    lib = _lib.load()
    with _lib.on_device(x.device):
        rc = lib.pli_online_softmax(x2.data_ptr(), out.data_ptr(), x2.shape[0], n, x2.stride(0), out.stride(0),
                                    _lib.dtype_code(x.dtype), _lib.current_stream_ptr(x.device))
    _lib.check(rc)
    return out.view(x.shape)


def online_softmax_with_output(x: torch.Tensor, v: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """x (..., n), v (..., n, d_v) -> (o (..., d_v) = softmax(x) @ v, d (...) = sum_i exp(x_i - max x)), both by the
    running recurrence; d has x's dtype like the reference's."""
    _require_cuda("x", x)
    _require_cuda("v", v)
    if x.dtype != v.dtype or x.device != v.device:
        raise RuntimeError("x and v must share dtype and device")
    if v.dim() != x.dim() + 1 or v.shape[:-1] != x.shape or x.shape[-1] < 1 or x.numel() == 0:
        raise RuntimeError(f"expected x (..., n) and v (..., n, d_v); got {tuple(x.shape)} and {tuple(v.shape)}")
    n, dv = x.shape[-1], v.shape[-1]
    x2 = x.reshape(-1, n)
    v2 = v.reshape(-1, n, dv)
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    if v2.stride(-1) != 1:
        v2 = v2.contiguous()
    rows = x2.shape[0]
    o = torch.empty((rows, dv), dtype=x.dtype, device=x.device)
    d = torch.empty((rows,), dtype=torch.float32, device=x.device)
    lib = _lib.load()
    with _lib.on_device(x.device):
        rc = lib.pli_online_softmax_with_output(x2.data_ptr(), v2.data_ptr(), o.data_ptr(), d.data_ptr(), rows, n, dv,
                                                x2.stride(0), v2.stride(0), v2.stride(1), o.stride(0),
                                                _lib.dtype_code(x.dtype), _lib.current_stream_ptr(x.device))
    _lib.check(rc)
    return o.view(*x.shape[:-1], dv), d.to(x.dtype).view(x.shape[:-1])
