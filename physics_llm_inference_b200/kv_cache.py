"""Device-side mirrors of the reference's contiguous KV caches.

Same names, fields, layout and update semantics as
  * `KVCache`       ch02/kv_cache.py:9-51        (k_cache/v_cache (B, max_seq_len, Hkv, D), seq_len)
  * `LayerKVCache`  ch02/cached_generation.py:20-33 (k/v, seq_len)
so code written against the reference keeps working; the slice-assign of `update`
(ch02/kv_cache.py:45-46, ch02/cached_generation.py:30-31) runs as the `pli_kv_append` kernel.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import _lib


def kv_append(k_store: torch.Tensor, v_store: torch.Tensor, k_new: torch.Tensor, v_new: torch.Tensor,
              start_pos, *, block_tables: torch.Tensor | None = None, layer: int = 0, validate: bool = False) -> None:
    """Write k_new/v_new (B, n, Hkv, D) at positions start_pos[b] + i of each sequence.

    Contiguous storage (B, L, Hkv, D) when block_tables is None, else paged pools
    (P, layers, bs, Hkv, D) addressed through block_tables (B, max_pages) int32: token t goes to
    page block_tables[b, t // bs], slot t % bs (ch07/paged_memory.py:54,84-86).
    start_pos: int or (B,) int32 device tensor.  An int start is bound-checked on the host; a device tensor is only
    checked with validate=True (one synchronising copy): start_pos[b] + n must fit the cache / the table's pages.
    """
    if not (k_store.is_cuda and k_new.is_cuda):
        raise RuntimeError("kv_append runs on CUDA tensors only (no CPU fallback)")
    if k_new.dim() != 4 or k_new.shape != v_new.shape:
        raise RuntimeError(f"k_new/v_new must both be (B, n, Hkv, D); got {tuple(k_new.shape)} {tuple(v_new.shape)}")
    if not (k_store.dtype == v_store.dtype == k_new.dtype == v_new.dtype):
        raise RuntimeError("KV storage and new K/V must share a dtype")
    B, n, Hkv, D = k_new.shape
    if n == 0:
        return
    dev = k_store.device
    if isinstance(start_pos, int):
        start = torch.full((B,), start_pos, dtype=torch.int32, device=dev)
    else:
        start = start_pos.to(device=dev, dtype=torch.int32)
    if k_new.stride(-1) != 1:
        k_new = k_new.contiguous()
    if v_new.stride(-1) != 1:
        v_new = v_new.contiguous()
    if v_new.stride() != k_new.stride():
        v_new = v_new.contiguous()
        k_new = k_new.contiguous()
    if k_store.stride() != v_store.stride() or k_store.stride(-1) != 1:
        raise RuntimeError("k and v storage must have identical strides and a unit head_dim stride")
    if block_tables is None:
        if k_store.dim() != 4 or k_store.shape[0] != B or k_store.shape[2:] != (Hkv, D):
            raise RuntimeError(f"contiguous cache {tuple(k_store.shape)} does not match new K/V {tuple(k_new.shape)}")
        st = (k_store.stride(0), 0, k_store.stride(1), k_store.stride(2))
        table_ptr, bs, tstride = None, 0, 0
    else:
        if k_store.dim() != 5 or k_store.shape[3:] != (Hkv, D):
            raise RuntimeError(f"paged pool {tuple(k_store.shape)} does not match new K/V {tuple(k_new.shape)}")
        if block_tables.dtype != torch.int32 or not block_tables.is_cuda or block_tables.stride(-1) != 1:
            raise RuntimeError("block_tables must be a CUDA int32 tensor with unit inner stride")
        st = k_store.stride()[:4]
        table_ptr, bs, tstride = block_tables.data_ptr(), k_store.shape[2], block_tables.stride(0)
    capacity = k_store.shape[1] if block_tables is None else block_tables.shape[1] * k_store.shape[2]
    if isinstance(start_pos, int):
        if start_pos < 0 or start_pos + n > capacity:
            raise RuntimeError(f"KV append of {n} token(s) at position {start_pos} does not fit the capacity {capacity}")
    elif validate:
        hi, lo = int(start.max()), int(start.min())
        if lo < 0 or hi + n > capacity:
            raise ValueError(f"KV append of {n} token(s) at positions [{lo}, {hi}] does not fit the capacity {capacity}")
    lib = _lib.load()
    with _lib.on_device(dev):
        rc = lib.pli_kv_append(k_new.data_ptr(), v_new.data_ptr(), k_store.data_ptr(), v_store.data_ptr(), table_ptr,
                               start.data_ptr(), B, n, Hkv, D, bs, tstride, layer, _lib.i64(*k_new.stride()[:3]),
                               _lib.i64(*st), _lib.dtype_code(k_store.dtype), _lib.current_stream_ptr(dev))
    _lib.check(rc)


@dataclass
class KVCache:
    """ch02/kv_cache.py:9-51."""
    k_cache: torch.Tensor
    v_cache: torch.Tensor
    seq_len: int

    @classmethod
    def create(cls, batch_size: int, max_seq_len: int, num_kv_heads: int, head_dim: int,
               device: torch.device, dtype: torch.dtype) -> "KVCache":
        shape = (batch_size, max_seq_len, num_kv_heads, head_dim)
        return cls(k_cache=torch.zeros(shape, device=device, dtype=dtype),
                   v_cache=torch.zeros(shape, device=device, dtype=dtype), seq_len=0)

    def update(self, k: torch.Tensor, v: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        new_tokens = k.shape[1]
        start, end = self.seq_len, self.seq_len + new_tokens
        if end > self.k_cache.shape[1]:
            raise RuntimeError(f"KV cache overflow: {end} > max_seq_len {self.k_cache.shape[1]}")
        kv_append(self.k_cache, self.v_cache, k, v, start)
        self.seq_len = end
        return self.k_cache[:, :end], self.v_cache[:, :end]

    def memory_bytes(self) -> int:
        return self.k_cache.numel() * self.k_cache.element_size() * 2


@dataclass
class LayerKVCache:
    """ch02/cached_generation.py:20-33."""
    k: torch.Tensor  # (batch, max_seq_len, num_kv_heads, head_dim)
    v: torch.Tensor
    seq_len: int = 0

    def update(self, k_new: torch.Tensor, v_new: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        new_len = k_new.shape[1]
        if self.seq_len + new_len > self.k.shape[1]:
            raise RuntimeError(f"KV cache overflow: {self.seq_len + new_len} > max_seq_len {self.k.shape[1]}")
        kv_append(self.k, self.v, k_new, v_new, self.seq_len)
        self.seq_len += new_len
        return self.k[:, :self.seq_len], self.v[:, :self.seq_len]


def create_caches(num_layers: int, batch_size: int, max_seq_len: int, num_kv_heads: int, head_dim: int,
                  device, dtype) -> list[LayerKVCache]:
    """One zero-initialised LayerKVCache per layer (ch02/cached_generation.py:189-205)."""
    caches = []
    for _ in range(num_layers):
        k = torch.zeros(batch_size, max_seq_len, num_kv_heads, head_dim, device=device, dtype=dtype)
        caches.append(LayerKVCache(k=k, v=torch.zeros_like(k), seq_len=0))
    return caches
