"""Drop-in for the reference's `ch06.flash_attention_forward` on B200.

Keeps the reference signature and layouts (ch06/flash_attention.py:14-20: `q, k, v` of shape
(B, H, N, D), `scale=None` -> D**-0.5, `config=None`, returns a tensor shaped and typed like q)
and adds, keyword-only, what BASELINE.json's north_star asks for on top of ch06:

  causal=False      ch01/gqa.py:33-34 / ch02/cached_generation.py:85-91 mask (bottom-right aligned
                    when Nq < Nk); the default stays non-causal like the reference.
  return_lse=False  also return log-sum-exp (B, Hq, Nq) float32 (ch06 computes row_max/row_sum at
                    :71-72 and drops them).
  k, v may have Hkv heads with Hq % Hkv == 0: q-head h reads kv-head h // (Hq // Hkv)
                    (ch01/gqa.py:14,30-31) without materialising the repeat_interleave copy.

All arithmetic runs in the CUDA library behind include/pli_attention.h; there is no CPU path.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import _lib


@dataclass
class FlashAttentionConfig:
    """Tile hints of ch06/flash_attention.py:6-11.  The sm_100a kernels use fixed 128x128 tiles, so
    the values are accepted for compatibility and do not change the result (in the reference they
    do not change it either: any block_q/block_k computes the same function)."""
    block_q: int = 64
    block_k: int = 64
    num_warps: int = 4
    num_stages: int = 2


def _unit_inner(x: torch.Tensor) -> torch.Tensor:
    """Views the tensor-core kernels can read in place: unit head_dim stride and no broadcast (zero) stride on a
    dimension larger than one (TMA cannot broadcast; `k.expand(...)` would otherwise fall to the slow SIMT kernel)."""
    if x.stride(-1) != 1:
        return x.contiguous()
    for size, stride in zip(x.shape[:-1], x.stride()[:-1]):
        if size > 1 and stride == 0:
            return x.contiguous()
    return x


def validate_paged_args(seq_lens: torch.Tensor, block_tables: torch.Tensor, block_size: int, num_pages: int, *,
                        q_lens=None, min_len: int = 1, max_seq_len: int | None = None) -> None:
    """Opt-in (`validate=True`) host check of the DEVICE-side preconditions of the paged entry points; costs one
    synchronising copy of the small index tensors.  The kernels themselves do not check them (no error can cross a
    launch): seq_lens[b] in [max(q_len[b], min_len), table width * page size] and <= max_seq_len, and every block-table
    entry a sequence uses inside [0, num_pages)."""
    lens = seq_lens.detach().to("cpu", torch.int64)
    table = block_tables.detach().to("cpu", torch.int64)
    cap = table.shape[1] * block_size
    bound = cap if max_seq_len is None else min(cap, int(max_seq_len))
    if q_lens is None:
        q_lens = torch.zeros_like(lens)
    else:
        q_lens = torch.as_tensor(q_lens, dtype=torch.int64).reshape(-1).expand(lens.shape[0]) if not torch.is_tensor(q_lens) \
            else q_lens.detach().to("cpu", torch.int64)
    for b in range(lens.shape[0]):
        L, nq = int(lens[b]), int(q_lens[b])
        if L < max(nq, min_len):
            raise ValueError(f"sequence {b}: seq_len {L} is smaller than its {nq} query token(s) (rows without a visible key)")
        if L > bound:
            raise ValueError(f"sequence {b}: seq_len {L} exceeds the bound {bound} (block table width x page size, max_seq_len)")
        used = table[b, :(L + block_size - 1) // block_size]
        if used.numel() and (int(used.min()) < 0 or int(used.max()) >= num_pages):
            raise ValueError(f"sequence {b}: a block-table entry in use is outside [0, {num_pages})")


def _check_inputs(q, k, v):
    for name, x in (("q", q), ("k", k), ("v", v)):
        if not isinstance(x, torch.Tensor):
            raise TypeError(f"{name} must be a torch.Tensor")
        if x.dim() != 4:
            raise RuntimeError(f"{name} must have shape (B, H, N, D); got {tuple(x.shape)}")
        if not x.is_cuda:
            raise RuntimeError(
                f"{name} is on {x.device}: this is the B200 CUDA path of flash_attention_forward and has no "
                "CPU fallback (the CPU restatement lives in oracle/ and is test infrastructure only)")
    if not (q.dtype == k.dtype == v.dtype):
        raise RuntimeError(f"q, k, v must share a dtype; got {q.dtype}, {k.dtype}, {v.dtype}")
    if not (q.device == k.device == v.device):
        raise RuntimeError("q, k, v must be on the same device")
    B, Hq, Nq, D = q.shape
    if k.shape != v.shape:
        raise RuntimeError(f"k and v must have the same shape; got {tuple(k.shape)} and {tuple(v.shape)}")
    if k.shape[0] != B or k.shape[3] != D:
        raise RuntimeError(f"k/v shape {tuple(k.shape)} does not match q shape {tuple(q.shape)} in batch or head_dim")
    if Hq % k.shape[1] != 0:
        raise RuntimeError(f"num_heads ({Hq}) must be a multiple of num_kv_heads ({k.shape[1]})")
    if min(B, Hq, Nq, D, k.shape[2]) <= 0:
        raise RuntimeError(f"empty attention problem: q {tuple(q.shape)}, k {tuple(k.shape)}")


def flash_attention_forward(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    scale: float | None = None,
    config: FlashAttentionConfig | None = None,
    *,
    causal: bool = False,
    return_lse: bool = False,
    peer_out=None,
):
    """softmax(Q K^T * scale [+ causal mask]) V on the GPU; see the module docstring.

    peer_out (a `sharding.PeerOutput` built with seq_len=Nq): q, k, v are this rank's head shard; every finished O
    tile is TMA-stored into ALL ranks' full (B, Hq_total, Nq, D) buffers over NVLink, and the call returns that
    full tensor once every rank's tiles have landed (all-gather fused into the kernel, no NCCL call)."""
    if config is None:
        config = FlashAttentionConfig()
    _check_inputs(q, k, v)
    B, Hq, Nq, D = q.shape
    Hkv, Nk = k.shape[1], k.shape[2]
    if causal and Nq > Nk:
        raise ValueError(f"causal attention needs Nq <= Nk (got Nq={Nq}, Nk={Nk})")
    if scale is None:
        scale = D ** -0.5
    q, k, v = _unit_inner(q), _unit_inner(k), _unit_inner(v)
    lse = torch.empty((B, Hq, Nq), dtype=torch.float32, device=q.device) if return_lse else None
    lib = _lib.load()
    if peer_out is not None:
        import ctypes
        sh = peer_out.shard
        if len(peer_out.shape) != 4 or peer_out.dtype != q.dtype or \
                (B, Hq, Nq, D) != (sh.b_end - sh.b_start, sh.q_end - sh.q_start, peer_out.shape[2], peer_out.shape[3]):
            raise RuntimeError(f"local q {tuple(q.shape)} / dtype does not match the PeerOutput shard {peer_out.shape}")
        peer_out.use_mode("scatter")
        Bt, Ht, _, _ = peer_out.shape
        ps = _lib.PeerScatter()
        ps.n_peers, ps.rank = sh.world_size, sh.rank
        for r in range(sh.world_size):
            ps.peer_o[r] = peer_out.output_ptrs[r]
            ps.peer_flags[r] = peer_out.flag_ptrs[r]
        ps.epoch = peer_out.epoch_ptr
        ps.buffer_stride = peer_out.buffer_stride
        ps.slice_offset = peer_out.slice_offset
        with _lib.on_device(q.device):
            stream = _lib.current_stream_ptr(q.device)
            rc = lib.pli_prefill_fwd_scatter(
                q.data_ptr(), k.data_ptr(), v.data_ptr(), lse.data_ptr() if lse is not None else None, B, Hq, Hkv, Nq, Nk,
                D, _lib.i64(*q.stride()[:3]), _lib.i64(*k.stride()[:3]), _lib.i64(*v.stride()[:3]),
                _lib.i64(Ht * Nq * D, Nq * D, D), float(scale), int(bool(causal)), _lib.dtype_code(q.dtype), Bt, Ht,
                sh.b_start, sh.q_start, ctypes.byref(ps), stream)
            _lib.check(rc)
            _lib.check(lib.pli_peer_publish_wait(ctypes.byref(ps), stream))
            o = peer_out.finish_step(ps, stream)
        return (o, lse) if return_lse else o
    out = torch.empty_like(q)
    if out.stride(-1) != 1:
        out = torch.empty(q.shape, dtype=q.dtype, device=q.device)

    with _lib.on_device(q.device):
        rc = lib.pli_prefill_fwd(
            q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), lse.data_ptr() if lse is not None else None,
            B, Hq, Hkv, Nq, Nk, D,
            _lib.i64(*q.stride()[:3]), _lib.i64(*k.stride()[:3]), _lib.i64(*v.stride()[:3]), _lib.i64(*out.stride()[:3]),
            float(scale), int(bool(causal)), _lib.dtype_code(q.dtype), _lib.current_stream_ptr(q.device))
    _lib.check(rc)
    return (out, lse) if return_lse else out


# the name BASELINE.json's north_star uses
flash_attention = flash_attention_forward


def prefill_kernel_kind(q, k, v) -> str:
    """'tcgen05' or 'simt': which kernel family `flash_attention_forward` uses for these tensors."""
    _check_inputs(q, k, v)
    q, k, v = _unit_inner(q), _unit_inner(k), _unit_inner(v)
    out = torch.empty_like(q)
    kind = _lib.load().pli_prefill_kernel_kind(
        q.shape[3], _lib.dtype_code(q.dtype), _lib.i64(*q.stride()[:3]), _lib.i64(*k.stride()[:3]),
        _lib.i64(*v.stride()[:3]), _lib.i64(*out.stride()[:3]), q.data_ptr(), k.data_ptr(), v.data_ptr(),
        out.data_ptr())
    return _lib.KIND_NAMES[kind]


# ---- accounting conventions of ch06/attention_memory.py:64-76 and ch06/flash_attention.py:77-104 ----
def attention_flops(batch_size: int, num_heads: int, seq_len: int, head_dim: int) -> int:
    """2BHN^2D (QK^T) + 5BHN^2 (softmax) + 2BHN^2D (PV), as ch06/attention_memory.py:64-76."""
    qk = 2 * batch_size * num_heads * seq_len * seq_len * head_dim
    return qk + 5 * batch_size * num_heads * seq_len * seq_len + qk


def flash_attention_memory_bytes(batch_size: int, num_heads: int, seq_len: int, head_dim: int,
                                 block_size: int = 64, dtype_bytes: int = 2) -> dict:
    """HBM traffic of the tiled algorithm: Q, K, V, O once each (ch06/flash_attention.py:77-104)."""
    qkv = 3 * batch_size * num_heads * seq_len * head_dim * dtype_bytes
    out = batch_size * num_heads * seq_len * head_dim * dtype_bytes
    tile = block_size * head_dim * dtype_bytes
    sram = 4 * tile + block_size * block_size * dtype_bytes + block_size * 2 * dtype_bytes
    naive = qkv + out + batch_size * num_heads * seq_len * seq_len * dtype_bytes
    return {
        "hbm_bytes": qkv + out,
        "hbm_mb": (qkv + out) / 1024 / 1024,
        "sram_bytes_per_block": sram,
        "sram_kb_per_block": sram / 1024,
        "naive_hbm_bytes": naive,
        "memory_savings": f"{seq_len // block_size}x",
    }


def prefill_algorithmic_flops(B: int, Hq: int, Nq: int, Nk: int, D: int, causal: bool) -> float:
    """Roofline numerator: 4*B*Hq*Nq*Nk*D, x 1/2 for the causal square
    (FlashAttention-paper convention); a rectangular causal problem counts its full Nq x (Nk-Nq)
    block plus half of the Nq x Nq triangle."""
    pairs = Nq * Nk if not causal else Nq * (Nk - Nq) + Nq * Nq / 2
    return 4.0 * B * Hq * D * pairs


def flash_attention_paged(q: torch.Tensor, k_pool: torch.Tensor, v_pool: torch.Tensor, block_tables: torch.Tensor,
                          seq_lens: torch.Tensor, *, layer: int = 0, scale: float | None = None,
                          max_seq_len: int | None = None, return_lse: bool = False, validate: bool = False):
    """Chunked prefill over the ch07 paged pools, read in place (no gather).

    q (B, Hq, Nq, D): the Nq newest tokens of each sequence (their K/V already appended to the pools, e.g. with
    `PagedKVCache.append`); pools (P, layers, bs, Hkv, D); block_tables (B, max_pages) int32; seq_lens (B,) int32
    CUDA = cached length per sequence including the chunk.  Query i of sequence b sees key j iff
    j <= i + (seq_lens[b] - Nq): the offset mask of ch02/cached_generation.py:85-91, per sequence.
    Returns o like q (and lse (B, Hq, Nq) float32 when return_lse).
    """
    if not (q.is_cuda and k_pool.is_cuda and v_pool.is_cuda):
        raise RuntimeError("flash_attention_paged runs on CUDA tensors only (no CPU fallback)")
    if q.dim() != 4 or k_pool.dim() != 5:
        raise RuntimeError(f"q must be (B, Hq, Nq, D) and the pools (P, layers, bs, Hkv, D); got {tuple(q.shape)}, {tuple(k_pool.shape)}")
    if not (q.dtype == k_pool.dtype == v_pool.dtype):
        raise RuntimeError("q and the pools must share a dtype")
    if k_pool.shape != v_pool.shape or k_pool.stride() != v_pool.stride() or k_pool.stride(-1) != 1:
        raise RuntimeError("k and v pools must have the same shape/strides and a unit head_dim stride")
    B, Hq, Nq, D = q.shape
    P, n_layers, bs, Hkv, Dk = k_pool.shape
    if Dk != D or Hq % Hkv != 0:
        raise RuntimeError(f"pools {tuple(k_pool.shape)} do not match q {tuple(q.shape)}")
    if not 0 <= layer < n_layers:
        raise IndexError(f"layer {layer} out of range for {n_layers} layers")
    if block_tables.dtype != torch.int32 or not block_tables.is_cuda or block_tables.dim() != 2 or block_tables.shape[0] != B:
        raise RuntimeError("block_tables must be a (B, max_pages) CUDA int32 tensor")
    if seq_lens.dtype != torch.int32 or not seq_lens.is_cuda or seq_lens.shape != (B,):
        raise RuntimeError("seq_lens must be a (B,) CUDA int32 tensor")
    if block_tables.stride(-1) != 1:
        block_tables = block_tables.contiguous()
    seq_lens = seq_lens.contiguous()
    q = _unit_inner(q)
    cap = block_tables.shape[1] * bs
    max_seq_len = cap if max_seq_len is None else min(int(max_seq_len), cap)
    if Nq > max_seq_len:
        raise ValueError(f"Nq ({Nq}) exceeds the cached length bound ({max_seq_len})")
    if validate:
        validate_paged_args(seq_lens, block_tables, bs, P, q_lens=Nq, max_seq_len=max_seq_len)
    if scale is None:
        scale = D ** -0.5
    out = torch.empty_like(q)
    if out.stride(-1) != 1:
        out = torch.empty(q.shape, dtype=q.dtype, device=q.device)
    lse = torch.empty((B, Hq, Nq), dtype=torch.float32, device=q.device) if return_lse else None
    lib = _lib.load()
    with _lib.on_device(q.device):
        rc = lib.pli_prefill_paged_fwd(
            q.data_ptr(), k_pool.data_ptr(), v_pool.data_ptr(), block_tables.data_ptr(), seq_lens.data_ptr(),
            out.data_ptr(), lse.data_ptr() if lse is not None else None, B, Hq, Hkv, Nq, D, max_seq_len, bs,
            block_tables.stride(0), layer, P, _lib.i64(*q.stride()[:3]), _lib.i64(*k_pool.stride()[:4]),
            _lib.i64(*out.stride()[:3]), float(scale), _lib.dtype_code(q.dtype), _lib.current_stream_ptr(q.device))
    _lib.check(rc)
    return (out, lse) if return_lse else out


def flash_attention_varlen_paged(q: torch.Tensor, k_pool: torch.Tensor, v_pool: torch.Tensor, block_tables: torch.Tensor,
                                 seq_lens: torch.Tensor, cu_seqlens_q: torch.Tensor, max_q_len: int, *, layer: int = 0,
                                 scale: float | None = None, max_seq_len: int | None = None, return_lse: bool = False,
                                 validate: bool = False):
    """`flash_attention_paged` for ragged query lengths: q (total_q, Hq, D) packs the newest tokens of every sequence,
    rows [cu_seqlens_q[b], cu_seqlens_q[b+1]) belong to sequence b (cu_seqlens_q (B+1,) int32 CUDA, max_q_len a host
    bound of the per-sequence lengths).  One launch serves the prefill side of a mixed batch
    (ch08/mixed_batch.py:63-104).  Returns o (total_q, Hq, D) (and lse (Hq, total_q) float32 when return_lse)."""
    if not (q.is_cuda and k_pool.is_cuda and v_pool.is_cuda):
        raise RuntimeError("flash_attention_varlen_paged runs on CUDA tensors only (no CPU fallback)")
    if q.dim() != 3 or k_pool.dim() != 5:
        raise RuntimeError(f"q must be (total_q, Hq, D) and the pools (P, layers, bs, Hkv, D); got {tuple(q.shape)}, {tuple(k_pool.shape)}")
    if not (q.dtype == k_pool.dtype == v_pool.dtype):
        raise RuntimeError("q and the pools must share a dtype")
    if k_pool.shape != v_pool.shape or k_pool.stride() != v_pool.stride() or k_pool.stride(-1) != 1:
        raise RuntimeError("k and v pools must have the same shape/strides and a unit head_dim stride")
    T, Hq, D = q.shape
    P, n_layers, bs, Hkv, Dk = k_pool.shape
    B = seq_lens.shape[0]
    if Dk != D or Hq % Hkv != 0:
        raise RuntimeError(f"pools {tuple(k_pool.shape)} do not match q {tuple(q.shape)}")
    if not 0 <= layer < n_layers:
        raise IndexError(f"layer {layer} out of range for {n_layers} layers")
    if block_tables.dtype != torch.int32 or not block_tables.is_cuda or block_tables.dim() != 2 or block_tables.shape[0] != B:
        raise RuntimeError("block_tables must be a (B, max_pages) CUDA int32 tensor")
    for name, x, n in (("seq_lens", seq_lens, B), ("cu_seqlens_q", cu_seqlens_q, B + 1)):
        if x.dtype != torch.int32 or not x.is_cuda or x.shape != (n,):
            raise RuntimeError(f"{name} must be a ({n},) CUDA int32 tensor")
    if T == 0:
        raise RuntimeError("empty q")
    if not 0 < int(max_q_len) <= T:
        raise ValueError(f"max_q_len ({max_q_len}) outside (0, total_q={T}]")
    if block_tables.stride(-1) != 1:
        block_tables = block_tables.contiguous()
    seq_lens, cu_seqlens_q = seq_lens.contiguous(), cu_seqlens_q.contiguous()
    q = _unit_inner(q)
    cap = block_tables.shape[1] * bs
    max_seq_len = cap if max_seq_len is None else min(int(max_seq_len), cap)
    if validate:
        cu = cu_seqlens_q.detach().to("cpu", torch.int64)
        if int(cu[0]) != 0 or int(cu[-1]) != T or bool((cu[1:] < cu[:-1]).any()):
            raise ValueError("cu_seqlens_q must be non-decreasing with [0] = 0 and [B] = total_q")
        if int((cu[1:] - cu[:-1]).max()) > int(max_q_len):
            raise ValueError(f"a sequence brings more than max_q_len = {max_q_len} query tokens")
        validate_paged_args(seq_lens, block_tables, bs, P, q_lens=cu[1:] - cu[:-1], max_seq_len=max_seq_len)
    if scale is None:
        scale = D ** -0.5
    out = torch.empty((T, Hq, D), dtype=q.dtype, device=q.device)
    lse = torch.empty((Hq, T), dtype=torch.float32, device=q.device) if return_lse else None
    lib = _lib.load()
    with _lib.on_device(q.device):
        rc = lib.pli_prefill_varlen_paged_fwd(
            q.data_ptr(), k_pool.data_ptr(), v_pool.data_ptr(), block_tables.data_ptr(), seq_lens.data_ptr(),
            cu_seqlens_q.data_ptr(), out.data_ptr(), lse.data_ptr() if lse is not None else None, B, Hq, Hkv, T,
            int(max_q_len), D, max_seq_len, bs, block_tables.stride(0), layer, P, _lib.i64(*q.stride()[:2]),
            _lib.i64(*k_pool.stride()[:4]), _lib.i64(*out.stride()[:2]), float(scale), _lib.dtype_code(q.dtype),
            _lib.current_stream_ptr(q.device))
    _lib.check(rc)
    return (out, lse) if return_lse else out
