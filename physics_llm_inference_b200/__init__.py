"""B200-native attention hot path of Infatoshi/physics-llm-inference.

Drop-in surface (same names and call signatures as the reference packages it replaces):

  ch06  flash_attention_forward, FlashAttentionConfig, attention_flops, flash_attention_memory_bytes,
        online_softmax, online_softmax_with_output, standard_softmax, attention_memory_bytes,
        attention_arithmetic_intensity          (naive_attention stays on the checker side: oracle/)
  ch02  KVCache, LayerKVCache, create_caches            (+ flash_decode / decode_with_cache)
  ch07  BlockTable, PagedKVCache                        (+ decode_with_paged, prefill_with_paged, kv_append)

Everything numeric runs in `libpli_attention.so` (hand-written sm_100a CUDA behind the C ABI in
include/pli_attention.h).  Importing the package does not need a GPU; calling an op does, and
fails loudly if the library has not been built.
"""
from ._lib import LIB_PATH, PliError, launch_count, reset_launch_count
from .decode import (DecodePlan, decode_kernel_kind, decode_num_splits, decode_with_cache, decode_with_paged, decode_workspace,
                     flash_decode, mixed_batch_attention, paged_gather, prefill_with_paged)
from .flash_attention import (FlashAttentionConfig, attention_flops, flash_attention, flash_attention_forward,
                              flash_attention_memory_bytes, flash_attention_paged, flash_attention_varlen_paged,
                              prefill_algorithmic_flops,
                              prefill_kernel_kind)
from .kv_cache import KVCache, LayerKVCache, create_caches, kv_append
from .online_softmax import (AttentionMemoryStats, attention_arithmetic_intensity, attention_memory_bytes, online_softmax,
                             online_softmax_with_output, standard_softmax)
from .modules import CachedGQA, DecodeGraphRunner, GroupedQueryAttention, TensorParallelGQA
from .paged_memory import BlockTable, PagedKVCache
from .sharding import HeadShard, PeerOutput, gather_heads, init_distributed, make_shard, shard_kv_heads

__all__ = [
    "flash_attention_forward", "flash_attention", "flash_attention_paged", "flash_attention_varlen_paged",
    "FlashAttentionConfig", "attention_flops", "attention_memory_bytes", "attention_arithmetic_intensity",
    "AttentionMemoryStats", "online_softmax", "online_softmax_with_output", "standard_softmax",
    "flash_attention_memory_bytes", "prefill_algorithmic_flops", "prefill_kernel_kind",
    "flash_decode", "DecodePlan", "decode_with_cache", "decode_with_paged", "decode_num_splits", "decode_workspace",
    "decode_kernel_kind", "paged_gather", "prefill_with_paged", "mixed_batch_attention",
    "KVCache", "LayerKVCache", "create_caches", "kv_append", "BlockTable", "PagedKVCache",
    "GroupedQueryAttention", "CachedGQA", "DecodeGraphRunner", "TensorParallelGQA",
    "HeadShard", "PeerOutput", "make_shard", "shard_kv_heads", "gather_heads", "init_distributed",
    "PliError", "LIB_PATH", "launch_count", "reset_launch_count",
]
