"""Paged KV storage: the reference's allocator with a device read/write path behind it.

`BlockTable` and `PagedKVCache` keep the names, constructor arguments, methods and error behaviour
of ch07/paged_memory.py:7-13 and :16-137 (ceil-div page counts :54/:84-86, RuntimeError on
exhaustion :56-60/:88-92, KeyError for an unknown request :77-78, free returns the page count
:100-110).  The reference only ever allocates page *indices* (nothing reads or
writes k_cache/v_cache, and they are None without CUDA); here the pools exist on the GPU in the
reference's layout (num_blocks, num_layers, block_size, num_heads, head_dim) and are

  written by `append` (pli_kv_append)          token t -> page block_indices[t // bs], slot t % bs
  read    by `flash_decode(..., block_tables=)`  through `block_table_tensor`.

The allocator itself is host bookkeeping, exactly as in the reference (a Python set of free pages).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch

from .kv_cache import kv_append


@dataclass
class BlockTable:
    request_id: int
    block_indices: list[int] = field(default_factory=list)
    num_tokens: int = 0

    def num_blocks(self) -> int:
        return len(self.block_indices)


class PagedKVCache:
    def __init__(self, num_blocks: int, block_size: int, num_layers: int, num_heads: int, head_dim: int,
                 dtype: torch.dtype = torch.float16, device: str = "cuda"):
        self.num_blocks = num_blocks
        self.block_size = block_size
        self.num_layers = num_layers
        self.num_heads = num_heads
        self.head_dim = head_dim
        self.dtype = dtype
        self.device = device

        self.free_blocks: set[int] = set(range(num_blocks))
        self.block_tables: dict[int, BlockTable] = {}
        # pages referenced by more than one block table (prefix sharing); absent = 1 owner
        self.shared_refs: dict[int, int] = {}

        # ch07/paged_memory.py:38-51: tensors only when CUDA is there and asked for
        if torch.cuda.is_available() and str(device).startswith("cuda"):
            shape = (num_blocks, num_layers, block_size, num_heads, head_dim)
            self.k_cache = torch.zeros(shape, dtype=dtype, device=device)
            self.v_cache = torch.zeros(shape, dtype=dtype, device=device)
        else:
            self.k_cache = None
            self.v_cache = None

    # ---- allocator: ch07/paged_memory.py:53-110 ----
    def allocate_blocks(self, request_id: int, num_tokens: int) -> BlockTable:
        num_blocks_needed = (num_tokens + self.block_size - 1) // self.block_size
        if len(self.free_blocks) < num_blocks_needed:
            raise RuntimeError(
                f"Not enough free blocks: need {num_blocks_needed}, have {len(self.free_blocks)}")
        allocated = [self.free_blocks.pop() for _ in range(num_blocks_needed)]
        table = BlockTable(request_id=request_id, block_indices=allocated, num_tokens=num_tokens)
        self.block_tables[request_id] = table
        return table

    def extend_blocks(self, request_id: int, new_tokens: int) -> None:
        if request_id not in self.block_tables:
            raise KeyError(f"Request {request_id} not found")
        table = self.block_tables[request_id]
        new_total = table.num_tokens + new_tokens
        old_blocks = (table.num_tokens + self.block_size - 1) // self.block_size
        new_blocks = (new_total + self.block_size - 1) // self.block_size
        blocks_needed = new_blocks - old_blocks
        if blocks_needed > len(self.free_blocks):
            raise RuntimeError(
                f"Not enough free blocks for extension: need {blocks_needed}, have {len(self.free_blocks)}")
        for _ in range(blocks_needed):
            table.block_indices.append(self.free_blocks.pop())
        table.num_tokens = new_total

    def free_blocks_for_request(self, request_id: int) -> int:
        if request_id not in self.block_tables:
            return 0
        table = self.block_tables.pop(request_id)
        for block_idx in table.block_indices:
            refs = self.shared_refs.get(block_idx, 1) - 1
            if refs >= 1:                     # still referenced by another request's table
                if refs == 1:
                    del self.shared_refs[block_idx]
                else:
                    self.shared_refs[block_idx] = refs
            else:
                self.free_blocks.add(block_idx)
        return len(table.block_indices)

    def fork_request(self, parent_id: int, child_id: int, num_tokens: int | None = None) -> BlockTable:
        """Start `child_id` with the first `num_tokens` cached tokens of `parent_id` (a matched prefix, as
        `RadixCache.match_prefix` of ch07/radix_cache.py:72-103 reports it) WITHOUT recomputing or copying
        them: full pages are aliased (both block tables point at the same physical page, ref-counted),
        a partially filled last page is copied so the child can append to it.  The read path needs no
        change: the kernels only ever see a block table."""
        if parent_id not in self.block_tables:
            raise KeyError(f"Request {parent_id} not found")
        parent = self.block_tables[parent_id]
        n = parent.num_tokens if num_tokens is None else num_tokens
        if not 0 <= n <= parent.num_tokens:
            raise ValueError(f"cannot share {n} tokens of a request with {parent.num_tokens}")
        full, rem = divmod(n, self.block_size)
        if rem and not self.free_blocks:
            raise RuntimeError("Not enough free blocks: need 1, have 0")
        pages = list(parent.block_indices[:full])
        for pg in pages:
            self.shared_refs[pg] = self.shared_refs.get(pg, 1) + 1
        if rem:
            new_page = self.free_blocks.pop()
            if self.k_cache is not None:
                src = parent.block_indices[full]
                self.k_cache[new_page].copy_(self.k_cache[src])
                self.v_cache[new_page].copy_(self.v_cache[src])
            pages.append(new_page)
        table = BlockTable(request_id=child_id, block_indices=pages, num_tokens=n)
        self.block_tables[child_id] = table
        return table

    def kv_indices(self, request_id: int) -> list[int]:
        """Flat slot address `page * block_size + slot` of every cached token of a request, in token order: the
        per-token `kv_indices` a `RadixCache.insert(token_ids, kv_indices)` (ch07/radix_cache.py:20-70) records."""
        if request_id not in self.block_tables:
            raise KeyError(f"Request {request_id} not found")
        table = self.block_tables[request_id]
        bs = self.block_size
        return [table.block_indices[t // bs] * bs + t % bs for t in range(table.num_tokens)]

    def share_prefix(self, child_id: int, matched: int, kv_indices: list[int]) -> BlockTable:
        """Start request `child_id` from a radix-cache hit: `(matched, kv_indices)` exactly as
        `RadixCache.match_prefix(token_ids)` returns them (ch07/radix_cache.py:72-103: the number of matched tokens and
        one KV slot address per matched token, here `page * block_size + slot` as produced by `kv_indices()`).

        The per-token addresses are turned into page-granular sharing: every run of `block_size` tokens that occupies
        one whole physical page in order is ALIASED (the child's block table points at the same page, ref-counted —
        nothing is copied or recomputed); the remainder (a partly filled last page, or addresses that are not
        page-aligned) is gathered into a fresh page so the child can append behind it.  Returns the child's table with
        `num_tokens = min(matched, len(kv_indices))`; the caller computes K/V only for the tokens after it."""
        if child_id in self.block_tables:
            raise ValueError(f"Request {child_id} already has a block table")
        bs = self.block_size
        n = min(int(matched), len(kv_indices))
        if n < 0:
            raise ValueError("matched must be non-negative")
        chunks = []                                                # (alias page | None, [slot addresses])
        for a in range(0, n, bs):
            slots = [int(x) for x in kv_indices[a:min(a + bs, n)]]
            page = slots[0] // bs
            whole = len(slots) == bs and all(sl == page * bs + i for i, sl in enumerate(slots))
            for sl in slots:
                pg = sl // bs
                if not 0 <= pg < self.num_blocks or pg in self.free_blocks:
                    raise RuntimeError(f"kv index {sl} points at page {pg}, which is not allocated (stale prefix entry)")
            chunks.append((page if whole else None, slots))
        fresh = sum(1 for pg, _ in chunks if pg is None)
        if fresh > len(self.free_blocks):
            raise RuntimeError(f"Not enough free blocks: need {fresh}, have {len(self.free_blocks)}")
        pages = []
        for pg, slots in chunks:
            if pg is not None:
                self.shared_refs[pg] = self.shared_refs.get(pg, 1) + 1
                pages.append(pg)
                continue
            new_page = self.free_blocks.pop()
            if self.k_cache is not None:
                dev = self.k_cache.device
                src_pages = torch.tensor([sl // bs for sl in slots], dtype=torch.long, device=dev)
                src_slots = torch.tensor([sl % bs for sl in slots], dtype=torch.long, device=dev)
                # (n_tok, layers, heads, dim) -> the first n_tok slots of the new page, every layer
                self.k_cache[new_page, :, :len(slots)] = self.k_cache[src_pages, :, src_slots].transpose(0, 1)
                self.v_cache[new_page, :, :len(slots)] = self.v_cache[src_pages, :, src_slots].transpose(0, 1)
            pages.append(new_page)
        table = BlockTable(request_id=child_id, block_indices=pages, num_tokens=n)
        self.block_tables[child_id] = table
        return table

    def get_num_free_blocks(self) -> int:
        return len(self.free_blocks)

    def get_memory_usage(self) -> dict:
        used_blocks = self.num_blocks - len(self.free_blocks)
        bytes_per_block = 2 * self.num_layers * self.block_size * self.num_heads * self.head_dim * 2
        total_bytes = self.num_blocks * bytes_per_block
        used_bytes = used_blocks * bytes_per_block
        return {
            "total_blocks": self.num_blocks,
            "used_blocks": used_blocks,
            "free_blocks": len(self.free_blocks),
            "block_size_tokens": self.block_size,
            "bytes_per_block": bytes_per_block,
            "total_mb": total_bytes / 1024 / 1024,
            "used_mb": used_bytes / 1024 / 1024,
            "utilization": used_blocks / self.num_blocks if self.num_blocks > 0 else 0,
        }

    # ---- device path (new: the reference has no reader/writer for the pools) ----
    def block_table_tensor(self, request_ids, device=None):
        """(block_tables (B, max_pages) int32, seq_lens (B,) int32) for a batch of requests.
        Unused table entries are -1 and are never dereferenced by the kernels."""
        device = device if device is not None else self.device
        tables = [self.block_tables[r] for r in request_ids]
        width = max(1, max(t.num_blocks() for t in tables))
        rows = [t.block_indices + [-1] * (width - t.num_blocks()) for t in tables]
        bt = torch.tensor(rows, dtype=torch.int32).to(device)
        lens = torch.tensor([t.num_tokens for t in tables], dtype=torch.int32).to(device)
        return bt, lens

    def append(self, request_ids, k_new: torch.Tensor, v_new: torch.Tensor, layer: int = 0,
               extend: bool = True) -> None:
        """Append k_new/v_new (B, n, num_heads, head_dim) to each request's pages at `layer`.

        With extend=True the requests are grown by n tokens first (`extend_blocks`, or
        `allocate_blocks` for a new id); with extend=False the last n tokens of the already-sized
        tables are written (use this for layers > 0 of the same step)."""
        if self.k_cache is None:
            raise RuntimeError("PagedKVCache has no device pools (constructed without CUDA)")
        n = k_new.shape[1]
        starts = []
        for r in request_ids:
            if extend:
                if r in self.block_tables:
                    self.extend_blocks(r, n)
                else:
                    self.allocate_blocks(r, n)
            starts.append(self.block_tables[r].num_tokens - n)
        bt, _ = self.block_table_tensor(request_ids)
        start = torch.tensor(starts, dtype=torch.int32).to(self.k_cache.device)
        kv_append(self.k_cache, self.v_cache, k_new, v_new, start, block_tables=bt, layer=layer)
