"""CPU fp32 oracle for the attention hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain PyTorch on the CPU, the algorithm the reference
(Infatoshi/physics-llm-inference) implements for the path named in
BASELINE.json.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the
product package (``physics_llm_inference_b200``) never does.

Parity pinning: the restatement is checked against
  * the unmodified reference functions, imported from /root/reference in the
    build container, by ``oracle/make_golden.py`` (bit-equality for the
    non-causal ch06 recurrence, <=1e-6 for the composed cases), and
  * the committed outputs of those reference runs in ``tests/golden/*.npz``
    (``tests/test_oracle_golden.py``), which travel to the GPU box where
    /root/reference does not exist.

Reference lines each function follows (paths relative to /root/reference):
  expand_kv                   ch01/gqa.py:14,30-31   ch02/cached_generation.py:77-78
  flash_attention_oracle      ch06/flash_attention.py:14-74 (recurrence :38-72)
      + causal rule           ch01/gqa.py:33-34, ch02/cached_generation.py:85-91
  naive_attention_oracle      ch06/attention_memory.py:19-33 (+ same masks)
  cached_attention_oracle     ch02/cached_generation.py:72-94 (= ch02/kv_cache.py:81-98)
  gather_paged                ch07/paged_memory.py:38-48 (pool layout), :54,:84-86 (ceil-div
                              page count => token t lives in page table[t // bs], slot t % bs)
  paged_decode_oracle         gather_paged + cached_attention_oracle
  combine_splits_oracle       log-sum-exp merge of independent softmax partials
                              (ch06/online_softmax.py:28-53 applied to whole partials)
  online_softmax_oracle       ch06/online_softmax.py:13-25
  online_softmax_with_output_oracle   ch06/online_softmax.py:28-53
"""
from __future__ import annotations

import math

import torch

NEG_INF = float("-inf")


def expand_kv(x: torch.Tensor, num_groups: int) -> torch.Tensor:
    """(B,Hkv,N,D) -> (B,Hkv*G,N,D); q-head h reads kv-head h // G (ch01/gqa.py:30-31)."""
    return x if num_groups == 1 else x.repeat_interleave(num_groups, dim=1)


def causal_visible(nq: int, nk: int) -> torch.Tensor:
    """bool (nq,nk): True where key j is visible to query i.

    ch02/cached_generation.py:87-90 builds ``triu(ones(s,L), diagonal=L-s+1)`` as the
    *masked* set, i.e. visible iff j <= i + (L - s) (bottom-right aligned); for s == L
    this is ch01/gqa.py:33's ``triu(diagonal=1)``.
    """
    masked = torch.triu(torch.ones(nq, nk, dtype=torch.bool), diagonal=nk - nq + 1)
    return ~masked


def flash_attention_oracle(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    scale: float | None = None,
    block_q: int = 64,
    block_k: int = 64,
    causal: bool = False,
    skip_masked_blocks: bool = False,
):
    """ch06 tiled recurrence in fp32 with the ch01/ch02 mask and GQA map added.

    Follows ch06/flash_attention.py:38-72 statement by statement; the three edits are
    (1) K/V are expanded per ch01/gqa.py:30-31 and Nk is taken from k, (2) masked scores are
    set to -inf between :55 and :57, (3) (m, d) are kept and lse = m + log d is returned.
    ``skip_masked_blocks`` drops K blocks that are entirely masked for the Q block (their
    contribution is exactly zero); it only exists so the CPU baseline does not time dead work.
    Returns (O fp32 (B,Hq,Nq,D), lse fp32 (B,Hq,Nq)).
    """
    q = q.detach().to("cpu", torch.float32)
    k = k.detach().to("cpu", torch.float32)
    v = v.detach().to("cpu", torch.float32)
    B, Hq, Nq, D = q.shape
    Hkv, Nk = k.shape[1], k.shape[2]
    assert Hq % Hkv == 0, "num_heads must be a multiple of num_kv_heads (ch01/gqa.py:11)"
    k = expand_kv(k, Hq // Hkv)
    v = expand_kv(v, Hq // Hkv)
    if scale is None:
        scale = D ** -0.5
    off = Nk - Nq

    Dv = v.shape[-1]
    output = torch.zeros(B, Hq, Nq, Dv)
    row_max = torch.full((B, Hq, Nq), NEG_INF)
    row_sum = torch.zeros((B, Hq, Nq))

    for q_start in range(0, Nq, block_q):
        q_end = min(q_start + block_q, Nq)
        q_block = q[:, :, q_start:q_end, :]
        block_output = torch.zeros(B, Hq, q_end - q_start, Dv)
        block_max = torch.full((B, Hq, q_end - q_start), NEG_INF)
        block_sum = torch.zeros((B, Hq, q_end - q_start))
        qi = torch.arange(q_start, q_end).unsqueeze(-1)

        for k_start in range(0, Nk, block_k):
            k_end = min(k_start + block_k, Nk)
            if causal and skip_masked_blocks and k_start > (q_end - 1) + off:
                break
            k_block = k[:, :, k_start:k_end, :]
            v_block = v[:, :, k_start:k_end, :]

            scores = torch.matmul(q_block, k_block.transpose(-2, -1)) * scale
            if causal and k_end - 1 > q_start + off:   # blocks left of the diagonal have nothing to mask
                kj = torch.arange(k_start, k_end).unsqueeze(0)
                scores = scores.masked_fill(kj > qi + off, NEG_INF)

            new_max = torch.maximum(block_max, scores.max(dim=-1).values)
            scale_old = torch.exp(block_max - new_max)
            scale_new = torch.exp(scores - new_max.unsqueeze(-1))
            new_sum = block_sum * scale_old + scale_new.sum(dim=-1)
            block_output = (block_output * block_sum.unsqueeze(-1) * scale_old.unsqueeze(-1)
                            + torch.matmul(scale_new, v_block)) / new_sum.unsqueeze(-1)
            block_max = new_max
            block_sum = new_sum

        output[:, :, q_start:q_end, :] = block_output
        row_max[:, :, q_start:q_end] = block_max
        row_sum[:, :, q_start:q_end] = block_sum

    return output, row_max + torch.log(row_sum)


def naive_attention_oracle(q, k, v, scale=None, causal=False):
    """Materialised softmax(QK^T*scale)V, ch06/attention_memory.py:19-33, + mask + GQA."""
    q = q.detach().to("cpu", torch.float32)
    k = k.detach().to("cpu", torch.float32)
    v = v.detach().to("cpu", torch.float32)
    Hq, Hkv = q.shape[1], k.shape[1]
    k = expand_kv(k, Hq // Hkv)
    v = expand_kv(v, Hq // Hkv)
    if scale is None:
        scale = q.shape[-1] ** -0.5
    scores = torch.matmul(q, k.transpose(-2, -1)) * scale
    if causal:
        scores = scores.masked_fill(~causal_visible(q.shape[2], k.shape[2]), NEG_INF)
    lse = torch.logsumexp(scores, dim=-1)
    return torch.matmul(torch.softmax(scores, dim=-1), v), lse


def cached_attention_oracle(q, k_cache, v_cache, seq_len=None, scale=None):
    """Attention of q (B,Hq,s,D) over a contiguous cache (B,L>=seq_len,Hkv,D).

    ch02/cached_generation.py:72-94: slice [:, :seq_len] (:33), transpose to (B,Hkv,L,D)
    (:72-74), repeat_interleave (:77-78), QK^T / sqrt(D) (:82), offset causal mask only when
    s > 1 (:85-91), softmax, PV (:93-94).  ``seq_len`` may be an int or a (B,) tensor (ragged
    batches are evaluated row by row; the reference itself keeps one seq_len per cache).
    Returns (O fp32 (B,Hq,s,D), lse fp32 (B,Hq,s)).
    """
    q = q.detach().to("cpu", torch.float32)
    B, Hq, s, D = q.shape
    L_max = k_cache.shape[1]
    if seq_len is None:
        seq_len = L_max
    lens = [int(seq_len)] * B if not torch.is_tensor(seq_len) else [int(x) for x in seq_len.tolist()]
    outs, lses = [], []
    for b in range(B):
        L = lens[b]
        kf = k_cache[b:b + 1, :L].detach().to("cpu", torch.float32).transpose(1, 2)
        vf = v_cache[b:b + 1, :L].detach().to("cpu", torch.float32).transpose(1, 2)
        G = Hq // kf.shape[1]
        kf, vf = expand_kv(kf, G), expand_kv(vf, G)
        if scale is None:
            scores = torch.matmul(q[b:b + 1], kf.transpose(-2, -1)) / math.sqrt(D)
        else:
            scores = torch.matmul(q[b:b + 1], kf.transpose(-2, -1)) * scale
        if s > 1:
            scores = scores.masked_fill(~causal_visible(s, L), NEG_INF)
        lses.append(torch.logsumexp(scores, dim=-1))
        outs.append(torch.matmul(torch.softmax(scores, dim=-1), vf))
    return torch.cat(outs, 0), torch.cat(lses, 0)


def page_address(t: int, block_indices, block_size: int):
    """Logical token t -> (physical page, slot).  ch07/paged_memory.py:54,84-86."""
    return int(block_indices[t // block_size]), t % block_size


def gather_paged(pool: torch.Tensor, block_indices, num_tokens: int, layer: int = 0) -> torch.Tensor:
    """Gather one request's tokens from a pool (P, layers, bs, Hkv, D) -> (num_tokens, Hkv, D)."""
    bs = pool.shape[2]
    n_pages = (num_tokens + bs - 1) // bs
    idx = torch.as_tensor(list(block_indices[:n_pages]), dtype=torch.long)
    rows = pool[idx, layer]                      # (n_pages, bs, Hkv, D)
    return rows.reshape(n_pages * bs, *pool.shape[3:])[:num_tokens]


def paged_decode_oracle(q, k_pool, v_pool, block_tables, seq_lens, layer=0, scale=None):
    """Decode attention of q (B,Hq,s,D) over paged K/V.

    block_tables: (B, max_pages) int tensor or list of lists; seq_lens: (B,) ints.
    Gathers each request to contiguous (ch07 address rule) and applies ch02's maths.
    """
    q = q.detach().to("cpu", torch.float32)
    B = q.shape[0]
    if torch.is_tensor(block_tables):
        block_tables = block_tables.tolist()
    if torch.is_tensor(seq_lens):
        seq_lens = seq_lens.tolist()
    outs, lses = [], []
    for b in range(B):
        L = int(seq_lens[b])
        kg = gather_paged(k_pool.detach().cpu(), block_tables[b], L, layer).unsqueeze(0)
        vg = gather_paged(v_pool.detach().cpu(), block_tables[b], L, layer).unsqueeze(0)
        o, lse = cached_attention_oracle(q[b:b + 1], kg, vg, L, scale)
        outs.append(o)
        lses.append(lse)
    return torch.cat(outs, 0), torch.cat(lses, 0)


def combine_splits_oracle(o_parts: torch.Tensor, lse_parts: torch.Tensor):
    """Merge S independent softmax partials: o_parts (S,...,D) normalised, lse_parts (S,...)."""
    m = lse_parts.max(dim=0).values
    w = torch.exp(lse_parts - m)
    w = torch.where(torch.isnan(w), torch.zeros_like(w), w)
    den = w.sum(dim=0)
    o = (o_parts * w.unsqueeze(-1)).sum(dim=0) / den.unsqueeze(-1)
    return o, m + torch.log(den)


def online_softmax_oracle(x: torch.Tensor) -> torch.Tensor:
    """ch06/online_softmax.py:13-25, statement by statement: one pass with the running (m, d), then normalise."""
    x = x.detach().to("cpu", torch.float32)
    n = x.shape[-1]
    m = x[..., 0].clone()
    d = torch.ones_like(m)
    for i in range(1, n):
        m_new = torch.maximum(m, x[..., i])
        d = d * torch.exp(m - m_new) + torch.exp(x[..., i] - m_new)
        m = m_new
    return torch.exp(x - m.unsqueeze(-1)) / d.unsqueeze(-1)


def online_softmax_with_output_oracle(x: torch.Tensor, v: torch.Tensor):
    """ch06/online_softmax.py:28-53: the recurrence with the running weighted sum; returns (o, d)."""
    x = x.detach().to("cpu", torch.float32)
    v = v.detach().to("cpu", torch.float32)
    n = x.shape[-1]
    m = x[..., 0].clone()
    d = torch.ones_like(m)
    o = v[..., 0, :].clone()
    for i in range(1, n):
        m_new = torch.maximum(m, x[..., i])
        scale_old = torch.exp(m - m_new)
        scale_new = torch.exp(x[..., i] - m_new)
        d_new = d * scale_old + scale_new
        o = (o * d.unsqueeze(-1) * scale_old.unsqueeze(-1) + v[..., i, :] * scale_new.unsqueeze(-1)) / d_new.unsqueeze(-1)
        m = m_new
        d = d_new
    return o, d


def seeded_qkv(seed: int, B: int, Hq: int, Hkv: int, Nq: int, Nk: int, D: int, dtype=torch.float32):
    """Seeded N(0,1) inputs, generated in fp32 on the CPU then cast."""
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(B, Hq, Nq, D, generator=g).to(dtype)
    k = torch.randn(B, Hkv, Nk, D, generator=g).to(dtype)
    v = torch.randn(B, Hkv, Nk, D, generator=g).to(dtype)
    return q, k, v


def seeded_paged(seed: int, B: int, Hq: int, Hkv: int, D: int, block_size: int, seq_lens,
                 num_layers: int = 1, slack_pages: int = 3, dtype=torch.float32):
    """Seeded paged-decode case: q, pools, a random-permutation block table, seq_lens."""
    g = torch.Generator().manual_seed(seed)
    seq_lens = [int(x) for x in seq_lens]
    pages_per = [(L + block_size - 1) // block_size for L in seq_lens]
    max_pages = max(pages_per)
    P = sum(pages_per) + slack_pages
    q = torch.randn(B, Hq, 1, D, generator=g).to(dtype)
    k_pool = torch.randn(P, num_layers, block_size, Hkv, D, generator=g).to(dtype)
    v_pool = torch.randn(P, num_layers, block_size, Hkv, D, generator=g).to(dtype)
    perm = torch.randperm(P, generator=g)
    table = torch.full((B, max_pages), -1, dtype=torch.int32)
    at = 0
    for b, n in enumerate(pages_per):
        table[b, :n] = perm[at:at + n].to(torch.int32)
        at += n
    return q, k_pool, v_pool, table, torch.tensor(seq_lens, dtype=torch.int32)
