"""Generate tests/golden/*.npz by running the UNMODIFIED reference on the CPU.

TEST INFRASTRUCTURE ONLY.  Run in the build container, where /root/reference exists:

    python oracle/make_golden.py            # writes tests/golden/, asserts oracle == reference

The GPU box has no /root/reference, so the outputs are committed.  Every case stores the
seed/shape (inputs are regenerated from the seed), an input checksum that detects RNG drift,
and the reference's outputs.  While generating, the oracle restatement is checked against the
reference directly (bit-equal for the ch06 non-causal recurrence).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("PLI_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.dont_write_bytecode = True

from oracle import attention_oracle as orc  # noqa: E402

from ch01.gqa import GroupedQueryAttention  # noqa: E402
from ch02.cached_generation import CachedGQA, LayerKVCache  # noqa: E402
from ch02.kv_cache import GQAWithCache, KVCache  # noqa: E402
from ch06.attention_memory import attention_flops, naive_attention  # noqa: E402
from ch06.flash_attention import (FlashAttentionConfig, flash_attention_forward,  # noqa: E402
                                  flash_attention_memory_bytes)
from ch06.online_softmax import online_softmax_with_output  # noqa: E402
from ch07.paged_memory import BlockTable, PagedKVCache  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def checksum(*ts):
    return np.array([float(t.double().sum()) for t in ts] + [float((t.double() ** 2).sum()) for t in ts])


def identity_gqa_weights(mod, Hq, Hkv, D):
    """q_proj = I, k_proj selects dims [0,Hkv*D), v_proj selects [Hkv*D, 2*Hkv*D), o_proj = I.

    With these weights the reference module's output is exactly its attention maths applied to
    q = x, k = x[..., :Hkv*D], v = x[..., Hkv*D:2*Hkv*D] (x*1 + zeros is exact in fp32).
    """
    hid = Hq * D
    with torch.no_grad():
        mod.q_proj.weight.copy_(torch.eye(hid))
        mod.o_proj.weight.copy_(torch.eye(hid))
        mod.k_proj.weight.zero_()
        mod.v_proj.weight.zero_()
        mod.k_proj.weight[:, :Hkv * D] = torch.eye(Hkv * D)
        mod.v_proj.weight[:, Hkv * D:2 * Hkv * D] = torch.eye(Hkv * D)


def split_x(x, Hq, Hkv, D):
    B, N, _ = x.shape
    q = x.view(B, N, Hq, D).transpose(1, 2)
    k = x[..., :Hkv * D].reshape(B, N, Hkv, D).transpose(1, 2)
    v = x[..., Hkv * D:2 * Hkv * D].reshape(B, N, Hkv, D).transpose(1, 2)
    return q, k, v


def to_bhnd(y, Hq, D):
    B, N, _ = y.shape
    return y.view(B, N, Hq, D).transpose(1, 2).contiguous()


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_grad_enabled(False)

    # ---- ch06 flash_attention_forward (non-causal, MHA) ---------------------------------
    cases = {"small": (101, 2, 3, 96, 64), "ragged": (102, 1, 2, 150, 32), "c1": (0xC0FFEE + 1, 1, 8, 512, 64)}
    ch06 = {}
    for name, (seed, B, H, N, D) in cases.items():
        q, k, v = orc.seeded_qkv(seed, B, H, H, N, N, D)
        ref = flash_attention_forward(q, k, v)
        nav = naive_attention(q, k, v)
        o, lse = orc.flash_attention_oracle(q, k, v)
        assert torch.equal(o, ref), f"oracle restatement is not bit-equal to the reference ({name})"
        o2, lse2 = orc.naive_attention_oracle(q, k, v)
        assert (o - o2).abs().max() < 2e-6 and (lse - lse2).abs().max() < 2e-6
        stride = 8 if name == "c1" else 1
        ch06[f"{name}_meta"] = np.array([seed, B, H, N, D, stride])
        ch06[f"{name}_insum"] = checksum(q, k, v)
        ch06[f"{name}_flash"] = ref[:, :, ::stride].numpy()
        ch06[f"{name}_naive"] = nav[:, :, ::stride].numpy()
        ch06[f"{name}_outsum"] = checksum(ref)
    # odd tile hints must give the same function (ch06/flash_attention.py:28-29)
    q, k, v = orc.seeded_qkv(103, 1, 2, 2, 200, 200, 64)
    ref = flash_attention_forward(q, k, v, scale=0.2, config=FlashAttentionConfig(block_q=48, block_k=80))
    o, _ = orc.flash_attention_oracle(q, k, v, scale=0.2, block_q=48, block_k=80)
    assert torch.equal(o, ref)
    ch06["oddtile_meta"] = np.array([103, 1, 2, 200, 64, 1])
    ch06["oddtile_insum"] = checksum(q, k, v)
    ch06["oddtile_flash"] = ref.numpy()
    ch06["flops_c1"] = np.array([attention_flops(1, 8, 512, 64)])
    ch06["membytes_c1"] = np.array([flash_attention_memory_bytes(1, 8, 512, 64)["hbm_bytes"]])
    cfg = FlashAttentionConfig()
    ch06["config_defaults"] = np.array([cfg.block_q, cfg.block_k, cfg.num_warps, cfg.num_stages])
    np.savez_compressed(os.path.join(OUT, "ch06_flash.npz"), **ch06)

    # ---- ch06 online_softmax_with_output (scalar recurrence) -----------------------------
    g = torch.Generator().manual_seed(104)
    x = torch.randn(3, 5, 40, generator=g) * 3
    vv = torch.randn(3, 5, 40, 8, generator=g)
    o, d = online_softmax_with_output(x, vv)
    np.savez_compressed(os.path.join(OUT, "ch06_online.npz"), x=x.numpy(), v=vv.numpy(), o=o.numpy(), d=d.numpy())

    # ---- ch01 GQA causal (square) --------------------------------------------------------
    B, Hq, Hkv, N, D = 2, 8, 2, 80, 16
    g = torch.Generator().manual_seed(105)
    x = torch.randn(B, N, Hq * D, generator=g)
    mod = GroupedQueryAttention(Hq * D, Hq, Hkv)
    identity_gqa_weights(mod, Hq, Hkv, D)
    y_causal = to_bhnd(mod(x, causal=True), Hq, D)
    y_full = to_bhnd(mod(x, causal=False), Hq, D)
    q, k, v = split_x(x, Hq, Hkv, D)
    o, _ = orc.flash_attention_oracle(q, k, v, causal=True)
    assert (o - y_causal).abs().max() < 2e-6, (o - y_causal).abs().max()
    o, _ = orc.flash_attention_oracle(q, k, v, causal=False)
    assert (o - y_full).abs().max() < 2e-6
    np.savez_compressed(os.path.join(OUT, "ch01_gqa.npz"), meta=np.array([B, Hq, Hkv, N, D]), x=x.numpy(),
                        causal=y_causal.numpy(), full=y_full.numpy())

    # ---- ch02 CachedGQA: prefill 37, chunk 11 over the cache, then two decode steps ------
    B, Hq, Hkv, D, Lmax = 2, 8, 2, 16, 64
    g = torch.Generator().manual_seed(106)
    mod = CachedGQA(Hq * D, Hq, Hkv)
    identity_gqa_weights(mod, Hq, Hkv, D)
    cache = LayerKVCache(k=torch.zeros(B, Lmax, Hkv, D), v=torch.zeros(B, Lmax, Hkv, D))
    mod2 = GQAWithCache(Hq * D, Hq, Hkv)
    identity_gqa_weights(mod2, Hq, Hkv, D)
    cache2 = KVCache.create(B, Lmax, Hkv, D, torch.device("cpu"), torch.float32)
    store = {"meta": np.array([B, Hq, Hkv, D, Lmax])}
    pos = 0
    for step, s in enumerate([37, 11, 1, 1]):
        x = torch.randn(B, s, Hq * D, generator=g)
        y = to_bhnd(mod(x, cache, pos), Hq, D)
        y2, _ = mod2(x, cache2)
        assert torch.equal(to_bhnd(y2, Hq, D), y), "ch02 kv_cache.py and cached_generation.py disagree"
        pos += s
        assert cache.seq_len == pos and cache2.seq_len == pos
        q, _, _ = split_x(x, Hq, Hkv, D)
        o, _ = orc.cached_attention_oracle(q, cache.k, cache.v, cache.seq_len)
        assert (o - y).abs().max() < 2e-6, (step, (o - y).abs().max())
        if s > 1:  # the same thing through the tiled recurrence with the offset mask
            kk = cache.k[:, :pos].transpose(1, 2)
            vv = cache.v[:, :pos].transpose(1, 2)
            o3, _ = orc.flash_attention_oracle(q, kk, vv, causal=True, block_q=16, block_k=16)
            assert (o3 - y).abs().max() < 2e-6
        store[f"x{step}"] = x.numpy()
        store[f"y{step}"] = y.numpy()
    store["k_cache"] = cache.k.numpy()
    store["v_cache"] = cache.v.numpy()
    store["seq_len"] = np.array([cache.seq_len])
    np.savez_compressed(os.path.join(OUT, "ch02_cached.npz"), **store)

    # ---- ch07 allocator traces (counts are the contract; set-pop order is not) -----------
    trace = []
    c = PagedKVCache(num_blocks=20, block_size=16, num_layers=2, num_heads=4, head_dim=64, device="cpu")
    assert c.k_cache is None and c.v_cache is None           # D7: no tensors on the host
    ops = [("alloc", 1, 50), ("alloc", 2, 16), ("extend", 1, 14), ("extend", 1, 1), ("alloc", 3, 1),
           ("free", 2, 0), ("extend", 3, 47), ("alloc", 4, 400), ("extend", 9, 1), ("free", 9, 0),
           ("alloc", 5, 0), ("free", 1, 0)]
    for op, rid, n in ops:
        err = 0
        try:
            if op == "alloc":
                c.allocate_blocks(rid, n)
            elif op == "extend":
                c.extend_blocks(rid, n)
            else:
                c.free_blocks_for_request(rid)
        except RuntimeError:
            err = 1
        except KeyError:
            err = 2
        t = c.block_tables.get(rid)
        trace.append([{"alloc": 0, "extend": 1, "free": 2}[op], rid, n, err,
                      -1 if t is None else t.num_blocks(), -1 if t is None else t.num_tokens,
                      c.get_num_free_blocks()])
    usage = c.get_memory_usage()
    bt = BlockTable(request_id=7, block_indices=[3, 1, 2], num_tokens=40)
    np.savez_compressed(os.path.join(OUT, "ch07_paged.npz"), trace=np.array(trace),
                        usage=np.array([usage["total_blocks"], usage["used_blocks"], usage["free_blocks"],
                                        usage["block_size_tokens"], usage["bytes_per_block"]]),
                        bt=np.array([bt.request_id, bt.num_blocks(), bt.num_tokens]))

    # ---- paged decode: ch07 address rule + ch02 maths (composition; no reference consumer) --
    q, kp, vp, table, lens = orc.seeded_paged(107, 3, 8, 2, 32, 16, [77, 16, 1], num_layers=2)
    layer = 1
    o, lse = orc.paged_decode_oracle(q, kp, vp, table, lens, layer=layer)
    # cross-check through the reference module: gather pages -> LayerKVCache -> CachedGQA decode
    for b in range(3):
        L = int(lens[b])
        kg = orc.gather_paged(kp, table[b].tolist(), L, layer)
        for t in range(L):
            p, s = orc.page_address(t, table[b].tolist(), 16)
            assert torch.equal(kg[t], kp[p, layer, s])
    np.savez_compressed(os.path.join(OUT, "paged_decode.npz"), meta=np.array([107, 3, 8, 2, 32, 16, 2, layer]),
                        lens=lens.numpy(), table=table.numpy(), o=o.numpy(), lse=lse.numpy(),
                        insum=checksum(q, kp, vp))
    print("golden vectors written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print(f"  {f}: {os.path.getsize(os.path.join(OUT, f))} bytes")


if __name__ == "__main__":
    main()
