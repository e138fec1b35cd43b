"""Callers of the hot path: the reference's GQA blocks with the attention core on the
kernels, against the reference's own outputs (golden, identity projection weights), and the CUDA-graph
decode step against eager execution."""
import os

import numpy as np
import pytest
import torch

import physics_llm_inference_b200 as pli
from oracle import attention_oracle as orc

pytestmark = pytest.mark.gpu


def _identity_weights(mod, Hq, Hkv, D):
    hid = Hq * D
    with torch.no_grad():
        mod.q_proj.weight.copy_(torch.eye(hid))
        mod.o_proj.weight.copy_(torch.eye(hid))
        mod.k_proj.weight.zero_()
        mod.v_proj.weight.zero_()
        mod.k_proj.weight[:, :Hkv * D] = torch.eye(Hkv * D)
        mod.v_proj.weight[:, Hkv * D:2 * Hkv * D] = torch.eye(Hkv * D)


def _bhnd(y, Hq, D):
    B, N, _ = y.shape
    return y.view(B, N, Hq, D).transpose(1, 2)


def test_gqa_module_matches_reference_module(golden_dir):
    """ch01.gqa.GroupedQueryAttention, unmodified, produced tests/golden/ch01_gqa.npz with these weights."""
    g = np.load(os.path.join(golden_dir, "ch01_gqa.npz"))
    B, Hq, Hkv, N, D = [int(x) for x in g["meta"]]
    mod = pli.GroupedQueryAttention(Hq * D, Hq, Hkv).cuda()
    _identity_weights(mod, Hq, Hkv, D)
    x = torch.from_numpy(g["x"]).cuda()
    with torch.no_grad():
        torch.backends.cuda.matmul.allow_tf32 = False
        y_c = _bhnd(mod(x, causal=True), Hq, D).cpu()
        y_f = _bhnd(mod(x, causal=False), Hq, D).cpu()
    assert (y_c - torch.from_numpy(g["causal"])).abs().max().item() <= 1e-3
    assert (y_f - torch.from_numpy(g["full"])).abs().max().item() <= 1e-3
    assert mod.kv_cache_size_per_token(torch.float16) == 2 * Hkv * D * 2     # ch01/test_ch01.py:79-88


def test_cached_gqa_module_matches_reference_module(golden_dir):
    """ch02.cached_generation.CachedGQA: prefill 37, chunk 11, decode 1, decode 1 (tests/golden/ch02_cached.npz)."""
    g = np.load(os.path.join(golden_dir, "ch02_cached.npz"))
    B, Hq, Hkv, D, Lmax = [int(x) for x in g["meta"]]
    mod = pli.CachedGQA(Hq * D, Hq, Hkv).cuda()
    _identity_weights(mod, Hq, Hkv, D)
    cache = pli.create_caches(1, B, Lmax, Hkv, D, "cuda", torch.float32)[0]
    torch.backends.cuda.matmul.allow_tf32 = False
    pos = 0
    with torch.no_grad():
        for step in range(4):
            x = torch.from_numpy(g[f"x{step}"]).cuda()
            y = _bhnd(mod(x, cache, pos), Hq, D).cpu()
            pos += x.shape[1]
            assert cache.seq_len == pos                                      # ch02/test_ch02.py:165-204
            assert (y - torch.from_numpy(g[f"y{step}"])).abs().max().item() <= 1e-3, step
    assert torch.equal(cache.k.cpu(), torch.from_numpy(g["k_cache"]))


@pytest.mark.parametrize("paged", [False, True])
def test_decode_step_under_cuda_graph(paged):
    B, Hq, Hkv, D, bs, L0, steps = 4, 8, 2, 128, 16, 70, 5
    g = torch.Generator().manual_seed(41)
    dt = torch.bfloat16
    if paged:
        pages_per = 8
        P = B * pages_per
        kc = torch.zeros(P, 1, bs, Hkv, D, dtype=dt, device="cuda")
        vc = torch.zeros_like(kc)
        table = torch.randperm(P, generator=g).to(torch.int32).view(B, pages_per).cuda()
    else:
        kc = torch.zeros(B, 128, Hkv, D, dtype=dt, device="cuda")
        vc = torch.zeros_like(kc)
        table = None
    k_hist = torch.randn(B, L0 + steps, Hkv, D, generator=g).to(dt).cuda()
    v_hist = torch.randn(B, L0 + steps, Hkv, D, generator=g).to(dt).cuda()
    pli.kv_append(kc, vc, k_hist[:, :L0], v_hist[:, :L0], 0, block_tables=table)
    runner = pli.DecodeGraphRunner(kc, vc, Hq, block_tables=table)
    lens = torch.full((B,), L0, dtype=torch.int32, device="cuda")
    assert runner.capture(lens)
    for i in range(steps):
        q = torch.randn(B, Hq, 1, D, generator=g).to(dt).cuda()
        out = runner.run(q, k_hist[:, L0 + i:L0 + i + 1], v_hist[:, L0 + i:L0 + i + 1])
        L = L0 + i + 1
        ref, _ = orc.cached_attention_oracle(q, k_hist, v_hist, L)
        assert (out.float().cpu() - ref[:, :, 0]).abs().max().item() <= 2e-2, i
    assert int(runner.seq_lens[0]) == L0 + steps


@pytest.mark.parametrize("world", [1, 2, 4])
def test_tensor_parallel_gqa_shards_sum_to_the_full_block(world):
    """KV-head-sharded GQA block: the ranks' partial outputs (what the all-reduce adds up) sum to the unsharded
    block's output; bf16 weights, random init."""
    torch.manual_seed(3)
    Hq, Hkv, D, B, N = 16, 4, 128, 2, 300
    full = pli.GroupedQueryAttention(Hq * D, Hq, Hkv).cuda().to(torch.bfloat16)
    x = torch.randn(B, N, Hq * D, device="cuda", dtype=torch.bfloat16) * 0.5
    with torch.no_grad():
        ref = full(x, causal=True).float()
        parts = [pli.TensorParallelGQA.from_full(full, world, r).partial_forward(x).float() for r in range(world)]
        one = pli.TensorParallelGQA.from_full(full, 1, 0)(x).float()
    assert torch.equal(one, ref)
    tol = 2e-2 * max(1.0, ref.abs().max().item())
    assert (sum(parts) - ref).abs().max().item() <= tol
    with pytest.raises(ValueError):
        pli.TensorParallelGQA(Hq * D, Hq, Hkv, world_size=3, rank=0)


def test_cached_generate_on_the_kernels_gives_the_reference_tokens(golden_dir):
    """F4 tail: the reference's `cached_generate` loop (ch02/cached_generation.py:208-274) over a model whose attention
    block is this package's `CachedGQA` (prefill on the causal kernel, every decode step on the split-KV kernel, K/V
    appended by pli_kv_append), fp32, greedy (temperature 1e-6): token ids identical to what the UNMODIFIED reference
    model generated on the CPU (tests/golden/r2_ch02_generate.npz, oracle/make_golden_r2.py)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from tiny_model import TinyCachedModel, cached_generate
    g = np.load(os.path.join(golden_dir, "r2_ch02_generate.npz"))
    vocab, hidden, layers, heads, kv_heads, inter = [int(x) for x in g["cfg"]]
    seed, batch, prompt_len, n_new = [int(x) for x in g["meta"]]
    torch.manual_seed(seed)
    model = TinyCachedModel(pli.CachedGQA, vocab, hidden, layers, heads, kv_heads, inter)
    wsum = np.array([float(p.detach().double().sum()) for p in model.parameters()] +
                    [float((p.detach().double() ** 2).sum()) for p in model.parameters()])
    assert np.allclose(wsum, g["wsum"], rtol=0, atol=1e-9), "seeded weights differ from the reference model's"
    model = model.cuda()
    torch.backends.cuda.matmul.allow_tf32 = False
    prompt = torch.from_numpy(g["prompt"]).cuda()
    make = lambda b, n: pli.create_caches(layers, b, n, kv_heads, hidden // heads, "cuda", torch.float32)  # noqa: E731
    pli.reset_launch_count()
    tokens, logits = cached_generate(model, prompt, n_new, make, temperature=1e-6)
    assert pli.launch_count() >= layers * (2 * n_new)            # an append and an attention launch per layer per step
    assert torch.equal(tokens.cpu(), torch.from_numpy(g["tokens"]))
    assert (logits[:, -1, :].cpu() - torch.from_numpy(g["last_logits"])).abs().max().item() <= 1e-3


def test_radix_hit_shares_pages_and_decodes_like_an_unshared_request(golden_dir):
    """F3: `(matched, kv_indices)` as the reference's RadixCache.match_prefix returned them (tests/golden/
    r2_ch07_radix.json) -> PagedKVCache.share_prefix: whole pages aliased, the partly filled page copied; the child then
    appends its own tokens and decodes bit-identically to a request that holds all of its tokens in private pages."""
    import json
    with open(os.path.join(golden_dir, "r2_ch07_radix.json")) as f:
        r = json.load(f)
    bs, Hkv, Hq, D, layers = r["block_size"], 2, 8, 128, 2
    cache = pli.PagedKVCache(num_blocks=32, block_size=bs, num_layers=layers, num_heads=Hkv, head_dim=D, dtype=torch.bfloat16)
    # request A lives in the pages the scenario names
    for pg in r["pages_a"]:
        cache.free_blocks.remove(pg)
    cache.block_tables[1] = pli.BlockTable(request_id=1, block_indices=list(r["pages_a"]), num_tokens=0)
    g = torch.Generator(device="cuda").manual_seed(91)
    n_a = len(r["a_tokens"])
    ka = torch.randn(layers, 1, n_a, Hkv, D, device="cuda", generator=g).bfloat16()
    va = torch.randn(layers, 1, n_a, Hkv, D, device="cuda", generator=g).bfloat16()
    cache.block_tables[1].num_tokens = n_a
    for layer in range(layers):
        cache.append([1], ka[layer], va[layer], layer=layer, extend=False)
    assert cache.kv_indices(1) == r["a_kv"]
    hit = r["queries"]["b_shares_53"]
    matched, n_b = hit["matched"], len(hit["tokens"])
    free_before = cache.get_num_free_blocks()
    tb = cache.share_prefix(2, matched, hit["kv_indices"])
    assert tb.num_tokens == 53 and tb.block_indices[:3] == r["pages_a"][:3] and tb.block_indices[3] not in r["pages_a"]
    assert cache.get_num_free_blocks() == free_before - 1
    # B's own tokens after the shared prefix, and the same request built privately (prefix K/V recomputed = copied)
    kb = torch.randn(layers, 1, n_b - matched, Hkv, D, device="cuda", generator=g).bfloat16()
    vb = torch.randn(layers, 1, n_b - matched, Hkv, D, device="cuda", generator=g).bfloat16()
    for layer in range(layers):
        cache.append([2], kb[layer], vb[layer], layer=layer, extend=(layer == 0))
        cache.append([3], torch.cat([ka[layer][:, :matched], kb[layer]], 1), torch.cat([va[layer][:, :matched], vb[layer]], 1),
                     layer=layer, extend=(layer == 0))
    q = torch.randn(1, Hq, 1, D, device="cuda", generator=g).bfloat16()
    for layer in range(layers):
        o_shared = pli.decode_with_paged(q, cache, [2], layer=layer)
        o_private = pli.decode_with_paged(q, cache, [3], layer=layer)
        assert torch.equal(o_shared, o_private)
        ro, _ = orc.cached_attention_oracle(q, torch.cat([ka[layer][:, :matched], kb[layer]], 1),
                                            torch.cat([va[layer][:, :matched], vb[layer]], 1), n_b)
        assert (o_shared.float().cpu() - ro).abs().max().item() <= 2e-2
    # A is untouched by B's appends (B wrote into its private copy of the partly shared page), and freeing A keeps the
    # aliased pages alive for B
    o_a = pli.decode_with_paged(q, cache, [1], layer=0)
    ro, _ = orc.cached_attention_oracle(q, ka[0], va[0], n_a)
    assert (o_a.float().cpu() - ro).abs().max().item() <= 2e-2
    cache.free_blocks_for_request(1)
    assert all(pg not in cache.free_blocks for pg in r["pages_a"][:3]) and all(pg in cache.free_blocks for pg in r["pages_a"][3:])
    assert torch.equal(pli.decode_with_paged(q, cache, [2], layer=1), pli.decode_with_paged(q, cache, [3], layer=1))
    cache.free_blocks_for_request(2)
    assert all(pg in cache.free_blocks for pg in r["pages_a"])
