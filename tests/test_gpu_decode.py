"""GPU parity tests of the decode path (contiguous ch02 cache and ch07 paged pools), the KV append
and the page addressing, through the C ABI, against the CPU oracle."""
import os

import numpy as np
import pytest
import torch

import physics_llm_inference_b200 as pli
from oracle import attention_oracle as orc

pytestmark = pytest.mark.gpu
TOL = {torch.float32: 1e-3, torch.bfloat16: 2e-2, torch.float16: 2e-2}
LSE_TOL = 1e-3


def _paged_case(seed, B, Hq, Hkv, D, bs, lens, dtype, n_layers=1, layer=0, num_splits=None):
    q, kp, vp, table, lens_t = orc.seeded_paged(seed, B, Hq, Hkv, D, bs, lens, num_layers=n_layers)
    qd, kd, vd = q.to(dtype).cuda(), kp.to(dtype).cuda(), vp.to(dtype).cuda()
    o, lse = pli.flash_decode(qd, kd, vd, lens_t.cuda(), block_tables=table.cuda(), layer=layer, return_lse=True,
                              num_splits=num_splits, max_seq_len=max(lens))
    torch.cuda.synchronize()
    ro, rlse = orc.paged_decode_oracle(qd, kd, vd, table, lens_t, layer=layer)
    assert o.shape == qd.shape and o.dtype == dtype
    return (o.float().cpu() - ro).abs().max().item(), (lse.cpu() - rlse[:, :, 0]).abs().max().item(), kd, table


@pytest.mark.parametrize("bs,D,G", [(16, 32, 4), (16, 128, 4), (7, 48, 2), (32, 64, 1)])
def test_paged_decode_fp32_simt(bs, D, G):
    eo, el, kd, table = _paged_case(31, 3, 2 * G, 2, D, bs, [77, bs, 1], torch.float32, n_layers=2, layer=1)
    assert pli.decode_kernel_kind(kd, table) == "simt"
    assert eo <= TOL[torch.float32] and el <= LSE_TOL, (eo, el)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("bs,D,G,lens,splits", [
    (16, 128, 4, [4096, 100, 1, 17, 2048], None),      # C3-shaped pages, ragged batch
    (16, 128, 4, [1000, 999], 1),
    (16, 128, 4, [1000, 999], 7),
    (16, 64, 8, [513, 64, 65], None),
    (8, 128, 1, [300, 9], 2),                          # MHA, small pages
    (64, 128, 2, [700, 64, 1], None),
    (128, 64, 16, [900, 128, 129], 3),                 # pages larger than a stage, 16 q heads per kv head
    (256, 128, 4, [1025], None),
    (16, 128, 32, [555, 16], 2),                       # 32 q heads per kv head: two 16-row chunks
])
def test_paged_decode_tma(bs, D, G, lens, splits, dtype):
    eo, el, kd, table = _paged_case(32, len(lens), 2 * G, 2, D, bs, lens, dtype, n_layers=2, layer=1, num_splits=splits)
    assert pli.decode_kernel_kind(kd, table) == "mma_tma"
    assert eo <= TOL[dtype] and el <= LSE_TOL, (eo, el)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_contiguous_decode(dtype):
    B, Hq, Hkv, D, Lmax = 3, 8, 2, 128, 700
    g = torch.Generator().manual_seed(33)
    q = torch.randn(B, Hq, 1, D, generator=g).to(dtype).cuda()
    kc = torch.randn(B, Lmax, Hkv, D, generator=g).to(dtype).cuda()
    vc = torch.randn(B, Lmax, Hkv, D, generator=g).to(dtype).cuda()
    for L in (1, 63, 64, 65, 500, 700):
        o, lse = pli.flash_decode(q, kc, vc, L, return_lse=True)
        ro, rlse = orc.cached_attention_oracle(q, kc, vc, L)
        assert (o.float().cpu() - ro).abs().max().item() <= TOL[dtype], L
        assert (lse.cpu() - rlse[:, :, 0]).abs().max().item() <= LSE_TOL, L
    lens = torch.tensor([700, 3, 257], dtype=torch.int32)
    o = pli.flash_decode(q, kc, vc, lens.cuda())
    ro, _ = orc.cached_attention_oracle(q, kc, vc, lens)
    assert (o.float().cpu() - ro).abs().max().item() <= TOL[dtype]


def test_garbage_beyond_seq_len_is_never_used():
    """Storage past seq_len (rest of the last page, unused pages) may hold NaN/Inf: results must not change."""
    B, Hq, Hkv, D, bs = 2, 8, 2, 128, 16
    lens = [100, 37]
    q, kp, vp, table, lens_t = orc.seeded_paged(34, B, Hq, Hkv, D, bs, lens)
    kd, vd = kp.bfloat16().cuda(), vp.bfloat16().cuda()
    o0 = pli.flash_decode(q.bfloat16().cuda(), kd, vd, lens_t.cuda(), block_tables=table.cuda())
    used = torch.zeros(kd.shape[0], bs, dtype=torch.bool)
    for b, L in enumerate(lens):
        for t in range(L):
            used[int(table[b, t // bs]), t % bs] = True
    kd[:, 0][~used.cuda()] = float("nan")          # (P, bs) mask over the single layer
    vd[:, 0][~used.cuda()] = float("inf")
    o1 = pli.flash_decode(q.bfloat16().cuda(), kd, vd, lens_t.cuda(), block_tables=table.cuda())
    assert torch.equal(o0, o1)


def test_page_addressing_bit_exact():
    """token t -> page table[t // bs], slot t % bs (ch07/paged_memory.py:54,84-86), bit for bit."""
    B, Hkv, D, bs, n_layers = 3, 2, 64, 16, 3
    lens = [50, 16, 33]
    _, kp, _, table, lens_t = orc.seeded_paged(35, B, 4, Hkv, D, bs, lens, num_layers=n_layers)
    for dtype in (torch.float32, torch.bfloat16):
        pool = kp.to(dtype).cuda()
        for layer in range(n_layers):
            got = pli.paged_gather(pool, table.cuda(), lens_t.cuda(), max(lens), layer=layer).cpu()
            for b, L in enumerate(lens):
                ref = orc.gather_paged(pool.cpu(), table[b].tolist(), L, layer)
                assert torch.equal(got[b, :L], ref)
                assert torch.count_nonzero(got[b, L:]) == 0


def test_kv_append_roundtrip_and_reference_objects(golden_dir):
    """ch02 golden: prefill 37, chunk 11, decode 1, decode 1 through KVCache.update + the kernels."""
    g = np.load(os.path.join(golden_dir, "ch02_cached.npz"))
    B, Hq, Hkv, D, Lmax = [int(x) for x in g["meta"]]
    cache = pli.KVCache.create(B, Lmax, Hkv, D, torch.device("cuda"), torch.float32)
    layer_cache = pli.create_caches(1, B, Lmax, Hkv, D, "cuda", torch.float32)[0]
    for step in range(4):
        x = torch.from_numpy(g[f"x{step}"]).cuda()
        s = x.shape[1]
        q = x.view(B, s, Hq, D).transpose(1, 2)
        k_new = x[..., :Hkv * D].reshape(B, s, Hkv, D)
        v_new = x[..., Hkv * D:2 * Hkv * D].reshape(B, s, Hkv, D)
        k_full, v_full = cache.update(k_new, v_new)
        layer_cache.update(k_new, v_new)
        assert k_full.shape[1] == cache.seq_len == layer_cache.seq_len
        ref = torch.from_numpy(g[f"y{step}"])
        if s > 1:   # chunk over the cache: offset causal mask (ch02/cached_generation.py:85-91)
            o = pli.flash_attention_forward(q, k_full.transpose(1, 2), v_full.transpose(1, 2), causal=True)
        else:
            o = pli.decode_with_cache(q, cache)
            o2 = pli.decode_with_cache(q, layer_cache)
            assert torch.equal(o, o2)
        assert (o.cpu() - ref).abs().max().item() <= 1e-3, step
    assert torch.equal(cache.k_cache.cpu(), torch.from_numpy(g["k_cache"]))
    assert torch.equal(cache.v_cache.cpu(), torch.from_numpy(g["v_cache"]))
    assert torch.equal(layer_cache.k, cache.k_cache)


def test_golden_paged_decode(golden_dir):
    g = np.load(os.path.join(golden_dir, "paged_decode.npz"))
    seed, B, Hq, Hkv, D, bs, n_layers, layer = [int(x) for x in g["meta"]]
    lens = [int(x) for x in g["lens"]]
    q, kp, vp, table, lens_t = orc.seeded_paged(seed, B, Hq, Hkv, D, bs, lens, num_layers=n_layers)
    o, lse = pli.flash_decode(q.cuda(), kp.cuda(), vp.cuda(), lens_t.cuda(), block_tables=table.cuda(), layer=layer,
                              return_lse=True)
    assert (o.cpu() - torch.from_numpy(g["o"])).abs().max().item() <= 1e-3
    assert (lse.cpu() - torch.from_numpy(g["lse"])[:, :, 0]).abs().max().item() <= 1e-3


def test_paged_cache_object_write_then_read():
    """PagedKVCache mirror end to end: allocate -> append (kernel) -> decode (kernel) == oracle."""
    Hq, Hkv, D, bs, layers = 8, 2, 128, 16, 2
    cache = pli.PagedKVCache(num_blocks=64, block_size=bs, num_layers=layers, num_heads=Hkv, head_dim=D,
                             dtype=torch.bfloat16, device="cuda")
    g = torch.Generator().manual_seed(36)
    lens = {10: 45, 11: 16, 12: 1}
    kv = {}
    for rid, L in lens.items():
        ks = torch.randn(1, L, Hkv, D, generator=g).bfloat16().cuda()
        vs = torch.randn(1, L, Hkv, D, generator=g).bfloat16().cuda()
        cache.append([rid], ks[:, :L - 1], vs[:, :L - 1], layer=1) if L > 1 else None
        cache.append([rid], ks[:, L - 1:], vs[:, L - 1:], layer=1)          # grows by one token (extend_blocks)
        kv[rid] = (ks, vs)
        assert cache.block_tables[rid].num_tokens == L
        assert cache.block_tables[rid].num_blocks() == (L + bs - 1) // bs
    rids = list(lens)
    q = torch.randn(len(rids), Hq, 1, D, generator=g).bfloat16().cuda()
    o = pli.decode_with_paged(q, cache, rids, layer=1)
    for i, rid in enumerate(rids):
        ro, _ = orc.cached_attention_oracle(q[i:i + 1], kv[rid][0], kv[rid][1], lens[rid])
        assert (o[i:i + 1].float().cpu() - ro).abs().max().item() <= 2e-2
    assert torch.count_nonzero(cache.k_cache[:, 0]) == 0                    # layer 0 untouched
    freed = cache.free_blocks_for_request(10)
    assert freed == 3 and cache.get_num_free_blocks() == 64 - 1 - 1


def test_c3_full_size_properties():
    """BASELINE C3 (B64, ctx 4096, 32q/8kv, D128, 16-token pages): sampled sequences vs the oracle,
    V = 1 => O = 1, and invariance under a re-permutation of the physical pages."""
    B, Hq, Hkv, D, bs, L = 64, 32, 8, 128, 16, 4096
    pages_per = L // bs
    P = B * pages_per + 5
    g = torch.Generator(device="cuda").manual_seed(37)
    kp = torch.randn(P, 1, bs, Hkv, D, device="cuda", generator=g).bfloat16()
    vp = torch.randn(P, 1, bs, Hkv, D, device="cuda", generator=g).bfloat16()
    q = torch.randn(B, Hq, 1, D, device="cuda", generator=g).bfloat16()
    perm = torch.randperm(P, generator=torch.Generator().manual_seed(38))[:B * pages_per].to(torch.int32)
    table = perm.view(B, pages_per).cuda()
    lens = torch.full((B,), L, dtype=torch.int32, device="cuda")
    o, lse = pli.flash_decode(q, kp, vp, lens, block_tables=table, return_lse=True, max_seq_len=L)
    for b in (0, 31, 63):
        ro, rlse = orc.paged_decode_oracle(q[b:b + 1], kp, vp, table[b:b + 1].cpu(), [L])
        assert (o[b:b + 1].float().cpu() - ro).abs().max().item() <= 2e-2
        assert (lse[b:b + 1].cpu() - rlse[:, :, 0]).abs().max().item() <= 1e-3
    o1 = pli.flash_decode(q, kp, torch.ones_like(vp), lens, block_tables=table, max_seq_len=L)
    assert (o1.float() - 1).abs().max().item() <= 1e-2
    # move every page somewhere else: same logical cache, different physical layout => identical output
    perm2 = torch.randperm(P, generator=torch.Generator().manual_seed(39))
    inv = torch.empty_like(perm2)
    inv[perm2] = torch.arange(P)
    kp2, vp2 = kp[perm2.cuda()], vp[perm2.cuda()]          # new_pool[i] = old_pool[perm2[i]]
    table2 = inv.cuda()[table.long()].to(torch.int32)
    o2 = pli.flash_decode(q, kp2, vp2, lens, block_tables=table2, max_seq_len=L)
    assert torch.equal(o2, o)


def test_shared_prefix_pages_alias():
    """F3: two requests whose block tables alias the same physical prefix pages decode like unshared ones."""
    Hq, Hkv, D, bs = 8, 2, 128, 16
    cache = pli.PagedKVCache(num_blocks=32, block_size=bs, num_layers=1, num_heads=Hkv, head_dim=D,
                             dtype=torch.bfloat16, device="cuda")
    g = torch.Generator().manual_seed(51)
    L0, extra = 70, 9                                       # 4 full pages + 6 tokens
    k0 = torch.randn(1, L0, Hkv, D, generator=g).bfloat16().cuda()
    v0 = torch.randn(1, L0, Hkv, D, generator=g).bfloat16().cuda()
    cache.append([1], k0, v0)
    cache.fork_request(1, 2, 67)                            # child shares 67 tokens: 4 aliased pages + copied tail
    assert cache.block_tables[2].block_indices[:4] == cache.block_tables[1].block_indices[:4]
    k2 = torch.randn(1, extra, Hkv, D, generator=g).bfloat16().cuda()
    v2 = torch.randn(1, extra, Hkv, D, generator=g).bfloat16().cuda()
    cache.append([2], k2, v2)
    q = torch.randn(2, Hq, 1, D, generator=g).bfloat16().cuda()
    o = pli.decode_with_paged(q, cache, [1, 2])
    ref1, _ = orc.cached_attention_oracle(q[:1], k0, v0, L0)
    kk = torch.cat([k0[:, :67], k2], 1)
    vv = torch.cat([v0[:, :67], v2], 1)
    ref2, _ = orc.cached_attention_oracle(q[1:], kk, vv, 67 + extra)
    assert (o[:1].float().cpu() - ref1).abs().max().item() <= 2e-2
    assert (o[1:].float().cpu() - ref2).abs().max().item() <= 2e-2
    # the parent's pages were not disturbed by the child's appends
    got = pli.paged_gather(cache.k_cache, *cache.block_table_tensor([1]), L0)
    assert torch.equal(got[0], k0[0])


def test_mixed_prefill_decode_batch():
    """One step of a ch08 mixed batch: two prompts of different length prefilled while three requests decode."""
    torch.manual_seed(11)
    Hq, Hkv, D, bs = 8, 2, 128, 16
    dev = torch.device("cuda")
    cache = pli.PagedKVCache(num_blocks=256, block_size=bs, num_layers=1, num_heads=Hkv, head_dim=D,
                             dtype=torch.bfloat16, device="cuda")
    history = {}                                    # request id -> (k, v) of everything cached, (n, Hkv, D)

    def feed(rid, n):
        k = torch.randn(1, n, Hkv, D, device=dev, dtype=torch.bfloat16)
        v = torch.randn(1, n, Hkv, D, device=dev, dtype=torch.bfloat16)
        cache.append([rid], k, v)
        pk, pv = history.get(rid, (k[:, :0], v[:, :0]))
        history[rid] = (torch.cat([pk, k], 1), torch.cat([pv, v], 1))

    for rid, n in ((3, 500), (4, 77), (5, 1300)):   # decoding requests with some context already cached
        feed(rid, n)
    prefill_ids, prefill_lens, decode_ids = [10, 11], [333, 90], [3, 4, 5]
    for rid, n in zip(prefill_ids, prefill_lens):   # this step's K/V: whole prompts for the new requests ...
        feed(rid, n)
    for rid in decode_ids:                          # ... and one token for each decoding request
        feed(rid, 1)
    T = sum(prefill_lens) + len(decode_ids)
    q = torch.randn(T, Hq, D, device=dev, dtype=torch.bfloat16)
    o = pli.mixed_batch_attention(q, cache, prefill_ids, prefill_lens, decode_ids)
    assert o.shape == q.shape
    at = 0
    for rid, n in list(zip(prefill_ids, prefill_lens)) + [(r, 1) for r in decode_ids]:
        k, v = history[rid]
        ro, _ = orc.cached_attention_oracle(q[at:at + n].transpose(0, 1).unsqueeze(0).float().cpu(),
                                            k.float().cpu(), v.float().cpu(), k.shape[1])
        assert (o[at:at + n].float().cpu().transpose(0, 1) - ro[0]).abs().max().item() <= 2e-2, rid
        at += n


@pytest.mark.parametrize("splits", [None, 3, 4, 16])
def test_decode_peer_output_single_rank(splits):
    """The fused-gather entry with a world of one: same result as flash_decode, over several steps (buffers
    alternate by epoch), through both the direct-output kernel and the combine kernel."""
    B, Hq, Hkv, D, bs, L = 8, 8, 2, 128, 16, 700
    q, kp, vp, table, lens = orc.seeded_paged(31, B, Hq, Hkv, D, bs, [L] * B, dtype=torch.bfloat16)
    q, kp, vp, table, lens = q.cuda(), kp.cuda(), vp.cuda(), table.cuda(), lens.cuda()
    shard = pli.make_shard(0, 1, Hq, Hkv, B)
    po = pli.PeerOutput(B, Hq, D, torch.bfloat16, shard)
    ref = pli.flash_decode(q, kp, vp, lens, block_tables=table, max_seq_len=L, num_splits=splits)[:, :, 0]
    launches = pli.launch_count()
    for step in range(3):
        o, lse = pli.flash_decode(q, kp, vp, lens, block_tables=table, max_seq_len=L, num_splits=splits, peer_out=po,
                                  return_lse=True)
        assert o.shape == (B, Hq, D)
        assert torch.equal(o, ref), step
    assert po.epoch == 3 and po.mode == "gather"
    assert pli.launch_count() - launches == 6, "decode (+ merge) kernel and the programmatically launched publish / wait"
    # captured with a consumer in the same graph: one buffer at a fixed address, every replay is seen
    ws = pli.decode_workspace(B, Hq, D, splits or pli.decode_num_splits(B, Hkv, L), "cuda")
    consumed = torch.zeros(B, Hq, D, device="cuda")
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        pli.flash_decode(q, kp, vp, lens, block_tables=table, max_seq_len=L, num_splits=splits, peer_out=po, workspace=ws)
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            og = pli.flash_decode(q, kp, vp, lens, block_tables=table, max_seq_len=L, num_splits=splits, peer_out=po,
                                  workspace=ws)
            consumed.copy_(og.float() * 2)
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(3):
        consumed.zero_()
        g.replay()
        o = po.advance()
        torch.cuda.synchronize()
        assert torch.equal(o, ref) and torch.equal(consumed, ref.float() * 2)
    assert po.epoch == 7


def test_decode_plan_matches_flash_decode_and_replays_in_a_graph():
    """`DecodePlan`: the call marshalled once, launched many times; same bits as `flash_decode`, follows in-place updates
    of its buffers, and captures into a CUDA graph."""
    B, Hq, Hkv, D, bs = 5, 8, 2, 128, 16
    lens_l = [300, 17, 1024, 64, 999]
    q, kp, vp, table, lens = orc.seeded_paged(77, B, Hq, Hkv, D, bs, lens_l, dtype=torch.bfloat16)
    qd, kd, vd, td, ld = q.cuda(), kp.cuda(), vp.cuda(), table.cuda(), lens.cuda()
    ref, rl = pli.flash_decode(qd, kd, vd, ld, block_tables=td, max_seq_len=1024, return_lse=True)
    out = torch.empty(B, Hq, D, device="cuda", dtype=torch.bfloat16)
    plan = pli.DecodePlan(qd, kd, vd, ld, block_tables=td, max_seq_len=1024, out=out, return_lse=True)
    o, l = plan()
    assert o.data_ptr() == out.data_ptr() and torch.equal(o, ref) and torch.equal(l, rl)
    # in-place update of the planned buffers: shorter sequences, another query
    ld.copy_(torch.tensor([100, 17, 500, 1, 640], dtype=torch.int32))
    qd.mul_(-0.5)
    ref2 = pli.flash_decode(qd, kd, vd, ld, block_tables=td, max_seq_len=1024)
    o2, _ = plan()
    assert torch.equal(o2, ref2) and not torch.equal(ref2, ref)
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        plan()
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            plan()
    torch.cuda.current_stream().wait_stream(side)
    qd.add_(0.25)
    out.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out.unsqueeze(2), pli.flash_decode(qd, kd, vd, ld, block_tables=td, max_seq_len=1024))
    with pytest.raises(TypeError):
        pli.DecodePlan(qd, kd, vd, 300, block_tables=td)


@pytest.mark.parametrize("B,G,D,L,splits", [(1, 4, 128, 32768, 37), (3, 8, 64, 5000, 64), (2, 16, 128, 2000, 5),
                                            (4, 32, 128, 777, 3),
                                            (2, 16, 128, 4500, 64),      # 16 rows x 64 splits x 512 B: merged in three groups
                                            (2, 4, 128, 3000, 4), (2, 8, 64, 3000, 8), (1, 16, 128, 2000, 2),
                                            (3, 32, 128, 1500, 8),       # 2 / 4 / 8 splits: one cluster per unit (DSMEM merge)
                                            (1, 4, 128, 32768, 32), (2, 8, 64, 6000, 24), (2, 16, 128, 3000, 12),
                                            (3, 4, 128, 2500, 6), (1, 4, 128, 9000, 40)])   # clusters, then partials of clusters
def test_fused_combine_matches_the_two_pass_path_on_a_dirty_workspace(B, G, D, L, splits):
    """Several splits per sequence: the split-KV kernel's last-arriving CTA of every unit merges the partials itself (no
    combine launch).  Its arrival counters live in the caller's workspace and need no initialisation: the same answer
    from a workspace full of 0xFF bytes / random bits / the previous launch's state, under CUDA-graph replay too, and
    equal (to fp32 rounding of the merge order) to the two-pass C-ABI path pli_decode_splitkv + pli_decode_combine."""
    from physics_llm_inference_b200 import _lib
    Hkv, bs = 2, 16
    Hq = Hkv * G
    lens_l = [L] + [max(1, L // (i + 2)) for i in range(B - 1)]
    q, kp, vp, table, lens = orc.seeded_paged(91, B, Hq, Hkv, D, bs, lens_l, dtype=torch.bfloat16)
    qd, kd, vd, td, ld = q.cuda(), kp.cuda(), vp.cuda(), table.cuda(), lens.cuda()
    lib = _lib.load()
    need = int(lib.pli_decode_workspace_bytes(B, Hq, D, splits))
    # two-pass reference through the C ABI
    ws2 = torch.empty(need // 4 + 2, dtype=torch.float32, device="cuda")
    o2 = torch.empty(B, Hq, D, device="cuda", dtype=torch.bfloat16)
    l2 = torch.empty(B, Hq, device="cuda", dtype=torch.float32)
    stream = torch.cuda.current_stream().cuda_stream
    q3 = qd[:, :, 0, :]
    _lib.check(lib.pli_decode_splitkv(q3.data_ptr(), kd.data_ptr(), vd.data_ptr(), td.data_ptr(), ld.data_ptr(), B, Hq, Hkv, D,
                                      L, bs, td.stride(0), 0, kd.shape[0], _lib.i64(q3.stride(0), q3.stride(1)),
                                      _lib.i64(*kd.stride()[:4]), D ** -0.5, _lib.dtype_code(qd.dtype), splits,
                                      ws2.data_ptr(), ws2.numel() * 4, stream))
    _lib.check(lib.pli_decode_combine(ws2.data_ptr(), o2.data_ptr(), l2.data_ptr(), B, Hq, D, splits,
                                      _lib.i64(o2.stride(0), o2.stride(1)), _lib.dtype_code(qd.dtype), stream))
    ro, rlse = orc.paged_decode_oracle(qd, kd, vd, table, lens)
    assert (o2.float().cpu().unsqueeze(2) - ro).abs().max().item() <= 2e-2
    launches = pli.launch_count()
    for fill in ("ff", "random", "reuse", "reuse"):
        ws = torch.empty(need // 4 + 2, dtype=torch.float32, device="cuda")
        if fill == "ff":
            ws.view(torch.uint8).fill_(0xFF)
        elif fill == "random":
            ws.view(torch.int32).random_(-2**31, 2**31 - 1)
        else:
            ws = prev_ws                                              # noqa: F821  (state left by the previous launch)
        out = torch.full((B, Hq, D), float("nan"), device="cuda", dtype=torch.bfloat16)
        o, lse = pli.flash_decode(qd, kd, vd, ld, block_tables=td, max_seq_len=L, num_splits=splits, workspace=ws, out=out,
                                  return_lse=True)
        torch.cuda.synchronize()
        assert (o.float() - o2.float().unsqueeze(2)).abs().max().item() <= 1e-2 * o2.float().abs().max().item() + 1e-3, fill
        assert (lse - l2).abs().max().item() <= 1e-4, fill
        assert (lse.cpu() - rlse[:, :, 0]).abs().max().item() <= LSE_TOL
        prev_ws = ws                                                  # noqa: F841
    assert pli.launch_count() - launches == 4, "one launch per fused call (no combine pass)"
    # graph replays: the launch id is frozen in the graph, the last arriver leaves zero arrivals behind
    plan = pli.DecodePlan(qd, kd, vd, ld, block_tables=td, max_seq_len=L, num_splits=splits, workspace=prev_ws, out=out)
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        plan()
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            plan()
            plan()
    torch.cuda.current_stream().wait_stream(side)
    first = out.clone()
    for _ in range(3):
        out.fill_(float("nan"))
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, first)


def test_peer_output_on_the_simt_path_keeps_the_two_buffer_protocol():
    """f32 storage is served by the SIMT split-KV kernel + the combine kernel: the gather then runs the round-1 protocol
    (two buffers by step parity, publish / wait kernel behind the compute kernels) and the PeerOutput records that; mixing
    it with a single-buffer step on the same object is refused."""
    B, Hq, Hkv, D, bs, L = 4, 4, 2, 64, 16, 300
    q, kp, vp, table, lens = orc.seeded_paged(33, B, Hq, Hkv, D, bs, [L] * B, dtype=torch.float32)
    q, kp, vp, table, lens = q.cuda(), kp.cuda(), vp.cuda(), table.cuda(), lens.cuda()
    shard = pli.make_shard(0, 1, Hq, Hkv, B)
    po = pli.PeerOutput(B, Hq, D, torch.float32, shard)
    ref = pli.flash_decode(q, kp, vp, lens, block_tables=table, max_seq_len=L)[:, :, 0]
    for step in range(3):
        o = pli.flash_decode(q, kp, vp, lens, block_tables=table, max_seq_len=L, peer_out=po)
        assert torch.equal(o, ref), step
        assert o.data_ptr() == po.buffer((step + 1) & 1).data_ptr()
    assert po.mode == "scatter" and po.epoch == 3
    ro, _ = orc.paged_decode_oracle(q, kp, vp, table, lens)
    assert (ref.cpu().unsqueeze(2) - ro).abs().max().item() <= 1e-3
    pob = pli.PeerOutput(B, Hq, D, torch.bfloat16, shard)
    qb, kb, vb = q.bfloat16(), kp.bfloat16(), vp.bfloat16()
    pli.flash_decode(qb, kb, vb, lens, block_tables=table, max_seq_len=L, peer_out=pob)
    assert pob.mode == "gather"
    with pytest.raises(RuntimeError, match="protocol"):
        pob.use_mode("scatter")
