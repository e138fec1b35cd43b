"""GPU parity AT SIZE for the two large BASELINE configs, through the C ABI, against the pinned oracle.

C4  long-context causal prefill, N = 65 536 (ch06/flash_attention.py:38-72 + ch02/cached_generation.py:85-91 mask):
    the per-rank shard of the 8-GPU run (B1, 4 q / 1 kv heads) and the full 32 q / 8 kv problem on one GPU.
C5  paged decode B = 256, L = 32 768, 16-token pages, 524 288 pages in use (ch02/cached_generation.py:72-94 over
    ch07/paged_memory.py:54,84-86): page * page_stride exceeds 2^31 elements here, which is exactly where a 32-bit
    offset or a TMA coordinate would overflow silently.

The oracle cannot run these sizes whole, so rows / sequences are SAMPLED (first, last, tile edges, seeded random ones)
and each sample is evaluated with the oracle's ch02 maths (`cached_attention_oracle`: query row i of a causal prefill
is a one-token decode over keys [0, i]); size-independent properties cover the rest (V = 1 => O = 1, invariance under a
physical page permutation, causality by perturbation).
"""
import pytest
import torch

import physics_llm_inference_b200 as pli
from oracle import attention_oracle as orc

pytestmark = pytest.mark.gpu

N4 = 65536


def _sample_rows(n, seed):
    edges = [0, 1, 63, 64, 127, 128, 129, 255, 256, 8191, 8192, 32767, 32768, 32769, n - 257, n - 129, n - 128, n - 2, n - 1]
    g = torch.Generator().manual_seed(seed)
    rnd = torch.randint(0, n, (8,), generator=g).tolist()
    return sorted(set(int(r) for r in edges + rnd if 0 <= r < n))


def _check_rows(o, lse, q, k, v, rows, kv_heads):
    """o, lse, q: (1, Hq, N, D) / (1, Hq, N) on the GPU; k, v (1, Hkv, N, D).  Row i of q head h == one-token decode of
    q[h, i] over keys [0, i] of kv head h // G (oracle: ch02/cached_generation.py:72-94, no mask for one token)."""
    Hq, Hkv = q.shape[1], k.shape[1]
    G = Hq // Hkv
    worst_o = worst_l = 0.0
    for hk in kv_heads:
        kc = k[0, hk].float().cpu().unsqueeze(0).unsqueeze(2)           # (1, N, 1, D) cache layout of ch02
        vc = v[0, hk].float().cpu().unsqueeze(0).unsqueeze(2)
        for i in rows:
            qi = q[:, hk * G:(hk + 1) * G, i:i + 1].float().cpu()      # (1, G, 1, D)
            ro, rl = orc.cached_attention_oracle(qi, kc, vc, i + 1)
            worst_o = max(worst_o, (o[:, hk * G:(hk + 1) * G, i:i + 1].float().cpu() - ro).abs().max().item())
            worst_l = max(worst_l, (lse[:, hk * G:(hk + 1) * G, i:i + 1].cpu() - rl).abs().max().item())
    return worst_o, worst_l


def test_c4_rank_shard_at_size():
    """The shard one of eight GPUs runs for C4: B1, 4 q heads / 1 kv head, N 65 536, causal, bf16."""
    g = torch.Generator(device="cuda").manual_seed(0xC0FFEE + 4)
    q = torch.randn(1, 4, N4, 128, device="cuda", generator=g).bfloat16()
    k = torch.randn(1, 1, N4, 128, device="cuda", generator=g).bfloat16()
    v = torch.randn(1, 1, N4, 128, device="cuda", generator=g).bfloat16()
    o, lse = pli.flash_attention_forward(q, k, v, causal=True, return_lse=True)
    torch.cuda.synchronize()
    assert pli.prefill_kernel_kind(q, k, v) == "tcgen05"
    eo, el = _check_rows(o, lse, q, k, v, _sample_rows(N4, 1), [0])
    assert eo <= 2e-2 and el <= 1e-3, (eo, el)
    # V = 1 => O = 1 for every one of the 262 144 rows (normalisation, no row lost or double-stored at this size)
    o1 = pli.flash_attention_forward(q, k, torch.ones_like(v), causal=True)
    assert (o1.float() - 1).abs().max().item() <= 1e-2
    # causality: perturbing keys >= 40 000 leaves rows < 40 000 bit-identical
    k2, v2 = k.clone(), v.clone()
    k2[:, :, 40000:] += 1
    v2[:, :, 40000:] -= 3
    o2 = pli.flash_attention_forward(q, k2, v2, causal=True)
    assert torch.equal(o2[:, :, :40000], o[:, :, :40000])
    assert not torch.equal(o2[:, :, 40000:], o[:, :, 40000:])


def test_c4_full_problem_on_one_gpu():
    """All of C4 (32 q / 8 kv heads, N 65 536, 35.2 TFLOP) on one GPU: sampled rows of four of the eight KV groups,
    and GQA sharing: a q head copied onto its group neighbour gives bit-identical rows."""
    g = torch.Generator(device="cuda").manual_seed(0xC0FFEE + 44)
    q = torch.randn(1, 32, N4, 128, device="cuda", generator=g).bfloat16()
    k = torch.randn(1, 8, N4, 128, device="cuda", generator=g).bfloat16()
    v = torch.randn(1, 8, N4, 128, device="cuda", generator=g).bfloat16()
    q[:, 5] = q[:, 4]                                      # heads 4 and 5 share kv head 1
    o, lse = pli.flash_attention_forward(q, k, v, causal=True, return_lse=True)
    torch.cuda.synchronize()
    rows = [0, 127, 128, 4095, 4096, 32768, 65407, 65408, N4 - 1]
    eo, el = _check_rows(o, lse, q, k, v, rows, [0, 1, 6, 7])
    assert eo <= 2e-2 and el <= 1e-3, (eo, el)
    assert torch.equal(o[:, 5], o[:, 4]) and torch.equal(lse[:, 5], lse[:, 4])
    assert not torch.equal(o[:, 3], o[:, 4])


def test_c5_paged_decode_at_size():
    """All of C5's largest point on one GPU: B 256, L 32 768, 32 q / 8 kv heads, 16-token pages; 524 288 pages in use
    out of 524 288 + 37, random permutation table; K and V pools are 17.2 GB each."""
    B, Hq, Hkv, D, bs, L = 256, 32, 8, 128, 16, 32768
    pages_per = L // bs
    P = B * pages_per + 37
    free, _ = torch.cuda.mem_get_info()
    need = 2 * P * bs * Hkv * D * 2
    if free < 2.2 * need:
        pytest.skip(f"needs {2.2 * need / 2**30:.0f} GiB of free HBM")
    g = torch.Generator(device="cuda").manual_seed(0xC0FFEE + 5)
    kp = torch.empty(P, 1, bs, Hkv, D, device="cuda", dtype=torch.bfloat16).normal_(generator=g)
    vp = torch.empty(P, 1, bs, Hkv, D, device="cuda", dtype=torch.bfloat16).normal_(generator=g)
    q = torch.randn(B, Hq, 1, D, device="cuda", generator=g).bfloat16()
    perm = torch.randperm(P, generator=torch.Generator().manual_seed(55))[:B * pages_per].to(torch.int32)
    table = perm.view(B, pages_per).cuda()
    assert int(table.max()) * kp.stride(0) > 2 ** 31              # the point of the test
    lens_l = [L] * B
    lens_l[7], lens_l[100] = L - 5, 17                            # ragged: a partly filled last page, a short one
    lens = torch.tensor(lens_l, dtype=torch.int32, device="cuda")
    o, lse = pli.flash_decode(q, kp, vp, lens, block_tables=table, return_lse=True, max_seq_len=L)
    torch.cuda.synchronize()
    assert pli.decode_kernel_kind(kp, table) == "mma_tma"
    for b in (0, 7, 100, 255):
        pages = table[b].long()
        kg = kp[pages, 0].reshape(1, L, Hkv, D).cpu()              # ch07 address rule: token t -> page table[t // bs], slot t % bs
        vg = vp[pages, 0].reshape(1, L, Hkv, D).cpu()
        # the same gather through the oracle's own address function on a few tokens (bit-exact indexing)
        for t in (0, 15, 16, lens_l[b] - 1):
            pg, slot = orc.page_address(t, table[b].tolist(), bs)
            assert torch.equal(kg[0, t], kp[pg, 0, slot].cpu())
        ro, rl = orc.cached_attention_oracle(q[b:b + 1].cpu(), kg, vg, lens_l[b])
        assert (o[b:b + 1].float().cpu() - ro).abs().max().item() <= 2e-2, b
        assert (lse[b:b + 1].cpu() - rl[:, :, 0]).abs().max().item() <= 1e-3, b
    # physical page permutation invariance at full size: new_pool[i] = old_pool[perm2[i]], tables remapped
    perm2 = torch.randperm(P, generator=torch.Generator().manual_seed(56)).cuda()
    inv = torch.empty_like(perm2)
    inv[perm2] = torch.arange(P, device="cuda")
    table2 = inv[table.long()].to(torch.int32)
    kp2 = kp[perm2]
    del kp
    vp2 = vp[perm2]
    del vp
    o2, lse2 = pli.flash_decode(q, kp2, vp2, lens, block_tables=table2, return_lse=True, max_seq_len=L)
    assert torch.equal(o2, o) and torch.equal(lse2, lse)
    # head-sharded the way the 8-GPU run does it (one kv head + its 4 q heads): same rows as the unsharded call
    ks, vs = kp2[:, :, :, 3:4], vp2[:, :, :, 3:4]
    os_ = pli.flash_decode(q[:, 12:16], ks, vs, lens, block_tables=table2, max_seq_len=L)
    assert (os_.float() - o2[:, 12:16].float()).abs().max().item() <= 1e-2
