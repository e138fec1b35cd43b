"""The CPU oracle against the committed outputs of the UNMODIFIED reference (tests/golden/*.npz,
written by oracle/make_golden.py in the build container).  This is what pins the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import attention_oracle as orc


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _checksum(*ts):
    return np.array([float(t.double().sum()) for t in ts] + [float((t.double() ** 2).sum()) for t in ts])


@pytest.mark.parametrize("case", ["small", "ragged", "c1"])
def test_ch06_flash_bit_equal(golden_dir, case):
    g = _load(golden_dir, "ch06_flash.npz")
    seed, B, H, N, D, stride = [int(x) for x in g[f"{case}_meta"]]
    q, k, v = orc.seeded_qkv(seed, B, H, H, N, N, D)
    np.testing.assert_allclose(_checksum(q, k, v), g[f"{case}_insum"], rtol=1e-12,
                               err_msg="seeded inputs differ from the ones the golden vectors were made with")
    o, lse = orc.flash_attention_oracle(q, k, v)
    ref = torch.from_numpy(g[f"{case}_flash"])
    # the restatement was bit-equal to the reference when the vectors were written; allow 1 ulp-level
    # drift for a different BLAS/thread count on another host
    assert (o[:, :, ::stride] - ref).abs().max().item() <= 2e-6
    nav = torch.from_numpy(g[f"{case}_naive"])
    assert (o[:, :, ::stride] - nav).abs().max().item() <= 5e-6
    o2, lse2 = orc.naive_attention_oracle(q, k, v)
    assert (lse - lse2).abs().max().item() <= 5e-6


def test_ch06_odd_tiles_and_scale(golden_dir):
    g = _load(golden_dir, "ch06_flash.npz")
    seed, B, H, N, D, _ = [int(x) for x in g["oddtile_meta"]]
    q, k, v = orc.seeded_qkv(seed, B, H, H, N, N, D)
    o, _ = orc.flash_attention_oracle(q, k, v, scale=0.2, block_q=48, block_k=80)
    assert (o - torch.from_numpy(g["oddtile_flash"])).abs().max().item() <= 2e-6
    o64, _ = orc.flash_attention_oracle(q, k, v, scale=0.2)
    assert (o - o64).abs().max().item() <= 5e-6   # tile hints do not change the function


def test_ch06_online_softmax_recurrence(golden_dir):
    g = _load(golden_dir, "ch06_online.npz")
    x, v = torch.from_numpy(g["x"]), torch.from_numpy(g["v"])
    # scores x as a 1-dim "attention": q = 1, k = x, scale = 1, one key per block = the scalar recurrence
    lead = x.shape[:-1]
    n = x.shape[-1]
    q = torch.ones(int(np.prod(lead)), 1, 1, 1)
    k = x.reshape(-1, 1, n, 1)
    vv = v.reshape(-1, 1, n, v.shape[-1])
    o, lse = orc.flash_attention_oracle(q, k, vv, scale=1.0, block_q=1, block_k=1)
    assert (o.reshape(*lead, -1) - torch.from_numpy(g["o"])).abs().max().item() <= 2e-6
    # d of the reference is the running sum relative to the running max: lse = max + log d
    d_ref = torch.from_numpy(g["d"])
    assert (lse.reshape(lead) - (x.max(-1).values + torch.log(d_ref))).abs().max().item() <= 5e-6


def _split_x(x, Hq, Hkv, D):
    B, N, _ = x.shape
    q = x.view(B, N, Hq, D).transpose(1, 2)
    k = x[..., :Hkv * D].reshape(B, N, Hkv, D).transpose(1, 2)
    v = x[..., Hkv * D:2 * Hkv * D].reshape(B, N, Hkv, D).transpose(1, 2)
    return q, k, v


def test_ch01_gqa_causal(golden_dir):
    g = _load(golden_dir, "ch01_gqa.npz")
    B, Hq, Hkv, N, D = [int(x) for x in g["meta"]]
    q, k, v = _split_x(torch.from_numpy(g["x"]), Hq, Hkv, D)
    o, lse = orc.flash_attention_oracle(q, k, v, causal=True)
    assert (o - torch.from_numpy(g["causal"])).abs().max().item() <= 2e-6
    o, _ = orc.flash_attention_oracle(q, k, v, causal=False)
    assert (o - torch.from_numpy(g["full"])).abs().max().item() <= 2e-6
    # causality by perturbation, as ch01/test_ch01.py:22-39
    k2, v2 = k.clone(), v.clone()
    k2[:, :, -1] += 10.0
    v2[:, :, -1] -= 5.0
    o2, _ = orc.flash_attention_oracle(q, k2, v2, causal=True)
    o1, _ = orc.flash_attention_oracle(q, k, v, causal=True)
    assert torch.equal(o1[:, :, :-1], o2[:, :, :-1])


def test_ch02_cached_prefill_chunk_decode(golden_dir):
    g = _load(golden_dir, "ch02_cached.npz")
    B, Hq, Hkv, D, Lmax = [int(x) for x in g["meta"]]
    k_cache = torch.zeros(B, Lmax, Hkv, D)
    v_cache = torch.zeros(B, Lmax, Hkv, D)
    pos = 0
    for step in range(4):
        x = torch.from_numpy(g[f"x{step}"])
        s = x.shape[1]
        q, k_new, v_new = _split_x(x, Hq, Hkv, D)
        # ch02/kv_cache.py:45-46 append
        k_cache[:, pos:pos + s] = k_new.transpose(1, 2)
        v_cache[:, pos:pos + s] = v_new.transpose(1, 2)
        pos += s
        o, _ = orc.cached_attention_oracle(q, k_cache, v_cache, pos)
        assert (o - torch.from_numpy(g[f"y{step}"])).abs().max().item() <= 2e-6, step
        if s > 1:
            o3, _ = orc.flash_attention_oracle(q, k_cache[:, :pos].transpose(1, 2), v_cache[:, :pos].transpose(1, 2),
                                               causal=True, block_q=16, block_k=16)
            assert (o3 - o).abs().max().item() <= 2e-6
    assert pos == int(g["seq_len"][0])
    assert torch.equal(k_cache, torch.from_numpy(g["k_cache"]))
    assert torch.equal(v_cache, torch.from_numpy(g["v_cache"]))


def test_paged_decode_composition(golden_dir):
    g = _load(golden_dir, "paged_decode.npz")
    seed, B, Hq, Hkv, D, bs, n_layers, layer = [int(x) for x in g["meta"]]
    lens = [int(x) for x in g["lens"]]
    q, kp, vp, table, lens_t = orc.seeded_paged(seed, B, Hq, Hkv, D, bs, lens, num_layers=n_layers)
    np.testing.assert_allclose(_checksum(q, kp, vp), g["insum"], rtol=1e-12)
    assert torch.equal(table, torch.from_numpy(g["table"]))
    o, lse = orc.paged_decode_oracle(q, kp, vp, table, lens_t, layer=layer)
    assert (o - torch.from_numpy(g["o"])).abs().max().item() <= 2e-6
    assert (lse - torch.from_numpy(g["lse"])).abs().max().item() <= 2e-6
    # address rule, bit exact: token t -> page table[t // bs], slot t % bs
    for b in range(B):
        kg = orc.gather_paged(kp, table[b].tolist(), lens[b], layer)
        for t in range(lens[b]):
            page, slot = orc.page_address(t, table[b].tolist(), bs)
            assert torch.equal(kg[t], kp[page, layer, slot])


def test_split_combine_is_exact():
    q, k, v = orc.seeded_qkv(7, 2, 4, 2, 1, 300, 32)
    full, lse = orc.naive_attention_oracle(q, k, v)
    parts, lses = [], []
    for a, b in [(0, 128), (128, 192), (192, 300)]:
        o, l = orc.naive_attention_oracle(q, k[:, :, a:b], v[:, :, a:b])
        parts.append(o)
        lses.append(l)
    o, l = orc.combine_splits_oracle(torch.stack(parts), torch.stack(lses))
    assert (o - full).abs().max().item() <= 2e-6
    assert (l - lse).abs().max().item() <= 2e-6


def test_ch06_online_softmax_functions_bit_equal(golden_dir):
    """oracle restatements of ch06/online_softmax.py:13-53 == the reference's outputs (r2_ch06_surface.npz / ch06_online.npz)."""
    g = _load(golden_dir, "ch06_online.npz")
    s = _load(golden_dir, "r2_ch06_surface.npz")
    x, v = torch.from_numpy(g["x"]), torch.from_numpy(g["v"])
    assert torch.equal(orc.online_softmax_oracle(x), torch.from_numpy(s["online"]))
    o, d = orc.online_softmax_with_output_oracle(x, v)
    assert torch.equal(o, torch.from_numpy(g["o"])) and torch.equal(d, torch.from_numpy(g["d"]))
    # the reference's own property (ch06/test_ch06.py:84-88): online == standard to 1e-4
    assert (torch.from_numpy(s["online"]) - torch.from_numpy(s["standard"])).abs().max().item() <= 1e-4
