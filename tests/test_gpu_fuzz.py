"""Seeded random-shape parity sweep (prefill and decode) against the CPU oracle."""
import random

import pytest
import torch

import physics_llm_inference_b200 as pli
from oracle import attention_oracle as orc

pytestmark = pytest.mark.gpu


def test_prefill_random_shapes():
    rng = random.Random(1234)
    worst = 0.0
    for case in range(40):
        D = rng.choice([64, 128, 128])
        Hkv = rng.choice([1, 2, 3])
        G = rng.choice([1, 2, 4, 8])
        B = rng.choice([1, 2, 3])
        causal = rng.random() < 0.6
        Nk = rng.choice([1, 17, 64, 65, 127, 128, 129, 255, 256, 257, 383, 500, 640, 1000])
        Nq = rng.choice([1, 5, 64, 127, 128, 129, 256, 300, 512])
        if causal:
            Nq = min(Nq, Nk)
        dtype = rng.choice([torch.bfloat16, torch.float16])
        scale = rng.choice([None, None, 0.05, 0.3])
        q, k, v = orc.seeded_qkv(9000 + case, B, Hkv * G, Hkv, Nq, Nk, D, dtype=dtype)
        o, lse = pli.flash_attention_forward(q.cuda(), k.cuda(), v.cuda(), scale, causal=causal, return_lse=True)
        ro, rlse = orc.flash_attention_oracle(q, k, v, scale, causal=causal)
        eo = (o.float().cpu() - ro).abs().max().item()
        el = (lse.cpu() - rlse).abs().max().item()
        assert eo <= 2e-2 and el <= 1e-3, (case, B, Hkv, G, Nq, Nk, D, causal, dtype, scale, eo, el)
        worst = max(worst, eo)
    assert worst > 0


def test_decode_random_shapes():
    rng = random.Random(4321)
    for case in range(30):
        D = rng.choice([64, 128])
        Hkv = rng.choice([1, 2, 4])
        G = rng.choice([1, 2, 4, 5, 8, 16])
        B = rng.choice([1, 2, 5])
        bs = rng.choice([8, 16, 32, 64, 128])
        lens = [rng.choice([1, 2, 15, 16, 17, 63, 64, 65, 100, 257, 700, 1500]) for _ in range(B)]
        dtype = rng.choice([torch.bfloat16, torch.float16])
        splits = rng.choice([None, 1, 2, 5])
        n_layers = rng.choice([1, 3])
        layer = rng.randrange(n_layers)
        q, kp, vp, table, lens_t = orc.seeded_paged(8000 + case, B, Hkv * G, Hkv, D, bs, lens, num_layers=n_layers, dtype=dtype)
        o, lse = pli.flash_decode(q.cuda(), kp.cuda(), vp.cuda(), lens_t.cuda(), block_tables=table.cuda(), layer=layer,
                                  return_lse=True, num_splits=splits, max_seq_len=max(lens))
        ro, rlse = orc.paged_decode_oracle(q, kp, vp, table, lens_t, layer=layer)
        eo = (o.float().cpu() - ro).abs().max().item()
        el = (lse.cpu() - rlse[:, :, 0]).abs().max().item()
        assert eo <= 2e-2 and el <= 1e-3, (case, B, Hkv, G, D, bs, lens, dtype, splits, layer, eo, el)


def test_online_softmax_row_kernels(golden_dir):
    """a3: `online_softmax` / `online_softmax_with_output` (ch06/online_softmax.py:13-53) on the GPU against the
    reference's own outputs (ch06_online.npz, r2_ch06_surface.npz; the reference's tolerances: 1e-4 / 1e-3,
    ch06/test_ch06.py:84-120) and against the oracle on other shapes and dtypes."""
    import os

    import numpy as np
    g = np.load(os.path.join(golden_dir, "ch06_online.npz"))
    s = np.load(os.path.join(golden_dir, "r2_ch06_surface.npz"))
    x, v = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["v"]).cuda()
    sm = pli.online_softmax(x)
    assert sm.shape == x.shape and sm.dtype == x.dtype
    assert (sm.cpu() - torch.from_numpy(s["online"])).abs().max().item() <= 1e-4
    assert (sm.sum(-1) - 1).abs().max().item() <= 1e-5
    assert (pli.standard_softmax(x).cpu() - torch.from_numpy(s["standard"])).abs().max().item() <= 1e-6
    o, d = pli.online_softmax_with_output(x, v)
    assert o.shape == (3, 5, 8) and d.shape == (3, 5)
    assert (o.cpu() - torch.from_numpy(g["o"])).abs().max().item() <= 1e-3
    assert ((d.cpu() - torch.from_numpy(g["d"])) / torch.from_numpy(g["d"])).abs().max().item() <= 1e-5
    rng = torch.Generator().manual_seed(17)
    for shape, dv, dtype, tol in [((7,), 3, torch.float32, 1e-4), ((2, 3, 1), 5, torch.float32, 1e-4),
                                  ((5, 1000), 130, torch.float32, 1e-4), ((4, 33, 77), 64, torch.bfloat16, 2e-2),
                                  ((300, 17), 256, torch.float16, 2e-2)]:
        xs = (torch.randn(*shape, generator=rng) * 4).to(dtype)
        vs = torch.randn(*shape, dv, generator=rng).to(dtype)
        got = pli.online_softmax(xs.cuda())
        assert (got.float().cpu() - orc.online_softmax_oracle(xs)).abs().max().item() <= tol
        go, gd = pli.online_softmax_with_output(xs.cuda(), vs.cuda())
        ro, rd = orc.online_softmax_with_output_oracle(xs, vs)
        assert go.dtype == dtype and gd.dtype == dtype
        assert (go.float().cpu() - ro).abs().max().item() <= max(tol, 1e-3)
        assert ((gd.float().cpu() - rd) / rd).abs().max().item() <= (1e-5 if dtype == torch.float32 else 1e-2)
    # a strided view (last dim contiguous, rows not)
    big = torch.randn(6, 50, device="cuda")
    view = big[:, 10:30]
    assert (pli.online_softmax(view) - torch.softmax(view, -1)).abs().max().item() <= 1e-6
