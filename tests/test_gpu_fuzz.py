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
