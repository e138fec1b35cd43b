"""Bring-up check of the opt-in wide prefill kernel (flags bit 2 of pli_debug_prefill_trace; argv[1] = flags, default 4 =
three softmax warpgroups, 36 = two) against the default kernel and the oracle: python tools/wide_check.py [flags]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import physics_llm_inference_b200 as pli
from physics_llm_inference_b200 import _lib
from oracle import attention_oracle as orc

lib = _lib.load()
FLAGS = int(sys.argv[1]) if len(sys.argv) > 1 else 4
shapes = [((1, 2, 1, 128, 128, 128), True), ((1, 2, 1, 128, 128, 128), False), ((1, 4, 1, 256, 256, 128), True),
          ((1, 4, 1, 300, 300, 128), True), ((2, 8, 2, 1024, 1024, 128), True), ((2, 8, 2, 1024, 1024, 128), False),
          ((1, 8, 2, 129, 1000, 128), True), ((2, 16, 2, 384, 1000, 128), True), ((1, 8, 2, 2048, 2048, 128), True),
          ((4, 32, 8, 256, 256, 128), False), ((1, 6, 3, 513, 513, 128), True)]
bad = 0
for (B, Hq, Hkv, Nq, Nk, D), causal in shapes:
    q, k, v = orc.seeded_qkv(91, B, Hq, Hkv, Nq, Nk, D)
    k = k * torch.linspace(0.5, 4.0, Nk).view(1, 1, Nk, 1)
    qd, kd, vd = q.bfloat16().cuda(), k.bfloat16().cuda(), v.bfloat16().cuda()
    o0, l0 = pli.flash_attention_forward(qd, kd, vd, causal=causal, return_lse=True)
    torch.cuda.synchronize()
    try:
        _lib.check(lib.pli_debug_prefill_trace(None, 0, FLAGS))
        o1, l1 = pli.flash_attention_forward(qd, kd, vd, causal=causal, return_lse=True)
        torch.cuda.synchronize()
        same = True
        for _ in range(5):
            o2, l2 = pli.flash_attention_forward(qd, kd, vd, causal=causal, return_lse=True)
            torch.cuda.synchronize()
            same = same and torch.equal(o1, o2) and torch.equal(l1, l2)
    finally:
        _lib.check(lib.pli_debug_prefill_trace(None, 0, 0))
    ro, rl = orc.flash_attention_oracle(qd, kd, vd, causal=causal)
    eo = (o1.float().cpu() - ro).abs().max().item()
    el = (l1.cpu() - rl).abs().max().item()
    dd = (o1.float() - o0.float()).abs().max().item()
    ok = eo <= 2e-2 and el <= 1e-3 and same
    bad += not ok
    print(f"{'ok ' if ok else 'BAD'} B{B} {Hq}q/{Hkv}kv Nq{Nq} Nk{Nk} causal={int(causal)}: |o-oracle| {eo:.2e} |lse-oracle| {el:.2e} "
          f"|o-default| {dd:.2e} deterministic={same}", flush=True)
print("FAILED" if bad else "all ok")
sys.exit(1 if bad else 0)
