"""Short, fixed workloads for ncu captures (one GPU): `python tools/profile_target.py prefill|paged|decode|split [reps]`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import physics_llm_inference_b200 as pli

which = sys.argv[1] if len(sys.argv) > 1 else "prefill"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
torch.manual_seed(0)
if which == "prefill":
    B, Hq, Hkv, N, D = 4, 32, 8, 8192, 128
    q = torch.randn(B, Hq, N, D, device="cuda").bfloat16()
    k = torch.randn(B, Hkv, N, D, device="cuda").bfloat16()
    v = torch.randn(B, Hkv, N, D, device="cuda").bfloat16()
    for _ in range(reps):
        pli.flash_attention_forward(q, k, v, causal=True)
elif which == "paged":
    # C2 read in place from 16-token pages (chunk = whole prompt): the kPaged instance of the prefill kernel
    B, Hq, Hkv, N, D, bs = 4, 32, 8, 8192, 128, 16
    P = B * N // bs
    kp = torch.randn(P, 1, bs, Hkv, D, device="cuda").bfloat16()
    vp = torch.randn(P, 1, bs, Hkv, D, device="cuda").bfloat16()
    table = torch.randperm(P).to(torch.int32).view(B, N // bs).cuda()
    lens = torch.full((B,), N, dtype=torch.int32, device="cuda")
    q = torch.randn(B, Hq, N, D, device="cuda").bfloat16()
    for _ in range(reps):
        pli.flash_attention_paged(q, kp, vp, table, lens, max_seq_len=N)
else:
    # decode: C3, or ("split") one sequence of 32768 tokens = 37 splits per KV head merged by the kernel itself
    B, Hq, Hkv, D, L, bs = (1, 32, 8, 128, 32768, 16) if which == "split" else (64, 32, 8, 128, 4096, 16)
    P = B * L // bs
    kp = torch.randn(P, 1, bs, Hkv, D, device="cuda").bfloat16()
    vp = torch.randn(P, 1, bs, Hkv, D, device="cuda").bfloat16()
    table = torch.randperm(P)[:P].to(torch.int32).view(B, L // bs).cuda()
    lens = torch.full((B,), L, dtype=torch.int32, device="cuda")
    q = torch.randn(B, Hq, 1, D, device="cuda").bfloat16()
    for _ in range(reps):
        pli.flash_decode(q, kp, vp, lens, block_tables=table, max_seq_len=L)
torch.cuda.synchronize()
print("done", which)
