"""Host issue time of a planned decode step (plain and with the single-call gather into a PeerOutput): python tools/plan_overhead.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import physics_llm_inference_b200 as pli
B, Hq, Hkv, L, D, bs = 256, 4, 1, 1024, 128, 16
pages = B * L // bs
kp = torch.randn(pages, 1, bs, Hkv, D, device="cuda").bfloat16(); vp = torch.randn(pages, 1, bs, Hkv, D, device="cuda").bfloat16()
table = torch.randperm(pages).to(torch.int32).view(B, L // bs).cuda()
lens = torch.full((B,), L, dtype=torch.int32, device="cuda")
q = torch.randn(B, Hq, 1, D, device="cuda").bfloat16()
ws = pli.decode_workspace(B, Hq, D, 1, "cuda"); out = torch.empty(B, Hq, D, device="cuda", dtype=torch.bfloat16)
shard = pli.make_shard(0, 1, Hq, Hkv, B)
po = pli.PeerOutput(B, Hq, D, torch.bfloat16, shard)
plans = {"plain": pli.DecodePlan(q, kp, vp, lens, block_tables=table, max_seq_len=L, workspace=ws, out=out),
         "gather (world 1)": pli.DecodePlan(q, kp, vp, lens, block_tables=table, max_seq_len=L, workspace=ws, peer_out=po)}
for name, f in plans.items():
    for _ in range(20): f()
    torch.cuda.synchronize()
    n = 3000
    t0 = time.perf_counter()
    for _ in range(n): f()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{name:18s} CPU issue time per step {(t1 - t0) / n * 1e6:5.1f} us; per step incl. drain {(t2 - t0) / n * 1e6:5.1f} us", flush=True)
