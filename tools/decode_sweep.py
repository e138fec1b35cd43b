"""Per-GPU share of BASELINE C5 (decode B256, ctx 1k-32k, paged, KV heads sharded 1 per GPU => Hq 4 / Hkv 1),
plus C3, as GB/s of algorithmic bytes.  python tools/decode_sweep.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import physics_llm_inference_b200 as pli


def run(B, Hq, Hkv, L, D=128, bs=16, reps=20, splits=None):
    pages = B * L // bs
    npools = max(1, min(3, int(6e9 // (2 * pages * bs * Hkv * D * 2))))
    pools = [(torch.randn(pages, 1, bs, Hkv, D, device="cuda").bfloat16(), torch.randn(pages, 1, bs, Hkv, D, device="cuda").bfloat16())
             for _ in range(npools)]
    table = torch.randperm(pages)[:pages].to(torch.int32).view(B, L // bs).cuda()
    lens = torch.full((B,), L, dtype=torch.int32, device="cuda")
    q = torch.randn(B, Hq, 1, D, device="cuda").bfloat16()
    S = splits or pli.decode_num_splits(B, Hkv, L)
    ws = pli.decode_workspace(B, Hq, D, S, "cuda")
    out = torch.empty(B, Hq, D, device="cuda", dtype=torch.bfloat16)
    for i in range(3):
        pli.flash_decode(q, *pools[i % npools], lens, block_tables=table, max_seq_len=L, num_splits=S, workspace=ws, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        pli.flash_decode(q, *pools[i % npools], lens, block_tables=table, max_seq_len=L, num_splits=S, workspace=ws, out=out)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    nbytes = 2 * B * L * Hkv * D * 2 + 2 * B * Hq * D * 2 + 4 * B * (L // bs)
    print(f"B{B} Hq{Hq} Hkv{Hkv} L{L}: splits {S}  {us:8.1f} us  {nbytes / us / 1e3:7.0f} GB/s  ({nbytes / 1e9:.3f} GB, {npools} pools)", flush=True)
    del pools


if __name__ == "__main__":
    run(64, 32, 8, 4096)
    for L in (1024, 2048, 4096, 8192, 16384, 32768):
        run(256, 4, 1, L)
    for S in (1, 2, 3, 4):
        run(256, 4, 1, 1024, splits=S)
    run(1, 32, 8, 32768)
    run(8, 32, 8, 8192)
