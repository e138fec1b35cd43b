"""Prefill throughput over a grid of shapes (spot pathological cases):  python tools/prefill_sweep.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import physics_llm_inference_b200 as pli


def run(B, Hq, Hkv, Nq, Nk, D, causal, dtype=torch.bfloat16, reps=10):
    q = torch.randn(B, Hq, Nq, D, device="cuda").to(dtype)
    k = torch.randn(B, Hkv, Nk, D, device="cuda").to(dtype)
    v = torch.randn(B, Hkv, Nk, D, device="cuda").to(dtype)
    for _ in range(2):
        pli.flash_attention_forward(q, k, v, causal=causal)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        pli.flash_attention_forward(q, k, v, causal=causal)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = pli.prefill_algorithmic_flops(B, Hq, Nq, Nk, D, causal)
    print(f"B{B:3d} Hq{Hq:2d} Hkv{Hkv:2d} Nq{Nq:6d} Nk{Nk:6d} D{D:3d} causal={int(causal)} {str(dtype)[6:]:8s} {ms:8.3f} ms {fl / ms / 1e9:7.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    for causal in (True, False):
        for (B, N) in ((64, 512), (32, 1024), (16, 2048), (8, 4096), (4, 8192), (2, 16384), (1, 32768)):
            run(B, 32, 8, N, N, 128, causal)
    run(4, 32, 8, 8192, 8192, 64, True)
    run(4, 32, 8, 8192, 8192, 64, False)
    run(4, 32, 8, 8192, 8192, 128, True, torch.float16)
    run(4, 32, 32, 8192, 8192, 128, True)          # MHA (no cluster pairs)
    run(4, 32, 4, 8192, 8192, 128, True)           # G = 8
    run(1, 32, 8, 8192, 8192, 128, True)
    run(8, 32, 8, 512, 8192, 128, True)            # chunked prefill over a long cache
    run(64, 32, 8, 128, 4096, 128, True)           # one Q tile per sequence (dead second tile)
    run(1, 8, 8, 512, 512, 64, False, torch.float32, reps=5)   # C1 on the SIMT path
