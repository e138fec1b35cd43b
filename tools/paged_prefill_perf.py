"""Paged / ragged prefill against the contiguous kernel on the C2 shape (chunk = whole prompt) and on chunked
shapes: python tools/paged_prefill_perf.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import physics_llm_inference_b200 as pli


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run(B, Hq, Hkv, Nq, L, D=128, bs=16):
    pages = B * L // bs
    kp = torch.randn(pages, 1, bs, Hkv, D, device="cuda").bfloat16()
    vp = torch.randn(pages, 1, bs, Hkv, D, device="cuda").bfloat16()
    table = torch.randperm(pages).to(torch.int32).view(B, L // bs).cuda()
    lens = torch.full((B,), L, dtype=torch.int32, device="cuda")
    q = torch.randn(B, Hq, Nq, D, device="cuda").bfloat16()
    kc = pli.paged_gather(kp, table, lens, L).transpose(1, 2)
    vc = pli.paged_gather(vp, table, lens, L).transpose(1, 2)
    fl = pli.prefill_algorithmic_flops(B, Hq, Nq, L, D, True)
    t_c = timed(lambda: pli.flash_attention_forward(q, kc, vc, causal=True))
    t_p = timed(lambda: pli.flash_attention_paged(q, kp, vp, table, lens, max_seq_len=L))
    qv = q.transpose(1, 2).reshape(B * Nq, Hq, D).contiguous()
    cu = (torch.arange(B + 1, dtype=torch.int32) * Nq).cuda()
    t_v = timed(lambda: pli.flash_attention_varlen_paged(qv, kp, vp, table, lens, cu, Nq, max_seq_len=L))
    print(f"B{B} {Hq}q/{Hkv}kv Nq{Nq} L{L} bs{bs}: contiguous {fl / t_c / 1e9:7.1f}  paged {fl / t_p / 1e9:7.1f}  "
          f"ragged+paged {fl / t_v / 1e9:7.1f} TFLOP/s   ({t_c * 1e3:.0f} / {t_p * 1e3:.0f} / {t_v * 1e3:.0f} us)", flush=True)


if __name__ == "__main__":
    run(4, 32, 8, 8192, 8192)
    run(4, 32, 8, 8192, 8192, bs=32)
    run(4, 32, 8, 8192, 8192, bs=64)
    run(4, 32, 8, 8192, 8192, bs=128)
    run(8, 32, 8, 2048, 8192)
    run(16, 32, 8, 512, 4096)
    run(64, 32, 8, 128, 4096)
