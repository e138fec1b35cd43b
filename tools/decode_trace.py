"""Decode cases VERDICT r1 #6 names, timed three ways (eager DecodePlan, a 10-step CUDA graph over rotating pools, and
-- with a tuning build -- a per-CTA %globaltimer timeline that says where a short kernel's time goes):
    python tools/decode_trace.py            (product library: timings only)
    PLI_LIB_PATH=.../build/libpli_attention_tuning.so python tools/decode_trace.py     (+ timeline)
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import physics_llm_inference_b200 as pli
from physics_llm_inference_b200 import _lib

TUNING = "tuning" in os.environ.get("PLI_LIB_PATH", "")


def case(B, Hq, Hkv, L, D=128, bs=16, splits=None, target=None):
    pages = B * L // bs
    npools = max(2, min(4, int(5e9 // (2 * pages * bs * Hkv * D * 2))))
    pools = [(torch.randn(pages, 1, bs, Hkv, D, device="cuda").bfloat16(), torch.randn(pages, 1, bs, Hkv, D, device="cuda").bfloat16())
             for _ in range(npools)]
    table = torch.randperm(pages).to(torch.int32).view(B, L // bs).cuda()
    lens = torch.full((B,), L, dtype=torch.int32, device="cuda")
    q = torch.randn(B, Hq, 1, D, device="cuda").bfloat16()
    S = splits or pli.decode_num_splits(B, Hkv, L)
    ws = pli.decode_workspace(B, Hq, D, S, "cuda")
    out = torch.empty(B, Hq, D, device="cuda", dtype=torch.bfloat16)
    plans = [pli.DecodePlan(q, kp, vp, lens, block_tables=table, max_seq_len=L, num_splits=S, workspace=ws, out=out) for kp, vp in pools]
    nbytes = 2 * B * L * Hkv * D * 2 + 2 * B * Hq * D * 2 + 4 * B * (L // bs)

    def timed(fn, reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn(); torch.cuda.synchronize()
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3

    def eager():
        for pl in plans: pl()
    us_eager = timed(eager, 10) / npools
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    steps = 12 // npools * npools
    with torch.cuda.stream(side):
        eager(); torch.cuda.synchronize()
        with torch.cuda.graph(g):
            for i in range(steps): plans[i % npools]()
    torch.cuda.current_stream().wait_stream(side)
    us_graph = timed(g.replay, 10) / steps
    tag = f"B{B} Hq{Hq} Hkv{Hkv} L{L} splits {S}"
    line = (f"{tag:34s} eager {us_eager:7.1f} us {nbytes / us_eager / 1e3:6.0f} GB/s | graph {us_graph:7.1f} us "
            f"{nbytes / us_graph / 1e3:6.0f} GB/s")
    if target:
        line += f" | target {target} GB/s: {'MET' if nbytes / us_graph / 1e3 >= target else 'not met'}"
    print(line, flush=True)
    if TUNING:
        lib = _lib.load()
        cap = 4096
        buf = torch.zeros(cap * 16, dtype=torch.int64, device="cuda")
        for _ in range(3): plans[0]()
        torch.cuda.synchronize()
        lib.pli_debug_decode_trace(buf.data_ptr(), cap)
        plans[1]()
        torch.cuda.synchronize()
        lib.pli_debug_decode_trace(None, 0)
        t = buf.view(cap, 16).cpu()
        t = t[t[:, 7] > 0].double()
        start_ns = t[:, 7] - t[:, 7].min()
        names = ["cta_start", "loads_issued", "first_landed", "last_consumed", "written", "combined"]
        print(f"    {t.shape[0]} CTAs; CTA starts spread over {start_ns.max() / 1e3:.2f} us (%globaltimer); SM cycles after the "
              f"CTA's own start (min / median / max), 1 us ~ 1900 cycles:")
        for i, n in enumerate(names[1:], start=1):
            ok = t[:, i] > 0
            if ok.sum() == 0:
                continue
            col = (t[:, i] - t[:, 0])[ok]
            print(f"      {n:14s} {col.min():9.0f} {col.median():9.0f} {col.max():9.0f}")
        for i, n in ((12, "barriers_init"), (13, "page_ids_known"), (6, "before_1st_tma"), (8, "warp0_done"), (9, "warp1_done"), (10, "warp2_done"),
                     (11, "warp3_done"), (14, "merge_bar1"), (15, "merge_bar2")):
            col = (t[:, i] - t[:, 0])[t[:, i] > 0]
            if col.numel():
                print(f"      {n:14s} {col.min():9.0f} {col.median():9.0f} {col.max():9.0f}")
        print(f"      tail (last consumed -> written) median {(t[:, 4] - t[:, 3]).median():.0f} cycles")
        ok = t[:, 5] > 0
        if ok.sum():
            print(f"      merge by the last arriver (written -> combined) median {((t[:, 5] - t[:, 4])[ok]).median():.0f} cycles")
    del pools, plans


if __name__ == "__main__":
    case(64, 32, 8, 4096, target=6250)           # C3
    case(256, 4, 1, 1024, target=4900)           # C5 ctx 1024, the share of one of 8 GPUs
    case(1, 32, 8, 32768, target=5000)
    case(8, 32, 8, 8192, target=6000)
    case(256, 4, 1, 8192)
