"""Staged GPU bring-up report: runs each layer of the hot path in order, never stops at the first
failure, and prints enough to localise a wrong descriptor / barrier from one gpurun call.

    python tools/gpu_bringup.py [--quick]
"""
import os
import subprocess
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STAGES = {}


def stage(name):
    def deco(fn):
        STAGES[name] = fn
        return fn
    return deco


@stage("selftest")
def selftest():
    import torch
    from physics_llm_inference_b200 import _lib
    lib = _lib.load()
    for D in (128, 64):
        for dtype in (torch.bfloat16, torch.float16):
            g = torch.Generator().manual_seed(5)
            a = torch.randn(128, D, generator=g).to(dtype).cuda()
            b = torch.randn(128, D, generator=g).to(dtype).cuda()
            c = torch.randn(128, D, generator=g).to(dtype).cuda()
            s_out = torch.zeros(128, 128, device="cuda")
            o_out = torch.zeros(128, D, device="cuda")
            _lib.check(lib.pli_set_device(0))
            _lib.check(lib.pli_debug_umma_selftest(a.data_ptr(), b.data_ptr(), c.data_ptr(), s_out.data_ptr(),
                                                   o_out.data_ptr(), D, _lib.dtype_code(dtype),
                                                   torch.cuda.current_stream().cuda_stream))
            torch.cuda.synchronize()
            s_ref = a.float() @ b.float().T
            es = (s_out - s_ref).abs()
            p = (s_out * 0.015625).to(dtype).float()
            o_ref = p @ c.float()
            o_ref2 = (s_ref * 0.015625).to(dtype).float() @ c.float()
            eo = (o_out - o_ref).abs()
            print(f"  D={D} {dtype}: S err max {es.max():.3e} (row-wise max first 4 rows {es.max(1).values[:4].tolist()}); "
                  f"O err max {eo.max():.3e} (vs ref-S {(o_out - o_ref2).abs().max():.3e})")
            if es.max() > 1e-3:
                bad = (es > 1e-3)
                print(f"    S wrong in {int(bad.sum())} of {bad.numel()} entries; bad rows {bad.any(1).nonzero().flatten()[:16].tolist()} "
                      f"bad cols {bad.any(0).nonzero().flatten()[:16].tolist()}")
                print("    S[0,:8]", s_out[0, :8].tolist(), "ref", s_ref[0, :8].tolist())
            if eo.max() > 2e-3:
                bad = (eo > 2e-3)
                print(f"    O wrong in {int(bad.sum())} of {bad.numel()} entries; bad rows {bad.any(1).nonzero().flatten()[:16].tolist()} "
                      f"bad cols {bad.any(0).nonzero().flatten()[:16].tolist()}")
                print("    O[0,:8]", o_out[0, :8].tolist(), "ref", o_ref[0, :8].tolist())


def _prefill_case(B, Hq, Hkv, Nq, Nk, D, dtype, causal, seed=3):
    import torch
    import physics_llm_inference_b200 as pli
    from oracle import attention_oracle as orc
    q, k, v = orc.seeded_qkv(seed, B, Hq, Hkv, Nq, Nk, D, dtype=dtype)
    qd, kd, vd = q.cuda(), k.cuda(), v.cuda()
    kind = pli.prefill_kernel_kind(qd, kd, vd)
    t = time.time()
    o, lse = pli.flash_attention_forward(qd, kd, vd, causal=causal, return_lse=True)
    torch.cuda.synchronize()
    dt = time.time() - t
    ro, rlse = orc.flash_attention_oracle(q, k, v, causal=causal)
    eo = (o.float().cpu() - ro).abs()
    el = (lse.cpu() - rlse).abs()
    nan = int(torch.isnan(o.float()).sum())
    print(f"  [{kind}] B{B} Hq{Hq} Hkv{Hkv} Nq{Nq} Nk{Nk} D{D} {str(dtype)[6:]} causal={int(causal)}: "
          f"max|dO| {eo.max():.3e} max|dLSE| {el.max():.3e} nan {nan} ({dt * 1e3:.1f} ms)")
    if eo.max() > 2e-2 or nan:
        per_row = eo.amax(dim=(0, 1, 3))
        bad_rows = (per_row > 2e-2).nonzero().flatten()
        print(f"    bad rows: {bad_rows[:12].tolist()} ... {bad_rows[-4:].tolist()} ({len(bad_rows)} of {Nq}); "
              f"per-head max {eo.amax(dim=(0, 2, 3)).tolist()}")
    return eo.max().item()


@stage("simt")
def simt():
    import torch
    _prefill_case(1, 8, 8, 512, 512, 64, torch.float32, False)
    _prefill_case(1, 4, 2, 100, 130, 32, torch.float32, True)


@stage("prefill_small")
def prefill_small():
    import torch
    for (B, Hq, Hkv, Nq, Nk, D, causal) in [
        (1, 1, 1, 128, 128, 128, False), (1, 1, 1, 256, 256, 128, False), (1, 1, 1, 256, 256, 128, True),
        (1, 1, 1, 256, 512, 128, False), (1, 4, 1, 512, 512, 128, True), (1, 2, 2, 300, 300, 64, True),
        (2, 8, 2, 1024, 1024, 128, True), (1, 2, 1, 77, 333, 128, True),
    ]:
        _prefill_case(B, Hq, Hkv, Nq, Nk, D, torch.bfloat16, causal)
    _prefill_case(1, 8, 8, 512, 512, 64, torch.float16, False)


@stage("prefill_persistent")
def prefill_persistent():
    import torch
    # more work items than SMs: exercises the persistent loop, barrier phases across items, LPT order
    _prefill_case(2, 16, 4, 2048, 2048, 128, torch.bfloat16, True)
    _prefill_case(1, 32, 8, 1536, 1536, 64, torch.bfloat16, False)


@stage("decode")
def decode():
    import torch
    import physics_llm_inference_b200 as pli
    from oracle import attention_oracle as orc
    for (bs, D, G, lens, dtype, splits) in [
        (16, 32, 4, [77, 16, 1], torch.float32, None),
        (16, 128, 4, [64], torch.bfloat16, 1), (16, 128, 4, [1000, 999], torch.bfloat16, 1),
        (16, 128, 4, [4096, 100, 1, 17, 2048], torch.bfloat16, None), (16, 64, 8, [513, 64, 65], torch.float16, None),
        (8, 128, 1, [300, 9], torch.bfloat16, 2), (128, 64, 16, [900, 128, 129], torch.bfloat16, 3),
    ]:
        q, kp, vp, table, lens_t = orc.seeded_paged(32, len(lens), 2 * G, 2, D, bs, lens, num_layers=2, dtype=dtype)
        kind = pli.decode_kernel_kind(kp.cuda(), table.cuda())
        o, lse = pli.flash_decode(q.cuda(), kp.cuda(), vp.cuda(), lens_t.cuda(), block_tables=table.cuda(), layer=1,
                                  return_lse=True, num_splits=splits, max_seq_len=max(lens))
        torch.cuda.synchronize()
        ro, rlse = orc.paged_decode_oracle(q, kp, vp, table, lens_t, layer=1)
        eo = (o.float().cpu() - ro).abs()
        el = (lse.cpu() - rlse[:, :, 0]).abs()
        print(f"  [{kind}] bs{bs} D{D} G{G} lens{lens} {str(dtype)[6:]} splits={splits}: max|dO| {eo.max():.3e} "
              f"max|dLSE| {el.max():.3e} nan {int(torch.isnan(o.float()).sum())}")
        if eo.max() > 2e-2:
            print("    per-seq max", eo.amax(dim=(1, 2, 3)).tolist(), "per-head max", eo.amax(dim=(0, 2, 3)).tolist())


@stage("perf")
def perf():
    import torch
    import physics_llm_inference_b200 as pli
    B, Hq, Hkv, N, D = 4, 32, 8, 8192, 128
    q = torch.randn(B, Hq, N, D, device="cuda").bfloat16()
    k = torch.randn(B, Hkv, N, D, device="cuda").bfloat16()
    v = torch.randn(B, Hkv, N, D, device="cuda").bfloat16()
    for causal in (True, False):
        for _ in range(3):
            pli.flash_attention_forward(q, k, v, causal=causal)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            pli.flash_attention_forward(q, k, v, causal=causal)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = pli.prefill_algorithmic_flops(B, Hq, N, N, D, causal)
        print(f"  prefill C2 causal={int(causal)}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s")
    Bd, L, bs = 64, 4096, 16
    P = Bd * L // bs
    pools = [(torch.randn(P, 1, bs, Hkv, D, device="cuda").bfloat16(), torch.randn(P, 1, bs, Hkv, D, device="cuda").bfloat16())
             for _ in range(3)]
    table = torch.randperm(P)[:Bd * (L // bs)].to(torch.int32).view(Bd, L // bs).cuda()
    lens = torch.full((Bd,), L, dtype=torch.int32, device="cuda")
    qd = torch.randn(Bd, Hq, 1, D, device="cuda").bfloat16()
    nbytes = 2 * Bd * L * Hkv * D * 2 + 2 * Bd * Hq * D * 2 + 4 * Bd * (L // bs)
    for splits in (None, 1, 2, 4, 8):
        ws = pli.decode_workspace(Bd, Hq, D, splits or pli.decode_num_splits(Bd, Hkv, L), "cuda")
        for i in range(3):
            pli.flash_decode(qd, pools[i][0], pools[i][1], lens, block_tables=table, max_seq_len=L, num_splits=splits, workspace=ws)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(30):
            pli.flash_decode(qd, pools[i % 3][0], pools[i % 3][1], lens, block_tables=table, max_seq_len=L, num_splits=splits, workspace=ws)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 30 * 1e3
        print(f"  decode C3 splits={splits or pli.decode_num_splits(Bd, Hkv, L)}: {us:.1f} us  {nbytes / us / 1e3:.0f} GB/s")


def main():
    names = sys.argv[1:] or list(STAGES)
    if names == ["--child"]:
        return
    if "--stage" in sys.argv:
        name = sys.argv[sys.argv.index("--stage") + 1]
        STAGES[name]()
        return
    # each stage in its own process with a timeout: a trap or hang in one stage does not hide the others
    for name in names:
        print(f"=== {name} ===", flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--stage", name], timeout=240,
                               capture_output=True, text=True)
            print(r.stdout, end="")
            if r.returncode != 0:
                print(f"  !! stage exited with {r.returncode}\n" + "\n".join(r.stderr.strip().splitlines()[-12:]))
        except subprocess.TimeoutExpired as e:
            print((e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or ""))
            print("  !! stage timed out after 240 s")
        sys.stdout.flush()


if __name__ == "__main__":
    main()
