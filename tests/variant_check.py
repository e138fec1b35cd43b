"""Child process of tests/test_gpu_prefill.py: checks of the prefill kernel variants that only exist in the TUNING build of
the library (PLI_LIB_PATH points at build/libpli_attention_tuning.so).   python tests/variant_check.py pair|wide"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import physics_llm_inference_b200 as pli  # noqa: E402
from oracle import attention_oracle as orc  # noqa: E402
from physics_llm_inference_b200 import _lib  # noqa: E402

lib = _lib.load()
failures = []


def inputs(shape, dtype):
    B, Hq, Hkv, Nq, Nk, D = shape
    q, k, v = orc.seeded_qkv(91, B, Hq, Hkv, Nq, Nk, D)
    k = k * torch.linspace(0.5, 4.0, Nk).view(1, 1, Nk, 1)          # lazy rescales fire at many steps
    return q.to(dtype).cuda(), k.to(dtype).cuda(), v.to(dtype).cuda()


def check_pair():
    shapes = [((2, 8, 2, 1024, 1024, 128), True),      # C2 scaled down: two CTA pairs per KV group and row block
              ((1, 4, 1, 300, 300, 128), True),        # ragged last tile, a single pair
              ((2, 16, 2, 384, 1000, 128), True),      # chunk over cache (offset mask), group size 8
              ((4, 32, 8, 256, 256, 128), False),      # more pairs than the chip has SM pairs
              ((1, 8, 2, 2048, 2048, 128), True),      # lazy rescales across many steps
              ((1, 8, 2, 100, 100, 128), True)]        # a single short item per pair
    for dtype in (torch.bfloat16, torch.float16):
        for shape, causal in shapes:
            qd, kd, vd = inputs(shape, dtype)
            _lib.check(lib.pli_debug_prefill_trace(None, 0, 64))         # flags bit 6: per-CTA MMAs + TMA multicast
            o0, l0 = pli.flash_attention_forward(qd, kd, vd, causal=causal, return_lse=True)
            torch.cuda.synchronize()
            _lib.check(lib.pli_debug_prefill_trace(None, 0, 0))          # default: CTA-pair MMAs
            for _ in range(5):
                o1, l1 = pli.flash_attention_forward(qd, kd, vd, causal=causal, return_lse=True)
                torch.cuda.synchronize()
                if not (torch.equal(o1, o0) and torch.equal(l1, l0)):
                    failures.append(f"pair {shape} {causal} {dtype}: not bit-identical to the per-CTA MMA kernel")
                    break
            ro, rl = orc.flash_attention_oracle(qd, kd, vd, causal=causal)
            eo, el = (o1.float().cpu() - ro).abs().max().item(), (l1.cpu() - rl).abs().max().item()
            if eo > 2e-2 or el > 1e-3:
                failures.append(f"pair {shape} {causal} {dtype}: oracle error O {eo} LSE {el}")


def check_wide():
    shapes = [((1, 2, 1, 128, 128, 128), True), ((1, 4, 1, 300, 300, 128), True), ((2, 8, 2, 1024, 1024, 128), True),
              ((2, 8, 2, 1024, 1024, 128), False), ((1, 8, 2, 129, 1000, 128), True), ((1, 6, 3, 513, 513, 128), True),
              ((4, 32, 8, 256, 256, 128), False), ((1, 8, 2, 2048, 2048, 128), True)]
    for flags in (4, 36):                              # 4: three softmax warpgroups (576 threads), 36: two (512 threads)
        for shape, causal in shapes:
            qd, kd, vd = inputs(shape, torch.bfloat16)
            o0, l0 = pli.flash_attention_forward(qd, kd, vd, causal=causal, return_lse=True)
            torch.cuda.synchronize()
            try:
                _lib.check(lib.pli_debug_prefill_trace(None, 0, flags))
                o1, l1 = pli.flash_attention_forward(qd, kd, vd, causal=causal, return_lse=True)
                torch.cuda.synchronize()
                for _ in range(8):
                    o2, l2 = pli.flash_attention_forward(qd, kd, vd, causal=causal, return_lse=True)
                    torch.cuda.synchronize()
                    if not (torch.equal(o1, o2) and torch.equal(l1, l2)):
                        failures.append(f"wide {flags} {shape} {causal}: not reproducible")
                        break
            finally:
                _lib.check(lib.pli_debug_prefill_trace(None, 0, 0))
            ro, rl = orc.flash_attention_oracle(qd, kd, vd, causal=causal)
            eo, el = (o1.float().cpu() - ro).abs().max().item(), (l1.cpu() - rl).abs().max().item()
            if eo > 2e-2 or el > 1e-3:
                failures.append(f"wide {flags} {shape} {causal}: oracle error O {eo} LSE {el}")
            if (o1.float() - o0.float()).abs().max().item() > 4e-2:      # two bf16 roundings apart at most
                failures.append(f"wide {flags} {shape} {causal}: far from the product kernel")


if __name__ == "__main__":
    mode = sys.argv[1]
    {"pair": check_pair, "wide": check_wide}[mode]()
    for f in failures:
        print("FAIL", f)
    print(f"variant_check {mode}: {len(failures)} failure(s)")
    sys.exit(1 if failures else 0)
