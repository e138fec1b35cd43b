"""GPU parity tests of the prefill path, through the C ABI, against the CPU oracle.

Tolerances are BASELINE.json's: max-abs-error 2e-2 for bf16/f16 inputs, 1e-3 for f32 inputs, 1e-3
for log-sum-exp, all against the oracle run in fp32 on the same (already rounded) inputs."""
import os

import numpy as np
import pytest
import torch

import physics_llm_inference_b200 as pli
from oracle import attention_oracle as orc
from physics_llm_inference_b200 import _lib

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1e-3, torch.bfloat16: 2e-2, torch.float16: 2e-2}
LSE_TOL = 1e-3


def _run(q, k, v, causal, scale=None, dtype=torch.bfloat16):
    qd, kd, vd = (x.to(dtype).cuda() for x in (q, k, v))
    o, lse = pli.flash_attention_forward(qd, kd, vd, scale, causal=causal, return_lse=True)
    torch.cuda.synchronize()
    ro, rlse = orc.flash_attention_oracle(qd, kd, vd, scale, causal=causal)   # oracle sees the rounded inputs
    assert o.shape == qd.shape and o.dtype == dtype and lse.shape == qd.shape[:3]
    return (o.float().cpu() - ro).abs().max().item(), (lse.cpu() - rlse).abs().max().item()


def test_umma_selftest_descriptors():
    """One 128x128xD tile through the SS (K-major) and TS (MN-major B) MMA paths of the kernel."""
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
            assert (s_out - s_ref).abs().max().item() < 1e-3, f"SS K-major MMA wrong (D={D}, {dtype})"
            p = (s_out * 0.015625).to(dtype).float()
            o_ref = p @ c.float()
            assert (o_out - o_ref).abs().max().item() < 2e-3, f"TS MN-major MMA wrong (D={D}, {dtype})"


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("shape", [(1, 8, 8, 512, 512, 64), (2, 4, 2, 100, 100, 32), (1, 2, 1, 33, 77, 16),
                                   (1, 3, 3, 70, 70, 128), (1, 2, 2, 40, 40, 80)])
def test_fp32_simt_parity(shape, causal):
    B, Hq, Hkv, Nq, Nk, D = shape
    q, k, v = orc.seeded_qkv(0xC0FFEE + 1, B, Hq, Hkv, Nq, Nk, D)
    assert pli.prefill_kernel_kind(q.cuda(), k.cuda(), v.cuda()) == "simt"
    eo, el = _run(q, k, v, causal, dtype=torch.float32)
    assert eo <= TOL[torch.float32] and el <= LSE_TOL, (eo, el)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("shape", [
    (1, 8, 8, 512, 512, 64),        # C1 shape (book test shape) in 16-bit
    (2, 4, 4, 128, 128, 64),        # ch06/test_ch06.py:160-178
    (1, 4, 1, 256, 256, 128),       # one work item, two full Q tiles, GQA 4:1
    (2, 8, 2, 1024, 1024, 128),     # C2 scaled down
    (1, 4, 2, 300, 300, 128),       # ragged tiles
    (1, 2, 2, 1, 1, 128),           # single token
    (1, 4, 2, 77, 333, 64),         # Nq < Nk: bottom-right aligned mask (chunk over cache)
    (1, 8, 2, 129, 1000, 128),
    (3, 6, 3, 513, 513, 128),
])
def test_tcgen05_parity(shape, causal, dtype):
    B, Hq, Hkv, Nq, Nk, D = shape
    q, k, v = orc.seeded_qkv(0xC0FFEE + 2, B, Hq, Hkv, Nq, Nk, D)
    assert pli.prefill_kernel_kind(q.to(dtype).cuda(), k.to(dtype).cuda(), v.to(dtype).cuda()) == "tcgen05"
    eo, el = _run(q, k, v, causal, dtype=dtype)
    assert eo <= TOL[dtype] and el <= LSE_TOL, (eo, el)


def test_reference_cuda_tests_pass_against_the_drop_in():
    """ch06/test_ch06.py:160-189 verbatim (fp16, rtol=atol 0.01 / 0.02 vs naive attention)."""
    torch.manual_seed(0)
    for (B, H, N, D), tol in [((2, 4, 128, 64), 0.01), ((1, 8, 512, 64), 0.02), ((2, 8, 256, 64), 0.01)]:
        q = torch.randn(B, H, N, D, device="cuda", dtype=torch.float16)
        k = torch.randn(B, H, N, D, device="cuda", dtype=torch.float16)
        v = torch.randn(B, H, N, D, device="cuda", dtype=torch.float16)
        out = pli.flash_attention_forward(q, k, v)
        assert out.shape == (B, H, N, D)
        naive, _ = orc.naive_attention_oracle(q, k, v)
        torch.testing.assert_close(out.float().cpu(), naive, rtol=tol, atol=tol)


def test_golden_c1_fp32(golden_dir):
    """BASELINE config 1 against the stored output of the unmodified reference."""
    g = np.load(os.path.join(golden_dir, "ch06_flash.npz"))
    seed, B, H, N, D, stride = [int(x) for x in g["c1_meta"]]
    q, k, v = orc.seeded_qkv(seed, B, H, H, N, N, D)
    o = pli.flash_attention_forward(q.cuda(), k.cuda(), v.cuda())
    ref = torch.from_numpy(g["c1_flash"])
    assert (o.cpu()[:, :, ::stride] - ref).abs().max().item() <= 1e-3
    for dtype in (torch.bfloat16, torch.float16):       # same inputs rounded: 2e-2 against the fp32 reference run
        o = pli.flash_attention_forward(q.to(dtype).cuda(), k.to(dtype).cuda(), v.to(dtype).cuda())
        assert (o.float().cpu()[:, :, ::stride] - ref).abs().max().item() <= 2e-2


def test_golden_ch01_gqa_and_ch02_chunks(golden_dir):
    g = np.load(os.path.join(golden_dir, "ch01_gqa.npz"))
    B, Hq, Hkv, N, D = [int(x) for x in g["meta"]]
    x = torch.from_numpy(g["x"]).cuda()
    # the (B,N,H,D)->(B,H,N,D) transposed views of ch01/gqa.py:27-29 go in without a copy
    q = x.view(B, N, Hq, D).transpose(1, 2)
    k = x[..., :Hkv * D].reshape(B, N, Hkv, D).transpose(1, 2)
    v = x[..., Hkv * D:2 * Hkv * D].reshape(B, N, Hkv, D).transpose(1, 2)
    o = pli.flash_attention_forward(q, k, v, causal=True)
    assert (o.cpu() - torch.from_numpy(g["causal"])).abs().max().item() <= 1e-3
    o = pli.flash_attention_forward(q, k, v)
    assert (o.cpu() - torch.from_numpy(g["full"])).abs().max().item() <= 1e-3


def test_strided_views_bf16():
    """(B,N,H,D) storage viewed as (B,H,N,D): TMA descriptors are built from the strides."""
    B, N, Hq, Hkv, D = 2, 384, 8, 2, 128
    g = torch.Generator().manual_seed(9)
    qs = torch.randn(B, N, Hq, D, generator=g).bfloat16().cuda()
    ks = torch.randn(B, N, Hkv, D, generator=g).bfloat16().cuda()
    vs = torch.randn(B, N, Hkv, D, generator=g).bfloat16().cuda()
    q, k, v = qs.transpose(1, 2), ks.transpose(1, 2), vs.transpose(1, 2)
    assert pli.prefill_kernel_kind(q, k, v) == "tcgen05"
    o = pli.flash_attention_forward(q, k, v, causal=True)
    assert o.stride() == q.stride()
    ro, _ = orc.flash_attention_oracle(q, k, v, causal=True)
    assert (o.float().cpu() - ro).abs().max().item() <= 2e-2


def test_scale_and_config_arguments():
    q, k, v = orc.seeded_qkv(21, 1, 4, 4, 200, 200, 64)
    for dtype in (torch.float32, torch.bfloat16):
        qd, kd, vd = (x.to(dtype).cuda() for x in (q, k, v))
        for scale in (0.2, 1.0, -0.1):
            o = pli.flash_attention_forward(qd, kd, vd, scale, pli.FlashAttentionConfig(block_q=48, block_k=80))
            ro, _ = orc.flash_attention_oracle(qd, kd, vd, scale)
            assert (o.float().cpu() - ro).abs().max().item() <= (3e-2 if scale == 1.0 else TOL[dtype])


def test_normalisation_and_causality_properties_full_size():
    """Size-independent properties at BASELINE's C2 shape (B4, 32q/8kv, N8192, D128, bf16, causal)."""
    B, Hq, Hkv, N, D = 4, 32, 8, 8192, 128
    g = torch.Generator(device="cuda").manual_seed(3)
    q = torch.randn(B, Hq, N, D, device="cuda", generator=g).bfloat16()
    k = torch.randn(B, Hkv, N, D, device="cuda", generator=g).bfloat16()
    v = torch.randn(B, Hkv, N, D, device="cuda", generator=g).bfloat16()
    o, lse = pli.flash_attention_forward(q, k, v, causal=True, return_lse=True)
    # (1) V = 1 => O = 1 exactly up to bf16 rounding (ch06/test_ch06.py:67-73)
    ones = torch.ones_like(v)
    o1 = pli.flash_attention_forward(q, k, ones, causal=True)
    assert (o1.float() - 1).abs().max().item() <= 1e-2
    # (2) causality by perturbation (ch01/test_ch01.py:22-39): changing the last 100 keys/values leaves earlier rows bit-identical
    k2, v2 = k.clone(), v.clone()
    k2[:, :, -100:] += 3
    v2[:, :, -100:] -= 2
    o2 = pli.flash_attention_forward(q, k2, v2, causal=True)
    assert torch.equal(o2[:, :, :N - 100], o[:, :, :N - 100])
    assert not torch.equal(o2[:, :, N - 100:], o[:, :, N - 100:])
    # (3) sampled rows against the oracle (fp32, CPU): every 4th head of two batches, 24 rows spread over N
    rows = torch.tensor([0, 1, 127, 128, 129, 255, 256, 1000, 2047, 2048, 4095, 4096, 4097, 5000, 6143, 6144, 7000,
                         7935, 7936, 8063, 8064, 8100, 8190, 8191])
    G = Hq // Hkv
    for b in (0, 3):
        for h in range(0, Hq, 4):
            kk = k[b, h // G].float().cpu()
            vv = v[b, h // G].float().cpu()
            s = (q[b, h, rows.cuda()].float().cpu() @ kk.T) * D ** -0.5
            s = s.masked_fill(torch.arange(N)[None, :] > rows[:, None], float("-inf"))
            ref = torch.softmax(s, -1) @ vv
            assert (o[b, h, rows.cuda()].float().cpu() - ref).abs().max().item() <= 2e-2
            assert (lse[b, h, rows.cuda()].cpu() - torch.logsumexp(s, -1)).abs().max().item() <= LSE_TOL
    # (4) GQA map: q heads 4j..4j+3 with identical q rows give identical outputs (they share kv head j)
    qq = q.clone()
    qq[:, 1::4] = qq[:, 0::4]
    oo = pli.flash_attention_forward(qq, k, v, causal=True)
    assert torch.equal(oo[:, 1::4], oo[:, 0::4])


def test_errors_on_gpu_tensors():
    q = torch.randn(1, 6, 16, 64, device="cuda").bfloat16()
    k = torch.randn(1, 4, 16, 64, device="cuda").bfloat16()
    with pytest.raises(RuntimeError, match="multiple"):
        pli.flash_attention_forward(q, k, k)
    with pytest.raises(ValueError, match="Nq <= Nk"):
        pli.flash_attention_forward(q, q[:, :, :8], q[:, :, :8], causal=True)
    with pytest.raises(RuntimeError, match="dtype"):
        pli.flash_attention_forward(q, q.float(), q.float())


@pytest.mark.parametrize("shape,causal", [
    ((8, 16, 4, 256, 256, 128), True),       # many short items: 1-4 half-steps each, softmax runs ahead across items
    ((4, 32, 8, 128, 128, 128), False),      # one KV tile per item
    ((2, 32, 8, 2048, 2048, 128), True),     # more items than SMs, early-step O rescales
    ((1, 8, 8, 1024, 1024, 64), True),
])
def test_prefill_is_deterministic_and_race_free(shape, causal):
    """Bitwise run-to-run reproducibility (a missed barrier shows up as sporadic differences), with inputs whose
    scores grow along the sequence so that the lazy O rescale actually fires at many steps."""
    B, Hq, Hkv, Nq, Nk, D = shape
    q, k, v = orc.seeded_qkv(77, B, Hq, Hkv, Nq, Nk, D)
    k = k * torch.linspace(0.5, 6.0, Nk).view(1, 1, Nk, 1)     # later keys score higher: running max keeps moving
    qd, kd, vd = q.bfloat16().cuda(), k.bfloat16().cuda(), v.bfloat16().cuda()
    o0, l0 = pli.flash_attention_forward(qd, kd, vd, causal=causal, return_lse=True)
    ro, rl = orc.flash_attention_oracle(qd, kd, vd, causal=causal)
    assert (o0.float().cpu() - ro).abs().max().item() <= 2e-2
    assert (l0.cpu() - rl).abs().max().item() <= 1e-3
    for _ in range(25):
        o, l = pli.flash_attention_forward(qd, kd, vd, causal=causal, return_lse=True)
        assert torch.equal(o, o0) and torch.equal(l, l0)


def _run_variant_check(mode):
    """The kernel variants behind experiment switches exist only in the tuning build of the library (-DPLI_TUNING=1:
    the product build has one code path); tests/variant_check.py runs them in a child process that loads that build."""
    import subprocess
    import sys
    from physics_llm_inference_b200 import build
    lib = os.path.join(build.OBJ, "libpli_attention_tuning.so")
    if not os.path.exists(lib):
        build.build(variant="tuning", defines=["-DPLI_TUNING=1"])
    env = dict(os.environ, PLI_LIB_PATH=lib)
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "variant_check.py"), mode],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


def test_pair_mma_kernel_matches_per_cta_mma_kernel_bitwise():
    """The CTA-pair MMA kernel (tcgen05.mma.cta_group::2: M = 256 across the two CTAs of a cluster, each CTA holding half
    of K / V in shared memory, barriers collected in the leader CTA; the product's choice when a CTA pair shares its K/V)
    does the same arithmetic in the same order as the per-CTA MMA kernel with TMA multicast (tuning flag bit 6): outputs
    bit-identical, run after run, and within tolerance of the oracle."""
    _run_variant_check("pair")


def test_wide_kernel_variant_parity():
    """The experimental wide kernel (one 128-row Q tile per CTA, 128-key S tiles in three TMEM buffers; tuning flags 4 and
    36): oracle tolerance, closeness to the product kernel, bitwise run-to-run reproducibility."""
    _run_variant_check("wide")


def test_product_build_has_no_experiment_switches():
    """the product library has one code path; asking it for a kernel-selection flag is an error."""
    lib = _lib.load()
    if os.environ.get("PLI_LIB_PATH"):
        pytest.skip("a non-product library is loaded")
    assert lib.pli_debug_prefill_trace(None, 0, 0) == 0
    assert lib.pli_debug_prefill_trace(None, 0, 4) != 0


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("bs,D,Hkv,G,Nq,lens", [
    (16, 128, 2, 4, 128, [512, 300, 128]),          # C3-style pages, cluster pairs (G % 4 == 0), ragged cache lengths
    (16, 128, 1, 2, 200, [777, 200]),               # head pairs without cluster, ragged Nq tile
    (64, 64, 2, 1, 96, [1000, 97, 640]),            # MHA: row-pair items, big pages, D 64
    (128, 128, 2, 8, 256, [1500, 256]),
    (32, 128, 1, 4, 1, [33, 1, 64]),                # single new token through the prefill kernel
])
def test_paged_prefill_parity(bs, D, Hkv, G, Nq, lens, dtype):
    """Chunked prefill over the paged pools in place == oracle (page gather + ch02 maths with the offset mask)."""
    B, Hq = len(lens), Hkv * G
    _, kp, vp, table, lens_t = orc.seeded_paged(61, B, Hq, Hkv, D, bs, lens, num_layers=2, dtype=dtype)
    g = torch.Generator().manual_seed(62)
    q = torch.randn(B, Hq, Nq, D, generator=g).to(dtype)
    o, lse = pli.flash_attention_paged(q.cuda(), kp.cuda(), vp.cuda(), table.cuda(), lens_t.cuda(), layer=1,
                                       return_lse=True, max_seq_len=max(lens))
    ro, rlse = orc.paged_decode_oracle(q, kp, vp, table, lens_t, layer=1)
    assert (o.float().cpu() - ro).abs().max().item() <= 2e-2
    assert (lse.cpu() - rlse).abs().max().item() <= 1e-3
    # and the same arithmetic as gathering the pages first and calling the contiguous kernel.  Not torch.equal: a
    # masked key carries weight 2^-126 (not 0) on the polynomial-exp2 lanes, and past the sequence end the paged
    # path multiplies that with whatever the page holds where the contiguous path sees TMA zero fill (~1e-38).
    for b, L in enumerate(lens):
        kg = orc.gather_paged(kp, table[b].tolist(), L, 1).unsqueeze(0).transpose(1, 2).cuda()
        vg = orc.gather_paged(vp, table[b].tolist(), L, 1).unsqueeze(0).transpose(1, 2).cuda()
        o2 = pli.flash_attention_forward(q[b:b + 1].cuda(), kg, vg, causal=True)
        assert (o2.float() - o[b:b + 1].float()).abs().max().item() <= 1e-30, b


def test_chunked_prefill_through_paged_cache():
    """ch08-style chunked prefill: a prompt fed in chunks through PagedKVCache.append + prefill_with_paged gives the
    same rows as one causal pass over the whole prompt."""
    torch.manual_seed(5)
    B, Hq, Hkv, D, N, chunk = 2, 8, 2, 128, 640, 256
    dev = torch.device("cuda")
    q = torch.randn(B, Hq, N, D, device=dev, dtype=torch.bfloat16)
    k = torch.randn(B, N, Hkv, D, device=dev, dtype=torch.bfloat16)
    v = torch.randn(B, N, Hkv, D, device=dev, dtype=torch.bfloat16)
    full = pli.flash_attention_forward(q, k.transpose(1, 2), v.transpose(1, 2), causal=True)
    cache = pli.PagedKVCache(num_blocks=128, block_size=16, num_layers=1, num_heads=Hkv, head_dim=D,
                             dtype=torch.bfloat16, device="cuda")
    rids = [7, 9]
    for a in range(0, N, chunk):
        e = min(a + chunk, N)
        cache.append(rids, k[:, a:e], v[:, a:e])
        o = pli.prefill_with_paged(q[:, :, a:e], cache, rids)
        assert (o.float() - full[:, :, a:e].float()).abs().max().item() <= 2e-2


@pytest.mark.parametrize("bs,D,Hkv,G,q_lens,lens", [
    (16, 128, 2, 4, [128, 37, 300, 1], [512, 300, 300, 77]),       # CTA pairs; full, ragged, multi-tile and 1-row chunks
    (16, 128, 1, 2, [200, 129, 5], [777, 129, 900]),               # head pairs, tiles straddling sequence ends
    (64, 64, 2, 1, [96, 260, 31], [1000, 260, 640]),               # MHA row-pair items, D 64
    (32, 128, 1, 3, [64, 257], [64, 300]),                         # odd group (row pairs), whole prompt as the chunk
])
def test_varlen_paged_prefill_parity(bs, D, Hkv, G, q_lens, lens):
    """Ragged query lengths over the paged pools: every sequence matches the oracle run on it alone, and rows of
    the packed output that belong to no tile of a sequence are never clobbered by a neighbour's store."""
    dtype = torch.bfloat16
    B, Hq = len(lens), Hkv * G
    _, kp, vp, table, lens_t = orc.seeded_paged(71, B, Hq, Hkv, D, bs, lens, num_layers=2, dtype=dtype)
    g = torch.Generator().manual_seed(72)
    T = sum(q_lens)
    q = torch.randn(T, Hq, D, generator=g).to(dtype)
    cu = torch.tensor([0] + q_lens, dtype=torch.int32).cumsum(0, dtype=torch.int32)
    o, lse = pli.flash_attention_varlen_paged(q.cuda(), kp.cuda(), vp.cuda(), table.cuda(), lens_t.cuda(), cu.cuda(),
                                              max(q_lens), layer=1, return_lse=True, max_seq_len=max(lens))
    assert o.shape == (T, Hq, D) and lse.shape == (Hq, T)
    for b in range(B):
        a, e = int(cu[b]), int(cu[b + 1])
        qb = q[a:e].transpose(0, 1).unsqueeze(0)                                   # (1, Hq, nq, D)
        ro, rlse = orc.paged_decode_oracle(qb, kp, vp, table[b:b + 1], lens_t[b:b + 1], layer=1)
        assert (o[a:e].float().cpu().transpose(0, 1) - ro[0]).abs().max().item() <= 2e-2, b
        assert (lse[:, a:e].cpu() - rlse[0]).abs().max().item() <= 1e-3, b


def test_prefill_peer_output_single_rank():
    """The fused-gather prefill entry with a world of one: same bits as the plain call, buffers alternate by step."""
    torch.manual_seed(9)
    B, Hq, Hkv, N, D = 2, 8, 2, 500, 128
    q = torch.randn(B, Hq, N, D, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(B, Hkv, N, D, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(B, Hkv, N, D, device="cuda", dtype=torch.bfloat16)
    ref, rlse = pli.flash_attention_forward(q, k, v, causal=True, return_lse=True)
    po = pli.PeerOutput(B, Hq, D, torch.bfloat16, pli.make_shard(0, 1, Hq, Hkv, B), seq_len=N)
    for step in range(3):
        o, lse = pli.flash_attention_forward(q, k, v, causal=True, return_lse=True, peer_out=po)
        assert torch.equal(o, ref) and torch.equal(lse, rlse), step


def _poison_unused(kp, vp, table, lens, bs, layer):
    """NaN into K and Inf into V wherever no sequence has a token: the rest of each last page and every unused page
    (the reference's allocator recycles pages without clearing them, ch07/paged_memory.py:100-110)."""
    used = torch.zeros(kp.shape[0], bs, dtype=torch.bool)
    for b, L in enumerate(lens):
        for t in range(L):
            used[int(table[b, t // bs]), t % bs] = True
    kp[:, layer][~used.to(kp.device)] = float("nan")
    vp[:, layer][~used.to(vp.device)] = float("inf")


@pytest.mark.parametrize("bs,D,Hkv,G,Nq,lens", [
    (16, 128, 2, 4, 128, [500, 300, 129]),          # CTA pairs; last pages 4, 12 and 1 tokens full
    (16, 128, 1, 2, 70, [777, 71]),                 # head pairs, single CTA
    (64, 64, 2, 1, 96, [1000, 97, 641]),            # big pages: most of the last page is garbage
    (128, 128, 1, 4, 33, [130, 33]),                # one-page tiles
])
def test_paged_prefill_ignores_garbage_past_seq_len(bs, D, Hkv, G, Nq, lens):
    """Recycled pages are dirty: NaN/Inf beyond seq_lens[b] must not change a single bit of the paged prefill output
    (keys are masked by select, V rows past the end are zeroed in shared memory before the MMAs read them)."""
    B, Hq = len(lens), Hkv * G
    _, kp, vp, table, lens_t = orc.seeded_paged(81, B, Hq, Hkv, D, bs, lens, num_layers=2, dtype=torch.bfloat16)
    q = torch.randn(B, Hq, Nq, D, generator=torch.Generator().manual_seed(82)).bfloat16().cuda()
    kd, vd = kp.cuda(), vp.cuda()
    args = (table.cuda(), lens_t.cuda())
    o0, l0 = pli.flash_attention_paged(q, kd, vd, *args, layer=1, return_lse=True, max_seq_len=max(lens))
    _poison_unused(kd, vd, table, lens, bs, 1)
    o1, l1 = pli.flash_attention_paged(q, kd, vd, *args, layer=1, return_lse=True, max_seq_len=max(lens))
    assert torch.isfinite(o1.float()).all()
    assert torch.equal(o0, o1) and torch.equal(l0, l1)
    ro, _ = orc.paged_decode_oracle(q.cpu(), kp, vp, table, lens_t, layer=1)
    assert (o1.float().cpu() - ro).abs().max().item() <= 2e-2


def test_varlen_and_mixed_batch_ignore_garbage_past_seq_len():
    bs, D, Hkv, G = 16, 128, 2, 4
    Hq = Hkv * G
    q_lens, lens = [128, 37, 300, 1], [500, 300, 301, 77]
    _, kp, vp, table, lens_t = orc.seeded_paged(83, len(lens), Hq, Hkv, D, bs, lens, dtype=torch.bfloat16)
    T = sum(q_lens)
    q = torch.randn(T, Hq, D, generator=torch.Generator().manual_seed(84)).bfloat16().cuda()
    cu = torch.tensor([0] + q_lens, dtype=torch.int32).cumsum(0, dtype=torch.int32).cuda()
    kd, vd = kp.cuda(), vp.cuda()
    run = lambda: pli.flash_attention_varlen_paged(q, kd, vd, table.cuda(), lens_t.cuda(), cu, max(q_lens),  # noqa: E731
                                                   return_lse=True, max_seq_len=max(lens))
    o0, l0 = run()
    _poison_unused(kd, vd, table, lens, bs, 0)
    o1, l1 = run()
    assert torch.isfinite(o1.float()).all()
    assert torch.equal(o0, o1) and torch.equal(l0, l1)

    # mixed prefill/decode step over a PagedKVCache whose free pages and page tails are dirty
    cache = pli.PagedKVCache(num_blocks=96, block_size=bs, num_layers=1, num_heads=Hkv, head_dim=D, dtype=torch.bfloat16)
    g = torch.Generator(device="cuda").manual_seed(85)
    plens = {1: 200, 2: 77, 3: 431, 4: 18}
    for rid, n in plens.items():
        kn = torch.randn(1, n, Hkv, D, device="cuda", generator=g).bfloat16()
        vn = torch.randn(1, n, Hkv, D, device="cuda", generator=g).bfloat16()
        cache.append([rid], kn, vn)
    qm = torch.randn(150 + 77 + 2, Hq, D, device="cuda", generator=g).bfloat16()
    step = lambda: pli.mixed_batch_attention(qm, cache, [1, 2], [150, 77], [3, 4])  # noqa: E731
    m0 = step()
    used = torch.zeros(96, bs, dtype=torch.bool)
    for rid, n in plens.items():
        for t in range(n):
            used[cache.block_tables[rid].block_indices[t // bs], t % bs] = True
    cache.k_cache[:, 0][~used.cuda()] = float("nan")
    cache.v_cache[:, 0][~used.cuda()] = float("-inf")
    m1 = step()
    assert torch.isfinite(m1.float()).all() and torch.equal(m0, m1)


def test_expanded_zero_stride_views():
    """Broadcast (zero-stride) K/V such as k.expand(B, ...) or an MQA k[:, :1].expand(-1, H, -1, -1) are read at the
    right addresses (TMA cannot broadcast: the C layer routes them to the SIMT kernel, the wrapper materialises them
    first so the tensor-core kernel serves them)."""
    torch.manual_seed(11)
    B, Hq, N, D = 3, 8, 200, 128
    q = torch.randn(B, Hq, N, D, device="cuda", dtype=torch.bfloat16)
    k1 = torch.randn(1, 2, N, D, device="cuda", dtype=torch.bfloat16)
    v1 = torch.randn(1, 2, N, D, device="cuda", dtype=torch.bfloat16)
    ke, ve = k1.expand(B, -1, -1, -1), v1.expand(B, -1, -1, -1)                  # batch stride 0
    assert ke.stride(0) == 0
    o = pli.flash_attention_forward(q, ke, ve, causal=True)
    ref = pli.flash_attention_forward(q, ke.contiguous(), ve.contiguous(), causal=True)
    assert torch.equal(o, ref)
    km, vm = k1[:, :1].expand(B, Hq, -1, -1), v1[:, :1].expand(B, Hq, -1, -1)    # MQA: batch and head stride 0
    o = pli.flash_attention_forward(q, km, vm, causal=False)
    ro, _ = orc.naive_attention_oracle(q, km, vm)
    assert (o.float().cpu() - ro).abs().max().item() <= 2e-2
    # straight through the C ABI (no wrapper normalisation): zero strides select the SIMT kernel, not a wrong tensor map
    assert _lib.load().pli_prefill_kernel_kind(D, _lib.PLI_BF16, _lib.i64(*q.stride()[:3]), _lib.i64(*ke.stride()[:3]),
                                               _lib.i64(*ve.stride()[:3]), _lib.i64(*q.stride()[:3]), q.data_ptr(),
                                               ke.data_ptr(), ve.data_ptr(), q.data_ptr()) == _lib.PLI_KIND_SIMT
    out = torch.empty_like(q)
    rc = _lib.load().pli_prefill_fwd(q.data_ptr(), ke.data_ptr(), ve.data_ptr(), out.data_ptr(), None, B, Hq, 2, N, N, D,
                                     _lib.i64(*q.stride()[:3]), _lib.i64(*ke.stride()[:3]), _lib.i64(*ve.stride()[:3]),
                                     _lib.i64(*out.stride()[:3]), D ** -0.5, 1, _lib.PLI_BF16,
                                     torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    assert (out.float() - ref.float()).abs().max().item() <= 2e-2


def test_rows_without_a_visible_key_give_zero_output_and_minus_inf_lse():
    """A caller error the kernels survive: seq_lens[b] < Nq leaves the first Nq - seq_lens[b] query rows
    of that sequence with no visible key.  They come out as O = 0, LSE = -inf (not NaN), the other rows are unaffected,
    and `validate=True` reports the violation on the host instead."""
    bs, D, Hkv, G, Nq = 16, 128, 2, 4, 64
    lens = [200, 40]                                   # sequence 1 is shorter than its 64 query tokens
    _, kp, vp, table, lens_t = orc.seeded_paged(95, 2, Hkv * G, Hkv, D, bs, lens, dtype=torch.bfloat16)
    q = torch.randn(2, Hkv * G, Nq, D, generator=torch.Generator().manual_seed(96)).bfloat16()
    o, lse = pli.flash_attention_paged(q.cuda(), kp.cuda(), vp.cuda(), table.cuda(), lens_t.cuda(), return_lse=True,
                                       max_seq_len=max(lens))
    dead = Nq - lens[1]
    assert torch.count_nonzero(o[1, :, :dead]) == 0 and bool(torch.isinf(lse[1, :, :dead]).all()) and bool((lse[1, :, :dead] < 0).all())
    assert torch.isfinite(o.float()).all() and torch.isfinite(lse[1, :, dead:]).all() and torch.isfinite(lse[0]).all()
    ro, rl = orc.paged_decode_oracle(q[:1], kp, vp, table[:1], lens_t[:1])
    assert (o[:1].float().cpu() - ro).abs().max().item() <= 2e-2
    # live rows of the short sequence: query i (i >= dead) sees keys j <= i - dead
    ro1, _ = orc.paged_decode_oracle(q[1:, :, dead:], kp, vp, table[1:], lens_t[1:])
    assert (o[1:, :, dead:].float().cpu() - ro1).abs().max().item() <= 2e-2
    with pytest.raises(ValueError):
        pli.flash_attention_paged(q.cuda(), kp.cuda(), vp.cuda(), table.cuda(), lens_t.cuda(), validate=True)
    bad = table.clone()
    bad[0, 1] = kp.shape[0] + 7
    with pytest.raises(ValueError):
        pli.flash_decode(q[:, :, :1].cuda(), kp.cuda(), vp.cuda(), lens_t.cuda(), block_tables=bad.cuda(), validate=True)
    with pytest.raises(RuntimeError):
        pli.kv_append(kp.cuda(), vp.cuda(), kp[:2, 0, :4].cuda(), vp[:2, 0, :4].cuda(), 10 ** 6, block_tables=table.cuda())
