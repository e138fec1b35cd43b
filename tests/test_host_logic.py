"""CPU tests of the host side: the reference-facing interface, the allocator mirror against the
reference's golden trace, accounting conventions, the C-ABI exports, and loud failure without CUDA."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

import physics_llm_inference_b200 as pli
from physics_llm_inference_b200 import _lib


def test_library_exports_every_header_symbol():
    lib = _lib.load()
    names = _lib.header_symbols()
    assert len(names) >= 14
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in include/pli_attention.h but not exported"
        assert name in _lib._SIGNATURES, f"{name} has no ctypes signature"
    assert lib.pli_abi_version() == 1
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for name in names:
        assert f" T {name}" in out


def test_library_is_sm100a_native():
    """The shipped cubin carries tcgen05 / TMA instructions (SASS mnemonics from B200_PROFILING.md)."""
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "UTCHMMA" in sass          # tcgen05.mma
    assert "LDTM" in sass and "STTM" in sass   # tcgen05.ld / st
    assert "UTMALDG" in sass and "UTMASTG" in sass  # TMA load / store


def test_workspace_and_split_helpers_need_no_gpu():
    lib = _lib.load()
    # partials (O and LSE, f32) + one 16-byte arrival-counter pair per (b, q head) upper bound, 8-byte aligned
    assert lib.pli_decode_workspace_bytes(64, 32, 128, 4) == 64 * 32 * 4 * 129 * 4 + 64 * 32 * 16
    assert lib.pli_decode_workspace_bytes(1, 1, 64, 1) == 264 + 16
    assert lib.pli_decode_workspace_bytes(0, 32, 128, 4) == 0
    # split counts (148 SMs assumed without a device): one wave of CTAs, >= 256 tokens per split, counts above eight are
    # multiples of eight (eight splits at a time merge inside a thread-block cluster)
    assert lib.pli_decode_num_splits(256, 1, 1024) == 1
    assert lib.pli_decode_num_splits(64, 8, 4096) == 1
    assert lib.pli_decode_num_splits(8, 8, 8192) == 4
    assert lib.pli_decode_num_splits(4, 8, 16384) == 8
    assert lib.pli_decode_num_splits(1, 8, 32768) == 32
    assert lib.pli_decode_num_splits(1, 8, 700) == 3


def test_config_defaults_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "ch06_flash.npz"))
    cfg = pli.FlashAttentionConfig()
    assert [cfg.block_q, cfg.block_k, cfg.num_warps, cfg.num_stages] == [int(x) for x in g["config_defaults"]]
    assert pli.attention_flops(1, 8, 512, 64) == int(g["flops_c1"][0])
    assert pli.flash_attention_memory_bytes(1, 8, 512, 64)["hbm_bytes"] == int(g["membytes_c1"][0])
    mem = pli.flash_attention_memory_bytes(batch_size=1, num_heads=8, seq_len=1024, head_dim=64)
    assert mem["hbm_bytes"] < mem["naive_hbm_bytes"]            # ch06/test_ch06.py:143-151


def test_algorithmic_flops_convention():
    # C2 = 2.199 TFLOP, C1 = 0.537 GFLOP (0.268 causal)
    assert abs(pli.prefill_algorithmic_flops(4, 32, 8192, 8192, 128, True) - 2.199e12) < 1e9
    assert abs(pli.prefill_algorithmic_flops(1, 8, 512, 512, 64, False) - 0.537e9) < 1e6
    assert abs(pli.prefill_algorithmic_flops(1, 8, 512, 512, 64, True) - 0.268e9) < 1e6


def test_cpu_tensors_fail_loudly():
    q = torch.randn(1, 2, 8, 16)
    with pytest.raises(RuntimeError, match="no\\s+CPU fallback"):
        pli.flash_attention_forward(q, q, q)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pli.flash_decode(q[:, :, :1], torch.zeros(1, 8, 2, 16), torch.zeros(1, 8, 2, 16), 8)
    with pytest.raises(RuntimeError):
        pli.KVCache.create(1, 8, 2, 16, torch.device("cpu"), torch.float32).update(torch.zeros(1, 1, 2, 16),
                                                                                     torch.zeros(1, 1, 2, 16))


def test_signature_matches_reference():
    import inspect
    sig = inspect.signature(pli.flash_attention_forward)
    names = list(sig.parameters)
    assert names[:5] == ["q", "k", "v", "scale", "config"]        # ch06/flash_attention.py:14-20
    assert sig.parameters["scale"].default is None and sig.parameters["config"].default is None
    assert sig.parameters["causal"].kind is inspect.Parameter.KEYWORD_ONLY
    assert sig.parameters["causal"].default is False              # default stays non-causal
    assert sig.parameters["return_lse"].kind is inspect.Parameter.KEYWORD_ONLY
    assert pli.flash_attention is pli.flash_attention_forward


def test_shape_validation_errors():
    # validation happens before any CUDA call, so the messages can be checked on CPU tensors too
    q = torch.randn(2, 8, 4)
    with pytest.raises(RuntimeError, match=r"\(B, H, N, D\)"):
        pli.flash_attention_forward(q, q, q)


# ---- ch07 allocator mirror vs the reference's recorded behaviour ----
def test_paged_allocator_matches_reference_trace(golden_dir):
    g = np.load(os.path.join(golden_dir, "ch07_paged.npz"))
    c = pli.PagedKVCache(num_blocks=20, block_size=16, num_layers=2, num_heads=4, head_dim=64, device="cpu")
    assert c.k_cache is None and c.v_cache is None
    for op, rid, n, err, nblocks, ntok, nfree in g["trace"].tolist():
        got = 0
        try:
            if op == 0:
                c.allocate_blocks(rid, n)
            elif op == 1:
                c.extend_blocks(rid, n)
            else:
                c.free_blocks_for_request(rid)
        except RuntimeError:
            got = 1
        except KeyError:
            got = 2
        assert got == err, (op, rid, n)
        t = c.block_tables.get(rid)
        assert (-1 if t is None else t.num_blocks()) == nblocks
        assert (-1 if t is None else t.num_tokens) == ntok
        assert c.get_num_free_blocks() == nfree
    u = c.get_memory_usage()
    assert [u["total_blocks"], u["used_blocks"], u["free_blocks"], u["block_size_tokens"], u["bytes_per_block"]] == \
        [int(x) for x in g["usage"]]
    bt = pli.BlockTable(request_id=7, block_indices=[3, 1, 2], num_tokens=40)
    assert [bt.request_id, bt.num_blocks(), bt.num_tokens] == [int(x) for x in g["bt"]]


def test_paged_reference_unit_tests():
    """ch07/test_ch07.py:228-323 restated against the mirror."""
    c = pli.PagedKVCache(num_blocks=100, block_size=16, num_layers=2, num_heads=4, head_dim=64, device="cpu")
    t = c.allocate_blocks(request_id=1, num_tokens=50)
    assert (t.request_id, t.num_tokens, len(t.block_indices)) == (1, 50, 4)
    assert len(set(t.block_indices)) == 4
    c.extend_blocks(1, 20)
    assert c.block_tables[1].num_tokens == 70 and len(c.block_tables[1].block_indices) == 5
    free0 = c.get_num_free_blocks()
    assert c.free_blocks_for_request(1) == 5 and c.get_num_free_blocks() == free0 + 5
    assert c.free_blocks_for_request(1) == 0
    u = c.get_memory_usage()
    assert u["total_blocks"] == 100 and u["free_blocks"] == 100 and u["utilization"] == 0.0
    small = pli.PagedKVCache(num_blocks=2, block_size=16, num_layers=2, num_heads=4, head_dim=64, device="cpu")
    with pytest.raises(RuntimeError):
        small.allocate_blocks(request_id=1, num_tokens=100)
    with pytest.raises(KeyError):
        small.extend_blocks(5, 1)


def test_head_sharding_maths():
    for W in (1, 2, 4, 8):
        shards = [pli.make_shard(r, W, 32, 8, 4) for r in range(W)]
        assert sum(s.units for s in shards) == 8 * 4
        assert [s.kv_start for s in shards] == [r * 8 // W for r in range(W)]
        for s in shards:
            assert s.q_start == s.kv_start * 4 and s.q_end == s.kv_end * 4      # consecutive q heads share a kv head
    s = pli.make_shard(5, 8, 8, 2, 8)                                           # more ranks than kv heads: split batch
    assert (s.kv_start, s.kv_end, s.b_start, s.b_end) == (1, 2, 2, 4)
    with pytest.raises(ValueError):
        pli.make_shard(0, 3, 32, 8, 4)


def test_missing_library_is_an_error(tmp_path, monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.PliError, match="no CPU fallback"):
        _lib.load()


def test_prefix_sharing_refcounts():
    """F3: page-granular prefix sharing on top of the reference allocator (host bookkeeping)."""
    c = pli.PagedKVCache(num_blocks=10, block_size=16, num_layers=1, num_heads=2, head_dim=8, device="cpu")
    c.allocate_blocks(1, 40)                                  # 3 pages, last one holds 8 tokens
    t2 = c.fork_request(1, 2, 35)                             # 2 full pages shared + 1 private copy page
    assert t2.num_tokens == 35 and t2.num_blocks() == 3
    assert t2.block_indices[:2] == c.block_tables[1].block_indices[:2]
    assert t2.block_indices[2] != c.block_tables[1].block_indices[2]
    assert c.get_num_free_blocks() == 10 - 3 - 1
    c.extend_blocks(2, 20)                                    # child grows on its own pages
    assert c.block_tables[2].num_blocks() == 4
    assert c.free_blocks_for_request(1) == 3                  # parent leaves: shared pages stay alive
    assert c.get_num_free_blocks() == 10 - 4
    t3 = c.fork_request(2, 3)                                 # share everything (55 tokens: 3 full + copy)
    assert t3.block_indices[:3] == c.block_tables[2].block_indices[:3]
    c.free_blocks_for_request(2)
    c.free_blocks_for_request(3)
    assert c.get_num_free_blocks() == 10 and not c.shared_refs
    with pytest.raises(KeyError):
        c.fork_request(9, 4)
    with pytest.raises(ValueError):
        c.allocate_blocks(5, 10) and c.fork_request(5, 6, 11)


def test_ch06_accounting_surface_matches_reference(golden_dir):
    """attention_memory_bytes / AttentionMemoryStats (ch06/attention_memory.py:6-16,36-61) and
    attention_arithmetic_intensity (:78-86): same integers as the unmodified reference (r2_ch06_surface.npz)."""
    g = np.load(os.path.join(golden_dir, "r2_ch06_surface.npz"))
    for cfg, ref, mb in zip(g["mem_cfgs"], g["mem"], g["mem_total_mb"]):
        s = pli.attention_memory_bytes(*[int(x) for x in cfg])
        assert isinstance(s, pli.AttentionMemoryStats)
        got = [s.batch_size, s.num_heads, s.seq_len, s.head_dim, s.qk_bytes, s.softmax_bytes, s.output_bytes, s.total_bytes]
        assert got == [int(x) for x in ref]
        assert s.total_mb == float(mb)
    s = pli.attention_memory_bytes(batch_size=1, num_heads=8, seq_len=512, head_dim=64)      # default dtype_bytes = 2
    assert s.qk_bytes == 1 * 8 * 512 * 512 * 2
    for cfg, ref in zip(g["ai_cfgs"], g["ai"]):
        assert pli.attention_arithmetic_intensity(int(cfg[0]), int(cfg[1])) == float(ref)
    with pytest.raises(RuntimeError):
        pli.online_softmax(torch.randn(4, 8))                      # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        pli.online_softmax_with_output(torch.randn(4, 8), torch.randn(4, 8, 2))


def test_share_prefix_from_a_radix_hit_is_page_granular(golden_dir):
    """F3 bookkeeping on the host (no pools needed): `(matched, kv_indices)` of the reference's RadixCache.match_prefix
    (r2_ch07_radix.json) -> aliased whole pages + one fresh page for the partly filled remainder, ref-counted."""
    import json
    with open(os.path.join(golden_dir, "r2_ch07_radix.json")) as f:
        r = json.load(f)
    bs = r["block_size"]
    cache = pli.PagedKVCache(num_blocks=32, block_size=bs, num_layers=1, num_heads=2, head_dim=8, device="cpu")
    for pg in r["pages_a"]:
        cache.free_blocks.remove(pg)
    cache.block_tables[1] = pli.BlockTable(request_id=1, block_indices=list(r["pages_a"]), num_tokens=len(r["a_tokens"]))
    assert cache.kv_indices(1) == r["a_kv"]
    assert r["inserted_a"] == len(r["a_tokens"])
    free0 = cache.get_num_free_blocks()
    b = r["queries"]["b_shares_53"]
    tb = cache.share_prefix(2, b["matched"], b["kv_indices"])
    assert tb.num_tokens == 53 and tb.block_indices[:3] == r["pages_a"][:3] and len(tb.block_indices) == 4
    assert cache.get_num_free_blocks() == free0 - 1 and all(cache.shared_refs[p] == 2 for p in r["pages_a"][:3])
    c = r["queries"]["c_shares_32"]
    tc = cache.share_prefix(3, c["matched"], c["kv_indices"])
    assert tc.num_tokens == 32 and tc.block_indices == r["pages_a"][:2]          # page-aligned hit: nothing copied
    assert cache.get_num_free_blocks() == free0 - 1 and cache.shared_refs[r["pages_a"][0]] == 3
    d = r["queries"]["d_shares_0"]
    td = cache.share_prefix(4, d["matched"], d["kv_indices"])
    assert td.num_tokens == 0 and td.block_indices == []
    e = r["queries"]["e_whole"]
    te = cache.share_prefix(5, e["matched"], e["kv_indices"])
    assert te.num_tokens == 70 and te.block_indices[:4] == r["pages_a"][:4] and te.block_indices[4] != r["pages_a"][4]
    # the child grows like any request (ch07/paged_memory.py:76-98) and frees without releasing shared pages
    cache.extend_blocks(2, 20)
    assert cache.block_tables[2].num_tokens == 73 and cache.block_tables[2].num_blocks() == 5
    with pytest.raises(ValueError):
        cache.share_prefix(2, 1, [0])                              # id in use
    with pytest.raises(RuntimeError):
        cache.share_prefix(9, 16, list(range(31 * bs, 32 * bs)) if 31 in cache.free_blocks else [999999] * 16)   # stale
    for rid in (1, 3, 4, 5):
        cache.free_blocks_for_request(rid)
    assert all(p not in cache.free_blocks for p in r["pages_a"][:3])             # still referenced by request 2
    cache.free_blocks_for_request(2)
    assert cache.get_num_free_blocks() == 32 and not cache.shared_refs


def test_combine_pass_reads_partials_only_behind_the_dependency_wait():
    """The split-KV combine kernel is launched programmatically behind the kernel that writes the partials: in its SASS
    no global load may precede ACQBULK (griddepcontrol.wait).  Round 2 found `const __restrict__` loads hoisted above the
    wait (stale partials); this disassembles the built library so the regression shows up without a GPU."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", "-fun", "decode_combine_kernel", _lib.LIB_PATH], capture_output=True, text=True)
    if sass.returncode != 0 or "Function" not in sass.stdout:
        sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True)
    kernels, cur = {}, None
    for line in sass.stdout.splitlines():
        if "Function :" in line:
            cur = line.split("Function :")[1].strip() if "decode_combine_kernel" in line else None
            if cur:
                kernels[cur] = []
        elif cur and "/*" in line:
            kernels[cur].append(line)
    assert len(kernels) >= 3, "combine kernel instantiations not found in the library"
    for name, lines in kernels.items():
        wait = next((i for i, l in enumerate(lines) if "ACQBULK" in l), None)
        assert wait is not None, f"{name}: no griddepcontrol.wait (ACQBULK)"
        early = [l for l in lines[:wait] if " LDG" in l or "LD.E" in l]
        assert not early, f"{name}: global loads in front of the dependency wait:\n" + "\n".join(early)
