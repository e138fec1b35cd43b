"""world_size-2 gloo test of the multi-GPU launcher logic on CPU: each rank computes its head shard
(with the CPU oracle standing in for the kernels) and `gather_heads` must reassemble the full tensor."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import attention_oracle as orc


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, Hq, Hkv, B, out_dir):
    import physics_llm_inference_b200 as pli
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = pli.init_distributed("gloo")
    assert (r, w) == (rank, world)
    q, k, v = orc.seeded_qkv(11, B, Hq, Hkv, 40, 40, 16)
    shard = pli.make_shard(rank, world, Hq, Hkv, B)
    qs, ks, vs = pli.shard_kv_heads(q, k, v, shard)
    o_loc, lse_loc = orc.flash_attention_oracle(qs, ks, vs, causal=True)
    o = pli.gather_heads(o_loc, shard)
    lse = pli.gather_heads(lse_loc, shard)
    full, full_lse = orc.flash_attention_oracle(q, k, v, causal=True)
    ok = torch.equal(o, full) and torch.equal(lse, full_lse) and o.shape == full.shape
    torch.save(torch.tensor(int(ok)), os.path.join(out_dir, f"ok{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("Hq,Hkv,B", [(8, 2, 2), (4, 1, 4)])
def test_gloo_world2_shard_and_gather(tmp_path, Hq, Hkv, B):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, Hq, Hkv, B, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert int(torch.load(os.path.join(tmp_path, f"ok{r}.pt"))) == 1
