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


def _tp_worker(rank, world, port, out_dir):
    import physics_llm_inference_b200 as pli
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    pli.init_distributed("gloo")
    torch.manual_seed(0)
    full = pli.GroupedQueryAttention(64, 4, 2)
    tp = pli.TensorParallelGQA.from_full(full, world, rank)
    # host logic only (no attention call on CPU): the shard holds this rank's rows / columns, and reduce() sums
    nq = tp.num_heads * tp.head_dim
    ok = torch.equal(tp.q_proj.weight, full.q_proj.weight[rank * nq:(rank + 1) * nq])
    ok = ok and torch.equal(tp.o_proj.weight, full.o_proj.weight[:, rank * nq:(rank + 1) * nq])
    ok = ok and (tp.num_heads, tp.num_kv_heads) == (2, 1)
    part = torch.full((2, 3), float(rank + 1))
    ok = ok and torch.equal(tp.reduce(part), torch.full((2, 3), float(sum(range(1, world + 1)))))
    torch.save(torch.tensor(int(ok)), os.path.join(out_dir, f"tp{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_tensor_parallel_gqa_plumbing(tmp_path):
    port = _free_port()
    mp.spawn(_tp_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert int(torch.load(os.path.join(tmp_path, f"tp{r}.pt"))) == 1
