"""Multi-GPU launcher pieces: shard attention over KV-head groups, gather only on request.

The reference has no distributed code (ch09's tensor parallelism is shape-only and
`nccl_primitives.py` is a cost model).  Attention needs no exchange step: every (batch, KV-head
group) is independent because softmax never crosses heads or batch and a KV head is shared only
by its own G = Hq/Hkv query heads (ch01/gqa.py:14,30-31).  So each rank (one process per GPU) owns
a contiguous slice of KV heads with their query heads — what a tensor-parallel server holds
anyway — and runs the single-GPU kernels on it.  NCCL is used only by `gather_heads`, when the
caller wants the full (B, Hq, N, D) output on every rank.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class HeadShard:
    """Which heads (and, if there are more ranks than KV heads, which batch rows) a rank owns."""
    rank: int
    world_size: int
    num_heads: int
    num_kv_heads: int
    batch: int
    kv_start: int
    kv_end: int
    b_start: int
    b_end: int

    @property
    def group_size(self) -> int:
        return self.num_heads // self.num_kv_heads

    @property
    def q_start(self) -> int:
        return self.kv_start * self.group_size

    @property
    def q_end(self) -> int:
        return self.kv_end * self.group_size

    @property
    def units(self) -> int:
        """(batch, KV-head group) units on this rank."""
        return (self.kv_end - self.kv_start) * (self.b_end - self.b_start)


def make_shard(rank: int, world_size: int, num_heads: int, num_kv_heads: int, batch: int) -> HeadShard:
    if num_heads % num_kv_heads != 0:
        raise ValueError(f"num_heads ({num_heads}) must be a multiple of num_kv_heads ({num_kv_heads})")
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    if num_kv_heads % world_size == 0:
        per = num_kv_heads // world_size
        return HeadShard(rank, world_size, num_heads, num_kv_heads, batch, rank * per, (rank + 1) * per, 0, batch)
    if world_size % num_kv_heads == 0 and batch % (world_size // num_kv_heads) == 0:
        ways = world_size // num_kv_heads            # ranks per KV head: split the batch between them
        per_b = batch // ways
        head, part = rank // ways, rank % ways
        return HeadShard(rank, world_size, num_heads, num_kv_heads, batch, head, head + 1, part * per_b,
                         (part + 1) * per_b)
    raise ValueError(
        f"cannot shard {num_kv_heads} KV heads x batch {batch} evenly over {world_size} ranks "
        "(need Hkv % W == 0, or W % Hkv == 0 with batch divisible by W / Hkv)")


def shard_kv_heads(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, shard: HeadShard):
    """Slice full (B, H, N, D) q/k/v down to this rank's heads (views, no copy)."""
    qs = q[shard.b_start:shard.b_end, shard.q_start:shard.q_end]
    ks = k[shard.b_start:shard.b_end, shard.kv_start:shard.kv_end]
    vs = v[shard.b_start:shard.b_end, shard.kv_start:shard.kv_end]
    return qs, ks, vs


def gather_heads(o_local: torch.Tensor, shard: HeadShard, group=None) -> torch.Tensor:
    """All-gather head-sharded outputs (B_loc, Hq_loc, ...) into the full (B, Hq, ...) tensor.

    One `all_gather_into_tensor` over NCCL (NVLink 5 / NVSwitch) on GPU tensors, gloo on CPU tensors
    in tests; the rank-major result is permuted back to head order locally."""
    W = shard.world_size
    if W == 1:
        return o_local
    x = o_local.contiguous()
    # the flat collective wants the output as the concatenation along dim 0 (gloo insists on that shape)
    flat = torch.empty((W * x.shape[0], *x.shape[1:]), dtype=x.dtype, device=x.device)
    buf = flat.view(W, *x.shape)
    try:
        dist.all_gather_into_tensor(flat, x, group=group)
    except NotImplementedError:
        # a backend without the flat collective (never NCCL): the list form.  A RuntimeError (an NCCL failure, a shape
        # mismatch between ranks) is an error and propagates — it must not turn into a second, slower collective.
        parts = [torch.empty_like(x) for _ in range(W)]
        dist.all_gather(parts, x, group=group)
        buf = torch.stack(parts, 0)
    if shard.num_kv_heads % W == 0:
        # (W, B, Hq/W, ...) -> (B, W, Hq/W, ...) -> (B, Hq, ...)
        return buf.transpose(0, 1).reshape(x.shape[0], W * x.shape[1], *x.shape[2:])
    ways = W // shard.num_kv_heads
    # rank = head * ways + part: (Hkv, ways, B/ways, G, ...) -> (ways, B/ways, Hkv, G, ...) -> (B, Hq, ...)
    buf = buf.reshape(shard.num_kv_heads, ways, *x.shape)
    buf = buf.permute(1, 2, 0, 3, *range(4, buf.dim()))
    return buf.reshape(shard.batch, shard.num_heads, *x.shape[2:])


class PeerOutput:
    """Decode output that every rank holds IN FULL, filled by all ranks' decode kernels over NVLink peer memory
    (`flash_decode(..., peer_out=...)` -> pli_decode_fwd_scatter + pli_peer_publish_wait): the all-gather that
    would follow a head-sharded decode step is done by the kernel's own stores.

    The buffers live in torch symmetric memory (one allocation per rank, mapped into every process of the group):
    [done words, one per rank | ready words at byte 64 | step counter at byte 256, CTA counter at 260 | pad to 512]
    [buffer 0 (B, Hq, D)][buffer 1].
    The step counter is read on the device, so the step can be captured in a CUDA graph; `epoch` mirrors it on
    the host (call `advance()` after replaying a captured step).

    Two protocols (include/pli_attention.h), fixed by the first call that uses the object (`mode`):
      "gather"   decode on the TMA kernel: ONE launch, ONE buffer (buffer 0, a fixed address: captures into a CUDA graph
                 next to its consumers without a copy); a `ready` credit from every receiver replaces the second buffer,
                 the last CTA of the grid publishes / waits for the `done` words (pli_decode_fwd_gather);
      "scatter"  prefill, and decode on the SIMT kernel: step e writes buffer e & 1 so that a rank never overwrites data
                 a slower peer is still reading, and a one-warp kernel publishes / waits behind the compute kernel."""

    def __init__(self, batch: int, num_heads: int, head_dim: int, dtype: torch.dtype, shard: HeadShard, *, group=None,
                 device=None, seq_len: int | None = None, graph_safe: bool = False):
        """Output of a decode step (B, Hq, D), or with `seq_len` of a prefill call (B, Hq, seq_len, D).

        graph_safe: also allocate a FIXED output tensor.  Under CUDA-graph capture the double buffer a replay writes
        alternates with the device step counter, so a consumer captured in the same graph (o_proj, the next layer) would
        read a stale buffer on every other replay; with graph_safe=True a captured step ends with a device-side copy
        from the buffer of that step (parity read on the device) into `stable`, and the call returns `stable`.  Capturing
        a step without it raises.  Costs one extra output-sized buffer and one pass over it per captured step.

        Stream contract: every kernel that reads the output of step e must be enqueued on the stream that issues step
        e + 1 (or be ordered before it by an event): a rank passes the flag wait of step e + 1 only after all peers have
        run, in stream order, the readers of step e — that is what makes two buffers enough."""
        if shard.world_size > 8:
            raise ValueError("PeerOutput supports up to 8 ranks (one NVSwitch domain)")
        self.shard, self.dtype = shard, dtype
        self.shape = (batch, num_heads, head_dim) if seq_len is None else (batch, num_heads, seq_len, head_dim)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.esz = torch.empty((), dtype=dtype).element_size()
        self.numel = batch * num_heads * head_dim * (1 if seq_len is None else seq_len)
        if (self.numel * self.esz) % 16:
            raise ValueError("PeerOutput needs an output of a multiple of 16 bytes")
        self.buf_bytes = -(-self.numel * self.esz // 256) * 256
        self.header_bytes = 512
        total = self.header_bytes + 2 * self.buf_bytes
        if shard.world_size > 1:
            import torch.distributed._symmetric_memory as symm
            self.storage = symm.empty(total, dtype=torch.uint8, device=self.device)
            self.storage.zero_()
            torch.cuda.synchronize(self.device)
            grp = dist.group.WORLD if group is None else group
            self.handle = symm.rendezvous(self.storage, grp.group_name)
            self.base_ptrs = [int(p) for p in self.handle.buffer_ptrs]
            self.handle.barrier()                             # every rank has zeroed its flags before anyone writes
        else:
            self.storage = torch.zeros(total, dtype=torch.uint8, device=self.device)
            self.handle = None
            self.base_ptrs = [self.storage.data_ptr()]
        self.epoch = 0                                        # host mirror of the device step counter
        self.mode = None                                      # "gather" | "scatter", see the class docstring
        self.graph_safe = graph_safe
        self._stable = None                                   # the fixed output of graph_safe "scatter" steps, on demand

    @property
    def stable(self):
        if self._stable is None and self.graph_safe:
            self._stable = torch.empty(self.shape, dtype=self.dtype, device=self.device)
        return self._stable

    def use_mode(self, mode: str) -> None:
        if self.mode is None:
            self.mode = mode
        elif self.mode != mode:
            raise RuntimeError(f"this PeerOutput has been used with the '{self.mode}' protocol; a '{mode}' step on the same "
                               "object would mix one-buffer and two-buffer steps (use a second PeerOutput)")

    def finish_step(self, ps, stream: int) -> torch.Tensor:
        """Called by the fused entry points after pli_peer_publish_wait: the tensor the caller gets for this step."""
        from . import _lib
        if not torch.cuda.is_current_stream_capturing():
            _lib.raise_on_device_fault()                      # e.g. a peer that never arrived in an EARLIER step
            return self.advance()
        if self.mode == "gather":
            return self.buffer(0)                             # one buffer, a fixed address: nothing to copy
        if self.stable is None:
            raise RuntimeError(
                "a fused-gather step is being captured in a CUDA graph, but this PeerOutput has no fixed output: the "
                "double buffer a replay writes alternates with the step parity.  Build it with PeerOutput(..., "
                "graph_safe=True) (the captured step then copies into `.stable`), and call advance(n) after replays")
        import ctypes
        _lib.check(_lib.load().pli_peer_select_copy(ctypes.byref(ps), self.stable.data_ptr(), self.numel * self.esz,
                                                    self.esz, stream))
        return self.stable

    def buffer(self, index: int) -> torch.Tensor:
        """This rank's copy of output buffer `index`, shaped (B, Hq, D) or (B, Hq, N, D)."""
        a = self.header_bytes + index * self.buf_bytes
        return self.storage[a:a + self.buf_bytes].view(self.dtype)[:self.numel].view(self.shape)

    def advance(self, steps: int = 1) -> torch.Tensor:
        """Account for `steps` steps launched on the device (a direct call does this itself; call it after replaying
        a CUDA graph that holds captured steps).  Returns the buffer of the latest step."""
        self.epoch += steps
        return self.buffer(0 if self.mode == "gather" else self.epoch & 1)

    @property
    def output_ptrs(self):
        return [p + self.header_bytes for p in self.base_ptrs]

    @property
    def flag_ptrs(self):
        return list(self.base_ptrs)

    @property
    def ready_ptrs(self):
        return [p + 64 for p in self.base_ptrs]

    @property
    def epoch_ptr(self) -> int:
        return self.base_ptrs[self.shard.rank] + 256

    @property
    def cta_counter_ptr(self) -> int:
        return self.base_ptrs[self.shard.rank] + 260

    @property
    def buffer_stride(self) -> int:
        return self.buf_bytes // self.esz

    @property
    def slice_offset(self) -> int:
        """Element offset of this rank's first (batch row, head) inside the full tensor."""
        per_head = self.numel // (self.shape[0] * self.shape[1])
        return (self.shard.b_start * self.shape[1] + self.shard.q_start) * per_head


def init_distributed(backend: str | None = None):
    """Join the process group described by RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (torchrun).
    Returns (rank, world_size, local_rank).  Single-process runs return (0, 1, 0) without a group."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local
