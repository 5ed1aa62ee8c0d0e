"""Utterance-batch sharding of the mel->waveform path across the GPUs of one box.

Utterances never interact (no batch statistics anywhere on the path; batch is dim 0 of every
conv -- SURVEY 8e), so the path shards by contiguous blocks of utterances with NO data-path
collective.  NCCL is used only at the edges, as the north_star asks: scatter mel (+lengths) from
rank 0, gather PCM to rank 0.  The reference has no multi-GPU inference at all; the
process model mirrors its training launcher (one process per GPU, env:// rendezvous, train.py:49,61).

`sharded_infer` is backend-agnostic: tests run it over gloo on CPU tensors with a stand-in
`infer_fn`; bench.py runs it over NCCL with `SynthesizerTrn.infer`.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_utterances: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous [start, end) utterance ranges, sizes differing by at most one (first ranks larger)."""
    base, extra = divmod(n_utterances, world_size)
    out, s = [], 0
    for r in range(world_size):
        e = s + base + (1 if r < extra else 0)
        out.append((s, e))
        s = e
    return out


def _scatter(buf_local: torch.Tensor, chunks: Optional[Sequence[torch.Tensor]], src: int, group=None):
    """dist.scatter with uneven chunks emulated by point-to-point sends (NCCL grouped send/recv)."""
    rank = dist.get_rank(group)
    ops = []
    if rank == src:
        for r, c in enumerate(chunks):
            if r == src:
                buf_local.copy_(c)
            elif c.numel():
                ops.append(dist.P2POp(dist.isend, c.contiguous(), r, group))
    elif buf_local.numel():
        ops.append(dist.P2POp(dist.irecv, buf_local, src, group))
    if ops:  # one batch = one NCCL group call (ncclGroupStart/End around the sends / the recv)
        for q in dist.batch_isend_irecv(ops):
            q.wait()


def _gather(local: torch.Tensor, outs: Optional[Sequence[torch.Tensor]], dst: int, group=None):
    rank = dist.get_rank(group)
    ops = []
    if rank == dst:
        for r, o in enumerate(outs):
            if r == dst:
                o.copy_(local)
            elif o.numel():
                ops.append(dist.P2POp(dist.irecv, o, r, group))
    elif local.numel():
        ops.append(dist.P2POp(dist.isend, local.contiguous(), dst, group))
    if ops:
        for q in dist.batch_isend_irecv(ops):
            q.wait()


def sharded_infer(infer_fn: Callable[[torch.Tensor, torch.Tensor], torch.Tensor], mel: Optional[torch.Tensor],
                  lengths: Optional[torch.Tensor], n_utterances: int, n_mel: int, T: int, samples_per_frame: int,
                  device, root: int = 0, group=None) -> Optional[torch.Tensor]:
    """Scatter `mel [N, n_mel, T]` / `lengths [N]` (present on `root` only, already on `device`),
    run `infer_fn(mel_shard, lengths_shard) -> pcm [n, 1, samples_per_frame*T]` on every rank, gather
    PCM on `root`.  Returns the full `[N, 1, samples_per_frame*T]` tensor on root, None elsewhere."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bounds = shard_bounds(n_utterances, world)
    s, e = bounds[rank]
    mel_local = torch.empty(e - s, n_mel, T, device=device, dtype=torch.float32)
    len_local = torch.empty(e - s, device=device, dtype=torch.int64)
    if rank == root:
        _scatter(mel_local, [mel[a:b] for a, b in bounds], root, group)
        _scatter(len_local, [lengths[a:b] for a, b in bounds], root, group)
    else:
        _scatter(mel_local, None, root, group)
        _scatter(len_local, None, root, group)
    pcm_local = infer_fn(mel_local, len_local) if e > s else torch.empty(0, 1, samples_per_frame * T, device=device)
    if rank == root:
        full = torch.empty(n_utterances, 1, samples_per_frame * T, device=device, dtype=torch.float32)
        _gather(pcm_local, [full[a:b] for a, b in bounds], root, group)
        return full
    _gather(pcm_local, None, root, group)
    return None
