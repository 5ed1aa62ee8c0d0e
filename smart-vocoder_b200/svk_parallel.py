"""Utterance-batch sharding of the mel->waveform path across the GPUs of one box.

Utterances never interact (no batch statistics anywhere on the path; batch is dim 0 of every
conv -- SURVEY 8e), so the path shards by contiguous blocks of utterances with NO data-path
collective.  NCCL is used only at the edges, as the north_star asks: scatter mel (+lengths) from
rank 0, gather PCM to rank 0.  The reference has no multi-GPU inference at all; the
process model mirrors its training launcher (one process per GPU, env:// rendezvous, train.py:49,61).

`sharded_infer` is backend-agnostic: tests run it over gloo on CPU tensors with a stand-in
`infer_fn`; bench.py runs it over NCCL with `SynthesizerTrn.infer`.

`time_sharded_infer` shards ONE batch of very long utterances along time instead: rank r
synthesises frames [t0_r, t1_r) from a window widened by `halo` frames of context per side (svk_halo_frames),
which reproduces the whole-utterance result exactly, so again there is no data-path collective -- only the
scatter of the overlapping input windows and the gather of the PCM segments.
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


def time_shard_bounds(T: int, world_size: int, halo: int) -> List[Tuple[int, int, int, int]]:
    """Per rank (t0, t1, a, b): frames [t0, t1) are the rank's own, [a, b) is the window it needs as input
    (own frames + `halo` frames of context per side, clipped to [0, T))."""
    out = []
    for t0, t1 in shard_bounds(T, world_size):
        if t1 > t0:
            out.append((t0, t1, max(0, t0 - halo), min(T, t1 + halo)))
        else:
            out.append((t0, t1, t0, t1))
    return out


def time_sharded_infer(window_fn: Callable[[torch.Tensor, torch.Tensor, torch.Tensor, int, int], torch.Tensor],
                       mel: Optional[torch.Tensor], lengths: Optional[torch.Tensor], eps: Optional[torch.Tensor],
                       B: int, n_mel: int, n_lat: int, T: int, samples_per_frame: int, halo: int, device, root: int = 0,
                       group=None) -> Optional[torch.Tensor]:
    """Shard `mel [B, n_mel, T]` / `eps [B, n_lat, T]` / `lengths [B]` (on `root`) along TIME.  Every rank receives its
    halo-widened window and the lengths as seen from the window start, and calls
    `window_fn(mel_w, lengths_w, eps_w, lo, hi) -> pcm [B, 1, samples_per_frame * (hi - lo)]`, where [lo, hi) are its own
    frames in window coordinates (SynthesizerTrn.infer_window does exactly this).  Root returns `[B, 1, spf * T]`."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bounds = time_shard_bounds(T, world, halo)
    t0, t1, a, b = bounds[rank]
    w = b - a
    mel_w = torch.empty(B, n_mel, w, device=device, dtype=torch.float32)
    eps_w = torch.empty(B, n_lat, w, device=device, dtype=torch.float32)
    len_w = torch.empty(B if w else 0, device=device, dtype=torch.int64)
    if rank == root:
        _scatter(mel_w, [mel[:, :, x:y].contiguous() for _, _, x, y in bounds], root, group)
        _scatter(eps_w, [eps[:, :, x:y].contiguous() for _, _, x, y in bounds], root, group)
        _scatter(len_w, [(lengths - x).clamp(0, y - x) if y > x else lengths[:0] for _, _, x, y in bounds], root, group)
    else:
        _scatter(mel_w, None, root, group)
        _scatter(eps_w, None, root, group)
        _scatter(len_w, None, root, group)
    spf = samples_per_frame
    pcm = window_fn(mel_w, len_w, eps_w, t0 - a, t1 - a) if t1 > t0 else torch.empty(B, 1, 0, device=device)
    if rank == root:
        parts = [torch.empty(B, 1, spf * (y - x), device=device, dtype=torch.float32) for x, y, _, _ in bounds]
        _gather(pcm, parts, root, group)
        return torch.cat(parts, dim=2)
    _gather(pcm, None, root, group)
    return None


def length_buckets(lengths: Sequence[int], boundaries: Sequence[int], batch_size: int) -> List[List[int]]:
    """Serving-side counterpart of the reference's DistributedBucketSampler (data_utils.py:130-226): batches of at most
    `batch_size` utterance indices whose lengths fall into the same (boundaries[i], boundaries[i+1]] group, so that a padded
    batch wastes little decoder work (the decoder runs over the padded length).  Unlike the training sampler
    nothing is dropped, duplicated or shuffled: every utterance appears exactly once (lengths outside the boundaries go
    to an extra first / last group), longest first inside a group so a batch's padding is set by its first item."""
    if batch_size < 1:
        raise ValueError("batch_size must be positive")
    b = sorted(int(x) for x in boundaries)
    groups: List[List[int]] = [[] for _ in range(len(b) + 1)]
    for i, n in enumerate(lengths):
        g = 0
        while g < len(b) and int(n) > b[g]:
            g += 1
        groups[g].append(i)
    out: List[List[int]] = []
    for g in groups:
        g.sort(key=lambda i: (-int(lengths[i]), i))
        out += [g[k:k + batch_size] for k in range(0, len(g), batch_size)]
    return out
