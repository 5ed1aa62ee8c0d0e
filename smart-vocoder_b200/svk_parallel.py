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

import os
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
                  device, root: int = 0, group=None, micro_batches: int = 1,
                  out_host: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """Scatter `mel [N, n_mel, T]` / `lengths [N]` (present on `root` only), run
    `infer_fn(mel_shard, lengths_shard) -> pcm [n, 1, samples_per_frame*T]` on every rank, gather PCM on `root`.
    Returns the full `[N, 1, samples_per_frame*T]` tensor on root, None elsewhere.

    `mel` / `lengths` may live on `device` or (root, CUDA) in pinned host memory; with `out_host` (root: a pinned CPU
    tensor `[N, 1, spf*T]`) the PCM is also copied to the host and `out_host` is returned after synchronising.

    `micro_batches` > 1 splits every rank's shard into that many contiguous pieces and pipelines them on CUDA: the
    H2D + NCCL scatter of piece m+1 and the NCCL gather + D2H of piece m-1 run on side streams under the kernels of
    piece m, so the root's serial PCIe / gather leg (all PCM of the box funnels through rank 0) hides behind compute
    instead of following it.  Every rank issues the same sequence of NCCL group calls.  On CPU (gloo) the pieces
    simply run one after another -- same result."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bounds = shard_bounds(n_utterances, world)
    M = max(1, int(micro_batches))
    L = samples_per_frame * T
    cuda = torch.device(device).type == "cuda"
    # piece m of rank r = utterances pieces[r][m]
    pieces = [[(a + x, a + y) for x, y in shard_bounds(b - a, M)] for a, b in bounds]
    if cuda:
        cur = torch.cuda.current_stream(device)
        comm, copy = torch.cuda.Stream(device), torch.cuda.Stream(device)
        comm.wait_stream(cur), copy.wait_stream(cur)
    full = torch.empty(n_utterances, 1, L, device=device, dtype=torch.float32) if rank == root else None
    mel_loc, len_loc, ev_sc = [], [], []
    # ---- stage the inputs: (root) H2D of piece m on the copy stream, then the scatter of piece m on the comm stream
    for m in range(M):
        a, b = pieces[rank][m]
        ml = torch.empty(b - a, n_mel, T, device=device, dtype=torch.float32)
        ll = torch.empty(b - a, device=device, dtype=torch.int64)
        chunks_m = chunks_l = None
        if rank == root:
            if cuda and not mel.is_cuda:
                with torch.cuda.stream(copy):
                    chunks_m = [mel[x:y].to(device, non_blocking=True) for x, y in (pieces[r][m] for r in range(world))]
                    chunks_l = [lengths[x:y].to(device, non_blocking=True) for x, y in (pieces[r][m] for r in range(world))]
                comm.wait_stream(copy)
            else:
                chunks_m = [mel[x:y] for x, y in (pieces[r][m] for r in range(world))]
                chunks_l = [lengths[x:y] for x, y in (pieces[r][m] for r in range(world))]
        if cuda:
            with torch.cuda.stream(comm):
                _scatter(ml, chunks_m, root, group)
                _scatter(ll, chunks_l, root, group)
                ev = torch.cuda.Event()
                ev.record(comm)
            for t_ in (chunks_m or []) + (chunks_l or []) + [ml, ll]:
                t_.record_stream(comm)
            ev_sc.append(ev)
        else:
            _scatter(ml, chunks_m, root, group)
            _scatter(ll, chunks_l, root, group)
        mel_loc.append(ml), len_loc.append(ll)
    # ---- compute piece m while piece m-1 is gathered (and copied to the host)
    for m in range(M):
        a, b = pieces[rank][m]
        if cuda:
            cur.wait_event(ev_sc[m])
        pcm = infer_fn(mel_loc[m], len_loc[m]) if b > a else torch.empty(0, 1, L, device=device)
        outs = [full[x:y] for x, y in (pieces[r][m] for r in range(world))] if rank == root else None
        if cuda:
            comm.wait_stream(cur)
            with torch.cuda.stream(comm):
                _gather(pcm, outs, root, group)
            pcm.record_stream(comm)
            if rank == root and out_host is not None:
                copy.wait_stream(comm)
                with torch.cuda.stream(copy):
                    for x, y in (pieces[r][m] for r in range(world)):
                        if y > x:
                            out_host[x:y].copy_(full[x:y], non_blocking=True)
        else:
            _gather(pcm, outs, root, group)
            if rank == root and out_host is not None:
                out_host.copy_(full) if m == M - 1 else None
    if cuda:
        cur.wait_stream(comm), cur.wait_stream(copy)
        if full is not None:
            full.record_stream(comm), full.record_stream(copy)
    if rank == root and out_host is not None:
        if cuda:
            torch.cuda.current_stream(device).synchronize()
        return out_host
    return full


class SharedHostBuffer:
    """A host array that every rank process of the box maps (POSIX shared memory) and registers with CUDA as pinned
    memory, so each GPU can DMA its own shard of a batch in or out over ITS OWN PCIe link -- instead of the whole batch
    funnelling through rank 0's link and an NCCL gather.  Rank `root` creates it, the others attach by name."""

    def __init__(self, name: str, shape, dtype: torch.dtype, create: bool, register: bool = True):
        from multiprocessing import shared_memory
        import numpy as np
        self._np_dtype = {torch.float32: np.float32, torch.int64: np.int64, torch.int16: np.int16}[dtype]
        nbytes = int(np.prod(shape)) * np.dtype(self._np_dtype).itemsize
        if create:
            # tmpfs does not fail at ftruncate: a segment larger than what /dev/shm has left dies with SIGBUS on the first
            # write.  Refuse up front instead (containers often mount /dev/shm with 64 MB).
            try:
                st = os.statvfs("/dev/shm")
                free = st.f_bavail * st.f_frsize
            except OSError:
                free = None
            if free is not None and nbytes > 0.9 * free:
                raise MemoryError(f"SharedHostBuffer {name}: {nbytes} bytes requested, /dev/shm has {free} free")
        self._shm = shared_memory.SharedMemory(name=name, create=create, size=max(nbytes, 1))
        if not create:
            # Python < 3.13 registers attached segments with this process's resource tracker too, which then unlinks
            # them at exit under the owner's feet (and warns); only the creating rank owns the name
            try:
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self._shm._name, "shared_memory")
            except Exception:  # pragma: no cover
                pass
        self.array = np.ndarray(tuple(shape), dtype=self._np_dtype, buffer=self._shm.buf)
        self.tensor = torch.from_numpy(self.array)
        self._registered, self._owner = False, create
        if register and torch.cuda.is_available() and nbytes:
            rc = torch.cuda.cudart().cudaHostRegister(self.tensor.data_ptr(), nbytes, 0)
            self._registered = int(rc) == 0
        self.name = name

    def close(self):
        if self._registered:
            torch.cuda.cudart().cudaHostUnregister(self.tensor.data_ptr())
            self._registered = False
        self.tensor = self.array = None
        try:
            self._shm.close()
            if self._owner:
                self._shm.unlink()
        except Exception:  # pragma: no cover
            pass


def sharded_infer_direct(infer_fn: Callable[[torch.Tensor, torch.Tensor], torch.Tensor], mel_host: torch.Tensor,
                         lengths_host: torch.Tensor, out_host: torch.Tensor, n_utterances: int, device, group=None,
                         micro_batches: int = 1) -> None:
    """The same sharding with NO NCCL on the data path: `mel_host [N, n_mel, T]`, `lengths_host [N]` and
    `out_host [N, 1, L]` are SharedHostBuffer tensors visible to every rank; rank r copies its own utterances in
    (H2D), runs `infer_fn`, and copies its PCM out (D2H) -- N PCIe links in parallel.  With `micro_batches` > 1 the D2H
    of piece m runs on a side stream under the kernels of piece m+1.  Ends with a barrier: afterwards `out_host` is
    complete on every rank.  For callers that can hand the library shared host buffers; `sharded_infer` is the form
    that needs nothing but rank 0's tensors."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    a, b = shard_bounds(n_utterances, world)[rank]
    cuda = torch.device(device).type == "cuda"
    if cuda:
        cur = torch.cuda.current_stream(device)
        copy = torch.cuda.Stream(device)
        copy.wait_stream(cur)
    for x, y in shard_bounds(b - a, max(1, int(micro_batches))):
        if y <= x:
            continue
        m = mel_host[a + x:a + y].to(device, non_blocking=True)
        l = lengths_host[a + x:a + y].to(device, non_blocking=True)
        pcm = infer_fn(m, l)
        if cuda:
            copy.wait_stream(cur)
            with torch.cuda.stream(copy):
                out_host[a + x:a + y].copy_(pcm, non_blocking=True)
            pcm.record_stream(copy)
        else:
            out_host[a + x:a + y].copy_(pcm)
    if cuda:
        copy.synchronize()
        cur.synchronize()
    dist.barrier(group)


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
