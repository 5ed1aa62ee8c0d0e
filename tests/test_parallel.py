"""world_size-2 gloo tests (CPU) for the utterance sharding plumbing used by bench.py --gpus N."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_infer(mel, lengths):
    # deterministic stand-in with the right shape: 4 samples per frame, depends on mel and length
    B, C, T = mel.shape
    pcm = mel.sum(1, keepdim=True).repeat_interleave(4, dim=2)
    return pcm + lengths.view(B, 1, 1).float()


def _worker(rank, world, port, n_utt, q):
    import sys
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    import svk_parallel as P
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        T, C = 6, 3
        g = torch.Generator().manual_seed(0)
        mel = torch.randn(n_utt, C, T, generator=g)
        lengths = torch.arange(n_utt, dtype=torch.int64) + 1
        out = P.sharded_infer(_fake_infer, mel if rank == 0 else None, lengths if rank == 0 else None, n_utt, C, T, 4,
                              torch.device("cpu"))
        # pipelined form (micro-batches + host egress): same result, same NCCL-call sequence on every rank
        host = torch.zeros(n_utt, 1, 4 * T) if rank == 0 else None
        out2 = P.sharded_infer(_fake_infer, mel if rank == 0 else None, lengths if rank == 0 else None, n_utt, C, T, 4,
                               torch.device("cpu"), micro_batches=3, out_host=host)
        # direct form: every rank reads / writes its own rows of buffers shared between the rank processes
        name = f"svk_test_{port}"
        shapes = ((n_utt, C, T), torch.float32), ((n_utt,), torch.int64), ((n_utt, 1, 4 * T), torch.float32)
        if rank == 0:
            bufs = [P.SharedHostBuffer(f"{name}_{i}", sh, dt, create=True) for i, (sh, dt) in enumerate(shapes)]
            bufs[0].tensor.copy_(mel), bufs[1].tensor.copy_(lengths), bufs[2].tensor.zero_()
        dist.barrier()
        if rank != 0:
            bufs = [P.SharedHostBuffer(f"{name}_{i}", sh, dt, create=False) for i, (sh, dt) in enumerate(shapes)]
        P.sharded_infer_direct(_fake_infer, bufs[0].tensor, bufs[1].tensor, bufs[2].tensor, n_utt, torch.device("cpu"),
                               micro_batches=2)
        out3 = bufs[2].tensor.clone()
        dist.barrier()
        for bf in bufs:
            bf.close()
        if rank == 0:
            want = _fake_infer(mel, lengths)
            q.put(bool(torch.equal(out, want)) and bool(torch.equal(out2, want)) and out2 is host and bool(torch.equal(out3, want)))
        else:
            assert out is None and out2 is None
            assert torch.equal(out3, _fake_infer(mel, lengths))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_utt", [4, 5, 1])
def test_sharded_infer_gloo_world2(n_utt):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_utt, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_shard_bounds():
    import svk_parallel as P
    assert P.shard_bounds(512, 8) == [(64 * i, 64 * (i + 1)) for i in range(8)]
    assert P.shard_bounds(5, 2) == [(0, 3), (3, 5)]
    assert P.shard_bounds(1, 4) == [(0, 1), (1, 1), (1, 1), (1, 1)]
    for n in range(0, 20):
        for w in range(1, 9):
            b = P.shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))


# ---------------------------------------------------------------------------------- time-axis sharding (8(f) rank 3)
HALO = 5


def _fake_full(mel, lengths, eps):
    """Stand-in synthesiser with receptive field HALO frames per side and zero padding at the true ends."""
    B, C, T = mel.shape
    mask = (torch.arange(T)[None, :] < lengths[:, None]).float()[:, None, :]
    x = (mel.sum(1, keepdim=True) + eps.sum(1, keepdim=True)) * mask
    k = torch.arange(1, 2 * HALO + 2, dtype=torch.float32).view(1, 1, -1)
    y = torch.nn.functional.conv1d(x, k, padding=HALO) * mask
    return y.repeat_interleave(4, dim=2)


def _fake_window(mel_w, len_w, eps_w, lo, hi):
    return _fake_full(mel_w, len_w, eps_w)[:, :, 4 * lo:4 * hi].contiguous()


def _time_worker(rank, world, port, T, q):
    import sys
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    import svk_parallel as P
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, C, CL = 3, 4, 2
        g = torch.Generator().manual_seed(1)
        mel = torch.randn(B, C, T, generator=g)
        eps = torch.randn(B, CL, T, generator=g)
        lengths = torch.tensor([T, max(T - 7, 0), T // 2], dtype=torch.int64)
        out = P.time_sharded_infer(_fake_window, mel if rank == 0 else None, lengths if rank == 0 else None,
                                   eps if rank == 0 else None, B, C, CL, T, 4, HALO, torch.device("cpu"))
        if rank == 0:
            q.put(bool(torch.equal(out, _fake_full(mel, lengths, eps))))
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("T", [64, 33, 3, 1])
def test_time_sharded_infer_gloo_world2(T):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_time_worker, args=(r, 2, port, T, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_time_shard_bounds():
    import svk_parallel as P
    assert P.time_shard_bounds(1000, 2, 110) == [(0, 500, 0, 610), (500, 1000, 390, 1000)]
    assert P.time_shard_bounds(1, 2, 110) == [(0, 1, 0, 1), (1, 1, 1, 1)]
    for T in (1, 7, 300):
        for w in (1, 2, 8):
            b = P.time_shard_bounds(T, w, 110)
            assert b[0][0] == 0 and b[-1][1] == T
            assert all(x[2] <= x[0] and x[1] <= x[3] and x[2] >= 0 and x[3] <= T for x in b)


def test_length_buckets_cover_every_utterance_once():
    import svk_parallel as P
    lengths = [300, 32, 999, 301, 500, 1, 300, 700, 1001, 64]
    batches = P.length_buckets(lengths, [32, 300, 400, 500, 1000], batch_size=2)
    flat = [i for b in batches for i in b]
    assert sorted(flat) == list(range(len(lengths)))          # nothing dropped or duplicated (inference, not training)
    assert all(1 <= len(b) <= 2 for b in batches)
    bounds = [32, 300, 400, 500, 1000]
    grp = lambda n: sum(n > x for x in bounds)                # noqa: E731
    for b in batches:
        assert len({grp(lengths[i]) for i in b}) == 1         # one length group per batch (data_utils.py:133-135)
        assert [lengths[i] for i in b] == sorted((lengths[i] for i in b), reverse=True)
    assert P.length_buckets([], [10], 4) == []
    with pytest.raises(ValueError):
        P.length_buckets([1], [10], 0)
