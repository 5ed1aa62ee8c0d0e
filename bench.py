#!/usr/bin/env python
"""Benchmark of the B200-native SMART-Vocoder mel->waveform path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU PyTorch path (rank 0 only)
    python bench.py --impl torch-gpu --steps K ...            # the reference's modules on the same B200 through
                                                              # PyTorch eager + cuDNN, TF32 off and on

One "step" = one pass of SynthesizerTrn.infer (MelEncoder + inverse flow + HiFi-GAN decoder) over one
batch of synthetic 80 x T mels per GPU.  Workload at N=1: BASELINE.json configs[2] -- iitp_base.json,
batch 16, 80x1024 mel, fp32, full path (configs[1] is a single-kernel case and configs[0] the CPU
plumbing case; both are parity tests, not bench lines).  For N>1 every rank runs the same per-GPU batch
(weak scaling, utterance sharding, no data-path collective).

Prints ONE JSON line (rank 0).  `value` = whole-job audio samples/s with inputs resident in HBM;
`e2e` = the same through the reference-facing `SynthesizerTrn.infer` call from pinned HOST buffers
(H2D of mel+lengths and D2H of the PCM inside the timed region; NCCL scatter/gather for N>1);
`roofline` = the dominant kernel against the pipe it runs on; `cpu_baseline` = the UNMODIFIED reference
(oracle/_ref, vendored by oracle/vendor_ref.py; oracle/torch_port.py only if that is absent) on the box's host
cores.  Extra objects in the same line: `gpu_eager_baseline` (the reference on this GPU through PyTorch eager),
`latency_b1` (BASELINE configs[1] shape, one utterance), `c4_bf16` (configs[3]), `c5_shard` (configs[4]'s
64x1024 per GPU).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "smart-vocoder_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import svk_weights as W  # noqa: E402

SAMPLE_RATE = 22050
FLOP_PER_FRAME = 657_479_680       # 2 x 328,739,840 MACs, convolutions only (SURVEY 8(d) / BASELINE.md 3)
BYTES_MIN_PER_FRAME = 2112         # mel 320 + eps 768 + PCM 1024
WEIGHT_BYTES_FP32 = 142_603_776
METRIC = "audio_samples_per_sec_22.05kHz_mel80_to_wav"
UNIT = "samples/s"
NOISE_SCALE = 0.667                # inference.ipynb:118


def load_cfg():
    return json.load(open(os.path.join(ROOT, "configs", "iitp_base.json")))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        d["_source"] = "measured (MEASURED_PEAKS.json)"
        return d
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "sm_max_mhz": 1965.0,
            "_source": "fallback (B200_PROFILING.md)"}


def synth_inputs(B, T, seed=0):
    """mel ~ N(-5, 2) like real log-mels, full-length utterances."""
    g = torch.Generator().manual_seed(seed)
    mel = torch.randn(B, 80, T, generator=g) * 2.0 - 5.0
    lengths = torch.full((B,), T, dtype=torch.int64)
    return mel, lengths


# ----------------------------------------------------------------------------------- CPU / eager side
def seeded_state_dict(dims):
    return W.make_state_dict(dims, seed=1234)


def make_reference_runner(device="cpu"):
    """-> (fn(mel, lengths) -> o, kind, description).  The unmodified reference (oracle/_ref, vendored by
    oracle/vendor_ref.py) when present, else oracle/torch_port.py (its torch operators restated)."""
    from oracle import ref_loader
    cfg = load_cfg()
    dims = W.dims_from_model_kwargs(513, **cfg["model"])
    sd = seeded_state_dict(dims)
    net = ref_loader.build_reference_net(cfg["model"], sd, 513, cfg["train"]["segment_size"] // cfg["data"]["hop_length"],
                                         cfg["data"]["n_speakers"])
    if net is not None:
        net = net.to(device)

        def fn(mel, lengths):
            return net.infer(mel, lengths, noise_scale=NOISE_SCALE)[0]
        return fn, "reference", ("unmodified reference SynthesizerTrn.infer (models.py:331-339) imported from oracle/_ref "
                                 "(weight_norm recomputed per call, as the reference does)")
    from oracle import torch_port
    sdt = {k: torch.from_numpy(v).to(device) for k, v in sd.items()}

    def fn2(mel, lengths):
        return torch_port.infer(sdt, dims, mel, lengths, None, NOISE_SCALE, None)[0]
    return fn2, "port", "oracle/torch_port.py (the reference's torch operators restated; oracle/_ref not vendored)"


def cpu_reference_run(B, T, steps, warmup, threads=None):
    """Time the reference path on the host cores -> (seconds per step, kind, description)."""
    if threads:
        torch.set_num_threads(threads)
    fn, kind, desc = make_reference_runner("cpu")
    mel, lengths = synth_inputs(B, T)
    torch.manual_seed(1)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            o = fn(mel, lengths)
            dt = time.perf_counter() - t0
            assert o.shape == (B, 1, 256 * T)
            if i >= warmup:
                times.append(dt)
    return times, kind, desc


def run_reference_arm(args):
    """`--impl reference`: the reference's own CPU implementation of the path, all host threads, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    T, Bfull = args.frames, args.batch_per_gpu
    fn, kind, desc = make_reference_runner("cpu")
    W_, K = max(args.warmup, 1), args.steps
    # bounded sample: as many utterances of the per-GPU batch as keep the whole run near `--ref-budget-s` seconds
    # (utterances are independent; the reference is FASTER per sample at small batch -- SURVEY 6 -- so a sample flatters it)
    torch.manual_seed(1)
    mel1, len1 = synth_inputs(1, T)
    with torch.no_grad():
        fn(mel1, len1)  # first call: thread pool / primitive-cache start-up, not a step
        t0 = time.perf_counter()
        fn(mel1, len1)
        t1 = time.perf_counter() - t0
    Bs = int(max(1, min(Bfull, args.ref_budget_s / max(t1 * (K + W_), 1e-6))))
    mel, lengths = synth_inputs(Bs, T)
    times = []
    with torch.no_grad():
        for i in range(W_ + K):
            t0 = time.perf_counter()
            o = fn(mel, lengths)
            dt = time.perf_counter() - t0
            if i >= W_:
                times.append(dt)
    assert o.shape == (Bs, 1, 256 * T)
    total = sum(times)
    value = Bs * 256 * T * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": W_, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "rtf": (total / len(times)) / (Bs * 256 * T / SAMPLE_RATE),
        "config": {"workload": f"iitp_base.json full path, {Bfull}x80x{T} mel per GPU, fp32 (BASELINE configs[2]); the "
                               f"reference arm times {Bs} of the {Bfull} utterances per step (independent units, throughput "
                               f"scales linearly; the reference is faster per sample at small batch)",
                   "frames": T, "batch_per_gpu": Bfull, "sample_batch": Bs},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                         "sample": f"{desc}; {Bs}x80x{T} mel per step ({Bs}/{Bfull} of the step), mean of {len(times)} steps"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def gpu_eager_baseline(B, T, dev, steps=5, warmup=3):
    """The reference's modules on the SAME GPU through PyTorch eager + cuDNN (the kernel-level bar), TF32
    off (an fp32 result) and on (torch's default).  Device-resident inputs, CUDA events, L2 flushed."""
    fn, kind, desc = make_reference_runner(dev)
    mel, lengths = synth_inputs(B, T)
    mel, lengths = mel.to(dev), lengths.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {"kind": kind, "what": desc + f", .cuda(), {B}x80x{T}, PyTorch {torch.__version__} eager, cuDNN "
                                      f"{torch.backends.cudnn.version()}, cudnn.benchmark off (stock)", "unit": UNIT}
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        for tag, tf32 in (("tf32_off", False), ("tf32_on", True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            with torch.no_grad():
                for i in range(warmup):
                    flush.fill_(i)
                    o = fn(mel, lengths)
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(steps):
                    flush.fill_(i)
                    o = fn(mel, lengths)
                e1.record()
                torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / steps
            out[tag] = {"value": B * 256 * T / (ms / 1e3), "ms_per_step": ms, "finite": bool(torch.isfinite(o).all())}
            del o
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
    torch.cuda.empty_cache()
    return out


def run_torch_gpu_arm(args):
    """`--impl torch-gpu`: one JSON line for the reference on the GPU through PyTorch eager (rank 0, one GPU)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    if not torch.cuda.is_available():
        emit({"impl": "torch-gpu", "unavailable": "no CUDA device"})
        return 0
    dev = torch.device("cuda", 0)
    B, T = args.batch_per_gpu, args.frames
    g = gpu_eager_baseline(B, T, dev, steps=args.steps, warmup=max(args.warmup, 3))
    off = g["tf32_off"]
    emit({"impl": "torch-gpu", "metric": METRIC, "value": off["value"], "unit": UNIT, "n_gpus": 1, "steps": args.steps,
          "warmup": max(args.warmup, 3), "ms_per_step": off["ms_per_step"], "higher_is_better": True, "scaling": "weak",
          "vs_baseline": None, "dtype": "f32", "data": "synthetic",
          "config": {"workload": f"iitp_base.json full path, {B}x80x{T} mel, fp32 (TF32 off) through PyTorch eager + cuDNN"},
          "gpu_eager_baseline": g, "gpu_launches": 0})
    return 0


# ----------------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.proc, self.path = None, f"/tmp/svk_clocks_{os.getpid()}.csv"
        try:
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            ident = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", ident, f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def wait_started(self, timeout=8.0):
        """nvidia-smi's start-up (driver attach) stalls kernel launches for a few 100 ms: let it finish and
        deliver its first sample before the timed region begins."""
        if self.proc is None:
            return
        t0 = time.time()
        while time.time() - t0 < timeout:
            try:
                if os.path.getsize(self.path) > 0:
                    break
            except OSError:
                pass
            time.sleep(0.05)
        time.sleep(0.3)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, smax, power, reasons = [], [], [], set()
        for ln in open(self.path):
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[0])), smax.append(float(p[1])), power.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if sm:
            # median SM clock UNDER LOAD: the sampler also sees the idle stretch before the warm-up
            load = [c for c, w in zip(sm, power) if w >= 0.5 * max(power)] or sm
            out.update(sm_mhz=statistics.median(load), sm_max_mhz=max(smax), reasons=sorted(reasons), samples=len(sm),
                       samples_under_load=len(load), power_w_max=max(power))
        return out


# ----------------------------------------------------------------------------------- our arm
def aggregate_profile(records, steps):
    """Group per-launch records into kernel families (engine, layer kind, Cin, taps); pick the one with most time."""
    fam, by_layer, by_engine = {}, {}, {}
    for r in records:
        key = (r["engine"], r["layer"].startswith("resblock"), r["cin"], r["k"])
        f = fam.setdefault(key, {"ms": 0.0, "n": 0, "flops": 0.0, "bytes": 0.0, "layers": set()})
        for q in ("ms", "flops", "bytes"):
            f[q] += r[q]
        f["n"] += 1
        f["layers"].add(r["layer"])
        by_layer[r["layer"]] = by_layer.get(r["layer"], 0.0) + r["ms"]
        e = by_engine.setdefault(r["engine"], {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "n": 0})
        e["ms"] += r["ms"]
        e["flops"] += r["flops"]
        e["bytes"] += r["bytes"]
        e["n"] += 1
    total_ms = sum(f["ms"] for f in fam.values())
    top_key, top = max(fam.items(), key=lambda kv: kv[1]["ms"])
    return total_ms, top_key, top, by_layer, by_engine


def _timed_steps(fn, K, Wm, flush, barrier, dev, world):
    """Wm warm-up + K timed calls of fn() with the L2 flush in between, CUDA events, max over ranks -> ms per step."""
    import torch.distributed as dist
    for i in range(Wm):
        flush.fill_(i & 0xFF)
        fn()
    barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        flush.fill_(i & 0xFF)
        fn()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / K


def run_extras(args, net, dims, cfg, dev, world, rank, barrier, flush):
    """Measurements beside the headline, in the same process on the same box (VERDICT r1 'missing' 3/4, N2, N3)."""
    from models import SynthesizerTrn
    out = {}
    T = args.frames
    net.range_check = False
    # ---- BASELINE configs[4]: 64 utterances x 1024 frames per GPU (512 over 8 GPUs), device-resident, weak sharding
    B5 = args.c5_batch
    if B5 > 0:
        mel5, len5 = synth_inputs(B5, T, seed=100 + rank)
        mel5, len5 = mel5.to(dev), len5.to(dev)
        K5 = max(2, min(args.steps, 5))
        ms5 = _timed_steps(lambda: net.infer(mel5, len5, noise_scale=NOISE_SCALE), K5, 1, flush, barrier, dev, world)
        net.check_range()
        out["c5_shard"] = {"value": B5 * world * dims.hop * T / (ms5 / 1e3), "unit": UNIT, "ms_per_step": ms5, "steps": K5,
                           "warmup": 1, "workload": f"BASELINE configs[4] per-GPU shape: {B5}x80x{T} mel per GPU, "
                           f"{B5 * world} utterances over {world} GPU(s), fp32-class engine, inputs resident in HBM",
                           "workspace_bytes": int(net._handle.workspace_bytes(B5, T, T))}
        del mel5, len5
        if world > 1:
            # ... and end to end at that size: 64 x world utterances from rank 0's pinned host memory and back
            import svk_parallel as P
            N5 = B5 * world
            if rank == 0:
                mel5h = torch.cat([synth_inputs(B5, T, seed=100 + r)[0] for r in range(world)]).pin_memory()
                len5h = torch.full((N5,), T, dtype=torch.int64).pin_memory()
                pcm5h = torch.empty(N5, 1, dims.hop * T).pin_memory()
            fn5 = lambda m, l: net.infer(m, l, noise_scale=NOISE_SCALE)[0]  # noqa: E731
            times5 = []
            for i in range(1 + 2):
                barrier()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                P.sharded_infer(fn5, mel5h if rank == 0 else None, len5h if rank == 0 else None, N5, 80, T, dims.hop, dev,
                                micro_batches=max(args.micro_batches, 4), out_host=pcm5h if rank == 0 else None)
                torch.cuda.synchronize()
                barrier()
                if i >= 1:
                    times5.append(time.perf_counter() - t0)
            net.check_range()
            t5 = torch.tensor([min(times5)], device=dev, dtype=torch.float64)
            import torch.distributed as dist
            dist.all_reduce(t5, op=dist.ReduceOp.MAX)
            out["c5_shard"]["e2e"] = {"value": N5 * dims.hop * T / float(t5.item()), "unit": UNIT, "ms_per_step": 1e3 * float(t5.item()),
                                      "h2d_bytes_per_step": N5 * 80 * T * 4 + N5 * 8, "d2h_bytes_per_step": N5 * dims.hop * T * 4,
                                      "path": "sharded_infer, 4 micro-batches per rank, rank 0 pinned host in / out, best of 2"}
    if world > 1:
        net.range_check = True
        return out
    # ---- BASELINE configs[1] shape / the notebook's actual use: ONE utterance of 1024 frames, latency
    mel1, len1 = synth_inputs(1, T, seed=7)
    mel1, len1 = mel1.to(dev), len1.to(dev)
    lat = []
    for i in range(5 + 40):  # eager calls are bound by the host's launch rate at this size: enough calls for a stable median
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        net.infer(mel1, len1, noise_scale=NOISE_SCALE)
        e1.record()
        torch.cuda.synchronize()
        if i >= 5:
            lat.append(e0.elapsed_time(e1))
    net.check_range()
    out["latency_b1"] = {"median_ms": statistics.median(lat), "min_ms": min(lat), "max_ms": max(lat), "calls": len(lat),
                         "launches": int(net.last_launch_count()),
                         "rtf": statistics.median(lat) / 1e3 / (dims.hop * T / SAMPLE_RATE),
                         "workload": f"1x80x{T} mel (inference.ipynb:114-118), device-resident, one infer per measurement, "
                                     f"CUDA events around the call, L2 warm"}
    glat = []
    try:
        for i in range(3 + 20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            net.infer_graph(mel1, len1, noise_scale=NOISE_SCALE)
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                glat.append(e0.elapsed_time(e1))
        gr = next(iter(net._graphs.values()))
        out["latency_b1"].update({"cuda_graph_median_ms": statistics.median(glat), "cuda_graph_min_ms": min(glat),
                                  "cuda_graph_kernel_nodes": gr.kernel_nodes, "cuda_graph_programmatic_edges": gr.programmatic_edges,
                                  "cuda_graph_note": "SynthesizerTrn.infer_graph: svk_graph_launch replay + the three input "
                                                     "copies (mel, lengths, eps draw) into the graph's buffers"})
    except Exception as e:  # pragma: no cover
        out["latency_b1"]["cuda_graph_error"] = str(e)[:200]
    # ---- the reference's modules on this GPU through PyTorch eager + cuDNN (the kernel-level bar)
    try:
        out["gpu_eager_baseline"] = gpu_eager_baseline(args.batch_per_gpu, T, dev, steps=3, warmup=2)
    except Exception as e:  # pragma: no cover
        out["gpu_eager_baseline"] = {"error": str(e)[:300]}
    # ---- BASELINE configs[3]: 64 x 80x512, bf16 operands / fp32 accumulate
    if args.engine != "bf16" and args.c4_batch > 0:
        net16 = SynthesizerTrn(513, 32, n_speakers=cfg["data"]["n_speakers"], engine="bf16", range_check=False, **cfg["model"])
        net16.load_state_dict({k: torch.from_numpy(v) for k, v in seeded_state_dict(dims).items()})
        net16 = net16.cuda().eval()
        B4, T4 = args.c4_batch, 512
        mel4, len4 = synth_inputs(B4, T4, seed=9)
        mel4, len4 = mel4.to(dev), len4.to(dev)
        K4 = max(2, min(args.steps, 5))
        ms4 = _timed_steps(lambda: net16.infer(mel4, len4, noise_scale=NOISE_SCALE), K4, 2, flush, barrier, dev, world)
        out["c4_bf16"] = {"value": B4 * dims.hop * T4 / (ms4 / 1e3), "unit": UNIT, "ms_per_step": ms4, "steps": K4, "warmup": 2,
                          "dtype": "bf16", "workload": f"BASELINE configs[3]: {B4}x80x{T4} mel, bf16 operands (activation images "
                          f"and weights), one tcgen05 pass per conv, fp32 accumulate; tolerance stated in tests/test_gpu_parity.py",
                          "frac_of_bf16_ceiling": (FLOP_PER_FRAME * B4 * T4 / (ms4 / 1e3) / 1e12) / measured_peaks()["bf16_tflops_sustained"]}
        del net16, mel4, len4
        torch.cuda.empty_cache()
    net.range_check = True
    return out


def run_b200_arm(args):
    import torch.distributed as dist

    import svk_parallel as P
    from models import SynthesizerTrn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch ourselves the way the driver does
        port = os.environ.get("MASTER_PORT", "29577")
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", port, os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    if not torch.cuda.is_available():
        emit({"error": "no CUDA device: the B200 arm has no CPU fallback"})
        return 2
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    cfg = load_cfg()
    dims = W.dims_from_model_kwargs(513, **cfg["model"])
    net = SynthesizerTrn(513, cfg["train"]["segment_size"] // cfg["data"]["hop_length"], n_speakers=cfg["data"]["n_speakers"],
                         engine=args.engine, **cfg["model"])
    net.load_state_dict({k: torch.from_numpy(v) for k, v in W.make_state_dict(dims, seed=1234).items()})
    net = net.cuda().eval()
    net.range_check = False  # device-resident legs: keep infer() asynchronous, check once after the timed region

    B, T, K, Wm = args.batch_per_gpu, args.frames, args.steps, max(args.warmup, 3)
    mel_h, len_h = synth_inputs(B, T, seed=rank)
    mel_d, len_d = mel_h.to(dev), len_h.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    torch.manual_seed(1 + rank)
    samples_per_step_rank = B * dims.hop * T

    with torch.no_grad():
        # The clock sampler starts BEFORE the warm-up: nvidia-smi's cold start takes seconds on a fresh box, and a GPU
        # left idle that long drops its SM / memory clocks -- the warm-up steps must be the last thing before the timed
        # region, not the sampler's start-up.
        sampler = ClockSampler(local_rank) if rank == 0 and not os.environ.get("SVK_BENCH_NO_SAMPLER") else None  # (A/B of the sampler's own cost)
        if sampler:
            sampler.wait_started()
        barrier()
        for i in range(Wm):  # identical to a timed step (the L2 flush too: its first use loads a torch module)
            flush.fill_(i & 0xFF)
            o = net.infer(mel_d, len_d, noise_scale=NOISE_SCALE)[0]
        torch.cuda.synchronize()
        launches_per_step = net.last_launch_count()

        # ---------------- timed region: K steps, inputs resident in HBM ----------------
        def timed_pass(profile):
            if profile:
                net._handle.profile_begin(max_records=(launches_per_step + 8) * K)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            torch.cuda.synchronize()
            marks = [torch.cuda.Event(enable_timing=True) for _ in range(K)]  # diagnostic only: where a slow step sits
            ev0.record()
            for i in range(K):
                flush.fill_(i & 0xFF)  # evict L2 between steps
                out = net.infer(mel_d, len_d, noise_scale=NOISE_SCALE)[0]
                marks[i].record()
            ev1.record()
            torch.cuda.synchronize()
            barrier()
            recs = net._handle.profile_end() if profile else None
            ends = [ev0.elapsed_time(m) for m in marks]
            step_ms_list[:] = [round(b - a, 3) for a, b in zip([0.0] + ends[:-1], ends)]
            return ev0.elapsed_time(ev1), recs, out

        step_ms_list = []
        elapsed_ms, _, o = timed_pass(False)
        step_ms = list(step_ms_list)
        # the same K steps again with a CUDA-event pair around every launch (roofline numbers); kept out of
        # `value` because event pairs serialise the stream at every launch boundary
        profiled_ms, records, _ = timed_pass(True)
        clocks = sampler.stop() if sampler else None
        net.check_range()  # raises if any step produced a non-finite sample (svk_check_range)
        assert torch.isfinite(o).all()
        net.range_check = True  # the end-to-end leg is the call a user makes: check included

        t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())

        # ---------------- e2e: pinned host -> (scatter) -> infer -> (gather) -> pinned host ----------------
        Ke = max(1, min(K, args.e2e_steps))
        N_total = B * world
        if rank == 0:
            mel_all_h = torch.cat([synth_inputs(B, T, seed=r)[0] for r in range(world)]).pin_memory()
            len_all_h = torch.full((N_total,), T, dtype=torch.int64).pin_memory()
            pcm_h = torch.empty(N_total, 1, dims.hop * T).pin_memory()
        infer_fn = lambda m, l: net.infer(m, l, noise_scale=NOISE_SCALE)[0]  # noqa: E731

        def e2e_step(micro=1):
            if world == 1:
                m = mel_all_h.to(dev, non_blocking=True)
                l = len_all_h.to(dev, non_blocking=True)
                pcm_h.copy_(infer_fn(m, l), non_blocking=True)
                torch.cuda.synchronize()
            else:
                # rank 0: pinned host mel -> H2D -> NCCL scatter; every rank: infer; NCCL gather -> D2H pinned host.
                # micro > 1: the gather + D2H of one piece of a rank's shard runs under the kernels of the next.
                net.range_check = False
                P.sharded_infer(infer_fn, mel_all_h if rank == 0 else None, len_all_h if rank == 0 else None, N_total, 80, T,
                                dims.hop, dev, micro_batches=micro, out_host=pcm_h if rank == 0 else None)
                torch.cuda.synchronize()
                net.check_range()
                net.range_check = True

        def time_e2e(step_fn):
            for _ in range(3):  # warm: the caching allocator (new tensor sizes -> cudaMalloc stalls of 50-140 ms in the first
                step_fn()       # two or three calls, tools/shard1_check.py), NCCL channels, pinned-copy paths
            barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(Ke):
                step_fn()
            torch.cuda.synchronize()
            barrier()
            te = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            return float(te[0].item())  # wall clock bracketed by syncs: includes the host-side copies' latency

        e2e_serial_s = time_e2e(lambda: e2e_step(1))
        e2e_s, e2e_mode = e2e_serial_s, "serial"
        e2e_alt = {}
        if world > 1:
            t_mb = time_e2e(lambda: e2e_step(args.micro_batches))
            e2e_alt["nccl_micro_batched"] = {"value": N_total * dims.hop * T * Ke / t_mb, "ms_per_step": 1e3 * t_mb / Ke,
                                             "micro_batches": args.micro_batches}
            if t_mb < e2e_s:
                e2e_s, e2e_mode = t_mb, "nccl_micro_batched"
            # direct: every rank moves its own shard over its own PCIe link through host buffers shared by the rank
            # processes (svk_parallel.SharedHostBuffer); no NCCL on the data path, one barrier per step
            def all_ok(flag):
                t_ = torch.tensor([1 if flag else 0], device=dev, dtype=torch.int32)
                dist.all_reduce(t_, op=dist.ReduceOp.MIN)
                return bool(t_.item())

            tag = f"svk_bench_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}"
            shapes = (((N_total, 80, T), torch.float32), ((N_total,), torch.int64), ((N_total, 1, dims.hop * T), torch.float32))
            shm, good = [], True
            for phase in (0, 1):  # rank 0 creates, then the others attach; every rank learns whether all succeeded
                try:
                    if good and (rank == 0) == (phase == 0):
                        shm = [P.SharedHostBuffer(f"{tag}_{i}", sh, dt, create=rank == 0) for i, (sh, dt) in enumerate(shapes)]
                        if rank == 0:
                            shm[0].tensor.copy_(mel_all_h), shm[1].tensor.copy_(len_all_h)
                    mine = True
                except Exception as ex:  # pragma: no cover
                    mine, e2e_alt["direct_shared_host"] = False, {"error": str(ex)[:300]}
                good = all_ok(good and mine)
            if good:
                def direct_step():
                    net.range_check = False
                    P.sharded_infer_direct(infer_fn, shm[0].tensor, shm[1].tensor, shm[2].tensor, N_total, dev)
                    net.check_range()
                    net.range_check = True
                t_d = time_e2e(direct_step)
                ok_d = all_ok(bool(torch.isfinite(shm[2].tensor).all()) if rank == 0 else True)
                e2e_alt["direct_shared_host"] = {"value": N_total * dims.hop * T * Ke / t_d, "ms_per_step": 1e3 * t_d / Ke, "finite": ok_d,
                                                 "path": "every rank: H2D of its own shard from a host buffer shared by the rank processes "
                                                         "-> infer -> D2H of its PCM into the shared output (svk_parallel.sharded_infer_direct)"}
                if t_d < e2e_s and ok_d:
                    e2e_s, e2e_mode = t_d, "direct_shared_host"
            else:
                e2e_alt.setdefault("direct_shared_host", {"error": "shared host buffers unavailable on some rank"})
            barrier()
            for bf in shm:
                bf.close()

        # Pipelined end to end (N = 1): the same per-step copies, but the H2D of step i+1 and the D2H of step i-1 overlap
        # the kernels of step i (SynthesizerTrn.pipeline -> svk_pipeline_*).  eps is drawn on the
        # device (svk_randn), as torch.randn_like draws it on the device in the serial leg.
        if world == 1:
            pipe = net.pipeline(B, T, depth=2)
            mel_np, len_np = mel_all_h.numpy(), len_all_h.numpy()
            outs = [torch.empty(N_total, 1, dims.hop * T).pin_memory() for _ in range(2)]
            outs_np = [o_.numpy() for o_ in outs]
            tk = pipe.submit(mel_np, len_np, outs_np[0], seed=1, noise_scale=NOISE_SCALE)  # warm
            pipe.wait(tk)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            prev = None
            for i in range(Ke):
                tk = pipe.submit(mel_np, len_np, outs_np[i & 1], seed=1, noise_scale=NOISE_SCALE)
                if prev is not None:
                    pipe.wait(prev)
                prev = tk
            pipe.wait(prev)
            e2e_pipe_s = time.perf_counter() - t0
            assert np.isfinite(outs_np[(Ke - 1) & 1]).all()
            pipe.close()
            e2e_s, e2e_mode = e2e_pipe_s, "pipelined"

        # ---------------- extra measurements (same run, same box) ----------------
        extras = {}
        if not args.no_extras:
            extras = run_extras(args, net, dims, cfg, dev, world, rank, barrier, flush)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---------------- report ----------------
    peaks = measured_peaks()
    prop = torch.cuda.get_device_properties(dev)
    n_sm = prop.multi_processor_count
    sm_max_mhz = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz", 1965.0)
    ffma_peak_tf = n_sm * 128 * 2 * sm_max_mhz * 1e6 / 1e12
    total_samples = samples_per_step_rank * world * K
    value = total_samples / (elapsed_ms / 1e3)
    ms_per_step = elapsed_ms / K

    prof_total_ms, top_key, top, by_layer, by_engine = aggregate_profile(records, K)
    eng, _, top_cin, top_k = top_key
    avg_ms = top["ms"] / top["n"]
    achieved_tf = (top["flops"] / top["n"]) / (avg_ms / 1e3) / 1e12
    hbm_gbs = (top["bytes"] / top["n"]) / (avg_ms / 1e3) / 1e9
    traffic, traffic_src, traffic_alg = None, None, None
    tpath = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            ent = tj.get("families", {}).get(f"{eng}:cin{top_cin}:k{top_k}")
            if ent:
                traffic, traffic_src = ent.get("dram_bytes_per_launch"), tj.get("source")
                traffic_alg = ent.get("algorithmic_bytes_per_launch_average", ent.get("algorithmic_bytes_of_this_launch"))
        except Exception:
            traffic = None
    tc_peak = peaks["bf16_tflops_sustained"]
    if eng == "tc":
        bound, peak_tf = "tensor", tc_peak
        kname = f"conv_tc_kernel (tcgen05 kind::f16, 3-product fp16 split) Cin={top_cin} k={top_k}"
        peak_src = (f"{peaks['_source']}: cuBLAS bf16 sustained (kernel timed inside a long step); achieved counts USEFUL "
                    f"fp32-equivalent FLOPs, the tensor pipe executes 3x that (tensor_raw), so the scheme's ceiling is frac=1/3")
    else:
        bound, peak_tf = "tensor", tc_peak
        kname = f"conv_ffma_kernel Cin={top_cin} k={top_k}"
        peak_src = f"{peaks['_source']}; FFMA-pipe kernel, fp32 peak {ffma_peak_tf:.1f} TFLOP/s"
    step_tf = FLOP_PER_FRAME * B * T / (ms_per_step / 1e3) / 1e12
    roofline = {
        "bound": bound, "kernel": kname + f" ({'+'.join(sorted(top['layers']))})",
        "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
        "tensor_raw": {"achieved": 3 * achieved_tf, "frac": 3 * achieved_tf / peak_tf} if eng == "tc" else None,
        "peak_source": peak_src,
        "traffic": traffic, "traffic_source": traffic_src,
        "traffic_launch_algorithmic_bytes": traffic_alg,  # average over the family's launches as captured in the model
                                                          # (profiles/dominant_kernel_traffic.json lists each of them)
        "timing": "CUDA-event pair around every launch on the launching stream, second pass of the same K steps "
                  "(profiled_pass_ms_per_step); `value` comes from the first pass without per-launch events",
        "launches_in_group": top["n"], "avg_launch_ms": avg_ms, "share_of_step": top["ms"] / prof_total_ms,
        "algorithmic_flops_per_launch": top["flops"] / top["n"], "algorithmic_bytes_per_launch": top["bytes"] / top["n"],
        "hbm": {"achieved_gbs": hbm_gbs, "peak_gbs": peaks["hbm_gbs"], "frac": hbm_gbs / peaks["hbm_gbs"],
                "peak_source": peaks["_source"]},
        "whole_step": {"tflops": step_tf, "frac_of_tensor_peak": step_tf / tc_peak,
                       "frac_of_ffma_peak": step_tf / ffma_peak_tf,
                       "hbm_compulsory_frac": (BYTES_MIN_PER_FRAME * B * T + WEIGHT_BYTES_FP32) / (ms_per_step / 1e3) / 1e9 / peaks["hbm_gbs"],
                       "unfused_layer_traffic_gbs": sum(e["bytes"] for e in by_engine.values()) / K / (ms_per_step / 1e3) / 1e9,
                       "kernel_time_share_by_layer": {k: v / prof_total_ms for k, v in sorted(by_layer.items(), key=lambda kv: -kv[1])},
                       "kernel_time_share_by_engine": {k: v["ms"] / prof_total_ms for k, v in by_engine.items()},
                       "profiled_kernel_ms_per_step": prof_total_ms / K,
                       "profiled_pass_ms_per_step": profiled_ms / K,
                       "profiled_gap_ms_per_step": sum(r["gap_ms"] for r in records) / K},
    }

    # honest traffic accounting: what the launches move, split into what an fp32-only layer-by-layer
    # schedule would move too and what exists only because tensors are kept twice (fp32 + fp16 hi/lo operand image)
    alg_b = sum(r["bytes"] for r in records) / K
    dup_b = sum(r["dup_bytes"] for r in records) / K
    yard = 4_282_688 * B * T  # unfused fp32 conv-boundary traffic per frame, measured with hooks on the reference
    roofline["traffic_accounting"] = {
        "algorithmic_bytes_per_step": alg_b, "operand_image_duplicate_bytes_per_step": dup_b,
        "compulsory_bytes_per_step": alg_b - dup_b, "unfused_fp32_yardstick_bytes_per_step": yard,
        "ratio_to_yardstick": alg_b / yard, "hbm_floor_ms_per_step": alg_b / (peaks["hbm_gbs"] * 1e9) * 1e3,
        "note": "sum over launches of input + outputs + residual operands + weights; an operand image counts 4 B/element "
                "(two fp16 planes); `duplicate` = images written beside an fp32 copy of the same values + split_image launches"}
    # the most HBM-bound family beside the dominant tensor-bound one
    fam_hbm = {}
    for r in records:
        if r["engine"] != "tc" or r["ms"] <= 0:
            continue
        f = fam_hbm.setdefault((r["layer"], r["cin"], r["k"]), {"ms": 0.0, "bytes": 0.0, "n": 0})
        f["ms"] += r["ms"]; f["bytes"] += r["bytes"]; f["n"] += 1
    if fam_hbm:
        (hl, hc, hk), hf = max(fam_hbm.items(), key=lambda kv: kv[1]["bytes"] / kv[1]["ms"])
        hg = hf["bytes"] / hf["ms"] / 1e6
        roofline["most_hbm_bound_family"] = {"layer": hl, "cin": hc, "k": hk, "avg_launch_ms": hf["ms"] / hf["n"],
                                             "achieved_gbs": hg, "peak_gbs": peaks["hbm_gbs"], "frac": hg / peaks["hbm_gbs"]}

    if args.dump_profile:
        groups = {}
        for r in records:
            g = groups.setdefault((r["engine"], r["layer"], r["cin"], r["cout"], r["k"], r["dilation"], r["length"]),
                                  {"ms": 0.0, "n": 0, "flops": 0.0, "bytes": 0.0, "gap": 0.0})
            g["ms"], g["n"], g["flops"], g["bytes"] = g["ms"] + r["ms"], g["n"] + 1, g["flops"] + r["flops"], g["bytes"] + r["bytes"]
            g["gap"] += r["gap_ms"]
        with open(args.dump_profile, "w") as f:
            f.write("engine,layer,cin,cout,k,dil,length,launches_per_step,avg_ms,ms_per_step,useful_tflops,alg_gbs,avg_gap_ms\n")
            for key, g in sorted(groups.items(), key=lambda kv: -kv[1]["ms"]):
                avg = g["ms"] / g["n"]
                f.write(",".join(str(x) for x in key) + f",{g['n'] / K:.1f},{avg:.4f},{g['ms'] / K:.3f},"
                        f"{g['flops'] / g['n'] / avg / 1e9:.1f},{g['bytes'] / g['n'] / avg / 1e6:.0f},{g['gap'] / g['n']:.4f}\n")

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        times, kind, desc = cpu_reference_run(1, T, steps=args.cpu_steps, warmup=1, threads=cores)
        best = min(times)
        cpu_baseline = {"value": 256 * T / best, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                        "sample": f"{desc} on 1x80x{T} mel = 1/{B} of the step; "
                                  f"best of {len(times)} after 1 warm-up; mean {statistics.mean(times):.2f}s",
                        "rtf": best / (256 * T / SAMPLE_RATE)}

    h2d = (mel_all_h.numel() * 4 + len_all_h.numel() * 8)
    d2h = pcm_h.numel() * 4
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm, "step_ms": step_ms,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.engine == "bf16" else "f32", "data": "synthetic",
        "rtf": (ms_per_step / 1e3) / (samples_per_step_rank / SAMPLE_RATE),
        "config": {"workload": f"iitp_base.json full path (MelEncoder+flow^-1+Generator), {B}x80x{T} mel per GPU, "
                               + ("bf16 operands / fp32 accumulate (BASELINE configs[3] arithmetic)" if args.engine == "bf16" else
                                  f"fp32 (BASELINE configs[2]{'; configs[4] sharding, weak' if world > 1 else ''})"),
                   "engine": args.engine,
                   "batch_per_gpu": B, "global_batch": B * world, "frames": T, "noise_scale": NOISE_SCALE,
                   "weights": "seeded recipe svk_weights.make_state_dict(seed=1234), random init (no checkpoint exists)",
                   "sharding": "utterances, no data-path collective" if world > 1 else "none",
                   "l2": "256 MiB device write between steps (inside the timed region, <0.1 ms each); per-step working "
                         "set ~2.3 GB >> 126 MB L2"},
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "e2e": {"value": N_total * dims.hop * T * Ke / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": Ke, "ms_per_step": 1e3 * e2e_s / Ke, "mode": e2e_mode, "alternatives": e2e_alt,
                "serial": {"value": N_total * dims.hop * T * Ke / e2e_serial_s, "ms_per_step": 1e3 * e2e_serial_s / Ke,
                           "path": "pinned host -> H2D" + (" -> NCCL scatter" if world > 1 else "") + " -> SynthesizerTrn.infer (libsvk) -> "
                                   + ("NCCL gather -> " if world > 1 else "") + "D2H pinned host, one step at a time"},
                "path": ("SynthesizerTrn.pipeline (svk_pipeline_submit / _wait, depth 2): every step copies its mel + lengths from "
                         "pinned host memory and its PCM back; the copies of neighbouring steps overlap this step's kernels; eps "
                         "drawn on the device" if e2e_mode == "pipelined" else
                         "best of: svk_parallel.sharded_infer (rank 0 pinned host -> H2D -> NCCL scatter -> SynthesizerTrn.infer -> "
                         "NCCL gather -> D2H pinned host; `serial` = one piece per rank, `nccl_micro_batched` = pieces pipelined) and "
                         "sharded_infer_direct (`direct_shared_host`); `mode` says which, `alternatives` + `serial` list all"),
                "note": "timed with a host clock around the K steps (copies + kernels + waits)"},
        "gpu_launches": int(launches_per_step * K * world),
        "clocks": clocks,
        "device": prop.name,
    }
    line.update(extras)
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line goes to the process's original stdout; everything else any library prints
    (e.g. NCCL's version banner) was diverted to stderr in main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "torch-gpu"])
    ap.add_argument("--ref-budget-s", type=float, default=150.0, help="reference arm: wall-clock budget that sizes its per-step sample")
    ap.add_argument("--no-extras", action="store_true", help="skip gpu_eager_baseline / latency_b1 / c4_bf16 / c5_shard")
    ap.add_argument("--batch-per-gpu", type=int, default=16)
    ap.add_argument("--frames", type=int, default=1024)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--engine", default="tc", choices=["tc", "fp32", "bf16"],
                    help="tc = tcgen05 fp16x3 (fp32-class, the default and the headline), fp32 = FFMA, "
                         "bf16 = BASELINE configs[3] arithmetic (use with --batch-per-gpu 64 --frames 512)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--micro-batches", type=int, default=2, help="N>1 end-to-end leg: pieces per rank pipelined by sharded_infer")
    ap.add_argument("--c5-batch", type=int, default=64, help="utterances per GPU of the c5_shard extra (0 = skip)")
    ap.add_argument("--c4-batch", type=int, default=64, help="utterances of the c4_bf16 extra (0 = skip)")
    ap.add_argument("--dump-profile", default=None, help="write the per-layer launch table (CUDA-event times) to this file")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.impl == "torch-gpu":
        return run_torch_gpu_arm(args)
    return run_b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())
