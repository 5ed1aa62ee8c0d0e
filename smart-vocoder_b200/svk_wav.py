"""PCM egress: the step right after the path in the reference's callers
(`inference.ipynb:118,124` plays `audio`; the demo files under `generated_files*/` are what it saved).

Those demo files are IEEE-float32 WAV: RIFF / `fmt ` chunk of 18 bytes (format tag 3, mono, 22 050 Hz,
cbSize 0) / `fact` chunk (sample count) / `data` -- the layout scipy.io.wavfile.write produces for
float32 input.  `write_wav_float32` writes exactly that layout from the waveform `infer` returns
(`tests/golden/demo_wav_headers.json` pins every header field against the reference's files).
No resampling, scaling or clipping: the reference saves the network output as is.
"""
from __future__ import annotations

import struct

import numpy as np

SAMPLE_RATE = 22050  # configs/iitp_base.json: data.sampling_rate


def wav_float32_bytes(pcm, sample_rate: int = SAMPLE_RATE) -> bytes:
    """pcm: 1-D (or [1, n] / [1, 1, n]) float array or CPU/GPU torch tensor -> complete WAV file as bytes."""
    if hasattr(pcm, "detach"):
        pcm = pcm.detach().float().cpu().numpy()
    a = np.ascontiguousarray(np.asarray(pcm, dtype="<f4").reshape(-1))
    n = a.size
    fmt = struct.pack("<4sIHHIIHHH", b"fmt ", 18, 3, 1, sample_rate, sample_rate * 4, 4, 32, 0)
    fact = struct.pack("<4sII", b"fact", 4, n)
    data_hdr = struct.pack("<4sI", b"data", n * 4)
    body = b"WAVE" + fmt + fact + data_hdr + a.tobytes()
    return b"RIFF" + struct.pack("<I", len(body)) + body


def write_wav_float32(path: str, pcm, sample_rate: int = SAMPLE_RATE) -> int:
    """Write one utterance; returns the number of samples written."""
    blob = wav_float32_bytes(pcm, sample_rate)
    with open(path, "wb") as f:
        f.write(blob)
    return (len(blob) - 58) // 4


write_wav = write_wav_float32  # the name INTEGRATION.md's table uses

MAX_WAV_VALUE = 32768.0  # configs/iitp_base.json:23


def to_int16(pcm, max_wav_value: float = MAX_WAV_VALUE):
    """Float waveform -> int16 samples: saturate(round_half_even(x * max_wav_value)), the inverse of the
    `audio / 32768.0` of inference.ipynb cell 4.  A CUDA tensor is converted on the device (svk_pcm_to_int16: the D2H
    copy then moves half the bytes) and stays there; numpy input is converted with numpy (same arithmetic)."""
    if hasattr(pcm, "is_cuda") and pcm.is_cuda:
        import torch
        import svk_runtime as rt
        x = pcm.detach().to(torch.float32).contiguous()
        out = torch.empty(x.shape, dtype=torch.int16, device=x.device)
        with torch.cuda.device(x.device):
            rt.check(rt.lib().svk_pcm_to_int16(x.data_ptr(), x.numel(), float(max_wav_value), out.data_ptr(),
                                               torch.cuda.current_stream(x.device).cuda_stream))
        return out
    if hasattr(pcm, "detach"):
        pcm = pcm.detach().float().cpu().numpy()
    a = np.asarray(pcm, dtype=np.float32) * np.float32(max_wav_value)
    return np.clip(np.rint(a), -32768, 32767).astype(np.int16)


def wav_int16_bytes(pcm, sample_rate: int = SAMPLE_RATE, max_wav_value: float = MAX_WAV_VALUE) -> bytes:
    """16-bit PCM WAV (format tag 1) of a float waveform or of int16 samples."""
    if hasattr(pcm, "detach"):
        pcm = pcm.detach().cpu().numpy()
    a = np.asarray(pcm)
    if a.dtype != np.int16:
        a = to_int16(a, max_wav_value)
    a = np.ascontiguousarray(a.reshape(-1).astype("<i2"))
    fmt = struct.pack("<4sIHHIIHH", b"fmt ", 16, 1, 1, sample_rate, sample_rate * 2, 2, 16)
    body = b"WAVE" + fmt + struct.pack("<4sI", b"data", a.size * 2) + a.tobytes()
    return b"RIFF" + struct.pack("<I", len(body)) + body


def write_wav_int16(path: str, pcm, sample_rate: int = SAMPLE_RATE) -> int:
    blob = wav_int16_bytes(pcm, sample_rate)
    with open(path, "wb") as f:
        f.write(blob)
    return (len(blob) - 44) // 2


def wav_header_info(blob: bytes) -> dict:
    """Parse the chunk structure of a WAV file (enough of it to compare with the reference's demo files)."""
    riff, _, wave = struct.unpack("<4sI4s", blob[:12])
    if riff != b"RIFF" or wave != b"WAVE":
        raise ValueError("not a RIFF/WAVE file")
    pos, info = 12, {"file_bytes": len(blob)}
    while pos + 8 <= len(blob):
        cid, csz = struct.unpack("<4sI", blob[pos:pos + 8])
        if cid == b"fmt ":
            tag, ch, sr, br, ba, bits = struct.unpack("<HHIIHH", blob[pos + 8:pos + 24])
            info.update(format_tag=tag, channels=ch, sample_rate=sr, byte_rate=br, block_align=ba, bits=bits,
                        fmt_chunk_size=csz)
        elif cid == b"data":
            info.update(data_bytes=csz, data_offset=pos + 8)
            break
        pos += 8 + csz + (csz & 1)
    return info
