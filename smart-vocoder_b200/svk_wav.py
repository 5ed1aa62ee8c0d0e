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
