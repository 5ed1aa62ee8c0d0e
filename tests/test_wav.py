"""PCM egress format: our float32 WAV writer against the header facts of the reference's
own demo files (`generated_files*/`, extracted once into tests/golden/demo_wav_headers.json)."""
import json
import os

import numpy as np

import svk_wav

HERE = os.path.dirname(os.path.abspath(__file__))


def test_wav_layout_matches_reference_demo_files(tmp_path):
    golden = json.load(open(os.path.join(HERE, "golden", "demo_wav_headers.json")))
    assert len(golden) == 10
    rng = np.random.default_rng(0)
    for name, g in golden.items():
        n = g["data_bytes"] // 4
        assert n % 256 == 0  # hop_length * frames: these are raw network outputs
        pcm = rng.standard_normal(n).astype(np.float32) * 0.1
        path = tmp_path / "o.wav"
        assert svk_wav.write_wav_float32(str(path), pcm) == n
        blob = path.read_bytes()
        assert svk_wav.wav_header_info(blob) == g, name
        assert np.array_equal(np.frombuffer(blob[g["data_offset"]:], "<f4"), pcm)  # samples untouched


def test_wav_accepts_infer_shaped_output():
    o = np.linspace(-1, 1, 512, dtype=np.float32).reshape(1, 1, 512)
    blob = svk_wav.wav_float32_bytes(o[0])
    info = svk_wav.wav_header_info(blob)
    assert info["data_bytes"] == 2048 and info["format_tag"] == 3 and info["sample_rate"] == 22050


def test_int16_wav_round_trip(tmp_path):
    """16-bit egress: saturating round-half-even of x * 32768 and a canonical 44-byte PCM header (readable by scipy)."""
    from scipy.io import wavfile
    x = np.array([0.0, 0.5, -0.5, 1.0, -1.0, 1.5, 3.0517578125e-05 * 0.5, 3.0517578125e-05 * 1.5, -2.0], np.float32)
    q = svk_wav.to_int16(x)
    assert q.tolist() == [0, 16384, -16384, 32767, -32768, 32767, 0, 2, -32768]
    path = tmp_path / "i16.wav"
    assert svk_wav.write_wav_int16(str(path), x) == x.size
    sr, data = wavfile.read(str(path))
    assert sr == 22050 and data.dtype == np.int16 and np.array_equal(data, q)
    info = svk_wav.wav_header_info(path.read_bytes())
    assert info["format_tag"] == 1 and info["bits"] == 16 and info["data_bytes"] == 2 * x.size
