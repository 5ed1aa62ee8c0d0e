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
