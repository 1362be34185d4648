"""WAV ingest (reference reader.py:67-85): header parsing on the CPU, the device path on the GPU."""
import io
import os
import struct

import numpy as np
import pytest
from scipy.io import wavfile

import dspfe


def _wav_bytes(rate, data):
    b = io.BytesIO()
    wavfile.write(b, rate, data)
    return b.getvalue()


def test_wav_header_parser_against_scipy():
    rng = np.random.default_rng(0)
    for rate, ch, n in ((16000, 2, 1234), (44100, 1, 77), (48000, 2, 0), (8000, 6, 19)):
        data = rng.integers(-32768, 32767, size=(n, ch) if ch > 1 else (n,), dtype=np.int16)
        img = _wav_bytes(rate, data)
        info = dspfe.wav_info(img)
        assert (info["rate"], info["channels"], info["bits"], info["n_frames"]) == (rate, ch, 16, n)
        got = np.frombuffer(img, dtype=np.int16, offset=info["data_offset"], count=n * ch)
        np.testing.assert_array_equal(got, data.reshape(-1))
    # an extra chunk before 'data' (LIST), odd-sized with its pad byte
    data = rng.integers(-100, 100, size=(50, 2), dtype=np.int16)
    img = _wav_bytes(16000, data)
    extra = b"LIST" + struct.pack("<I", 5) + b"abcde" + bytes(1)
    img2 = img[:36] + extra + img[36:]
    img2 = img2[:4] + struct.pack("<I", len(img2) - 8) + img2[8:]
    info = dspfe.wav_info(img2)
    assert info["n_frames"] == 50 and info["data_offset"] == 44 + len(extra)
    with pytest.raises(dspfe.DspfeError):
        dspfe.wav_info(b"RIFFxxxxWAVEjunk")
    with pytest.raises(dspfe.DspfeError) as e:
        dspfe.wav_info(_wav_bytes(16000, rng.random(10).astype(np.float32)))
    assert e.value.code == -2


def test_scan_paths_counts_frames_beyond_the_header_probe(tmp_path):
    rng = np.random.default_rng(2)
    paths, want = [], []
    for i, (rate, ch, n) in enumerate(((16000, 2, 20000), (16000, 1, 333), (44100, 2, 1), (16000, 2, 0), (48000, 3, 4097), (8000, 1, 2026))):
        data = rng.integers(-32768, 32767, size=(n, ch) if ch > 1 else (n,), dtype=np.int16)
        p = os.path.join(tmp_path, f"s{i}.wav")
        wavfile.write(p, rate, data)
        paths.append(p); want.append((rate, n))
    off, rates = dspfe.wav_scan_paths(paths)
    assert list(np.diff(off)) == [w[1] for w in want] and list(rates) == [w[0] for w in want]
    with pytest.raises(dspfe.DspfeError):
        dspfe.wav_scan_paths([os.path.join(tmp_path, "missing.wav")])


@pytest.mark.gpu
def test_ingest_matches_reader_semantics(tmp_path):
    rng = np.random.default_rng(1)
    paths, want = [], []
    for i, (rate, ch, n) in enumerate(((16000, 2, 20000), (16000, 1, 333), (44100, 2, 1), (16000, 2, 0), (48000, 3, 4097))):
        data = rng.integers(-32768, 32767, size=(n, ch) if ch > 1 else (n,), dtype=np.int16)
        p = os.path.join(tmp_path, f"p{i:03d}-{i:02d}-01.wav")
        wavfile.write(p, rate, data)
        paths.append(p)
        r, sig = wavfile.read(p)                            # reader.py:76
        want.append((r, sig[:, 0] if sig.ndim == 2 else sig))   # reader.py:80 (the reference indexes [:,0]; mono files are taken as they are)
    pcm, off, rates = dspfe.ingest_wavs(paths)
    pcm = pcm.cpu().numpy()
    assert list(rates) == [w[0] for w in want]
    for i, (_, sig) in enumerate(want):
        np.testing.assert_array_equal(pcm[off[i]:off[i + 1]], sig)
    # the packed batch feeds the kernels directly
    import torch
    lr = dspfe.EndpointPlan().detect(torch.from_numpy(pcm).cuda(), torch.from_numpy(off).cuda())
    assert lr.shape == (5, 2)


def test_scan_paths_finds_a_data_chunk_behind_a_large_list_chunk(tmp_path):
    """A valid WAV whose `data` chunk sits beyond the first 4 KB (a large LIST chunk first): scipy.io.wavfile.read reads it
    (reader.py:76), so does the scanner (it walks the chunk headers of the file instead of probing a fixed prefix)."""
    import struct
    from scipy.io import wavfile
    import dspfe
    x = (np.arange(6000).reshape(-1, 2) % 977 - 400).astype(np.int16)
    fmt = struct.pack("<4sIHHIIHH", b"fmt ", 16, 1, 2, 16000, 16000 * 4, 4, 16)
    junk = struct.pack("<4sI", b"LIST", 10001) + b"\0" * 10001 + b"\0"        # odd length: one pad byte
    data = struct.pack("<4sI", b"data", x.nbytes) + x.tobytes()
    body = b"WAVE" + fmt + junk + data
    p = tmp_path / "late_data.wav"
    p.write_bytes(b"RIFF" + struct.pack("<I", len(body)) + body)
    rate, ref = wavfile.read(str(p))
    assert rate == 16000 and np.array_equal(ref, x)
    off, rates = dspfe.wav_scan_paths([str(p)])
    assert list(off) == [0, len(x)] and list(rates) == [16000]
