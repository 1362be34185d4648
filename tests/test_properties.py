"""Property tests (hypothesis) of the host-side index arithmetic: framing, decimation, sharding.  CPU only."""
import numpy as np
from hypothesis import given, settings, strategies as st

import dspfe
from dspfe import shard
from oracle import ref_features as O


@settings(max_examples=200, deadline=None)
@given(n=st.integers(0, 200000), flen=st.integers(1, 2048), step=st.integers(1, 1024))
def test_num_frames_equals_framesig_rows(n, flen, step):
    want = 1 if n <= flen else 1 + int(np.ceil((n - flen) / step))          # sigproc.py:79-82
    assert dspfe.num_frames(n, flen, step) == want == O.num_frames(n, flen, step)
    # the padded length covers the signal, and one frame fewer would not
    assert (want - 1) * step + flen >= n
    assert want == 1 or (want - 2) * step + flen < n


@settings(max_examples=60, deadline=None)
@given(n=st.integers(1, 6000), rate=st.sampled_from([8000, 16000, 22050, 32000, 44100, 48000, 10000, 9000]))
def test_decimated_length_and_frames_match_the_reference_loop(n, rate):
    idx = O.downsample_indices(n, rate, 10000)
    nf, ld = dspfe.pitch_num_frames_host(n, samplerate=rate, frame_len=300, frame_step=100, method=1)
    assert ld == len(idx)
    assert nf == O.num_frames(len(idx), 300, 100)


@settings(max_examples=50, deadline=None)
@given(lengths=st.lists(st.integers(1, 80000), min_size=1, max_size=300), world=st.sampled_from([1, 2, 3, 4, 8]))
def test_lpt_partition_properties(lengths, world):
    ln = np.array(lengths)
    parts = shard.lpt_partition(ln, world)
    assert sorted(np.concatenate(parts).tolist()) == list(range(len(ln)))
    loads = np.array([ln[p].sum() for p in parts])
    assert loads.max() - loads.min() <= ln.max()                  # greedy LPT: within one utterance of each other


def test_unsupported_rate_ratio_is_reported():
    import pytest
    with pytest.raises(dspfe.DspfeError) as e:       # 11025 -> 10000 needs a 400-entry pattern (limit 256)
        dspfe.pitch_num_frames_host(1000, samplerate=11025)
    assert e.value.code == -2
