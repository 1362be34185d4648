"""GPU tests of the whole-front-end entry points (dspfe_frontend / dspfe_frontend_host, csrc/frontend.cu): one call equals
the chain of the separate entry points bit for bit, slabbing does not change a bit, the host-buffer path equals the device
path, and the union of two ranks' LPT shards equals the single-GPU output bit for bit."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_batch(n=40, seed=5):
    from dspfe import synth
    lengths = synth.ragged_lengths(n, seed=seed, lo=6000, hi=50000)
    lengths[:4] = [1, 500, 8000, 80000]
    return synth.synth_batch(lengths, seed0=900 + seed)


def to_np(o, tot):
    rows, fc, fa = tot
    return dict(lr=o["lr"].cpu().numpy(), mfcc=o["mfcc"][:rows].cpu().numpy(), mfcc_frame_off=o["mfcc_frame_off"].cpu().numpy(),
                cep_pitch=o["cep_pitch"][:fc].cpu().numpy(), cep_lag=o["cep_lag"][:fc].cpu().numpy(),
                cep_frame_off=o["cep_frame_off"].cpu().numpy(), cep_feat=o["cep_feat"].cpu().numpy(),
                acr_pitch=o["acr_pitch"][:fa].cpu().numpy(), acr_lag=o["acr_lag"][:fa].cpu().numpy(),
                acr_frame_off=o["acr_frame_off"].cpu().numpy())


def assert_same(a, b, what):
    for k in a:
        np.testing.assert_array_equal(a[k], b[k], err_msg=f"{what}: {k}")


def run_device(pcm, off, **kw):
    import torch
    import dspfe
    dev = torch.device("cuda:0")
    fe = dspfe.FrontendPlan(**kw)
    o = fe.alloc(len(pcm), len(off) - 1, device=dev)
    tot = fe.run(torch.from_numpy(pcm).to(dev), torch.from_numpy(off).to(dev), off, o)
    torch.cuda.synchronize()
    return to_np(o, tot)


def test_frontend_equals_separate_entry_points():
    """model.py:52-95 / pitch_model.py:38-41 chained by dspfe_frontend == the same chain through dspfe_endpoint,
    dspfe_mfcc_delta, dspfe_pitch called one after the other."""
    import torch
    import dspfe
    pcm, off = make_batch()
    got = run_device(pcm, off)
    dev = torch.device("cuda:0")
    pcm_d, off_d = torch.from_numpy(pcm).to(dev), torch.from_numpy(off).to(dev)
    lr = dspfe.EndpointPlan().detect(pcm_d, off_d)
    out, fo = dspfe.MfccPlan(delta_n=2).mfcc_delta(pcm_d, off_d, trim=lr)
    cep = dspfe.PitchPlan(method=0, preemph=0.97).detect(pcm_d, off_d, trim=lr, want_feat=True)
    acr = dspfe.PitchPlan(method=1, frame_len=300).detect(pcm_d, off_d, trim=lr)
    torch.cuda.synchronize()
    fo = fo.cpu().numpy(); cfo = cep["frame_off"].cpu().numpy(); afo = acr["frame_off"].cpu().numpy()
    want = dict(lr=lr.cpu().numpy(), mfcc=out[: fo[-1]].cpu().numpy(), mfcc_frame_off=fo,
                cep_pitch=cep["pitch"][: cfo[-1]].cpu().numpy(), cep_lag=cep["lag"][: cfo[-1]].cpu().numpy(), cep_frame_off=cfo,
                cep_feat=cep["feat"].cpu().numpy(), acr_pitch=acr["pitch"][: afo[-1]].cpu().numpy(),
                acr_lag=acr["lag"][: afo[-1]].cpu().numpy(), acr_frame_off=afo)
    assert_same(got, want, "front-end vs separate entry points")


@pytest.mark.parametrize("slab", [20000, 100000, 700001])
def test_slabs_do_not_change_a_bit(slab):
    """The slab loop (bounded workspaces, BASELINE config 5) cuts the batch at arbitrary, unaligned utterance starts."""
    pcm, off = make_batch(seed=6)
    whole = run_device(pcm, off)
    assert_same(run_device(pcm, off, slab_samples=slab), whole, f"slab_samples={slab}")


def test_host_path_equals_device_path():
    import dspfe
    pcm, off = make_batch(seed=7)
    want = run_device(pcm, off)
    for hs in (30000, 0):
        fe = dspfe.FrontendPlan(host_slab_samples=hs)
        o = fe.alloc(len(pcm), len(off) - 1, device=None, pinned=True)
        tot = fe.run_host(pcm, off, o)
        assert_same(to_np(o, tot), want, f"host path, host_slab_samples={hs}")
        tot2 = fe.run_host(pcm, off, o)                     # buffers and events are reused
        assert tot2 == tot
        assert_same(to_np(o, tot2), want, "host path, second call")
    # offsets that do not start at zero: the batch is a window into a larger buffer
    fe = dspfe.FrontendPlan(host_slab_samples=50000)
    k = 5
    o = fe.alloc(len(pcm), len(off) - 1 - k, device=None)
    tot = fe.run_host(pcm, off[k:], o)
    sub = run_device(pcm[off[k]:].copy(), off[k:] - off[k])
    assert_same(to_np(o, tot), sub, "host path on a window of the buffer")


def test_timing_facility_names_every_kernel():
    import torch
    import dspfe
    pcm, off = make_batch(n=12, seed=8)
    dev = torch.device("cuda:0")
    fe = dspfe.FrontendPlan()
    o = fe.alloc(len(pcm), len(off) - 1, device=dev)
    pcm_d, off_d = torch.from_numpy(pcm).to(dev), torch.from_numpy(off).to(dev)
    fe.run(pcm_d, off_d, off, o)
    n0 = dspfe.launch_count()
    dspfe.timing_begin()
    fe.run(pcm_d, off_d, off, o)
    marks = dspfe.timing_end()
    names = [n for n, _ in marks]
    assert dspfe.launch_count() - n0 == len(marks)
    for k in ("ep_block_kernel", "mfcc_delta_kernel", "pitch_clip_kernel", "pitch_frame_kernel<0>", "pitch_frame_kernel<2>",
              "pitch_track_kernel", "pitch_feature_kernel", "fe_finish_kernel"):
        assert k in names, (k, names)
    assert all(ms >= 0 for _, ms in marks)


def test_two_rank_shard_union_equals_single_gpu(tmp_path):
    """Multi-GPU correctness: every rank processes its LPT shard of one global utterance list
    on its own device (ranks share cuda:0 when the box has a single GPU); the union of the per-rank MFCC rows, endpoints and
    pitch tracks, put back in global utterance order, equals the single-GPU output bit for bit."""
    import torch
    from dspfe import shard, synth
    U = 64
    lengths = synth.ragged_lengths(U, seed=77, lo=6000, hi=60000)
    utts = [synth.synth_utterance(7000 + u, int(n)) for u, n in enumerate(lengths)]
    pcm, off = shard.pack_shard(utts, np.arange(U))
    whole = run_device(pcm, off)
    world = 2
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r % max(torch.cuda.device_count(), 1)),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT="29617")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "rank_worker.py"), str(U), str(tmp_path)], env=env))
    for p in procs:
        assert p.wait(timeout=600) == 0
    parts = shard.lpt_partition(lengths, world)
    assert sorted(np.concatenate(parts).tolist()) == list(range(U))
    got = {k: [None] * U for k in ("lr", "mfcc", "cep_pitch", "cep_lag", "cep_feat", "acr_pitch", "acr_lag")}
    for r in range(world):
        d = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        for j, u in enumerate(parts[r]):
            got["lr"][u] = d["lr"][j]; got["cep_feat"][u] = d["cep_feat"][j]
            for key, fo in (("mfcc", "mfcc_frame_off"), ("cep_pitch", "cep_frame_off"), ("cep_lag", "cep_frame_off"),
                            ("acr_pitch", "acr_frame_off"), ("acr_lag", "acr_frame_off")):
                got[key][u] = d[key][d[fo][j]:d[fo][j + 1]]
    np.testing.assert_array_equal(np.stack(got["lr"]), whole["lr"])
    np.testing.assert_array_equal(np.stack(got["cep_feat"]), whole["cep_feat"])
    for key in ("mfcc", "cep_pitch", "cep_lag", "acr_pitch", "acr_lag"):
        np.testing.assert_array_equal(np.concatenate(got[key]), whole[key], err_msg=key)
    # the optional final gather (dspfe.shard.gather_rows over torch.distributed) put the MFCC rows in global order on every rank
    for r in range(world):
        d = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        np.testing.assert_array_equal(d["gathered_mfcc"], whole["mfcc"])
        np.testing.assert_array_equal(d["gathered_off"], whole["mfcc_frame_off"])
