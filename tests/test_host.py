"""Host-side logic: the nn.Module surface (state_dict schema, load errors, CPU refusal), the
utterance sharder, the time-chunk planner and — with the CPU oracle standing in for the CUDA
forward — the chunk/halo algebra and the world_size-2 gloo exchange.  CPU only."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import fixtures as fx
from oracle import torch_oracle
from oracle.common import max_abs
from tts_king_b200 import parallel
from tts_king_b200.hifi.vocoder.utils import get_padding

from _util import golden, make_generator, stored_state

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_state_dict_schema_v1():
    m = make_generator(fx.V1, fold=False)
    sd = m.state_dict()
    assert len(sd) == 234
    assert list(sd.keys())[:3] == ["conv_pre.bias", "conv_pre.weight_g", "conv_pre.weight_v"]
    assert tuple(sd["ups.0.weight_g"].shape) == (512, 1, 1)  # ConvTranspose: norm axis is C_in
    assert tuple(sd["ups.0.weight_v"].shape) == (512, 256, 16)
    assert tuple(sd["resblocks.11.convs2.2.weight_v"].shape) == (32, 32, 11)
    assert tuple(sd["conv_post.weight_v"].shape) == (1, 32, 7)
    m2 = make_generator(fx.V1, fold=True)
    assert len(m2.state_dict()) == 156
    assert "conv_pre.weight" in m2.state_dict()


def test_load_state_dict_both_layouts_and_strictness():
    g = golden("tiny_rb1")
    sd = stored_state(g)
    m = make_generator(fx.TINY_RB1, seed=99, fold=False)
    m.load_state_dict(sd)  # g/v layout
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd[k])
    mf = make_generator(fx.TINY_RB1, seed=99, fold=True)
    mf.load_state_dict(torch_oracle.fold_state_dict(sd))  # folded layout
    with pytest.raises(RuntimeError):
        mf.load_state_dict(sd)  # strict: g/v keys into a folded module
    bad = dict(sd)
    bad["conv_pre.bias"] = torch.zeros(7)
    with pytest.raises(RuntimeError):
        m.load_state_dict(bad)


def test_module_survives_deepcopy_and_pickle():
    """The reference's AttrDict (`self.__dict__ = self`) loses its attributes when copied; the module keeps
    the hyper-parameters it needs as plain values, so a copied / pickled generator still builds its plan."""
    import copy
    import pickle

    m = make_generator(fx.TINY_RB1)
    for clone in (copy.deepcopy(m), pickle.loads(pickle.dumps(m))):
        assert clone.hop_length == 256 and clone.halo_frames == 13
        c = clone._native_config()
        assert c.upsample_initial_channel == 32 and c.num_upsamples == 4 and c.resblock_type == 1
        assert list(c.upsample_rates)[:4] == [8, 8, 2, 2] and list(c.resblock_kernel_sizes)[:3] == [3, 7, 11]
        for (k1, v1), (k2, v2) in zip(m.state_dict().items(), clone.state_dict().items()):
            assert k1 == k2 and torch.equal(v1, v2)


def test_get_padding():
    assert [get_padding(k, d) for k in (3, 7, 11) for d in (1, 3, 5)] == [1, 3, 5, 3, 9, 15, 5, 15, 25]


def test_cpu_input_is_refused_not_emulated():
    m = make_generator(fx.TINY_RB1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 80, 4))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_hifiapi_refuses_without_cuda():
    from tts_king_b200.hifiapi import AttrDict, HIFIapi

    cfg = AttrDict(hifi=fx.make_h(fx.TINY_RB1), model_config=AttrDict(vocoder=AttrDict(use_cpu=True)))
    cfg.hifi["weights_path"] = None
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        HIFIapi(cfg, "cpu")


# ------------------------------------------------------------------ sharding / chunk planning
def test_halo_is_13_frames_for_v1():
    h = fx.make_h(fx.V1)
    assert parallel.receptive_reach_samples(h) == 3258
    assert parallel.halo_frames(h) == 13
    assert parallel.hop_length(h) == 256


def test_shard_utterances_balances_and_partitions():
    rng = np.random.default_rng(0)
    lengths = rng.integers(400, 2000, size=64).tolist()
    for world in (1, 2, 4, 8):
        shards = parallel.shard_utterances(lengths, world)
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(64))
        loads = [sum(lengths[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(lengths)
    assert parallel.shard_utterances([5, 5, 5], 8)[3:] == [[]] * 5
    assert parallel.shard_utterances([], 2) == [[], []]


def test_plan_time_chunks_covers_exactly():
    for T, parts, halo in ((310078, 8, 13), (100, 8, 13), (5, 8, 13), (1, 1, 13), (27, 2, 13)):
        chunks = parallel.plan_time_chunks(T, parts, halo)
        assert len(chunks) == parts
        assert chunks[0].start == 0 and chunks[-1].stop == T
        for a, b in zip(chunks, chunks[1:]):
            assert a.stop == b.start
        for c in chunks:
            assert 0 <= c.lo <= c.start <= c.stop <= c.hi <= T
            if c.frames:
                assert c.lo == max(0, c.start - halo) and c.hi == min(T, c.stop + halo)
    c8 = parallel.plan_time_chunks(310078, 8, 13)
    assert max(c.frames for c in c8) == 38760


def _oracle_fn(cfg, sd):
    return lambda mel: torch_oracle.forward(cfg, sd, mel)


def test_chunked_forward_equals_full_forward():
    """T6: chunk + 13-frame halo == monolithic forward (exact in exact arithmetic)."""
    cfg = fx.TINY_RB1
    sd = stored_state(golden("tiny_rb1"), "alive.")
    h = fx.make_h(cfg)
    halo, hop = parallel.halo_frames(h), parallel.hop_length(h)
    assert halo == 13
    mel = fx.synthetic_mel(1, 90, seed=3).double()
    sd64 = {k: v.double() for k, v in sd.items()}
    fn = _oracle_fn(cfg, sd64)
    full = fn(mel)
    for chunk_frames in (90, 45, 30, 17):
        y = parallel.chunked_forward(fn, mel, chunk_frames, halo, hop)
        assert y.shape == full.shape
        assert max_abs(y.numpy(), full.numpy()) <= 1e-13
    # a halo that is too short must show up (halo 8 leaves ~1e-6)
    y_bad = parallel.chunked_forward(fn, mel, 30, 4, hop)
    assert max_abs(y_bad.numpy(), full.numpy()) > 1e-9


def test_direct_store_chunking_index_algebra(monkeypatch):
    """chunked_forward_into / sharded_long_form_into: every chunk's window lands at its place in the final
    buffer (the oracle stands in for Generator.forward_into; ranks are emulated one after the other)."""
    cfg = fx.TINY_RB1
    sd = {k: v.double() for k, v in stored_state(golden("tiny_rb1"), "alive.").items()}

    class Gen:
        h = fx.make_h(cfg)
        hop_length = 256

        def forward_into(self, x, out, skip, keep, max_wav_value=None):
            out.copy_(torch_oracle.forward(cfg, sd, x)[:, :, skip * 256:(skip + keep) * 256])
            return out

    T = 70
    mel = fx.synthetic_mel(1, T, seed=3).double()
    full = torch_oracle.forward(cfg, sd, mel)
    out = torch.zeros_like(full)
    parallel.chunked_forward_into(Gen(), mel, 23, 13, out)
    assert max_abs(out.numpy(), full.numpy()) <= 1e-13
    out = torch.zeros_like(full)
    for a, b in ((0, 25), (25, 48), (48, 70)):
        lo, hi = max(0, a - 13), min(T, b + 13)
        monkeypatch.setattr(parallel, "exchange_halo", lambda local, halo, group=None, lo=lo, hi=hi, a=a, b=b: (mel[:, :, lo:hi], a - lo, hi - b))
        parallel.sharded_long_form_into(Gen(), mel[:, :, a:b], 13, out, a, chunk_frames=9)
    assert max_abs(out.numpy(), full.numpy()) <= 1e-13


# ------------------------------------------------------------------ ragged batches (N2)
def test_length_buckets_partition_and_cost():
    from tts_king_b200 import ragged

    rng = np.random.default_rng(5)
    for n in (1, 2, 7, 16, 33):
        frames = [int(v) for v in rng.integers(1, 900, size=n)]
        t_max = max(frames)
        buckets = ragged.plan_length_buckets(frames, 13, t_max)
        assert sorted(i for b in buckets for i in b) == list(range(n))  # a partition
        assert buckets == ragged.plan_length_buckets(frames, 13, t_max)  # deterministic
        cost = sum(len(b) * ragged.bucket_extent(frames, b, 13, t_max) + 400 for b in buckets)
        assert cost <= n * t_max + 400  # never worse than the padded batch
        for b in buckets:  # contiguous in sorted order: no bucket's range straddles another's
            lo, hi = min(frames[i] for i in b), max(frames[i] for i in b)
            assert all(not (lo < frames[j] < hi) for c in buckets if c is not b for j in c)
    assert ragged.plan_length_buckets([800] * 16, 13, 800) == [list(range(16))]  # nothing to skip
    assert ragged.plan_length_buckets([], 13, 800) == []
    assert len(ragged.plan_length_buckets([100, 800] * 4, 13, 800, max_buckets=1)) == 1
    assert len(ragged.plan_length_buckets([100, 800] * 4, 13, 800)) == 2
    with pytest.raises(ValueError):
        ragged.plan_length_buckets([3, 0], 13, 800)


class _OracleGenerator:
    """The CPU oracle behind the two attributes and the call ragged_generate uses."""

    def __init__(self, cfg, sd):
        self.cfg, self.sd, self.h = cfg, sd, fx.make_h(cfg)
        self.hop_length = parallel.hop_length(self.h)
        self.calls = []

    def __call__(self, mel):
        self.calls.append(tuple(mel.shape))
        return torch_oracle.forward(self.cfg, self.sd, mel)


def test_ragged_generate_equals_padded_forward_inside_valid_ranges():
    """N2: length buckets + 13-frame tail == the padded forward on every kept sample."""
    from tts_king_b200 import ragged

    cfg = fx.TINY_RB1
    sd = {k: v.double() for k, v in stored_state(golden("tiny_rb1"), "alive.").items()}
    gen = _OracleGenerator(cfg, sd)
    T = 96
    mel = fx.synthetic_mel(5, T, seed=11).double()  # padding frames deliberately NOT zero
    keep = [96 * 256, 20 * 256 + 17, 21 * 256, 1, 60 * 256 - 3]
    full = gen(mel)
    gen.calls.clear()
    parts = ragged.ragged_generate(gen, mel, keep, out_int16=False, launch_cost=8)
    assert len(gen.calls) >= 2 and sum(b * t for b, _, t in gen.calls) < 5 * T  # padding was skipped
    for i, n in enumerate(keep):
        assert parts[i].shape == (n,)
        assert max_abs(parts[i].numpy(), full[i, 0, :n].numpy()) <= 1e-13
    with pytest.raises(ValueError):
        ragged.ragged_generate(gen, mel, keep[:-1], out_int16=False)


def test_vocoder_infer_rejects_other_vocoders():
    from tts_king_b200.fs_two.utils.model import vocoder_infer

    with pytest.raises(NotImplementedError):
        vocoder_infer(torch.zeros(1, 80, 4), None, {"vocoder": {"model": "MelGAN"}},
                      {"preprocessing": {"audio": {"max_wav_value": 32768.0}}})


# ------------------------------------------------------------------ world_size-2 gloo
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    cfg = fx.TINY_RB1
    sd = {k: v.double() for k, v in stored_state(golden("tiny_rb1"), "alive.").items()}
    h = fx.make_h(cfg)
    halo, hop = parallel.halo_frames(h), parallel.hop_length(h)
    fn = _oracle_fn(cfg, sd)
    mel = fx.synthetic_mel(1, 61, seed=5).double()
    chunks = parallel.plan_time_chunks(mel.shape[-1], world, 0)
    local = mel[:, :, chunks[rank].start:chunks[rank].stop].contiguous()
    wav_local = parallel.sharded_long_form(fn, local, halo, hop)
    assert wav_local.shape[-1] == chunks[rank].frames * hop
    wav = parallel.gather_wav(wav_local, dst=0)
    # utterance sharding: every rank takes its own utterances, nothing is exchanged on the data path
    lengths = [20, 33, 14, 27, 9]
    mine = parallel.shard_utterances(lengths, world)[rank]
    outs = {i: fn(fx.synthetic_mel(1, lengths[i], seed=100 + i).double()) for i in mine}
    gathered = [None] * world
    dist.all_gather_object(gathered, {i: o.numpy() for i, o in outs.items()})
    if rank == 0:
        full = fn(mel)
        merged = {}
        for d in gathered:
            merged.update(d)
        ok_utt = sorted(merged) == list(range(len(lengths))) and all(
            np.array_equal(merged[i], fn(fx.synthetic_mel(1, lengths[i], seed=100 + i).double()).numpy())
            for i in merged)
        np.savez(out_path, err=max_abs(wav.numpy(), full.numpy()), n=wav.shape[-1], n_full=full.shape[-1], ok_utt=ok_utt)
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_halo_exchange_and_gather(tmp_path):
    """T7 on CPU: time-sharded 2-rank run (halo exchange + gather) == single-process run, and
    utterance sharding reassembles every utterance bit-identically."""
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = np.load(out)
    assert int(r["n"]) == int(r["n_full"])
    assert float(r["err"]) <= 1e-13
    assert bool(r["ok_utt"])
