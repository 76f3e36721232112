"""T3-T7 — end-to-end parity of the CUDA generator against the reference's own outputs
(tests/golden/, produced by tools/make_golden.py from /root/reference) and against the CPU oracle.
Needs a B200: run with `-m gpu` under gpurun.

Tolerances are the north_star's: fp32 path max-abs <= 1e-4, bf16 path SNR >= 40 dB.
"""
import numpy as np
import pytest
import torch

from oracle import fixtures as fx
from oracle import torch_oracle
from oracle.common import ac_snr_db, max_abs, snr_db
from tts_king_b200 import parallel

from _util import golden, make_generator, stored_state

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4      # north_star: fp32 max-abs error <= 1e-4
BF16_SNR_DB = 40.0   # north_star: bf16 path SNR >= 40 dB against the fp32 reference
FULL = [("v1", fx.V1), ("v2_narrow", fx.V2_NARROW), ("v3_rb2", fx.V3_RB2)]


def gpu_forward(m, mel):
    with torch.no_grad():
        y = m(mel.cuda())
    torch.cuda.synchronize()
    return y.cpu()


@pytest.mark.parametrize("prec", ["fp32", "fp32_ffma"])
@pytest.mark.parametrize("name,cfg", FULL)
def test_fp32_paths_match_reference(name, cfg, prec):
    """T3: same random-init weights (seed 1234, digest-checked), same synthetic mels."""
    g = golden(name + "_seed1234")
    m = make_generator(cfg, precision=prec)
    assert fx.state_digest(m.state_dict()) == str(g["digest_folded"])
    m.cuda()
    ya = gpu_forward(m, torch.from_numpy(g["mel_a"]))
    assert ya.shape == g["y_a"].shape
    assert max_abs(ya.numpy(), g["y_a"]) <= FP32_TOL
    # ragged batch of log-mel-like values, passed as the non-contiguous transpose of a time-major
    # tensor exactly as tts_king.py:48 does
    mel_b = torch.from_numpy(g["mel_b"]).transpose(1, 2).contiguous().cuda().transpose(1, 2)
    assert not mel_b.is_contiguous()
    yb = gpu_forward(m, mel_b)
    assert max_abs(yb.numpy(), g["y_b"]) <= FP32_TOL


@pytest.mark.parametrize("name,cfg", FULL)
def test_fp32_alive_weights(name, cfg):
    """T3b: trained-like gains expose rounding that the attenuating default init hides
    (single-pass TF32 fails here)."""
    g = golden(name + "_seed1234")
    m = make_generator(cfg, precision="fp32")
    m.load_state_dict(fx.alive_state(cfg))
    m.cuda()
    y = gpu_forward(m, torch.from_numpy(g["mel_a"]))
    assert max_abs(y.numpy(), g["y_alive_a"]) <= FP32_TOL


@pytest.mark.parametrize("name,cfg", FULL)
def test_bf16_snr(name, cfg):
    """T4: bf16 operands, fp32 accumulate + residual stream."""
    g = golden(name + "_seed1234")
    m = make_generator(cfg, precision="bf16").cuda()
    y = gpu_forward(m, torch.from_numpy(g["mel_a"])).numpy()
    assert snr_db(g["y_a"], y) >= BF16_SNR_DB, (snr_db(g["y_a"], y), ac_snr_db(g["y_a"], y))
    m.load_state_dict(fx.alive_state(cfg))
    ya = gpu_forward(m, torch.from_numpy(g["mel_a"])).numpy()
    assert snr_db(g["y_alive_a"], ya) >= BF16_SNR_DB, snr_db(g["y_alive_a"], ya)


@pytest.mark.parametrize("name,cfg", [("tiny_rb1", fx.TINY_RB1), ("tiny_rb2", fx.TINY_RB2)])
def test_stored_weights_and_edge_shapes(name, cfg):
    """T5: load_state_dict of a reference-produced g/v state_dict; T = 1; unbatched input."""
    g = golden(name)
    m = make_generator(cfg, seed=5, fold=False)
    m.load_state_dict(stored_state(g))
    m.cuda()
    y = gpu_forward(m, torch.from_numpy(g["mel"]))
    assert max_abs(y.numpy(), g["y"]) <= FP32_TOL
    y1 = gpu_forward(m, torch.from_numpy(g["mel_T1"]))
    assert y1.shape == g["y_T1"].shape and max_abs(y1.numpy(), g["y_T1"]) <= FP32_TOL
    yu = gpu_forward(m, torch.from_numpy(g["mel"][0]))  # [80,T] -> [1,T*hop]
    assert yu.shape == (1, g["y"].shape[-1]) and max_abs(yu.numpy(), g["y"][0]) <= FP32_TOL
    # folded-layout load and the int16 tail of HIFIapi.generate on a full-scale signal
    mf = make_generator(cfg, seed=6, fold=True)
    mf.load_state_dict(stored_state(g, "alive."))
    mf.cuda()
    ya = gpu_forward(mf, torch.from_numpy(g["mel"]))
    assert max_abs(ya.numpy(), g["y_alive"]) <= FP32_TOL
    i16 = mf.generate_int16(torch.from_numpy(g["mel"]).cuda()).cpu().numpy()
    assert i16.dtype == np.int16 and i16.shape == g["y_alive_int16"].shape
    assert np.abs(i16.astype(np.int32) - g["y_alive_int16"].astype(np.int32)).max() <= 4  # 1e-4 * 32768 = 3.3 LSB


def test_wrong_channel_count_raises():
    m = make_generator(fx.TINY_RB1).cuda()
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 79, 5, device="cuda"))
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 80, 5))  # CPU tensor into a CUDA module


def test_hifiapi_drop_in():
    """T5: the wrapper end to end — generate() int16 and __call__ float against the reference's."""
    from tts_king_b200.hifiapi import AttrDict, HIFIapi

    g = golden("hifiapi_v1")
    cfg = AttrDict(hifi=fx.make_h(fx.V1), model_config=AttrDict(vocoder=AttrDict(use_cpu=True)))
    cfg.hifi["weights_path"] = None
    torch.manual_seed(1234)
    api = HIFIapi(cfg, "cpu")
    mel = torch.from_numpy(g["mel"])
    wav = api.generate(mel)
    assert isinstance(wav, np.ndarray) and wav.dtype == np.int16 and wav.shape == g["generate_int16"].shape
    assert np.abs(wav.astype(np.int32) - g["generate_int16"].astype(np.int32)).max() <= 4
    y = api(mel)
    assert y.device.type == "cpu" and max_abs(y.numpy(), g["call_f32"]) <= FP32_TOL


def test_checkpoint_file_roundtrip(tmp_path):
    """hifiapi.py:20-22: torch.load(path)["generator"] in the g/v layout."""
    from tts_king_b200.hifiapi import AttrDict, HIFIapi

    g = golden("tiny_rb1")
    path = str(tmp_path / "hifi.pth")
    torch.save({"generator": stored_state(g)}, path)
    cfg = AttrDict(hifi=fx.make_h(fx.TINY_RB1), model_config=AttrDict(vocoder=AttrDict(use_cpu=False)))
    cfg.hifi["weights_path"] = path
    api = HIFIapi(cfg, "cuda")
    y = api(torch.from_numpy(g["mel"]))
    assert y.is_cuda and max_abs(y.cpu().numpy(), g["y"]) <= FP32_TOL


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_chunked_equals_full_on_gpu(prec):
    """T6: 13-frame halo chunking reproduces the monolithic forward."""
    m = make_generator(fx.V1, precision=prec).cuda()
    h = fx.make_h(fx.V1)
    halo, hop = parallel.halo_frames(h), parallel.hop_length(h)
    mel = fx.synthetic_mel(1, 150, seed=9).cuda()
    with torch.no_grad():
        full = m(mel)
        y = parallel.chunked_forward(m, mel, 40, halo, hop)
    assert y.shape == full.shape
    assert max_abs(y.cpu().numpy(), full.cpu().numpy()) <= 1e-6


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_batch_items_are_independent(prec):
    """T7: utterance sharding is exact — an item's result does not depend on its batch."""
    m = make_generator(fx.V1, precision=prec).cuda()
    mel = fx.synthetic_mel(3, 64, seed=21).cuda()
    with torch.no_grad():
        yb = m(mel)
        ys = torch.cat([m(mel[i:i + 1]) for i in range(3)], dim=0)
    assert torch.equal(yb, ys)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_ragged_batch_is_bit_identical_inside_valid_ranges(prec):
    """N2: length buckets skip the padding the reference computes (fs_two/utils/tools.py:257-268) and
    change nothing on the samples vocoder_infer keeps."""
    from tts_king_b200 import ragged

    m = make_generator(fx.V1, precision=prec).cuda()
    T = 120
    mel = fx.synthetic_mel(6, T, seed=31).cuda()  # padding frames are not zero: they must not matter
    keep = [T * 256, 37 * 256 + 5, 38 * 256, 1, 90 * 256 - 1, 64 * 256]
    assert m._get_engine().halo_frames() == parallel.halo_frames(m.h) == 13
    with torch.no_grad():
        full = m(mel)
        full16 = m.generate_int16(mel)
        for mode in ("kernel", "buckets"):
            parts = ragged.ragged_generate(m, mel, keep, out_int16=False, launch_cost=8, mode=mode)
            parts16 = ragged.ragged_generate(m, mel, keep, out_int16=True, launch_cost=8, mode=mode)
            for i, n in enumerate(keep):
                assert torch.equal(parts[i], full[i, 0, :n]), (mode, i)
                assert torch.equal(parts16[i], full16[i, 0, :n]), (mode, i)
        # the kernel path on poisoned scratch: whatever the skipped rows hold must not leak into kept samples
        frames = [-(-n // 256) for n in keep]
        m(torch.full_like(mel, float("nan")))  # leaves NaN in every workspace buffer
        y = m.forward_ragged(mel, frames)
        for i, n in enumerate(keep):
            assert torch.equal(y[i, 0, :n], full[i, 0, :n]), i
            assert (y[i, 0, frames[i] * 256:] == 0).all(), i  # the skipped tail reads as silence, not scratch
    with pytest.raises(ValueError):
        m.forward_ragged(mel, frames[:-1])
    with pytest.raises(RuntimeError):
        m.forward_ragged(mel, [0] + frames[1:])  # frames must be in [1, T]


def test_ragged_batches_larger_than_the_kernel_limit_are_split():
    """hg_forward_ragged takes at most 64 utterances (the tile table travels in the kernel parameters);
    ragged_generate splits larger batches into groups of similar length."""
    from tts_king_b200 import ragged

    m = make_generator(fx.V1, precision="bf16").cuda()
    B, T = 70, 24
    mel = fx.synthetic_mel(B, T, seed=77).cuda()
    keep = [((7 * i) % T + 1) * 256 - (i % 3) for i in range(B)]
    with torch.no_grad():
        full = m(mel)
        parts = ragged.ragged_generate(m, mel, keep, out_int16=False)
        with pytest.raises(ValueError):
            m.forward_ragged(mel, [T] * B)
    for i, n in enumerate(keep):
        assert torch.equal(parts[i], full[i, 0, :n]), i


@pytest.mark.parametrize("name,cfg", [("v2_narrow", fx.V2_NARROW), ("v3_rb2", fx.V3_RB2)])
@pytest.mark.parametrize("prec", ["fp32", "bf16", "fp32_ffma"])
def test_ragged_other_configs_and_paths(name, cfg, prec):
    """The compacted tile space goes through every kernel family (narrow, FFMA, non-fused tcgen05)."""
    m = make_generator(cfg, precision=prec).cuda()
    hop = m.hop_length
    mel = fx.synthetic_mel(4, 70, seed=17).cuda()
    frames = [70, 9, 33, 1]
    with torch.no_grad():
        full = m(mel)
        y = m.forward_ragged(mel, frames)
    for i, f in enumerate(frames):
        assert torch.equal(y[i, 0, :f * hop], full[i, 0, :f * hop]), i


def test_vocoder_infer_drop_in():
    """N1/N2: fs_two/utils/model.py:85-100 — list of int16 arrays, trimmed to `lengths`."""
    from tts_king_b200.fs_two.utils.model import vocoder_infer

    mc = {"vocoder": {"model": "HiFi-GAN"}}
    pc = {"preprocessing": {"audio": {"max_wav_value": 32768.0}}}
    # against the reference's own (forward * 32768).astype(int16) (tools/make_golden.py)
    g = golden("tiny_rb1")
    m = make_generator(fx.TINY_RB1, seed=6, fold=True)
    m.load_state_dict(stored_state(g, "alive."))
    m.cuda()
    ref16 = g["y_alive_int16"]
    wavs = vocoder_infer(torch.from_numpy(g["mel"]), m, mc, pc)
    assert isinstance(wavs, list) and len(wavs) == ref16.shape[0]
    for i, w in enumerate(wavs):
        assert w.dtype == np.int16 and w.shape == ref16[i, 0].shape
        assert np.abs(w.astype(np.int32) - ref16[i, 0].astype(np.int32)).max() <= 4  # 1e-4 * 32768 = 3.3 LSB
    # with lengths: same samples, padding not computed
    m = make_generator(fx.V1).cuda()
    mel = fx.synthetic_mel(5, 100, seed=41)
    wavs = vocoder_infer(mel, m, mc, pc)
    lengths = torch.tensor([100 * 256, 31 * 256 + 9, 1, 70 * 256, 30 * 256])
    trimmed = vocoder_infer(mel, m, mc, pc, lengths=lengths)
    assert len(trimmed) == 5
    for i, w in enumerate(trimmed):
        n = int(lengths[i])
        assert w.dtype == np.int16 and w.shape == (n,)
        assert np.array_equal(w, wavs[i][:n])  # bit-identical to the padded run


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_forward_into_writes_only_the_window(prec):
    """hg_forward_window: the last kernel stores the owned samples of a haloed chunk in place."""
    m = make_generator(fx.V1, precision=prec).cuda()
    mel = fx.synthetic_mel(2, 50, seed=13).cuda()
    with torch.no_grad():
        full = m(mel)
        full16 = m.generate_int16(mel)
        big = torch.full((2, 1, 40 * 256 + 77), 7.0, device="cuda")          # a window of a larger buffer
        m.forward_into(mel, big[:, :, 11:11 + 30 * 256], skip_frames=9, keep_frames=30)
        assert torch.equal(big[:, :, 11:11 + 30 * 256], full[:, :, 9 * 256:39 * 256])
        assert (big[:, :, :11] == 7.0).all() and (big[:, :, 11 + 30 * 256:] == 7.0).all()  # nothing outside it
        o16 = torch.zeros((2, 1, 50 * 256), dtype=torch.int16, device="cuda")
        m.forward_into(mel, o16)
        assert torch.equal(o16, full16)
        # chunked long-form without the concatenation
        out = torch.empty_like(full)
        parallel.chunked_forward_into(m, mel, 17, parallel.halo_frames(m.h), out)
        assert torch.equal(out, full)
    with pytest.raises(ValueError):
        m.forward_into(mel, torch.empty((2, 1, 10 * 256), device="cuda"), skip_frames=45, keep_frames=10)
    with pytest.raises(ValueError):
        m.forward_into(mel, torch.empty((2, 1, 10 * 256 + 1), device="cuda"), skip_frames=0, keep_frames=10)


def test_profile_reports_what_each_launch_ran():
    m = make_generator(fx.V1, precision="bf16").cuda()
    # a short input runs the latency schedule: 256-channel convs in 64-column N tiles on the single-CTA kernel
    small = {r["name"]: r for r in m.profile_layers(fx.synthetic_mel(1, 64, seed=5).cuda())}
    assert small["resblocks.2.convs1.0"]["kernel"] == "tcgen05" and small["resblocks.2.convs1.0"]["n_tile"] == 64
    mel = fx.synthetic_mel(8, 200, seed=5).cuda()
    rows = m.profile_layers(mel)
    assert len(rows) == m.kernel_launches(8, 200) == 61
    assert all(r["ms"] > 0 for r in rows)
    by_name = {r["name"]: r for r in rows}
    assert by_name["mel_to_operand"]["kernel"] == "repack" and by_name["conv_post"]["kernel"] == "conv_post"
    assert by_name["resblocks.2.convs1.0"]["kernel"] == "tcgen05 cta_group::2"      # 256 ch, k = 11
    assert by_name["resblocks.3.convs1.0"]["kernel"] == "tcgen05"                   # 128 ch, k = 3: resident weights
    assert by_name["resblocks.8.convs2.1"]["kernel"] == "tcgen05 fused pair" and "resblocks.8.convs1.1" not in by_name
    m.precision = "fp32_ffma"
    assert {r["kernel"] for r in m.profile_layers(mel)} == {"repack", "cuda-core", "conv_post"}


def test_full_size_cross_check_against_ffma():
    """BASELINE cfg-2 scale (16 x 800 frames): the tensor-core fp32 path against the exact-fp32
    CUDA-core path on the device (the CPU oracle would take minutes), plus bf16 SNR."""
    mel = fx.synthetic_mel(16, 800, seed=7).cuda()
    m = make_generator(fx.V1, precision="fp32_ffma").cuda()
    with torch.no_grad():
        ref = m(mel).cpu().numpy()
        m.precision = "fp32"
        y32 = m(mel).cpu().numpy()
        m.precision = "bf16"
        y16 = m(mel).cpu().numpy()
    assert max_abs(y32, ref) <= FP32_TOL
    assert snr_db(ref, y16) >= BF16_SNR_DB
    # one utterance of that batch against the CPU oracle (the reference's arithmetic)
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    cpu = torch_oracle.forward(fx.V1, sd, mel[3:4, :, :200].cpu())
    with torch.no_grad():
        m.precision = "fp32"
        part = m(mel[3:4, :, :200]).cpu()
    assert max_abs(part.numpy(), cpu.numpy()) <= FP32_TOL


def test_cuda_graph_replay_matches_eager():
    """Graph-captured forward (the cfg-1 latency path): bit-identical to the eager launch sequence,
    also for a non-contiguous input and after the weights' workspace has been reused."""
    g = golden("v1_seed1234")
    m = make_generator(fx.V1, precision="fp32").cuda()
    mel = torch.from_numpy(g["mel_a"]).cuda()
    with torch.no_grad():
        eager = m(mel).clone()
    run = m.make_graphed(1, mel.shape[-1])
    y1 = run(mel).clone()
    y2 = run(mel.transpose(1, 2).contiguous().transpose(1, 2)).clone()
    with torch.no_grad():
        m(fx.synthetic_mel(2, 50, seed=1).cuda())  # an unrelated eager call in between
    y3 = run(mel).clone()
    assert torch.equal(y1, eager) and torch.equal(y2, eager) and torch.equal(y3, eager)
    assert max_abs(y1.cpu().numpy(), g["y_a"]) <= FP32_TOL
    with pytest.raises(RuntimeError):
        run(torch.zeros(1, 80, 7, device="cuda"))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (gpurun --gpus 2)")
def test_multi_gpu_sharding_matches_single_gpu():
    """T7 on real GPUs: utterance sharding is bitwise exact, the NCCL halo exchange reproduces the
    monolithic forward (tools/multi_gpu_check.py, one process per GPU)."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    n = min(torch.cuda.device_count(), 8)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29531",
                        os.path.join(root, "tools", "multi_gpu_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    for prec in ("fp32", "bf16"):
        assert res[prec]["utterance_sharding_bitwise_equal"]
        assert res[prec]["long_form_max_abs_vs_single_gpu"] <= 1e-6
        assert res[prec]["direct_p2p_store_bitwise_equal"]  # gather fused into conv_post's stores (NVLink P2P)


@pytest.mark.parametrize("env", [{"HG_TC2": "0"}, {"HG_FUSE_PAIRS": "0"}, {"HG_EPI_TMA": "0"}, {"HG_FOLD": "0"}, {"HG_FOLD": "2"}, {"HG_PAD_NARROW": "0"}, {"HG_CHAIN": "1"},
                                 {"HG_TC2": "0", "HG_FUSE_PAIRS": "0", "HG_EPI_TMA": "0"}, {"HG_FORCE_FFMA": "1"},
                                 {"HG_EPI_TMA_CONVT": "0"}, {"HG_TC2_CONVT": "0"}, {"HG_TILE_ORDER": "0", "HG_CONCURRENT_KELEMS": "0"}])
def test_alternative_kernel_paths_keep_parity(env):
    """Every layer has more than one kernel path (CTA-pair / single-CTA tcgen05, fused / unfused
    ResBlock pairs, TMA / generic epilogue, CUDA-core).  Each combination must meet the same
    tolerances; the switches are read at plan creation, so each runs in its own process."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, json, torch, numpy as np\n"
        f"sys.path.insert(0, {root!r}); sys.path.insert(0, {os.path.join(root, 'tests')!r})\n"
        "from oracle import fixtures as fx\n"
        "from oracle.common import max_abs, snr_db\n"
        "from _util import golden, make_generator\n"
        "out = {'fp32': [0.0, 1e9], 'bf16': [0.0, 1e9]}\n"
        "for name, cfg in (('v1', fx.V1), ('v2_narrow', fx.V2_NARROW)):\n"
        "    g = golden(name + '_seed1234')\n"
        "    for prec in ('fp32', 'bf16'):\n"
        "        m = make_generator(cfg, precision=prec).cuda()\n"
        "        with torch.no_grad():\n"
        "            y = m(torch.from_numpy(g['mel_b']).cuda()).cpu().numpy()\n"
        "        out[prec] = [max(out[prec][0], max_abs(y, g['y_b'])), min(out[prec][1], snr_db(g['y_b'], y))]\n"
        "print(json.dumps(out))\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env={**os.environ, **env})
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert res["fp32"][0] <= FP32_TOL, res
    assert res["bf16"][1] >= BF16_SNR_DB, res


def test_concurrent_resblocks_are_bit_identical_to_the_serial_schedule():
    """Short inputs run the ResBlocks of a stage on separate streams (api.cu::concurrent_stage): same kernels, same
    MRF summation order, so the waveform must be the serial schedule's bit for bit — eager, ragged, int16, as a CUDA
    graph, and with only some of the stages under the size threshold.  The switch is read at plan creation, so the
    two schedules are two generators in one subprocess."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import os, sys, json, torch\n"
        f"sys.path.insert(0, {root!r}); sys.path.insert(0, {os.path.join(root, 'tests')!r})\n"
        "from oracle import fixtures as fx\n"
        "from _util import make_generator\n"
        "res = {}\n"
        "for name, cfg in (('v1', fx.V1), ('v3_rb2', fx.V3_RB2), ('v2_narrow', fx.V2_NARROW)):\n"
        "    for prec in ('bf16', 'fp32'):\n"
        "        gens = {}\n"
        "        for tag, kelems in (('serial', '0'), ('conc', '2560'), ('stage0', '600')):\n"
        "            os.environ['HG_CONCURRENT_KELEMS'] = kelems\n"
        "            gens[tag] = make_generator(cfg, precision=prec).cuda()\n"
        "        ok = True\n"
        "        with torch.no_grad():\n"
        "            for B, T in ((1, 64), (1, 256), (3, 37), (2, 300)):\n"
        "                mel = fx.synthetic_mel(B, T, seed=7).cuda()\n"
        "                ref = gens['serial'](mel)\n"
        "                for tag in ('conc', 'stage0'):\n"
        "                    ok &= bool(torch.equal(gens[tag](mel), ref))\n"
        "                    ok &= bool(torch.equal(gens[tag](mel), ref))  # again: buffers and events are reused\n"
        "                ok &= bool(torch.equal(gens['conc'].generate_int16(mel), gens['serial'].generate_int16(mel)))\n"
        "                if B > 1:\n"
        "                    fr = [T, max(1, T // 3), T - 1][:B]\n"
        "                    ok &= bool(torch.equal(gens['conc'].forward_ragged(mel, fr), gens['serial'].forward_ragged(mel, fr)))\n"
        "            mel = fx.synthetic_mel(1, 128, seed=9).cuda()\n"
        "            run = gens['conc'].make_graphed(1, 128)\n"
        "            ok &= bool(torch.equal(run(mel), gens['serial'](mel)))\n"
        "            ok &= bool(torch.equal(run(mel), gens['serial'](mel)))\n"
        "        res[name + '/' + prec] = ok\n"
        "print(json.dumps(res))\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert all(res.values()), res


@pytest.mark.parametrize("rates,kernels", [((5, 2), (10, 4)), ((2, 2), (8, 4)), ((8, 2), (15, 4))])
def test_unsupported_upsampler_shapes_fail_loudly(rates, kernels):
    """hg_plan_create rejects ConvTranspose1d shapes whose output is not exactly rate x input (odd k - u,
    hifi/models.py:169) or whose polyphase form needs more than L + 1 GEMM rows (k > 3u) instead of running
    with mis-sized buffers."""
    from oracle.common import GenConfig

    cfg = GenConfig(32, rates, kernels, (3,), ((1, 3, 5),), "1")
    m = make_generator(cfg).cuda()
    with pytest.raises(RuntimeError, match="unsupported"):
        m(torch.zeros(1, 80, 4, device="cuda"))


def test_data_updates_need_invalidate_and_load_state_dict_does_not():
    """The engine's cache key cannot see `param.data` updates (the reference's own init_weights idiom,
    hifi/vocoder/utils.py:24-27); invalidate() re-folds, load_state_dict() invalidates by itself."""
    g = golden("tiny_rb1")
    m = make_generator(fx.TINY_RB1, seed=5, fold=True).cuda()
    mel = torch.from_numpy(g["mel"]).cuda()
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    with torch.no_grad():
        y0 = m(mel).clone()
        m.conv_post.bias.data.add_(0.25)
        m.invalidate()
        y1 = m(mel).clone()
        assert not torch.equal(y0, y1)
        m.load_state_dict(sd0)
        assert torch.equal(m(mel), y0)


def test_resblock_type_follows_the_block_class_built():
    """`resblock: 1` (int, unquoted YAML) is not == "1": the reference builds ResBlock2 (hifi/models.py:155)
    and so must the native plan and the halo."""
    from tts_king_b200.hifi.models import Generator, ResBlock2

    h = fx.make_h(fx.TINY_RB2)
    h["resblock"] = 1
    h.resblock = 1
    torch.manual_seed(3)
    m = Generator(h)
    assert all(isinstance(b, ResBlock2) for b in m.resblocks)
    m.remove_weight_norm()
    m.eval().cuda()
    y = m(fx.synthetic_mel(1, 6, seed=2).cuda())
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    ref = torch_oracle.forward(fx.TINY_RB2, sd, fx.synthetic_mel(1, 6, seed=2))
    assert max_abs(y.cpu().numpy(), ref.numpy()) <= FP32_TOL
    assert m.halo_frames == parallel.halo_frames(fx.make_h(fx.TINY_RB2))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (gpurun --gpus 2)")
def test_one_process_two_gpus():
    """one process may drive several GPUs.  Function attributes (opt-in shared memory)
    are per device, and no entry point may leave the caller's current device changed."""
    mel = fx.synthetic_mel(2, 40, seed=3)
    outs = []
    for prec in ("bf16", "fp32"):
        m0 = make_generator(fx.V1, precision=prec).to("cuda:0")
        m1 = make_generator(fx.V1, precision=prec).to("cuda:1")
        torch.cuda.set_device(0)
        with torch.no_grad():
            y1 = m1(mel.to("cuda:1"))          # first use of every kernel on device 1 while device 0 is current
            assert torch.cuda.current_device() == 0
            y0 = m0(mel.to("cuda:0"))
            y1b = m1(mel.to("cuda:1"))
        torch.cuda.synchronize(0)
        torch.cuda.synchronize(1)
        assert torch.equal(y0.cpu(), y1.cpu()) and torch.equal(y1.cpu(), y1b.cpu())
        del m1                                  # plan destruction on device 1 ...
        import gc
        gc.collect()
        assert torch.cuda.current_device() == 0  # ... must not move the caller
        outs.append(y0)
