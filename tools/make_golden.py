#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference in this container.

The reference (diff7/tts-king) ships no tests or golden vectors for its HiFi-GAN path
(SURVEY.md §8c), so parity is pinned on outputs of the reference itself: this script imports
``hifi.models.Generator`` and ``hifiapi.HIFIapi`` from /root/reference (read-only), runs them on
seeded weights and seeded synthetic mels, and stores inputs/outputs.  /root/reference does not
exist on the GPU box, so only the committed .npz files travel.

    python tools/make_golden.py            # rewrites tests/golden/

Needs a 3-line ``matplotlib`` stub because hifi/vocoder/utils.py:1-8 imports it at module top
and it is not installed here.
"""
from __future__ import annotations

import os
import sys
import types
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("TTS_KING_REFERENCE", "/root/reference")


def import_reference():
    m = types.ModuleType("matplotlib")
    m.use = lambda *a, **k: None
    p = types.ModuleType("matplotlib.pylab")
    m.pylab = p
    sys.modules.setdefault("matplotlib", m)
    sys.modules.setdefault("matplotlib.pylab", p)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    warnings.filterwarnings("ignore")
    import hifi.models as ref_models  # noqa: E402
    import hifiapi as ref_api  # noqa: E402

    return ref_models, ref_api


def import_reference_postnet():
    """fs_two/transformer/Layers.py on its own: the package __init__ pulls in text front-ends whose
    dependencies (unidecode, ...) are not installed, and Layers.py itself only needs torch plus the
    two attention classes of SubLayers for FFTBlock, which PostNet does not touch."""
    import importlib.util

    pkg = types.ModuleType("_ref_transformer")
    pkg.__path__ = []
    sys.modules["_ref_transformer"] = pkg
    sub = types.ModuleType("_ref_transformer.SubLayers")
    sub.MultiHeadAttention = object
    sub.PositionwiseFeedForward = object
    sys.modules["_ref_transformer.SubLayers"] = sub
    spec = importlib.util.spec_from_file_location("_ref_transformer.Layers", os.path.join(REF, "fs_two", "transformer", "Layers.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["_ref_transformer.Layers"] = mod
    spec.loader.exec_module(mod)
    return mod


def make_postnet_golden(out_dir):
    """tests/golden/postnet.npz — the reference's PostNet (fs_two/transformer/Layers.py:71-143) in eval
    mode and the mel_linear + postnet + add tail of fs_two/model/fastspeech2.py:101-104."""
    from oracle import fixtures as fx

    ref_layers = import_reference_postnet()
    blob = {}
    # full size (80 -> 512 x3 -> 80): weights are seeded, not stored; their digest is
    torch.manual_seed(1234)
    full = ref_layers.PostNet(**fx.POSTNET_FULL)
    blob["full.digest_fresh"] = np.array(fx.state_digest({k: v for k, v in full.state_dict().items() if v.dtype.is_floating_point}))
    sd = fx.alive_batchnorm_({k: v.clone() for k, v in full.state_dict().items()})
    full.load_state_dict(sd)
    full.eval()
    blob["full.digest_alive"] = np.array(fx.state_digest({k: v for k, v in sd.items() if v.dtype.is_floating_point}))
    torch.manual_seed(77)
    lin = torch.nn.Linear(256, 80)  # fastspeech2.py:26-30
    torch.nn.init.xavier_normal_(lin.weight)
    dec = torch.randn(3, 41, 256, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        output = lin(dec)
        post = full(output)
        blob["full.lin_w"], blob["full.lin_b"] = lin.weight.detach().numpy(), lin.bias.detach().numpy()
        blob["full.decoder_output"] = dec.numpy()
        blob["full.output"] = output.numpy()
        blob["full.postnet"] = post.numpy()
        blob["full.postnet_output"] = (post + output).numpy()
        x1 = torch.randn(1, 1, 80, generator=torch.Generator().manual_seed(6))
        blob["full.x_T1"], blob["full.y_T1"] = x1.numpy(), full(x1).numpy()
    # tiny (80 -> 32 x3 -> 80): state dict stored, for load_state_dict parity
    torch.manual_seed(4321)
    tiny = ref_layers.PostNet(**fx.POSTNET_TINY)
    sd = fx.alive_batchnorm_({k: v.clone() for k, v in tiny.state_dict().items()}, seed=98)
    tiny.load_state_dict(sd)
    tiny.eval()
    for k, v in sd.items():
        blob["tiny.sd." + k] = v.numpy()
    x = torch.randn(2, 23, 80, generator=torch.Generator().manual_seed(9))
    with torch.no_grad():
        blob["tiny.x"], blob["tiny.y"] = x.numpy(), tiny(x).numpy()
    np.savez_compressed(os.path.join(out_dir, "postnet.npz"), **blob)
    print("postnet.npz:", {k: getattr(v, "shape", None) for k, v in blob.items() if not k.startswith("tiny.sd.")})


def main():
    from oracle import fixtures as fx

    if "--postnet-only" in sys.argv:
        make_postnet_golden(os.path.join(ROOT, "tests", "golden"))
        return

    ref_models, ref_api = import_reference()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)

    def ref_generator(cfg, seed=1234):
        torch.manual_seed(seed)
        g = ref_models.Generator(fx.make_h(cfg))
        sd_gv = {k: v.clone() for k, v in g.state_dict().items()}
        g.remove_weight_norm()
        g.eval()
        return g, sd_gv

    def run(g, mel):
        with torch.no_grad():
            return g(mel)

    # ---- BASELINE.json shapes on V1 (cfg-1: 1 x 256 frames; one 800-frame utterance of cfg-2): reference
    # outputs at the sizes the bench runs.  Mels are regenerated from their seed in the tests (sha256 kept
    # here); only the waveforms are stored.
    if "--skip-baseline-shapes" not in sys.argv:
        g, _ = ref_generator(fx.V1)
        blob = dict(digest_folded=np.array(fx.state_digest({k: v.clone() for k, v in g.state_dict().items()})))
        for tag, frames in (("cfg1", 256), ("utt800", 800)):
            mel = fx.synthetic_mel(1, frames, seed=7)
            blob[f"{tag}.frames"] = np.array(frames)
            blob[f"{tag}.mel_sha256"] = np.array(fx.tensor_digest(mel))
            blob[f"{tag}.y"] = run(g, mel).numpy()
        np.savez_compressed(os.path.join(out_dir, "v1_baseline_shapes.npz"), **blob)
        print("v1_baseline_shapes", {k: getattr(v, "shape", None) for k, v in blob.items()})
    if "--baseline-shapes-only" in sys.argv:
        return

    # ---- full-size configs: weights are NOT stored (55 MB); their sha256 is, and tests rebuild them
    # with the same seed through the new package's Generator(h) and compare digests.
    for name, cfg in (("v1", fx.V1), ("v2_narrow", fx.V2_NARROW), ("v3_rb2", fx.V3_RB2)):
        g, sd_gv = ref_generator(cfg)
        sd_folded = {k: v.clone() for k, v in g.state_dict().items()}
        mel_a = fx.synthetic_mel(1, 32, seed=7)
        mel_b = fx.synthetic_mel(2, 17, seed=8, kind="logmel")
        stage_out = []
        hooks = []
        # record the first channels of every upsampler output (cheap intermediate pins)
        for up in g.ups:
            hooks.append(up.register_forward_hook(lambda m, i, o: stage_out.append(o[0, :4, :256].clone())))
        y_a = run(g, mel_a)
        for h in hooks:
            h.remove()
        y_b = run(g, mel_b)
        # non-contiguous, time-major caller input as tts_king.py:48 produces it
        mel_tm = mel_b.transpose(1, 2).contiguous()
        y_tm = run(g, mel_tm.transpose(1, 2))
        assert torch.equal(y_tm, y_b)
        blob = dict(
            digest_gv=np.array(fx.state_digest(sd_gv)),
            digest_folded=np.array(fx.state_digest(sd_folded)),
            n_tensors_gv=np.array(len(sd_gv)),
            n_tensors_folded=np.array(len(sd_folded)),
            mel_a=mel_a.numpy(), y_a=y_a.numpy(), mel_b=mel_b.numpy(), y_b=y_b.numpy(),
        )
        for i, s in enumerate(stage_out):
            blob[f"ups{i}_head"] = s.numpy()
        # trained-like gains (T3b)
        alive = fx.alive_state(cfg)
        g.load_state_dict(alive)
        blob["alive_digest"] = np.array(fx.state_digest(alive))
        blob["y_alive_a"] = run(g, mel_a).numpy()
        np.savez_compressed(os.path.join(out_dir, f"{name}_seed1234.npz"), **blob)
        print(name, "y_a max", float(y_a.abs().max()), "alive max", float(np.abs(blob["y_alive_a"]).max()))

    # ---- tiny configs: the full g/v state_dict is stored, so the oracle is pinned without any
    # dependence on reproducing the init.
    for name, cfg in (("tiny_rb1", fx.TINY_RB1), ("tiny_rb2", fx.TINY_RB2)):
        g, sd_gv = ref_generator(cfg)
        mel = fx.synthetic_mel(2, 9, seed=11)
        y = run(g, mel)
        mel1 = fx.synthetic_mel(1, 1, seed=12)  # T = 1 edge case
        y1 = run(g, mel1)
        blob = {f"sd.{k}": v.numpy() for k, v in sd_gv.items()}
        blob.update(mel=mel.numpy(), y=y.numpy(), mel_T1=mel1.numpy(), y_T1=y1.numpy())
        alive = fx.alive_state(cfg, gain=1.0)
        g.load_state_dict(alive)
        y_alive = run(g, mel)
        blob.update({f"alive.{k}": v.numpy() for k, v in alive.items()})
        blob["y_alive"] = y_alive.numpy()
        # HIFIapi.generate tail on the alive output (hifiapi.py:50-51)
        blob["y_alive_int16"] = (y_alive * 32768).cpu().numpy().astype("int16")
        np.savez_compressed(os.path.join(out_dir, f"{name}.npz"), **blob)
        print(name, "y max", float(y.abs().max()), "alive max", float(y_alive.abs().max()))

    # ---- HIFIapi wrapper end to end (hifiapi.py:11-52) on V1
    cfg_obj = fx.AttrDict(
        hifi=fx.make_h(fx.V1),
        model_config=fx.AttrDict(vocoder=fx.AttrDict(use_cpu=True)),
    )
    cfg_obj.hifi["weights_path"] = None
    torch.manual_seed(1234)
    api = ref_api.HIFIapi(cfg_obj, "cpu")
    mel = fx.synthetic_mel(1, 16, seed=7)
    wav_i16 = api.generate(mel)
    wav_f32 = api(mel).detach()
    np.savez_compressed(os.path.join(out_dir, "hifiapi_v1.npz"), mel=mel.numpy(), generate_int16=wav_i16,
                        call_f32=wav_f32.numpy())
    # the int16 cast edge cases numpy produces on this platform
    edge = np.array([0.0, 0.5, -0.5, 0.99999, -1.0, 1.0, 32767.4 / 32768, -32768.9 / 32768, 1e-9, -1e-9],
                    dtype=np.float32)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        edge_i16 = (torch.from_numpy(edge) * 32768).numpy().astype("int16")
    np.savez_compressed(os.path.join(out_dir, "int16_cast.npz"), x=edge, y=edge_i16)
    make_postnet_golden(out_dir)
    print("hifiapi", wav_i16.shape, wav_i16.dtype, int(np.abs(wav_i16).max()), "edge", edge_i16.tolist())


if __name__ == "__main__":
    main()
