"""``HIFIapi`` — mirror of the reference wrapper ``hifiapi.py:11-52``.

Same constructor and methods: ``HIFIapi(config, device)``, ``__call__(x)`` (float wav) and
``generate(mel_specs)`` (int16 numpy).  ``config`` is the object the reference passes around
(OmegaConf / nested AttrDict): ``config.hifi`` is the ``h`` of ``Generator(h)``,
``config.hifi.weights_path`` an optional ``{"generator": state_dict}`` checkpoint,
``config.hifi.MAX_WAV_VALUE`` the int16 scale and ``config.model_config["vocoder"]["use_cpu"]`` the
reference's CPU switch.

Device semantics.  The reference computes wherever ``device`` says and ships ``use_cpu: true``.  This
implementation computes on a B200 only, so ``device`` keeps its meaning for the caller-facing side —
where ``__call__`` expects its input and returns its output — while the arithmetic always runs on
``compute_device`` (default ``cuda:<current>``).  With ``use_cpu: true`` a caller therefore still
passes and receives CPU tensors, exactly as before, and the copies happen inside.
"""
from __future__ import annotations

import torch

from .hifi.models import Generator


class AttrDict(dict):
    """dict whose keys are also attributes (the helper the reference keeps next to its wrapper,
    hifiapi.py:5-8) — configs built by hand use it in place of OmegaConf.  Attribute access is routed
    to the items, so copies and pickles of a config (and of a module holding one) keep working."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name) from None

    def __setattr__(self, name, value):
        self[name] = value

    def __delattr__(self, name):
        try:
            del self[name]
        except KeyError:
            raise AttributeError(name) from None


class HIFIapi:
    def __init__(self, config, device="gpu", compute_device=None, precision="fp32"):
        caller_side = "cpu" if config.model_config["vocoder"]["use_cpu"] else device
        if caller_side == "gpu":  # the reference's default argument is not a torch device name
            caller_side = "cuda"
        if not torch.cuda.is_available():
            raise RuntimeError("tts_king_b200.HIFIapi needs a CUDA (sm_100a) device; there is no CPU fallback")
        if compute_device is None:
            compute_device = caller_side if str(caller_side).startswith("cuda") else "cuda"
        compute_device = torch.device(compute_device)
        if compute_device.index is None:
            compute_device = torch.device("cuda", torch.cuda.current_device())

        generator = Generator(config.hifi, precision=precision)
        ckpt_path = config.hifi.weights_path  # optional {"generator": state_dict} file (hifiapi.py:17-22)
        if ckpt_path is not None:
            state = torch.load(ckpt_path, map_location="cpu")["generator"]
            generator.load_state_dict(state)
        generator.to(compute_device)
        generator.remove_weight_norm()

        self.model = generator.eval()
        self.cfg = config
        self.device = caller_side
        self.compute_device = compute_device

    def train(self):
        raise NotImplementedError(" Train for HiFi was not implemented yet")

    def __call__(self, x):
        # kept for compatibility with other vocoders / callers (fs_two/utils/model.py:90)
        y = self.model(x.to(self.compute_device))
        return y.to(self.device)

    def generate(self, mel_specs):
        """
        Converts a batch of mel spectrograms [B,80,T] into int16 audio [B,1,T*hop] on the host
        (numpy), like the reference: wav * MAX_WAV_VALUE, truncating cast.  The scale and cast are
        fused into the last kernel, so only 2 bytes per sample cross PCIe.
        """
        scale = float(self.cfg.hifi.MAX_WAV_VALUE)
        with torch.no_grad():
            mel_dev = mel_specs.to(self.compute_device, non_blocking=True)
            audio = self.model.eval().generate_int16(mel_dev, scale)
            # device -> pinned host at full PCIe rate (a pageable .cpu() is staged and ~3x slower); the
            # pinned block comes from torch's caching host allocator and is owned by the returned array
            host = torch.empty(audio.shape, dtype=audio.dtype, pin_memory=True)
            host.copy_(audio, non_blocking=True)
            torch.cuda.current_stream(self.compute_device).synchronize()
        return host.numpy()
