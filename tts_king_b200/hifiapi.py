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
    def __init__(self, *args, **kwargs):
        super(AttrDict, self).__init__(*args, **kwargs)
        self.__dict__ = self


class HIFIapi:
    def __init__(self, config, device="gpu", compute_device=None, precision="fp32"):
        if config.model_config["vocoder"]["use_cpu"]:
            device = "cpu"
        if device == "gpu":  # the reference's default is not a torch device name
            device = "cuda"
        if compute_device is None:
            compute_device = device if str(device).startswith("cuda") else "cuda"
        if not torch.cuda.is_available():
            raise RuntimeError("tts_king_b200.HIFIapi needs a CUDA (sm_100a) device; there is no CPU fallback")
        compute_device = torch.device(compute_device)
        if compute_device.index is None:
            compute_device = torch.device("cuda", torch.cuda.current_device())

        # Load checkpoint if exists
        weights_path = config.hifi.weights_path

        self.model = Generator(config.hifi, precision=precision)
        if weights_path is not None:
            checkpoint = torch.load(weights_path, map_location="cpu")
            self.model.load_state_dict(checkpoint["generator"])

        self.cfg = config
        self.device = device
        self.compute_device = compute_device

        self.model.to(compute_device)
        self.model.remove_weight_norm()
        self.model.eval()

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
        self.model.eval()
        with torch.no_grad():
            mel_dev = mel_specs.to(self.compute_device, non_blocking=True)
            audio = self.model.generate_int16(mel_dev, float(self.cfg.hifi.MAX_WAV_VALUE))
            # device -> pinned host at full PCIe rate (a pageable .cpu() is staged and ~3x slower); the
            # pinned block comes from torch's caching host allocator and is owned by the returned array
            host = torch.empty(audio.shape, dtype=audio.dtype, pin_memory=True)
            host.copy_(audio, non_blocking=True)
            torch.cuda.current_stream(self.compute_device).synchronize()
        return host.numpy()
