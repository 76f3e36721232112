"""``vocoder_infer`` — mirror of reference ``fs_two/utils/model.py:85-100`` for the HiFi-GAN branch.

Same signature, same return value (a list of int16 numpy arrays, trimmed to ``lengths``).  Differences
are all on the inside:

* ``* max_wav_value`` and the truncating int16 cast run in the last kernel's epilogue, so 2 bytes
  per sample cross PCIe instead of 4 and the host-side multiply/cast disappears;
* with ``lengths`` the batch is split into length buckets (``tts_king_b200.ragged``) and the padding
  the reference computes and then throws away is not computed.  Kept samples are bit-identical.
"""
from __future__ import annotations

import numpy as np
import torch

from ... import ragged


def vocoder_infer(mels, vocoder, model_config, preprocess_config, lengths=None):
    name = model_config["vocoder"]["model"]
    if name != "HiFi-GAN":
        raise NotImplementedError(f"vocoder {name!r}: only the HiFi-GAN branch of the reference is implemented")
    scale = float(preprocess_config["preprocessing"]["audio"]["max_wav_value"])
    gen = getattr(vocoder, "model", vocoder)  # a Generator, or a HIFIapi wrapping one
    dev = next(gen.parameters()).device
    with torch.no_grad():
        mels = mels.to(dev)
        if lengths is None:
            wavs = gen.generate_int16(mels, scale).squeeze(1)
            host = torch.empty(wavs.shape, dtype=wavs.dtype, pin_memory=True)
            host.copy_(wavs, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            return [w for w in host.numpy()]
        lens = [int(n) for n in (lengths.tolist() if hasattr(lengths, "tolist") else lengths)]
        parts = ragged.ragged_generate(gen, mels, lens, out_int16=True, max_wav_value=scale)
        offs = np.cumsum([0] + [int(p.numel()) for p in parts])
        host = torch.empty((int(offs[-1]),), dtype=torch.int16, pin_memory=True)
        for p, a, b in zip(parts, offs[:-1], offs[1:]):
            host[a:b].copy_(p, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        flat = host.numpy()
        return [flat[a:b] for a, b in zip(offs[:-1], offs[1:])]
