"""The tail of ``FastSpeech2.forward`` — mirror of reference ``fs_two/model/fastspeech2.py:101-104``:

    output = self.mel_linear(output)                    # nn.Linear(decoder_hidden, 80)
    postnet_output = self.postnet(output) + output

and the hand-off to the vocoder (``tts_king.py:48``: ``mel_spec.transpose(1, 2)``).  The acoustic
model proper (encoder, variance adaptor, decoder) is out of scope and stays in
PyTorch; this module takes its decoder output [B,T,hidden] on the GPU and keeps everything after it
in the native kernels and time-major: mel_linear is a k = 1 conv, the PostNet is a conv stack, the
residual add sits in the last epilogue, and ``Generator`` reads the [B,T,80] result in place
through its strides — no transpose copy anywhere.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn

from ..transformer.Layers import PostNet, _StackEngine, _stack_input
from ... import _native


class MelLinear(nn.Linear):
    """``nn.Linear(decoder_hidden, n_mel_channels)`` (fastspeech2.py:26-30) whose forward runs as a
    k = 1 conv stack on the GPU.  Same parameters / state_dict as nn.Linear."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True, precision: str = "fp32"):
        super().__init__(in_features, out_features, bias)
        if precision not in _native.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_native.PRECISIONS)}")
        self.precision = precision
        self._engine: Optional[_StackEngine] = None
        self._engine_key = None

    def invalidate(self):
        """Drop the packed device weights (rebuilt on the next forward); call after a ``.data`` update."""
        if self._engine is not None:
            self._engine.close()
        self._engine = None
        self._engine_key = None

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate()
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self.invalidate()
        return out

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_engine"] = None
        state["_engine_key"] = None
        return state

    def _get_engine(self) -> _StackEngine:
        device = self.weight.device
        if device.type != "cuda":
            raise RuntimeError("tts_king_b200.MelLinear runs only on a CUDA sm_100a device (there is no CPU fallback)")
        tensors = [self.weight] + ([self.bias] if self.bias is not None else [])
        key = (device, tuple((t.data_ptr(), t._version) for t in tensors))
        if self._engine is None or self._engine_key != key:
            if self._engine is not None:
                self._engine.close()
            w = self.weight.detach().float().cpu().contiguous().unsqueeze(-1)  # [out, in, 1]
            b = (self.bias.detach().float().cpu().contiguous() if self.bias is not None
                 else torch.zeros(self.out_features))
            self._engine = _StackEngine(device, [(w, b, _native.ACT_NONE, 0.0)])
            self._engine_key = key
        return self._engine

    def forward(self, x):
        eng = self._get_engine()
        return eng.forward(_stack_input(x, eng.c_in, eng.device), None, _native.PRECISIONS[self.precision])


@torch.no_grad()
def mel_tail(decoder_output: torch.Tensor, mel_linear: MelLinear, postnet: PostNet) -> Tuple[torch.Tensor, torch.Tensor]:
    """(output, postnet_output) of fastspeech2.py:101-104, both [B,T,80] time-major on the GPU."""
    output = mel_linear(decoder_output)
    return output, postnet.forward_residual(output)


def mel_to_vocoder(mel_time_major: torch.Tensor) -> torch.Tensor:
    """tts_king.py:48: the [B,80,T] view the vocoder API expects — a stride trick, not a copy;
    ``Generator`` reads it in place."""
    return mel_time_major.transpose(1, 2)
