"""``ConvNorm`` and ``PostNet`` — mirror of reference ``fs_two/transformer/Layers.py:36-143``
(the five Conv1d(k5) + BatchNorm1d (+ tanh) layers between FastSpeech2's
decoder and the vocoder).

Same constructors, same sub-module tree, therefore the same ``state_dict`` keys
(``convolutions.{i}.0.conv.{weight,bias}``, ``convolutions.{i}.1.{weight,bias,running_mean,
running_var,num_batches_tracked}``) and the same seeded initialisation as the reference.  ``forward``
runs on the B200 through ``hg_stack_forward`` (include/hifigan_b200.h): the eval-mode BatchNorm is
an affine map per channel and is folded into each conv's weight and bias when the plan is built;
the input stays time-major [B,T,80] end to end, so the reference's two transposes (:134,:142)
disappear.  Inference only — in training mode (batch statistics, dropout) ``forward`` refuses.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from ... import _native


class ConvNorm(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, padding=None, dilation=1, bias=True,
                 w_init_gain="linear"):
        super().__init__()
        if padding is None:
            assert kernel_size % 2 == 1
            padding = int(dilation * (kernel_size - 1) / 2)
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding,
                              dilation=dilation, bias=bias)

    def forward(self, signal):
        raise RuntimeError("tts_king_b200 ConvNorm is a parameter container; PostNet.forward runs the whole stack natively")


class _StackEngine:
    """One conv-stack plan (hg_stack_create) with its device weights and scratch."""

    def __init__(self, device: torch.device, layers: List[Tuple[torch.Tensor, torch.Tensor, int, float]]):
        # layers: (folded weight [C_out,C_in,k] fp32 cpu, bias [C_out] fp32 cpu, act, slope)
        self.L = _native.lib()
        self.device = device
        self.plan = ctypes.c_void_p()
        desc = (_native.HgStackLayer * len(layers))()
        for d, (w, _, act, slope) in zip(desc, layers):
            d.c_out, d.c_in, d.k = (int(v) for v in w.shape)
            d.dilation, d.act, d.slope = 1, int(act), float(slope)
        _native.check(self.L.hg_stack_create(desc, len(layers), device.index, ctypes.byref(self.plan)))
        try:
            for i, (w, b, _, _) in enumerate(layers):
                shape = (ctypes.c_int64 * 3)(*w.shape)
                _native.check(self.L.hg_plan_upload_weight(self.plan, str(i).encode(), w.data_ptr(), shape, 3,
                                                           b.data_ptr(), b.numel()))
            _native.check(self.L.hg_plan_finalize(self.plan))
        except Exception:
            self.close()
            raise
        self.c_in = int(layers[0][0].shape[1])
        self.c_out = int(layers[-1][0].shape[0])
        self._ws: Dict[int, torch.Tensor] = {}

    def close(self):
        if self.plan:
            self.L.hg_plan_destroy(self.plan)
            self.plan = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def forward(self, x: torch.Tensor, residual: Optional[torch.Tensor], prec: int) -> torch.Tensor:
        """x [B,T,c_in] fp32 (any strides) -> [B,T,c_out] fp32 contiguous (+ residual)."""
        B, T, C = x.shape
        out = torch.empty((B, T, self.c_out), device=x.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            need = ctypes.c_size_t()
            _native.check(self.L.hg_stack_workspace_bytes(self.plan, B, T, prec, ctypes.byref(need)))
            buf = self._ws.get(st)
            if buf is None or buf.numel() < need.value + 1024:
                self._ws.pop(st, None)
                buf = torch.empty(need.value + 1024, dtype=torch.uint8, device=self.device)
                self._ws[st] = buf
            base = (buf.data_ptr() + 1023) & ~1023
            sB, sT, sC = x.stride()
            _native.check(self.L.hg_stack_forward(self.plan, x.data_ptr(), sB, sC, sT, B, T,
                                                  residual.data_ptr() if residual is not None else None, out.data_ptr(),
                                                  prec, base, buf.numel() - (base - buf.data_ptr()), st))
        return out


def _stack_input(x: torch.Tensor, channels: int, device: torch.device) -> torch.Tensor:
    if not isinstance(x, torch.Tensor) or x.dim() != 3 or x.shape[2] != channels:
        raise RuntimeError(f"expected input[B, T, {channels}], got {list(getattr(x, 'shape', []))}")
    if x.device != device:
        raise RuntimeError(f"input is on {x.device} but the module's weights are on {device}")
    if x.shape[0] == 0 or x.shape[1] == 0:
        raise RuntimeError("empty batch / zero-length sequence")
    x = x.detach()
    return x if x.dtype == torch.float32 else x.float()


class PostNet(nn.Module):
    """
    PostNet: Five 1-d convolution with 512 channels and kernel size 5
    """

    def __init__(self, n_mel_channels=80, postnet_embedding_dim=512, postnet_kernel_size=5, postnet_n_convolutions=5,
                 precision: str = "fp32"):
        super().__init__()
        self.convolutions = nn.ModuleList()
        pad = int((postnet_kernel_size - 1) / 2)
        dims = [n_mel_channels] + [postnet_embedding_dim] * (postnet_n_convolutions - 1) + [n_mel_channels]
        for i in range(postnet_n_convolutions):
            gain = "linear" if i == postnet_n_convolutions - 1 else "tanh"
            self.convolutions.append(nn.Sequential(
                ConvNorm(dims[i], dims[i + 1], kernel_size=postnet_kernel_size, stride=1, padding=pad, dilation=1,
                         w_init_gain=gain),
                nn.BatchNorm1d(dims[i + 1])))
        if precision not in _native.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_native.PRECISIONS)}")
        self.precision = precision
        self._engine: Optional[_StackEngine] = None
        self._engine_key = None

    # ------------------------------------------------------------------ reference API
    def forward(self, x):
        """x [B,T,n_mel] -> [B,T,n_mel]  (Layers.py:133-143, eval mode)."""
        return self._run(x, residual=False)

    # ------------------------------------------------------------------ extension
    def forward_residual(self, x):
        """``postnet(x) + x`` (fastspeech2.py:104) with the add in the last conv's epilogue."""
        return self._run(x, residual=True)

    def invalidate(self):
        """Drop the packed device weights (rebuilt on the next forward).  Needed after an update through
        ``param.data`` / ``buffer.data``, which does not bump the tensor version the cache key watches;
        ``load_state_dict`` and ``.to()`` call it themselves."""
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

    # ------------------------------------------------------------------ internals
    def __getstate__(self):
        state = self.__dict__.copy()
        state["_engine"] = None
        state["_engine_key"] = None
        return state

    def folded_layers(self) -> List[Tuple[torch.Tensor, torch.Tensor]]:
        """[(weight, bias)] per conv with its eval-mode BatchNorm folded in, fp32 on the host."""
        out = []
        for seq in self.convolutions:
            conv, bn = seq[0].conv, seq[1]
            s = (bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps))
            w = conv.weight.detach().double() * s[:, None, None]
            b = (conv.bias.detach().double() - bn.running_mean.detach().double()) * s + bn.bias.detach().double()
            out.append((w.float().cpu().contiguous(), b.float().cpu().contiguous()))
        return out

    def _get_engine(self) -> _StackEngine:
        tensors = list(self.parameters()) + list(self.buffers())
        device = tensors[0].device
        if device.type != "cuda":
            raise RuntimeError("tts_king_b200.PostNet runs only on a CUDA sm_100a device; move the module with "
                               ".to('cuda') (there is no CPU fallback)")
        key = (device, tuple((t.data_ptr(), t._version) for t in tensors))
        if self._engine is None or self._engine_key != key:
            if self._engine is not None:
                self._engine.close()
            folded = self.folded_layers()
            n = len(folded)
            self._engine = _StackEngine(device, [(w, b, _native.ACT_TANH if i < n - 1 else _native.ACT_NONE, 0.0)
                                                 for i, (w, b) in enumerate(folded)])
            self._engine_key = key
        return self._engine

    def _run(self, x, residual: bool):
        if self.training:
            raise RuntimeError("tts_king_b200.PostNet is inference-only (eval-mode BatchNorm, no dropout): call .eval()")
        eng = self._get_engine()
        x = _stack_input(x, eng.c_in, eng.device)
        res = None
        if residual:
            res = x if x.is_contiguous() else x.contiguous()
        return eng.forward(x, res, _native.PRECISIONS[self.precision])
