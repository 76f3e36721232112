"""HiFi-GAN generator with the reference's Python surface and a CUDA engine underneath.

Mirror of the inference half of the reference's ``hifi/models.py`` (``Generator`` :146-210,
``ResBlock1`` :12-101, ``ResBlock2`` :104-143): same constructor, same module tree and therefore the
same ``state_dict`` keys (``weight_g`` / ``weight_v`` / ``bias`` before ``remove_weight_norm``,
``weight`` / ``bias`` after), same ``forward(mel[B,80,T]) -> wav[B,1,T*prod(rates)]``.

What differs is everything below that surface.  The ``torch.nn`` modules are parameter containers
only; ``forward`` folds the weight norm, hands the tensors to ``libhifigan_b200.so`` (C ABI in
``include/hifigan_b200.h``) and runs the whole network as ~80 hand-written sm_100a kernels
(tcgen05 implicit-GEMM convolutions with fused bias / leaky-ReLU / residual / MRF epilogues).
Inference only: the output never requires grad.  CUDA only: a CPU tensor raises.
"""
from __future__ import annotations

import ctypes
import types
from typing import Sequence, Dict, List, Optional, Tuple

import torch
import torch.nn as nn
from torch.nn import Conv1d, ConvTranspose1d
from torch.nn.utils import remove_weight_norm, weight_norm

from .. import _native
from .vocoder.utils import get_padding, init_weights

LRELU_SLOPE = 0.1
NUM_MELS = 80  # the reference hard-codes Conv1d(80, ...) at hifi/models.py:153


def _wn_conv(channels: int, kernel_size: int, dilation: int) -> nn.Module:
    return weight_norm(Conv1d(channels, channels, kernel_size, 1, dilation=dilation,
                              padding=get_padding(kernel_size, dilation)))


def _folded(m: nn.Module) -> torch.Tensor:
    """The effective weight of a (possibly still weight-normed) conv: torch._weight_norm(v, g, 0) —
    exactly what remove_weight_norm bakes in (reference hifi/models.py:97-101,203-210)."""
    if hasattr(m, "weight_g"):
        return torch._weight_norm(m.weight_v.detach(), m.weight_g.detach(), 0)
    return m.weight.detach()


def _run_op_conv(m: Conv1d, x_cl: torch.Tensor, in_slope: float, residual: Optional[torch.Tensor],
                 precision: str) -> torch.Tensor:
    """One Conv1d through the op-level C entry point; x_cl is channels-last [B, L, C] on CUDA."""
    L = _native.lib()
    w = _folded(m).float().cpu().contiguous()
    b = m.bias.detach().float().cpu().contiguous()
    B, n, cin = x_cl.shape
    y = torch.empty((B, n, m.out_channels), dtype=torch.float32, device=x_cl.device)
    with torch.cuda.device(x_cl.device):
        st = torch.cuda.current_stream().cuda_stream
        _native.check(L.hg_op_conv1d(x_cl.device.index, _native.PRECISIONS[precision], x_cl.data_ptr(), B, n, cin,
                                     w.data_ptr(), b.data_ptr(), m.out_channels, m.kernel_size[0], m.dilation[0],
                                     float(in_slope), residual.data_ptr() if residual is not None else None,
                                     y.data_ptr(), st))
    return y


class _ResBlockBase(nn.Module):
    precision = "fp32"

    def _check(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("tts_king_b200 ResBlock runs only on CUDA (sm_100a); there is no CPU fallback")
        return x.detach().float().transpose(1, 2).contiguous()  # [B, L, C]


class ResBlock1(_ResBlockBase):
    """Three (dilated conv, conv) pairs with pre-activation residuals — reference hifi/models.py:12-101."""

    def __init__(self, h, channels, kernel_size=3, dilation=(1, 3, 5)):
        super().__init__()
        self.h = h
        self.convs1 = nn.ModuleList([_wn_conv(channels, kernel_size, dilation[m]) for m in range(3)])
        self.convs1.apply(init_weights)
        self.convs2 = nn.ModuleList([_wn_conv(channels, kernel_size, 1) for _ in range(3)])
        self.convs2.apply(init_weights)

    def forward(self, x):
        # stand-alone use (block-level parity tests); Generator.forward runs the fused schedule
        r = self._check(x)
        for c1, c2 in zip(self.convs1, self.convs2):
            xt = _run_op_conv(c1, r, LRELU_SLOPE, None, self.precision)
            r = _run_op_conv(c2, xt, LRELU_SLOPE, r, self.precision)
        return r.transpose(1, 2).contiguous()

    def remove_weight_norm(self):
        for m in list(self.convs1) + list(self.convs2):
            remove_weight_norm(m)


class ResBlock2(_ResBlockBase):
    """Two dilated convs with pre-activation residuals — reference hifi/models.py:104-143."""

    def __init__(self, h, channels, kernel_size=3, dilation=(1, 3)):
        super().__init__()
        self.h = h
        self.convs = nn.ModuleList([_wn_conv(channels, kernel_size, dilation[m]) for m in range(2)])
        self.convs.apply(init_weights)

    def forward(self, x):
        r = self._check(x)
        for c in self.convs:
            r = _run_op_conv(c, r, LRELU_SLOPE, r, self.precision)
        return r.transpose(1, 2).contiguous()

    def remove_weight_norm(self):
        for m in self.convs:
            remove_weight_norm(m)


class _Engine:
    """Owns one HgPlan (device weights in kernel layout) and the scratch workspaces."""

    def __init__(self, cfg: _native.HgConfig, device: torch.device, weights: List[Tuple[str, torch.Tensor, torch.Tensor]]):
        self.L = _native.lib()
        self.device = device
        self.plan = ctypes.c_void_p()
        _native.check(self.L.hg_plan_create(ctypes.byref(cfg), device.index, ctypes.byref(self.plan)))
        try:
            for name, w, b in weights:
                shape = (ctypes.c_int64 * w.dim())(*w.shape)
                _native.check(self.L.hg_plan_upload_weight(self.plan, name.encode(), w.data_ptr(), shape, w.dim(),
                                                           b.data_ptr(), b.numel()))
            _native.check(self.L.hg_plan_finalize(self.plan))
        except Exception:
            self.close()
            raise
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

    def workspace(self, B: int, T: int, prec: int, stream: int) -> Tuple[int, int]:
        need = ctypes.c_size_t()
        _native.check(self.L.hg_workspace_bytes(self.plan, B, T, prec, ctypes.byref(need)))
        buf = self._ws.get(stream)
        if buf is None or buf.numel() < need.value + 1024:
            self._ws.pop(stream, None)
            buf = None
            buf = torch.empty(need.value + 1024, dtype=torch.uint8, device=self.device)
            self._ws[stream] = buf
        base = (buf.data_ptr() + 1023) & ~1023
        return base, buf.numel() - (base - buf.data_ptr())

    def launches(self, B: int, T: int, prec: int) -> int:
        n = ctypes.c_int()
        _native.check(self.L.hg_forward_launches(self.plan, B, T, prec, ctypes.byref(n)))
        return n.value

    def forward(self, mel: torch.Tensor, out: torch.Tensor, out_dtype: int, out_scale: float, prec: int,
                frames: Optional[Sequence[int]] = None):
        B, _, T = mel.shape
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            ws, ws_bytes = self.workspace(B, T, prec, st)
            sB, sC, sT = mel.stride()
            if frames is None:
                _native.check(self.L.hg_forward(self.plan, mel.data_ptr(), sB, sC, sT, B, T, out.data_ptr(), out_dtype,
                                                float(out_scale), prec, ws, ws_bytes, st))
            else:
                fr = (ctypes.c_int32 * B)(*[int(v) for v in frames])
                _native.check(self.L.hg_forward_ragged(self.plan, mel.data_ptr(), sB, sC, sT, B, T, fr, out.data_ptr(),
                                                       out_dtype, float(out_scale), prec, ws, ws_bytes, st))

    def forward_window(self, mel: torch.Tensor, out: torch.Tensor, skip: int, keep: int, out_dtype: int, out_scale: float,
                       prec: int):
        """hg_forward_window: samples [skip, skip+keep) of every item -> out[b, 0, :keep]; `out` may live on a
        peer GPU (its stores then go over NVLink)."""
        B, _, T = mel.shape
        if out.device != mel.device:
            _native.check(self.L.hg_enable_peer_access(mel.device.index, out.device.index))
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            ws, ws_bytes = self.workspace(B, T, prec, st)
            sB, sC, sT = mel.stride()
            _native.check(self.L.hg_forward_window(self.plan, mel.data_ptr(), sB, sC, sT, B, T, out.data_ptr(), out.stride(0),
                                                   skip, keep, out_dtype, float(out_scale), prec, ws, ws_bytes, st))

    def halo_frames(self) -> int:
        n = ctypes.c_int()
        _native.check(self.L.hg_halo_frames(self.plan, ctypes.byref(n)))
        return n.value


    def layer_table(self, prec: int) -> List[_native.HgLayerInfo]:
        n = ctypes.c_int()
        _native.check(self.L.hg_layer_count(self.plan, ctypes.byref(n)))
        out = []
        for i in range(n.value):
            info = _native.HgLayerInfo()
            _native.check(self.L.hg_layer_info(self.plan, i, prec, ctypes.byref(info)))
            out.append(info)
        return out

    def profile(self, mel: torch.Tensor, out: torch.Tensor, prec: int) -> List[Tuple[int, float]]:
        """One forward with CUDA events around every launch: [(plan layer index | -1, ms), ...]."""
        B, _, T = mel.shape
        cap = 4096
        idx = (ctypes.c_int * cap)()
        ms = (ctypes.c_float * cap)()
        n = ctypes.c_int()
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            ws, ws_bytes = self.workspace(B, T, prec, st)
            sB, sC, sT = mel.stride()
            _native.check(self.L.hg_profile_forward(self.plan, mel.data_ptr(), sB, sC, sT, B, T, out.data_ptr(),
                                                    _native.OUT_F32, 1.0, prec, ws, ws_bytes, st, idx, ms, cap,
                                                    ctypes.byref(n)))
        return [(idx[i], ms[i]) for i in range(n.value)]

    def launch_info(self, launch: int) -> _native.HgLayerInfo:
        """Kernel family and tiling launch `launch` of this thread's last profile() actually used."""
        info = _native.HgLayerInfo()
        _native.check(self.L.hg_profile_launch_info(self.plan, launch, ctypes.byref(info)))
        return info


class Generator(nn.Module):
    """``Generator(h)`` — reference hifi/models.py:146-210.

    ``h`` needs the attributes the reference reads: ``resblock`` ("1" or "2"), ``upsample_rates``,
    ``upsample_kernel_sizes``, ``upsample_initial_channel``, ``resblock_kernel_sizes``,
    ``resblock_dilation_sizes``.

    ``precision`` (attribute, not part of the reference API) selects the arithmetic of the
    contraction: ``"fp32"`` (default; bf16x3 split products on tensor cores, fp32-accurate),
    ``"bf16"`` (bf16 operands, fp32 accumulate and residual stream) or ``"fp32_ffma"`` (exact fp32
    on CUDA cores, slow).
    """

    def __init__(self, h, precision: str = "fp32"):
        super().__init__()
        self.h = h
        # the hyper-parameters this module reads, as plain values: `h` is whatever object the caller has
        # (OmegaConf node, AttrDict, namespace) and may not survive a deepcopy / pickle of the module
        self._hp = types.SimpleNamespace(
            resblock=str(h.resblock), upsample_initial_channel=int(h.upsample_initial_channel),
            upsample_rates=tuple(int(u) for u in h.upsample_rates),
            upsample_kernel_sizes=tuple(int(k) for k in h.upsample_kernel_sizes),
            resblock_kernel_sizes=tuple(int(k) for k in h.resblock_kernel_sizes),
            resblock_dilation_sizes=tuple(tuple(int(d) for d in ds) for ds in h.resblock_dilation_sizes))
        self.num_kernels = len(h.resblock_kernel_sizes)
        self.num_upsamples = len(h.upsample_rates)
        uic = h.upsample_initial_channel
        self.conv_pre = weight_norm(Conv1d(NUM_MELS, uic, 7, 1, padding=3))
        block = ResBlock1 if h.resblock == "1" else ResBlock2
        # everything downstream (native plan, halo) follows the block class actually built: an unquoted
        # YAML `resblock: 1` is the int 1, which the reference's `== "1"` (hifi/models.py:155) sends to ResBlock2
        self._hp.resblock = "1" if block is ResBlock1 else "2"

        self.ups = nn.ModuleList()
        for i, (u, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)):
            self.ups.append(weight_norm(ConvTranspose1d(uic // (2 ** i), uic // (2 ** (i + 1)), k, u,
                                                        padding=(k - u) // 2)))

        self.resblocks = nn.ModuleList()
        ch = uic
        for i in range(len(self.ups)):
            ch = uic // (2 ** (i + 1))
            for k, d in zip(h.resblock_kernel_sizes, h.resblock_dilation_sizes):
                self.resblocks.append(block(h, ch, k, d))

        self.conv_post = weight_norm(Conv1d(ch, 1, 7, 1, padding=3))
        self.ups.apply(init_weights)
        self.conv_post.apply(init_weights)

        if precision not in _native.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_native.PRECISIONS)}")
        self.precision = precision
        self._engine: Optional[_Engine] = None
        self._engine_key = None

    # ------------------------------------------------------------------ reference API
    def remove_weight_norm(self):
        print("Removing weight norm for inference HIFI GAN...")
        for m in self.ups:
            remove_weight_norm(m)
        for blk in self.resblocks:
            blk.remove_weight_norm()
        remove_weight_norm(self.conv_pre)
        remove_weight_norm(self.conv_post)

    def forward(self, x):
        return self._run(x, _native.OUT_F32, 1.0)

    # ------------------------------------------------------------------ extensions
    def invalidate(self):
        """Drop the packed device weights; the next forward re-folds and re-uploads them.  The engine is
        rebuilt automatically when a parameter is replaced or updated through autograd-visible in-place ops
        (``load_state_dict``, ``.to()``, ``remove_weight_norm``), but an update through ``param.data`` — the
        reference's own ``init_weights`` idiom (hifi/vocoder/utils.py:24-27), EMA swaps — does not bump the
        tensor version and cannot be seen: call this after such an update."""
        if self._engine is not None:
            self._engine.close()
        self._engine = None
        self._engine_key = None

    refresh_weights = invalidate

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate()
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self.invalidate()
        return out

    @torch.no_grad()
    def generate_int16(self, x, max_wav_value: float = 32768.0, frames: Optional[Sequence[int]] = None):
        """forward + ``* MAX_WAV_VALUE`` + numpy-style truncating int16 cast, fused into the last
        kernel (the tail of reference hifiapi.py:47-51).  Returns a device int16 tensor [B,1,N].
        ``frames``: see ``forward_ragged``."""
        return self._run(x, _native.OUT_I16, float(max_wav_value), frames)

    @torch.no_grad()
    def forward_ragged(self, x, frames: Sequence[int]):
        """``forward`` for a padded batch [B,80,T] (B <= 64) in which utterance ``i`` only keeps its first
        ``frames[i]`` mel frames (the vocoder call of synth_samples, fs_two/utils/tools.py:257-268).
        Rows past ``frames[i]`` + the receptive halo (13 frames for V1) are not computed at any
        layer; the returned [B,1,T*hop] tensor is bit-identical to ``forward(x)`` on
        ``[i, 0, :frames[i]*hop]`` and zero beyond (from the next multiple of 256 samples on, when hop is
        not a multiple of the last kernel's 256-sample tile)."""
        return self._run(x, _native.OUT_F32, 1.0, frames)

    @torch.no_grad()
    def forward_into(self, x, out, skip_frames: int = 0, keep_frames: Optional[int] = None, max_wav_value: float = 32768.0):
        """Run ``forward(x)`` and store only the samples of mel frames ``[skip_frames, skip_frames +
        keep_frames)`` into ``out`` [B,1,keep_frames*hop] (float32, or int16 for the fused
        ``* max_wav_value`` + cast) — the epilogue of the last kernel writes them in place, nothing else
        is materialised.  This is how a time chunk computed with halo frames delivers its owned samples
        straight into the final waveform (``parallel.chunked_forward_into``).  ``out`` may be a view into a
        larger buffer (unit stride along time) and may live on ANOTHER GPU of the node, peer-mapped into
        this process (``parallel.share_output_buffer``): the stores then travel over NVLink and no gather
        collective is needed."""
        if not isinstance(x, torch.Tensor) or x.dim() != 3 or x.shape[1] != NUM_MELS:
            raise RuntimeError(f"expected input[B, {NUM_MELS}, T], got {list(getattr(x, 'shape', []))}")
        eng = self._get_engine()
        if x.device != eng.device:
            raise RuntimeError(f"input is on {x.device} but the generator's weights are on {eng.device}")
        B, _, T = x.shape
        keep = T - skip_frames if keep_frames is None else int(keep_frames)
        hop = self.hop_length
        if skip_frames < 0 or keep < 1 or skip_frames + keep > T:
            raise ValueError(f"frames [{skip_frames}, {skip_frames + keep}) outside the {T} frames of the input")
        if out.dtype not in (torch.float32, torch.int16) or out.device.type != "cuda":
            raise ValueError("out must be a CUDA float32 or int16 tensor")
        if tuple(out.shape) != (B, 1, keep * hop) or out.stride(2) != 1:
            raise ValueError(f"out must be [B={B}, 1, {keep * hop}] with unit stride along time, got {list(out.shape)} / {out.stride()}")
        x = x.detach()
        if x.dtype != torch.float32:
            x = x.float()
        i16 = out.dtype == torch.int16
        eng.forward_window(x, out, skip_frames * hop, keep * hop, _native.OUT_I16 if i16 else _native.OUT_F32,
                           float(max_wav_value) if i16 else 1.0, _native.PRECISIONS[self.precision])
        return out

    @torch.no_grad()
    def make_graphed(self, B: int, T: int, out_int16: bool = False, max_wav_value: float = 32768.0):
        """Capture one forward for a fixed [B, 80, T] shape into a CUDA graph and return a callable
        ``run(mel) -> wav``.  The reference has nothing comparable (eager, default
        stream); this is for the small-T latency regime (BASELINE cfg-1), where the ~60 kernel
        launches of a forward cost more than their execution.  ``run`` copies ``mel`` (any strides,
        same device) into a static buffer, replays the graph and returns the static output tensor —
        clone it if it must survive the next call."""
        eng = self._get_engine()
        dev = eng.device
        static_in = torch.zeros((B, NUM_MELS, T), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):  # warm-up outside capture: tensor maps, function attributes, workspace
                    static_out = self._run(static_in, _native.OUT_I16 if out_int16 else _native.OUT_F32, max_wav_value)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = self._run(static_in, _native.OUT_I16 if out_int16 else _native.OUT_F32, max_wav_value)

        def run(mel: torch.Tensor) -> torch.Tensor:
            if tuple(mel.shape) != (B, NUM_MELS, T):
                raise RuntimeError(f"graphed generator expects input[{B}, {NUM_MELS}, {T}], got {list(mel.shape)}")
            static_in.copy_(mel, non_blocking=True)
            graph.replay()
            return static_out

        run.graph = graph  # keep the graph (and its private workspace) alive with the callable
        return run

    @property
    def hop_length(self) -> int:
        n = 1
        for u in self._hp.upsample_rates:
            n *= u
        return n

    @property
    def halo_frames(self) -> int:
        """Mel frames of context a sample needs on each side (13 for V1): the receptive reach of the
        stack in frames — what time chunks and ragged batches add to the frames they own."""
        from .. import parallel

        return parallel.halo_frames(self._hp)

    def kernel_launches(self, B: int, T: int) -> int:
        """Kernels one forward enqueues (bench.py reports it as gpu_launches)."""
        eng = self._get_engine()
        return eng.launches(B, T, _native.PRECISIONS[self.precision])

    @torch.no_grad()
    def profile_layers(self, x):
        """Per-launch device times of one forward: list of dicts (name, kind, shape, tiling, ms)."""
        eng = self._get_engine()
        prec = _native.PRECISIONS[self.precision]
        x = x.detach().float()
        B, _, T = x.shape
        out = torch.empty((B, 1, T * self.hop_length), device=x.device, dtype=torch.float32)
        rows = []
        for i, (li, ms) in enumerate(eng.profile(x, out, prec)):
            t = eng.launch_info(i)  # what the launch actually ran (kernel family, tiling)
            if li < 0:
                rows.append(dict(name="mel_to_operand", kind=-1, kernel=_native.KERNEL_PATHS[t.kernel_path], ms=ms))
                continue
            rows.append(dict(name=t.name.decode(), kind=t.kind, c_in=t.c_in, c_out=t.c_out, k=t.k, dilation=t.dilation,
                             stride=t.stride, kernel=_native.KERNEL_PATHS[t.kernel_path], tensor_core=bool(t.tensor_core),
                             n_tile=t.n_tile, k_chunk=t.k_chunk, m_subtiles=t.m_subtiles, stages=t.stages,
                             smem_bytes=t.smem_bytes, weights_resident=bool(t.weights_resident),
                             slab_buffers=t.slab_buffers, ms=ms))
        return rows

    # ------------------------------------------------------------------ internals
    def __getstate__(self):
        # the native plan is process-local: drop it on pickle / deepcopy, it is rebuilt on demand
        state = self.__dict__.copy()
        state["_engine"] = None
        state["_engine_key"] = None
        return state

    def _convs(self) -> List[Tuple[str, nn.Module]]:
        out: List[Tuple[str, nn.Module]] = [("conv_pre", self.conv_pre)]
        out += [(f"ups.{i}", m) for i, m in enumerate(self.ups)]
        for n, blk in enumerate(self.resblocks):
            if isinstance(blk, ResBlock1):
                out += [(f"resblocks.{n}.convs1.{m}", c) for m, c in enumerate(blk.convs1)]
                out += [(f"resblocks.{n}.convs2.{m}", c) for m, c in enumerate(blk.convs2)]
            else:
                out += [(f"resblocks.{n}.convs.{m}", c) for m, c in enumerate(blk.convs)]
        out.append(("conv_post", self.conv_post))
        return out

    def _native_config(self) -> _native.HgConfig:
        h = self._hp
        c = _native.HgConfig()
        c.num_mels = NUM_MELS
        c.upsample_initial_channel = int(h.upsample_initial_channel)
        c.num_upsamples = self.num_upsamples
        for i, (u, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)):
            c.upsample_rates[i] = int(u)
            c.upsample_kernel_sizes[i] = int(k)
        c.num_kernels = self.num_kernels
        rb1 = h.resblock == "1"
        for j, (k, ds) in enumerate(zip(h.resblock_kernel_sizes, h.resblock_dilation_sizes)):
            c.resblock_kernel_sizes[j] = int(k)
            for m in range(3 if rb1 else 2):
                c.resblock_dilation_sizes[j][m] = int(ds[m])
        c.resblock_type = 1 if rb1 else 2
        return c

    def _get_engine(self) -> _Engine:
        params = list(self.parameters())
        device = params[0].device
        if device.type != "cuda":
            raise RuntimeError(
                "tts_king_b200.Generator runs only on a CUDA sm_100a device; move the module with .to('cuda') "
                "(there is no CPU fallback)")
        key = (device, tuple((p.data_ptr(), p._version) for p in params))
        if self._engine is None or self._engine_key != key:
            if self._engine is not None:
                self._engine.close()
            weights = [(name, _folded(m).float().cpu().contiguous(), m.bias.detach().float().cpu().contiguous())
                       for name, m in self._convs()]
            self._engine = _Engine(self._native_config(), device, weights)
            self._engine_key = key
        return self._engine

    def _run(self, x, out_dtype: int, out_scale: float, frames: Optional[Sequence[int]] = None):
        if not isinstance(x, torch.Tensor):
            raise TypeError("expected a Tensor")
        squeeze = x.dim() == 2
        if squeeze:
            x = x.unsqueeze(0)
        if x.dim() != 3 or x.shape[1] != NUM_MELS:
            raise RuntimeError(f"expected input[B, {NUM_MELS}, T] (or [{NUM_MELS}, T]), got {list(x.shape)}")
        eng = self._get_engine()
        if x.device != eng.device:
            raise RuntimeError(f"input is on {x.device} but the generator's weights are on {eng.device}")
        if x.shape[0] == 0 or x.shape[2] == 0:
            raise RuntimeError("empty batch / zero-length mel")
        x = x.detach()
        if x.dtype != torch.float32:
            x = x.float()
        B, _, T = x.shape
        if frames is not None:
            frames = [int(v) for v in (frames.tolist() if hasattr(frames, "tolist") else frames)]
            if len(frames) != B:
                raise ValueError(f"{len(frames)} frame counts for a batch of {B}")
            if B > _native.MAX_RAGGED_ITEMS:
                raise ValueError(f"a ragged batch holds at most {_native.MAX_RAGGED_ITEMS} utterances, got {B}")
        alloc = torch.empty if frames is None else torch.zeros  # ragged: the skipped tail reads as silence
        out = alloc((B, 1, T * self.hop_length), device=x.device,
                    dtype=torch.float32 if out_dtype == _native.OUT_F32 else torch.int16)
        eng.forward(x, out, out_dtype, out_scale, _native.PRECISIONS[self.precision], frames)
        return out.squeeze(0) if squeeze else out
