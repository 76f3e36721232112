"""Multi-GPU partitioning of the vocoder path (one process per GPU, ``torch.distributed``).

The reference has no parallelism at all; these are the two partitionings the
generator's structure allows:

* **utterance sharding** — batch items never interact (no normalisation layers), so a set of
  utterances is split across ranks by frame count with no data-path collective;
* **time-chunk sharding of one long mel** — the generator is a finite-receptive-field stack
  (reference hifi/models.py:185-201: k7 conv, 4 transposed convs, dilated ResBlocks, k7 conv), so a
  chunk computed with ``halo`` extra mel frames on each interior side reproduces the full-length
  result exactly on its own samples.  The only communication is the neighbour exchange of
  ``halo`` x 80 floats (4 160 B for V1) and the final gather of the waveform.

Everything here is host logic plus ``torch.distributed`` calls; it works with the ``nccl`` backend
on GPUs and with ``gloo`` on CPU tensors (which is how tests/ exercises it).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import torch


# ----------------------------------------------------------------------------- receptive field
def receptive_reach_samples(h) -> int:
    """Upper bound on how far (in output samples, per side) one output sample's dependence reaches
    beyond the input frames that 'own' it (3 258 for V1)."""
    reach = 3  # conv_pre: k7, padding 3 (hifi/models.py:152-154), in mel frames
    rb1 = str(h.resblock) == "1"
    nd = 3 if rb1 else 2
    for u, k in zip(h.upsample_rates, h.upsample_kernel_sizes):
        reach = reach * int(u) + (int(k) - int(u) + 1) // 2
        blk = 0
        for rk, ds in zip(h.resblock_kernel_sizes, h.resblock_dilation_sizes):
            half = (int(rk) - 1) // 2
            r = sum(half * int(d) for d in list(ds)[:nd])
            if rb1:
                r += nd * half  # the dilation-1 second conv of each pair
            blk = max(blk, r)
        reach += blk
    return reach + 3  # conv_post: k7, padding 3


def hop_length(h) -> int:
    return int(math.prod(int(u) for u in h.upsample_rates))


def halo_frames(h) -> int:
    """Mel frames of context a time chunk needs on each interior side (13 for V1)."""
    return -(-receptive_reach_samples(h) // hop_length(h))


# ----------------------------------------------------------------------------- utterance sharding
def shard_utterances(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Greedy longest-first bin packing of utterance indices onto ranks by frame count.
    Deterministic (ties broken by index), every index appears exactly once."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    loads = [0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda j: (loads[j], j))
        shards[r].append(i)
        loads[r] += int(lengths[i])
    for s in shards:
        s.sort()
    return shards


# ----------------------------------------------------------------------------- time chunking
@dataclass(frozen=True)
class Chunk:
    start: int      # first owned frame
    stop: int       # one past the last owned frame
    lo: int         # first frame actually fed to the generator (start - halo, clipped to 0)
    hi: int         # one past the last frame fed (stop + halo, clipped to T)

    @property
    def frames(self) -> int:
        return self.stop - self.start


def plan_time_chunks(T: int, parts: int, halo: int) -> List[Chunk]:
    """Split [0,T) into `parts` contiguous chunks (sizes differ by at most one frame; trailing
    chunks may be empty when T < parts).  True sequence ends get no halo — the generator's own
    zero padding applies there, exactly as in the monolithic forward."""
    if T < 0 or parts < 1 or halo < 0:
        raise ValueError("bad arguments")
    base, extra = divmod(T, parts)
    chunks, a = [], 0
    for r in range(parts):
        n = base + (1 if r < extra else 0)
        b = a + n
        chunks.append(Chunk(a, b, max(0, a - halo), min(T, b + halo)) if n > 0 else Chunk(a, a, a, a))
        a = b
    return chunks


def forward_chunk(forward_fn: Callable[[torch.Tensor], torch.Tensor], mel: torch.Tensor, c: Chunk, hop: int) -> torch.Tensor:
    """Run one planned chunk of mel [B,80,T] and keep only the samples it owns."""
    y = forward_fn(mel[:, :, c.lo:c.hi])
    return y[..., (c.start - c.lo) * hop:(c.stop - c.lo) * hop]


def chunked_forward(forward_fn: Callable[[torch.Tensor], torch.Tensor], mel: torch.Tensor, chunk_frames: int,
                    halo: int, hop: int) -> torch.Tensor:
    """Single-device long-form synthesis: equal to forward_fn(mel) but with activation memory bounded
    by chunk_frames (a 60-minute mel needs > 10 GB per activation tensor monolithically)."""
    T = mel.shape[-1]
    parts = max(1, -(-T // max(1, chunk_frames)))
    outs = [forward_chunk(forward_fn, mel, c, hop) for c in plan_time_chunks(T, parts, halo) if c.frames > 0]
    return torch.cat(outs, dim=-1)


def chunked_forward_into(generator, mel: torch.Tensor, chunk_frames: int, halo: int, out: torch.Tensor,
                         max_wav_value: float = 32768.0) -> torch.Tensor:
    """``chunked_forward`` without the concatenation: every chunk's last kernel stores its owned samples
    directly at their place in ``out`` [B,1,T*hop] (``Generator.forward_into``), the halo samples are
    never written.  ``out`` may be a window of a buffer on another GPU (see ``share_output_buffer``)."""
    T = mel.shape[-1]
    hop = generator.hop_length
    parts = max(1, -(-T // max(1, chunk_frames)))
    for c in plan_time_chunks(T, parts, halo):
        if c.frames > 0:
            generator.forward_into(mel[:, :, c.lo:c.hi], out[:, :, c.start * hop:c.stop * hop], c.start - c.lo, c.frames,
                                   max_wav_value)
    return out


# ----------------------------------------------------------------------------- distributed pieces
def exchange_halo(local_mel: torch.Tensor, halo: int, group=None) -> tuple:
    """Neighbour exchange for a mel that is already time-sharded across ranks.

    local_mel: [B,80,T_r] — rank r's contiguous slice, ranks ordered in time, every T_r >= halo
    (except that empty ranks are not supported).  Returns (padded_mel, left, right) where left/right
    are the numbers of halo frames received (0 at the true ends).  Traffic: halo*80*4 bytes per side.
    """
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1 or halo == 0:
        return local_mel, 0, 0
    B, C, T = local_mel.shape
    if T < halo:
        raise ValueError(f"rank {rank}: local slice ({T} frames) shorter than the halo ({halo})")
    ops, left_buf, right_buf = [], None, None
    send_l = local_mel[:, :, :halo].contiguous()
    send_r = local_mel[:, :, T - halo:].contiguous()
    if rank > 0:
        left_buf = torch.empty_like(send_l)
        ops += [dist.P2POp(dist.isend, send_l, _global_rank(rank - 1, group), group),
                dist.P2POp(dist.irecv, left_buf, _global_rank(rank - 1, group), group)]
    if rank < world - 1:
        right_buf = torch.empty_like(send_r)
        ops += [dist.P2POp(dist.isend, send_r, _global_rank(rank + 1, group), group),
                dist.P2POp(dist.irecv, right_buf, _global_rank(rank + 1, group), group)]
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    parts = [p for p in (left_buf, local_mel, right_buf) if p is not None]
    return torch.cat(parts, dim=2), (halo if left_buf is not None else 0), (halo if right_buf is not None else 0)


def _global_rank(group_rank: int, group) -> int:
    import torch.distributed as dist

    return group_rank if group is None else dist.get_global_rank(group, group_rank)


def sharded_long_form(forward_fn: Callable[[torch.Tensor], torch.Tensor], local_mel: torch.Tensor, halo: int, hop: int,
                      group=None) -> torch.Tensor:
    """Rank-local waveform of a time-sharded mel: halo exchange, one forward, trim the halo samples."""
    padded, left, right = exchange_halo(local_mel, halo, group)
    y = forward_fn(padded)
    n = y.shape[-1]
    return y[..., left * hop:n - right * hop]


class _PeerMapping:
    """A buffer of another process's GPU mapped for this rank's GPU (hg_ipc_import); exposes
    __cuda_array_interface__ so torch can alias it, closes the mapping when collected."""

    def __init__(self, device_index: int, handle: bytes, offset: int, shape, typestr: str):
        import ctypes

        from . import _native

        self._L = _native.lib()
        self._device = device_index
        base, ptr = ctypes.c_void_p(), ctypes.c_void_p()
        _native.check(self._L.hg_ipc_import(device_index, handle, offset, ctypes.byref(base), ctypes.byref(ptr)))
        self._base = base
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr.value, False), "version": 2}

    def __del__(self):
        try:
            if self._base:
                self._L.hg_ipc_close(self._device, self._base)
                self._base = None
        except Exception:
            pass


_TYPESTR = {torch.float32: "<f4", torch.int16: "<i2"}


def share_output_buffer(shape, dtype=torch.float32, owner: int = 0, group=None) -> torch.Tensor:
    """One waveform buffer on rank ``owner``'s GPU, writable by the kernels of every rank of the node:
    the owner allocates it and exports it through CUDA IPC (``hg_ipc_export``), the handle travels over
    the process group, and every other rank maps it for ITS OWN GPU (``hg_ipc_import``) and gets a
    tensor that aliases the owner's memory (labelled with the local device; loads and stores go over
    NVLink).  ``Generator.forward_into`` can then store waveform samples straight into it, which
    replaces the gather collective of ``gather_wav``.  Single node only.  Keep the returned tensor
    alive on every rank, and ``dist.barrier()`` after the last writer has synchronised, before the
    owner reads."""
    import ctypes

    import torch.distributed as dist

    from . import _native

    rank = dist.get_rank(group)
    local = torch.cuda.current_device()
    box = [None]
    buf = None
    if rank == owner:
        buf = torch.empty(*shape, dtype=dtype, device=torch.device("cuda", local))
        handle = ctypes.create_string_buffer(64)
        off = ctypes.c_int64()
        _native.check(_native.lib().hg_ipc_export(buf.data_ptr(), handle, ctypes.byref(off)))
        box[0] = (handle.raw, off.value)
    dist.broadcast_object_list(box, src=_global_rank(owner, group), group=group)
    if rank != owner:
        handle, off = box[0]
        mapping = _PeerMapping(local, handle, off, shape, _TYPESTR[dtype])
        buf = torch.as_tensor(mapping, device=torch.device("cuda", local))
        buf._hg_peer_mapping = mapping  # the mapping lives as long as the tensor that aliases it
    return buf


def sharded_long_form_into(generator, local_mel: torch.Tensor, halo: int, out_full: torch.Tensor, first_frame: int,
                           chunk_frames: int = 16384, group=None, max_wav_value: float = 32768.0) -> None:
    """Time-sharded long-form synthesis with the gather fused into the last kernel: halo exchange with the
    neighbours, then this rank's frames ``[first_frame, first_frame + T_r)`` are vocoded chunk by chunk and
    every chunk's samples are stored directly into ``out_full`` [B,1,T_total*hop] (``share_output_buffer``),
    wherever that buffer lives."""
    padded, left, right = exchange_halo(local_mel, halo, group)
    hop = generator.hop_length
    T_r = local_mel.shape[-1]
    own = out_full[:, :, first_frame * hop:(first_frame + T_r) * hop]
    Tp = padded.shape[-1]
    parts = max(1, -(-T_r // max(1, chunk_frames)))
    for c in plan_time_chunks(T_r, parts, halo):
        if c.frames == 0:
            continue
        # chunk c owns local frames [c.start, c.stop); in `padded` they sit `left` frames further right, and the
        # neighbours' halo frames stand in for the chunk halos at this rank's two ends
        lo = max(0, c.start + left - halo)
        hi = min(Tp, c.stop + left + halo)
        generator.forward_into(padded[:, :, lo:hi], own[:, :, c.start * hop:c.stop * hop], c.start + left - lo, c.frames,
                               max_wav_value)


def gather_wav(local_wav: torch.Tensor, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """Concatenate rank-local waveforms [B,1,N_r] along time on rank `dst` (None elsewhere).
    Chunks may differ in length: sizes are exchanged first, payloads are padded to the maximum."""
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1:
        return local_wav
    n = torch.tensor([local_wav.shape[-1]], dtype=torch.int64, device=local_wav.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    nmax = max(sizes)
    pad = torch.zeros(*local_wav.shape[:-1], nmax, dtype=local_wav.dtype, device=local_wav.device)
    pad[..., :local_wav.shape[-1]] = local_wav
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=_global_rank(dst, group), group=group)
    if rank != dst:
        return None
    return torch.cat([b[..., :s] for b, s in zip(bufs, sizes)], dim=-1)
