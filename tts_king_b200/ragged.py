"""Length-aware scheduling for padded batches.

The reference pads every utterance of a batch to the batch maximum and runs the vocoder over the
padding (``fs_two/utils/tools.py:257-268`` -> ``fs_two/utils/model.py:85-100``: the waveforms are
trimmed to ``lengths`` only afterwards).  The generator is a finite-receptive-field stack, so the
samples of utterance ``i`` inside ``[0, lengths[i])`` depend on mel frames ``[0, len_i + halo)`` only
(``halo`` = 13 frames for V1, ``parallel.halo_frames``) and batch items never interact.  Two ways to
stop computing the padding, both bit-identical to the padded run on every kept sample:

* **in the kernels** (``Generator.forward_ragged`` -> ``hg_forward_ragged``): every launch gets a
  compacted tile index space — item ``b`` contributes only the tiles covering its first
  ``(len_b + halo) * rows_per_frame`` rows — so one forward runs at full occupancy over exactly the
  valid work.  This is what ``ragged_generate`` uses (in groups of 64 utterances, the kernels' limit).
* **length buckets** (``plan_length_buckets``): split the batch into contiguous length groups, each
  run as its own shorter dense forward; chosen by dynamic programming over
  ``sum_g (n_g * (max_len_g + halo) + launch_cost)``.  Needs nothing from the generator but a
  callable, which is how tests/ check the algebra against the CPU oracle; on the GPU it loses to the
  in-kernel path because small forwards underfill the machine (profiles/r1_ragged_bench.json).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import parallel


def plan_length_buckets(frames: Sequence[int], halo: int, t_max: int, launch_cost: int = 400,
                        max_buckets: Optional[int] = None) -> List[List[int]]:
    """Partition utterance indices into length buckets.

    frames[i]  mel frames utterance i needs (>= 1); halo: right context in frames; t_max: the padded
    extent (a bucket never runs longer than this).  Returns a list of index lists, longest bucket
    first; every index appears exactly once; deterministic."""
    n = len(frames)
    if n == 0:
        return []
    if any(int(f) < 1 for f in frames):
        raise ValueError("every utterance needs at least one frame")
    order = sorted(range(n), key=lambda i: (int(frames[i]), i))
    ext = [min(int(t_max), int(frames[i]) + halo) for i in order]  # extent a bucket ending at this item runs at
    kmax = n if max_buckets is None else max(1, min(n, int(max_buckets)))
    INF = float("inf")
    # best[k][j]: minimum cost of the first j items in exactly k buckets
    best = [[INF] * (n + 1) for _ in range(kmax + 1)]
    cut = [[0] * (n + 1) for _ in range(kmax + 1)]
    best[0][0] = 0.0
    for k in range(1, kmax + 1):
        for j in range(1, n + 1):
            for i in range(k - 1, j):
                if best[k - 1][i] == INF:
                    continue
                c = best[k - 1][i] + (j - i) * ext[j - 1] + launch_cost
                if c < best[k][j]:
                    best[k][j], cut[k][j] = c, i
    k = min(range(1, kmax + 1), key=lambda q: (best[q][n], q))
    buckets, j = [], n
    while k > 0:
        i = cut[k][j]
        buckets.append(sorted(order[i:j]))
        j, k = i, k - 1
    return buckets  # longest first


def bucket_extent(frames: Sequence[int], bucket: Sequence[int], halo: int, t_max: int) -> int:
    return min(int(t_max), max(int(frames[i]) for i in bucket) + halo)


def _halo_of(generator) -> int:
    h = getattr(generator, "halo_frames", None)
    return int(h) if h is not None else parallel.halo_frames(generator.h)


@torch.no_grad()
def ragged_generate(generator, mels: torch.Tensor, sample_lengths: Sequence[int], out_int16: bool = True,
                    max_wav_value: float = 32768.0, launch_cost: int = 400, mode: str = "auto") -> List[torch.Tensor]:
    """Vocode a padded batch ``mels`` [B,80,T] computing only what ``sample_lengths`` (samples per
    utterance, as ``vocoder_infer``'s ``lengths``) keeps.  Returns one 1-D device tensor per utterance,
    already trimmed — equal to ``generator(mels)[i, 0, :sample_lengths[i]]`` (int16 path: to the
    fused scale-and-truncate of ``generate_int16``).

    mode: "kernel" (compacted tiles inside one forward), "buckets" (dense forward per length bucket)
    or "auto" (kernel when the generator has it)."""
    B, _, T = mels.shape
    if len(sample_lengths) != B:
        raise ValueError(f"{len(sample_lengths)} lengths for a batch of {B}")
    if mode not in ("auto", "kernel", "buckets"):
        raise ValueError(f"unknown mode {mode!r}")
    hop = generator.hop_length
    total = T * hop
    keep = [max(0, min(int(n), total)) for n in sample_lengths]
    frames = [max(1, -(-k // hop)) for k in keep]
    out: List[Optional[torch.Tensor]] = [None] * B
    if mode == "kernel" or (mode == "auto" and hasattr(generator, "forward_ragged")):
        from . import _native

        order = sorted(range(B), key=lambda i: (-frames[i], i))  # groups of similar length when B > 64
        for g0 in range(0, B, _native.MAX_RAGGED_ITEMS):
            group = sorted(order[g0:g0 + _native.MAX_RAGGED_ITEMS])
            whole = len(group) == B
            x = mels if whole else mels.index_select(0, torch.as_tensor(group, device=mels.device))
            tg = max(frames[i] for i in group)
            tg = min(T, tg + _halo_of(generator))
            fr = [frames[i] for i in group]
            y = (generator.generate_int16(x[:, :, :tg], max_wav_value, frames=fr) if out_int16
                 else generator.forward_ragged(x[:, :, :tg], fr))
            for row, i in enumerate(group):
                out[i] = y[row, 0, :keep[i]]
        return out  # type: ignore[return-value]
    halo = _halo_of(generator)
    run = (lambda x: generator.generate_int16(x, max_wav_value)) if out_int16 else generator
    for bucket in plan_length_buckets(frames, halo, T, launch_cost):
        tb = bucket_extent(frames, bucket, halo, T)
        idx = torch.as_tensor(bucket, device=mels.device)
        y = run(mels.index_select(0, idx)[:, :, :tb])
        for row, i in enumerate(bucket):
            out[i] = y[row, 0, :keep[i]]
    return out  # type: ignore[return-value]
