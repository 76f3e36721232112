#!/usr/bin/env python
"""BASELINE.json configs 3 and 5 on N GPUs (one process per GPU, torchrun):

  cfg-3  64 utterances x 2000 frames, utterance-sharded (greedy longest-first) — strong scaling;
         plus a ragged variant (lengths ~U[400,2000]) that exercises the sharder
  cfg-5  one 60-minute mel (310 078 frames) time-chunked across the ranks with a 13-frame halo
         exchanged over NCCL, waveform gathered on rank 0

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29541 tools/scaling_configs.py [--precision bf16]

Timing: CUDA events on every rank around the rank-local work (inputs resident in HBM), barrier +
synchronize on both sides, max over ranks.  Rank 0 prints one JSON line per config.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import fixtures as fx  # noqa: E402  (weight / input recipes only)
from tts_king_b200 import parallel  # noqa: E402
from _util import make_generator  # noqa: E402

SR, HOP = 22050, 256


def timed(fn, dev, world, reps=3):
    best = None
    for _ in range(reps):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record()
        torch.cuda.synchronize(); dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = float(t) if best is None else min(best, float(t))
    return best, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--skip-long", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    m = make_generator(fx.V1, precision=args.precision).to(dev)
    h = fx.make_h(fx.V1)
    halo, hop = parallel.halo_frames(h), parallel.hop_length(h)

    # ---------------- cfg-3: 64 x 2000 frames, equal lengths and ragged
    for name, lengths in (("cfg3_64x2000", [2000] * 64),
                          ("cfg3_ragged", np.random.default_rng(3).integers(400, 2001, size=64).tolist())):
        mine = parallel.shard_utterances(lengths, world)[rank]
        # the reference pads a batch to its longest item and computes on the padding; here
        # each rank batches its own utterances by exact length groups (no padded frames)
        groups = {}
        for i in mine:
            groups.setdefault(lengths[i], []).append(i)
        mels = {L: torch.randn(len(ix), 80, L, generator=torch.Generator().manual_seed(L)).to(dev) for L, ix in groups.items()}

        def run():
            with torch.no_grad():
                return [m(x) for x in mels.values()]

        run()
        ms, _ = timed(run, dev, world)
        audio_s = sum(lengths) * HOP / SR
        if rank == 0:
            print(json.dumps({"config": name, "n_gpus": world, "precision": args.precision, "ms": ms,
                              "audio_s": audio_s, "audio_sec_per_sec": audio_s / (ms * 1e-3),
                              "frames_per_rank_max": max(sum(lengths[i] for i in s) for s in parallel.shard_utterances(lengths, world))}))

    # ---------------- cfg-3 ragged again, the way synth_samples batches it (padded to the longest item), with the
    # padding skipped inside the kernels (Generator.forward_ragged) — one forward per rank
    lengths = np.random.default_rng(3).integers(400, 2001, size=64).tolist()
    mine = parallel.shard_utterances(lengths, world)[rank]
    fr = [lengths[i] for i in mine]
    padded = torch.zeros(len(mine), 80, max(fr))
    for row, L in enumerate(fr):
        padded[row, :, :L] = torch.randn(80, L, generator=torch.Generator().manual_seed(L))
    padded = padded.to(dev)

    def run_ragged():
        return m.forward_ragged(padded, fr)

    run_ragged()
    ms, _ = timed(run_ragged, dev, world)
    if rank == 0:
        audio_s = sum(lengths) * HOP / SR
        print(json.dumps({"config": "cfg3_ragged_padded_batch_kernel_skip", "n_gpus": world, "precision": args.precision, "ms": ms,
                          "audio_s": audio_s, "audio_sec_per_sec": audio_s / (ms * 1e-3)}))

    # ---------------- cfg-5: one 60-minute mel, time-chunked with halo exchange
    if not args.skip_long:
        T = 310078
        chunks = parallel.plan_time_chunks(T, world, 0)
        c = chunks[rank]
        g = torch.Generator().manual_seed(1000 + rank)
        local_mel = torch.randn(1, 80, c.frames, generator=g).to(dev)  # the mel is already time-sharded
        sub = 16384  # frames per forward on one GPU: bounds activation memory (the long chunk is itself chunked)

        def run_long():
            with torch.no_grad():
                padded, left, right = parallel.exchange_halo(local_mel, halo)
                # rank-local chunking of the (haloed) slice; interior cuts get their own 13-frame halos
                y = parallel.chunked_forward(m, padded, sub, halo, hop)
                y = y[..., left * hop: y.shape[-1] - right * hop]
                return parallel.gather_wav(y, dst=0)

        run_long()
        ms, wav = timed(run_long, dev, world, reps=2)
        if rank == 0:
            audio_s = T * HOP / SR
            print(json.dumps({"config": "cfg5_60min_time_chunked", "n_gpus": world, "precision": args.precision, "ms": ms,
                              "audio_s": audio_s, "audio_sec_per_sec": audio_s / (ms * 1e-3), "frames": T,
                              "frames_per_rank": c.frames, "halo_frames": halo, "halo_bytes_per_side": halo * 80 * 4,
                              "wav_samples_gathered": int(wav.shape[-1]), "includes": "halo exchange + forward + gather to rank 0"}))
        # the same with the gather fused into the last kernel: every rank's conv_post stores its samples straight
        # into rank 0's buffer through a peer mapping (NVLink), float32 like the NCCL variant and int16 (the
        # HIFIapi.generate tail, half the bytes)
        for dt, label in ((torch.float32, "f32"), (torch.int16, "i16")):
            full = parallel.share_output_buffer((1, 1, T * hop), dt, owner=0)

            def run_direct():
                with torch.no_grad():
                    parallel.sharded_long_form_into(m, local_mel, halo, full, c.start, chunk_frames=sub)

            run_direct()
            ms, _ = timed(run_direct, dev, world, reps=2)
            if rank == 0:
                print(json.dumps({"config": f"cfg5_60min_direct_p2p_store_{label}", "n_gpus": world, "precision": args.precision, "ms": ms,
                                  "audio_s": audio_s, "audio_sec_per_sec": audio_s / (ms * 1e-3),
                                  "includes": "halo exchange + forward with conv_post storing into rank 0's buffer over NVLink (no gather collective)"}))
            dist.barrier()
            del full
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
