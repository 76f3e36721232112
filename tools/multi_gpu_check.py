#!/usr/bin/env python
"""Multi-GPU parity check, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/multi_gpu_check.py

1. utterance sharding: a ragged set of utterances split across ranks (greedy longest-first) must
   reproduce, bit for bit, what one GPU computes for the same utterances;
2. one long mel time-sharded across ranks: 13-frame halo exchange over NCCL + gather must equal the
   single-GPU forward of the whole mel;
3. the same with the gather fused into the last kernel (every rank stores its samples straight into
   rank 0's buffer through a peer mapping, over NVLink): bit-identical, float and int16.
Rank 0 prints one JSON line.
"""
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


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    out = {"world": world}
    for prec in ("fp32", "bf16"):
        m = make_generator(fx.V1, precision=prec).to(dev)
        h = fx.make_h(fx.V1)
        halo, hop = parallel.halo_frames(h), parallel.hop_length(h)
        # ---- 1. utterance sharding
        rng = np.random.default_rng(0)
        lengths = rng.integers(40, 200, size=3 * world + 1).tolist()
        mine = parallel.shard_utterances(lengths, world)[rank]
        with torch.no_grad():
            outs = {i: m(fx.synthetic_mel(1, lengths[i], seed=100 + i).to(dev)).cpu().numpy() for i in mine}
        gathered = [None] * world
        dist.all_gather_object(gathered, outs)
        # ---- 2. one long mel, time-sharded
        T = 64 * world + 37
        mel = fx.synthetic_mel(1, T, seed=5)
        chunks = parallel.plan_time_chunks(T, world, 0)
        local_mel = mel[:, :, chunks[rank].start:chunks[rank].stop].contiguous().to(dev)
        with torch.no_grad():
            wav_local = parallel.sharded_long_form(m, local_mel, halo, hop)
        wav = parallel.gather_wav(wav_local, dst=0)
        # ---- 3. the same, with the gather fused into the last kernel: every rank's conv_post stores its samples
        # straight into rank 0's buffer over NVLink (CUDA IPC peer mapping), int16 tail included
        direct = parallel.share_output_buffer((1, 1, T * hop), torch.float32, owner=0)
        direct16 = parallel.share_output_buffer((1, 1, T * hop), torch.int16, owner=0)
        with torch.no_grad():
            parallel.sharded_long_form_into(m, local_mel, halo, direct, chunks[rank].start, chunk_frames=29)
            parallel.sharded_long_form_into(m, local_mel, halo, direct16, chunks[rank].start, chunk_frames=64)
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            merged = {}
            for d in gathered:
                merged.update(d)
            with torch.no_grad():
                ok = sorted(merged) == list(range(len(lengths))) and all(
                    np.array_equal(merged[i], m(fx.synthetic_mel(1, lengths[i], seed=100 + i).to(dev)).cpu().numpy())
                    for i in merged)
                full = m(mel.to(dev))
                full16 = m.generate_int16(mel.to(dev))
            err = float((wav - full).abs().max())
            out[prec] = {"direct_p2p_store_bitwise_equal": bool(torch.equal(direct, full)) and bool(torch.equal(direct16, full16)),
                         "direct_buffer_device_on_this_rank": str(direct.device),"utterance_sharding_bitwise_equal": bool(ok), "n_utterances": len(lengths),
                         "long_form_frames": T, "long_form_max_abs_vs_single_gpu": err,
                         "halo_frames": halo, "halo_bytes_per_side": halo * 80 * 4}
        dist.barrier()
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
