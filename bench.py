#!/usr/bin/env python
"""bench.py — HiFi-GAN V1 vocoder throughput (generated audio-seconds per second) on N B200s.

    python bench.py --gpus 1 --steps 10 --warmup 3                    # this repo's CUDA path
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W         # N > 1, one rank per GPU
    python bench.py --impl reference --steps 3 --warmup 1             # the reference's CPU arithmetic

Workload (BASELINE.json configs[1]): HiFi-GAN V1, random-init weights (seed 1234), batch 16 x 800
synthetic mel frames (22.05 kHz, hop 256) PER GPU — utterance-sharded, so per-GPU work is fixed as N
grows ("weak" scaling) and no collective touches the data path.  One "step" is one forward over the
batch.  `value` is timed with CUDA events with the mels already resident in HBM; `e2e` goes through
the public wrapper (HIFIapi.generate) with pinned HOST mels and a HOST int16 result every step.

The JSON line also carries `roofline` (tcgen05 conv kernels: algorithmic FLOPs / CUDA-event time,
against the measured bf16 peak in MEASURED_PEAKS.json) and `cpu_baseline` (the oracle's torch port
of the reference arithmetic timed on this box's host cores on a bounded sample).
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SR, HOP = 22050, 256
B_PER_GPU, T_FRAMES = 16, 800
FLOP_PER_FRAME_V1 = 614_105_088  # SURVEY.md §8d / BASELINE.md §2 (2 x 307 052 544 MAC)
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def v1_h():
    from tts_king_b200.hifiapi import AttrDict

    return AttrDict(resblock="1", upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4],
                    upsample_initial_channel=512, resblock_kernel_sizes=[3, 7, 11],
                    resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]], MAX_WAV_VALUE=32768, weights_path=None)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        d["_source"] = "measured"
        return d
    d = dict(FALLBACK_PEAKS)
    d["_source"] = "fallback"
    return d


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_step(n_utt: int, frames: int, threads: int):
    """The reference's CPU arithmetic (oracle.torch_oracle: the ATen calls hifi/models.py dispatches,
    folded weights, no_grad — what HIFIapi.generate runs, hifiapi.py:47-49) on `n_utt` x `frames`."""
    import torch

    from oracle import fixtures as fx
    from oracle import torch_oracle
    from tts_king_b200.hifi.models import Generator

    torch.set_num_threads(threads)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.manual_seed(1234)
        m = Generator(fx.make_h(fx.V1))
        with contextlib.redirect_stdout(io.StringIO()):
            m.remove_weight_norm()
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    mel = fx.synthetic_mel(n_utt, frames, seed=7)

    def step():
        y = torch_oracle.forward(fx.V1, sd, mel)
        step.last_float, step.mel = y, mel  # kept for the fp32-accuracy record (checker use, outside any timed GPU region)
        return (y * 32768).numpy().astype("int16")

    return step, n_utt * frames * HOP / SR


def run_reference(args):
    """--impl reference: the reference's CPU path on this box's host cores.  The reference itself is
    a Python package that cannot travel to the GPU box (and is not pip-installable: no setup.py), so
    this arm times the oracle's port of it — kind "port" — with every host thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    step, audio_s = cpu_reference_step(1, T_FRAMES, threads)  # bounded sample: 1 utterance of the batch
    for _ in range(max(1, args.warmup)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = audio_s / dt
    sample = f"1 utterance x {T_FRAMES} frames per step (of the 16 x {T_FRAMES} batch); CPU throughput falls with batch"
    print(json.dumps({
        "impl": "reference", "metric": "audio_sec_per_sec", "value": val, "unit": "audio-s/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"HiFi-GAN V1 batch {B_PER_GPU} x {T_FRAMES} frames per GPU (22.05 kHz, hop 256)",
                   "sample": sample},
        "cpu_baseline": {"value": val, "unit": "audio-s/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def conv_flops(row, B, frames_in):
    """Algorithmic FLOPs of one launch (2 x MACs, all taps incl. padded positions;
    ConvTranspose counted as L_in*C_in*C_out*k)."""
    return 2.0 * B * frames_in * row["c_in"] * row["c_out"] * row["k"]


def layer_bytes(row, names, B, T, rates, e_a=2):
    """Algorithmic HBM bytes of one launch under this implementation's dataflow: operand
    copies `a` are e_a bytes per element (2 = one bf16 plane, 4 = hi + lo planes), the residual stream is
    fp32; unique elements read + written, halo re-reads and weights excluded."""
    import math

    n = row["name"]
    if row["kind"] < 0:
        return B * T * (80 * 4 + 128 * e_a)
    cin, cout = row["c_in"], row["c_out"]
    if n == "conv_pre":
        return B * T * (128 * e_a + cout * e_a)
    if n.startswith("ups."):
        i = int(n.split(".")[1])
        lin = T * math.prod(rates[:i])
        return B * (lin * cin * e_a + lin * row["stride"] * cout * (4 + e_a))
    if n == "conv_post":
        return B * T * math.prod(rates) * (cin * 4 + 4)
    blk = int(n.split(".")[1])
    nk = 3  # ResBlocks per stage (num_kernels of V1 / V2-style configs)
    i, j = blk // nk, blk % nk
    e = B * T * math.prod(rates[:i + 1]) * cout
    if ".convs1." in n:
        return e * (e_a + e_a)                       # a in, t out
    by = e * (e_a + 4)                               # operand in + residual in
    if not n.endswith(".2"):
        return by + e * (4 + e_a)                    # x out + a out
    by += e * (4 if j > 0 else 0)                    # last pair of the block: MRF running sum in
    by += e * (4 if (j < nk - 1 or i == len(rates) - 1) else e_a)  # sum / x out, or the next stage's operand
    return by


def layer_bytes_8d(row, names, B, T, rates, e=2):
    """Algorithmic HBM bytes of one launch by SURVEY.md §8(d)'s own rule: (unique input + unique output
    elements) x the storage dtype (e = 2 for the bf16 path, 4 for fp32) + one more C*L*e for a fused residual
    read and for a fused MRF accumulate read; a fused ResBlock-pair launch counts its own in / out / residual
    only (the intermediate never leaves the SM).  Weights and halo re-reads excluded."""
    import math

    n = row["name"]
    if row["kind"] < 0:
        return B * T * 80 * (4 + e)
    cin, cout = row["c_in"], row["c_out"]
    if n == "conv_pre":
        return B * T * (80 + cout) * e
    if n.startswith("ups."):
        i = int(n.split(".")[1])
        lin = T * math.prod(rates[:i])
        return B * lin * (cin + row["stride"] * cout) * e
    if n == "conv_post":
        return B * T * math.prod(rates) * (cin + 1) * e
    blk = int(n.split(".")[1])
    nk = 3
    i, j = blk // nk, blk % nk
    el = B * T * math.prod(rates[:i + 1]) * cout
    if ".convs1." in n:
        return 2 * el * e
    by = 3 * el * e                                   # in + out + fused residual read
    if n.endswith(".2") and j > 0:
        by += el * e                                  # fused MRF accumulate read
    return by


def timed_max_over_ranks(fn, reps, world, dist, dev):
    """Mean device time of `reps` calls of fn (CUDA events on the current stream, barrier + synchronize on
    both sides), MAX over ranks; returns (ms per call, last result)."""
    import torch

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = None
    for _ in range(reps):
        out = fn()
    e1.record()
    sync_all()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), out


def run_cfg3_strong(gen, dev, world, rank, dist, reps=3):
    """BASELINE cfg-3: 64 utterances x 2000 frames in TOTAL, utterance-sharded over the ranks (greedy
    longest-first, tts_king_b200.parallel.shard_utterances) — strong scaling, no data-path collective."""
    import torch

    from tts_king_b200 import parallel

    lengths = [2000] * 64
    mine = parallel.shard_utterances(lengths, world)[rank]
    mel = torch.randn(len(mine), 80, 2000, generator=torch.Generator().manual_seed(300 + rank)).to(dev)
    with torch.no_grad():
        y = gen(mel)                                  # warm-up (workspace, descriptors)
        ok = bool(torch.equal(y[:1], gen(mel[:1])))   # an item's samples do not depend on its batch
        del y
        sampler = ClockSampler(dev.index)
        if rank == 0:
            sampler.start()
        ms, _ = timed_max_over_ranks(lambda: gen(mel), reps, world, dist, dev)
        clocks = sampler.stop() if rank == 0 else None
    flags = torch.tensor([1 if ok else 0], device=dev)
    if world > 1:
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    audio_s = sum(lengths) * HOP / SR
    del mel
    torch.cuda.empty_cache()
    return {"workload": "64 utterances x 2000 frames in total, utterance-sharded", "scaling": "strong", "n_gpus": world,
            "utterances_per_rank": len(mine), "ms": ms, "reps": reps, "audio_sec_per_sec": audio_s / (ms * 1e-3),
            "item_bitwise_equal_to_its_single_forward": bool(flags.item()), "clocks": clocks}


def run_long_form(gen, dev, world, rank, dist, reps=2, T=310078, sub=16384):
    """BASELINE cfg-5: ONE 60-minute mel (310 078 frames) time-sharded over the ranks.  Every step: 13-frame
    mel halo exchange with the neighbours (NCCL isend/irecv), the rank's frames vocoded in 16 384-frame
    chunks, and the waveform collected on rank 0 — (a) with an NCCL gather, (b) with the gather fused into
    conv_post's stores (rank 0's buffer peer-mapped over NVLink, CUDA IPC).  Rank 0 then recomputes, on
    its own GPU alone, the samples around every rank boundary and both ends and compares bit for bit."""
    import torch

    from tts_king_b200 import parallel

    halo, hop = gen.halo_frames, gen.hop_length
    mel_full = torch.randn(1, 80, T, generator=torch.Generator().manual_seed(5000))  # same on every rank
    chunks = parallel.plan_time_chunks(T, world, 0)
    c = chunks[rank]
    local_mel = mel_full[:, :, c.start:c.stop].contiguous().to(dev)  # the mel arrives already time-sharded
    audio_s = T * hop / SR
    rec = {"workload": f"one mel of {T} frames (60 min at 22.05 kHz), time-sharded", "scaling": "strong", "n_gpus": world,
           "frames_per_rank": c.frames, "chunk_frames": sub, "halo_frames": halo,
           "halo_bytes": 0 if world == 1 else 2 * (world - 1) * halo * 80 * 4,
           "gather_bytes": 0 if world == 1 else int(sum(ch.frames for ch in chunks[1:])) * hop * 4, "reps": reps}

    def check(wav):
        """rank 0: windows straddling every rank boundary and both ends, recomputed on one GPU."""
        ok = True
        with torch.no_grad():
            for e in [0] + [ch.stop for ch in chunks[:-1]] + [T]:
                a, b = max(0, e - 40), min(T, e + 40)
                lo, hi = max(0, a - halo), min(T, b + halo)
                w = gen(mel_full[:, :, lo:hi].to(dev))
                ok = ok and bool(torch.equal(wav[:, :, a * hop:b * hop], w[:, :, (a - lo) * hop:(b - lo) * hop]))
        return ok

    with torch.no_grad():
        if world == 1:
            out = torch.empty((1, 1, T * hop), device=dev)
            fn = lambda: parallel.chunked_forward_into(gen, local_mel, sub, halo, out)  # noqa: E731
            fn()
            ms, _ = timed_max_over_ranks(fn, reps, world, dist, dev)
            rec.update(ms=ms, audio_sec_per_sec=audio_s / (ms * 1e-3), variant="single GPU, chunked, conv_post stores in place",
                       bitwise_equal_to_single_gpu=check(out))
            return rec

        def run_nccl():
            padded, left, right = parallel.exchange_halo(local_mel, halo)
            y = parallel.chunked_forward(gen, padded, sub, halo, hop)
            y = y[..., left * hop: y.shape[-1] - right * hop]
            return parallel.gather_wav(y, dst=0)

        run_nccl()
        ms_nccl, wav = timed_max_over_ranks(run_nccl, reps, world, dist, dev)
        ok_nccl = check(wav) if rank == 0 else True
        del wav
        full = parallel.share_output_buffer((1, 1, T * hop), torch.float32, owner=0)
        run_direct = lambda: parallel.sharded_long_form_into(gen, local_mel, halo, full, c.start, chunk_frames=sub)  # noqa: E731
        run_direct()
        ms_direct, _ = timed_max_over_ranks(run_direct, reps, world, dist, dev)
        ok_direct = check(full) if rank == 0 else True
        dist.barrier()
        del full
    flags = torch.tensor([1 if (ok_nccl and ok_direct) else 0], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    ms = min(ms_nccl, ms_direct)
    rec.update(ms=ms, audio_sec_per_sec=audio_s / (ms * 1e-3), ms_nccl_halo_and_gather=ms_nccl,
               ms_halo_and_direct_nvlink_store=ms_direct,
               variant="nccl gather" if ms_nccl <= ms_direct else "direct NVLink store from conv_post",
               bitwise_equal_to_single_gpu=bool(flags.item()))
    torch.cuda.empty_cache()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the cfg-3 / cfg-5 / fp32 records")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from tts_king_b200.hifiapi import AttrDict, HIFIapi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = AttrDict(hifi=v1_h(), model_config=AttrDict(vocoder=AttrDict(use_cpu=True)))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.manual_seed(1234)  # the reference's hifi.seed (config.yaml:23): same random-init weights
        with contextlib.redirect_stdout(io.StringIO()):
            api = HIFIapi(cfg, "cpu", compute_device=dev, precision=args.precision)
    gen = api.model
    g = torch.Generator().manual_seed(7 + rank)
    mel_host = torch.randn(B_PER_GPU, 80, T_FRAMES, generator=g).pin_memory()
    mel_dev = mel_host.to(dev)
    audio_s_step = B_PER_GPU * T_FRAMES * HOP / SR  # per rank

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---------------- device-resident throughput (`value`)
    with torch.no_grad():
        for _ in range(args.warmup):
            y = gen(mel_dev)
        sync_all()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            y = gen(mel_dev)
        e1.record()
        sync_all()
        ms_total = e0.elapsed_time(e1)
        clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = world * audio_s_step / (ms_step * 1e-3)

    # ---------------- end to end through the public wrapper, host buffers both ways
    for _ in range(2):
        wav = api.generate(mel_host)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        wav = api.generate(mel_host)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e = world * audio_s_step / (float(t.item()) / args.steps)
    h2d = mel_host.numel() * 4
    d2h = int(wav.size) * 2

    # ---------------- per-launch roofline pass (CUDA events around every launch, rank 0)
    roofline, layers_out = None, None
    if rank == 0:
        peaks = load_peaks()
        rows = gen.profile_layers(mel_dev)  # warm pass
        rows = gen.profile_layers(mel_dev)
        rates = [8, 8, 2, 2]
        tc_flops = tc_ms = all_ms = 0.0
        names = {r["name"] for r in rows}
        for r in rows:
            all_ms += r["ms"]
            if r["kind"] < 0:
                continue
            name = r["name"]
            if name == "conv_pre":
                frames_in = T_FRAMES
            elif name.startswith("ups."):
                i = int(name.split(".")[1]); frames_in = T_FRAMES * int(__import__("math").prod(rates[:i]))
            elif name == "conv_post":
                frames_in = T_FRAMES * HOP
            else:
                i = int(name.split(".")[1]) // 3; frames_in = T_FRAMES * int(__import__("math").prod(rates[:i + 1]))
            r["flops"] = conv_flops(r, B_PER_GPU, frames_in)
            # a fused ResBlock-pair launch is reported under its second conv: credit the first conv too
            if ".convs2." in name and name.replace(".convs2.", ".convs1.") not in names:
                r["flops"] *= 2.0
                r["fused_pair"] = True
            r["tflops"] = r["flops"] / (r["ms"] * 1e-3) / 1e12 if r["ms"] > 0 else None
            if r.get("tensor_core"):
                tc_flops += r["flops"]; tc_ms += r["ms"]
        # every launch against its own roofline: max(FLOPs / tensor peak, algorithmic bytes / HBM peak)
        e_a = 4 if args.precision == "fp32" else 2
        tf_peak = float(peaks.get("bf16_tflops_sustained", FALLBACK_PEAKS["bf16_tflops_sustained"])) * 1e12
        bw_peak = float(peaks.get("hbm_gbs", 6555.8)) * 1e9
        roof_ms = roof8d_ms = 0.0
        for r in rows:
            r["bytes"] = layer_bytes(r, names, B_PER_GPU, T_FRAMES, rates, e_a)
            r["roofline_ms"] = max(r.get("flops", 0.0) / tf_peak, r["bytes"] / bw_peak) * 1e3
            roof_ms += r["roofline_ms"]
            r["bytes_8d"] = layer_bytes_8d(r, names, B_PER_GPU, T_FRAMES, rates, e_a)
            r["roofline_8d_ms"] = max(r.get("flops", 0.0) / tf_peak, r["bytes_8d"] / bw_peak) * 1e3
            roof8d_ms += r["roofline_8d_ms"]
        mult = 3.0 if args.precision == "fp32" else 1.0  # bf16x3: compensation passes are overhead, not credited
        peak = float(peaks.get("bf16_tflops_sustained", FALLBACK_PEAKS["bf16_tflops_sustained"]))
        # duration of the tcgen05 launches inside the TIMED region: the step time (CUDA events around the K steps) times
        # their share of the step; the per-launch event pass itself is slower than the step it dissects (an event after
        # every launch serialises what programmatic dependent launch overlaps) and is reported beside it
        share = tc_ms / all_ms if all_ms else 0.0
        ach_serial = tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
        ach = tc_flops / (ms_step * share * 1e-3) / 1e12 if share > 0 else 0.0
        traffic = None  # DRAM bytes of the tcgen05 launches of one step, from the committed ncu launch list
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", f"traffic_{args.precision}.json")))
            traffic = tj["tcgen05_kernels"]["dram_bytes_per_step"]
        except Exception:
            pass
        roofline = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "traffic": traffic, "traffic_note": "DRAM bytes per step summed over the tcgen05 launches (ncu), not per launch",
                    "hbm_floor_ms": (traffic / (float(peaks.get("hbm_gbs", 6555.8)) * 1e9) * 1e3) if traffic else None,
                    "tensor_floor_ms": tc_flops / (peak * 1e12) * 1e3, "measured_ms": ms_step * share,
                    "per_launch_event_pass": {"measured_ms": tc_ms, "achieved": ach_serial, "frac": ach_serial / peak,
                                              "note": "same launches timed one by one (an event after each): no launch overlap"}, "kernel": "conv_tc2_kernel + conv_pair_fold_kernel + conv_pair_tc_kernel + conv_tc_kernel (tcgen05 implicit-GEMM convs, all 59 launches of a step)",
                    "share_of_step": tc_ms / all_ms if all_ms else None, "peak_source": peaks["_source"] + " sustained bf16",
                    "mma_passes_per_product": mult,
                    "per_layer": {"sum_of_launch_rooflines_ms": roof_ms, "sum_of_launch_times_ms": all_ms,
                                  "frac": roof_ms / all_ms if all_ms else None,
                                  "note": "each launch's roofline = max(algorithmic FLOPs / tensor peak, bytes / HBM peak); `frac` counts "
                                          "the bytes of THIS implementation's dataflow (fp32 residual stream + a separate bf16 operand copy), "
                                          "`frac_8d` counts SURVEY.md §8(d)'s algorithmic bytes (every tensor once, in the path's storage dtype)",
                                  "sum_of_launch_rooflines_8d_ms": roof8d_ms, "frac_8d": roof8d_ms / all_ms if all_ms else None}}
        layers_out = rows
        try:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"bench_layers_{args.precision}.json"), "w"), indent=1)
        except Exception:
            pass

    # ---------------- CPU baseline on this box's host cores (rank 0, N = 1 only)
    cpu_baseline = step = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        step, a_s = cpu_reference_step(1, T_FRAMES, threads)
        step()
        best = 1e30
        t_end = time.perf_counter() + 20.0
        reps = 0
        while reps < 3 or (time.perf_counter() < t_end and reps < 8):
            t0 = time.perf_counter(); step(); best = min(best, time.perf_counter() - t0); reps += 1
        cpu_baseline = {"value": a_s / best, "unit": "audio-s/s", "cores": threads, "kind": "port",
                        "sample": f"1 utterance x {T_FRAMES} frames, best of {reps} (torch/oneDNN fp32, "
                                  "oracle.torch_oracle = the ATen calls hifi/models.py dispatches)"}

    # ---------------- the other BASELINE configs, so that the driver's N = 1/2/4/8 runs carry them: the
    # fp32-accurate arithmetic mode on the same workload, cfg-3 (strong scaling over utterances) and cfg-5
    # (one long mel, halo exchange + gather) — every rank takes part
    fp32_rec = cfg3 = long_form = None
    if not args.no_extras:
        if args.precision == "bf16":
            gen.precision = "fp32"
            with torch.no_grad():
                for _ in range(3):
                    gen(mel_dev)
                ms32, _ = timed_max_over_ranks(lambda: gen(mel_dev), 10, world, dist if world > 1 else None, dev)
                err = None
                if cpu_baseline is not None:  # N = 1: the oracle's waveform of the timed CPU sample is at hand
                    y = gen(step.mel.to(dev)).cpu()
                    err = float((y - step.last_float).abs().max())
            gen.precision = "bf16"
            peaks = load_peaks()
            tf = world * B_PER_GPU * T_FRAMES * FLOP_PER_FRAME_V1 / (ms32 * 1e-3) / 1e12
            pk = float(peaks.get("bf16_tflops_sustained", FALLBACK_PEAKS["bf16_tflops_sustained"])) * world
            fp32_rec = {"value": world * audio_s_step / (ms32 * 1e-3), "unit": "audio-s/s", "ms_per_step": ms32, "steps": 10,
                        "dtype": "bf16x3 split products, fp32 accumulate (fp32-accurate mode, north_star <= 1e-4)",
                        "max_abs_vs_oracle": err, "max_abs_sample": "1 utterance x 800 frames (seed 7) vs oracle.torch_oracle" if err is not None else None,
                        "tflops_algorithmic": tf, "roofline": {"bound": "tensor", "achieved": tf, "peak": pk, "unit": "TFLOP/s", "frac": tf / pk,
                                                               "note": "algorithmic FLOPs only; the 3 bf16 passes per product are overhead"}}
        cfg3 = run_cfg3_strong(gen, dev, world, rank, dist if world > 1 else None)
        long_form = run_long_form(gen, dev, world, rank, dist if world > 1 else None)

    if rank == 0:
        launches = gen.kernel_launches(B_PER_GPU, T_FRAMES)
        print(json.dumps({
            "metric": "audio_sec_per_sec", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "bf16x3(fp32-accurate)",
            "data": "synthetic",
            "config": {"workload": f"HiFi-GAN V1 batch {B_PER_GPU} x {T_FRAMES} frames per GPU (22.05 kHz, hop 256), "
                                   "random-init seed 1234", "parallelism": f"utterance-sharded x{world}, no data-path collective",
                       "precision": args.precision,
                       "l2": "no flush: per-step activation working set (~2.5 GB) is 20x the 126 MB L2"},
            "tflops_algorithmic": world * B_PER_GPU * T_FRAMES * FLOP_PER_FRAME_V1 / (ms_step * 1e-3) / 1e12,
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "audio-s/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "HIFIapi.generate(host mel) -> host int16"},
            "gpu_launches": launches * args.steps,
            "roofline": roofline, "cpu_baseline": cpu_baseline,
            "fp32": fp32_rec, "cfg3_strong": cfg3, "long_form": long_form,
        }))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
