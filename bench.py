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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
        roof_ms = 0.0
        for r in rows:
            r["bytes"] = layer_bytes(r, names, B_PER_GPU, T_FRAMES, rates, e_a)
            r["roofline_ms"] = max(r.get("flops", 0.0) / tf_peak, r["bytes"] / bw_peak) * 1e3
            roof_ms += r["roofline_ms"]
        mult = 3.0 if args.precision == "fp32" else 1.0  # bf16x3: compensation passes are overhead, not credited
        peak = float(peaks.get("bf16_tflops_sustained", FALLBACK_PEAKS["bf16_tflops_sustained"]))
        ach = tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
        traffic = None  # DRAM bytes of the tcgen05 launches of one step, from the committed ncu launch list
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", f"traffic_{args.precision}.json")))
            traffic = tj["tcgen05_kernels"]["dram_bytes_per_step"]
        except Exception:
            pass
        roofline = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "traffic": traffic, "traffic_note": "DRAM bytes per step summed over the tcgen05 launches (ncu), not per launch",
                    "hbm_floor_ms": (traffic / (float(peaks.get("hbm_gbs", 6555.8)) * 1e9) * 1e3) if traffic else None,
                    "tensor_floor_ms": tc_flops / (peak * 1e12) * 1e3, "measured_ms": tc_ms, "kernel": "conv_tc2_kernel + conv_pair_tc_kernel + conv_tc_kernel (tcgen05 implicit-GEMM convs, all 59 launches of a step)",
                    "share_of_step": tc_ms / all_ms if all_ms else None, "peak_source": peaks["_source"] + " sustained bf16",
                    "mma_passes_per_product": mult,
                    "per_layer": {"sum_of_launch_rooflines_ms": roof_ms, "sum_of_launch_times_ms": all_ms,
                                  "frac": roof_ms / all_ms if all_ms else None,
                                  "note": "each launch's roofline = max(algorithmic FLOPs / tensor peak, algorithmic bytes / HBM peak)"}}
        layers_out = rows
        try:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"bench_layers_{args.precision}.json"), "w"), indent=1)
        except Exception:
            pass

    # ---------------- CPU baseline on this box's host cores (rank 0, N = 1 only)
    cpu_baseline = None
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
        }))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
