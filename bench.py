#!/usr/bin/env python
"""bench.py -- headline benchmark of the VSC22 hot path on B200 (contract in the task statement).

Primary line (BASELINE.json configs[1]): ViT-B/16@224 bf16 frame encoder, 10 000 synthetic frames per
step per GPU, metric ``frame-descriptors/sec``; a step = one pass over the 10k-frame job in batches of
256 through the encoder plan.  Secondary object ``"sim"`` (configs[2]): 10k x 40k x 512 score
normalisation + top-10, metric ``sim-pairs/sec``.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload both|encoder|sim]

N > 1: launched by torchrun, one rank per GPU; frames shard over ranks for encoding (no data-path
collective, weak scaling) and the reference bank shards for similarity (one all-gather of partial
top-k).  Timing: CUDA events on the launching stream, barrier + synchronize on both sides, MAX over
ranks.  ``--impl reference`` times the CPU port of the reference's path (oracle/, numpy/torch on the
host cores) on a bounded sample of the same workload; /root/reference itself does not exist on the
GPU box.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

N_FRAMES = 10_000          # configs[1]: 10k synthetic frames
BATCH = 512                # encoder plan chunk (max_frames): the 10k-frame job is walked in chunks of 512 frames
                           # (measured: 256 -> 22.2k, 384 -> 22.7k, 512 -> 23.1k, 768 -> 23.1k frames/s; tools/sweep_batch.py)
SIM_NQ, SIM_NR, SIM_NZ, SIM_D, SIM_K = 10_000, 40_000, 40_000, 512, 10   # configs[2]
SIM_NQ_C, SIM_NR_C = 10_000, 40_000                                     # candidate generation on the same shapes


def load_traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed ncu --set full
    captures (profiles/r02_traffic.json, else round 1's; numbers measured under the profiler, never timings)."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            v = json.load(open(os.path.join(REPO, "profiles", name))).get(kernel)
        except Exception:
            v = None
        if v is not None:
            return v
    return None


def load_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9 or not (t0 - 0.1 <= ts <= t1 + 0.3):
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def dist_setup(n_gpus):
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    else:
        torch.cuda.set_device(0)
    return world, rank, local


def barrier_sync(world):
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(ms, world):
    import torch
    if world == 1:
        return ms
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ------------------------------------------------------------------------------------------------ encoder
def bench_encoder(args, world, rank, peaks):
    import numpy as np
    import torch

    from vsc22_submission_b200 import _lib
    from vsc22_submission_b200.encoder import B200ViTEncoder, VIT_B16_224_GEM, random_weights
    spec = VIT_B16_224_GEM
    dev = torch.device("cuda", torch.cuda.current_device())
    enc = B200ViTEncoder(spec, random_weights(spec, seed=0), max_frames=BATCH).to(dev).eval()
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    frames = torch.empty((N_FRAMES, 3, spec.img, spec.img), dtype=torch.float32, device=dev)
    for i in range(0, N_FRAMES, 1000):     # 6.0 GB resident in HBM: >> L2, nothing is cached between steps
        frames[i:i + 1000] = torch.randn((min(1000, N_FRAMES - i), 3, spec.img, spec.img), generator=g, device=dev).clamp_(-1, 1)
    out = None
    for _ in range(args.warmup):
        out = enc(frames)
    barrier_sync(world)
    clocks = ClockSampler(torch.cuda.current_device())
    clocks.start()
    time.sleep(0.3)
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record()
    for _ in range(args.steps):
        out = enc(frames)
    e1.record()
    barrier_sync(world)
    t_wall1 = time.time()
    ms = max_over_ranks(e0.elapsed_time(e1), world)
    launches = _lib.launch_count() - n0
    clk = clocks.stop(t_wall0, t_wall1)
    ms_per_step = ms / args.steps
    value = N_FRAMES * world / (ms_per_step / 1e3)
    # ---- roofline pass: the same K steps again with a CUDA event pair around every launch (the event records cost
    #      ~3 % of the step, so they are kept out of the pass that defines `value`)
    _lib.prof_collect()
    _lib.prof_enable(True)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(args.steps):
        out = enc(frames)
    p1.record()
    barrier_sync(world)
    prof_ms = p0.elapsed_time(p1)
    _lib.prof_enable(False)
    prof = _lib.prof_collect()

    # ---- e2e: host (pinned) frames in, host descriptors out, copies inside the timed region
    e2e_frames = N_FRAMES
    host = None
    while host is None and e2e_frames >= BATCH:
        try:
            host = torch.empty((e2e_frames, 3, spec.img, spec.img), dtype=torch.float32, pin_memory=True)
        except RuntimeError:
            e2e_frames //= 2
    for i in range(0, e2e_frames, 1000):
        host[i:i + 1000].copy_(frames[i:i + 1000])
    torch.cuda.synchronize()
    host_np = host.numpy()
    for _ in range(min(args.warmup, 2)):
        enc.forward_host(host_np[:4 * BATCH], dev)
    barrier_sync(world)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out_host = enc.forward_host(host_np, dev)
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3, world) / args.steps
    barrier_sync(world)
    e2e = {"value": e2e_frames * world / (e2e_ms / 1e3), "unit": "frame-descriptors/sec",
           "h2d_bytes_per_step": int(e2e_frames * 3 * spec.img * spec.img * 4 * world),
           "d2h_bytes_per_step": int(e2e_frames * spec.out_dim * 4 * world), "ms_per_step": e2e_ms,
           "frames_per_step": e2e_frames * world, "api": "B200ViTEncoder.forward_host (vscb200_vit_forward_host)"}
    dev_vs_host = float(np.abs(out_host[:64] - out[:64].cpu().numpy()).max())

    gm = prof["gemm"]
    roof_peak = peaks["bf16_tflops_sustained"]
    gemm_tflops = gm["work"] / (gm["ms"] / 1e3) / 1e12 if gm["ms"] > 0 else 0.0
    roofline = {"bound": "tensor", "kernel": "gemm_bf16_kernel (tcgen05, all projections + patch embed)",
                "achieved": gemm_tflops, "peak": roof_peak, "unit": "TFLOP/s", "frac": gemm_tflops / roof_peak,
                "peak_source": f"{peaks['source']} bf16_tflops_sustained (kernel timed inside a long step)",
                "traffic": load_traffic("gemm_bf16_kernel"), "launches": gm["launches"],
                "share_of_step": gm["ms"] / prof_ms if prof_ms > 0 else None,
                "timed_with": "CUDA event pair around every launch, K steps right after the timed K steps (same inputs)",
                "ms_per_step_with_events": prof_ms / args.steps,
                "whole_step_tflops": value / world * spec.flops_per_frame() / 1e12,
                "whole_step_frac": value / world * spec.flops_per_frame() / 1e12 / roof_peak,
                "kernel_ms": {k: round(v["ms"] / args.steps, 3) for k, v in prof.items() if v["launches"]}}
    return {"value": value, "ms_per_step": ms_per_step, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk,
            "roofline": roofline, "dev_vs_host_maxabs": dev_vs_host, "enc": enc, "frames": frames, "out": out}


def bench_swin(args):
    """Second encoder family on the reference's path (swinv2_v106/107/115: SwinV2-B 256x256, config_v106.py:8-24):
    device-resident frames, 1024 synthetic frames per step, plan chunk 256, in both arithmetic modes; parity of 8 frames:
    the bf16 mode against the matched-precision oracle (and the fp32 figure beside it), the fp32-equivalent mode against
    the fp32 oracle, which reproduces the reference class bit for bit -- the 1e-3 contract."""
    import dataclasses

    import numpy as np
    import torch

    from oracle import swin_ref
    from vsc22_submission_b200.swin_encoder import B200SwinEncoder, SWINV2_B_256, random_weights
    dev = torch.device("cuda", torch.cuda.current_device())
    w = random_weights(SWINV2_B_256, seed=0)
    n = 1024
    frames = torch.randn((n, 3, 256, 256), generator=torch.Generator(device=dev).manual_seed(7), device=dev).clamp_(-1, 1)
    x8 = frames[:8].cpu()
    refs = {p: swin_ref.forward(swin_ref.SWINV2_B_256, w, x8, precision=p).numpy() for p in ("fp32", "bf16")}
    rel = lambda a, b: float((np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)).max())
    out = {}
    for mode in ("bf16", "fp32"):
        enc = B200SwinEncoder(dataclasses.replace(SWINV2_B_256, precision=mode), w, max_frames=256).to(dev).eval()
        for _ in range(2):
            y = enc(frames)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            y = enc(frames)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        got = y[:8].cpu().numpy()
        fps = n / (ms / 1e3)
        out[mode] = {"value": fps, "ms_per_step": ms, "tflops": fps * SWINV2_B_256.flops_per_frame() / 1e12,
                     "parity_rel_l2_max_vs_fp32_oracle": rel(got, refs["fp32"]),
                     "parity_rel_l2_max_vs_matched_precision_oracle": rel(got, refs["bf16"]) if mode == "bf16" else None}
        del enc, y
        torch.cuda.empty_cache()
    return {"metric": "frame-descriptors/sec", "value": out["bf16"]["value"], "ms_per_step": out["bf16"]["ms_per_step"],
            "frames_per_step": n, "tflops": out["bf16"]["tflops"], "flops_per_frame": SWINV2_B_256.flops_per_frame(),
            "parity_rel_l2_max_vs_fp32_oracle": out["bf16"]["parity_rel_l2_max_vs_fp32_oracle"],
            "modes": out, "parity_tolerance": 1e-3, "parity_frames": 8,
            "parity_contract": "fp32-equivalent mode <= 1e-3 vs the fp32 oracle (== the reference class); the bf16 mode sits on "
                               "the bf16-operand floor of this 24-block res-post-norm network",
            "config": {"workload": "SwinV2-B 256x256 window 16 (config_v106.py), random init, 1 GPU"}}


def bench_ingest(n=512, h=360, w=640, out=224):
    """SURVEY.md 8f row f4 (after JPEG decode): Resize(bicubic) + ToTensor + Normalize of D/infer/src/transform.py:20-43 for
    decoded 360 x 640 frames -- device kernels (uint8 frames uploaded from pinned memory inside the timed region) next to
    Pillow + numpy on one host core (what each DataLoader worker of the reference runs)."""
    import numpy as np
    import torch
    from PIL import Image

    from vsc22_submission_b200 import ingest
    rng = np.random.default_rng(0)
    frames = torch.from_numpy(rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)).pin_memory()
    pre = ingest.sscd_transform(out, out)
    dev_frames = frames.cuda()
    pre(dev_frames[:8])
    torch.cuda.synchronize()
    td, te = [], []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        y = pre(dev_frames)
        e1.record()
        torch.cuda.synchronize()
        td.append(e0.elapsed_time(e1))
        t0 = time.perf_counter()
        y2 = pre(frames)
        torch.cuda.synchronize()
        te.append((time.perf_counter() - t0) * 1e3)
    sample = 48
    mean, std = np.array(ingest.IMAGENET_MEAN, np.float32)[:, None, None], np.array(ingest.IMAGENET_STD, np.float32)[:, None, None]
    t0 = time.perf_counter()
    for i in range(sample):
        im = np.asarray(Image.fromarray(frames[i].numpy()).resize((out, out), Image.BICUBIC))
        ref = (im.transpose(2, 0, 1).astype(np.float32) / np.float32(255) - mean) / std
    cpu = sample / (time.perf_counter() - t0)
    same = bool(np.array_equal(ref, y[sample - 1].cpu().numpy()))
    bytes_alg = n * (h * w * 3 + out * out * 3 * 4)
    del dev_frames, y, y2
    torch.cuda.empty_cache()
    return {"workload": f"{n} decoded {h}x{w} RGB frames -> bicubic {out}x{out} + ToTensor + Normalize (float32 NCHW)",
            "frames_per_sec": n / (min(td) / 1e3), "ms": min(td),
            "hbm_frac_algorithmic": bytes_alg / (min(td) / 1e3) / 1e9 / load_peaks()["hbm_gbs"],
            "e2e": {"frames_per_sec": n / (min(te) / 1e3), "ms": min(te), "h2d_bytes": n * h * w * 3,
                    "api": "ingest.sscd_transform(224, 224)(pinned uint8 frames)"},
            "cpu_baseline": {"value": cpu, "unit": "frames/sec", "cores": 1, "kind": "reference",
                             "sample": f"{sample} frames through Pillow Image.resize(BICUBIC) + numpy ToTensor/Normalize"},
            "parity": {"bit_identical_to_pillow": same}}


def bench_jpeg(n=256, h=360, w=640, out=224):
    """SURVEY.md 8f row f4, the decode half: the JPEG files of n 360 x 640 frames (quality 90, 4:2:0, as ffmpeg writes them)
    -> RGB on the device (csrc/jpeg.cu) -> resize + normalise; host bytes in, [n, 3, 224, 224] float32 on the device out.
    Beside it: Pillow (libjpeg-turbo) decode + resize on one host core -- what a DataLoader worker of the reference runs
    (D/infer/src/dataset.py:137-141) -- on a bounded sample, and bit-identity of the decoded frames."""
    import io

    import numpy as np
    import torch
    from PIL import Image

    from vsc22_submission_b200 import ingest
    rng = np.random.default_rng(1)
    yy, xx = np.mgrid[0:h, 0:w]
    files = []
    for i in range(n):                              # smooth content + noise: ~60-80 KB per frame like real video frames
        img = np.stack([128 + 90 * np.sin(xx / (17.0 + i % 7) + yy / 29.0), 128 + 90 * np.cos(xx / 23.0 - yy / (11.0 + i % 5)),
                        (xx * 2 + yy * 3 + 5 * i) % 256], axis=-1) + rng.normal(0, 10, (h, w, 3))
        buf = io.BytesIO()
        Image.fromarray(np.clip(img, 0, 255).astype(np.uint8)).save(buf, format="JPEG", quality=90, subsampling=2)
        files.append(buf.getvalue())
    pre = ingest.sscd_transform(out, out)
    pre(files[:8])
    torch.cuda.synchronize()
    t_dec, t_all = [], []
    for _ in range(3):
        t0 = time.perf_counter()
        rgb = ingest.decode_jpeg_frames(files)
        torch.cuda.synchronize()
        t_dec.append((time.perf_counter() - t0) * 1e3)
        t0 = time.perf_counter()
        y = pre(files)
        torch.cuda.synchronize()
        t_all.append((time.perf_counter() - t0) * 1e3)
    files4 = files * 4                               # a segment is a serial bit stream: throughput grows with frames per call
    ingest.decode_jpeg_frames(files4[:8])
    t_big = []
    for _ in range(2):
        t0 = time.perf_counter()
        big = ingest.decode_jpeg_frames(files4)
        torch.cuda.synchronize()
        t_big.append((time.perf_counter() - t0) * 1e3)
    del big
    sample = 32
    t0 = time.perf_counter()
    for i in range(sample):
        ref = np.asarray(Image.open(io.BytesIO(files[i])).convert("RGB"))
        Image.fromarray(ref).resize((out, out), Image.BICUBIC)
    cpu = sample / (time.perf_counter() - t0)
    same = bool(np.array_equal(ref, rgb[sample - 1].cpu().numpy()))
    nbytes = sum(len(f) for f in files)
    return {"workload": f"{n} JPEG frames {h}x{w} (quality 90, 4:2:0, {nbytes // n} B each) -> RGB -> bicubic {out}x{out} + Normalize",
            "decode_frames_per_sec": n / (min(t_dec) / 1e3), "decode_ms": min(t_dec),
            "decode_resize_frames_per_sec": n / (min(t_all) / 1e3), "decode_resize_ms": min(t_all),
            "decode_frames_per_sec_1024_per_call": 4 * n / (min(t_big) / 1e3), "decode_ms_1024_per_call": min(t_big),
            "h2d_bytes": nbytes, "api": "ingest.sscd_transform(224, 224)(list of JPEG bytes)  [host bytes in, CUDA tensor out]",
            "cpu_baseline": {"value": cpu, "unit": "frames/sec", "cores": 1, "kind": "reference",
                             "sample": f"{sample} frames through Pillow Image.open + convert('RGB') + resize(BICUBIC)"},
            "parity": {"decoded_bit_identical_to_pillow": same}}


def h2d_bandwidth_gbs():
    """Pinned host -> device copy rate of this box (the ceiling of every e2e number that ships fp32 frames)."""
    import torch
    host = torch.empty(1 << 28, dtype=torch.uint8, pin_memory=True)
    dev = torch.empty(1 << 28, dtype=torch.uint8, device="cuda")
    dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        dev.copy_(host, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return 4 * (1 << 28) / (e0.elapsed_time(e1) / 1e3) / 1e9


def cpu_baseline_encoder(frames_cpu_fn, seconds_budget=20.0):
    """The oracle port of the reference ViT forward (torch fp32, all host cores) on a bounded sample."""
    import torch

    from oracle import vit_ref
    torch.set_num_threads(os.cpu_count() or 1)
    spec = vit_ref.CLIP_B16_224
    w = vit_ref.init_weights(spec, seed=0)
    x = frames_cpu_fn(8)
    vit_ref.forward(spec, w, x)          # warm-up batch
    n, t0 = 0, time.perf_counter()
    while n < 64 or (time.perf_counter() - t0 < seconds_budget and n < 512):
        vit_ref.forward(spec, w, x)
        n += 8
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "frame-descriptors/sec", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n} frames of the same 224x224 workload in batches of 8, fp32, oracle/vit_ref.py "
                      f"(restated CLIPModel.forward + gem tail), {dt:.1f} s"}


# ------------------------------------------------------------------------------------------------ similarity
def make_sim_data(device, world=1, rank=0):
    """Queries (replicated) and THIS RANK's shard of the reference / noise banks.  Weak scaling: every rank holds
    SIM_NR + SIM_NZ bank rows, the global banks have world x that many (ids offset by rank * SIM_NR)."""
    import torch
    def unit(n, seed):
        g = torch.Generator(device=device).manual_seed(seed)
        x = torch.randn((n, SIM_D), generator=g, device=device)
        return x / x.norm(dim=1, keepdim=True)
    return unit(SIM_NQ, 2), unit(SIM_NR, 3 + 100 * rank), unit(SIM_NZ, 4 + 100 * rank)


def sim_step_device(Q, R_shard, Z_shard, world, rank, row0_r):
    """score_normalize(beta=1.2, nk=1) + top-10 with Q, R, Z resident in HBM.  With world > 1 the noise bank and the
    reference bank are sharded by rows and the step has TWO collectives: one all-reduce of the noise bank's column
    moments and ONE all-gather in which the partial top-nk of the noise search and the partial top-k of the reference
    search travel together.  That is possible because the score-normalisation bias of a query is a constant added to all
    of its scores (score_normalization.py:96-101: last feature column = bias on the query side, 1 on the reference
    side): it does not change the ranking of the query's references, so the reference search runs on the un-biased
    query features and the merge kernel adds the bias afterwards."""
    import torch

    from vsc22_submission_b200 import search, sharding
    if world == 1:
        return search.score_normalized_search(Q, R_shard, Z_shard, SIM_K, beta=1.2, nk=1)
    lvd = _global_low_var_dim(Z_shard, world)
    q_0 = search.sn_transform(Q, lvd, True, fill=0.0)
    zi = search.DeviceIndex(SIM_D, search.METRIC_INNER_PRODUCT)
    zi.set_id_offset(rank * Z_shard.shape[0])          # global ids: the merged keys of a row stay distinct
    zi.add_sn(Z_shard, lvd, True, fill=0.0)
    Dz, Iz = zi.search(q_0, 1)
    ri = search.DeviceIndex(SIM_D, search.METRIC_INNER_PRODUCT)
    ri.set_id_offset(row0_r)
    ri.add_sn(R_shard, lvd, True, fill=1.0)
    D0, I0 = ri.search(q_0, SIM_K)                      # last query column 0: q.r without the bias
    keys = sharding.gather_partial_topk_multi([(Dz, Iz), (D0, I0)])
    Dz_g, _ = search.merge_packed_topk_cols(keys, 0, 1, 1)
    bias = search.bias_from_topk(Dz_g, 1.2, 1)
    return search.merge_packed_topk_cols(keys, 1, SIM_K, SIM_K, bias=bias)


def _merge_topk(D, I, k):
    """ONE all-gather of the [nq, k] partial results per rank (score and id packed into a 64-bit key), then the k-way merge
    kernel (csrc/merge.cu)."""
    from vsc22_submission_b200 import sharding
    return sharding.merge_partial_topk(D, I, k)


def _global_low_var_dim(Z_shard, world):
    """Column statistics of the row-sharded noise bank: device kernels + two all-reduces of a float64 [d] vector; the
    dropped dimension stays on the device (no host synchronisation inside the step)."""
    from vsc22_submission_b200 import sharding
    return sharding.global_low_var_dim(Z_shard, n_total=Z_shard.shape[0] * world)


def bench_sim(args, world, rank, peaks):
    import numpy as np
    import torch

    from vsc22_submission_b200 import _lib
    dev = torch.device("cuda", torch.cuda.current_device())
    Q, R, Z = make_sim_data(dev, world, rank)
    R_s, Z_s, row0 = R, Z, rank * SIM_NR
    # Everything the timed loop touches exists BEFORE the warm-up, and the warm-up runs the timed loop's exact body
    # (flush, events, profiler on), so the K timed steps find every allocator and event pool in steady state.  A 4 ms step
    # is sensitive to one-off host stalls: right after the encoder phase tears down (6 GB of pinned and device buffers
    # released) single cudaMalloc calls were seen to take 40-120 ms, and one such call inside a timed step tripled the
    # average of K = 3 steps.
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > L2: flushed between steps
    for _ in range(args.warmup + 2):
        flush.zero_()
        barrier_sync(world)
        e0.record()
        D, I = sim_step_device(Q, R_s, Z_s, world, rank, row0)
        e1.record()
        torch.cuda.synchronize()
    barrier_sync(world)
    # host-side cost of the two primitives a step is made of (diagnostic: the step is GPU-bound only while the host
    # enqueues faster than the device executes)
    from vsc22_submission_b200 import search as _s
    tiny = torch.zeros((4, 8), device=dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        torch.empty((1000, 512), dtype=torch.float32, device=dev)
    t1 = time.perf_counter()
    for _ in range(200):
        _s.bias_from_topk(tiny, 1.0, 1)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    host_probe = {"torch_empty_us": (t1 - t0) / 200 * 1e6, "tiny_launch_us": (t2 - t1) / 200 * 1e6}
    _lib.prof_collect()
    n0 = _lib.launch_count()
    total, host_issue, step_ms = 0.0, 0.0, []
    for _ in range(args.steps):
        flush.zero_()
        barrier_sync(world)
        e0.record()
        th = time.perf_counter()
        D, I = sim_step_device(Q, R_s, Z_s, world, rank, row0)
        host_issue += time.perf_counter() - th
        e1.record()
        torch.cuda.synchronize()
        total += e0.elapsed_time(e1)
        step_ms.append(round(e0.elapsed_time(e1), 3))
    barrier_sync(world)
    ms = max_over_ranks(total, world) / args.steps
    launches = _lib.launch_count() - n0
    # the per-kernel split (roofline): the same K steps again with the library's CUDA-event pair around every launch --
    # kept out of the timed steps above, where the event records would sit between the launches
    _lib.prof_collect()
    _lib.prof_enable(True)
    for _ in range(args.steps):
        flush.zero_()
        barrier_sync(world)
        sim_step_device(Q, R_s, Z_s, world, rank, row0)
        torch.cuda.synchronize()
    _lib.prof_enable(False)
    prof = _lib.prof_collect()
    barrier_sync(world)
    pairs = SIM_NQ * (SIM_NR + SIM_NZ) * world
    value = pairs / (ms / 1e3)
    sim_parity = sim_parity_check(Q, R_s, Z_s, D, I, world, rank) if world > 1 else None
    # e2e through the reference-facing API: lists of per-video host feature arrays through score_normalize
    # (score_normalization.py:33-104), the normalised features added video by video to a faiss-style index
    # (vsc/index.py:87-94) and one search over all query rows -- host buffers in, host results out
    e2e = None
    stream = None
    dense = None
    cand = None
    loc = None
    if world == 1:
        import dataclasses

        from vsc22_submission_b200 import faiss_compat as faiss, search

        @dataclasses.dataclass
        class VF:
            video_id: str
            feature: np.ndarray

        Qh, Rh, Zh = (x.cpu().numpy() for x in (Q, R, Z))
        vids = lambda pre, x, per: [VF(f"{pre}{i}", x[i:i + per]) for i in range(0, x.shape[0], per)]
        qv, rv, zv = vids("Q", Qh, 100), vids("R", Rh, 400), vids("N", Zh, 400)

        def host_step():
            q2, r2 = search.score_normalize(qv, rv, zv, beta=1.2, nk=1)
            ri = faiss.IndexFlat(SIM_D, faiss.METRIC_INNER_PRODUCT)
            for r in r2:
                ri.add(r.feature)
            return ri.search(np.concatenate([q.feature for q in q2], axis=0), SIM_K)
        host_step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            Dh, Ih = host_step()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
        e2e = {"value": pairs / (e2e_ms / 1e3), "unit": "sim-pairs/sec", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int((SIM_NQ * 2 + SIM_NR * 2 + SIM_NZ) * SIM_D * 4),
               "d2h_bytes_per_step": int((SIM_NQ + SIM_NR) * SIM_D * 4 + SIM_NQ * SIM_K * 12),
               "api": "search.score_normalize(lists of per-video host arrays) + faiss_compat.IndexFlat.add per video + "
                      ".search on numpy arrays",
               "idx_agree_with_device_path": float((Ih == I.cpu().numpy()).mean())}
        stream = bench_sim_stream(peaks)
        dense = bench_sim_dense(peaks)
        cand = bench_candidates(peaks)
        loc = bench_localization()
    flops = 2.0 * SIM_D * pairs / world            # per GPU
    bytes_alg = (SIM_NQ + SIM_NR + SIM_NZ) * SIM_D * 4 + SIM_NQ * SIM_K * 12
    sc = prof["scores"]
    roof = {"bound": "tensor", "kernel": "similarity scores kernel (both banks)",
            "achieved": (sc["work"] / (sc["ms"] / 1e3) / 1e12) if sc["ms"] > 0 else 0.0, "peak": peaks["bf16_tflops"],
            "unit": "TFLOP/s", "peak_source": f"{peaks['source']} bf16_tflops (burst)", "traffic": load_traffic("sim1_topk_kernel"),
            "whole_step_tensor_frac": flops / (ms / 1e3) / 1e12 / peaks["bf16_tflops"],
            "whole_step_hbm_frac": bytes_alg / (ms / 1e3) / 1e9 / peaks["hbm_gbs"],
            "kernel_ms": {k: round(v["ms"] / args.steps, 3) for k, v in prof.items() if v["launches"]}}
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof["note"] = ("selection on ONE bf16 MMA per product (sim1_topk_kernel, CTA pairs) with a proven error margin; the "
                    "survivors are rescored in exact fp32, so reported scores and indices are those of an fp32 brute force; "
                    "config 3 as stated is tensor-bound, the HBM-bound form is 'stream'")
    roof["kernel"] = "sim1_topk_kernel (single bf16 pass, both banks)"
    return {"metric": "sim-pairs/sec", "value": value, "unit": "sim-pairs/sec", "ms_per_step": ms,
            "host_issue_ms_per_step": host_issue * 1e3 / args.steps, "host_probe": host_probe, "step_ms": step_ms, "e2e": e2e,
            "stream": stream, "dense": dense, "candidates": cand, "localization": loc,
            "gpu_launches": int(launches), "roofline": roof, "dtype": "f32", "parity": sim_parity,
            "config": {"workload": "configs[2]: 10k query x 40k ref 512-D cosine sim + score-norm (40k noise bank, "
                                   "beta=1.2, nk=1) + top-10", "nq": SIM_NQ, "nr": SIM_NR, "nz": SIM_NZ, "d": SIM_D,
                       "k": SIM_K, "l2_flush": "256 MiB write between steps",
                       "scaling": "weak (every rank holds a 40k + 40k row bank shard; global banks = world x that)",
                       "sharding": "bank rows over ranks + all-gather of partial top-k" if world > 1 else "single GPU"}}


def sim_parity_check(Q, R_s, Z_s, D, I, world, rank, sample=64):
    """N > 1: the merged top-k of a 64-query sample against a brute force over the GATHERED banks on rank 0 (fp64 scores
    of the score-normalised descriptors; the all-gather here is test plumbing, outside every timed region)."""
    import torch
    import torch.distributed as dist
    from oracle import score_norm_np
    R_all = [torch.empty_like(R_s) for _ in range(world)]
    Z_all = [torch.empty_like(Z_s) for _ in range(world)]
    dist.all_gather(R_all, R_s)
    dist.all_gather(Z_all, Z_s)
    if rank != 0:
        return None
    Rg, Zg = torch.cat(R_all).cpu().numpy(), torch.cat(Z_all).cpu().numpy()
    q = Q[:sample].cpu().numpy()
    q2, r2, _ = score_norm_np.score_normalize(q, Rg, Zg, beta=1.2, nk=1)
    S = q2.astype("float64") @ r2.astype("float64").T
    order = (-S).argsort(axis=1, kind="stable")[:, :SIM_K]
    ref_d = S[[[i] for i in range(sample)], order]
    got_i, got_d = I[:sample].cpu().numpy(), D[:sample].cpu().numpy()
    return {"queries_checked": sample, "bank_rows": int(Rg.shape[0]), "noise_rows": int(Zg.shape[0]),
            "topk_index_agreement": float((got_i == order).mean()),
            "max_abs_score_err": float(abs(got_d - ref_d).max()), "checker": "oracle/score_norm_np.py + fp64 brute force"}


def bench_sim_stream(peaks, nq=40, nr=1_000_000, k=10, iters=10):
    """The reference's real call pattern (one index.search per query video, score_normalization.py:93-98;
    SURVEY.md 8d(i)): a few query rows against a large resident bank.  HBM-bound: algorithmic bytes =
    nr * d * 4 (the bank is read once), roofline = measured copy bandwidth."""
    import torch

    from vsc22_submission_b200 import _lib, search
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=dev).manual_seed(5)
    R = torch.nn.functional.normalize(torch.randn((nr, SIM_D), generator=g, device=dev))
    Qs = torch.nn.functional.normalize(torch.randn((nq, SIM_D), generator=g, device=dev))
    ix = search.DeviceIndex(SIM_D)
    ix.add(R)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        ix.search(Qs, k)
    call_ms, kern_ms = [], []
    for _ in range(iters):                       # whole call: library profiler off (its event records sit between the launches)
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        D, I = ix.search(Qs, k)
        e1.record()
        torch.cuda.synchronize()
        call_ms.append(e0.elapsed_time(e1))
    for _ in range(iters):                       # the dominant kernel alone: CUDA events around its launch (ProfScope)
        flush.zero_()
        torch.cuda.synchronize()
        _lib.prof_collect()
        _lib.prof_enable(True)
        D, I = ix.search(Qs, k)
        torch.cuda.synchronize()
        _lib.prof_enable(False)
        kern_ms.append(_lib.prof_collect()["scores"]["ms"])
    ref = (Qs @ R.T).topk(k, dim=1)
    exact = float((ref.indices == I).float().mean().item())
    call, kern = sum(call_ms) / iters, sum(kern_ms) / iters
    bytes_alg = nr * SIM_D * 4
    del ix, R
    torch.cuda.empty_cache()
    return {"workload": f"streaming form: {nq} query rows x {nr} bank rows x {SIM_D}-D, top-{k}, bank resident in HBM, "
                        "256 MiB L2 flush between calls", "ms_per_call": call, "pairs_per_sec": nq * nr / (call / 1e3),
            "topk_equal_torch_fp32": exact,
            "roofline": {"bound": "hbm", "kernel": "sim_stream_kernel (bank streamed once; group maxima out)",
                         "achieved": bytes_alg / (kern / 1e3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": bytes_alg / (kern / 1e3) / 1e9 / peaks["hbm_gbs"],
                         "peak_source": f"{peaks['source']} hbm_gbs (copy bandwidth)", "traffic": load_traffic("sim_stream_kernel"),
                         "algorithmic_bytes": bytes_alg,
                         "kernel_ms": kern, "whole_call_achieved": bytes_alg / (call / 1e3) / 1e9,
                         "whole_call_frac": bytes_alg / (call / 1e3) / 1e9 / peaks["hbm_gbs"]}}


def bench_sim_dense(peaks, n=40_000, iters=3):
    """BASELINE configs[4] form: the dense n x n frame-similarity matrix (fp32, written to
    HBM: n*n*4 B) and the per-row top-5 that feeds the temporal-network alignment (vta.py:262-265)."""
    import torch

    from vsc22_submission_b200 import _lib, search
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=dev).manual_seed(6)
    X = torch.nn.functional.normalize(torch.randn((n, SIM_D), generator=g, device=dev))
    ix = search.DeviceIndex(SIM_D)
    ix.add(X)
    S = ix.scores(X[:1024])
    del S
    ts, tk = [], []
    for _ in range(iters):
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        S = ix.scores(X)
        e1.record()
        D, I = ix.search(X, 5)
        e2.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1)); tk.append(e1.elapsed_time(e2))
        del S
    ms, ms_k = min(ts), min(tk)
    self_hit = float((I[:, 0] == torch.arange(n, device=dev)).float().mean().item())
    torch.cuda.empty_cache()
    return {"workload": f"dense {n} x {n} x {SIM_D} similarity matrix (fp32 out, {n * n * 4 / 1e9:.1f} GB) + per-row top-5",
            "dense_ms": ms, "top5_ms": ms_k, "pairs_per_sec_dense": n * n / (ms / 1e3), "top1_is_self": self_hit,
            "roofline": {"bound": "hbm-write / tensor (balanced)", "hbm_write_frac": n * n * 4 / (ms / 1e3) / 1e9 / peaks["hbm_gbs"],
                         "tensor_frac": 2.0 * n * n * SIM_D / (ms / 1e3) / 1e12 / peaks["bf16_tflops"],
                         "note": "3 bf16 MMAs per product: tensor ceiling 1/3"}}


def bench_candidates(peaks, nq=SIM_NQ_C, nr=SIM_NR_C, rows_q=40, rows_r=50, iters=3):
    """Video-level candidate generation (sscd_baseline.py:87-101): the global top-K frame
    pairs over ALL query rows (K = 1200 per query video, vsc/index.py:142-165) reduced to (query video, ref video)
    candidates -- device-resident, then through the reference-facing mirror with per-video host arrays."""
    import dataclasses

    import numpy as np
    import torch

    from vsc22_submission_b200 import candidates, search
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=dev).manual_seed(8)
    unit = lambda n: torch.nn.functional.normalize(torch.randn((n, SIM_D), generator=g, device=dev))
    Q, R = unit(nq), unit(nr)
    K = 1200 * (nq // rows_q)
    ix = search.DeviceIndex(SIM_D)
    ix.add(R)
    qo, ro = torch.arange(0, nq + 1, rows_q), torch.arange(0, nr + 1, rows_r)
    ix.global_search(Q, K)
    ts, tv = [], []
    for _ in range(iters):
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        s, qi, ri = ix.global_search(Q, K)
        e1.record()
        sc, qv, rv = ix.global_video_pairs(qo, ro)
        e2.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1)); tv.append(e1.elapsed_time(e2))
    ms, ms_v = min(ts), min(tv)

    @dataclasses.dataclass
    class VF:
        video_id: str
        feature: np.ndarray
        timestamps: np.ndarray = None

    Qh, Rh = Q.cpu().numpy(), R.cpu().numpy()
    qv_l = [VF(f"Q{i}", Qh[i:i + rows_q]) for i in range(0, nq, rows_q)]
    rv_l = [VF(f"R{i}", Rh[i:i + rows_r]) for i in range(0, nr, rows_r)]

    def host_step():
        cg = candidates.CandidateGeneration(rv_l, candidates.MaxScoreAggregation())
        return cg.query(qv_l, global_k=K, limit=25 * len(qv_l))        # sscd_baseline.py:99-100 keeps 25 per query
    host_step()
    t0 = time.perf_counter()
    cands = host_step()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    del ix
    torch.cuda.empty_cache()
    return {"workload": f"global top-{K} frame pairs of {nq} x {nr} x {SIM_D} (1200 per query video of {rows_q} rows) -> "
                        f"(query video, ref video) candidates by best frame pair, sorted",
            "global_search_ms": ms, "video_pairs_ms": ms_v, "pairs_per_sec": nq * nr / ((ms + ms_v) / 1e3),
            "frame_pairs_kept": int(s.numel()), "candidates": int(sc.numel()),
            "e2e": {"ms": e2e_ms, "pairs_per_sec": nq * nr / (e2e_ms / 1e3), "candidates_returned": len(cands),
                    "h2d_bytes": int((nq + nr) * SIM_D * 4), "d2h_bytes": int(sc.numel() * 20),
                    "api": "candidates.CandidateGeneration(refs, MaxScoreAggregation()).query(queries, global_k) on "
                           "per-video host arrays (index build included)"},
            "roofline": {"bound": "tensor", "tensor_frac": 2.0 * nq * nr * SIM_D / (ms / 1e3) / 1e12 / peaks["bf16_tflops"],
                         "note": "one bf16 MMA per product with the threshold emission in the GEMM epilogue (no dense "
                                 "block); radius from a 1/64 column-sample GEMM, verified by the K-th best survivor; "
                                 "survivors rescored in exact fp32 and radix-sorted"}}


def bench_localization(n_q=2000, n_r=8000, per_q=5):
    """SURVEY.md 8f row f1 (sscd_baseline.py:107-152): for 5 candidates per query video, the frame-similarity matrix
    (+0.5 bias), its per-row top-5 and the temporal-network alignment vcsl.vta.tn(tn_max_step=5, min_length=4) with MaxSim
    scoring -- the reference runs this in a 16-process pool over networkx; here one device pass for all pairs."""
    import dataclasses

    import numpy as np
    import torch

    from oracle import tn_np
    from vsc22_submission_b200.localization import VCSLLocalizationMaxSim

    @dataclasses.dataclass
    class VF:
        video_id: str
        feature: np.ndarray
        timestamps: np.ndarray

    @dataclasses.dataclass
    class Cand:
        query_id: str
        ref_id: str
        score: float = 0.0

    rng = np.random.default_rng(0)
    unit = lambda x: (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)
    mk = lambda pre, i, f: VF(f"{pre}{i}", f, np.arange(len(f), dtype=np.float32))
    refs = [mk("R", i, unit(rng.standard_normal((int(rng.integers(20, 80)), SIM_D)))) for i in range(n_r)]
    queries = []
    for i in range(n_q):
        n = int(rng.integers(10, 60))
        f = rng.standard_normal((n, SIM_D))
        src = refs[i].feature
        L = min(n, len(src), 25)
        f[:L] = src[:L] + 0.3 * rng.standard_normal((L, SIM_D)) / np.sqrt(SIM_D)      # a copied segment per query
        queries.append(mk("Q", i, unit(f)))
    cands = [Cand(f"Q{i}", f"R{(i + j * 7) % n_r}") for i in range(n_q) for j in range(per_q)]
    loc = VCSLLocalizationMaxSim(queries, refs, model_type="TN", tn_max_step=5, min_length=4, similarity_bias=0.5)
    loc.align(cands[:64])
    ts, te = [], []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        boxes, nb, _ = loc.align(cands)
        t1 = time.perf_counter()
        matches = loc.localize_all(cands)
        t2 = time.perf_counter()
        ts.append(t1 - t0); te.append(t2 - t1)
    sims = loc.similarities(cands[:300])
    t0 = time.perf_counter()
    same = sum(tn_np.tn(s, tn_max_step=5, min_length=4) == boxes[i, :nb[i]].tolist() for i, (_, s) in enumerate(sims))
    cpu = 300 / (time.perf_counter() - t0)
    torch.cuda.empty_cache()
    return {"workload": f"{len(cands)} candidate pairs ({per_q} per query video; 10-60 x 20-80 frames x {SIM_D}): sims + top-5 + "
                        "temporal network + MaxSim scores",
            "align_ms": min(ts) * 1e3, "pairs_per_sec": len(cands) / min(ts), "boxes": int(nb.sum()),
            "e2e": {"ms": min(te) * 1e3, "pairs_per_sec": len(cands) / min(te), "matches": len(matches),
                    "api": "localization.VCSLLocalizationMaxSim(...).localize_all(candidates) -> Match tuples"},
            "cpu_baseline": {"value": cpu, "unit": "pairs/sec", "cores": 1, "kind": "port",
                             "sample": "oracle tn on 300 of the pairs (similarity matrices taken from the device)"},
            "parity": {"pairs_checked": 300, "boxes_identical": int(same)}}


def cpu_baseline_candidates(nq=1000, nr=SIM_NR_C, rows_q=40, rows_r=50):
    """oracle port of CandidateGeneration on a bounded sample (1k of the 10k query rows, full bank)."""
    import numpy as np

    from oracle import candidates_np, faiss_np
    faiss_np.set_accumulate("f32")
    rng = np.random.default_rng(8)
    unit = lambda n: (lambda x: x / np.linalg.norm(x, axis=1, keepdims=True))(rng.standard_normal((n, SIM_D)).astype(np.float32))
    Q, R = unit(nq), unit(nr)
    t0 = time.perf_counter()
    c = candidates_np.candidates(Q, R, [rows_q] * (nq // rows_q), [rows_r] * (nr // rows_r), 1200 * (nq // rows_q))
    dt = time.perf_counter() - t0
    return {"value": nq * nr / dt, "unit": "sim-pairs/sec", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{nq} of {SIM_NQ_C} query rows x {nr} bank rows, global_k = 1200 per video ({len(c)} candidates), {dt:.1f} s"}


def cpu_baseline_sim():
    import numpy as np

    from oracle import faiss_np, score_norm_np
    faiss_np.set_accumulate("f32")     # plain sgemm, what faiss-CPU IndexFlat executes
    rng = np.random.default_rng(2)
    nq = 1000                          # bounded sample: 1k of the 10k queries against the full banks
    unit = lambda n: (lambda x: x / np.linalg.norm(x, axis=1, keepdims=True))(rng.standard_normal((n, SIM_D)).astype(np.float32))
    Q, R, Z = unit(nq), unit(SIM_NR), unit(SIM_NZ)
    t0 = time.perf_counter()
    q2, r2, _ = score_norm_np.score_normalize(Q, R, Z, beta=1.2, nk=1)
    ix = faiss_np.IndexFlat(SIM_D, faiss_np.METRIC_INNER_PRODUCT)
    ix.add(r2)
    ix.search(q2, SIM_K)
    dt = time.perf_counter() - t0
    faiss_np.set_accumulate("f64")
    return {"value": nq * (SIM_NR + SIM_NZ) / dt, "unit": "sim-pairs/sec", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{nq} of the 10k queries against the full 40k+40k banks, numpy sgemm + argsort "
                      f"(oracle/score_norm_np.py + oracle/faiss_np.py), {dt:.1f} s"}


# ------------------------------------------------------------------------------------------------ main
# ------------------------------------------------------------------------------------------------ BASELINE configs[3]
def bench_config4(args, world, rank, peaks):
    """configs[3]: "Swin-L + ViT-L ensemble 384^2, 200k frames + 200k x 1M sim, sharded 8 x B200".  STRONG scaling: the job
    is fixed (--c4-frames query frames through BOTH encoders -- SwinV2-L/w24 @ 384 and ViT-L/16 @ 384 -- then the ensemble
    tail, then every query descriptor against a --c4-bank-row reference bank with score normalisation and top-10);
    frames and bank rows are sharded over the ranks.  Data-path collectives: ONE all-gather of the query descriptors
    (each rank needs all queries for its bank shard) and ONE all-gather of the packed partial top-k.  The frames come
    from a pool of 1024 distinct synthetic frames per rank (1.8 GB, >> L2) that is walked repeatedly -- 200k frames
    do not fit the HBM of one GPU at N = 1."""
    import numpy as np
    import torch

    from vsc22_submission_b200 import _lib, encoder, search, sharding, swin_encoder
    from vsc22_submission_b200.ensemble import B200PCA
    dev = torch.device("cuda", torch.cuda.current_device())
    n_frames, n_bank, n_noise, d_out, k = args.c4_frames, args.c4_bank_rows, 40_000, 512, 10
    f0, f1 = sharding.shard_range(n_frames, world, rank)
    b0, b1 = sharding.shard_range(n_bank, world, rank)
    z0, z1 = sharding.shard_range(n_noise, world, rank)
    sw_spec, vit_spec = swin_encoder.SWINV2_L_384, encoder.VIT_L16_384
    sw = swin_encoder.B200SwinEncoder(sw_spec, swin_encoder.random_weights(sw_spec, seed=0), max_frames=64).to(dev).eval()
    vit = encoder.B200ViTEncoder(vit_spec, encoder.random_weights(vit_spec, seed=0), max_frames=128).to(dev).eval()
    g = torch.Generator(device=dev).manual_seed(11 + rank)
    pool_n = 1024
    pool = torch.randn((pool_n, 3, 384, 384), generator=g, device=dev).clamp_(-1, 1)
    d_cat = sw_spec.out_dim + vit_spec.width
    gp = torch.Generator().manual_seed(5)
    comp = torch.linalg.qr(torch.randn((d_cat, d_out), generator=gp))[0].T.contiguous()
    pca = B200PCA((torch.randn(d_cat, generator=gp) * 0.01).numpy(), comp.numpy(), device=dev)
    unit = lambda n, seed: torch.nn.functional.normalize(
        torch.randn((n, d_out), generator=torch.Generator(device=dev).manual_seed(seed), device=dev))
    R_s, Z_s = unit(b1 - b0, 300 + rank), unit(z1 - z0, 400 + rank)
    n_mine = f1 - f0
    desc = torch.empty((n_mine, d_out), dtype=torch.float32, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def step(n_local):
        ev[0].record()
        for c0 in range(0, n_local, pool_n):
            nc = min(pool_n, n_local - c0)
            a = sw(pool[:nc])
            b = vit(pool[:nc])[:, 0]
            desc[c0:c0 + nc] = pca.transform_parts([a, b])
        ev[1].record()
        # every rank needs all query descriptors for its bank shard: one all-gather (NVLink)
        Q = sharding.gather_descriptors(desc[:n_local], n_local * world if n_local != n_mine else n_frames) if world > 1 else desc[:n_local]
        ev[2].record()
        lvd = search.low_var_dim_device(Z_s) if world == 1 else sharding.global_low_var_dim(Z_s, n_total=n_noise)
        zi = search.DeviceIndex(d_out)
        zi.add(search.sn_transform(Z_s, lvd, True, fill=0.0))
        Dz, _ = zi.search(search.sn_transform(Q, lvd, True, fill=0.0), 1)
        if world > 1:
            Dz = sharding.merge_partial_topk(Dz, torch.zeros_like(Dz, dtype=torch.int64), 1)[0]
        q_t = search.sn_transform(Q, lvd, True, bias=search.bias_from_topk(Dz, 1.2, 1))
        ri = search.DeviceIndex(d_out)
        ri.set_id_offset(b0)
        ri.add(search.sn_transform(R_s, lvd, True, fill=1.0))
        D, I = ri.search(q_t, k)
        if world > 1:
            D, I = sharding.merge_partial_topk(D, I, k)
        ev[3].record()
        return D, I

    warm_n = min(n_mine, 2 * pool_n)
    for _ in range(args.warmup):                   # warm-up steps walk 2 pool passes per rank, the timed steps the whole share
        step(warm_n)
    barrier_sync(world)
    clocks = ClockSampler(torch.cuda.current_device())
    clocks.start()
    n0 = _lib.launch_count()
    tot = [0.0, 0.0, 0.0, 0.0]
    t_wall0 = time.time()
    for _ in range(args.steps):
        barrier_sync(world)
        D, I = step(n_mine)
        torch.cuda.synchronize()
        for i in range(3):
            tot[i] += ev[i].elapsed_time(ev[i + 1])
        tot[3] += ev[0].elapsed_time(ev[3])
    barrier_sync(world)
    t_wall1 = time.time()
    ck = clocks.stop(t_wall0, t_wall1)
    ms = max_over_ranks(tot[3], world) / args.steps
    enc_ms = max_over_ranks(tot[0], world) / args.steps
    launches = _lib.launch_count() - n0
    fl_frame = sw_spec.flops_per_frame() + vit_spec.flops_per_frame()
    pairs = float(n_frames) * (n_bank + n_noise)
    return {"metric": "frame-descriptors/sec", "unit": "frame-descriptors/sec", "value": n_frames / (ms / 1e3), "ms_per_step": ms,
            "scaling": "strong", "gpu_launches": int(launches), "clocks": ck,
            "breakdown_ms": {"encode_both_models_plus_pca": enc_ms, "query_all_gather": tot[1] / args.steps,
                             "score_norm_and_search": tot[2] / args.steps},
            "encode_frames_per_sec": n_frames / (enc_ms / 1e3), "sim_pairs_per_sec": pairs / (tot[2] / args.steps / 1e3),
            "roofline": {"bound": "tensor", "kernel": "whole step (both encoders + similarity)",
                         "achieved": (n_frames * fl_frame + 2.0 * d_out * pairs) / (ms / 1e3) / 1e12 / world,
                         "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s per GPU",
                         "frac": (n_frames * fl_frame + 2.0 * d_out * pairs) / (ms / 1e3) / 1e12 / world / peaks["bf16_tflops_sustained"],
                         "peak_source": f"{peaks['source']} bf16_tflops_sustained", "traffic": None},
            "config": {"workload": "configs[3]: SwinV2-L/w24@384 + ViT-L/16@384 ensemble (concat, PCA to 512), "
                                   f"{n_frames} frames + {n_frames} x {n_bank} similarity (score-norm, 40k noise rows, top-10)",
                       "frames_total": n_frames, "bank_rows_total": n_bank, "flops_per_frame": fl_frame,
                       "frame_pool": f"{pool_n} distinct 384x384 frames per rank walked repeatedly (1.8 GB >> L2)",
                       "parallelism": f"frames and bank rows sharded over {world} ranks; collectives: all-gather of the query "
                                      "descriptors, all-gather of the packed partial top-k, two all-reduces of the noise moments"}}



def run_reference(args):
    """--impl reference: the CPU port of the reference's path on the box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    g = torch.Generator().manual_seed(1)
    sample_fn = lambda n: torch.randn((n, 3, 224, 224), generator=g).clamp_(-1, 1)
    vals = []
    base = None
    for _ in range(max(1, min(args.steps, 3))):
        base = cpu_baseline_encoder(sample_fn, seconds_budget=15.0)
        vals.append(base["value"])
    value = sum(vals) / len(vals)
    base["value"] = value
    sim = cpu_baseline_sim() if args.workload in ("both", "sim") else None
    line = {"impl": "reference", "metric": "frame-descriptors/sec", "value": value, "unit": "frame-descriptors/sec",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": None,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: ViT-B/16 224x224 frame encoder (CLIPModel(224,16,768,12,12) + gem/Linear "
                                   "tail), CPU port on host cores, bounded sample", "batch": 8},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": "frame-descriptors/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if sim:
        line["sim"] = {"metric": "sim-pairs/sec", "value": sim["value"], "unit": "sim-pairs/sec", "cpu_baseline": sim}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="both", choices=["both", "encoder", "sim", "config4"])
    ap.add_argument("--c4-frames", type=int, default=200_000, help="config4: query frames of the whole job")
    ap.add_argument("--c4-bank-rows", type=int, default=1_000_000, help="config4: reference bank rows of the whole job")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    world, rank, local = dist_setup(args.gpus)
    peaks = load_peaks()
    line = {"metric": "frame-descriptors/sec", "unit": "frame-descriptors/sec", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic"}
    if args.workload == "config4":
        res = bench_config4(args, world, rank, peaks)
        if rank == 0:
            line.update(res)
            line["e2e"] = None
            print(json.dumps(line), flush=True)
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            dist.destroy_process_group()
        return
    enc_res = None
    if args.workload in ("both", "encoder"):
        enc_res = bench_encoder(args, world, rank, peaks)
        line.update(value=enc_res["value"], ms_per_step=enc_res["ms_per_step"], e2e=enc_res["e2e"],
                    gpu_launches=enc_res["gpu_launches"], clocks=enc_res["clocks"], roofline=enc_res["roofline"])
        line["config"] = {"workload": "configs[1]: ViT-B/16 224x224 bf16 frame encoder, 10k synthetic frames per GPU per step",
                          "arch": "CLIPModel(224,16,768,12,12) + gem(p=3)/Linear(768->512) tail, random init (seed 0)",
                          "frames_per_step_per_gpu": N_FRAMES, "plan_chunk_frames": BATCH, "tokens": 197,
                          "flops_per_frame": enc_res["enc"].spec.flops_per_frame(),
                          "cache": "6.0 GB of input frames per step, far larger than the 126 MB L2 (no flush needed)",
                          "parallelism": f"dp{world} (frames sharded over ranks, no data-path collective)",
                          "dev_vs_host_api_maxabs": enc_res["dev_vs_host_maxabs"]}
    if args.workload in ("both", "sim"):
        if enc_res is not None:
            enc_res["enc"], enc_res["frames"] = None, None
            torch.cuda.empty_cache()
        sim = bench_sim(args, world, rank, peaks)
        if enc_res is None:
            line.update(metric="sim-pairs/sec", unit="sim-pairs/sec", value=sim["value"], ms_per_step=sim["ms_per_step"],
                        e2e=sim["e2e"], gpu_launches=sim["gpu_launches"], roofline=sim["roofline"], dtype="f32",
                        config=sim["config"], stream=sim.get("stream"), dense=sim.get("dense"),
                        candidates=sim.get("candidates"), localization=sim.get("localization"),
                        host_issue_ms_per_step=sim.get("host_issue_ms_per_step"), parity=sim.get("parity"),
                        step_ms=sim.get("step_ms"))
        else:
            line["sim"] = sim
    if rank == 0 and world == 1 and args.workload == "both":
        line["swin"] = bench_swin(args)
        line["ingest"] = bench_ingest()
        line["ingest"]["jpeg"] = bench_jpeg()
        line["h2d_pinned_gbs"] = h2d_bandwidth_gbs()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        g = torch.Generator().manual_seed(1)
        if args.workload in ("both", "encoder"):
            line["cpu_baseline"] = cpu_baseline_encoder(lambda n: torch.randn((n, 3, 224, 224), generator=g).clamp_(-1, 1))
            # parity beside the timing: the first 32 frames of the job vs the oracle on the same weights, at the north-star
            # tolerance (1e-3 relative L2 per frame): the benched bf16 mode against the fp32 oracle AND the matched-precision
            # oracle, and the fp32-equivalent mode (split-bf16 GEMMs, fp32 attention) with its own throughput
            import dataclasses

            from oracle import vit_ref
            from vsc22_submission_b200.encoder import B200ViTEncoder, VIT_B16_224_GEM, random_weights
            w = random_weights(VIT_B16_224_GEM, seed=0)
            gd = torch.Generator(device="cuda").manual_seed(1)
            xd = torch.randn((1000, 3, 224, 224), generator=gd, device="cuda").clamp_(-1, 1)
            x = xd[:32].cpu()
            rel = lambda a, b: ((a - b).norm(dim=1) / b.norm(dim=1)).max().item()
            ref32 = vit_ref.forward(vit_ref.CLIP_B16_224, w, x)
            ref16 = vit_ref.forward(vit_ref.CLIP_B16_224, w, x, precision="bf16")
            got = enc_res["out"][:32].cpu()
            enc32 = B200ViTEncoder(dataclasses.replace(VIT_B16_224_GEM, precision="fp32"), w, max_frames=256).cuda().eval()
            got32 = enc32(xd[:32]).cpu()
            for _ in range(2):
                enc32(xd)
            torch.cuda.synchronize()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            enc32(xd)
            f1.record()
            torch.cuda.synchronize()
            fps32 = 1000 / (f0.elapsed_time(f1) / 1e3)
            line["parity"] = {"encoder_rel_l2_max_vs_fp32_oracle": rel(got, ref32),
                              "encoder_rel_l2_max_vs_matched_precision_oracle": rel(got, ref16),
                              "fp32_mode_rel_l2_max_vs_fp32_oracle": rel(got32, ref32),
                              "fp32_mode_frames_per_sec": fps32,
                              "fp32_mode_tflops_algorithmic": fps32 * VIT_B16_224_GEM.flops_per_frame() / 1e12,
                              "frames": 32, "tolerance": 1e-3,
                              "ok": bool(rel(got, ref32) <= 1e-3 and rel(got32, ref32) <= 1e-3)}
            del enc32, xd
            torch.cuda.empty_cache()
        if args.workload in ("both", "sim"):
            cb = cpu_baseline_sim()
            tgt = line["sim"] if "sim" in line else line
            if tgt.get("candidates"):
                tgt["candidates"]["cpu_baseline"] = cpu_baseline_candidates()
            if "sim" in line:
                line["sim"]["cpu_baseline"] = cb
            else:
                line["cpu_baseline"] = cb
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        if enc_res:
            line.pop("enc", None)
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
