#!/usr/bin/env python
"""Benchmark of the Disentangled-VAE training hot path (BASELINE.json metric: training mel-frames/sec, fwd+bwd).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision bf16|tf32]

One "step" = `model(x1, x2)` -> `loss_functionGVAE2` -> `LOSS.backward()` (+ bucketed NCCL gradient all-reduce when
N > 1) on one batch of synthetic 80-bin mel pairs.  N = 1 workload = BASELINE config 2: 256 pairs x 128 frames =
two [512, 80, 64] tensors per step (the network is locked to 64-frame chunks); every rank of an N-GPU run
carries that same shape (weak scaling: N = 8 is BASELINE config 3, global batch 2048).  mel-frames/s counts both pair
members: 2 * pairs * frames per step per GPU.

Printed JSON line: `value` = device-resident throughput (CUDA events, max over ranks); `e2e` = the same step through
the public trainer-facing API with pinned-host inputs copied H2D and the loss read back D2H every step; `roofline` =
the dominant tensor-core kernel timed live; `cpu_baseline` = the fp32 oracle (a port of the reference's PyTorch CPU path)
on the host cores, bounded sample.  `--impl reference` times that CPU path alone (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "disentangle-vae-for-vc_b200"))

METRIC = "train_mel_frames_per_sec_fwd_bwd"
UNIT = "mel-frames/s"
T_CHUNK = 64


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("DVAE_B200_PRECISION", "fp16"), choices=["fp16", "bf16", "tf32"])
    ap.add_argument("--pairs", type=int, default=256, help="pairs per step per GPU (BASELINE config 2: 256)")
    ap.add_argument("--frames", type=int, default=128, help="frames per segment (split into 64-frame chunks)")
    ap.add_argument("--cpu-rows", type=int, default=32, help="rows per call of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(args):
    rows = args.pairs * args.frames // T_CHUNK
    return (f"BASELINE config 2 per GPU: {args.pairs} pairs x {args.frames} frames = 2 x [{rows},80,64] "
            f"(64-frame chunks), latent 32, speaker_size 4, fwd+loss+bwd")


# ----------------------------------------------------------------------------- CPU path (oracle port of the reference)
def cpu_step_time(rows, steps, warmup):
    """Seconds per fwd+loss+bwd step of the fp32 oracle on all host cores, [rows,80,64] x 2."""
    import torch
    from oracle import dvae_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    sd = O.clone_sd(O.synth_state_dict(0), requires_grad=True)
    x1, x2, eps = O.synth_inputs(rows)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.train_step(sd, x1, x2, eps, batch_size=rows)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return sum(times) / len(times), torch.get_num_threads()


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path (oracle port), host cores only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rows = args.cpu_rows
    # bound the whole run to a few minutes: probe one step, shrink the sample if needed
    probe, cores = cpu_step_time(rows, 1, 0)
    budget = 150.0
    while rows > 4 and probe * (args.steps + args.warmup) > budget:
        rows //= 2
        probe, cores = cpu_step_time(rows, 1, 0)
    sec, cores = cpu_step_time(rows, args.steps, args.warmup)
    value = 2 * rows * T_CHUNK / sec
    sample = f"{rows} of {args.pairs * args.frames // T_CHUNK} rows per call per step (fp32, torch {cores} threads)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self._stop, self._th = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([f.strip() for f in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._th.join(timeout=6)

    def summary(self):
        sm = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


# ----------------------------------------------------------------------------- ours
def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (dvae_b200 has no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    os.environ["DVAE_B200_PRECISION"] = args.precision
    from dvae_b200 import lib, ops
    from dvae_b200.parallel import GradBuckets
    from model.disentangled_vae import ConvolutionalMulVAE
    from oracle import dvae_oracle as O   # only for the cpu_baseline leg and the deterministic synthetic weights

    R = args.pairs * args.frames // T_CHUNK
    frames_per_step = 2 * args.pairs * args.frames
    torch.manual_seed(1234 + rank)
    trainer = ConvolutionalMulVAE("VCTK", 64, 80, 32, 1e-4, 0.01, 500, False, batch_size=R, speaker_size=4, device=dev,
                                  latent_dim=32, beta=0.1, mse_cof=10, kl_cof=10, style_cof=0.1)
    model = trainer.model
    model.train()
    params = list(model.parameters())
    if world > 1:
        for p in params:                       # identical replicas: broadcast rank 0's random init
            dist.broadcast(p.data, 0)
        model._engine.buckets = GradBuckets([(n, tuple(p.shape)) for n, p in model.named_parameters()], dev)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x1 = torch.rand(R, 80, 64, device=dev, generator=g)
    x2 = torch.rand(R, 80, 64, device=dev, generator=g)
    noise_dev = [torch.randn(R, 28, device=dev, generator=g), torch.randn(R, 28, device=dev, generator=g),
                 torch.randn(R, 4, device=dev, generator=g)]

    def fwd_bwd(a, b):
        for p in params:
            p.grad = None
        out = model(a, b)
        losses = trainer.loss_functionGVAE2(a, b, *out)
        losses[0].backward()
        return losses

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    # ---- device-resident throughput: inputs + noise already in HBM
    it = [0]

    def dev_noise(shape):
        k = it[0] % 3
        it[0] += 1
        return noise_dev[k]
    model.noise_hook = dev_noise
    warm = max(args.warmup, 3)
    with ClockSampler(local) as clocks:
        lib.LAUNCHES = 0
        total_ms = timed(lambda: fwd_bwd(x1, x2), args.steps, warm)
        launches = lib.LAUNCHES * args.steps // (args.steps + warm)
    ms_per_step = total_ms / args.steps
    value = world * frames_per_step / (ms_per_step * 1e-3)

    # ---- end to end: pinned host inputs -> H2D, CPU-drawn noise (as the reference does), loss read back
    model.noise_hook = None
    x1_h, x2_h = x1.cpu().pin_memory(), x2.cpu().pin_memory()

    # the host side is the package's own input pipeline (dvae_b200.data): the copy of step k+1's batch overlaps step k, and
    # the loss of step k is read back (4 bytes, pinned) while step k+1 runs -- every step still moves its own inputs H2D
    # and its own loss D2H inside the timed region
    from dvae_b200.data import AsyncScalars, DevicePrefetcher

    def host_batches():
        while True:
            yield x1_h, x2_h, None
    feed = iter(DevicePrefetcher(host_batches(), dev, depth=2))
    readback = AsyncScalars(1, dev)
    last_loss = [None]

    def e2e_step():
        a, b, _ = next(feed)
        losses = fwd_bwd(a, b)
        got = readback.push(losses[0])
        if got is not None:
            last_loss[0] = got[0]
    e2e_ms = timed(e2e_step, args.steps, warm) / args.steps
    last_loss[0] = readback.flush()[0]
    e2e_value = world * frames_per_step / (e2e_ms * 1e-3)
    noise_bytes = (2 * R * 28 + R * 4) * 4
    h2d = x1_h.numel() * 4 * 2 + noise_bytes

    # ---- optimizer step (reported separately: the metric is fwd+bwd)
    fwd_bwd(x1, x2)
    opt_ms = timed(lambda: trainer.optimizer.step(), 5, 2) / 5

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant tensor-core kernel, timed live: the 512->512 k=5 implicit-GEMM conv (8 of 11 convs,
    # fwd + dgrad + wgrad all run this kernel family) at the step's own shape [2R, 64, 512]
    dt = {"bf16": lib.BF16, "fp16": lib.F16, "tf32": lib.TF32}[args.precision]
    ad = ops.act_dtype(dt)
    xs = [torch.randn(2 * R, 64, 512, device=dev).to(ad) for _ in range(3)]   # rotate buffers: 3 x 64 MB (bf16) > L2 with outputs
    wk = (torch.randn(512, 5, 512, device=dev) * 0.02).to(ad)
    bias = torch.zeros(512, device=dev)
    for i in range(3):
        ops.conv5_fwd(dt, xs[i], wk, bias)
    torch.cuda.synchronize(dev)
    reps = 12
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        ops.conv5_fwd(dt, xs[i % 3], wk, bias)
    e1.record()
    torch.cuda.synchronize(dev)
    conv_ms = e0.elapsed_time(e1) / reps
    conv_flops = 2.0 * (2 * R * 64) * 512 * (5 * 512)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops", 1590.0) * (0.5 if args.precision == "tf32" else 1.0)
    achieved = conv_flops / (conv_ms * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": "tc_gemm_persistent_kernel<BLOCK_N=256, cta_group::2 pairs> conv5 fwd 512->512", "achieved": achieved,
                "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this shape from the ncu --set full capture
                # (profiles/r01_ncu_gemm_v1.txt): 102.8 MB + 35.3 MB; algorithmic = 67.1 (x) + 2.6 (w) + 67.1 (y) MB
                "traffic": 128.1e6 if (args.precision != "tf32" and R == 512) else None,   # dram rd+wr, profiles/r01_ncu_targets_v3.txt
                "peak_source": ("MEASURED_PEAKS.json bf16_tflops (burst; kernel timed alone)" if peaks else "fallback 1.59 PF")
                + (" x 0.5 for kind::tf32" if args.precision == "tf32" else ""),
                "flops_per_launch": conv_flops, "ms_per_launch": conv_ms}
    step_flops = 168.56e6 * frames_per_step
    sustained = peaks.get("bf16_tflops_sustained", 1400.0) * (0.5 if args.precision == "tf32" else 1.0)
    step_frac = step_flops / (ms_per_step * 1e-3) / 1e12 / sustained

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        rows = args.cpu_rows
        sec, cores = cpu_step_time(rows, 2, 1)
        cpu_baseline = {"value": 2 * rows * T_CHUNK / sec, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{rows} of {R} rows per call, 1 warm-up + 2 timed fwd+loss+bwd steps of the fp32 oracle "
                                  f"(port of the reference's PyTorch CPU path), {cores} torch threads"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.precision, "data": "synthetic",
        "config": {"workload": workload_name(args), "rows_per_call": R, "frames_per_step_per_gpu": frames_per_step,
                   "parallelism": f"dp{world} (whole speaker groups per rank, bucketed NCCL all-reduce)" if world > 1 else "single GPU",
                   "l2": "inputs+activations per step (>2 GB) exceed the 126 MB L2; no explicit flush",
                   "weights": "random init (reference initialisers), no checkpoint"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms,
                "pipeline": "dvae_b200.data: H2D of step k+1 (pinned, side stream) overlaps step k; loss of step k read back during step k+1"},
        "gpu_launches": launches,
        "roofline": roofline,
        "step_tensor_frac": {"algorithmic_tflop_per_step": step_flops / 1e12, "achieved_tflops": step_flops / (ms_per_step * 1e-3) / 1e12,
                             "of_sustained_peak": step_frac},
        "optimizer_ms": opt_ms,
        "cpu_baseline": cpu_baseline,
        "clocks": clocks.summary(),
    }
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


# stdout carries exactly ONE line, the JSON result: libraries that chat on fd 1 (NCCL prints its version banner there at
# init) are sent to stderr for the whole run and the result is written to the saved descriptor.
_STDOUT_FD = None


def _quiet_stdout():
    global _STDOUT_FD
    if _STDOUT_FD is None:
        sys.stdout.flush()
        _STDOUT_FD = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _STDOUT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_STDOUT_FD, data)


if __name__ == "__main__":
    _quiet_stdout()
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
