#!/usr/bin/env python
"""Benchmark of the Disentangled-VAE training hot path (BASELINE.json metric: training mel-frames/sec, fwd+bwd).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp16|bf16|tf32]
                    [--config 2|4|5] [--scaling weak|strong] [--lean]

One "step" = `model(x1, x2)` -> `loss_functionGVAE2` -> `LOSS.backward()` (+ bucketed NCCL gradient all-reduce when
N > 1) on one batch of synthetic 80-bin mel pairs.  N = 1 workload = BASELINE config 2: 256 pairs x 128 frames =
two [512, 80, 64] tensors per step (the network is locked to 64-frame chunks).  `--scaling weak` (default):
every rank of an N-GPU run carries that same shape (N = 8 is BASELINE config 3, global batch 2048); `--scaling strong`:
config 3's fixed global batch of 2048 pairs is divided over the ranks.  mel-frames/s counts both pair members:
2 * pairs * frames per step per GPU.

Printed JSON line (one, on stdout):
  value / ms_per_step   device-resident fwd+loss+bwd (CUDA events, max over ranks) at --precision (default fp16: the
                        fastest storage type that meets north_star's parity tolerance; `modes` times the others)
  e2e                   the same step through the trainer-facing API: pinned-host inputs copied H2D, noise drawn on the
                        CPU like the reference, loss read back D2H, every step
  full_train_step       fwd+loss+bwd + fused Adam + re-derivation of the tensor-core weight copies (a real training step)
  parity                this run's own check of the timed dtype against the fp32 oracle at R = 512 (N = 1 only)
  gpu_eager_baseline    the reference's op mix (the oracle's ATen path: cuDNN conv / BN / LSTM, cuBLAS) on this same GPU,
                        fp32 with TF32 on and bf16 autocast, same shape (N = 1 only)
  roofline              the dominant tensor-core kernel timed live; roofline_kernels: the other families (incl. the
                        LSTM step kernels, the least efficient); roofline_hbm: the memory-bound tail against HBM peak
  cpu_baseline          the fp32 oracle (a port of the reference's PyTorch CPU path) on the host cores, bounded sample
`--impl reference` times that CPU path alone (rank 0 only).  `--config 4` / `--config 5` time the secondary BASELINE
configurations (many-to-many conversion; AutoVC generator fwd+bwd).
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
CFG3_GLOBAL_PAIRS = 2048


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("DVAE_B200_PRECISION", "fp16"), choices=["fp16", "bf16", "tf32", "fp32"],
                    help="activation storage / MMA mode; fp32 = the strict checking mode on the CUDA cores (use with --lean)")
    ap.add_argument("--config", type=int, default=2, choices=[2, 4, 5], help="BASELINE config: 2 training step (3 at N > 1), "
                    "4 many-to-many conversion, 5 AutoVC generator fwd+bwd")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--pairs", type=int, default=256, help="pairs per step per GPU (BASELINE config 2: 256)")
    ap.add_argument("--frames", type=int, default=128, help="frames per segment (split into 64-frame chunks)")
    ap.add_argument("--cpu-rows", type=int, default=64, help="rows per call of the CPU sample (64 = BASELINE config 1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lean", action="store_true", help="only value / e2e / clocks (what N > 1 runs print)")
    return ap.parse_args()


def workload_name(args):
    rows = args.pairs * args.frames // T_CHUNK
    return (f"BASELINE config 2 per GPU: {args.pairs} pairs x {args.frames} frames = 2 x [{rows},80,64] "
            f"(64-frame chunks), latent 32, speaker_size 4, fwd+loss+bwd")


def base_config(args, world=1):
    R = args.pairs * args.frames // T_CHUNK
    return {"workload": workload_name(args), "rows_per_call": R, "frames_per_step_per_gpu": 2 * args.pairs * args.frames,
            "parallelism": f"dp{world} (whole speaker groups per rank, bucketed NCCL all-reduce)" if world > 1 else "single GPU",
            "l2": "inputs+activations per step (>2 GB) exceed the 126 MB L2; no explicit flush",
            "weights": "random init (reference initialisers), no checkpoint"}


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


# ----------------------------------------------------------------------------- CPU path (oracle port of the reference)
def cpu_step_time(rows, steps, warmup, threads=None):
    """Seconds per fwd+loss+bwd step of the fp32 oracle on the host cores, [rows,80,64] x 2."""
    import torch
    from oracle import dvae_oracle as O
    torch.set_num_threads(threads or os.cpu_count() or 1)
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
    """`--impl reference`: the reference's own CPU implementation of the path (the oracle port: the reference is pure Python
    over the same ATen / MKLDNN kernels), host cores only.  Each timed step is BASELINE config 1 (the reference's
    CPU-runnable case: 32 pairs x 128 frames = 64 rows per call) on all host cores; on top of the K timed steps, ONE full
    config-2 step (512 rows per call) and the as-shipped thread count (torch.set_num_threads(4)) are timed once."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rows = args.cpu_rows
    probe, cores = cpu_step_time(rows, 1, 0)
    budget = 150.0
    while rows > 4 and probe * (args.steps + args.warmup) > budget:   # bound the whole run to a few minutes
        rows //= 2
        probe, cores = cpu_step_time(rows, 1, 0)
    sec, cores = cpu_step_time(rows, args.steps, args.warmup)
    value = 2 * rows * T_CHUNK / sec
    R_full = args.pairs * args.frames // T_CHUNK
    extra = {}
    try:
        sec4, _ = cpu_step_time(rows, 1, 0, threads=4)
        extra["as_shipped_4_threads"] = {"value": 2 * rows * T_CHUNK / sec4, "unit": UNIT, "cores": 4, "rows_per_call": rows,
                                         "ms_per_step": sec4 * 1e3, "steps": 1}
        secf, _ = cpu_step_time(R_full, 1, 0)
        extra["full_config2_step"] = {"value": 2 * R_full * T_CHUNK / secf, "unit": UNIT, "cores": cores, "rows_per_call": R_full,
                                      "ms_per_step": secf * 1e3, "steps": 1}
    except Exception as e:   # the K timed steps above are the result; these are context
        extra["extra_error"] = f"{type(e).__name__}: {e}"
    sample = (f"{rows} of {R_full} rows per call per step = BASELINE config 1 shape (32 pairs x 128 frames), fp32, torch {cores} threads; "
              f"one full {R_full}-row step and a 4-thread step timed once (extra keys)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": base_config(args, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    line.update(extra)
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
            self._stop.wait(0.1)

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


def event_time(fn, steps, warmup, dev):
    """ms per call of fn(i): CUDA events on the current stream, `warmup` untimed calls first."""
    import torch
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / steps


# ----------------------------------------------------------------------------- per-kernel rooflines (rank 0, N = 1)
def tensor_rooflines(args, dev, R, peaks):
    """Live CUDA-event timing of one member of every tensor-core kernel family at the step's own shapes, rotating
    > L2 operand sets.  Returns (headline roofline dict, list of the others)."""
    import torch
    from dvae_b200 import lib, ops
    dt = {"bf16": lib.BF16, "fp16": lib.F16, "tf32": lib.TF32, "fp32": lib.F32}[args.precision]
    ad = ops.act_dtype(dt)
    scale = 0.5 if args.precision == "tf32" else 1.0
    peak_tf = peaks.get("bf16_tflops", 1590.0) * scale
    src = ("MEASURED_PEAKS.json bf16_tflops (burst; kernel timed alone)" if peaks else "fallback 1.59 PF") + \
          (" x 0.5 for kind::tf32" if args.precision == "tf32" else "")
    R2, T = 2 * R, 64
    M = R2 * T
    out = []

    def entry(kernel, flops, ms, launches=1, note=None):
        ach = flops / (ms * 1e-3) / 1e12
        e = {"bound": "tensor", "kernel": kernel, "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
             "peak_source": src, "flops_per_launch": flops / launches, "ms_per_launch": ms / launches}
        if note:
            e["note"] = note
        return e

    xs = [torch.randn(R2, T, 512, device=dev).to(ad) for _ in range(3)]   # 3 x 64 MB (16-bit) rotate through L2 with the outputs
    wk = (torch.randn(512, 5, 512, device=dev) * 0.02).to(ad)
    bias = torch.zeros(512, device=dev)
    conv_flops = 2.0 * M * 512 * (5 * 512)
    ms = event_time(lambda i: ops.conv5_fwd(dt, xs[i % 3], wk, bias), 12, 3, dev)
    head = entry("tc_gemm_persistent_kernel<BLOCK_N=256, cta_group::2 pairs> conv5 fwd 512->512", conv_flops, ms)
    # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this shape (ncu --set full, profiles/r01_ncu_targets_v3.txt);
    # algorithmic = 67.1 (x) + 2.6 (w) + 67.1 (y) MB
    head["traffic"] = 128.1e6 if (args.precision != "tf32" and R == 512) else None
    ms = event_time(lambda i: ops.conv5_dgrad(dt, xs[i % 3], wk), 12, 3, dev)
    out.append(entry("same kernel, conv5 dgrad 512->512 (B operand MN-major)", conv_flops, ms))
    dwk = torch.zeros(512, 5, 512, device=dev)
    ms = event_time(lambda i: ops.conv5_wgrad(dt, xs[i % 3], xs[(i + 1) % 3], dwk), 12, 3, dev)
    out.append(entry("same kernel, conv5 wgrad 512->512 (both operands MN-major, split-K, fp32 red.add epilogue)", conv_flops, ms))
    # LSTM input projection 65536 x 4096 x 1024 (dec_lstm2 layer 1)
    h_in = [torch.randn(M, 1024, device=dev).to(ad) for _ in range(2)]
    wih = (torch.randn(4096, 1024, device=dev) * 0.02).to(ad)
    b4 = torch.zeros(4096, device=dev)
    ms = event_time(lambda i: ops.linear_fwd(dt, h_in[i % 2], wih, b4), 8, 2, dev)
    out.append(entry("same kernel, LSTM input projection [65536 x 4096 x 1024], TMA-store epilogue", 2.0 * M * 4096 * 1024, ms))
    del h_in
    # the recurrences: 64 time steps of dec_lstm2 (H = 1024) and dec_lstm1 (H = 512), forward and backward
    for H in (1024, 512):
        xg = torch.randn(R2, T, 4 * H, device=dev).to(ad)
        whh = (torch.randn(1, 4 * H, H, device=dev) * (1.0 / H ** 0.5)).to(ad)
        keep = {}

        def fwd(i):
            keep["h"], keep["c"] = ops.lstm_fwd(dt, xg, whh, H, 1)
        ms = event_time(fwd, 3, 1, dev)
        n_f = lib._lib.dvae_lstm_launches_for(dt, R2, T, H, 1, 0)
        out.append(entry(f"LSTM recurrence fwd H={H}: {n_f} launch(es) for {T} steps, fused cell epilogue",
                         2.0 * R2 * 4 * H * H * (T - 1), ms, launches=n_f,
                         note="ms_per_launch / flops_per_launch are per launch; achieved is over the whole recurrence"))
        dh = torch.randn(R2, T, H, device=dev).to(ad) * 0.01
        ms = event_time(lambda i: ops.lstm_bwd(dt, dh, xg, keep["c"], whh, H, 1), 3, 1, dev)
        n_b = lib._lib.dvae_lstm_launches_for(dt, R2, T, H, 1, 1)
        out.append(entry(f"LSTM recurrence bwd H={H}: {n_b} launch(es) for {T} steps (dh_rec GEMM + cell backward)",
                         2.0 * R2 * 4 * H * H * (T - 1), ms, launches=n_b,
                         note="ms_per_launch / flops_per_launch are per launch; achieved is over the whole recurrence"))
        del xg, dh
    return head, out


def hbm_rooflines(dev, R, peaks):
    """The memory-bound tail against the measured HBM copy bandwidth: algorithmic bytes / CUDA-event time."""
    import torch
    from dvae_b200 import lib, ops
    from dvae_b200.lib import call, ptr, stream
    peak = peaks.get("hbm_gbs", 6650.0)
    src = "MEASURED_PEAKS.json hbm_gbs (copy)" if peaks else "fallback 6.65 TB/s"
    res = []

    def entry(kernel, nbytes, ms, note=None):
        gbs = nbytes / (ms * 1e-3) / 1e9
        e = {"bound": "hbm", "kernel": kernel, "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
             "bytes_per_launch": nbytes, "ms_per_launch": ms, "peak_source": src}
        if note:
            e["note"] = note
        return e
    g = torch.Generator(device=dev).manual_seed(0)
    # ---- speaker-group accumulate + finalize at north_star's standalone size
    B, D, G = 1 << 22, 32, 1 << 17
    mu = torch.randn(B, D, device=dev, generator=g)
    lv = torch.randn(B, D, device=dev, generator=g) * 0.5
    labels = torch.arange(B, device=dev, dtype=torch.int64) // (B // G)
    gid, _ = ops.segment_ids_sorted(labels)
    acc = torch.zeros(G, 2, D, device=dev)
    cnt = torch.zeros(G, device=dev)
    out_a, out_b = torch.empty_like(mu), torch.empty_like(mu)
    table = torch.empty_like(acc)

    def accumulate(i):
        acc.zero_(), cnt.zero_()
        call("dvae_group_accumulate", ops.MODE_POG, ptr(mu), ptr(lv), ptr(gid), ptr(acc), ptr(cnt), B, D, stream())
    t_zero = event_time(lambda i: (acc.zero_(), cnt.zero_()), 10, 3, dev)
    t_acc = event_time(accumulate, 10, 3, dev) - t_zero
    t_fin = event_time(lambda i: call("dvae_group_finalize", ops.MODE_POG, ptr(acc), ptr(cnt), ptr(gid), ptr(table), ptr(out_a),
                                      ptr(out_b), B, G, D, stream()), 10, 3, dev)
    nb = 2 * B * D * 4 + B * 4
    res.append(entry("group_accumulate (product of Gaussians, 2^22 rows, D = 32, 2^17 groups)", nb, t_acc))
    res.append(entry("group_finalize (per-group table + scatter back to rows)", nb, t_fin))
    res.append(entry("group accumulate + finalize", 2 * nb, t_acc + t_fin))
    del mu, lv, labels, gid, out_a, out_b
    # ---- fused loss forward at the step's shape (6 x [R,80,64] fp32 read)
    n = R * 80 * 64
    ts = [torch.rand(R, 80, 64, device=dev, generator=g) for _ in range(6)]
    q = [torch.randn(R, 32, device=dev, generator=g) * 0.1 for _ in range(4)]
    s2 = [torch.randn(R, 4, device=dev, generator=g) * 0.1 for _ in range(2)]
    ms = event_time(lambda i: ops.loss_fwd(*ts, *q, *s2, float(R), 10.0, 10.0), 20, 3, dev)
    res.append(entry("fused loss forward (4 L1 sums + 3 KL, one launch)", 6 * n * 4, ms,
                     note="63 MB per launch: fits L2 (126 MB), so this is an L2-resident figure, not an HBM one"))
    # ---- BatchNorm finalize + apply at [65536, 512]
    rows, C = 2 * R * 64, 512
    dt, eb = lib.F16, 2
    ys = [torch.randn(rows, C, device=dev).to(torch.float16) for _ in range(3)]
    gam, bet = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    rm, rv, nbt = torch.zeros(C, device=dev), torch.ones(C, device=dev), torch.zeros((), device=dev, dtype=torch.long)
    st = {}

    def bnf(i):
        st["o"], st["s"] = ops.bn_train_fwd(dt, ys[i % 3], gam, bet, rm, rv, nbt, 2, lib.ACT_RELU, 1e-5, 0.1)
    ms = event_time(bnf, 9, 3, dev)
    res.append(entry("BatchNorm train forward, 16-bit [65536, 512]: statistics + finalize + apply (3 launches)", rows * C * eb * 3, ms))
    ms = event_time(lambda i: ops.bn_train_bwd(dt, ys[(i + 1) % 3], ys[i % 3], st["s"], 2, lib.ACT_RELU), 9, 3, dev)
    res.append(entry("BatchNorm train backward, 16-bit [65536, 512]: reduce + finalize + apply (3 launches)", rows * C * eb * 5, ms))
    return res


# ----------------------------------------------------------------------------- GPU-eager reference (rank 0, N = 1)
def gpu_eager_baseline(dev, R, steps=10, warmup=3):
    """The reference's own op mix on this GPU: the oracle's ATen path (F.conv1d / batch_norm / torch.lstm / F.linear ->
    cuDNN + cuBLAS, autograd backward), same shape, same synthetic data, CUDA events.  Two precisions: fp32 storage with
    TF32 tensor cores allowed, and bf16 autocast."""
    import torch
    from oracle import dvae_oracle as O
    sd = O.clone_sd(O.synth_state_dict(0), requires_grad=True, device=dev)
    x1, x2, eps = O.synth_inputs(R)
    x1, x2, eps = x1.to(dev), x2.to(dev), [e.to(dev) for e in eps]
    flags = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    res = {"impl": "oracle's ATen path (cuDNN conv / BN / LSTM, cuBLAS linear, autograd), torch " + torch.__version__,
           "rows_per_call": R, "steps": steps, "warmup": warmup}
    frames = 2 * R * T_CHUNK
    try:
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True
        ms = event_time(lambda i: O.train_step(sd, x1, x2, eps, batch_size=R), steps, warmup, dev)
        res["fp32_tf32"] = {"ms_per_step": ms, "value": frames / (ms * 1e-3), "unit": UNIT}

        def amp(i):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                O.train_step(sd, x1, x2, eps, batch_size=R)
        ms = event_time(amp, steps, warmup, dev)
        res["bf16_autocast"] = {"ms_per_step": ms, "value": frames / (ms * 1e-3), "unit": UNIT}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = flags
    return res


# ----------------------------------------------------------------------------- ours: config 2 / 3 (training step)
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
    from dvae_b200 import lib
    from dvae_b200.parallel import GradBuckets
    from model.disentangled_vae import ConvolutionalMulVAE

    if args.scaling == "strong":
        assert CFG3_GLOBAL_PAIRS % world == 0
        args.pairs = CFG3_GLOBAL_PAIRS // world      # BASELINE config 3: the global batch is fixed, the ranks share it
    R = args.pairs * args.frames // T_CHUNK
    frames_per_step = 2 * args.pairs * args.frames
    lean = args.lean or world > 1
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
        for _ in range(warm):
            fwd_bwd(x1, x2)
        lib.LAUNCHES = 0
        total_ms = timed(lambda: fwd_bwd(x1, x2), args.steps, 0)
        launches = lib.LAUNCHES
    ms_per_step = total_ms / args.steps
    value = world * frames_per_step / (ms_per_step * 1e-3)

    # ---- host enqueue time of one step (queue empty at the start, nothing waits on the GPU)
    enq = []
    for _ in range(3):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        fwd_bwd(x1, x2)
        enq.append((time.perf_counter() - t0) * 1e3)
    torch.cuda.synchronize(dev)
    host_enqueue_ms = min(enq)

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

    # ---- a real training step: fwd + loss + bwd + fused Adam + re-derivation of the tensor-core weight copies (the next
    # forward refreshes them because the optimizer moved the parameters).  lr = 0: the timing is the same, the weights stay.
    model.noise_hook = dev_noise
    for grp in trainer.optimizer.param_groups:
        grp["lr"] = 0.0

    def train_step():
        fwd_bwd(x1, x2)
        trainer.optimizer.step()
    full_ms = timed(train_step, args.steps, 2) / args.steps
    fwd_bwd(x1, x2)
    opt_ms = timed(lambda: trainer.optimizer.step(), 5, 2) / 5

    # ---- the same step replayed as ONE CUDA graph (dvae_b200.graph.GraphedTrainStep: weight refresh + forward + loss + backward;
    # single GPU: the bucketed all-reduce is not captured).  Device time and the host time of a replay.
    graph_info = None
    if world == 1 and not args.lean:
        try:
            from dvae_b200.graph import GraphedTrainStep
            model.noise_hook = dev_noise
            gstep = GraphedTrainStep(trainer, x1, x2)
            g_ms = timed(lambda: gstep(x1, x2), args.steps, 3) / args.steps
            g_enq = []
            for _ in range(3):
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                gstep(x1, x2)
                g_enq.append((time.perf_counter() - t0) * 1e3)
            torch.cuda.synchronize(dev)
            graph_info = {"ms_per_step": g_ms, "value": frames_per_step / (g_ms * 1e-3), "unit": UNIT,
                          "host_ms_per_step": min(g_enq),
                          "includes": "refresh of the tensor-core weight copies + fwd + loss + bwd as one graph launch; inputs and noise "
                                      "copied into the graph's static tensors every step (device to device here)"}
            del gstep
        except Exception as e:   # reported, never fatal: the eager numbers above are the bench
            graph_info = {"error": f"{type(e).__name__}: {e}"[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    step_flops = 168.56e6 * frames_per_step
    sustained = peaks.get("bf16_tflops_sustained", 1400.0) * (0.5 if args.precision == "tf32" else 1.0)
    cfg = base_config(args, world)
    if args.scaling == "strong":
        cfg["workload"] = (f"BASELINE config 3 (strong scaling): {CFG3_GLOBAL_PAIRS} pairs x {args.frames} frames over {world} GPU(s) = "
                           f"2 x [{R},80,64] per rank, fwd+loss+bwd + gradient all-reduce")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": args.precision, "data": "synthetic", "config": cfg,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms,
                "pipeline": "dvae_b200.data: H2D of step k+1 (pinned, side stream) overlaps step k; loss of step k read back during step k+1"},
        "gpu_launches": launches,
        "gpu_launches_per_step": launches / args.steps,
        "host_enqueue_ms_per_step": host_enqueue_ms,
        "cuda_graph_step": graph_info,
        "full_train_step": {"ms_per_step": full_ms, "value": world * frames_per_step / (full_ms * 1e-3), "unit": UNIT,
                            "includes": "fwd + loss + bwd + fused Adam (one launch) + refresh of the tensor-core weight copies"},
        "optimizer_ms": opt_ms,
        "step_tensor_frac": {"algorithmic_tflop_per_step": step_flops / 1e12, "achieved_tflops": step_flops / (ms_per_step * 1e-3) / 1e12,
                             "of_sustained_peak": step_flops / (ms_per_step * 1e-3) / 1e12 / sustained},
        "clocks": clocks.summary(),
    }
    if not lean:
        del trainer, model, params
        torch.cuda.empty_cache()
        head, others = tensor_rooflines(args, dev, R, peaks)
        line["roofline"] = head
        line["roofline_kernels"] = others
        line["roofline_hbm"] = hbm_rooflines(dev, R, peaks)
        torch.cuda.empty_cache()
        from oracle.parity import parity_report       # the oracle as the checker of the timed dtype, in this same run
        line["parity"] = parity_report(args.precision, R)
        torch.cuda.empty_cache()
        line["gpu_eager_baseline"] = gpu_eager_baseline(dev, R)
        best = min(v["ms_per_step"] for k, v in line["gpu_eager_baseline"].items() if isinstance(v, dict))
        line["gpu_eager_baseline"]["speedup_of_value_over_fastest_eager"] = best / ms_per_step
        torch.cuda.empty_cache()
        modes = {}
        for other in ("fp16", "bf16", "tf32"):
            if other == args.precision:
                continue
            os.environ["DVAE_B200_PRECISION"] = other
            tr = ConvolutionalMulVAE("VCTK", 64, 80, 32, 1e-4, 0.01, 500, False, batch_size=R, speaker_size=4, device=dev,
                                     latent_dim=32, beta=0.1, mse_cof=10, kl_cof=10, style_cof=0.1)
            tr.model.train()
            tr.model.noise_hook = dev_noise
            ps = list(tr.model.parameters())

            def step_other(i):
                for p in ps:
                    p.grad = None
                o = tr.model(x1, x2)
                tr.loss_functionGVAE2(x1, x2, *o)[0].backward()
            ms = event_time(step_other, 5, 3, dev)
            modes[other] = {"ms_per_step": ms, "value": frames_per_step / (ms * 1e-3), "unit": UNIT}
            del tr, ps
            torch.cuda.empty_cache()
        os.environ["DVAE_B200_PRECISION"] = args.precision
        line["modes"] = modes
        if not args.no_cpu_baseline:
            rows = args.cpu_rows
            sec, cores = cpu_step_time(rows, 3, 1)
            sec4, _ = cpu_step_time(rows, 1, 0, threads=4)
            line["cpu_baseline"] = {
                "value": 2 * rows * T_CHUNK / sec, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{rows} of {R} rows per call (= BASELINE config 1 shape), 1 warm-up + 3 timed fwd+loss+bwd steps of the fp32 "
                          f"oracle (port of the reference's PyTorch CPU path), {cores} torch threads",
                "as_shipped_4_threads": {"value": 2 * rows * T_CHUNK / sec4, "unit": UNIT, "cores": 4, "steps": 1}}
        else:
            line["cpu_baseline"] = None
    else:
        line["roofline"] = None
        line["cpu_baseline"] = None
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------- ours: secondary configs
def run_secondary(args):
    """BASELINE config 4 (many-to-many conversion inference, 512 utterances x 512 frames) or config 5 (AutoVC-style
    generator fwd+bwd, batch 256 x 128 frames) on one GPU."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (dvae_b200 has no CPU path)")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    os.environ["DVAE_B200_PRECISION"] = args.precision
    from dvae_b200 import lib
    torch.manual_seed(0)
    warm = max(args.warmup, 3)
    if args.config == 5:
        from autovc_replicate.proposed_autovc import Generator
        R = args.pairs * args.frames // T_CHUNK
        gen = Generator().to(dev)
        gen.train()
        x = torch.rand(R, 80, 64, device=dev)
        x_h = x.cpu().pin_memory()
        n_el = float(x.numel())

        def step(inp):
            for p in gen.parameters():
                p.grad = None
            mel, post = gen(inp)
            t = inp.transpose(1, 2).unsqueeze(1)
            loss = ((mel - t).abs().sum() + (post - t).abs().sum()) / n_el      # l1(mel, x^T) + l1(mel_postnet, x^T), SURVEY 8d
            loss.backward()
            return loss
        with ClockSampler(0) as clocks:
            for _ in range(warm):
                step(x)
            lib.LAUNCHES = 0
            ms = event_time(lambda i: step(x), args.steps, 0, dev)
            launches = lib.LAUNCHES

        def e2e(i):
            step(x_h.to(dev, non_blocking=True)).item()
        e2e_ms = event_time(e2e, args.steps, warm, dev)
        frames = R * T_CHUNK
        line = {"metric": "autovc_generator_train_mel_frames_per_sec_fwd_bwd", "value": frames / (ms * 1e-3), "unit": UNIT,
                "config": {"workload": f"BASELINE config 5: autovc_replicate.proposed_autovc.Generator fwd+bwd, [{R},80,64] "
                                       f"(batch {args.pairs} x {args.frames} frames), L1 loss on both outputs"},
                "e2e": {"value": frames / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": x_h.numel() * 4, "d2h_bytes_per_step": 4,
                        "ms_per_step": e2e_ms}}
    else:
        from model.disentangled_vae import ConvolutionalMulVAE
        w = ConvolutionalMulVAE("VCTK", 64, 80, 32, 1e-4, 0.01, 500, False, batch_size=8, speaker_size=4, device=dev, latent_dim=32)
        w.model.eval()
        U, frames_u = 512, 512
        src = torch.rand(U, 80, frames_u, device=dev)
        trg = torch.rand(U, 80, frames_u, device=dev)
        src_h, trg_h = src.cpu().pin_memory(), trg.cpu().pin_memory()

        def step(a, b):
            return w.convert_utterances(a, b)
        with ClockSampler(0) as clocks:
            for _ in range(warm):
                step(src, trg)
            lib.LAUNCHES = 0
            ms = event_time(lambda i: step(src, trg), args.steps, 0, dev)
            launches = lib.LAUNCHES
        out_h = torch.empty((U, 80, (frames_u // 64 + 1) * 64), pin_memory=True)

        def e2e(i):
            _, cv = step(src_h.to(dev, non_blocking=True), trg_h.to(dev, non_blocking=True))
            out_h.copy_(cv, non_blocking=True)
        e2e_ms = event_time(e2e, args.steps, warm, dev)
        torch.cuda.synchronize(dev)
        line = {"metric": "conversion_utterances_per_sec", "value": U / (ms * 1e-3), "unit": "utterances/s",
                "mel_frames_per_s": U * frames_u / (ms * 1e-3),
                "config": {"workload": f"BASELINE config 4: many-to-many conversion, {U} utterances x {frames_u} frames = {U * 9} source + "
                                       f"{U * 9} target 64-frame chunks (chunking_mel incl. the zero chunk): encode both, per-utterance "
                                       "style mean, decode, postnet, time-concat, clamp"},
                "e2e": {"value": U / (e2e_ms * 1e-3), "unit": "utterances/s", "h2d_bytes_per_step": 2 * src_h.numel() * 4,
                        "d2h_bytes_per_step": out_h.numel() * 4, "ms_per_step": e2e_ms}}
    line.update({"n_gpus": 1, "steps": args.steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                 "vs_baseline": None, "dtype": args.precision, "data": "synthetic", "gpu_launches": launches,
                 "roofline": None, "cpu_baseline": None, "clocks": clocks.summary()})
    _emit(line)


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
    elif a.config == 2:
        run_ours(a)
    else:
        run_secondary(a)
