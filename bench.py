#!/usr/bin/env python
"""bench.py — HealNet fusion forward throughput (samples/s) on B200, BASELINE.json's metric and config.

  python bench.py [--gpus N] [--steps K] [--warmup W]                 this repo's CUDA path
  python bench.py --impl reference [--gpus N --steps K --warmup W]    the reference's CPU forward (oracle port)
  torchrun ... bench.py --gpus N ...                                   one rank per GPU, batch sharded (weak scaling)

Workload (config.workload = "cfg1"): BASELINE.json configs[0] — README synthetic 3-modality example
(tab 1x2000, img 224x224x3, vol 12x224x224x3), latent 512x512, depth 3, out_dims 4, batch 4 per GPU, fp32 I/O.
A "step" is one forward over one batch of synthetic inputs (torch.rand, seed 0; default-init weights, seed 0).

Prints ONE JSON line (see the task contract): `value` = samples/s with inputs resident in HBM; `e2e` = the same
through the public module call with pinned HOST inputs (H2D + D2H inside the timed region); `roofline` for the
dominant kernel (the volume modality's streaming cross-attention), timed live with CUDA events on its launch
stream through the library's measurement hook; `cpu_baseline` = the oracle port timed on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (constructor kwargs, per-sample input shapes, per-GPU batch)
    "cfg1": (dict(n_modalities=3, channel_dims=[2000, 3, 3], num_spatial_axes=[1, 2, 3], out_dims=4, l_c=512,
                  l_d=512),
             [(1, 2000), (224, 224, 3), (12, 224, 224, 3)], 4),
    "cfg2": (dict(n_modalities=2, channel_dims=[2000, 1024], num_spatial_axes=[1, 1], out_dims=4, l_c=256, l_d=512),
             [(1, 2000), (4096, 1024)], 8),
    # cfg 3 of BASELINE.json: parameters and inputs in bf16 (the module computes in split fp16 / fp32 either way)
    "cfg3": (dict(n_modalities=3, channel_dims=[2000, 3, 3], num_spatial_axes=[1, 2, 3], out_dims=4, l_c=512,
                  l_d=1024, depth=8),
             [(1, 2000), (224, 224, 3), (12, 224, 224, 3)], 16),
    "cfg4": (dict(n_modalities=2, channel_dims=[2000, 768], num_spatial_axes=[1, 1], out_dims=4, l_c=512, l_d=512),
             [(1, 2000), (8192, 768)], 4),
    "cfg5": (dict(n_modalities=1, channel_dims=[512], num_spatial_axes=[1], out_dims=4, l_c=512, l_d=512),
             [(65536, 512)], 8),
    "tiny": (dict(n_modalities=3, channel_dims=[200, 3, 3], num_spatial_axes=[1, 2, 3], out_dims=4, l_c=128,
                  l_d=128),
             [(1, 200), (64, 64, 3), (4, 64, 64, 3)], 2),
}
METRIC = "HealNet forward samples/sec (3-modality, latent 512x512)"


def flops_per_sample(kwargs, shapes) -> float:
    """As-written algorithmic FLOPs of one forward for one sample: K/V projection, Q, QK^T + PV,
    output projection, cross feed-forward, and the latent self-attention + feed-forward after every modality.
    `shapes` are the per-sample input shapes (*axes, channels); hyper-parameters default as HealNet.__init__ does."""
    g = lambda k, dflt: kwargs.get(k, dflt)
    L, D, depth = g("l_c", 128), g("l_d", 128), g("depth", 3)
    I = g("x_heads", 8) * g("cross_dim_head", 64)
    lI = g("l_heads", 8) * g("latent_dim_head", 64)
    spc = 1 if g("self_per_cross_attn", 1) > 0 else 0
    feats = 2 * g("num_freq_bands", 2) + 1 if g("fourier_encode_data", True) else 0
    tot = 0.0
    for m, s in enumerate(shapes):
        n = _prod(s[:-1])
        c = kwargs["channel_dims"][m] + kwargs["num_spatial_axes"][m] * feats
        tot += 4 * n * c * I + 2 * L * D * I + 4 * L * n * I + 2 * L * I * D + 24 * L * D * D
        tot += spc * (6 * L * D * lI + 4 * L * L * lI + 2 * L * lI * D + 24 * L * D * D)
    return depth * tot + 2 * D * kwargs["out_dims"]


def load_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/), or None."""
    p = os.path.join(ROOT, "profiles", "r1_final_attn_small_kernel_summary.json")
    try:
        return float(json.load(open(p))["dram_bytes_per_launch"])
    except (OSError, KeyError, ValueError):
        return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops=d.get("bf16_tflops_sustained", 1375.1), hbm=d.get("hbm_gbs", 6545.3), src="measured",
                    sm_max_mhz=d.get("sm_max_mhz", 1965.0))
    return dict(tflops=1400.0, hbm=6650.0, src="fallback", sm_max_mhz=1965.0)


class ClockSampler:
    """Samples nvidia-smi SM clocks / throttle reasons every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_forward_timed(kwargs, shapes, budget_s: float, threads: int, repeats: int = 1):
    """Times the oracle port (reference algorithm: materialised K/V and attention matrices, torch CPU ops) on a
    BOUNDED sample: one sample whose volume / image token axes are cut to the leading `frac` of their first
    spatial axis so that one forward fits `budget_s`; the measured time is scaled to a full sample by the ratio of
    the as-written algorithmic FLOPs (the cost is linear in the token count)."""
    import torch
    from oracle import healnet_oracle as O
    torch.set_num_threads(threads)
    cfg = O.OracleConfig(**{k: v for k, v in kwargs.items() if k in O.OracleConfig.__dataclass_fields__})
    torch.manual_seed(0)
    from healnet_b200 import HealNet
    sd = {k: v.detach() for k, v in HealNet(**kwargs).state_dict().items()}
    full_flops = O.flops_per_sample(cfg, [s[:-1] for s in shapes])

    def run(sample_shapes):
        g = torch.Generator().manual_seed(0)
        xs = [torch.rand((1,) + tuple(s), generator=g) for s in sample_shapes]
        t0 = time.perf_counter()
        with torch.no_grad():
            out = O.forward(sd, cfg, xs, head_chunk=2)
        return time.perf_counter() - t0, out

    # calibrate on a thin slab of the largest modality, then pick the largest slab that fits the budget
    big = max(range(len(shapes)), key=lambda i: _prod(shapes[i][:-1]))
    ax0 = shapes[big][0]
    probe = [tuple(s) for s in shapes]
    probe[big] = (1,) + tuple(shapes[big][1:])
    t_probe, _ = run(probe)
    f_probe = O.flops_per_sample(cfg, [s[:-1] for s in probe])
    rate = f_probe / t_probe
    keep = ax0
    while keep > 1 and O.flops_per_sample(cfg, [s[:-1] for s in _cut(shapes, big, keep)]) / rate > budget_s:
        keep -= 1
    sample = _cut(shapes, big, keep)
    f_sample = O.flops_per_sample(cfg, [s[:-1] for s in sample])
    times = []
    for _ in range(repeats):
        t, out = run(sample)
        times.append(t)
    t_step = statistics.median(times)
    t_full = t_step * full_flops / f_sample
    desc = (f"1 sample, depth {cfg.depth}, modality {big} cut to {keep}/{ax0} of its first axis "
            f"({f_sample / full_flops * 100:.1f}% of a full sample's FLOPs), time scaled by the FLOP ratio; "
            f"oracle port, head_chunk=2, fp32")
    return dict(samples_per_s=1.0 / t_full, step_s=t_step, sample=desc, frac=f_sample / full_flops, times=times)


def _prod(t):
    p = 1
    for v in t:
        p *= v
    return p


def _cut(shapes, big, keep):
    out = [tuple(s) for s in shapes]
    out[big] = (keep,) + tuple(shapes[big][1:])
    return out


def run_reference_gpu_eager(args):
    """Opt-in (`--impl reference --ref-device cuda`): the reference algorithm as PyTorch eager ops on ONE GPU (the
    oracle port on CUDA tensors: materialised K/V and attention matrices, TF32 flags at torch defaults), the
    denominator of the north star's '>= 10x the reference single-GPU eager forward'. Not part of the driver's arms."""
    import torch
    from oracle import healnet_oracle as O
    from healnet_b200 import HealNet
    kwargs, shapes, per_gpu = WORKLOADS[args.workload]
    batch = args.batch or per_gpu
    dev = torch.device("cuda", 0)
    cfg = O.OracleConfig(**{k: v for k, v in kwargs.items() if k in O.OracleConfig.__dataclass_fields__})
    torch.manual_seed(0)
    sd = {k: v.detach().to(dev) for k, v in HealNet(**kwargs).state_dict().items()}
    g = torch.Generator().manual_seed(0)
    xs = [torch.rand((batch,) + tuple(s), generator=g).to(dev) for s in shapes]
    # the (b*h, L, N) attention matrix of the volume is 9.87 GB per sample in fp32: evaluate two heads at a time
    chunk = 2

    def step():
        with torch.no_grad():
            return O.forward(sd, cfg, xs, head_chunk=chunk)

    for _ in range(max(1, min(args.warmup, 2))):
        step()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = max(1, min(args.steps, 5))
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    line = dict(metric=METRIC, value=batch / (ms * 1e-3), unit="samples/s", n_gpus=1, steps=steps, warmup=args.warmup,
                ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                impl="reference", device="cuda-eager",
                config=dict(workload=args.workload, batch_per_gpu=batch, shapes=[list(s) for s in shapes],
                            head_chunk=chunk, **{k: kwargs[k] for k in ("l_c", "l_d")}),
                peak_mem_gb=torch.cuda.max_memory_allocated(dev) / 2 ** 30)
    print(json.dumps(line), flush=True)
    return 0


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if args.ref_device == "cuda":
        return run_reference_gpu_eager(args)
    kwargs, shapes, per_gpu = WORKLOADS[args.workload]
    threads = len(os.sched_getaffinity(0))
    total = max(1, args.steps + args.warmup)
    budget = max(2.0, min(30.0, 150.0 / total))
    import torch
    times, res = [], None
    for i in range(total):
        res = cpu_forward_timed(kwargs, shapes, budget, threads)
        if i >= args.warmup:
            times.append(res["step_s"] / res["frac"])
    t_full = statistics.median(times) if times else res["step_s"] / res["frac"]
    v = 1.0 / t_full
    line = dict(metric=METRIC, value=v, unit="samples/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=t_full * 1e3 * per_gpu, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic", impl="reference",
                config=dict(workload=args.workload, batch_per_gpu=per_gpu, shapes=[list(s) for s in shapes],
                            **{k: kwargs[k] for k in ("l_c", "l_d")}),
                cpu_baseline=dict(value=v, unit="samples/s", cores=threads, kind="port", sample=res["sample"]),
                e2e=dict(value=v, unit="samples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from healnet_b200 import HealNet
    from healnet_b200.distributed import gather_rows

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    kwargs, shapes, per_gpu = WORKLOADS[args.workload]
    batch = args.batch or per_gpu
    peaks = load_peaks()

    torch.manual_seed(0)
    model = HealNet(**kwargs).eval().to(dev)
    io_dtype = torch.bfloat16 if args.workload == "cfg3" else torch.float32
    if io_dtype != torch.float32:
        model = model.to(io_dtype)
    token_shard = args.shard == "tokens" and world > 1
    # batch sharding (default): every rank owns its own samples; token sharding: every rank sees the SAME samples and
    # streams 1/world of each long token axis (strong scaling of a batch too small to spread over the GPUs)
    g = torch.Generator().manual_seed(0 if token_shard else rank)
    host = [torch.rand((batch,) + tuple(s), generator=g).to(io_dtype).pin_memory() for s in shapes]
    resident = [t.to(dev) for t in host]
    global_batch = batch if token_shard else batch * world
    if token_shard:
        model.enable_token_sharding(min_tokens=8192, max_batch=batch)

    def step_resident():
        out = model(list(resident))
        return gather_rows(out, global_batch) if (world > 1 and not token_shard) else out

    # end to end: pinned host tensors in, logits read on the host, every step. Run the way a serving loop runs it:
    # step i+1 is enqueued (its H2D copies included) before step i's logits are waited for, so the GPU never idles on
    # the host; every step's result still lands in pinned host memory and is read there.
    out_shape = (batch, kwargs["out_dims"])
    host_out = [torch.empty(out_shape, dtype=io_dtype).pin_memory() for _ in range(2)]
    out_ready = [torch.cuda.Event() for _ in range(2)]

    def run_e2e(steps):
        model.keep_output_on_device = True
        checksum = 0.0
        for i in range(steps):
            out = model(list(host))                      # H2D of this step's inputs + forward, asynchronous
            host_out[i % 2].copy_(out, non_blocking=True)  # D2H of this step's logits
            out_ready[i % 2].record()
            if i >= 1:                                   # read the previous step's logits on the host
                out_ready[(i - 1) % 2].synchronize()
                checksum += float(host_out[(i - 1) % 2].float().sum())
        out_ready[(steps - 1) % 2].synchronize()
        checksum += float(host_out[(steps - 1) % 2].float().sum())
        model.keep_output_on_device = False
        return host_out[(steps - 1) % 2], checksum

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out

    for _ in range(max(args.warmup, 1)):
        step_resident()
    model.enable_kernel_timing(True)
    big = max(range(len(shapes)), key=lambda i: _prod(shapes[i][:-1]))
    with ClockSampler(local_rank) as clocks:
        ms, out = timed(step_resident, args.steps)
        kt = model.read_kernel_timing(big)
    model.enable_kernel_timing(False)
    launches = model.last_launch_count * args.steps
    run_e2e(2)
    # wall clock here on purpose: the timed region ends when the last step's logits have been read on the host
    barrier()
    t0 = time.perf_counter()
    out_h, _ = run_e2e(args.steps)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())

    if rank == 0:
        ms_per_step = ms / args.steps
        value = global_batch / (ms_per_step * 1e-3)
        k_ms = kt["ms"] / max(kt["launches"], 1)
        achieved = kt["flops"] / max(kt["launches"], 1) / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
        csum = clocks.summary()
        sm_hz = (csum["sm_mhz"] or peaks["sm_max_mhz"]) * 1e6
        exp_rate = kt["exps"] / max(kt["ms"], 1e-9) / 1e-3
        line = dict(
            metric=METRIC, value=value, unit="samples/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
            ms_per_step=ms_per_step, higher_is_better=True, scaling="strong" if token_shard else "weak",
            vs_baseline=None, dtype="f32" if io_dtype == torch.float32 else "bf16 I/O, fp16-split / fp32 arithmetic",
            data="synthetic",
            config=dict(workload=args.workload, batch_per_gpu=batch, global_batch=global_batch,
                        shapes=[list(s) for s in shapes], depth=kwargs.get("depth", 3), l_c=kwargs["l_c"],
                        l_d=kwargs["l_d"],
                        parallelism=(f"token-axis sharded x{world} (partials merged over NVLink peer memory)"
                                     if token_shard else f"batch-sharded x{world}"), operands="fp16 (split hi/lo on the latent side), fp32 accumulate",
                        l2="per-step working set (standardised context rows + inputs) exceeds the 126 MB L2"),
            clocks=dict(sm_mhz=csum["sm_mhz"], sm_max_mhz=csum["sm_max_mhz"], reasons=csum["reasons"]),
            e2e=dict(value=global_batch * args.steps / e2e_s, unit="samples/s",
                     h2d_bytes_per_step=sum(t.numel() * t.element_size() for t in host),
                     d2h_bytes_per_step=out_h.numel() * out_h.element_size(),
                     pipeline="depth 2: step i+1 is enqueued before step i's logits are read on the host"),
            gpu_launches=launches,
            roofline=dict(bound="tensor", achieved=achieved, peak=peaks["tflops"], unit="TFLOP/s",
                          frac=achieved / peaks["tflops"],
                          traffic=load_traffic() if args.workload == "cfg1" and batch == 4 else None,
                          peak_source=peaks["src"],
                          kernel="attn_small_kernel<32,3,6> (volume cross-attention, xattn_small.cu)", kernel_ms=k_ms,
                          kernel_share_of_step=kt["ms"] / ms_per_step if ms_per_step > 0 else None,
                          flops="executed (reassociated small-context form, padded tiles)",
                          exp_per_s=exp_rate, exp_frac_of_mufu=exp_rate / (148 * 16 * sm_hz)),
            algorithmic=dict(tflop_per_sample=flops_per_sample(kwargs, shapes) / 1e12,
                             tflops_as_written=value * flops_per_sample(kwargs, shapes) / 1e12),
        )
        if world == 1 and not args.no_cpu:
            threads = len(os.sched_getaffinity(0))
            cb = cpu_forward_timed(kwargs, shapes, args.cpu_budget, threads)
            line["cpu_baseline"] = dict(value=cb["samples_per_s"], unit="samples/s", cores=threads, kind="port",
                                        sample=cb["sample"])
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg1", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the workload's)")
    ap.add_argument("--shard", default="batch", choices=["batch", "tokens"],
                    help="multi-GPU partitioning: batch (weak scaling, default) or tokens (strong scaling of one batch)")
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"],
                    help="--impl reference only: cpu (default, the driver's arm) or cuda (PyTorch eager on one GPU)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU work for the cpu_baseline sample")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
