#!/usr/bin/env python
"""Benchmark of the DL_VQA training step (BASELINE.json: "train samples/sec at 1/2/4/8 B200").

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...  # the UNMODIFIED reference (oracle/_ref) on the host cores

One step = forward + soft-target loss + backward + Adam on one synthetic batch of 256 samples per GPU at the
config.yaml shapes (BASELINE.json configs[1]), dropout 0.3 active (train mode).  Prints ONE JSON line.
"""
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

# algorithmic FLOP per sample of the contraction kernels, forward; dgrad/wgrad equal
CONV_FLOP = {0: 0.170e9, 1: 1.752e9, 2: 1.595e9}
WORKLOAD = ("single-B200 training step (BASELINE.json configs[1]): full VqaNet fwd + soft-target loss "
            "+ bwd + Adam, config.yaml shapes, dropout 0.3, random init, V=15000, T=23")      # both arms name the same workload
STEP_FLOP_PER_SAMPLE = 13.00e9          # fwd + bwd at T = 23


def _ncu_traffic(tag):
    """DRAM bytes per launch of kernel `tag` from the newest committed ncu capture (profiles/r*_ncu_traffic.json)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_traffic.json")))
    if not files:
        return None, None
    d = json.load(open(files[-1]))
    k = d.get("kernels", {}).get(tag)
    if not k:
        return None, None
    return k["dram_bytes_read"] + k["dram_bytes_write"], os.path.basename(files[-1])


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the unmodified reference step (oracle/_ref; oracle port if absent), timed on the host cores
# --------------------------------------------------------------------------------------------------
def _use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm is meant to use the host's cores."""
    import torch
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def cpu_reference_steps(steps, warmup, batch=32):
    """The reference's CPU path of the step on the host cores: the UNMODIFIED reference from oracle/_ref (models/model.py
    VqaNet, train.py run_batch / update_learning_rate, torch.optim.Adam as train.py:55) when that copy is present
    (kind "reference"), else the oracle port of the same arithmetic (kind "port").  fp32, train() mode with the
    config.yaml dropouts, config.yaml shapes (BASELINE.json configs[0])."""
    import warnings
    import torch
    from oracle import vqa_oracle as O
    from oracle import ref_loader
    threads = _use_all_host_threads()
    cfg = O.DEFAULT_CFG
    V = 15000
    v, q, q_len, a_idx, a_val, a_len = O.synthetic_batch(batch, cfg, V, seed=1)
    times = []
    if ref_loader.available():
        ref = ref_loader.load()
        torch.manual_seed(1)
        model = ref.VqaNet(cfg, V).train(True)
        opt = torch.optim.Adam(model.parameters(), lr=5e-4)            # reference train.py:55
        log_softmax = torch.nn.LogSoftmax(dim=1)
        # run_batch moves its inputs with `if torch.cuda.is_available(): v = v.cuda()` (train.py:183-187); this leg times the
        # CPU path on a box that has a GPU, so CUDA is hidden from the reference for its duration (its code is untouched)
        real_is_available = torch.cuda.is_available
        torch.cuda.is_available = lambda: False
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                for it in range(warmup + steps):
                    t0 = time.perf_counter()
                    loss, _ = ref.train.run_batch(model, log_softmax, (v, q, a_idx, a_val, a_len, None, q_len), cfg["max_answers"])
                    opt.zero_grad()
                    ref.train.update_learning_rate(opt, it, 5e-4)
                    loss.backward()
                    opt.step()
                    dt = time.perf_counter() - t0
                    if it >= warmup:
                        times.append(dt)
        finally:
            torch.cuda.is_available = real_is_available
        return times, batch, threads, "reference"
    sd = O.random_params(cfg, V, seed=1)
    leaves = {k: t.clone().requires_grad_(True) for k, t in sd.items()}
    opt = torch.optim.Adam(list(leaves.values()), lr=5e-4)
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        logits = O.forward(leaves, cfg, v, q, q_len)
        loss = O.soft_target_loss_dense(logits, a_idx, a_val)
        opt.zero_grad()
        for g in opt.param_groups:
            g["lr"] = O.learning_rate(5e-4, it)
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return times, batch, threads, "port"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    # bounded sample: B=32 per step (about 3 s of CPU work each); cap the total at a few minutes
    steps_eff = min(steps, 40)
    warm = min(args.warmup, 3)
    times, batch, threads, kind = cpu_reference_steps(steps_eff, warm, batch=32)
    ms = 1000.0 * sum(times) / len(times)
    val = batch / (ms / 1000.0)
    what = ("unmodified reference (oracle/_ref: models/model.py VqaNet + train.py run_batch) + torch Adam" if kind == "reference"
            else "oracle port of models/model.py fwd + train.py loss + bwd + torch Adam")
    sample = (f"{what}, fp32, dropout 0.3, batch {batch} per step (bounded sample of the 256-sample step), "
              f"{steps_eff} timed steps, {threads} torch threads, os.cpu_count()={os.cpu_count()}")
    line = {"impl": "reference", "metric": "train samples/sec", "value": val, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": steps_eff, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            # the SAME workload as the GPU arm's line; each step is a bounded sample of it (batch 32 of the 256-sample step)
            "config": {"workload": WORKLOAD, "batch_per_gpu": 256, "global_batch": 256 * max(1, args.gpus),
                       "parallelism": f"dp{max(1, args.gpus)}",
                       "sample": f"batch {batch} per step on the host CPU (rank 0 only): a bounded sample of the 256-sample step"},
            "cpu_baseline": {"value": val, "unit": "samples/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


def _bind_to_gpu_numa_node(local):
    """Pin this rank's host threads (and therefore its pinned staging buffers, first touch) to the CPUs that are local
    to its GPU's PCIe root: with 4-8 ranks copying 154 MB per step each, remote-socket staging memory otherwise caps the
    host-to-device rate.  Best effort: silently skipped when sysfs does not expose the topology."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        dev = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        cpus = open(f"/sys/bus/pci/devices/{dev}/local_cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            if "-" in part:
                a, b = part.split("-"); ids.update(range(int(a), int(b) + 1))
            elif part:
                ids.add(int(part))
        ids &= set(os.sched_getaffinity(0))
        if ids:
            os.sched_setaffinity(0, ids)
            return f"{dev}: {cpus}"
    except Exception:
        pass
    return None


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import dl_vqa_b200 as D
    from dl_vqa_b200 import lib, synth
    from dl_vqa_b200.dp import GradientAllReduce

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = _bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib.load()
    if args.conv_cta_group:
        lib.call("vqa_tc_conv_set_cta_group", args.conv_cta_group)

    B = args.batch
    cfg = synth.default_cfg()                       # config.yaml defaults, dropout 0.3
    torch.manual_seed(1)
    model = D.VqaNet(cfg, synth.DEFAULT_TOKENS, compute_dtype=args.dtype).to(dev).train(True)
    opt = D.FusedAdam(model.parameters(), lr=5e-4)
    model.use_gradient_arena(True)        # gradients live in one persistent arena (all-reduced in place, bucket by bucket, for N > 1)
    if args.dtype == "bfloat16":
        model.use_weight_shadows(opt)     # FusedAdam keeps the bf16 GEMM weight shadows current (no per-step re-casts)
    ddp = GradientAllReduce(model)
    ddp.broadcast_parameters()
    # the loop body of train.py:69-81 as one CUDA graph per input buffer set (--no-graph: enqueue kernel by kernel)
    gstep = D.GraphedTrainStep(model, opt, cfg["max_answers"], ddp=ddp if world > 1 else None, lr=5e-4,
                               enabled=not args.no_graph)

    host = synth.make_batch(B, cfg, seed=1 + rank, pin=True)
    hv, hq, hai, hav, hal, _, hql = host
    hv16 = hv.to(torch.float16).pin_memory()          # the reference's on-disk image dtype (preprocessing/preprocess_images.py:40)
    assert torch.equal(hv16.float(), hv)
    small_bytes = sum(t.numel() * t.element_size() for t in (hq, hai, hav, hql))
    loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()

    def step(dbatch):
        return gstep(dbatch)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    hostt = {"enqueue_ms": 0.0}

    def timed(fn, n):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        for _ in range(n):
            fn()
        b.record()
        hostt["enqueue_ms"] = 1000.0 * (time.perf_counter() - t0)      # host time to enqueue the region (no sync inside)
        barrier()
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    dbatch = tuple(t.to(dev, non_blocking=True) for t in (hv, hq, hai, hav, hal)) + (None, hql.to(dev, non_blocking=True))
    for _ in range(max(3, args.warmup)):      # call 1 runs eagerly, call 2 captures + replays, the rest replay
        step(dbatch)
    torch.cuda.synchronize()

    # ---- (1) device-resident throughput (no per-kernel events inside this region)
    tags = [f"conv{i}_{k}" for i in range(3) for k in ("fwd", "dgrad", "wgrad")] + \
           ["v_conv", "v_conv_dgrad", "v_conv_wgrad", "vqa_attention_fwd", "vqa_attention_bwd", "lstm_step_fwd",
            "lstm_step_bwd", "lstm_recurrence_fwd", "lstm_bwd_pointwise", "lstm_bwd_persistent", "lstm_inproj", "lstm_whh_wgrad",
            "lstm_wih_wgrad", "lstm_inproj_dgrad", "lin1", "lin2", "q_lin", "lin1_dgrad", "lin2_dgrad", "lin1_wgrad", "lin2_wgrad",
            "q_lin_dgrad", "q_lin_wgrad", "act_cast", "vqa_adam_multi", "vqa_adam_multi_dev", "act_transpose", "unpool", "w_cast",
            "w_transpose", "embed_fwd", "embed_bwd", "vqa_dropnorm_fwd", "vqa_softloss_fwd_bwd", "vqa_colsum", "vqa_zero"]
    clocks = ClockSampler(local)
    clocks.start()
    n0, r0 = lib.launch_count(), gstep.replays
    if world > 1:
        gstep.wait_events = []          # CUDA events around the compute stream's wait for the gradient all-reduces
    if args.profile_mode:            # `ncu --profile-from-start off` captures exactly the timed steps
        torch.cuda.cudart().cudaProfilerStart()
    ms_dev = timed(lambda: step(dbatch), args.steps)
    if args.profile_mode:
        torch.cuda.cudart().cudaProfilerStop()
    wait_ms = sorted(a.elapsed_time(b) for a, b in (gstep.wait_events or []))
    gstep.wait_events = None
    host_ms_dev = hostt["enqueue_ms"] / args.steps
    # kernels of this library inside the timed region: counted at enqueue time, plus (replayed graphs) x (kernels captured per graph)
    launches = (lib.launch_count() - n0) + (gstep.replays - r0) * gstep.launches_per_replay

    if args.only_device:             # quick A/B runs (NCCL settings): device-resident number and the communication split only
        clocks.stop()
        comm = None
        if world > 1:
            model.grad_ready_hook = None
            g2 = D.GraphedTrainStep(model, opt, cfg["max_answers"], ddp=None, lr=5e-4, enabled=not args.no_graph)
            for _ in range(3):
                g2(dbatch)
            ms_nocomm = timed(lambda: g2(dbatch), args.steps)
            comm = {"finish_wait_ms_median": wait_ms[len(wait_ms) // 2] if wait_ms else None,
                    "ms_per_step_without_allreduce": ms_nocomm / args.steps}
        if rank == 0:
            _emit({"only_device": True, "n_gpus": world, "ms_per_step": ms_dev / args.steps,
                   "value": world * B / (ms_dev / args.steps / 1000.0), "exposed_comm": comm,
                   "nccl_env": {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}})
        if world > 1:
            dist.destroy_process_group()
        return

    if args.profile_mode:            # under ncu: no second timed region, no breakdown pass, no CPU leg
        clocks.stop()
        if rank == 0:
            _emit({"profile_mode": True, "ms_per_step": ms_dev / args.steps, "graph": not args.no_graph})
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- (2) end to end through the public API: every step's inputs come from pinned host memory through
    # DevicePrefetcher (copy of batch i+1 on a side stream while batch i computes, two fixed device buffer sets ->
    # two captured graphs) and the loss / score of every step are read back to the host.  One H2D copy of the full
    # batch per step happens inside the timed region.  Images are handed over as float16, the dtype the reference's
    # preprocessing writes (preprocessing/preprocess_images.py:40; its Dataset widens on the host at
    # data_preprocessing.py:174); the first-layer kernels read float16 directly.  `e2e_fp32_input` is the same loop fed
    # with the widened float32 images (twice the host->device bytes).
    class HostBatches:
        def __init__(self, v):
            self.v, self.n = v, 0

        def __iter__(self):
            for _ in range(self.n):
                yield (self.v, hq, hai, hav, hal, None, hql)

    def make_e2e(v):
        src = HostBatches(v)
        pf = D.DevicePrefetcher(src, dev, depth=args.prefetch_depth)

        def run(n):
            src.n = n
            for dbt in pf:
                loss, score = step(dbt)
                loss_host[0].copy_(loss, non_blocking=True)
                loss_host[1].copy_(score, non_blocking=True)
        return run

    e2e16, e2e32 = make_e2e(hv16), make_e2e(hv)
    e2e16(2 * args.prefetch_depth)                        # every buffer set: first sighting (eager) + capture
    ms_e2e = timed(lambda: e2e16(args.steps), 1)
    host_ms_e2e = hostt["enqueue_ms"] / args.steps
    clk = clocks.stop()          # sampled over both timed regions (device-resident and end-to-end)
    h2d_bytes = small_bytes + hv16.numel() * 2
    e2e32(2 * args.prefetch_depth)
    ms_e2e32 = timed(lambda: e2e32(args.steps), 1)
    h2d32_bytes = small_bytes + hv.numel() * 4

    # ---- (3) per-kernel breakdown: a separate pass, kernel by kernel (no graph), with CUDA events around every tagged
    # C-ABI call (the events add launch gaps, so this pass is not part of `value` / `e2e`).  Per-step collection and the
    # MEDIAN over steps per kernel: the events bracket host enqueue too, so one host hiccup (GC, the clock sampler) while
    # the queue is empty would otherwise be booked on whatever kernel came next.
    nb = min(args.steps, 7)
    per_step_times = []
    for _ in range(nb):
        lib.enable_kernel_timing(tags)
        gstep.eager_step(dbatch)
        per_step_times.append(lib.collect_kernel_timing())
    ktimes = {}
    for k in per_step_times[0]:
        ms = sorted(t[k][1] for t in per_step_times if k in t)
        ktimes[k] = (per_step_times[0][k][0] * nb, ms[len(ms) // 2] * nb)

    # ---- (4) exposed communication (N > 1): (a) the time the compute stream spends waiting for the two gradient
    # all-reduces before Adam may start (CUDA events around the waits, inside the timed region of (1)); (b) the same
    # graph-replayed step with the all-reduces removed -- the difference to `ms_per_step` is everything communication
    # costs: exposed waits plus the slow-down of the kernels it overlaps.  Run last: without the exchange the replicas
    # drift apart.
    comm = None
    if world > 1:
        model.grad_ready_hook = None
        g2 = D.GraphedTrainStep(model, opt, cfg["max_answers"], ddp=None, lr=5e-4, enabled=not args.no_graph)
        for _ in range(3):
            g2(dbatch)
        ms_nocomm = timed(lambda: g2(dbatch), args.steps)
        comm = {"finish_wait_ms_median": wait_ms[len(wait_ms) // 2] if wait_ms else None,
                "finish_wait_ms_max": wait_ms[-1] if wait_ms else None,
                "ms_per_step_without_allreduce": ms_nocomm / args.steps,
                "ms_per_step_delta": (ms_dev - ms_nocomm) / args.steps,
                "note": "finish_wait = compute-stream wait for the gradient all-reduces before Adam (rank 0, CUDA events inside "
                        "the timed region; includes the 1.5 MB image-bucket all-reduce, which cannot overlap anything); "
                        "delta = graph step with minus without all-reduce (max over ranks)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = _peaks()
    per_step = ms_dev / args.steps
    value = world * B / (per_step / 1000.0)
    e2e_val = world * B / (ms_e2e / args.steps / 1000.0)

    breakdown = {k: {"calls_per_step": n / nb, "ms_per_step": ms / nb} for k, (n, ms) in ktimes.items()}
    if "vqa_adam_multi_dev" in breakdown:
        breakdown["vqa_adam_multi"] = breakdown.pop("vqa_adam_multi_dev")
    # roofline of the dominant kernel: the tagged kernel with the largest device time whose algorithmic work is known
    esz = 2 if args.dtype in ("bf16", "bfloat16") else 4
    work = {}
    for i in (1, 2):
        for k in ("fwd", "dgrad", "wgrad"):
            work[f"conv{i}_{k}"] = ("tensor", CONV_FLOP[i] * B)
    for k in ("v_conv", "v_conv_dgrad", "v_conv_wgrad"):
        work[k] = ("tensor", 2.0 * B * 676 * 256 * 1024)
    # conv0 (K = 27): HBM-bound.  fwd reads the NCHW image (fp32 in the device-resident loop) and writes pooled bf16 +
    # uint8 mask [B,111,111,64]; the fused backward reads the image, the pooled gradient and the mask
    work["conv0_fwd"] = ("hbm", B * (3 * 224 * 224 * 4 + 111 * 111 * 64 * (esz + 1)))
    work["conv0_wgrad"] = ("hbm", B * (3 * 224 * 224 * 4 + 111 * 111 * 64 * (esz + 1)))
    work["vqa_attention_fwd"] = ("hbm", B * ((676 * 1024 + 676 * 256 + 512) * esz + 1024 * 4 + 2 * 676 * 4))
    work["vqa_attention_bwd"] = ("hbm", B * ((2 * 676 * 1024 + 2 * 676 * 256 + 512) * esz + 2 * 1024 * 4 + 2 * 676 * 4 + 2 * 1024 * 4))
    work["vqa_adam_multi"] = ("hbm", 23793242 * (4 * 7 + 2 * 0.98))       # p, g, m, v read; p, m, v written; bf16 shadows of the GEMM weights
    # LSTM recurrences: dense T = 23 FLOP of the recurrent GEMMs (what the kernels execute; mean length is 12)
    work["lstm_recurrence_fwd"] = ("tensor", 2.0 * 2 * B * 23 * 1024 * 4096)
    work["lstm_bwd_persistent"] = ("tensor", 2.0 * 2 * B * 22 * 1024 * 4096)
    roofline = None
    known = {k: v for k, v in breakdown.items() if k in work and v["calls_per_step"] > 0}
    if known:
        top = max((k for k in known if not k.startswith("lstm_")), key=lambda k: known[k]["ms_per_step"])
        bound, amount = work[top]
        dur_ms = known[top]["ms_per_step"] / known[top]["calls_per_step"]
        traffic, traffic_src = _ncu_traffic(top)
        if bound == "tensor":
            achieved, peak, unit, psrc = amount / (dur_ms / 1000.0) / 1e12, peaks["tflops"], "TFLOP/s", peaks["src"] + " sustained bf16"
        else:
            achieved, peak, unit, psrc = amount / (dur_ms / 1000.0) / 1e9, peaks["hbm_gbs"], "GB/s", peaks["src"] + " copy bandwidth"
        roofline = {"kernel": top, "bound": bound, "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak,
                    "traffic": traffic, "traffic_source": traffic_src, "peak_source": psrc,
                    "work_per_launch": amount, "ms_per_launch": dur_ms}
    per_kernel_frac = {}
    for k, v in known.items():
        bound, amount = work[k]
        d = v["ms_per_step"] / v["calls_per_step"] / 1000.0
        per_kernel_frac[k] = round((amount / d / 1e12) / peaks["tflops"] if bound == "tensor" else (amount / d / 1e9) / peaks["hbm_gbs"], 4)
    att = breakdown.get("vqa_attention_fwd")
    att_roof = None
    if att:
        byt = B * ((676 * 1024 + 676 * 256 + 512) * esz + 1024 * 4 + 2 * 676 * 4)
        gbs = byt / (att["ms_per_step"] / 1000.0) / 1e9
        att_roof = {"kernel": "vqa_attention_fwd", "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"],
                    "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"], "bytes_per_launch": byt,
                    "traffic": _ncu_traffic("vqa_attention_fwd")[0]}
    lstm = {}
    for k, steps_ in (("lstm_recurrence_fwd", 23), ("lstm_bwd_persistent", 23)):
        if k in breakdown and breakdown[k]["calls_per_step"] > 0:
            ms = breakdown[k]["ms_per_step"]
            lstm[k] = {"ms": ms, "us_per_time_step": 1000.0 * ms / steps_, "TFLOPs_dense": work[k][1] / ms / 1e9,
                       "frac_of_tensor_peak": work[k][1] / ms / 1e9 / peaks["tflops"]}

    micro_att = micro_lstm = None
    if world == 1 and not args.no_micro:
        del dbatch
        torch.cuda.empty_cache()
        from tools import microbench
        micro_att = microbench.attention_microbench(hbm_peak_gbs=peaks["hbm_gbs"])
        micro_lstm = microbench.lstm_microbench(tflops_peak=peaks["tflops"])

    # ---- the reference's own evaluated configuration (config/config_eval.yaml:52-69: stride 2, do_option '*', dropout 0.4):
    # its convolutions are outside the direct 3x3 / stride-1 kernels and run as im2col + tcgen05 GEMM + pool (im2col.cu);
    # the same step with those layers on the SIMT implicit-GEMM kernels (VQA_CONV_IM2COL=0) is timed beside it
    eval_cfg_line = None
    if world == 1 and not args.no_micro:
        import copy
        cfg2 = copy.deepcopy(cfg)
        cfg2["image"]["stride"] = 2
        cfg2["attention"]["do_option"] = "*"
        for k in ("text", "image", "attention", "classifier"):
            cfg2[k]["dropout"] = 0.4
        eval_cfg_line = {"config": "config_eval.yaml train block: stride 2, do_option '*', dropout 0.4; batch 256, bf16 arm, fwd+loss+bwd+Adam"}
        db2 = tuple(t.to(dev, non_blocking=True) for t in (hv16, hq, hai, hav, hal)) + (None, hql.to(dev, non_blocking=True))
        for key, env in (("im2col_tcgen05", "1"), ("simt_convolutions", "0")):
            os.environ["VQA_CONV_IM2COL"] = env
            torch.manual_seed(1)
            m2 = D.VqaNet(cfg2, synth.DEFAULT_TOKENS, compute_dtype=args.dtype).to(dev).train(True)
            o2 = D.FusedAdam(m2.parameters(), lr=5e-4)
            m2.use_gradient_arena(True)
            g2s = D.GraphedTrainStep(m2, o2, cfg2["max_answers"], ddp=None, lr=5e-4, enabled=not args.no_graph)
            for _ in range(3):
                g2s(db2)
            ms2 = timed(lambda: g2s(db2), 10) / 10
            eval_cfg_line[key] = {"ms_per_step": ms2, "samples_per_s": B / (ms2 / 1000.0)}
            del m2, o2, g2s
        os.environ.pop("VQA_CONV_IM2COL", None)
        torch.cuda.empty_cache()

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        times, cb, threads, kind = cpu_reference_steps(4, 1, batch=32)
        cms = sum(times) / len(times)
        cpu = {"value": cb / cms, "unit": "samples/s", "cores": threads, "kind": kind,
               "sample": f"{'unmodified reference (oracle/_ref)' if kind == 'reference' else 'oracle port'}, fp32, dropout 0.3, "
                         f"batch {cb}, 1 warm-up + 4 timed steps of fwd+loss+bwd+Adam ({cms:.2f} s/step), "
                         f"os.cpu_count()={os.cpu_count()}"}

    line = {"metric": "train samples/sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.dtype in ("bf16", "bfloat16") else "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
                       "step_launch": "one CUDA graph replay per step (GraphedTrainStep)" if not args.no_graph else "kernel by kernel",
                       "l2_policy": "inputs + activations per step (>2 GB) far exceed the 126 MB L2; no explicit flush"},
            "e2e": {"value": e2e_val, "unit": "samples/s", "h2d_bytes_per_step": h2d_bytes * world,
                    "d2h_bytes_per_step": 8 * world, "ms_per_step": ms_e2e / args.steps,
                    "input": "float16 images (the reference's stored dtype), int64 questions / answers, pinned host memory, "
                             f"DevicePrefetcher with {args.prefetch_depth} device buffer sets, loss + score read back every step"},
            "e2e_fp32_input": {"value": world * B / (ms_e2e32 / args.steps / 1000.0), "unit": "samples/s",
                               "h2d_bytes_per_step": h2d32_bytes * world, "ms_per_step": ms_e2e32 / args.steps,
                               "note": "same loop, images widened to float32 on the host as the reference's Dataset does"},
            "gpu_launches": int(launches), "gpu_launches_per_step": launches / args.steps,
            "host_enqueue_ms_per_step": {"device_resident": host_ms_dev, "e2e": host_ms_e2e},
            "host_cpu_binding": numa,
            "clocks": clk, "roofline": roofline, "attention_roofline": att_roof, "lstm_recurrence": lstm,
            "step_tensor_frac": (STEP_FLOP_PER_SAMPLE * B / (per_step / 1000.0) / 1e12) / peaks["tflops"],
            "roofline_frac_by_kernel": per_kernel_frac, "kernels": breakdown,
            "exposed_comm": comm, "attention_microbench": micro_att, "lstm_microbench": micro_lstm,
            "config_eval_yaml_step": eval_cfg_line, "cpu_baseline": cpu}
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


def _emit(line):
    """The ONE JSON line goes to the real stdout; everything else written to fd 1 during the run (NCCL's version
    banner, library chatter) was redirected to stderr in main()."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32", "bfloat16", "float32"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="enqueue the step kernel by kernel instead of replaying a CUDA graph")
    ap.add_argument("--no-micro", action="store_true", help="skip the configs[2] / configs[3] micro-benchmarks")
    ap.add_argument("--only-device", action="store_true", help="device-resident number + communication split only (A/B runs)")
    ap.add_argument("--prefetch-depth", type=int, default=3, help="device buffer sets of the end-to-end input pipeline")
    ap.add_argument("--conv-cta-group", type=int, default=0, help="override the conv kernels' tcgen05 cta_group (1 or 2)")
    ap.add_argument("--profile-mode", action="store_true", help="warm-up + timed steps only (for ncu)")
    args = ap.parse_args()
    if args.dtype == "fp32":
        args.dtype = "float32"
    if args.dtype == "bf16":
        args.dtype = "bfloat16"
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
