#!/usr/bin/env python3
"""Benchmark of the MMTM hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one forward+backward of the three MMTM blocks of the 2-view ResNet-18
(128x28^2, 256x14^2, 512x7^2) at batch 1024 per GPU, fp32 -- the top of BASELINE.json configs[1]
("MMTM fwd/bwd microbench sweep ... batch 32-1024, single B200"), i.e. the steady-state regime; the
`sweep` key repeats the measurement at batch 32 and 256 (256 = the `training_guided.gin` batch of
configs[2], which the `train` key times end to end).
Prints ONE JSON line (rank 0):

  value      algorithmic GB/s (40*N*C*HW bytes per block) over all ranks, inputs
             resident in HBM, the C-ABI calls of one step replayed from a CUDA graph
  e2e        the same metric through the public Python API (MMTM_mitigate modules) with the
             feature maps coming from PINNED HOST memory every step and the gates read back
  roofline   dominant kernel class: algorithmic bytes / CUDA-event time per launch vs the
             measured HBM copy peak (MEASURED_PEAKS.json)
  cpu_baseline  the reference's own `MMTM_mitigate` (unmodified, from the git-ignored mirror baseline/_ref;
             the oracle port where the mirror is absent) on the host cores: a bounded sample of the same
             workload at the SAME batch, all host threads
  train      guided 2-view training step end to end (cuDNN backbone + CUDA MMTM + one-call
             learning-speed statistic), samples/s, next to the CPU reference-path step

  stats      K4: the one-call (two-launch) learning-speed reduction over the real model's 142 parameters + gradients
             (190 MB), device time / GB/s, next to the reference's per-tensor loop on the host
  dp_parity  (N > 1) one block computed batch-sharded over the ranks vs rank 0's oracle on the whole batch

`--impl reference` times only the CPU reference path (the reference itself when mirrored) on the same metric/config.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHAPES = ((128, 28), (256, 14), (512, 7))  # (C, H) of mmtm2/3/4 (src/model.py:58-60)
METRIC = "mmtm_fwd_bwd_algorithmic_hbm_throughput"
UNIT = "GB/s"


def block_bytes(n, c, h):
    """Algorithmic bytes of one block fwd+bwd: 10 u, u = N*C*HW*4."""
    return 40 * n * c * h * h


def step_bytes(n):
    return sum(block_bytes(n, c, h) for c, h in SHAPES)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------
def synth_linear_params(torch, c, dev, seed):
    """nn.Linear-style uniform(-1/sqrt(fan_in), 1/sqrt(fan_in)) weights of one block: w_sq [D, 2C], b_sq, w_v [C, D],
    b_v, w_s, b_s with D = C (ratio 4).  Generated here so that the product arm never touches oracle/."""
    g = torch.Generator(device="cpu").manual_seed(seed)

    def u(shape, fan_in):
        k = 1.0 / fan_in ** 0.5
        return ((torch.rand(*shape, generator=g) * 2 - 1) * k).to(dev)

    d = c
    return [u((d, 2 * c), 2 * c), u((d,), 2 * c), u((c, d), d), u((c,), d), u((c, d), d), u((c,), d)]


class BlockBuffers:
    """Device-resident buffers of one MMTM block for direct C-ABI calls."""

    def __init__(self, torch, lib_mod, n, c, h, dev, seed):
        self.n, self.c, self.h, self.d = n, c, h, c
        g = torch.Generator(device=dev).manual_seed(seed)
        r = lambda *s: torch.randn(*s, device=dev, generator=g)
        self.a, self.b, self.go_a, self.go_b = r(n, c, h, h), r(n, c, h, h), r(n, c, h, h), r(n, c, h, h)
        self.a_out, self.b_out = torch.empty_like(self.a), torch.empty_like(self.b)
        self.d_a, self.d_b = torch.empty_like(self.a), torch.empty_like(self.b)
        self.w = synth_linear_params(torch, c, dev, seed)  # w_sq, b_sq, w_v, b_v, w_s, b_s
        self.dw = [torch.empty_like(t) for t in self.w]
        f = lambda *s: torch.empty(*s, device=dev)
        self.z, self.hid, self.g_a, self.g_b = f(n, 2 * c), f(n, c), f(n, c), f(n, c)
        self.gate_sum, self.run_v, self.run_s = f(c), torch.zeros(c, device=dev), torch.zeros(c, device=dev)
        self.dims = lib_mod.MMTMDims(n, c, c, h * h, h * h, c)
        lib = lib_mod.load()
        self.ws_bytes = lib.gml_mmtm_bwd_workspace_bytes(self.dims)
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
        self.fws_bytes = lib.gml_mmtm_fwd_workspace_bytes(self.dims)
        self.fws = torch.empty(self.fws_bytes, dtype=torch.uint8, device=dev)

    def fwd_bwd(self, lib, lib_mod, stream, flags=0):
        P = lambda t: t.data_ptr()
        w = self.w
        lib_mod.check(lib.gml_mmtm_fwd(P(self.a), P(self.b), P(self.a_out), P(self.b_out), P(w[0]), P(w[1]), P(w[2]),
                                       P(w[3]), P(w[4]), P(w[5]), P(self.z), P(self.hid), P(self.g_a), P(self.g_b),
                                       P(self.gate_sum), P(self.run_v), P(self.run_s), 0, None, None, P(self.fws),
                                       self.fws_bytes, self.dims, 0, 1.0, flags, stream), "gml_mmtm_fwd")
        dw = self.dw
        lib_mod.check(lib.gml_mmtm_bwd(P(self.go_a), P(self.go_b), P(self.a), P(self.b), P(w[0]), P(w[2]), P(w[4]),
                                       P(self.z), P(self.hid), P(self.g_a), P(self.g_b), None, None, None, None,
                                       P(self.d_a), P(self.d_b), P(dw[0]), P(dw[1]), P(dw[2]), P(dw[3]), P(dw[4]),
                                       P(dw[5]), P(self.ws), self.ws_bytes, self.dims, 0, 1.0, flags, stream),
                      "gml_mmtm_bwd")


def time_events(torch, fn, steps, warmup, sync_ranks):
    """W untimed + exactly K timed calls of fn between CUDA events on the current stream."""
    for _ in range(warmup):
        fn()
    sync_ranks()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    sync_ranks()
    return e0.elapsed_time(e1) / steps  # ms per step


def cpu_reference_blocks(torch, n):
    """The three MMTM blocks of the CPU reference path at batch n.  The UNMODIFIED reference `MMTM_mitigate`
    (baseline/_ref mirror, imported behind the gin/argh stubs of oracle/ref_loader.py) when it is there, else the oracle
    port.  Returns (step function, kind, description)."""
    from oracle import mmtm_oracle as mo
    from oracle import ref_loader
    g = torch.Generator().manual_seed(n)
    work = []
    use_ref = ref_loader.reference_available()
    ref = ref_loader.load_reference() if use_ref else None
    for c, h in SHAPES:
        x = [torch.randn(n, c, h, h, generator=g) for _ in range(4)]  # A, B, grad_A', grad_B'
        p = mo.synth_params(c, c, c)
        if use_ref:
            with ref_loader.cuda_to_cpu():
                m = ref.balanced_mmtm.MMTM_mitigate(c, c, 4)
            # the constructor pins the running means to cuda:0 (src/balanced_mmtm.py:30-31): this is the CPU path
            m.running_avg_weight_visual = m.running_avg_weight_visual.cpu()
            m.running_avg_weight_skeleton = m.running_avg_weight_skeleton.cpu()
            with torch.no_grad():
                for dst, src in zip((m.fc_squeeze.weight, m.fc_squeeze.bias, m.fc_visual.weight, m.fc_visual.bias,
                                     m.fc_skeleton.weight, m.fc_skeleton.bias), p.tensors()):
                    dst.copy_(src)
            work.append((m, x))
        else:
            work.append(((p, mo.MMTMState.zeros(c)), x))

    def once():
        for m, (a, b, ga, gb) in work:
            if use_ref:
                a_ = a.detach().requires_grad_(True)
                b_ = b.detach().requires_grad_(True)
                for q in m.parameters():
                    q.grad = None
                ao, bo, _, _ = m(a_, b_)
                torch.autograd.backward([ao, bo], [ga, gb])
            else:
                mo.forward_backward(a, b, m[0], m[1], ga, gb)

    what = ("reference src/balanced_mmtm.py MMTM_mitigate.forward + autograd backward (unmodified, CPU)" if use_ref
            else "oracle/mmtm_oracle.forward_backward (torch CPU fp32, autograd)")
    return once, ("reference" if use_ref else "port"), what


def cpu_baseline_mmtm(torch, n=1024, steps=5, warmup=1, budget_s=45.0):
    """CPU reference path on the host cores: fwd+bwd of the three blocks at batch n, `steps` timed iterations
    (fewer if the time budget runs out), all host threads."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    once, kind, what = cpu_reference_blocks(torch, n)
    t_start = time.perf_counter()
    for _ in range(warmup):
        once()
    times = []
    while len(times) < steps and (len(times) < 2 or time.perf_counter() - t_start < budget_s):
        t0 = time.perf_counter()
        once()
        times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return {"value": step_bytes(n) / med / 1e9, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "%s, 3 blocks at batch %d, median of %d iterations, %.1f ms each" % (what, n, len(times), med * 1e3),
            "ms_per_step": med * 1e3, "batch": n, "steps": len(times)}


def cpu_baseline_train(torch, batch=8, iters=3):
    """Reference-path training step on CPU: mirror model with the ORACLE MMTM + oracle statistic."""
    import greedy_multimodal_learning_b200 as pkg
    from oracle.mmtm_module import OracleMMTM
    from oracle import stats_oracle as so
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(777)
    model = pkg.MMTM_MVCNN(mmtm_cls=OracleMMTM)
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    x = torch.randn(batch, 2, 3, 224, 224)
    y = torch.randint(0, 40, (batch,))
    times = []
    for i in range(iters + 1):
        t0 = time.perf_counter()
        opt.zero_grad()
        fused, views, _, _ = model(x)
        loss = so.blend_loss(views, y)
        loss.backward()
        so.sqnorm_buckets(((n, p, p.grad) for n, p in model.named_parameters()), ["net_view_0", "net_view_1"],
                          ["visual", "skeleton"])
        opt.step()
        for t in [fused] + views:
            float(so.acc(t, y))
        float(loss)
        if i:
            times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return {"samples_per_s": batch / med, "batch": batch, "ms_per_step": med * 1e3, "cores": os.cpu_count(),
            "kind": "port", "sample": "oracle-backed MMTM_MVCNN train step (fwd+loss+bwd+learning-speed+SGD), "
                                      "batch %d, 224x224, median of %d" % (batch, iters)}


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res = cpu_baseline_mmtm(torch, n=args.batch, steps=args.steps, warmup=min(args.warmup, 3), budget_s=150.0)
    out = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
           "steps": res["steps"], "warmup": min(args.warmup, 3), "ms_per_step": res["ms_per_step"],
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "MMTM fwd+bwd, blocks 128x28^2+256x14^2+512x7^2 (mmtm2/3/4 of 2-view ResNet-18), "
                                  "batch %d, normal mode" % args.batch,
                      "batch_per_gpu": args.batch, "algorithmic_bytes_per_step_per_gpu": step_bytes(args.batch),
                      "host": "CPU reference path on %d host threads (rank 0 only)" % res["cores"]},
           "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
           "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)


PATHS = {  # --path -> (flags, tunables)
    "auto": (0, {}),
    "tile": ("F_FORCE_TILE", {}),
    "old": (0, {"tile_kind": 2}),        # round-1 selection: cluster kernels / streaming, no tile pipeline
    "streaming": ("F_FORCE_STREAMING", {}),
    "fused": ("F_FORCE_FUSED", {}),
}
L2_BYTES = 126 << 20


def run_ours(args):
    import torch
    import torch.distributed as dist

    import greedy_multimodal_learning_b200 as pkg
    from greedy_multimodal_learning_b200 import _lib as L, dist as gdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        gdist.init_from_env("nccl")
    lib = L.load()
    if lib.gml_device_is_blackwell() != 1:
        print("warning: not a compute-capability-10 device", file=sys.stderr)
    flags, tunables = PATHS[args.path]
    flags = getattr(L, flags) if isinstance(flags, str) else flags
    for k, v in tunables.items():
        L.check(lib.gml_set_tunable(k.encode(), v))

    def sync_ranks():
        if world > 1:
            dist.barrier()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peak, peak_src = measured_peak()
    stream = torch.cuda.Stream()

    def graph_of_steps(nb):
        """CUDA graph of `sets` consecutive steps, each on its own buffer set: enough sets that what one step
        leaves in the 126 MB L2 is gone when its buffers come round again (8u per set)."""
        ws = 8 * sum(nb * c * h * h * 4 for c, h in SHAPES)
        sets = max(1, -(-2 * L2_BYTES // ws))
        bufs = [[BlockBuffers(torch, L, nb, c, h, dev, seed=c + rank + 97 * i) for c, h in SHAPES] for i in range(sets)]
        with torch.cuda.stream(stream):
            before = lib.gml_launch_count(-1)
            for b_ in bufs[0]:
                b_.fwd_bwd(lib, L, stream.cuda_stream, flags)
            launches = lib.gml_launch_count(-1) - before
            for bs in bufs[1:]:
                for b_ in bs:
                    b_.fwd_bwd(lib, L, stream.cuda_stream, flags)
            stream.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                for bs in bufs:
                    for b_ in bs:
                        b_.fwd_bwd(lib, L, stream.cuda_stream, flags)
        return g, bufs, sets, launches

    def time_groups(g, sets, steps, warmup, min_total_s=0.5, min_groups=5):
        """median over >= 5 groups of (max over ranks of) the CUDA-event time of EXACTLY `steps` steps"""
        reps = -(-steps // sets)  # one replay = `sets` steps
        with torch.cuda.stream(stream):
            first = max_over_ranks(time_events(torch, g.replay, reps, -(-warmup // sets), sync_ranks))
            groups = [first]
            n_groups = max(min_groups, int(min_total_s / max(first * reps * 1e-3, 1e-6)) + 1)
            n_groups = min(n_groups, 400)
            if world > 1:  # every rank must run the same number of groups
                t = torch.tensor([n_groups], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                n_groups = int(t.item())
            for _ in range(n_groups - 1):
                groups.append(max_over_ranks(time_events(torch, g.replay, reps, 0, sync_ranks)))
        return statistics.median(groups) / sets, len(groups), reps * sets

    # ---- value: C-ABI step replayed from a CUDA graph, inputs resident in HBM -------------------
    n = args.batch
    graph, bufsets, sets, launches_per_step = graph_of_steps(n)
    blocks = bufsets[0]
    with ClockSampler(local) as clk:
        ms, n_groups, steps_per_group = time_groups(graph, sets, args.steps, args.warmup)
    clocks = clk.summary()
    total_bytes = step_bytes(n) * world
    value = total_bytes / (ms * 1e-3) / 1e9
    u_total = sum(n * c * h * h * 4 for c, h in SHAPES)  # one modality, all three blocks

    # ---- the rest of the configs[1] sweep (same protocol) -------------------------------------------
    sweep = {}
    for nb in (32, 256):
        if nb == n:
            continue
        g2, b2, s2, l2 = graph_of_steps(nb)
        ms_b, _, _ = time_groups(g2, s2, max(args.steps, 10), 3, min_total_s=0.2)
        sweep["batch_%d" % nb] = {"ms_per_step": ms_b, "value": step_bytes(nb) * world / (ms_b * 1e-3) / 1e9,
                                  "unit": UNIT, "frac_of_measured_hbm_peak": step_bytes(nb) / (ms_b * 1e-3) / 1e9 / peak,
                                  "launches_per_step": int(l2), "buffer_sets_rotated": s2}
        del g2, b2
        torch.cuda.empty_cache()

    # ---- roofline: per-kernel-class CUDA-event times, block by block (eager, profiled) ----------
    # algorithmic bytes per launch in units of u = N*C*HW*4 (one modality): SURVEY 8d / DESIGN.md
    alg_units = {"plane_mean": 2, "plane_scale_fwd": 4, "plane_dgate": 4, "plane_scale_bwd": 4, "fused_fwd": 4,
                 "fused_bwd": 6}
    prof_steps = max(3, min(args.steps, 10))
    kernels, dominant, step_prof_ms, per_block = {}, None, 0.0, []
    tfile = os.path.join(ROOT, "profiles", "traffic.json")
    traffic_db = json.load(open(tfile)) if os.path.isfile(tfile) else {}
    flush = torch.empty(2 * L2_BYTES, dtype=torch.uint8, device=dev)
    with torch.cuda.stream(stream):
        for blk in blocks:
            u_blk = blk.n * blk.c * blk.h * blk.h * 4
            lib.gml_profile_reset()
            lib.gml_profile_enable(1)
            for _ in range(prof_steps):
                flush.zero_()
                blk.fwd_bwd(lib, L, stream.cuda_stream, flags)
            stream.synchronize()
            lib.gml_profile_enable(0)
            fwd_ms = bwd_ms = 0.0
            for tag in range(lib.gml_kernel_tag_count()):
                tot, cnt = ctypes.c_double(), ctypes.c_int64()
                lib.gml_profile_read(tag, ctypes.byref(tot), ctypes.byref(cnt))
                if not cnt.value:
                    continue
                name = lib.gml_kernel_tag_name(tag).decode()
                key = "%dx%d^2/%s" % (blk.c, blk.h, name)
                per_step_ms = tot.value / prof_steps
                step_prof_ms += per_step_ms
                entry = {"launches_per_step": cnt.value / prof_steps, "ms_per_step": per_step_ms}
                if name in alg_units:
                    entry["algorithmic_bytes_per_launch"] = alg_units[name] * u_blk / (cnt.value / prof_steps)
                    entry["algorithmic_gbs"] = alg_units[name] * u_blk / (per_step_ms * 1e-3) / 1e9
                    tr = traffic_db.get("%s@%d" % (key, n), {})
                    entry["dram_bytes_per_launch"] = tr.get("traffic_bytes_per_launch")
                    if dominant is None or per_step_ms > kernels[dominant]["ms_per_step"]:
                        dominant = key
                kernels[key] = entry
            # per block and direction: the single tile / cluster kernel, or the sum of the streaming kernels
            for d_, names, units in (("fwd", ("fused_fwd", "plane_mean", "plane_scale_fwd"), 4),
                                     ("bwd", ("fused_bwd", "plane_dgate", "plane_scale_bwd"), 6)):
                t_ = sum(kernels.get("%dx%d^2/%s" % (blk.c, blk.h, nm), {}).get("ms_per_step", 0.0) for nm in names)
                dram = [kernels.get("%dx%d^2/%s" % (blk.c, blk.h, nm), {}).get("dram_bytes_per_launch") for nm in names]
                dram = [x for x in dram if x]
                if t_ > 0:
                    per_block.append({"shape": "%dx%d^2" % (blk.c, blk.h), "dir": d_, "algorithmic_bytes": units * u_blk,
                                      "dram_bytes": sum(dram) if dram else None, "ms": t_,
                                      "frac": units * u_blk / (t_ * 1e-3) / 1e9 / peak})
    del flush
    roof = None
    if dominant:
        k = kernels[dominant]
        roof = {"bound": "hbm", "kernel": dominant, "achieved": k["algorithmic_gbs"], "peak": peak, "unit": "GB/s",
                "frac": k["algorithmic_gbs"] / peak, "traffic": k.get("dram_bytes_per_launch"),
                "peak_source": peak_src + ", burst copy",
                "algorithmic_bytes_per_launch": k["algorithmic_bytes_per_launch"],
                "avg_launch_ms": k["ms_per_step"] / k["launches_per_step"],
                "share_of_step": k["ms_per_step"] / step_prof_ms, "whole_step_frac": value / world / peak,
                "per_block": per_block,
                "note": "per-launch CUDA-event times with L2 flushed before every call (cold inputs); traffic = ncu "
                        "dram read+write bytes of the same kernel from profiles/traffic.json (null where not captured)"}

    # ---- e2e: public Python API, feature maps from pinned host memory, gates read back ----------
    mods = []
    for b in blocks:
        m = pkg.MMTM_mitigate(b.c, b.c, 4, kernel_flags=flags)
        with torch.no_grad():
            for dst, src in zip((m.fc_squeeze.weight, m.fc_squeeze.bias, m.fc_visual.weight, m.fc_visual.bias,
                                 m.fc_skeleton.weight, m.fc_skeleton.bias), b.w):
                dst.copy_(src)
        m.sync_running_stats = False  # the microbenchmark has no collective (per-sample operator, batch sharded)
        mods.append(m.to(dev))
    host = [(b.a.cpu().pin_memory(), b.b.cpu().pin_memory()) for b in blocks]
    h2d = sum(a.numel() * 4 + bb.numel() * 4 for a, bb in host)
    d2h = sum(2 * b.n * b.c * 4 for b in blocks)

    copy_stream = torch.cuda.Stream(device=dev)

    def e2e_step():
        # the step's six host->device copies go out on a copy stream, block by block; the compute stream picks each pair
        # up through an event, so block i's kernels overlap block i+1's copies (what framework.DevicePrefetcher does for
        # training batches).  Every byte still crosses PCIe inside the timed region, every step.
        cur = torch.cuda.current_stream(dev)
        copy_stream.wait_stream(cur)
        staged = []
        with torch.cuda.stream(copy_stream):
            for ha, hb in host:
                a = ha.to(dev, non_blocking=True)
                bb = hb.to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                staged.append((a, bb, ev))
        gates = []
        for m, b, (a, bb, ev) in zip(mods, blocks, staged):
            cur.wait_event(ev)
            a.record_stream(cur)
            bb.record_stream(cur)
            a.requires_grad_(True)
            bb.requires_grad_(True)
            a_out, b_out, scales, _ = m(a, bb, True)          # return_scale=True: gates come back to the host
            torch.autograd.backward([a_out, b_out], [b.go_a, b.go_b])
            gates.append(scales)
        return gates

    e2e_steps = max(3, min(args.steps, 10))
    ms_e2e = max_over_ranks(time_events(torch, e2e_step, e2e_steps, min(args.warmup, 3), sync_ranks))
    e2e = {"value": total_bytes / (ms_e2e * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e, "steps": e2e_steps,
           "api": "MMTM_mitigate.forward(return_scale=True) + autograd.backward, inputs from pinned host memory on a copy "
                  "stream (block i+1 copies while block i computes), gates read back"}
    del host, mods, bufsets, blocks, graph
    torch.cuda.empty_cache()

    # ---- K4: the learning-speed reduction on the real model ------------------------------------------
    stats = bench_stats(torch, pkg, L, lib, dev, peak, with_cpu=(rank == 0 and world == 1)) if not args.no_stats else None
    # ---- data-parallel correctness carried by the scaling run ---------------------------------------
    dp_parity = bench_dp_parity(torch, pkg, dev, world, rank, flags) if world > 1 else None
    # ---- train: guided 2-view training step end to end ----------------------------------------------
    train = train_strong = util = None
    if not args.no_train:
        train = bench_train(torch, pkg, gdist, dev, world, rank, args, sync_ranks, max_over_ranks, args.train_batch)
        if world > 1 and args.train_global_batch % world == 0 and args.train_global_batch // world <= 1024:
            # BASELINE configs[4]: global batch 2048 sharded over 2/4/8 GPUs (strong scaling)
            train_strong = bench_train(torch, pkg, gdist, dev, world, rank, args, sync_ranks, max_over_ranks,
                                       args.train_global_batch // world)
            train_strong["scaling"] = "strong"
        util = bench_utilization(torch, pkg, gdist, dev, world, rank, args, sync_ranks, max_over_ranks)

    if rank == 0:
        # CPU baselines are a single-process, N=1 measurement (torchrun pins OMP threads for N>1)
        cpu = cpu_baseline_mmtm(torch, n=n) if world == 1 else None
        if train is not None and not args.no_cpu_train and world == 1:
            train["cpu_reference"] = cpu_baseline_train(torch)
            train["speedup_vs_cpu_reference"] = train["samples_per_s"] / train["cpu_reference"]["samples_per_s"]
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": "MMTM fwd+bwd, blocks 128x28^2+256x14^2+512x7^2 (mmtm2/3/4 of 2-view ResNet-18), "
                                      "batch %d per GPU, normal mode" % n,
                          "batch_per_gpu": n, "algorithmic_bytes_per_step_per_gpu": step_bytes(n),
                          "cache": "per-step working set 8u = %.2f GB per GPU > 126 MB L2 (no flush needed)"
                                   % (8 * u_total / 1e9),
                          "kernel_path": args.path, "launch": "CUDA graph replay of the C-ABI calls",
                          "timing": "median of %d groups of exactly %d steps (CUDA events, max over ranks per group), "
                                    "clock sampler running during all of them" % (n_groups, steps_per_group)},
               "clocks": clocks, "gpu_launches": int(launches_per_step * args.steps), "launches_per_step":
               int(launches_per_step), "e2e": e2e, "roofline": roof, "kernels": kernels,
               "cpu_baseline": ({k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")} if cpu else None),
               "frac_of_measured_hbm_peak": value / world / peak, "sweep": sweep, "stats": stats,
               "dp_parity": dp_parity, "train": train, "train_strong": train_strong, "utilization": util,
               "train_samples_per_s": train["samples_per_s"] if train else None,
               "train_strong_samples_per_s": train_strong["samples_per_s"] if train_strong else None,
               "e2e_value": e2e["value"]}
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def bench_stats(torch, pkg, L, lib, dev, peak, with_cpu):
    """K4: sum of squares of the 142 parameters and 142 gradients of MMTM_MVCNN (2 x 23,773,008 fp32 =
    190.2 MB) in one call = a scan launch + a one-cluster fold launch (programmatic dependent launch); device time from
    CUDA events around the C-ABI call, L2 flushed before each call.  The
    reference's loop (src/callbacks.py:203-205: two reductions + two .item() per tensor) is timed on the host beside it."""
    BR, MM = ["net_view_0", "net_view_1"], ["visual", "skeleton"]
    torch.manual_seed(777)
    model = pkg.MMTM_MVCNN().to(dev)
    gen = torch.Generator(device=dev).manual_seed(1)
    for p in model.parameters():
        p.grad = torch.randn(p.shape, device=dev, generator=gen) * 0.01
    sq = pkg.MultiTensorSqnorm(model.named_parameters(), BR, MM)
    got = sq.measure()  # builds the table; one full call incl. the read-back
    ptrs, numel, masks, kinds, nt = sq._table
    nbytes = 4 * sum(int(x) for x in numel)
    flush = torch.empty(2 * L2_BYTES, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream(dev)
    times = []
    for i in range(12):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(lib.gml_multi_tensor_sqnorm(ptrs, numel, masks, kinds, nt, sq._out.data_ptr(), None, sq._ws.data_ptr(),
                                            sq._ws_bytes, st.cuda_stream))
        e1.record()
        e1.synchronize()
        if i >= 2:
            times.append(e0.elapsed_time(e1))
    t0 = time.perf_counter()
    for _ in range(5):
        sq.measure()
    full_ms = (time.perf_counter() - t0) / 5 * 1e3
    ms = statistics.median(times)
    out = {"tensors": nt, "algorithmic_bytes": nbytes, "device_ms": ms, "gbs": nbytes / (ms * 1e-3) / 1e9,
           "frac_of_measured_hbm_peak": nbytes / (ms * 1e-3) / 1e9 / peak, "call_ms_incl_readback": full_ms,
           "launches_per_call": 2,
           "note": "gml_multi_tensor_sqnorm over MMTM_MVCNN (src/callbacks.py:203-223), cold L2; device_ms spans both "
                   "launches (scan over all tensors, then a one-cluster fold); profiles/r2_sqnorm.md has the per-kernel "
                   "ncu times (scan alone 32.9 us = 0.88 of the measured HBM peak)"}
    if with_cpu:
        from oracle import stats_oracle as so
        cpu_named = [(n_, p_.detach().cpu(), p_.grad.cpu()) for n_, p_ in model.named_parameters()]
        torch.set_num_threads(os.cpu_count() or 1)
        so.sqnorm_buckets(iter(cpu_named), BR, MM)
        t0 = time.perf_counter()
        for _ in range(5):
            want = so.sqnorm_buckets(iter(cpu_named), BR, MM)
        out["cpu_reference_loop_ms"] = (time.perf_counter() - t0) / 5 * 1e3
        out["max_rel_err_vs_reference_loop"] = max(abs(got[k][i] - want[k][i]) / want[k][i] for k in want for i in (0, 1))
    return out


def bench_dp_parity(torch, pkg, dev, world, rank, flags):
    """One block (256x14^2, global batch 64 * world) computed batch-sharded over the ranks -- gate-sum all-reduce
    inside the forward, weight gradients summed over ranks -- against rank 0's oracle on the whole batch."""
    import torch.distributed as dist
    c, h, per = 256, 14, 64
    n = per * world
    g = torch.Generator().manual_seed(4242)
    full = [torch.randn(n, c, h, h, generator=g) for _ in range(4)]
    w = synth_linear_params(torch, c, "cpu", 11)
    m = pkg.MMTM_mitigate(c, c, 4, kernel_flags=flags)
    with torch.no_grad():
        for dst, src in zip((m.fc_squeeze.weight, m.fc_squeeze.bias, m.fc_visual.weight, m.fc_visual.bias,
                             m.fc_skeleton.weight, m.fc_skeleton.bias), w):
            dst.copy_(src)
    m = m.to(dev)
    lo, hi = rank * per, (rank + 1) * per
    a = full[0][lo:hi].to(dev).requires_grad_(True)
    b = full[1][lo:hi].to(dev).requires_grad_(True)
    a_out, b_out, _, _ = m(a, b)
    torch.autograd.backward([a_out, b_out], [full[2][lo:hi].to(dev), full[3][lo:hi].to(dev)])
    grads = [p.grad.clone() for p in m.parameters()]
    for t in grads:
        dist.all_reduce(t)
    worst = 0.0
    if rank == 0:
        from oracle import mmtm_oracle as mo
        p = mo.MMTMParams(*w)
        st = mo.MMTMState.zeros(c)
        o = mo.forward_backward(full[0], full[1], p, st, full[2], full[3])
        rel = lambda x, r: float((x.detach().cpu().double() - r.double()).abs().max() / r.double().abs().max())
        checks = {"A_out": rel(a_out, o["A_out"][lo:hi]), "B_out": rel(b_out, o["B_out"][lo:hi]),
                  "dA": rel(a.grad, o["dA"][lo:hi]), "dB": rel(b.grad, o["dB"][lo:hi]),
                  "run_v": rel(m.running_avg_weight_visual, st.run_v)}
        for t, k in zip(grads, ("dWsq", "dbsq", "dWv", "dbv", "dWs", "dbs")):
            checks[k] = rel(t, o[k])
        worst = max(checks.values())
        return {"max_rel_err": worst, "per_output": checks, "tolerance": 1e-5, "ok": worst <= 1e-5,
                "config": "256x14^2, %d samples per rank, %d ranks, vs oracle on the concatenated batch" % (per, world)}
    return None


def bench_train(torch, pkg, gdist, dev, world, rank, args, sync_ranks, max_over_ranks, bsz):
    """training_guided.gin step: batch `bsz` per GPU, data parallel when world > 1."""
    gdist.seed_everything(777)
    torch.backends.cudnn.benchmark = True  # let cuDNN pick its convolution algorithms during the warm-up steps
    model = pkg.MMTM_MVCNN().to(dev)
    if world > 1:
        model, opt, reducer = gdist.setup_model(model, lambda p: torch.optim.SGD(p, lr=0.1, momentum=0, weight_decay=0))
    else:
        opt, reducer = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0, weight_decay=0), None
    cb = pkg.Bias_Mitigation_Strong(0.01, 5, ["net_view_0", "net_view_1"], 1)
    cb.set_model(model, ignore=False)
    engine = pkg.Model_(model, opt, pkg.blend_loss, 2, metrics=[pkg.acc], data_parallel=reducer).to(dev)
    cbs = pkg.CallbackList([cb])
    cbs.set_model_pytoune(engine)
    cbs.on_train_begin({})
    cbs.on_epoch_begin(1, {})
    g = torch.Generator().manual_seed(rank)
    x = torch.randn(bsz, 2, 3, 224, 224, generator=g).pin_memory()
    y = torch.randint(0, 40, (bsz,), generator=g).pin_memory()
    model.train(True)
    state = {"i": 0}

    def endless():
        while True:
            yield None, x, y

    # every step's batch is copied host->device from pinned memory inside the timed region; the copy of
    # batch i+1 runs on a side stream under the compute of batch i (DevicePrefetcher)
    batches = iter(pkg.DevicePrefetcher(endless(), dev))

    def step():
        state["i"] += 1
        s = {"number": state["i"], "indices": None}
        _, xd, yd = next(batches)
        engine.train_step(s, xd, yd, cbs)  # loss/accuracy read-back inside

    steps = max(3, min(args.steps, 8))
    ms = max_over_ranks(time_events(torch, step, steps, 3, sync_ranks))
    sps = bsz * world / (ms * 1e-3)
    peak, _ = measured_peak()
    roofline_sps = peak * 1e9 / 7_024_640 * world
    out = {"samples_per_s": sps, "ms_per_step": ms, "global_batch": bsz * world, "batch_per_gpu": bsz, "steps": steps,
           "scaling": "weak",
           "config": "training_guided.gin: 2-view ResNet-18 + MMTM, 224x224, SGD lr 0.1, Bias_Mitigation_Strong "
                     "eps 0.01 window 5, fp32 (cuDNN TF32 default on, cudnn.benchmark), per-step H2D of the batch (prefetched one step ahead) "
                     "and loss/acc read-back timed",
           "frac_of_mmtm_memory_roofline": sps / roofline_sps,
           "curation_mode_at_end": bool(engine.curation_mode)}
    del engine, model, opt, batches, x, y
    torch.cuda.empty_cache()
    return out


def bench_utilization(torch, pkg, gdist, dev, world, rank, args, sync_ranks, max_over_ranks):
    """BASELINE configs[3] (recording.gin + eval.gin): squeeze-mean recording over the ranks' shards (fp64 device sums,
    one all-reduce at the end) and the flow-cut evaluation with those means; samples/s of each pass, and whether every
    rank ends up with bit-identical dataset means."""
    import torch.distributed as dist
    bsz, n_batches = 64, 4
    gdist.seed_everything(777)
    model = pkg.MMTM_MVCNN().to(dev).eval()
    g = torch.Generator().manual_seed(100 + rank)
    xs = [torch.randn(bsz, 2, 3, 224, 224, generator=g).pin_memory() for _ in range(n_batches)]
    ys = [torch.randint(0, 40, (bsz,), generator=g) for _ in range(n_batches)]
    rec = pkg.SqueezeMeanRecorder(model.mmtm_blocks())

    def record_pass():
        with torch.no_grad():
            for x in xs:
                model(x.to(dev, non_blocking=True))
                rec.update(None)

    record_pass()  # warm-up (cuDNN autotune); its sums are discarded
    rec = pkg.SqueezeMeanRecorder(model.mmtm_blocks())
    ms_rec = max_over_ranks(time_events(torch, record_pass, 1, 0, sync_ranks))
    means = rec.result(device=dev)
    identical = True
    if world > 1:
        flat = torch.cat([v.flatten() for blk in means[1:] for v in blk])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        identical = all(torch.equal(gathered[0], t) for t in gathered)
    off = pkg.MMTM_MVCNN(mmtm_off=True, mmtm_rescale=means).to(dev).eval()
    off.load_state_dict(model.state_dict())
    counts = torch.zeros(3, dtype=torch.int64, device=dev)

    def eval_pass():
        with torch.no_grad():
            for x, y in zip(xs, ys):
                fused, views, _, _ = off(x.to(dev, non_blocking=True))
                yd = y.to(dev)
                counts[0] += (fused.argmax(1) == yd).sum()
                counts[1] += (views[0].argmax(1) == yd).sum()
                counts[2] += (views[1].argmax(1) == yd).sum()

    eval_pass()
    counts.zero_()
    ms_eval = max_over_ranks(time_events(torch, eval_pass, 1, 0, sync_ranks))
    if world > 1:
        dist.all_reduce(counts)
    total = bsz * n_batches * world
    out = {"recording_samples_per_s": total / (ms_rec * 1e-3), "flow_cut_eval_samples_per_s": total / (ms_eval * 1e-3),
           "samples": total, "dataset_means_rank_identical": bool(identical),
           "acc_flow_cut": [100.0 * float(c) / total for c in counts.tolist()],
           "config": "recording.gin + eval.gin on synthetic 224x224 views: %d samples per rank, SqueezeMeanRecorder "
                     "(gml_squeeze_accumulate, fp64 sums, one all-reduce), then mmtm_off evaluation" % (bsz * n_batches)}
    del model, off, xs
    torch.cuda.empty_cache()
    return out


_REAL_STDOUT = None


def emit(obj):
    """The ONE JSON line goes to the real stdout; everything libraries print (e.g. NCCL's version banner)
    was redirected to stderr in main()."""
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, line)
    else:
        sys.stdout.write(line.decode())
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="MMTM batch per GPU")
    ap.add_argument("--train-batch", type=int, default=256, help="training batch per GPU")
    ap.add_argument("--train-global-batch", type=int, default=2048, help="strong-scaling leg: global batch over all GPUs")
    ap.add_argument("--path", default="auto", choices=list(PATHS))
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--no-stats", action="store_true")
    ap.add_argument("--no-cpu-train", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
