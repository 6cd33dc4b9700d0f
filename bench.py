#!/usr/bin/env python
"""bench.py — belief propagate + chance-check steps/s (BASELINE.json metric) on N B200s.

A "step" is one pass of the fused propagateWhileValid kernel over one batch of synthetic
beliefs (BASELINE.json config 5: B beliefs x T=50 steps x M=256 obstacles, linear model,
PCCBlackmore checker, seed 20240501).  Units = executed belief-steps (one propagate + one
validity check against all M obstacles).

  value     device-resident inputs, CUDA-event timed, K launches, max over ranks
  e2e       the same metric through the host-pointer C-ABI call (pinned host buffers,
            H2D + kernel + D2H inside the timed region)
  roofline  F_alg (oracle-counted algorithmic flops) / kernel time vs the FP64
            pipe peak measured live; plus the kernel's algorithmic HBM bytes vs MEASURED_PEAKS
  cpu_baseline  the CPU oracle ("port": the reference cannot be built here) on a bounded sample

`--impl reference` times the oracle alone on the host cores (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

T_STEPS = 50
M_OBS = 256
METRIC = "belief propagate+chance-check steps/sec"
UNIT = "belief-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2-beliefs", type=int, default=24, help="beliefs per GPU (weak scaling)")
    ap.add_argument("--variant", default="all_survive", choices=["all_survive", "realistic"])
    ap.add_argument("--model", default="linear", choices=["linear", "unicycle"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-realistic", action="store_true", help="skip the secondary realistic-variant measurement")
    ap.add_argument("--no-solve", action="store_true", help="skip the short live CCK-CBS solve-time sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs under ncu)")
    return ap.parse_args()


def workload_name(args, B):
    return (f"config5 synthetic sweep: {B} beliefs/GPU x {T_STEPS} steps x {M_OBS} obstacles, "
            f"{'2DUncertainLinear' if args.model == 'linear' else 'Uncertain-Unicycle'}, PCCBlackmore, {args.variant}, seed 20240501")


# ----------------------------------------------------------------------------- CPU oracle legs
def oracle_setup(args):
    from oracle import oracle as O
    import cck_b200  # noqa: F401  (workload generator only; no GPU use)
    from cck_b200 import workloads as W
    centres = W.synthetic_obstacles(M_OBS, W.SEED, args.variant)
    world = O.World.synthetic(centres, float(W.WORLD), float(W.WORLD), 4, 0.9)
    dim = 2 if args.model == "linear" else 4
    return O, W, O.Svc(world, O.SVC_BLACKMORE, dim)


def oracle_time(args, O, W, svc, n, threads, offset=1000):
    bel, ctrl = W.synthetic_beliefs(n, args.model, variant=args.variant, offset=offset)
    aos_b, aos_c = np.ascontiguousarray(bel.T), np.ascontiguousarray(ctrl.T)
    steps = np.full(n, T_STEPS, dtype=np.int32)
    t0 = time.perf_counter()
    _, valid, cnt = svc.rollout(aos_b, aos_c, steps, nthreads=threads)
    dt = time.perf_counter() - t0
    return dt, int(cnt[0]), int(cnt[1])


def reference_sources_timing(args, W):
    """Informational: the reference's OWN sources (oracle/_ref, compiled against the stand-in headers of oracle/shim/)
    on a small sample of the same workload, one thread (the reference is single-threaded).  Not the official
    baseline: the stand-in Matrix heap-allocates where Eigen's fixed-size types do not, so this under-states a real
    build of the reference; the (faster) oracle port above is the conservative number."""
    try:
        import tempfile
        from oracle import ref as R
        if args.model != "linear" or not R.available():
            return None
        centres = W.synthetic_obstacles(M_OBS, W.SEED, args.variant)
        width = int(centres[:, 0].max()) + 2 if args.variant == "all_survive" else W.WORLD    # the parser sets x_max = map width
        with tempfile.TemporaryDirectory() as d:
            rows = [["."] * width for _ in range(W.WORLD)]
            for x, y in centres:
                rows[int(y)][int(x)] = "@"
            mp, sc = os.path.join(d, "synthetic.map"), os.path.join(d, "synthetic.scen")
            open(mp, "w").write(f"type octile\nheight {W.WORLD}\nwidth {width}\nmap\n" + "\n".join("".join(r) for r in rows) + "\n")
            open(sc, "w").write("version 1\n" + "".join(f"{i}\tsynthetic.map\t{width}\t{W.WORLD}\t{2 + i}\t2\t{5 + i}\t5\tRectangle\t2D-Uncertain-Linear-Model\t1.0\n" for i in range(4)))
            world = R.World(mp, sc, 4, 0.9)
            svc = R.Svc(world)
        n = 512
        bel, ctrl = W.synthetic_beliefs(n, "linear", variant=args.variant, offset=3000)
        steps = np.full(n, T_STEPS, dtype=np.int32)
        t0 = time.perf_counter()
        _, valid = svc.rollout(np.ascontiguousarray(bel.T), np.ascontiguousarray(ctrl.T), steps)
        dt = time.perf_counter() - t0
        units = int(np.minimum(valid.astype(np.int64) + 1, T_STEPS).sum())
        return {"single_thread_value": units / dt, "unit": UNIT, "sample": f"{n} beliefs ({units} belief-steps, {dt:.1f} s)",
                "note": "reference sources + stand-in Eigen/Boost/OMPL headers (oracle/_ref); informational, see DESIGN.md section 4"}
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)[:200]}


def cpu_baseline(args):
    """Oracle timed on all host cores on a bounded sample of the same workload."""
    O, W, svc = oracle_setup(args)
    cores = os.cpu_count() or 1
    dt, nsteps, _ = oracle_time(args, O, W, svc, 4096, cores)
    rate = nsteps / dt
    n = int(min(1 << 20, max(8192, rate * args.cpu_seconds / T_STEPS)))
    dt, nsteps, nhp = oracle_time(args, O, W, svc, n, cores)
    dt1, nsteps1, _ = oracle_time(args, O, W, svc, max(2048, n // (4 * cores)), 1)
    return {"value": nsteps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} beliefs x {T_STEPS} steps x {M_OBS} obstacles ({nsteps} belief-steps, {dt:.1f} s, std::thread over beliefs)",
            "single_thread_value": nsteps1 / dt1, "halfplane_tests_per_belief_step": nhp / nsteps,
            "reference_sources": reference_sources_timing(args, W)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    O, W, svc = oracle_setup(args)
    cores = os.cpu_count() or 1
    dt, nsteps, _ = oracle_time(args, O, W, svc, 4096, cores)
    n = int(min(1 << 19, max(8192, (nsteps / dt) * 8.0 / T_STEPS)))     # ~8 s of CPU work per step
    n = min(n, max(8192, int((nsteps / dt) * 150.0 / T_STEPS / max(1, args.steps + args.warmup))))
    times, units = [], 0
    for i in range(args.warmup + args.steps):
        dt, nsteps, _ = oracle_time(args, O, W, svc, n, cores, offset=2000 + i)
        if i >= args.warmup:
            times.append(dt); units += nsteps
    total = sum(times)
    val = units / total
    sample = f"{n} beliefs x {T_STEPS} steps x {M_OBS} obstacles per step"
    B = 1 << args.log2_beliefs
    print(json.dumps({
        "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "impl": "reference",
        "config": {"workload": workload_name(args, B), "sample_per_step": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ----------------------------------------------------------------------------- clocks sampler
class Clocks:
    """SM clock + throttle reasons sampled DURING the timed region: NVML in-process every 5 ms (the timed region of the
    default run is ~0.1 s, too short for a freshly spawned nvidia-smi), nvidia-smi -lms as the fallback."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.lines, self.proc, self.nv = [], None, None
        self.sm, self.mx, self.reasons, self.run = [], None, set(), True
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nv
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}   # nvmlClocksEventReasons*
        while self.run:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                r = int(get(self.h))
                for k, b in bits.items():
                    if r & b:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.nv:
            self.run = False
            self.th.join(timeout=1)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx, "reasons": sorted(self.reasons),
                    "samples": len(self.sm), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0])); mx = float(p[1])
            except ValueError:
                continue
            for nme, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


# ----------------------------------------------------------------------------- GPU leg
def run_ours(args):
    import torch
    import torch.distributed as dist
    import cck_b200 as K
    from cck_b200 import workloads as W

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B = 1 << args.log2_beliefs
    dim = 2 if args.model == "linear" else 4
    Wd = K.width(dim)
    ctx = K.Context(local)
    W.install_synthetic(ctx, args.model, M_OBS, variant=args.variant)

    # host inputs (pinned) — each rank owns a disjoint shard (weak scaling, no data-path collective)
    bel_np, ctrl_np = W.synthetic_beliefs(B, args.model, variant=args.variant, offset=rank)
    h_bel = torch.from_numpy(bel_np).pin_memory()
    h_ctrl = torch.from_numpy(ctrl_np).pin_memory()
    h_steps = torch.full((B,), T_STEPS, dtype=torch.int32).pin_memory()
    h_out = torch.empty((Wd, B), dtype=torch.float64).pin_memory()
    h_valid = torch.empty((B,), dtype=torch.int32).pin_memory()
    del bel_np, ctrl_np

    d_bel, d_ctrl, d_steps = h_bel.to(dev), h_ctrl.to(dev), h_steps.to(dev)
    d_out = torch.empty((Wd, B), dtype=torch.float64, device=dev)
    d_valid = torch.empty((B,), dtype=torch.int32, device=dev)
    d_sum = torch.zeros(3, dtype=torch.int64, device=dev)
    from cck_b200 import sharding

    # a non-default torch stream: its handle is non-NULL (NULL would select the context's own stream)
    # and torch.cuda.Event records on it, so events bracket exactly the kernels launched below
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    def step():
        ctx.propagate_while_valid_dev(stream, B, d_bel, B, d_ctrl, B, d_steps, d_out, B, d_valid)
        ctx.rollout_summary_dev(stream, B, d_steps, d_valid, d_sum)
        if world > 1:   # only the per-rank result record crosses NVLink
            sharding.gather_summary(d_sum)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    exec_steps = int(d_sum[1].item())

    # ---- timed region: device-resident inputs (1.7 GB/step of operands >> 126 MB L2) ----
    barrier()
    clocks = Clocks(local)
    l0 = ctx.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    ev[0].record()
    for i in range(args.steps):
        kev[i][0].record()
        ctx.propagate_while_valid_dev(stream, B, d_bel, B, d_ctrl, B, d_steps, d_out, B, d_valid)
        kev[i][1].record()
        ctx.rollout_summary_dev(stream, B, d_steps, d_valid, d_sum)
        if world > 1:
            sharding.gather_summary(d_sum)
        ev[i + 1].record()
    barrier()
    launches = ctx.launch_count - l0
    total_ms = ev[0].elapsed_time(ev[-1])
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    clk = clocks.stop()
    t = torch.tensor([total_ms, kern_ms], dtype=torch.float64, device=dev)
    units = torch.tensor([exec_steps], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(units, op=dist.ReduceOp.SUM)
    total_ms, kern_ms = float(t[0]), float(t[1])
    units_all = int(units[0])
    value = units_all * args.steps / (total_ms * 1e-3)

    # ---- e2e: host-pointer C-ABI call, pinned host buffers, H2D + kernel + D2H timed ----
    e2e = None
    if not args.no_e2e:
        n_e2e = max(1, min(args.steps, 3))
        L = ctx.L
        def e2e_call():
            ctx._ck(L.cck_propagate_while_valid(ctx.h, B, h_bel.data_ptr(), B, h_ctrl.data_ptr(), B, h_steps.data_ptr(),
                                                h_out.data_ptr(), B, h_valid.data_ptr(), None, 0, 0))
        e2e_call()
        barrier()
        l1 = ctx.launch_count
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_call()
        torch.cuda.synchronize()
        te = time.perf_counter() - t0
        launches += ctx.launch_count - l1
        tt = torch.tensor([te], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_units = int(np.minimum(h_valid.numpy().astype(np.int64) + 1, T_STEPS).sum())
        eu = torch.tensor([e2e_units], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(eu, op=dist.ReduceOp.SUM)
        e2e = {"value": int(eu[0]) * n_e2e / float(tt[0]), "unit": UNIT,
               "h2d_bytes_per_step": int(B * (Wd * 8 + dim * 8 + 4)), "d2h_bytes_per_step": int(B * (Wd * 8 + 4)),
               "calls": n_e2e, "timing": "host wall clock around the blocking C-ABI call, device synchronised both sides"}
        assert torch.equal(h_valid.to(dev), d_valid), "e2e and device-resident paths disagree"

    # ---- secondary: the realistic variant (obstacle field inside the world, early exits) on the same B ----
    realistic = None
    if args.variant == "all_survive" and not args.no_realistic:
        ctx2 = K.Context(local)
        W.install_synthetic(ctx2, args.model, M_OBS, variant="realistic")
        rb, rc = W.synthetic_beliefs(B, args.model, variant="realistic", offset=rank)
        d_rb, d_rc = torch.from_numpy(rb).to(dev), torch.from_numpy(rc).to(dev)
        del rb, rc
        d_rsum = torch.zeros(3, dtype=torch.int64, device=dev)
        for _ in range(3):
            ctx2.propagate_while_valid_dev(stream, B, d_rb, B, d_rc, B, d_steps, d_out, B, d_valid)
        ctx2.rollout_summary_dev(stream, B, d_steps, d_valid, d_rsum)
        barrier()
        r_units = int(d_rsum[1].item())
        hist = torch.bincount(d_valid.to(torch.int64), minlength=T_STEPS + 1).tolist()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            ctx2.propagate_while_valid_dev(stream, B, d_rb, B, d_rc, B, d_steps, d_out, B, d_valid)
        e1.record()
        barrier()
        launches += ctx2.launch_count - 4
        rt = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        ru = torch.tensor([r_units], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(rt, op=dist.ReduceOp.MAX)
            dist.all_reduce(ru, op=dist.ReduceOp.SUM)
        realistic = {"value": int(ru[0]) * args.steps / (float(rt[0]) * 1e-3), "unit": UNIT, "belief_steps_per_launch": r_units,
                     "kernel_ms": float(rt[0]) / args.steps,
                     "survival_histogram_valid_steps": {"0": hist[0], "1-9": sum(hist[1:10]), "10-24": sum(hist[10:25]),
                                                        "25-49": sum(hist[25:50]), "50": hist[50]},
                     "note": "same B/T/M with the obstacle field inside the world (SURVEY 8d variant i): ~2/3 of the rollouts end early"}
        del d_rb, d_rc
        ctx2.close()

    # ---- secondary: M = 0 trajectory mode (config-1 shape: bounds-only checker, every intermediate belief written
    #      for the plan validator) -- the HBM-bound regime of the path ----
    m0 = None
    if args.variant == "all_survive" and args.model == "linear" and not args.no_realistic:
        Bm = min(B, 1 << 22)
        ctx3 = K.Context(local)
        W.install_synthetic(ctx3, "linear", 0, variant="all_survive")
        d_traj = torch.empty((T_STEPS * Wd, Bm), dtype=torch.float64, device=dev)
        for _ in range(3):
            ctx3.propagate_while_valid_dev(stream, Bm, d_bel, B, d_ctrl, B, d_steps, d_out, B, d_valid, traj=d_traj, traj_ld=Bm, traj_T=T_STEPS)
        barrier()
        assert int((d_valid[:Bm] == T_STEPS).sum().item()) == Bm
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            ctx3.propagate_while_valid_dev(stream, Bm, d_bel, B, d_ctrl, B, d_steps, d_out, B, d_valid, traj=d_traj, traj_ld=Bm, traj_T=T_STEPS)
        e1.record()
        barrier()
        launches += args.steps
        mt = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(mt, op=dist.ReduceOp.MAX)
        m_ms = float(mt[0]) / args.steps
        m_rate = Bm * T_STEPS / (m_ms * 1e-3)
        bytes_per_step = Wd * 8 + (Wd * 8 + dim * 8 + 4 + Wd * 8 + 4) / T_STEPS          # 80 + 184/T
        peaks0 = {}
        try:
            peaks0 = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hp = peaks0.get("hbm_gbs", 6650.0)
        m0 = {"value": m_rate * world, "unit": UNIT, "beliefs_per_gpu": Bm, "kernel_ms": m_ms,
              "roofline": {"bound": "hbm", "achieved": m_rate * bytes_per_step / 1e9, "peak": hp, "unit": "GB/s",
                           "frac": m_rate * bytes_per_step / 1e9 / hp, "bytes_per_belief_step": bytes_per_step},
              "note": "M = 0 (bounds-only checker, config-1 shape), trajectory mode: 80 B written per belief-step"}
        del d_traj
        ctx3.close()

    # ---- PCIe context for e2e: pinned H2D / D2H copy bandwidth of the same buffers ----
    pcie = None
    if not args.no_e2e:
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(); d_bel.copy_(h_bel, non_blocking=True); e1.record(); h_out.copy_(d_out, non_blocking=True); e2.record()
        torch.cuda.synchronize()
        pcie = {"h2d_gbs": h_bel.numel() * 8 / (e0.elapsed_time(e1) * 1e-3) / 1e9, "d2h_gbs": h_out.numel() * 8 / (e1.elapsed_time(e2) * 1e-3) / 1e9}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        fp64_peak = ctx.measure_fp64_peak()
        if args.no_cpu:
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "skipped (--no-cpu)",
                   "halfplane_tests_per_belief_step": 533.99 if args.variant == "all_survive" else 400.0}
        else:
            cpu = cpu_baseline(args)
        hps = cpu["halfplane_tests_per_belief_step"]
        f_prop = 60.0 if dim == 2 else 700.0
        f_alg_per_step = f_prop + 15.0 * hps          # F_alg = N_steps*F_prop + H_total*15
        per_launch_units = exec_steps
        achieved_tf = per_launch_units * f_alg_per_step / (kern_ms * 1e-3) / 1e12
        bytes_per_belief = (Wd * 8 + dim * 8 + 4) + (Wd * 8 + 4)     # 184 B linear, 616 B unicycle
        hbm_gbs = B * bytes_per_belief / (kern_ms * 1e-3) / 1e9
        # dram__bytes_read+write of one launch from the committed ncu --set full capture (per belief, scaled to B)
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            key = f"{args.model}/{args.variant}"
            if key in tj:
                traffic = {"bytes_per_launch": tj[key]["dram_bytes_per_belief"] * B, "algorithmic_bytes_per_launch": B * bytes_per_belief,
                           "source": tj[key]["source"]}
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        # executed-work view: what must stay fp64 for bit-exact results is the Kalman recursion + the covariance sums
        # (linear: 75 FP64-pipe instructions per belief-step on the hot path - 70 of lin_step_diag, which skips the exact
        # zeros of the reference's diagonal Q, R, A_cl, plus Sigma+Lambda and |P01|+|P10|; counted from the ncu source
        # view, profiles/r1o_*.  The dense step of the reference's accumulation order is 95.)  No FMA contraction is
        # allowed, so the attainable rate is the FP64 ISSUE rate = DFMA peak / 2 flop; the chance check itself runs in
        # the fp32 broad phase
        lin_dense = os.environ.get("CCK_LIN_DENSE", "") == "1"
        # unicycle: 316 DMUL + 251 DADD + 32 DFMA + 10 DSETP per warp-step in the ncu source view of k_rollout_uni2 (gpurun r1g capture)
        inst_per_step = (95.0 if lin_dense else 75.0) if dim == 2 else 609.0
        exec_tinst = per_launch_units * inst_per_step / (kern_ms * 1e-3) / 1e12
        inst_peak = fp64_peak / 2.0      # one DFMA = 2 flop = 1 FP64-pipe instruction per lane
        # ---- second half of BASELINE.json's metric: CCK-CBS solve time, short live sample (full distribution:
        #      tools/solve_bench.py -> profiles/r1_solve_times.json) ----
        solve = None
        if world == 1 and not args.no_solve and not args.no_cpu:
            try:
                out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "solve_bench.py"), "--seeds", "5", "--time", "20", "--only", "k=4"],
                                     capture_output=True, text=True, timeout=300)
                solve = {}
                for ln in out.stdout.splitlines():
                    if ln.startswith("config"):
                        lab, js = ln.split(" {", 1)
                        d = json.loads("{" + js)
                        solve[lab] = {k: d[k] for k in ("seeds", "solved", "seconds_median", "seconds_p90", "K", "candidates_median", "cbs_nodes_median")}
                solve["how"] = "K-CBS over the batched GPU BSST (host/cck_plan), instances rebuilt from tests/golden; wall seconds per solve"
            except Exception as e:  # noqa: BLE001
                solve = {"error": str(e)[:200]}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args, B), "beliefs_per_gpu": B, "T": T_STEPS, "M": M_OBS,
                       "l2": "inputs+outputs per launch (%.1f GB) exceed the 126 MB L2; no flush needed" % (B * bytes_per_belief / 1e9),
                       "parallelism": f"beliefs sharded over {world} GPU(s); only a 24-byte result record per rank is all-gathered (NCCL)"},
            "kernel_ms": kern_ms, "belief_steps_per_launch": per_launch_units,
            # The kernel is FP64-issue bound (neither HBM nor tensor): `roofline` is the executed-work view, the one that says
            # how well the silicon is used.  One DADD/DMUL = one flop = one FP64-pipe instruction per lane; FMA contraction is
            # not allowed (bit-exact results), so the attainable rate is the FP64 issue rate = DFMA peak / 2.
            "roofline": {"bound": "fp64", "achieved": exec_tinst, "peak": inst_peak, "unit": "TFLOP/s", "frac": exec_tinst / inst_peak,
                         "traffic": traffic, "fp64_inst_per_belief_step": inst_per_step,
                         "how": "achieved = fp64 flops the kernel must execute for bit-exact results (%.0f DADD/DMUL per belief-step on the hot "
                                "path: Kalman recursion + covariance sums, counted in the ncu source view) x belief-steps per launch / CUDA-event "
                                "kernel time; peak = DFMA-chain microbenchmark run live on this GPU / 2 (no FMA contraction; "
                                "MEASURED_PEAKS.json has no FP64 entry)" % inst_per_step},
            # SURVEY 8(d)'s algorithmic figure: the reference's un-strength-reduced flop count.  The fp32 broad phase never
            # evaluates most half-planes, so this credited rate exceeds the machine peak; it measures the strength reduction.
            "roofline_algorithmic": {"bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved_tf / fp64_peak,
                                     "how": "oracle-counted F_alg (%.0f flop per belief-step = %.0f + 15 x %.2f half-plane tests) x belief-steps per "
                                            "launch / CUDA-event kernel time vs the measured DFMA peak" % (f_alg_per_step, f_prop, hps)},
            "roofline_hbm": {"bound": "hbm", "achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_gbs / hbm_peak,
                             "peak_source": "MEASURED_PEAKS.json (burst copy)" if "hbm_gbs" in peaks else "fallback 6650",
                             "bytes_per_belief": bytes_per_belief},
            "realistic": realistic, "m0_trajectory": m0, "solve_time": solve, "pcie": pcie,
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
