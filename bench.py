#!/usr/bin/env python
"""Benchmark of the patch-graph BA hot path (BASELINE.json metric: BA iterations/s, edges/s per window, % HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c5]

A step is one `fastba.BA(..., iterations=2)` call over one batch of synthetic input:
  c2 (default)  one sliding window of default_cdvo.yaml shape: 22 frames x 96 patches, 37 824 edges, 10 free poses
  c5            64 independent c2-shaped windows (EuRoC intrinsics) solved by one batched call per GPU
Multi-GPU (torchrun, one process per GPU): every rank solves its own independent windows (replicas only, no
collective on the data path; NCCL is used for the barrier and the max-over-ranks of the device time only).

Printed JSON (one line, rank 0): see the task contract.  `value` = BA iterations/s with inputs resident in HBM
(CUDA-graph replay of the call, CUDA-event timed, L2 flushed between steps); `e2e` = the same through the public
python API with pinned HOST buffers (H2D of all inputs and D2H of the updated poses/patches inside the timed
region); `roofline` = algorithmic bytes of the dominant kernel (linearize+Schur) / its event-timed duration against
the measured HBM peak; `cpu_baseline` = the torch/CPU restatement of the reference's ba.py (oracle/ba_torch_port.py)
timed on the host cores.  `--impl reference` times only that CPU path (the reference's own CPU implementation of the
BA is cdvslam/ba.py; it cannot be imported on the GPU box).
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

REPO = os.path.dirname(os.path.abspath(__file__))
for _p in (REPO, os.path.join(REPO, "cdv-slam_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np   # noqa: E402
import torch         # noqa: E402

ITERATIONS = 2
METRIC = "ba_iterations_per_s"


# ------------------------------------------------------------------------------------------------- workloads
def make_workload(name, rank, windows):
    from cdvslam_b200 import synth
    if name == "c2":
        return [synth.config_c2(seed=1234 + rank)]
    if name == "c5":
        return [synth.config_c5_window(rank * windows + s) for s in range(windows)]
    raise SystemExit("unknown workload %s" % name)


def workload_desc(name, probs):
    p = probs[0]
    return {"workload": {"c2": "c2: fastba.BA window, 22 frames x 96 patches, 37824 edges, 10 free poses, 2 iterations",
                         "c5": "c5: %d independent c2-shaped EuRoC windows per GPU, one batched call, 2 iterations"
                         % len(probs)}[name],
            "windows_per_gpu": len(probs), "edges_per_window": int(p.E), "frames": int(p.poses.shape[0]),
            "patches_per_frame": int(p.M), "free_poses": int(p.N), "iterations_per_step": ITERATIONS}


def algorithmic_bytes_linearize(p):
    """SURVEY.md 8(d), assembly (A3) per iteration and window: read E*(3*8+8+8) + Mu*12 + F*28 + 16,
    write 4*(36 N^2 + 6 N Mu + 2 Mu + 6 N)."""
    E, Mu, F, N = p.E, len(np.unique(p.kk)), p.poses.shape[0], p.N
    return E * 40 + Mu * 12 + F * 28 + 16 + 4 * (36 * N * N + 6 * N * Mu + 2 * Mu + 6 * N)


# ------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.thr = [], None, None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thr = threading.Thread(target=self._read, daemon=True)
        self.thr.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------- CPU baseline
def cpu_reference_arm(prob, steps, warmup, threads=None):
    """The reference's CPU path for this hot path (torch ba.py semantics, oracle/ba_torch_port.py) on one c2 window."""
    from oracle import ba_torch_port
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    t = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32)[None]
    poses, patches, intr, target, weight = t(prob.poses), t(prob.patches), t(prob.intrinsics), t(prob.target), t(prob.weight)
    ii, jj, kk = (torch.as_tensor(x) for x in (prob.ii, prob.jj, prob.kk))
    fx, fy, cx, cy = prob.intrinsics[0]
    args = (poses, patches, intr, target, weight, prob.lmbda, ii, jj, kk, prob.t0, [-64, -64, 2 * cx + 64, 2 * cy + 64])
    for _ in range(warmup):
        ba_torch_port.run(args, ITERATIONS, 1.0)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        ba_torch_port.run(args, ITERATIONS, 1.0)
        ts.append(time.perf_counter() - t0)
    total = sum(ts)
    return {"value": ITERATIONS * steps / total, "unit": "BA iterations/s", "cores": threads, "kind": "port",
            "sample": "%d calls of the torch ba.py restatement (2 GN iterations each) on one c2 window, fp32, "
                      "torch.set_num_threads(%d)" % (steps, threads),
            "ms_per_step": 1e3 * total / steps, "edges_per_s": prob.E * ITERATIONS * steps / total}


# ------------------------------------------------------------------------------------------------- GPU arm
class GpuArm:
    def __init__(self, probs, device):
        from cdvslam_b200 import native
        self.native = native
        self.dev = device
        self.B = len(probs)
        p0 = probs[0]
        self.p0 = p0
        f32 = lambda arrs: torch.as_tensor(np.stack([np.asarray(a, np.float32) for a in arrs]))
        i64 = lambda arrs: torch.as_tensor(np.stack([np.asarray(a, np.int64) for a in arrs]))
        # host (pinned) copies in the API layouts, batch-leading
        self.h = {"poses": f32([p.poses for p in probs]), "patches": f32([p.patches for p in probs]),
                  "intrinsics": f32([p.intrinsics for p in probs]), "target": f32([p.target for p in probs]),
                  "weight": f32([p.weight for p in probs]), "lmbda": torch.tensor([p0.lmbda], dtype=torch.float32),
                  "ii": i64([p.ii for p in probs]), "jj": i64([p.jj for p in probs]), "kk": i64([p.kk for p in probs])}
        self.h = {k: v.pin_memory() for k, v in self.h.items()}
        self.d = {k: v.to(device) for k, v in self.h.items()}
        self.pristine = {k: self.d[k].clone() for k in ("poses", "patches")}
        self.out_h = {k: torch.empty_like(self.h[k]).pin_memory() for k in ("poses", "patches")}
        self.flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=device)   # > 126 MB L2

    def restore(self):
        self.d["poses"].copy_(self.pristine["poses"])
        self.d["patches"].copy_(self.pristine["patches"])

    def flush_l2(self):
        if os.environ.get("BENCH_NO_FLUSH") != "1":       # diagnostics only: the reported numbers always flush
            self.flush_buf.zero_()

    def call(self, d=None):
        """The public API call a user makes."""
        from cdvslam_b200 import fastba
        d = d or self.d
        p = self.p0
        if self.B == 1:
            fastba.BA(d["poses"], d["patches"], d["intrinsics"], d["target"], d["weight"], d["lmbda"], d["ii"][0],
                      d["jj"][0], d["kk"][0], p.t0, p.t1, M=p.M, iterations=ITERATIONS, eff_impl=False)
        else:
            fastba.BA_batched(d["poses"], d["patches"], d["intrinsics"], d["target"], d["weight"], d["lmbda"], d["ii"],
                              d["jj"], d["kk"], p.t0, p.t1, M=p.M, iterations=ITERATIONS)

    def capture(self):
        self.call()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        n0 = self.native.lib().pgba_launch_count()
        with torch.cuda.graph(g):
            self.call()
        self.launches_per_step = self.native.lib().pgba_launch_count() - n0
        return g

    def timed_resident(self, graph, steps):
        evs = []
        for _ in range(steps):
            self.restore()
            self.flush_l2()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            graph.replay()
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in evs]

    def timed_e2e(self, steps):
        """Public API with host buffers: H2D of every input from pinned memory, the call, D2H of the result."""
        evs = []
        d = self.d
        for _ in range(steps):
            self.flush_l2()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for k in self.h:
                d[k].copy_(self.h[k], non_blocking=True)
            self.call(d)
            self.out_h["poses"].copy_(d["poses"], non_blocking=True)
            self.out_h["patches"].copy_(d["patches"], non_blocking=True)
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        h2d = sum(v.numel() * v.element_size() for v in self.h.values())
        d2h = sum(v.numel() * v.element_size() for v in self.out_h.values())
        return [a.elapsed_time(b) for a, b in evs], h2d, d2h

    def timed_e2e_host(self, steps, use_graph, arena=False):
        """Host-buffer entry point (fastba.BA_host -> pgba_ba_solve_host): pinned host tensors in, results written back
        into them; H2D of every input and D2H of poses / patches are inside the call.  The state is not reset between
        steps (the in-place host tensors keep being refined; the work per call does not depend on the values)."""
        from cdvslam_b200 import fastba
        p = self.p0
        if arena:       # the nine tensors as views of one pinned allocation (native.host_arena): 2 uploads + 1 download
            hh = self.native.host_arena(self.h["ii"].shape[-1], self.h["poses"].shape[1], self.h["patches"].shape[1],
                                        self.h["patches"].shape[-1])
            for k, v in self.h.items():
                hh[k].view(-1).copy_(v.reshape(-1)) if k not in ("ii", "jj", "kk") else hh[k].copy_(v[0])
            hh = {k: (v[None] if k in ("ii", "jj", "kk") else v) for k, v in hh.items()}
        else:
            hh = {k: v.clone().pin_memory() for k, v in self.h.items()}

        def call():
            fastba.BA_host(hh["poses"], hh["patches"], hh["intrinsics"], hh["target"], hh["weight"], hh["lmbda"],
                           hh["ii"][0], hh["jj"][0], hh["kk"][0], p.t0, p.t1, M=p.M, iterations=ITERATIONS,
                           eff_impl=False, device=self.dev)
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        run = call
        if use_graph:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                call()
            run = g.replay
        evs = []
        for _ in range(steps):
            self.flush_l2()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            run()
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        h2d = sum(hh[k].numel() * hh[k].element_size() for k in hh if k not in ("intrinsics", "_arena")) + \
            (hh["intrinsics"].numel() * 4 if arena else 16)
        d2h = hh["poses"].numel() * 4 + hh["patches"].numel() * 4
        return [a.elapsed_time(b) for a, b in evs], h2d, d2h

    def profiled(self, steps):
        """Per-stage device times of the same launch sequence (events between the kernels, library hook)."""
        native, d, p = self.native, self.d, self.p0
        L = native.lib()
        F, K, P, E = d["poses"].shape[1], d["patches"].shape[1], d["patches"].shape[-1], d["ii"].shape[-1]
        st = native.Strides()
        st.poses, st.patches, st.intrinsics = F * 7, K * 3 * P * P, F * 4
        st.target = st.weight = E * 2
        st.lmbda = 0
        st.ii = st.jj = st.kk = E
        nbytes = ctypes.c_size_t(0)
        native.check(L.pgba_ba_workspace_bytes(E, F, K, p.t0, p.t1, self.B, ctypes.byref(nbytes)), "workspace_bytes")
        ws = native.workspace(nbytes.value, d["poses"].device)
        n = 1 + 3 * ITERATIONS
        buf = (ctypes.c_float * n)()
        acc = np.zeros(n)
        for _ in range(steps):
            self.restore()
            self.flush_l2()
            torch.cuda.synchronize()
            rc = L.pgba_ba_solve_profiled(d["poses"].data_ptr(), d["patches"].data_ptr(), d["intrinsics"].data_ptr(),
                                          d["target"].data_ptr(), d["weight"].data_ptr(), d["lmbda"].data_ptr(),
                                          d["ii"].data_ptr(), d["jj"].data_ptr(), d["kk"].data_ptr(), ctypes.byref(st),
                                          self.B, E, F, K, P, p.t0, p.t1, ITERATIONS, ws.data_ptr(), ws.numel(),
                                          native.stream_ptr(d["poses"].device), buf)
            native.check(rc, "pgba_ba_solve_profiled")
            acc += np.array(list(buf))
        acc /= steps
        names = ["linearize_schur", "solve_retr", "backsub_retr"]
        stages = {"plan": float(acc[0])}
        for k, nme in enumerate(names):
            stages[nme] = float(np.mean([acc[1 + 3 * it + k] for it in range(ITERATIONS)]))
        return stages



def timed_events(fn, steps, before=None):
    """Median of CUDA-event timings (ms) of fn() on the current stream."""
    ts = []
    for _ in range(steps):
        if before is not None:
            before()
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        c.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(c))
    return float(statistics.median(ts))


def extra_corr_c3(dev, peak, flush):
    """BASELINE config c3: altcorr.corr on the c2 graph, 1/4-res 480x640 pyramid (120x160 + 30x40), radius 3.
    Algorithmic bytes per level: E*88 + K*9*C*s + Fj*C*H2*W2*s + E*441*s."""
    from cdvslam_b200 import synth, fastba, altcorr
    p = synth.config_c2()
    d = synth.to_torch(p, dev)
    out = {}
    for C, dt, s in ((24, torch.float16, 2), (128, torch.float16, 2), (128, torch.float32, 4)):
        gmap, pyr = synth.make_fmaps(p, C=C)
        g = torch.as_tensor(gmap, device=dev)[None].to(dt)
        f0 = torch.as_tensor(pyr[0], device=dev)[None].to(dt)
        f1 = torch.as_tensor(pyr[1], device=dev)[None].to(dt)
        coords = fastba.reproject(d["poses"], d["patches"], d["intrinsics"], d["ii"], d["jj"], d["kk"])
        kk, jj = d["kk"], d["jj"]
        fn = lambda: altcorr.corr_pyramid2(g, [f0, f1], coords, kk, jj, 3)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            fn()
        ms = timed_events(gr.replay, 20, before=flush)
        E, K, Fj = p.E, gmap.shape[0], len(np.unique(p.jj))
        alg = sum(E * 88 + K * 9 * C * s + Fj * C * h * w * s + E * 441 * s for (h, w) in ((120, 160), (30, 40))) - E * 88 - K * 9 * C * s
        out["C%d_%s" % (C, str(dt).split(".")[-1])] = {
            "ms": ms, "edges_per_s": E / (ms * 1e-3), "algorithmic_bytes": int(alg),
            "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (ms * 1e-3) / 1e9 / peak}}
        del g, f0, f1
    return {"workload": "c3: fused 2-level altcorr.corr on the c2 graph (37824 edges), radius 3, L2 flushed", **out}


def extra_update_loop(dev, flush, n_updates=12):
    """BASELINE config c3: the per-update hot path of slam.py:316-337, 470-496 on the c2 graph -- reproject -> two-level
    correlation lookup (C = 24 fp16, radius 3) -> synthetic network output (delta ~ N(0,1), weight ~ U(0,1)) -> BA, 2
    iterations -- captured as one CUDA graph and replayed `n_updates` times (the initialisation loop, slam.py:715-716)."""
    from cdvslam_b200 import synth, fastba, altcorr
    p = synth.config_c2()
    d = synth.to_torch(p, dev)
    gmap, pyr = synth.make_fmaps(p, C=24)
    g = torch.as_tensor(gmap, device=dev)[None].half()
    f0 = torch.as_tensor(pyr[0], device=dev)[None].half()
    f1 = torch.as_tensor(pyr[1], device=dev)[None].half()
    gen = torch.Generator(device=dev).manual_seed(1234)
    delta = torch.randn((1, p.E, 2), device=dev, generator=gen)
    weight = torch.rand((1, p.E, 2), device=dev, generator=gen)
    p0, q0 = d["poses"].clone(), d["patches"].clone()
    corr_out = {}

    def update():
        coords = fastba.reproject(d["poses"], d["patches"], d["intrinsics"], d["ii"], d["jj"], d["kk"], clamp_depth=True)
        corr_out["c"] = altcorr.corr_pyramid2(g, [f0, f1], coords, d["kk"], d["jj"], 3)
        target = coords[:, :, :, 1, 1] + delta                         # slam.py:493
        fastba.BA(d["poses"], d["patches"], d["intrinsics"], target, weight, d["lmbda"], d["ii"], d["jj"], d["kk"],
                  p.t0, p.t1, M=p.M, iterations=ITERATIONS, eff_impl=False)

    def reset():
        d["poses"].copy_(p0); d["patches"].copy_(q0); flush()
    for _ in range(2):
        reset(); update()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        update()

    def loop():
        for _ in range(n_updates):
            gr.replay()
    ms = timed_events(loop, 5, before=reset)
    return {"workload": "c3: reproject -> corr (2 levels, C=24 fp16, r=3) -> BA (2 iterations) on the c2 graph, "
                        "%d updates per step, one CUDA graph per update, L2 flushed before each step" % n_updates,
            "ms_per_update": ms / n_updates, "updates_per_s": n_updates / (ms * 1e-3),
            "edges_per_s": p.E * n_updates / (ms * 1e-3)}


def extra_c4(dev, flush):
    """BASELINE config c4: global BA, 1000 frames, 402 624 edges, 999 free poses, eff_impl=True, 2 iterations."""
    from cdvslam_b200 import synth, fastba
    p = synth.config_c4()
    d = synth.to_torch(p, dev)
    p0, q0 = d["poses"].clone(), d["patches"].clone()

    def call():
        fastba.BA(d["poses"], d["patches"], d["intrinsics"], d["target"], d["weight"], d["lmbda"], d["ii"], d["jj"],
                  d["kk"], p.t0, p.t1, M=p.M, iterations=ITERATIONS, eff_impl=True)

    def reset():
        d["poses"].copy_(p0); d["patches"].copy_(q0); flush()
    for _ in range(2):
        reset(); call()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        call()
    ms = timed_events(g.replay, 5, before=reset)
    return {"workload": "c4: global BA, 1000 frames x 96 patches, 402624 edges, 999 free poses, eff_impl, 2 iterations",
            "ms_per_call": ms, "value": ITERATIONS / (ms * 1e-3), "unit": "BA iterations/s",
            "edges_per_s": p.E * ITERATIONS / (ms * 1e-3)}


def hbm_peak():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    try:
        return json.load(open(os.path.join(REPO, "profiles", "traffic.json")))[workload]["linearize_kernel"]
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c5"])
    ap.add_argument("--windows", type=int, default=64, help="windows per GPU for c5")
    ap.add_argument("--no-extra", action="store_true", help="skip the additional batched (c5) measurement")
    ap.add_argument("--cpu-steps", type=int, default=60)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    warmup = max(args.warmup, 3)

    if args.impl == "reference":
        if rank != 0:
            return
        from cdvslam_b200 import synth
        prob = synth.config_c2(seed=1234)
        steps = args.steps if args.steps <= 100 else 100            # bounded sample: ~0.1-0.2 s per call
        res = cpu_reference_arm(prob, steps, min(warmup, 5))
        line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": "BA iterations/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": min(warmup, 5), "ms_per_step": res["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_desc("c2", [prob]), "edges_per_s": res["edges_per_s"],
                "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": res["value"], "unit": "BA iterations/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (the product path has no CPU fallback); use --impl reference for the CPU arm")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        from cdvslam_b200 import shard
        return shard.max_over_ranks(x, dev, dist)

    probs = make_workload(args.workload, rank, args.windows)
    arm = GpuArm(probs, dev)
    for _ in range(warmup):
        arm.restore()
        arm.call()
    graph = arm.capture()
    for _ in range(3):
        arm.restore()
        graph.replay()
    sampler = ClockSampler(local_rank)
    sampler.start()

    barrier()
    ms = arm.timed_resident(graph, args.steps)
    barrier()
    total_ms = max_over_ranks(sum(ms))
    e2e_modes = {}
    e2e_ms, h2d, d2h = arm.timed_e2e(args.steps)
    e2e_modes["device_api_with_torch_copies"] = sum(e2e_ms) / args.steps
    e2e_mode = "device_api_with_torch_copies"
    if arm.B == 1:
        for use_graph, arena in ((False, False), (True, False), (False, True), (True, True)):
            ms_h, h2d_h, d2h_h = arm.timed_e2e_host(args.steps, use_graph, arena)
            name = ("host_api_arena" if arena else "host_api") + ("_graph_replay" if use_graph else "_eager")
            e2e_modes[name] = sum(ms_h) / args.steps
            if sum(ms_h) < sum(e2e_ms):
                e2e_ms, h2d, d2h, e2e_mode = ms_h, h2d_h, d2h_h, name
    barrier()
    e2e_total = max_over_ranks(sum(e2e_ms))
    stages = arm.profiled(min(args.steps, 50))
    clocks = sampler.stop()

    W = len(probs)
    its = ITERATIONS * W * world * args.steps
    value = its / (total_ms * 1e-3)
    peak, peak_src = hbm_peak()
    alg = algorithmic_bytes_linearize(probs[0]) * W
    lin_ms = stages["linearize_schur"]
    achieved = alg / (lin_ms * 1e-3) / 1e9
    step_sum = stages["plan"] + ITERATIONS * sum(v for k, v in stages.items() if k != "plan")
    line = {"metric": METRIC, "value": value, "unit": "BA iterations/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_desc(args.workload, probs), l2="flushed between steps (256 MiB write)",
                           parallelism="replicas only: %d independent window set(s), no collective" % world,
                           timing="CUDA events around a CUDA-graph replay of the public API call, max over ranks"),
            "edges_per_s": probs[0].E * value,
            "e2e": {"value": its / (e2e_total * 1e-3), "unit": "BA iterations/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_total / args.steps, "mode": e2e_mode,
                    "ms_per_step_by_mode": e2e_modes},
            "gpu_launches": int(arm.launches_per_step * args.steps),
            "launches_per_step": int(arm.launches_per_step),
            "roofline": {"bound": "hbm", "kernel": "linearize_kernel (residual+Jacobian+assembly+Schur)",
                         "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                         "frac": achieved / peak, "algorithmic_bytes_per_launch": alg, "ms_per_launch": lin_ms,
                         "share_of_step": ITERATIONS * lin_ms / step_sum, "traffic": ncu_traffic(args.workload)},
            "stages_ms": stages, "clocks": clocks}

    if rank == 0 and world == 1:
        from cdvslam_b200 import synth
        line["cpu_baseline"] = {k: v for k, v in cpu_reference_arm(synth.config_c2(seed=1234), args.cpu_steps, 3).items()
                                if k in ("value", "unit", "cores", "kind", "sample")}
        if args.workload == "c2" and not args.no_extra:
            del arm, graph
            torch.cuda.empty_cache()
            p5 = make_workload("c5", 0, args.windows)
            arm5 = GpuArm(p5, dev)
            for _ in range(3):
                arm5.restore(); arm5.call()
            g5 = arm5.capture()
            n5 = max(10, min(args.steps // 4, 50))
            ms5 = arm5.timed_resident(g5, n5)
            st5 = arm5.profiled(min(n5, 20))
            alg5 = algorithmic_bytes_linearize(p5[0]) * len(p5)
            ach5 = alg5 / (st5["linearize_schur"] * 1e-3) / 1e9
            line["batched_c5"] = {"windows": len(p5), "steps": n5, "ms_per_step": sum(ms5) / n5,
                                  "value": ITERATIONS * len(p5) * n5 / (sum(ms5) * 1e-3), "unit": "BA iterations/s",
                                  "edges_per_s": p5[0].E * ITERATIONS * len(p5) * n5 / (sum(ms5) * 1e-3),
                                  "stages_ms": st5,
                                  "roofline": {"bound": "hbm", "kernel": "linearize_kernel", "achieved": ach5,
                                               "peak": peak, "unit": "GB/s", "frac": ach5 / peak,
                                               "algorithmic_bytes_per_launch": alg5,
                                               "ms_per_launch": st5["linearize_schur"],
                                               "traffic": ncu_traffic("c5")}}
    if rank == 0 and world == 1 and args.workload == "c2" and not args.no_extra:
        flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        try:
            line["corr_c3"] = extra_corr_c3(dev, peak, flush_buf.zero_)
            line["update_loop_c3"] = extra_update_loop(dev, flush_buf.zero_)
            line["global_c4"] = extra_c4(dev, flush_buf.zero_)
        except Exception as e:            # extras must never cost the headline line
            line["extras_error"] = repr(e)[:200]
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
