#!/usr/bin/env python3
"""bench.py -- NeRF train-step throughput (forward + backward + optimiser) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (host cores)

Workload (BASELINE.json configs[1], "C2"): the reference MLP 33->30->30->4 (5-band positional
encoding of xyz), 64 stratified samples per ray, 4096-ray batches, synthetic rays (data/lego is not
shipped).  A step = MLP forward, compositing, SSE loss, full reverse-mode gradient (d_ws, d_bs),
[N>1: NCCL all-reduce of the gradient buffer], Adam update -- per GPU on its own 4096-ray shard
(weak scaling).  `value` = samples/s over all GPUs with inputs resident in HBM; `e2e` = the same
step through the host-pointer C-ABI call (features, dists, targets copied from pinned host memory
every step, loss and gradients copied back).  One JSON line on stdout (rank 0).
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

METRIC = "nerf_train_samples_per_s"
UNIT = "samples/s"
WORKLOADS = {
    # name: (E bands, width, layers, rays per batch, samples per ray)
    "c2": dict(E=5, width=30, layers=3, R=4096, S=64),
    "c5": dict(E=10, width=256, layers=9, R=4096, S=192),
    # BASELINE config 1: the 2-D image fit of fit_img.py (mlp_fit: 22 -> 16 -> 16 -> 3, sigmoid head, SSE loss, SGD
    # 1e-4, fit_img.py:379-421,512-513); one step = one pass over a synthetic 256 x 256 image
    "c1": dict(fit=True, E=5, width=16, layers=3, R=65536, S=1),
}


def workload_config(name, n_gpus, extra=None):
    w = WORKLOADS[name]
    if w.get("fit"):
        c_in = 2 + 4 * w["E"]
        cfg = {"workload": "mlp-fit %s: MLP %d->%s->3 (fit_img.py), %d pixels per GPU per step" % (
            name.upper(), c_in, "x".join([str(w["width"])] * (w["layers"] - 1)), w["R"]),
            "pixels_per_gpu": w["R"], "global_pixels": w["R"] * n_gpus,
            "sharding": "pixels across %d GPU(s), gradient all-reduce" % n_gpus if n_gpus > 1 else "single GPU"}
        cfg.update(extra or {})
        return cfg
    c_in = 3 + 6 * w["E"]
    cfg = {"workload": "nerf-train %s: MLP %d->%s->4, %d rays x %d samples per GPU per step" % (
        name.upper(), c_in, "x".join([str(w["width"])] * (w["layers"] - 1)), w["R"], w["S"]),
        "rays_per_gpu": w["R"], "samples_per_ray": w["S"], "global_rays": w["R"] * n_gpus,
        "sharding": "rays across %d GPU(s), gradient all-reduce" % n_gpus if n_gpus > 1 else "single GPU"}
    cfg.update(extra or {})
    return cfg


def flops_per_sample(dims):
    return 3 * 2 * sum(dims[l] * dims[l + 1] for l in range(len(dims) - 1))  # F_train


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation on the host cores
# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import cpu_bench
    import multiprocessing as mp
    w = WORKLOADS[args.workload]
    cores = cpu_bench.host_cores()
    if args.workload == "c5":
        # paper-size network: the reference program rebuilt with larger static tapes (oracle/_ref/nerf_big.so),
        # one ray of 192 samples per call, 1.9 GB of stack per process -> at most 8 processes
        procs = min(cores, 8)
        if cpu_bench.run_big(1, 0) is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/nerf_big.so was not built"}))
            return 0
        tot_s, tot_n, steps = 0.0, 0, min(args.steps, 3)
        for _ in range(min(args.warmup, 1)):
            cpu_bench.run_big(procs, 1, S=w["S"])
        for _ in range(steps):
            r = cpu_bench.run_big(procs, 1, S=w["S"])
            tot_s += r["seconds"]
            tot_n += r["samples"]
        value = tot_n / tot_s
        sample = ("each step: %d processes x 1 call pair (forward + grad) x 1 ray x %d samples of the C5 network through "
                  "oracle/_ref/nerf_big.so; %d steps timed (each ~4 s), wall time of the slowest process" % (procs, w["S"], steps))
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * tot_s / steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": workload_config(args.workload, args.gpus),
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "reference", "sample": sample},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0
    chunks = 24   # per core per step: 24 chunks x 256 samples, ~0.25 s of C time
    fit = bool(w.get("fit"))
    with mp.get_context("fork").Pool(cores) as pool:
        run = (lambda n: cpu_bench.run_fit(cores, n, pool=pool)) if fit else (lambda n: cpu_bench.run(cores, n, S=w["S"], pool=pool))
        for _ in range(args.warmup):
            run(2)
        tot_s, tot_n = 0.0, 0
        for _ in range(args.steps):
            r = run(chunks)
            tot_s += r["seconds"]
            tot_n += r["samples"]
    value = tot_n / tot_s
    if fit:
        sample = ("each step: %d cores x %d chunks x 256 pixels of the C1 workload through mlp_fit / grad_mlp_fit, forward call + grad "
                  "call per chunk as fit_img.py does (zero-copy marshalling included, microseconds per call)" % (cores, chunks))
    else:
        sample = ("each step: %d cores x %d chunks x 256 samples (4 rays x 64) of the C2 workload, forward call + grad call "
                  "per chunk as train_nerf.py does, time inside the C calls only" % (cores, chunks))
    line = {"impl": "reference", "metric": "mlp_fit_train_samples_per_s" if fit else METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args.workload, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": r["kind"], "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def _dbg(msg):
    if os.environ.get("BENCH_DEBUG"):
        sys.stderr.write("[bench rank %s] %s\n" % (os.environ.get("RANK", "0"), msg))
        sys.stderr.flush()


class Env:
    """One process per GPU: device, stream, process group, measured peaks."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the CUDA library has no CPU fallback "
                             "(use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        self.device = torch.device("cuda", self.local)
        self.numa = None
        if self.world > 1 and not os.environ.get("LNB_BENCH_NO_AFFINITY"):
            # several GPUs on a multi-socket host: run this rank (and so first-touch its pinned host buffers) on the cores
            # next to its GPU, or the end-to-end legs measure the socket interconnect instead of PCIe
            try:
                import pynvml
                pynvml.nvmlInit()
                pr = torch.cuda.get_device_properties(self.local)
                bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
                pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode()))
                self.numa = "rank bound to its GPU's NUMA-local cores (NVML ideal CPU affinity, %d cores)" % len(os.sched_getaffinity(0))
            except Exception as e:      # no NVML, no permission: measure as launched
                self.numa = "unchanged (%s)" % type(e).__name__
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.device)
        self.stream = torch.cuda.current_stream(self.device)
        try:
            self.peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            self.peaks = {}

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.device)

    def max_over_ranks(self, x):
        t = self.torch.tensor([float(x)], device=self.device, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn):
        """Device time of fn() in ms: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks."""
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record(self.stream)
        fn()
        e1.record(self.stream)
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1))


class TrainRun:
    """One workload on this rank: synthetic batches (a pool larger than L2, rotated every step), a device-resident trainer,
    and the launch machinery (CUDA graphs of consecutive steps; NCCL mode: gradient graph, eager all-reduce, optimiser graph)."""
    CHUNK = 240

    def __init__(self, env, wname, path, inp, collective="peer", eager=False, rays_per_gpu=None, host_copies=True):
        import numpy as np
        from loma_nerf_b200 import api, synthetic
        torch = env.torch
        self.env, self.wname, self.path = env, wname, path
        w = dict(WORKLOADS[wname])
        if rays_per_gpu is not None:
            w["R"] = int(rays_per_gpu)
        self.w = w
        E, R, S = w["E"], w["R"], w["S"]
        self.fit = fit = bool(w.get("fit"))
        self.c_in = c_in = 2 + 4 * E if fit else 3 + 6 * E
        self.dims = dims = synthetic.mlp_dims(c_in, w["width"], w["layers"], 3 if fit else 4)
        self.N = N = R * S
        self.use_rays = use_rays = inp == "rays" and not fit
        self.ctx = ctx = api.Context(env.local)
        ctx.set_stream(env.stream)
        device = env.device
        rng = np.random.default_rng(215 + 1000 * env.rank)          # the reference seeds numpy with 215
        self.rng = rng
        if use_rays:
            self.bytes_per_batch = R * 3 * 8 * 2 + N * 8 + R * 12
        else:
            self.bytes_per_batch = N * c_in * 4 + (0 if fit else N * 4) + R * 12
        self.n_pool = n_pool = max(2, int(np.ceil(160e6 / self.bytes_per_batch)) + 1)
        self.batches, self.host_batches = [], []
        for b in range(n_pool if fit else 0):
            # pixel coordinates in [0,1)^2, 5-band positional encoding (pos_encoding.py:4-36), a smooth synthetic image
            xy = rng.uniform(0, 1, (N, 2))
            target = (0.5 + 0.4 * np.sin(6 * xy[:, :1] + np.array([0.0, 1.0, 2.0])) * np.cos(4 * xy[:, 1:])).clip(0, 1).astype(np.float32)
            X = ctx.pos_encoding(torch.as_tensor(xy, device=device), E)
            tgd = torch.as_tensor(target, device=device)
            self.batches.append(dict(X=X, target=tgd, path=path))
            if b < 2 and host_copies:
                self.host_batches.append(dict(features=dict(X=X.cpu().pin_memory(), target=tgd.cpu().pin_memory(), path=path)))
        for b in range(0 if fit else n_pool):
            o, d = synthetic.random_rays(rng, R)
            t = synthetic.stratified_t(rng, R, S)
            target = rng.uniform(0, 1, (R, 3)).astype(np.float32)
            od, dd, td = (torch.as_tensor(v, device=device) for v in (o, d, t))
            tgd = torch.as_tensor(target, device=device)
            if use_rays:
                self.batches.append(dict(rays=(od, dd, td), pe_bands=E, target=tgd, path=path))
            else:
                X, dists = ctx.sample_encode(od, dd, td, E)
                self.batches.append(dict(X=X, dists=dists, target=tgd, path=path))
            if b < 2 and host_copies:   # host copies (pinned) for the end-to-end legs
                X, dists = ctx.sample_encode(od, dd, td, E)
                cam_pose = np.eye(4)
                cam_pose[:3, 3], cam_pose[:3, :3] = synthetic.camera(rng)
                pix = torch.as_tensor(rng.integers(0, 800 * 800, R).astype(np.int32)).pin_memory()
                K = np.array([[synthetic.FOCAL, 0, 0.5], [0, synthetic.FOCAL, 0.5], [0, 0, 1.0]])
                self.host_batches.append(dict(
                    features=dict(X=X.cpu().pin_memory(), dists=dists.cpu().pin_memory(), target=tgd.cpu().pin_memory(), path=path),
                    rays=dict(rays=tuple(torch.as_tensor(v).pin_memory() for v in (o, d, t)), pe_bands=E,
                              target=tgd.cpu().pin_memory(), path=path),
                    camera=dict(camera=api.make_camera(cam_pose, K, 800, 800, near=synthetic.NEAR, far=synthetic.FAR, pixels=pix,
                                                       stratified=True, seed=215 + b),
                                S=S, pe_bands=E, target=tgd.cpu().pin_memory(), path=path)))
                del X, dists
        ws_np, bs_np = synthetic.init_mlp(np.random.default_rng(216), dims)   # same weights on every rank
        self.ws_np, self.bs_np = ws_np, bs_np
        self.nP = ws_np.size + bs_np.size
        if fit:
            self.trainer = api.Trainer(ctx, dims, ws_np, bs_np, head=api.L.HEAD_SIGMOID, optimizer="sgd", lr=1e-4)
        else:
            self.trainer = api.Trainer(ctx, dims, ws_np, bs_np, optimizer="adam", lr=5e-4)
        self.grads = self.trainer.grad_buffer()                           # [d_ws | d_bs | loss] on the device
        self.collective = "none"
        if env.world > 1:
            self.collective = "nccl"
            if collective == "peer" and path == "tc":
                try:
                    self.trainer.enable_peer_allreduce()
                    # a step the fused kernel cannot take (wide MLP, S > 128) has no in-kernel exchange: find out now
                    self.trainer.step(**self.batches[0])
                    self.collective = "peer"
                except Exception as e:
                    _dbg("peer all-reduce unavailable (%s); using NCCL" % str(e)[:120])
                    self.trainer.close()
                    self.trainer = api.Trainer(ctx, dims, ws_np, bs_np, optimizer="adam", lr=5e-4) if not fit else \
                        api.Trainer(ctx, dims, ws_np, bs_np, head=api.L.HEAD_SIGMOID, optimizer="sgd", lr=1e-4)
                    self.grads = self.trainer.grad_buffer()
        self.fused_step = env.world == 1 or self.collective == "peer"
        self.graphs, self.grad_graphs, self.apply_graph = {}, [], None
        self.launch_mode = "eager" if eager else "cuda-graph of consecutive steps"
        self.counter = 0
        if not eager:
            try:
                for b in range(2):
                    self.step_body(b)
                torch.cuda.synchronize(device)
                if self.fused_step:
                    self.graphs[1] = self._capture(lambda: self.step_body(0))      # proves capture works before the big ones
                else:
                    for b in range(n_pool):
                        self.grad_graphs.append(self._capture(lambda b=b: self.trainer.grad(**self.batches[b])))
                    self.apply_graph = self._capture(self.trainer.apply)
                    self.launch_mode = "cuda-graph (gradient) + NCCL all-reduce + cuda-graph (optimiser) per step"
            except Exception as e:  # report, never hide: fall back to eager launches
                self.graphs, self.grad_graphs, self.apply_graph = {}, [], None
                self.launch_mode = "eager (graph capture failed: %s)" % str(e)[:80]
                ctx.set_stream(env.stream)
                torch.cuda.synchronize(device)
        _dbg("%s: launch mode %s" % (wname, self.launch_mode))

    def close(self):
        self.graphs, self.grad_graphs, self.apply_graph = {}, [], None
        self.trainer.close()
        self.batches, self.host_batches = [], []
        self.env.torch.cuda.synchronize(self.env.device)
        self.ctx.close()
        self.env.torch.cuda.empty_cache()

    def step_body(self, b):
        if self.fused_step:
            self.trainer.step(**self.batches[b])                      # fused fwd+bwd kernel, then reduce + optimiser
        else:
            self.trainer.grad(**self.batches[b])
            self.env.dist.all_reduce(self.grads)                      # NCCL sum over NVLink: gradients + loss
            self.trainer.apply()

    def _capture(self, fn):
        torch = self.env.torch
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.ctx.set_stream(torch.cuda.current_stream(self.env.device))
            fn()
        self.ctx.set_stream(self.env.stream)
        return g

    def eager(self):
        return self.launch_mode.startswith("eager")

    def chunk_plan(self, T):
        """T consecutive steps as graph chunks; every chunk is a multiple of the pool size (so each replay walks the whole
        rotation from batch 0 and never revisits a batch while it can still be in L2) except a last remainder."""
        if T <= self.n_pool:
            return [T]
        C = self.n_pool * max(1, min(self.CHUNK, T) // self.n_pool)
        plan = [C] * (T // C)
        if T % C:
            plan.append(T % C)
        return plan

    def prepare(self, T):
        """Capture the graphs a run of T steps needs (outside any timed region: capture launches nothing)."""
        if self.eager() or not self.fused_step:
            return
        for c in set(self.chunk_plan(T)):
            if c not in self.graphs:
                self.graphs[c] = self._capture(lambda c=c: [self.step_body(j % self.n_pool) for j in range(c)])

    def run(self, T, marks=None):
        """T consecutive train steps.  marks (optional list) receives (event, steps so far) after every chunk."""
        torch = self.env.torch

        def mark(done):
            if marks is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record(self.env.stream)
                marks.append((e, done))
        done = 0
        if self.eager():
            every = max(1, T // 16)
            for i in range(T):
                self.step_body(self.counter % self.n_pool)
                self.counter += 1
                if (i + 1) % every == 0 or i + 1 == T:
                    mark(i + 1)
        elif self.fused_step:
            for c in self.chunk_plan(T):
                self.graphs[c].replay()
                done += c
                mark(done)
        else:
            every = max(1, T // 16)
            for i in range(T):
                self.grad_graphs[self.counter % self.n_pool].replay()
                self.env.dist.all_reduce(self.grads)
                self.apply_graph.replay()
                self.counter += 1
                if (i + 1) % every == 0 or i + 1 == T:
                    mark(i + 1)

    def measure(self, steps, warm, min_ms=50.0):
        """W warm-up steps, then reps x `steps` steps in ONE timed region of at least min_ms (the captured graphs replayed),
        device-timed, max over ranks.  Per-chunk events inside the region give the per-step median and spread."""
        env, torch = self.env, self.env.torch
        warm = max(warm, 3)
        self.prepare(warm)
        self.run(warm)
        env.barrier()
        probe = max(3, min(steps, 12))
        self.prepare(probe)
        est = env.timed(lambda: self.run(probe)) / probe
        reps = max(1, int(-(-(1.2 * min_ms) // max(est * steps, 1e-6))))   # 20 % margin: the probe runs cold
        T = reps * steps
        self.prepare(T)
        env.barrier()
        l0 = self.ctx.launches
        self.step_body(0)                               # count this library's kernels in one step (eagerly)
        per_step = self.ctx.launches - l0
        env.barrier()
        marks = []
        e0 = torch.cuda.Event(enable_timing=True)
        env.barrier()
        e0.record(env.stream)
        self.run(T, marks)
        env.barrier()
        total_ms = env.max_over_ranks(e0.elapsed_time(marks[-1][0]))
        per = []
        prev_e, prev_n = e0, 0
        for e, n in marks:
            if n > prev_n:
                per.append(prev_e.elapsed_time(e) / (n - prev_n))
            prev_e, prev_n = e, n
        per.sort()
        if self.collective == "peer":
            self.trainer.check_comm()
        return dict(ms_total=total_ms, steps_timed=T, reps=reps, ms_per_step=total_ms / T, launches_per_step=per_step,
                    value=env.world * self.N * T / (total_ms * 1e-3), warm=warm,
                    per_step_ms={"median": per[len(per) // 2], "min": per[0], "max": per[-1], "chunks": len(per)},
                    loss_last=float(self.grads[self.nP].item()))

    # ---- end to end: host buffers in, loss out, copies inside the timed region
    def e2e_sync(self, kind, n):
        env = self.env
        hb = [self.host_batches[0][kind], self.host_batches[1][kind]]
        for i in range(3):
            self.trainer.step_host(**hb[i % 2])
        env.barrier()
        t0 = time.perf_counter()
        for i in range(n):
            self.trainer.step_host(**hb[i % 2])
        env.torch.cuda.synchronize(env.device)
        sec = env.max_over_ranks(time.perf_counter() - t0)
        return env.world * self.N * n / sec

    def e2e_pipelined(self, kind, min_s=0.05, max_n=4000):
        """lnb_trainer_submit_host per step (H2D of batch i+1 under step i), one lnb_trainer_wait at the end: every step's
        batch crosses the bus and every step's loss comes back inside the timed region."""
        env = self.env
        hb = [self.trainer.prepare(**self.host_batches[0][kind]), self.trainer.prepare(**self.host_batches[1][kind])]   # marshalled once
        for i in range(4):
            self.trainer.submit_host(hb[i % 2])
        self.trainer.wait()
        env.barrier()
        t0 = time.perf_counter()
        for i in range(16):
            self.trainer.submit_host(hb[i % 2])
        self.trainer.wait()
        est = (time.perf_counter() - t0) / 16
        n = int(min(max_n, max(20, -(-min_s // est))))
        if env.world > 1:   # every rank must submit the same number of steps (the peer exchange runs in lockstep)
            n = int(env.max_over_ranks(n))
        env.barrier()
        t0 = time.perf_counter()
        for i in range(n):
            self.trainer.submit_host(hb[i % 2])
        losses = self.trainer.wait()
        sec = env.max_over_ranks(time.perf_counter() - t0)
        assert len(losses) == min(n, 4096) and all(l == l for l in losses), "e2e: a loss did not come back"
        return env.world * self.N * n / sec, n


def roofline_record(run, prof, m, peaks):
    """The dominant kernel of the step against the roof that bounds it (SURVEY.md 8d's per-unit bytes / FLOPs x the units
    one launch processes / that kernel's mean launch duration, CUDA events around every launch of it over `steps` steps)."""
    dims, N, R, c_in, fit, use_rays = run.dims, run.N, run.w["R"], run.c_in, run.fit, run.use_rays
    fl = flops_per_sample(dims)
    step_s = m["ms_per_step"] * 1e-3
    tpeak_s = peaks.get("bf16_tflops_sustained", 1389.4)
    alg_bytes = (N * 8 + R * 60) if use_rays else (N * (c_in * 4 + (0 if fit else 4)) + R * 12)
    pad = lambda v: (v + 63) // 64 * 64  # noqa: E731
    Lw = len(dims) - 1

    def ncu_traffic(csv_name, kernel, col_scale=1.0):
        p = os.path.join(ROOT, "profiles", csv_name)
        if run.wname != "c5" or not os.path.exists(p):
            return None, None
        try:
            import csv as _csv
            rows = [r for r in _csv.reader(open(p))][1:]
            g = [float(r[4]) + float(r[5]) for r in rows if r[2] == kernel]
            if g:
                return sum(g) / len(g) * col_scale, "profiles/%s (mean over the step's %d launches of %s)" % (csv_name, len(g), kernel)
        except Exception:
            pass
        return None, None

    if prof and prof["kernel"] == "chain_tc_kernel":
        # wide MLP, chained layers (wide_tc.cu chain_tc_kernel): two launches per step, the forward chain (all L layers) and
        # the adjoint chain (L-1 layers).  Activations are re-read from L2, so the launch is bound by the tensor pipe and
        # what feeds it (shared-memory bandwidth), not by HBM.
        sec = prof["ms_per_launch"] * 1e-3
        f_fwd = 2.0 * sum(dims[l] * dims[l + 1] for l in range(len(dims) - 1))
        f_adj = 2.0 * sum(dims[l] * dims[l + 1] for l in range(1, len(dims) - 1))
        flops_launch = N * (f_fwd + f_adj) / 2
        ach = flops_launch / sec / 1e12
        traffic, traffic_src = ncu_traffic("r02_wide_chain_ncu_full_per_launch.csv", "chain_tc_kernel")
        b_fwd = N * (pad(dims[0]) * 2 + sum(pad(dims[l + 1]) * 2 + pad(dims[l + 1]) // 8 for l in range(Lw - 1)) + 16)
        b_adj = N * (pad(dims[Lw]) * 2 + sum(pad(dims[l]) * 2 + pad(dims[l]) // 8 for l in range(1, Lw)))
        return {"bound": "tensor", "achieved": ach, "peak": tpeak_s, "unit": "TFLOP/s", "frac": ach / tpeak_s,
                "traffic": traffic, "traffic_source": traffic_src, "kernel": prof["kernel"], "us_per_launch": sec * 1e6,
                "launches_timed": prof["launches"], "launches_per_step": 2,
                "algorithmic_flops_per_launch": flops_launch, "algorithmic_bytes_per_launch": (b_fwd + b_adj) / 2,
                "hbm_gbs": (b_fwd + b_adj) / 2 / sec / 1e9,
                "step_tflops": N * fl / step_s / 1e12, "step_tensor_frac": N * fl / step_s / 1e12 / tpeak_s,
                "note": "the rest of the step is the HBM-bound weight-gradient kernels (dw_tc_kernel, one launch per layer); "
                        "step_tensor_frac is SURVEY 8d's F_train x samples / step time against the sustained bf16 peak",
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside the step)" if peaks else "fallback 1389 TFLOP/s"}
    if prof and prof["kernel"] == "gemm_tc_kernel":
        sec = prof["ms_per_launch"] * 1e-3
        gemms = [(pad(dims[l]), pad(dims[l + 1]), N * (pad(dims[l]) * 2 + pad(dims[l + 1]) * 2 + pad(dims[l + 1]) // 8)) for l in range(Lw - 1)]
        gemms.append((pad(dims[Lw - 1]), 16, N * (pad(dims[Lw - 1]) * 2 + 16)))
        gemms += [(pad(dims[l + 1]), pad(dims[l]), N * (pad(dims[l + 1]) * 2 + pad(dims[l]) * 2 + pad(dims[l]) // 8)) for l in range(Lw - 1, 0, -1)]
        bytes_launch = sum(g[2] for g in gemms) / len(gemms)
        flops_launch = sum(2.0 * N * g[0] * g[1] for g in gemms) / len(gemms)
        peak = peaks.get("hbm_gbs", 6650.0)
        ach = bytes_launch / sec / 1e9
        traffic, traffic_src = ncu_traffic("r01_wide_c5_ncu_full_per_launch.csv", "gemm_tc_kernel", 1e6)
        return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "traffic_source": traffic_src, "kernel": prof["kernel"], "us_per_launch": sec * 1e6,
                "launches_timed": prof["launches"], "launches_per_step": len(gemms),
                "algorithmic_bytes_per_launch": bytes_launch, "algorithmic_tflops": flops_launch / sec / 1e12,
                "step_tflops": N * fl / step_s / 1e12, "step_tensor_frac": N * fl / step_s / 1e12 / tpeak_s, "tensor_peak": tpeak_s,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy), bf16_tflops_sustained" if peaks else "fallback 6650 GB/s, 1389 TFLOP/s"}
    if prof and prof["kernel"].startswith("fused_f32"):
        # the exact fp32 CUDA-core kernel: by SURVEY 8d's bytes it is an HBM-bound step like the tensor-core one; what
        # actually bounds it is the FFMA pipe (F_train per sample on 148 SMs x 128 lanes x 2 FLOP x the SM clock)
        sec = prof["ms_per_launch"] * 1e-3
        peak = peaks.get("hbm_gbs", 6650.0)
        ach = alg_bytes / sec / 1e9
        ffma_peak = 148 * 128 * 2 * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e12
        return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                "kernel": prof["kernel"], "us_per_launch": sec * 1e6, "launches_timed": prof["launches"],
                "algorithmic_bytes_per_launch": alg_bytes, "ffma_tflops": N * fl / sec / 1e12, "ffma_peak_tflops": ffma_peak,
                "ffma_frac": N * fl / sec / 1e12 / ffma_peak,
                "note": "fp32 CUDA cores: ffma_frac = F_train x samples / launch time against 148 SMs x 128 FMA lanes at the maximum SM clock",
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s"}
    if prof:
        sec = prof["ms_per_launch"] * 1e-3
        if use_rays:
            peak = peaks.get("bf16_tflops", 1590.0)
            ach = N * fl / sec / 1e12
            return {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "traffic": None, "kernel": prof["kernel"], "us_per_launch": sec * 1e6,
                    "launches_timed": prof["launches"], "algorithmic_flops_per_launch": N * fl,
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst)" if peaks else "fallback 1590 TFLOP/s"}
        # the fused kernel reads pre-encoded features: HBM-bound by SURVEY.md 8d's per-unit bytes
        peak = peaks.get("hbm_gbs", 6650.0)
        ach = alg_bytes / sec / 1e9
        traffic, traffic_src = None, None
        for prof_csv in ("r02_mg_feat_ncu_full_summary.csv", "r01_fused_tc_features_ncu_full_summary.csv"):
            p = os.path.join(ROOT, "profiles", prof_csv)
            if run.wname == "c2" and traffic is None and os.path.exists(p):
                # DRAM bytes of one launch of this kernel on this workload, from the committed ncu --set full capture
                try:
                    vals = {}
                    for ln in open(p):
                        parts = [c.strip('"') for c in ln.strip().split('","')]
                        if len(parts) == 3 and parts[0].lstrip('"') in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(parts[1], None)
                            if mult:
                                vals[parts[0].lstrip('"')] = float(parts[2].rstrip('"')) * mult
                    if len(vals) == 2:
                        traffic, traffic_src = sum(vals.values()), "profiles/" + prof_csv
                except Exception:
                    pass
        return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "traffic_source": traffic_src, "kernel": prof["kernel"], "us_per_launch": sec * 1e6,
                "launches_timed": prof["launches"], "algorithmic_bytes_per_launch": alg_bytes,
                "algorithmic_tflops": N * fl / sec / 1e12, "tensor_frac": N * fl / sec / 1e12 / peaks.get("bf16_tflops", 1590.0),
                "note": "us_per_launch: CUDA events around each EAGER launch of the kernel (the events see ~3 us of launch gap that the graph "
                        "of consecutive steps hides under the previous kernel: there the whole step, reduce + Adam kernel included, is "
                        "ms_per_step); frac is computed from the eager figure",
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s"}
    # no single dominant kernel (layerwise kernels): report the whole step against HBM
    peak = peaks.get("hbm_gbs", 6650.0)
    ach = alg_bytes / step_s / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": None, "kernel": "whole step (layerwise kernels)",
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"}


def cpu_baseline_record(wname, S):
    """The reference's CPU implementation on this box's host cores, a bounded sample of the workload (rank 0, N = 1 only)."""
    from oracle import cpu_bench
    if wname == "c5":
        procs = min(cpu_bench.host_cores(), 8)
        r = cpu_bench.run_big(procs, 2, S=S)
        if r is None:
            return None
        return {"value": r["samples"] / r["seconds"], "unit": UNIT, "cores": procs, "kind": "reference",
                "sample": "%d processes x 2 call pairs x 1 ray x %d samples of this network through the reference program rebuilt "
                          "with larger tapes (oracle/_ref/nerf_big.so, 1.9 GB of stack each), forward + grad call, wall time of the "
                          "slowest process" % (procs, S)}
    cores = cpu_bench.host_cores()
    if WORKLOADS[wname].get("fit"):
        r = cpu_bench.run_fit(cores, 40)
        return {"value": r["samples"] / r["seconds"], "unit": UNIT, "cores": cores, "kind": r["kind"],
                "sample": "%d cores x 40 chunks x 256 pixels, mlp_fit + grad_mlp_fit per chunk (fit_img.py:423-532)" % cores}
    r = cpu_bench.run(cores, 40, S=S)
    r1 = cpu_bench.run(1, 40, S=S)
    return {"value": r["samples"] / r["seconds"], "unit": UNIT, "cores": cores, "kind": r["kind"],
            "sample": "%d cores x 40 chunks x 256 samples (4 rays x 64), forward + grad call per chunk, time inside the C calls only" % cores,
            "value_1core": r1["samples"] / r1["seconds"],
            "note": "the reference's own end-to-end rate is ~1e3 samples/s: its Python marshalling (mlp_utils.py:33-164) costs "
                    "~0.1 s per 120-sample chunk (SURVEY.md 6); excluded here"}


def dtype_of(path):
    return "f32" if path != "tc" else "bf16 operands, f32 accumulate (tcgen05)"


def sub_record(env, wname, path, inp, steps, collective, with_cpu, rays_per_gpu=None, min_ms=50.0):
    """A compact record of another BASELINE config (same timing rules as the main line)."""
    run = TrainRun(env, wname, path, inp, collective=collective, host_copies=False, rays_per_gpu=rays_per_gpu)
    try:
        m = run.measure(steps, 3, min_ms=min_ms)
        prof = run.ctx.profile_dominant(lambda: [run.step_body(i % run.n_pool) for i in range(min(steps, 10))])
        rf = roofline_record(run, prof, m, env.peaks)
        keep = ("bound", "kernel", "frac", "achieved", "peak", "unit", "us_per_launch", "launches_per_step", "step_tensor_frac",
                "step_tflops", "ffma_frac", "ffma_tflops", "tensor_frac")
        rec = {"workload": workload_config(wname, env.world)["workload"] if rays_per_gpu is None else
               "nerf-train %s shape, %d rays x %d samples per GPU per step" % (wname.upper(), run.w["R"], run.w["S"]),
               "metric": "mlp_fit_train_samples_per_s" if run.fit else METRIC, "value": m["value"], "unit": UNIT,
               "ms_per_step": m["ms_per_step"], "per_step_ms": m["per_step_ms"], "steps_timed": m["steps_timed"],
               "timed_region_ms": m["ms_total"], "path": path, "dtype": dtype_of(path), "input": "rays" if run.use_rays else "features",
               "collective": run.collective, "launch": run.launch_mode, "gpu_launches_per_step": m["launches_per_step"],
               "loss_last_step": m["loss_last"], "roofline": {k: rf[k] for k in keep if k in rf}}
        if with_cpu and env.rank == 0:
            rec["cpu_baseline"] = cpu_baseline_record(wname, run.w["S"])
        return rec
    finally:
        run.close()


def render_record(env, run, n_poses=120):
    """BASELINE config 4: forward-only 800 x 800 frames, 64 samples per ray, `n_poses` orbit poses dealt round-robin over the
    ranks (no collective).  Camera mode: a pose goes in, rays and linspace depths are generated inside the kernel
    (train_nerf.py:23-62, 589-605), a uint8 frame comes out.  `value`: frames stay on the device; `e2e`: every frame is
    copied to pinned host memory inside the timed region."""
    import numpy as np
    from loma_nerf_b200 import render, sharding, synthetic
    torch = env.torch
    ctx, dims, E, S = run.ctx, run.dims, run.w["E"], run.w["S"]
    Hh = 800
    ws_np, bs_np = run.trainer.read()[:2]
    ws_d, bs_d = torch.as_tensor(ws_np, device=env.device), torch.as_tensor(bs_np, device=env.device)
    K = np.array([[synthetic.FOCAL, 0, 0.5], [0, synthetic.FOCAL, 0.5], [0, 0, 1.0]])
    poses = [render.pose_spherical(th, -30.0, 4.0) for th in np.linspace(-180.0, 180.0, n_poses, endpoint=False)]
    mine = sharding.frames_for_rank(n_poses, env.world, env.rank)
    color = torch.empty((Hh * Hh, 3), dtype=torch.float32, device=env.device)
    u8 = [torch.empty((Hh * Hh, 3), dtype=torch.uint8, device=env.device) for _ in range(2)]
    host = [torch.empty((Hh * Hh, 3), dtype=torch.uint8).pin_memory() for _ in range(4)]   # a ring: a real host consumes each frame
    # the wide (layerwise) path keeps two bf16 activation tensors per call: bound a call to ~8 M samples
    ray_chunk = None if run.w["width"] <= 62 else max(1024, (8 << 20) // S)

    def frame(i, k, to_host):
        render.render_frame_device(ctx, dims, ws_d, bs_d, Hh, Hh, K, poses[i], S, E, near=synthetic.NEAR, far=synthetic.FAR,
                                   path=run.path, out_u8=u8[k % 2], color=color, rays_per_call=ray_chunk)
        if to_host:
            host[k % 4].copy_(u8[k % 2], non_blocking=True)

    for k, i in enumerate(mine[:2]):
        frame(i, k, True)
    env.barrier()
    l0 = ctx.launches
    ms = env.timed(lambda: [frame(i, k, False) for k, i in enumerate(mine)])
    launches = ctx.launches - l0
    env.barrier()
    t0 = time.perf_counter()
    for k, i in enumerate(mine):
        frame(i, k, True)
    torch.cuda.synchronize(env.device)
    sec = env.max_over_ranks(time.perf_counter() - t0)
    prof = ctx.profile_dominant(lambda: [frame(i, k, False) for k, i in enumerate(mine[:6])])
    rays = n_poses * Hh * Hh
    fwd_flops = 2.0 * sum(dims[l] * dims[l + 1] for l in range(len(dims) - 1))
    rec = {"metric": "nerf_render_rays_per_s", "value": rays / (ms * 1e-3), "unit": "rays/s", "poses": n_poses, "rays_per_frame": Hh * Hh,
           "samples_per_ray": S, "ms_per_frame_per_gpu": ms / max(1, len(mine)), "samples_per_s": rays * S / (ms * 1e-3),
           "frames_across_gpus": env.world, "gpu_launches": launches,
           "input": "camera mode: one pose per frame, rays + linspace depths + positional encoding generated in the kernel; uint8 frame out",
           "e2e": {"value": rays / sec, "unit": "rays/s", "h2d_bytes_per_frame": 0, "d2h_bytes_per_frame": Hh * Hh * 3,
                   "param_bytes_per_frame": 216, "api": "lnb_nerf_step (camera mode) + lnb_color_to_u8, frame copied to pinned host memory",
                   "seconds": sec}}
    if prof:
        s_l = prof["ms_per_launch"] * 1e-3
        per_launch = Hh * Hh if ray_chunk is None else min(ray_chunk, Hh * Hh)
        peak = env.peaks.get("bf16_tflops", 1590.0)
        ach = per_launch * S * fwd_flops / s_l / 1e12
        rec["roofline"] = {"bound": "tensor", "kernel": prof["kernel"], "us_per_launch": s_l * 1e6, "launches_timed": prof["launches"],
                           "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                           "algorithmic_flops_per_launch": per_launch * S * fwd_flops,
                           "output_gbs": per_launch * 12 / s_l / 1e9,
                           "note": "forward only: no per-ray input, 12 B of colour written per ray -- nothing for HBM to bound; the roof "
                                   "is the tensor pipe (SURVEY 8d F_fwd per sample), and what limits it is instruction issue (ncu: 75 % of the issue slots, tensor pipe 33 %; profiles/r02_render_fwd_*): N <= 32 MMAs and the per-sample prologue / compositing arithmetic around them"}
    del color, u8, host
    return rec


def parity_multi(env, run):
    """N > 1: ONE step through the fused peer-memory exchange and through NCCL on the same shards from the same weights;
    the summed gradients must agree and every rank must hold bit-identical results (what tests/test_gpu_multi.py checks,
    here on the driver's own multi-GPU run)."""
    from loma_nerf_b200 import api
    torch, dist = env.torch, env.dist
    if env.world == 1 or run.collective != "peer":
        return None
    batch = run.batches[0]
    mk = lambda: api.Trainer(run.ctx, run.dims, run.ws_np, run.bs_np, optimizer="adam", lr=5e-4)  # noqa: E731
    ta, tb = mk(), mk()
    try:
        ta.enable_peer_allreduce()
        ta.step(**batch)
        ga = ta.grad_buffer().clone()
        tb.grad(**batch)
        gb = tb.grad_buffer()
        dist.all_reduce(gb)
        tb.apply()
        torch.cuda.synchronize(env.device)
        nP = run.nP
        den = float(gb[:nP].abs().max().item())
        max_rel = float((ga[:nP] - gb[:nP]).abs().max().item()) / max(den, 1e-30)
        loss_rel = abs(float(ga[nP].item()) - float(gb[nP].item())) / max(abs(float(gb[nP].item())), 1e-30)
        wa = torch.as_tensor(ta.read()[0], device=env.device).flatten()
        wb = torch.as_tensor(tb.read()[0], device=env.device).flatten()
        both = torch.cat([ga, wa]).view(torch.int32)
        gathered = [torch.empty_like(both) for _ in range(env.world)]
        dist.all_gather(gathered, both)
        identical = all(bool(torch.equal(g, gathered[0])) for g in gathered)
        p_rel = float((wa - wb).abs().max().item()) / max(float(wb.abs().max().item()), 1e-30)
        status = ta.comm_status()
        return {"max_rel": env.max_over_ranks(max_rel), "loss_rel": env.max_over_ranks(loss_rel), "param_max_rel_after_adam": env.max_over_ranks(p_rel),
                "ranks_identical": identical, "comm_status": status, "ranks": env.world,
                "what": "one train step on the same per-rank shards: fused peer-memory all-reduce (inside tc_reduce_kernel) vs "
                        "lnb_trainer_grad + NCCL all-reduce + lnb_trainer_apply; max |a-b| / max |b| over [d_ws | d_bs]; "
                        "ranks_identical = gradients and updated weights bit-equal on every rank"}
    finally:
        ta.close()
        tb.close()


def run_ours(args):
    env = Env()
    torch = env.torch
    world, rank = env.world, env.rank
    wname, path = args.workload, args.path
    clocks = ClockSampler(env.local)
    run = TrainRun(env, wname, path, args.input, collective=args.collective, eager=args.eager)
    w, N, R, S, c_in, fit = run.w, run.N, run.w["R"], run.w["S"], run.c_in, run.fit
    _dbg("inputs ready")
    if rank == 0:
        clocks.start()
        time.sleep(0.15)
    m = run.measure(args.steps, args.warmup, min_ms=args.min_ms)
    _dbg("timed region done: %.3f ms over %d steps" % (m["ms_total"], m["steps_timed"]))
    # ---- dominant-kernel roofline: CUDA events around that kernel alone, same inputs
    prof = run.ctx.profile_dominant(lambda: [run.step_body(i % run.n_pool) for i in range(args.steps)])
    _dbg("profile pass done")
    # ---- e2e legs: host buffers -> loss, copies inside the timed region
    e2e_n = max(3, min(args.steps, 20))
    e2e_sync = run.e2e_sync("features", e2e_n)
    e2e_feat, n_feat = run.e2e_pipelined("features")
    e2e_rays = e2e_cam = None
    n_rays = n_cam = 0
    if not fit:
        e2e_rays, n_rays = run.e2e_pipelined("rays")
        if path == "tc" and w["width"] <= 62:
            e2e_cam, n_cam = run.e2e_pipelined("camera")
    h2d_feat = (N * c_in + (0 if fit else N) + R * 3) * 4
    h2d_rays = R * 3 * 8 * 2 + N * 8 + R * 12
    clk = clocks.stop() if rank == 0 else None
    _dbg("e2e done")
    render = None
    if not args.no_render and not fit:
        render = render_record(env, run, n_poses=120 if w["width"] <= 62 else 8)
        _dbg("render done")
    pm = parity_multi(env, run) if world > 1 else None
    cfg = workload_config(wname, world, {
        "path": path, "input": "rays + sample depths (PE fused in the kernel)" if run.use_rays else "pre-encoded features (the reference .so's layout)",
        "l2": "inputs rotate over %d batches (%.0f MB) > 126 MB L2" % (run.n_pool, run.n_pool * run.bytes_per_batch / 1e6),
        "optimizer": "SGD, step 1e-4 (fit_img.py:417,512-513) inside the step" if fit else "Adam (train_nerf.py:133-161) inside the step",
        "launch": run.launch_mode,
        "collective": {"none": "none (1 GPU)", "nccl": "NCCL all-reduce of [d_ws|d_bs|loss] between gradient and optimiser kernels",
                       "peer": "one-shot all-reduce over NVLink peer memory fused into the reduction+Adam kernel"}[run.collective],
        "loss_last_step": m["loss_last"]})
    line = {"metric": "mlp_fit_train_samples_per_s" if fit else METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": m["warm"], "ms_per_step": m["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": dtype_of(path), "data": "synthetic", "config": cfg,
            "timed": {"steps_timed": m["steps_timed"], "reps_of_steps": m["reps"], "region_ms": m["ms_total"], "per_step_ms": m["per_step_ms"],
                      "note": "ONE timed region: the captured graph(s) of --steps consecutive steps replayed reps_of_steps times so that the "
                              "region is >= %.0f ms; per_step_ms = median / min / max over the graph replays inside it" % args.min_ms},
            "clocks": clk, "gpu_launches": m["launches_per_step"] * m["steps_timed"],
            "roofline": roofline_record(run, prof, m, env.peaks),
            "e2e": {"value": e2e_feat, "unit": UNIT, "h2d_bytes_per_step": h2d_feat, "d2h_bytes_per_step": 4, "steps": n_feat,
                    "h2d_gbs": e2e_feat / (world * N) * h2d_feat / 1e9,
                    "sync_value": e2e_sync,
                    "api": "lnb_trainer_submit_host per step + lnb_trainer_wait: pre-encoded features (the reference .so's input layout) from pinned "
                           "host memory, copy of batch i+1 under step i, every loss copied back; PCIe-bound (h2d_gbs = achieved host->device GB/s "
                           "per GPU); sync_value = the synchronous lnb_trainer_step_host"},
            "e2e_rays": None if fit else {"value": e2e_rays, "unit": UNIT, "h2d_bytes_per_step": h2d_rays, "d2h_bytes_per_step": 4, "steps": n_rays,
                                         "api": "the same calls in rays mode: float64 rays + depths from pinned host memory, sample positions and "
                                                "positional encoding on the device"},
            "e2e_camera": None if e2e_cam is None else {"value": e2e_cam, "unit": UNIT, "h2d_bytes_per_step": R * 4 + R * 12, "d2h_bytes_per_step": 4,
                                                        "steps": n_cam, "api": "the same calls in camera mode: a pose, int32 pixel indices and "
                                                        "targets from pinned host memory; rays, stratified depths and encoding generated in the kernel"},
            "render": render}
    if pm is not None:
        line["parity_multi"] = pm
        line["host_affinity"] = env.numa
    if world == 1 and not args.no_cpu_baseline and rank == 0:
        line["cpu_baseline"] = cpu_baseline_record(wname, S)
    run.close()
    # ---- the other BASELINE configs, compact (default run only): exact fp32 arithmetic beside the headline, C1, C5, strong scaling
    if wname == "c2" and path == "tc" and not args.no_extra:
        with_cpu = world == 1 and not args.no_cpu_baseline
        for key, fn in (
                ("exact_f32", lambda: sub_record(env, "c2", "f32", "features", 20, "nccl", False)),
                ("c1", lambda: sub_record(env, "c1", "tc", "features", 50, args.collective, with_cpu)),
                ("c5", lambda: sub_record(env, "c5", "tc", "features", 5, "nccl", with_cpu)),
                ("strong", lambda: sub_record(env, "c2", "tc", "features", 20, args.collective, False,
                                              rays_per_gpu=32768 // world))):
            try:
                line[key] = fn()
            except Exception as e:    # a sub-record never takes the main line down; say what happened
                line[key] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
            _dbg(key + " done")
        if isinstance(line.get("strong"), dict) and "value" in line["strong"]:
            line["strong"].update({"scaling": "strong", "global_rays": 32768 // world * world,
                                   "note": "a FIXED batch of 32 768 rays x 64 samples split over the ranks (SURVEY.md 8e); compare `value` across N"})
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        env.dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--path", default="tc", choices=["tc", "f32"],
                    help="tc: fused tcgen05 kernel (bf16 operands); f32: exact fp32 CUDA-core kernels")
    ap.add_argument("--input", default="features", choices=["features", "rays"],
                    help="device-resident batch format: pre-encoded features, or rays (PE fused in the kernel)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"],
                    help="N>1: gradient exchange fused into the step over peer memory, or a separate NCCL all-reduce")
    ap.add_argument("--no-render", action="store_true", help="skip the forward-only frame-render measurement")
    ap.add_argument("--no-extra", action="store_true", help="skip the compact records of the other BASELINE configs (exact_f32, c1, c5, strong)")
    ap.add_argument("--min-ms", type=float, default=50.0, help="minimum length of the timed region (the --steps graph is replayed to reach it)")
    ap.add_argument("--eager", action="store_true", help="launch kernels eagerly instead of replaying CUDA graphs")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
