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


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from loma_nerf_b200 import api, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the CUDA library has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    w = WORKLOADS[args.workload]
    E, R, S = w["E"], w["R"], w["S"]
    fit = bool(w.get("fit"))
    c_in = 2 + 4 * E if fit else 3 + 6 * E
    dims = synthetic.mlp_dims(c_in, w["width"], w["layers"], 3 if fit else 4)
    N = R * S
    path = args.path
    use_rays = args.input == "rays" and not fit

    ctx = api.Context(local)
    stream = torch.cuda.current_stream(device)
    ctx.set_stream(stream)

    # ---- synthetic inputs: a pool of batches larger than L2 (126 MB), rotated every step
    rng = np.random.default_rng(215 + 1000 * rank)          # the reference seeds numpy with 215
    if use_rays:
        bytes_per_batch = R * 3 * 8 * 2 + N * 8 + R * 12
    else:
        bytes_per_batch = N * c_in * 4 + (0 if fit else N * 4) + R * 12
    n_pool = max(2, int(np.ceil(160e6 / bytes_per_batch)) + 1)
    batches, host_batches = [], []
    for b in range(n_pool if fit else 0):
        # pixel coordinates in [0,1)^2, 5-band positional encoding (pos_encoding.py:4-36), a smooth synthetic image
        xy = rng.uniform(0, 1, (N, 2))
        target = (0.5 + 0.4 * np.sin(6 * xy[:, :1] + np.array([0.0, 1.0, 2.0])) * np.cos(4 * xy[:, 1:])).clip(0, 1).astype(np.float32)
        X = ctx.pos_encoding(torch.as_tensor(xy, device=device), E)
        tgd = torch.as_tensor(target, device=device)
        batches.append(dict(X=X, target=tgd, path=path))
        if b < 2:
            host_batches.append(dict(features=dict(X=X.cpu().pin_memory(), target=tgd.cpu().pin_memory(), path=path)))
    for b in range(0 if fit else n_pool):
        o, d = synthetic.random_rays(rng, R)
        t = synthetic.stratified_t(rng, R, S)
        target = rng.uniform(0, 1, (R, 3)).astype(np.float32)
        od, dd, td = (torch.as_tensor(v, device=device) for v in (o, d, t))
        tgd = torch.as_tensor(target, device=device)
        if use_rays:
            batches.append(dict(rays=(od, dd, td), pe_bands=E, target=tgd, path=path))
        else:
            X, dists = ctx.sample_encode(od, dd, td, E)
            batches.append(dict(X=X, dists=dists, target=tgd, path=path))
        if b < 2:   # host copies (pinned) for the end-to-end legs
            X, dists = ctx.sample_encode(od, dd, td, E)
            host_batches.append(dict(
                features=dict(X=X.cpu().pin_memory(), dists=dists.cpu().pin_memory(), target=tgd.cpu().pin_memory(), path=path),
                rays=dict(rays=tuple(torch.as_tensor(v).pin_memory() for v in (o, d, t)), pe_bands=E,
                          target=tgd.cpu().pin_memory(), path=path)))
    _dbg("inputs ready")
    ws_np, bs_np = synthetic.init_mlp(np.random.default_rng(216), dims)   # same weights on every rank
    nP = ws_np.size + bs_np.size
    if fit:
        trainer = api.Trainer(ctx, dims, ws_np, bs_np, head=api.L.HEAD_SIGMOID, optimizer="sgd", lr=1e-4)
    else:
        trainer = api.Trainer(ctx, dims, ws_np, bs_np, optimizer="adam", lr=5e-4)
    grads = trainer.grad_buffer()                           # [d_ws | d_bs | loss] on the device
    collective = "none"
    if world > 1:
        collective = "nccl"
        if args.collective == "peer" and path == "tc":
            try:
                trainer.enable_peer_allreduce()
                collective = "peer"
            except Exception as e:
                _dbg("peer all-reduce unavailable (%s); using NCCL" % str(e)[:120])
    fused_step = world == 1 or collective == "peer"

    def step_body(b):
        if fused_step:
            trainer.step(**batches[b])                      # fused fwd+bwd kernel, then reduce + Adam
        else:
            trainer.grad(**batches[b])
            dist.all_reduce(grads)                          # NCCL sum over NVLink: gradients + loss
            trainer.apply()

    # CUDA graphs.  Fused step (N=1, or N>1 with the peer-memory all-reduce inside the kernel): ONE
    # graph holds a whole run of consecutive steps over the rotating batches, so the programmatic-
    # dependent-launch edges between kernels (prologue of kernel i+1 over the tail of kernel i) also
    # span step boundaries.  NCCL mode: per batch a gradient graph, an eager all-reduce, one shared
    # optimiser graph (NCCL stays outside capture).
    graph_cache, grad_graphs, apply_graph, launch_mode = {}, [], None, "cuda-graph of consecutive steps"
    CHUNK = 240

    def capture(fn):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            ctx.set_stream(torch.cuda.current_stream(device))
            fn()
        ctx.set_stream(stream)
        return g

    if not args.eager:
        try:
            for b in range(2):
                step_body(b)
            torch.cuda.synchronize(device)
            if fused_step:
                graph_cache[1] = capture(lambda: step_body(0))      # proves capture works before the big ones
            else:
                for b in range(n_pool):
                    grad_graphs.append(capture(lambda b=b: trainer.grad(**batches[b])))
                apply_graph = capture(trainer.apply)
                launch_mode = "cuda-graph (gradient) + NCCL all-reduce + cuda-graph (optimiser) per step"
        except Exception as e:  # report, never hide: fall back to eager launches
            graph_cache, grad_graphs, apply_graph = {}, [], None
            launch_mode = "eager (graph capture failed: %s)" % str(e)[:80]
            ctx.set_stream(stream)
            torch.cuda.synchronize(device)
    else:
        launch_mode = "eager"
    _dbg("launch mode: " + launch_mode)
    counter = {"i": 0}

    def run_steps(n):
        """n consecutive train steps, continuing the batch rotation."""
        if launch_mode == "eager" or launch_mode.startswith("eager"):
            for _ in range(n):
                step_body(counter["i"] % n_pool)
                counter["i"] += 1
        elif fused_step:
            while n > 0:
                c = min(n, CHUNK)
                key = (c, counter["i"] % n_pool)
                if key not in graph_cache:
                    i0 = counter["i"]
                    graph_cache[key] = capture(lambda: [step_body((i0 + j) % n_pool) for j in range(c)])
                graph_cache[key].replay()
                counter["i"] += c
                n -= c
        else:
            for _ in range(n):
                grad_graphs[counter["i"] % n_pool].replay()
                dist.all_reduce(grads)
                apply_graph.replay()
                counter["i"] += 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def timed(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        fn()
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    warm = max(args.warmup, 3)
    run_steps(warm)
    barrier()
    if fused_step and not launch_mode.startswith("eager"):
        # capture the timed run's graph(s) now, outside the timed region (capture launches nothing)
        i_save = counter["i"]
        n = args.steps
        while n > 0:
            c = min(n, CHUNK)
            key = (c, counter["i"] % n_pool)
            if key not in graph_cache:
                i0 = counter["i"]
                graph_cache[key] = capture(lambda: [step_body((i0 + j) % n_pool) for j in range(c)])
            counter["i"] += c
            n -= c
        counter["i"] = i_save
        barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.15)
    l0 = ctx.launches
    step_body(0)                               # count this library's kernels in one step (eagerly)
    per_step = ctx.launches - l0
    barrier()
    ms = timed(lambda: run_steps(args.steps))
    _dbg("timed region done: %.3f ms" % ms)
    launches = per_step * args.steps
    loss_now = float(grads[nP].item())
    value = world * N * args.steps / (ms * 1e-3)

    # ---- dominant-kernel roofline: CUDA events around that kernel alone, same inputs
    prof = ctx.profile_dominant(lambda: [step_body(i % n_pool) for i in range(args.steps)])

    # ---- e2e: the host-buffer C-ABI step (lnb_trainer_step_host): batch copied from pinned host
    # memory, step, loss copied back, every step inside the timed region
    def e2e(kind):
        hb = [host_batches[0][kind], host_batches[1][kind]]
        for i in range(3):
            trainer.step_host(**hb[i % 2])
        barrier()
        n = max(3, min(args.steps, 20))
        t0 = time.perf_counter()
        for i in range(n):
            trainer.step_host(**hb[i % 2])
        torch.cuda.synchronize(device)
        sec = torch.tensor([time.perf_counter() - t0], device=device)
        if world > 1:
            dist.all_reduce(sec, op=dist.ReduceOp.MAX)
        return world * N * n / float(sec.item()), n

    _dbg("profile pass done")
    # ---- render (BASELINE config 4 shape): forward-only 800x800 frames, 64 samples per ray, rays mode;
    # frames are disjoint across ranks (no collective)
    render = None
    if not args.no_render and not fit:
        Hh = 800
        frames = []
        for _ in range(2):
            fo, fd = synthetic.frame_rays(rng, Hh, Hh)
            ft = synthetic.stratified_t(rng, Hh * Hh, S)
            frames.append(tuple(torch.as_tensor(v, device=device) for v in (fo, fd, ft)))
        col = torch.zeros((Hh * Hh, 3), dtype=torch.float32, device=device)

        # the wide (layerwise) path keeps two bf16 activation tensors per call: bound a call to ~8 M samples
        ray_chunk = Hh * Hh if w["width"] <= 62 else max(1024, (8 << 20) // S)

        def render_frame(i):
            fo, fd, ft = frames[i % 2]
            for r0 in range(0, Hh * Hh, ray_chunk):
                r1 = min(Hh * Hh, r0 + ray_chunk)
                ctx.nerf_step_rays(dims, fo[r0:r1], fd[r0:r1], ft[r0:r1], E, ws_d, bs_d, target=None, grad=False, outputs=("color",),
                                   out={"color": col[r0:r1]}, path=path)

        ws_d, bs_d = (torch.as_tensor(v, device=device) for v in trainer.read()[:2])
        for i in range(2):
            render_frame(i)
        n_fr = 6
        rms = timed(lambda: [render_frame(i) for i in range(n_fr)])
        render = {"metric": "nerf_render_rays_per_s", "value": world * Hh * Hh * n_fr / (rms * 1e-3), "unit": "rays/s",
                  "rays_per_frame": Hh * Hh, "samples_per_ray": S, "frames_timed": n_fr, "ms_per_frame": rms / n_fr,
                  "samples_per_s": world * Hh * Hh * S * n_fr / (rms * 1e-3),
                  "input": "float64 rays + depths resident in HBM, PE on the device", "frames_across_gpus": world}
        del frames, col
    _dbg("render done")
    e2e_feat, e2e_n = e2e("features")
    e2e_rays = None if fit else e2e("rays")[0]
    h2d_feat = (N * c_in + (0 if fit else N) + R * 3) * 4
    h2d_rays = R * 3 * 8 * 2 + N * 8 + R * 12
    clk = clocks.stop() if rank == 0 else None
    _dbg("e2e done")

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": warm, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None,
                "dtype": "f32" if path != "tc" else "bf16 operands, f32 accumulate (tcgen05)",
                "data": "synthetic",
                "config": workload_config(args.workload, world, {
                    "path": path, "input": "rays + sample depths (PE fused in the kernel)" if use_rays else "pre-encoded features (the reference .so's layout)",
                    "l2": "inputs rotate over %d batches (%.0f MB) > 126 MB L2" % (n_pool, n_pool * bytes_per_batch / 1e6),
                    "optimizer": "Adam (train_nerf.py:133-161) inside the step", "launch": launch_mode,
                    "collective": {"none": "none (1 GPU)", "nccl": "NCCL all-reduce of [d_ws|d_bs|loss] between gradient and optimiser kernels",
                                   "peer": "one-shot all-reduce over NVLink peer memory fused into the reduction+Adam kernel"}[collective],
                    "loss_last_step": loss_now}),
                "clocks": clk, "gpu_launches": launches, "render": render,
                "e2e": {"value": e2e_feat, "unit": UNIT, "h2d_bytes_per_step": h2d_feat, "d2h_bytes_per_step": 4,
                        "steps": e2e_n, "api": "lnb_trainer_step_host, pre-encoded features from pinned host memory"},
                "e2e_rays": {"value": e2e_rays, "unit": UNIT, "h2d_bytes_per_step": h2d_rays, "d2h_bytes_per_step": 4,
                             "steps": e2e_n, "api": "lnb_trainer_step_host, rays mode: float64 rays + depths from pinned host "
                                                    "memory, sample positions and positional encoding on the device"}}
        fl = flops_per_sample(dims)
        alg_bytes = (N * 8 + R * 60) if use_rays else (N * (c_in * 4 + (0 if fit else 4)) + R * 12)
        if fit:
            line["metric"] = "mlp_fit_train_samples_per_s"
            line["config"]["optimizer"] = "SGD, step 1e-4 (fit_img.py:417,512-513) inside the step"
            line["e2e_rays"] = None
        if prof and prof["kernel"] == "chain_tc_kernel":
            # wide MLP, chained layers (wide_tc.cu chain_tc_kernel): two launches per step, the forward chain (all L
            # layers) and the adjoint chain (L-1 layers).  Activations are re-read from L2, so the launch is bound by
            # the tensor pipe and what feeds it (shared-memory bandwidth), not by HBM: SURVEY 8d's FLOPs per sample
            # (forward 2*sum(in*out); adjoint the same without layer 0) x samples / launch time against the bf16 peak.
            sec = prof["ms_per_launch"] * 1e-3
            f_fwd = 2.0 * sum(dims[l] * dims[l + 1] for l in range(len(dims) - 1))
            f_adj = 2.0 * sum(dims[l] * dims[l + 1] for l in range(1, len(dims) - 1))
            flops_launch = N * (f_fwd + f_adj) / 2
            tpeak = peaks.get("bf16_tflops_sustained", 1389.4)
            ach = flops_launch / sec / 1e12
            traffic, traffic_src = None, None
            prof_csv = os.path.join(ROOT, "profiles", "r01_wide_chain_ncu_full_per_launch.csv")
            if args.workload == "c5" and os.path.exists(prof_csv):
                try:
                    import csv as _csv
                    rows = [r for r in _csv.reader(open(prof_csv))][1:]
                    g = [float(r[4]) + float(r[5]) for r in rows if r[2] == "chain_tc_kernel"]   # bytes read + written
                    if g:
                        traffic, traffic_src = sum(g) / len(g), "profiles/r01_wide_chain_ncu_full_per_launch.csv (mean of the forward and the adjoint chain)"
                except Exception:
                    pass
            pad = lambda v: (v + 63) // 64 * 64  # noqa: E731
            Lw = len(dims) - 1
            b_fwd = N * (pad(dims[0]) * 2 + sum(pad(dims[l + 1]) * 2 + pad(dims[l + 1]) // 8 for l in range(Lw - 1)) + 16)
            b_adj = N * (pad(dims[Lw]) * 2 + sum(pad(dims[l]) * 2 + pad(dims[l]) // 8 for l in range(1, Lw)))
            line["roofline"] = {"bound": "tensor", "achieved": ach, "peak": tpeak, "unit": "TFLOP/s", "frac": ach / tpeak,
                                "traffic": traffic, "traffic_source": traffic_src, "kernel": prof["kernel"], "us_per_launch": sec * 1e6,
                                "launches_timed": prof["launches"], "launches_per_step": 2,
                                "algorithmic_flops_per_launch": flops_launch, "algorithmic_bytes_per_launch": (b_fwd + b_adj) / 2,
                                "hbm_gbs": (b_fwd + b_adj) / 2 / sec / 1e9,
                                "step_tflops": N * fl / (ms / args.steps * 1e-3) / 1e12,
                                "step_tensor_frac": N * fl / (ms / args.steps * 1e-3) / 1e12 / tpeak,
                                "note": "the rest of the step is the HBM-bound weight-gradient kernels (dw_tc_kernel, one launch per layer); "
                                        "step_tensor_frac is SURVEY 8d's F_train x samples / step time against the sustained bf16 peak",
                                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside the step)" if peaks else "fallback 1389 TFLOP/s"}
        elif prof and prof["kernel"] == "gemm_tc_kernel":
            # wide MLP: layerwise tensor-core GEMMs with bf16 activations in HBM (wide_tc.cu).  Per launch the
            # kernel reads A (rows x K bf16) [+ 32 B/row of ReLU bits when masked] and writes rows x 256 bf16
            # [+ 32 B/row of bits] (the head writes 16 B/row): summed over the step's launches below.
            sec = prof["ms_per_launch"] * 1e-3
            pad = lambda v: (v + 63) // 64 * 64  # noqa: E731
            Lw = len(dims) - 1
            gemms = [(pad(dims[l]), pad(dims[l + 1]), N * (pad(dims[l]) * 2 + pad(dims[l + 1]) * 2 + pad(dims[l + 1]) // 8)) for l in range(Lw - 1)]
            gemms.append((pad(dims[Lw - 1]), 16, N * (pad(dims[Lw - 1]) * 2 + 16)))
            gemms += [(pad(dims[l + 1]), pad(dims[l]), N * (pad(dims[l + 1]) * 2 + pad(dims[l]) * 2 + pad(dims[l]) // 8)) for l in range(Lw - 1, 0, -1)]
            bytes_launch = sum(g[2] for g in gemms) / len(gemms)
            flops_launch = sum(2.0 * N * g[0] * g[1] for g in gemms) / len(gemms)
            peak = peaks.get("hbm_gbs", 6650.0)
            ach = bytes_launch / sec / 1e9
            traffic, traffic_src = None, None
            prof_csv = os.path.join(ROOT, "profiles", "r01_wide_c5_ncu_full_per_launch.csv")
            if args.workload == "c5" and os.path.exists(prof_csv):
                try:
                    import csv as _csv
                    rows = [r for r in _csv.reader(open(prof_csv))][1:]
                    g = [float(r[4]) + float(r[5]) for r in rows if r[2] == "gemm_tc_kernel"]   # Mbyte read + written
                    if g:
                        traffic, traffic_src = sum(g) / len(g) * 1e6, "profiles/r01_wide_c5_ncu_full_per_launch.csv (mean of the step's %d launches)" % len(g)
                except Exception:
                    pass
            tpeak = peaks.get("bf16_tflops_sustained", 1389.4)
            line["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                "traffic": traffic, "traffic_source": traffic_src, "kernel": prof["kernel"], "us_per_launch": sec * 1e6,
                                "launches_timed": prof["launches"], "launches_per_step": len(gemms),
                                "algorithmic_bytes_per_launch": bytes_launch, "algorithmic_tflops": flops_launch / sec / 1e12,
                                "step_tflops": N * fl / (ms / args.steps * 1e-3) / 1e12,
                                "step_tensor_frac": N * fl / (ms / args.steps * 1e-3) / 1e12 / tpeak,
                                "tensor_peak": tpeak,
                                "note": "layerwise design: every GEMM streams its activations through HBM, so the kernel is bound by "
                                        "HBM (frac) although the step as a whole is compute-heavy; step_tensor_frac is SURVEY 8d's "
                                        "F_train x samples / step time against the sustained bf16 peak",
                                "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy), bf16_tflops_sustained" if peaks else "fallback 6650 GB/s, 1389 TFLOP/s"}
        elif prof:
            sec = prof["ms_per_launch"] * 1e-3
            if use_rays:
                peak = peaks.get("bf16_tflops", 1590.0)
                ach = N * fl / sec / 1e12
                line["roofline"] = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                                    "traffic": None, "kernel": prof["kernel"], "us_per_launch": sec * 1e6,
                                    "launches_timed": prof["launches"], "algorithmic_flops_per_launch": N * fl,
                                    "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst)" if peaks else "fallback 1590 TFLOP/s"}
            else:
                # the fused kernel reads pre-encoded features: HBM-bound by SURVEY.md 8d's per-unit bytes
                peak = peaks.get("hbm_gbs", 6650.0)
                ach = alg_bytes / sec / 1e9
                traffic, traffic_src = None, None
                prof_csv = os.path.join(ROOT, "profiles", "r01_fused_tc_features_ncu_full_summary.csv")
                if args.workload == "c2" and os.path.exists(prof_csv):
                    # DRAM bytes of one launch of this kernel on this workload, from the committed ncu --set full capture
                    try:
                        vals = {}
                        for ln in open(prof_csv):
                            parts = [c.strip('"') for c in ln.strip().split('","')]
                            if len(parts) == 3 and parts[0].lstrip('"') in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                                mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(parts[1], None)
                                if mult:
                                    vals[parts[0].lstrip('"')] = float(parts[2].rstrip('"')) * mult
                        if len(vals) == 2:
                            traffic, traffic_src = sum(vals.values()), "profiles/r01_fused_tc_features_ncu_full_summary.csv"
                    except Exception:
                        pass
                line["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                    "traffic": traffic, "traffic_source": traffic_src, "kernel": prof["kernel"], "us_per_launch": sec * 1e6,
                                    "launches_timed": prof["launches"], "algorithmic_bytes_per_launch": alg_bytes,
                                    "algorithmic_tflops": N * fl / sec / 1e12,
                                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s"}
        else:
            # no single dominant kernel on the layerwise path: report the whole step against HBM
            peak = peaks.get("hbm_gbs", 6650.0)
            ach = alg_bytes / (ms / args.steps * 1e-3) / 1e9
            line["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                "traffic": None, "kernel": "whole step (layerwise fp32 kernels)",
                                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"}
        if world == 1 and not args.no_cpu_baseline and args.workload == "c5":
            from oracle import cpu_bench
            procs = min(cpu_bench.host_cores(), 8)
            r = cpu_bench.run_big(procs, 2, S=S)
            if r is not None:
                line["cpu_baseline"] = {"value": r["samples"] / r["seconds"], "unit": UNIT, "cores": procs, "kind": "reference",
                                        "sample": "%d processes x 2 call pairs x 1 ray x %d samples of this network through the "
                                                  "reference program rebuilt with larger tapes (oracle/_ref/nerf_big.so, 1.9 GB of "
                                                  "stack each), forward + grad call, wall time of the slowest process" % (procs, S)}
        elif world == 1 and not args.no_cpu_baseline and fit:
            from oracle import cpu_bench
            cores = cpu_bench.host_cores()
            r = cpu_bench.run_fit(cores, 40)
            line["cpu_baseline"] = {"value": r["samples"] / r["seconds"], "unit": UNIT, "cores": cores, "kind": r["kind"],
                                    "sample": "%d cores x 40 chunks x 256 pixels, mlp_fit + grad_mlp_fit per chunk (fit_img.py:423-532)" % cores}
        elif world == 1 and not args.no_cpu_baseline:
            from oracle import cpu_bench
            cores = cpu_bench.host_cores()
            r = cpu_bench.run(cores, 40, S=S)
            r1 = cpu_bench.run(1, 40, S=S)
            line["cpu_baseline"] = {"value": r["samples"] / r["seconds"], "unit": UNIT, "cores": cores, "kind": r["kind"],
                                    "sample": "%d cores x 40 chunks x 256 samples (4 rays x 64), forward + grad call per chunk, "
                                              "time inside the C calls only" % cores,
                                    "value_1core": r1["samples"] / r1["seconds"],
                                    "note": "the reference's own end-to-end rate is ~1e3 samples/s: its Python marshalling "
                                            "(mlp_utils.py:33-164) costs ~0.1 s per 120-sample chunk (SURVEY.md 6); excluded here"}
        print(json.dumps(line))
    trainer.close()
    if world > 1:
        dist.destroy_process_group()
    ctx.close()
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
    ap.add_argument("--eager", action="store_true", help="launch kernels eagerly instead of replaying CUDA graphs")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
