"""oracle/cpu_bench.py -- times the reference's CPU implementation of the path on host cores.

TEST / MEASUREMENT INFRASTRUCTURE ONLY: used by bench.py's `cpu_baseline` leg and by
`bench.py --impl reference`, never by the product.  The thing timed is oracle/_ref/nerf.so (the
reference's own loma program compiled by its own compiler with gcc -O2; kind "reference") when it
was built, else the C restatement oracle/nerf_oracle.c (kind "port").  One worker process per
core, each running the reference's own chunk protocol (<= 256 samples per call: forward call then
grad call, train_nerf.py:275-478) on disjoint ray chunks; only the time inside the C calls is
counted (the reference's Python marshalling, ~0.1 s per chunk, is excluded -- SURVEY.md 8d).
"""
import ctypes
import multiprocessing as mp
import os
import time

import numpy as np

from . import oracle as O


def _worker(args):
    kind, seed, n_chunks, rays, S, train = args
    case = O.make_nerf_case(seed, rays, S)
    R = rays
    X, ws, bs, dims = case["X"], case["ws"], case["bs"], [int(v) for v in case["dims"]]
    target, dists = case["target"], case["dists"]
    t_c = 0.0
    if kind == "reference":
        ref = O.CompatCaller(ctypes.CDLL(os.path.join(O.REF_DIR, "nerf.so")), big_stack=False)
        box = {}

        def body():
            t = 0.0
            for _ in range(n_chunks):
                # CompatCaller times nothing itself; wrap the two C entry points
                lib = ref.lib
                f0, g0 = lib.nerf_evaluate_and_march, lib.grad_nerf_evaluate_and_march
                acc = [0.0]

                def tf(*a):
                    s = time.perf_counter(); r = f0(*a); acc[0] += time.perf_counter() - s
                    return r

                def tg(*a):
                    s = time.perf_counter(); r = g0(*a); acc[0] += time.perf_counter() - s
                    return r

                class Shim:
                    nerf_evaluate_and_march = staticmethod(tf)
                    grad_nerf_evaluate_and_march = staticmethod(tg)
                ref.lib = Shim
                try:
                    ref.nerf(X, ws, bs, dims, target, dists, R, S, g=("loss" if train else None))
                finally:
                    ref.lib = lib
                t += acc[0]
            box["t"] = t
        O.run_big_stack(body, 256 << 20)   # the grad function keeps 16 MB of tape on the stack
        t_c = box["t"]
    else:
        co = O.COracle()
        for _ in range(n_chunks):
            s = time.perf_counter()
            f = co.nerf_forward(X, ws, bs, dims, target, dists, R, S, rows=256)
            if train:
                co.nerf_backward(X, ws, bs, dims, target, dists, R, S, float(f["loss"]), rows=256)
            t_c += time.perf_counter() - s
    return t_c


def available_kind():
    return "reference" if O.have_ref("nerf") else "port"


def run(n_procs, chunks_per_proc, S=64, train=True, pool=None):
    """Every process evaluates `chunks_per_proc` chunks of (256 // S) rays x S samples.
    Returns dict(samples, seconds, kind, cores): seconds = max over workers of in-C time."""
    kind = available_kind()
    if kind == "port":
        O.build_c_oracle()
    rays = max(1, 256 // S)
    jobs = [(kind, 1000 + p, chunks_per_proc, rays, S, train) for p in range(n_procs)]
    if pool is not None:
        times = pool.map(_worker, jobs)
    elif n_procs == 1:
        times = [_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(n_procs) as pl:
            times = pl.map(_worker, jobs)
    return dict(samples=n_procs * chunks_per_proc * rays * S, seconds=max(times), kind=kind, cores=n_procs)


def _worker_big(args):
    """One call pair (forward + grad) of the paper-size reference build oracle/_ref/nerf_big.so on one ray of
    192 samples through the 63 -> 8 x 256 -> 4 network (its static tapes hold 1.9 GB on the stack)."""
    seed, n_calls, S = args
    case = O.make_nerf_case(seed, 1, S, E=10, width=256, n_layers=9)
    ref = O.load_ref("nerf_big")
    t0 = time.perf_counter()
    for _ in range(n_calls):
        ref.nerf(case["X"], case["ws"], case["bs"], [int(v) for v in case["dims"]], case["target"], case["dists"], 1, S, g="loss")
    return time.perf_counter() - t0


def run_big(n_procs, calls_per_proc=1, S=192):
    """Paper-size (BASELINE config 5) CPU baseline; needs oracle/_ref/nerf_big.so.  The time includes the
    zero-copy marshalling of the call (microseconds against seconds in the C code).  Returns None when the
    reference build is absent."""
    if not O.have_ref("nerf_big"):
        return None
    jobs = [(3000 + p, calls_per_proc, S) for p in range(n_procs)]
    if n_procs == 1:
        times = [_worker_big(jobs[0])]
    else:
        with mp.get_context("fork").Pool(n_procs) as pl:
            times = pl.map(_worker_big, jobs)
    return dict(samples=n_procs * calls_per_proc * S, seconds=max(times), kind="reference", cores=n_procs)


def _worker_fit(args):
    """The 2-D fit protocol of fit_img.py:423-532: chunks of 256 pixels, forward call then grad call."""
    kind, seed, n_chunks = args
    case = O.make_fit_case(seed, 256)
    X, ws, bs, dims, target = case["X"], case["ws"], case["bs"], [int(v) for v in case["dims"]], case["target"]
    box = {}
    if kind == "reference":
        ref = O.CompatCaller(ctypes.CDLL(os.path.join(O.REF_DIR, "mlp_fit.so")), big_stack=False)

        def body():
            t0 = time.perf_counter()
            for _ in range(n_chunks):
                ref.mlp_fit(X, ws, bs, dims, target, g="loss")
            box["t"] = time.perf_counter() - t0
        O.run_big_stack(body, 256 << 20)     # grad_mlp_fit keeps 14.7 MB of tape on the stack
        return box["t"]
    co = O.COracle()
    t0 = time.perf_counter()
    for _ in range(n_chunks):
        f = co.mlp_fit_forward(X, ws, bs, dims, target, rows=256)
        co.mlp_fit_backward(X, ws, bs, dims, target, float(f["loss"]), rows=256)
    return time.perf_counter() - t0


def run_fit(n_procs, chunks_per_proc, pool=None):
    """BASELINE config 1 on the host cores: every process runs `chunks_per_proc` chunks of 256 pixels through
    oracle/_ref/mlp_fit.so (or the C port).  The time includes the zero-copy marshalling of the calls."""
    kind = "reference" if O.have_ref("mlp_fit") else "port"
    if kind == "port":
        O.build_c_oracle()
    jobs = [(kind, 2000 + p, chunks_per_proc) for p in range(n_procs)]
    if pool is not None:
        times = pool.map(_worker_fit, jobs)
    elif n_procs == 1:
        times = [_worker_fit(jobs[0])]
    else:
        with mp.get_context("fork").Pool(n_procs) as pl:
            times = pl.map(_worker_fit, jobs)
    return dict(samples=n_procs * chunks_per_proc * 256, seconds=max(times), kind=kind, cores=n_procs)


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)
