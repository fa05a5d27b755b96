"""Replay of the reference hosts' call sequence against ANY `compiler`-like module (test helper).

train_nerf.py:209-499 and fit_img.py:355-532 are not runnable as files here (no data/lego, no
data/warren.jpeg, wandb, matplotlib), and /root/reference does not exist on the GPU box, so this
module restates what they DO between `compiler.compile(...)` and the optimiser update:

  * the per-row marshaller (mlp_utils.py:33-118: numpy -> .tolist() -> one ctypes array per row ->
    a table of row pointers; python floats become c_float, python ints c_int) and the readers
    (mlp_utils.py:120-164, element by element through the ctypes object);
  * the 4-ray x 30-sample chunk (train_nerf.py:190-203, 275-319), float64 scratch arrays,
    (3, 256, 256) intermediate_outputs with intermediate_shapes[l] = (256, out_l) (:225-238);
  * the forward call (:325-366), the grad call with _dreturn = the loss (:395-478), the NaN guard
    (:486-489) and the numpy Adam with its double bias correction (:133-161, :499);
  * fit_img.py: the mult_a_b known answer (:363-374), per chunk grad_mlp_fit with _dreturn = the
    previous loss or 0 (:468-498), SGD (:512-513), mlp_fit for the loss (:515-532).

`run_nerf_host` / `run_fit_host` return every number the host would have logged; the golden file
tests/golden/host_replay.npz holds what the REAL reference library produced for the same calls
(tests/golden/make_host_replay_golden.py).
"""
import ctypes
from ctypes import POINTER, cast

import numpy as np


# ---- mlp_utils.py:33-118 restated --------------------------------------------------------------
def to_ctypes(arr):
    """numpy / nested list -> ctypes the way the reference marshaller does it: every innermost row
    is its own ctypes array, rows are reached through tables of pointers."""
    if isinstance(arr, np.ndarray):
        arr = arr.tolist()
    depth, probe = 0, arr
    while isinstance(probe, list):
        depth, probe = depth + 1, probe[0]
    ct = {int: ctypes.c_int, float: ctypes.c_float}[type(probe)]
    if depth == 1:
        return (ct * len(arr))(*arr)
    p1 = POINTER(ct)
    if depth == 2:
        tab = (p1 * len(arr))()
        for i, row in enumerate(arr):
            tab[i] = (ct * len(row))(*row)
        return cast(tab, POINTER(p1))
    if depth == 3:
        p2 = POINTER(p1)
        outer = (p2 * len(arr))()
        for i, plane in enumerate(arr):
            tab = (p1 * len(plane))()
            for j, row in enumerate(plane):
                tab[j] = (ct * len(row))(*row)
            outer[i] = tab
        return cast(outer, POINTER(p2))
    raise ValueError("unsupported rank")


def from_ctypes(ptr, shape):
    """mlp_utils.py:120-164: read a float** / float*** back element by element."""
    out = np.zeros(shape, np.float32)
    for idx in np.ndindex(*shape):
        p = ptr
        for i in idx:
            p = p[i]
        out[idx] = p
    return out


class HostAdam:
    """train_nerf.py:133-161, bias correction applied twice (lr_t and m_hat/v_hat)."""

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta1, beta2, epsilon
        self.m = self.v = None
        self.t = 0

    def update(self, params, grads):
        if self.m is None:
            self.m = [np.zeros_like(p) for p in params]
            self.v = [np.zeros_like(p) for p in params]
        self.t += 1
        lr_t = self.lr * (np.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t))
        for i, (p, g) in enumerate(zip(params, grads)):
            self.m[i] = self.b1 * self.m[i] + (1 - self.b1) * g
            self.v[i] = self.b2 * self.v[i] + (1 - self.b2) * (g ** 2)
            m_hat = self.m[i] / (1 - self.b1 ** self.t)
            v_hat = self.v[i] / (1 - self.b2 ** self.t)
            params[i] -= lr_t * m_hat / (np.sqrt(v_hat) + self.eps)
        return params


def nerf_inputs(seed=215, n_chunks=3, rays_per_chunk=4):
    """Synthetic stand-in for dataset + get_rays (train_nerf.py:254-271): float64 rays, targets."""
    from oracle import oracle as O
    rng = np.random.default_rng(seed)
    o, d = O.synthetic_rays(rng, n_chunks * rays_per_chunk)
    target = rng.uniform(0, 1, (n_chunks * rays_per_chunk, 3)).astype(np.float32)
    dims = O.mlp_dims(33, 30, 3, 4)
    ws, bs = O.init_mlp(np.random.default_rng(seed + 1), dims, sigma_bias_shift=1.0)
    return o, d, target, dims, ws, bs


def run_nerf_host(compiler, call=None, n_chunks=3, seed=215, source=None):
    """The chunk loop of train_nerf.py:275-499 for `n_chunks` chunks of 4 rays x 30 samples.
    `compiler` is a module with the reference's compile(); `call(fn, *args)` lets the golden
    generator run the reference's grad function on a big stack."""
    from oracle import oracle as O
    call = call or (lambda fn, *a: fn(*a))
    S, E, L, in_ch, near, far, step_size = 30, 5, 3, 33, 2.0, 6.0, 5e-4
    src = source or ("def nerf_evaluate_and_march(layer_input):\n    pass\n\n"
                     "grad_nerf_evaluate_and_march = rev_diff(nerf_evaluate_and_march)\n")
    _, lib = compiler.compile(src, target="c", output_filename="_code/nerf")
    fwd, grad = lib.nerf_evaluate_and_march, lib.grad_nerf_evaluate_and_march
    o_all, d_all, tg_all, dims, ws_padded, bs_padded = nerf_inputs(seed, n_chunks)
    ws_shape = np.array([[dims[l], dims[l + 1]] for l in range(L)], np.int32)
    bs_shape = np.array([[dims[l + 1], 1] for l in range(L)], np.int32)
    intermediate_shapes = np.array([[256, dims[l + 1]] for l in range(L)], np.int32)
    intermediate_outputs = np.zeros((L, 256, 256), np.float32)
    opt = HostAdam(learning_rate=step_size)
    log = dict(loss=[], color=[], d_ws=[], d_bs=[])
    for c in range(n_chunks):
        ro, rd, tgt = o_all[4 * c:4 * c + 4], d_all[4 * c:4 * c + 4], tg_all[4 * c:4 * c + 4]
        depth = np.linspace(near, far, S)
        pts = ro[:, None, :] + rd[:, None, :] * depth[None, :, None]
        enc = O.positional_encoding(pts, E)                                     # (4, 30, 33) f32
        dists = np.concatenate((depth[1:] - depth[:-1], np.ones_like(depth[:1]) * 1e8))[None, :].repeat(4, axis=0)
        rgba, alpha = np.zeros((4, S, 4)), np.zeros((4, S))
        cum, wgt, acc = np.zeros((4, S)), np.zeros((4, S)), np.zeros((4, 3))
        acc_c = to_ctypes(acc)
        X = enc.reshape(-1, in_ch)
        loss = call(fwd, to_ctypes(X), ctypes.c_int(X.shape[0]), ctypes.c_int(in_ch), to_ctypes(ws_padded),
                    to_ctypes(bs_padded), to_ctypes(tgt), ctypes.c_int(4), ctypes.c_int(3), L, to_ctypes(ws_shape),
                    to_ctypes(bs_shape), to_ctypes(intermediate_shapes), to_ctypes(intermediate_outputs),
                    to_ctypes(rgba), ctypes.c_int(S), to_ctypes(dists), to_ctypes(alpha), to_ctypes(cum),
                    to_ctypes(wgt), acc_c)
        di = [ctypes.c_int(v) for v in (X.shape[0], in_ch, 4, 3, L, S)]
        d_ws, d_bs = to_ctypes(np.zeros_like(ws_padded)), to_ctypes(np.zeros_like(bs_padded))
        z = lambda a: to_ctypes(np.zeros_like(a))  # noqa: E731
        call(grad, to_ctypes(X), z(X), ctypes.c_int(X.shape[0]), ctypes.byref(di[0]), ctypes.c_int(in_ch),
             ctypes.byref(di[1]), to_ctypes(ws_padded), d_ws, to_ctypes(bs_padded), d_bs, to_ctypes(tgt), z(tgt),
             ctypes.c_int(4), ctypes.byref(di[2]), ctypes.c_int(3), ctypes.byref(di[3]), ctypes.c_int(L),
             ctypes.byref(di[4]), to_ctypes(ws_shape), z(ws_shape), to_ctypes(bs_shape), z(bs_shape),
             to_ctypes(intermediate_shapes), z(intermediate_shapes), to_ctypes(intermediate_outputs),
             z(intermediate_outputs), to_ctypes(rgba), z(rgba), ctypes.c_int(S), ctypes.byref(di[5]),
             to_ctypes(dists), z(dists), to_ctypes(alpha), z(alpha), to_ctypes(cum), z(cum), to_ctypes(wgt),
             z(wgt), to_ctypes(acc), z(acc), loss)
        d_ws_padded, d_bs_padded = from_ctypes(d_ws, ws_padded.shape), from_ctypes(d_bs, bs_padded.shape)
        color = from_ctypes(acc_c, acc.shape)
        assert not (np.isnan(d_ws_padded).any() or np.isnan(d_bs_padded).any()), "NaN in gradients"  # :486-489
        ws_padded, bs_padded = opt.update([ws_padded, bs_padded], [d_ws_padded, d_bs_padded])
        log["loss"].append(np.float32(loss)); log["color"].append(color)
        log["d_ws"].append(d_ws_padded); log["d_bs"].append(d_bs_padded)
    out = {k: np.stack(v) for k, v in log.items()}
    out.update(ws=ws_padded, bs=bs_padded)
    return out


def run_fit_host(compiler, call=None, n_chunks=2, chunk=256, seed=230, source=None):
    """fit_img.py:355-532 for `n_chunks` chunks of 16 x 16 pixels of a synthetic image."""
    from oracle import oracle as O
    call = call or (lambda fn, *a: fn(*a))
    src = source or ("def mlp_fit(layer_input):\n    pass\n\ndef mult_a_b(a):\n    pass\n\n"
                     "grad_mlp_fit = rev_diff(mlp_fit)\n")
    _, lib = compiler.compile(src, target="c", output_filename="_code/mlp_fit")
    f, mult_a_b, grad_f = lib.mlp_fit, lib.mult_a_b, lib.grad_mlp_fit
    a = np.array([[1, 2], [3, 4], [5, 6]], dtype=np.float32)
    b = np.array([[100], [200]], dtype=np.float32)
    c = np.array([[0], [0], [0]], dtype=np.float32)
    c_c = to_ctypes(c)
    mult_a_b(to_ctypes(a), a.shape[0], a.shape[1], to_ctypes(b), b.shape[0], b.shape[1], c_c)
    kat = from_ctypes(c_c, c.shape)
    assert np.allclose(kat, np.array([[500], [1100], [1700]], dtype=np.float32))            # fit_img.py:374
    case = O.make_fit_case(seed, n_chunks * chunk)
    X_all, tg_all, dims = case["X"], case["target"], [int(v) for v in case["dims"]]
    ws_padded, bs_padded = case["ws"].copy(), case["bs"].copy()
    L = len(dims) - 1
    ws_shape = np.array([[dims[l], dims[l + 1]] for l in range(L)], np.int32)
    bs_shape = np.array([[dims[l + 1], 1] for l in range(L)], np.int32)
    out_t = np.zeros((n_chunks * chunk, 3), np.float32)
    step_size, loss = 1e-4, []
    log = dict(d_ws=[], d_bs=[])
    for ci in range(n_chunks):
        X, tgt, o = X_all[ci * chunk:(ci + 1) * chunk], tg_all[ci * chunk:(ci + 1) * chunk], out_t[ci * chunk:(ci + 1) * chunk]
        ish = np.array([[X.shape[0], dims[l + 1]] for l in range(L)], np.int32)             # traced shapes, :434-441
        inter = np.zeros((L, int(ish.max()), int(ish.max())), np.float32)
        di = [ctypes.c_int(v) for v in (X.shape[0], X.shape[1], tgt.shape[0], tgt.shape[1], L)]
        d_ws, d_bs = to_ctypes(np.zeros_like(ws_padded)), to_ctypes(np.zeros_like(bs_padded))
        z = lambda v: to_ctypes(np.zeros_like(v))  # noqa: E731
        call(grad_f, to_ctypes(X), z(X), ctypes.c_int(X.shape[0]), ctypes.byref(di[0]), ctypes.c_int(X.shape[1]),
             ctypes.byref(di[1]), to_ctypes(o), z(o), to_ctypes(ws_padded), d_ws, to_ctypes(bs_padded), d_bs,
             to_ctypes(tgt), z(tgt), ctypes.c_int(tgt.shape[0]), ctypes.byref(di[2]), ctypes.c_int(tgt.shape[1]),
             ctypes.byref(di[3]), L, ctypes.byref(di[4]), to_ctypes(ws_shape), z(ws_shape), to_ctypes(bs_shape),
             z(bs_shape), to_ctypes(ish), z(ish), to_ctypes(inter), z(inter),
             loss[-1] if loss else ctypes.c_float(0))
        d_ws_padded, d_bs_padded = from_ctypes(d_ws, ws_padded.shape), from_ctypes(d_bs, bs_padded.shape)
        assert not (np.isnan(d_ws_padded).any() or np.isnan(d_bs_padded).any()), "NaN in the gradients"
        ws_padded -= step_size * d_ws_padded
        bs_padded -= step_size * d_bs_padded
        step_loss = call(f, to_ctypes(X), ctypes.c_int(X.shape[0]), ctypes.c_int(X.shape[1]), to_ctypes(o),
                         to_ctypes(ws_padded), to_ctypes(bs_padded), to_ctypes(tgt), ctypes.c_int(tgt.shape[0]),
                         ctypes.c_int(tgt.shape[1]), L, to_ctypes(ws_shape), to_ctypes(bs_shape), to_ctypes(ish),
                         to_ctypes(inter))
        loss.append(step_loss)
        log["d_ws"].append(d_ws_padded); log["d_bs"].append(d_bs_padded)
    out = {k: np.stack(v) for k, v in log.items()}
    out.update(loss=np.array(loss, np.float32), ws=ws_padded, bs=bs_padded, kat=kat)
    return out
