"""Host-side mirror of the flat API of libloma_nerf_b200.so.

`Context.nerf_step` / `Context.fit_step` evaluate the reference's two loma programs
(/root/reference/scripts/nerf.py:1-306, scripts/mlp_fit.py:1-174) on the GPU: MLP forward,
compositing / loss, and (grad=True) the reverse-mode gradients, with the reference's buffer
layouts (padded weights [L][max_in][max_out], features [N][C_in], dists [R][S] ...).

Inputs are either all torch CUDA tensors (device pointers, asynchronous on the context's stream)
or all numpy / CPU tensors (host pointers: staged through pinned memory, synchronous).  PyTorch is
only used as the owner of device memory; the arithmetic is in the CUDA library.
"""
import ctypes

import numpy as np

from . import _lib as L

try:  # torch is plumbing (device memory / streams); the host-pointer path works without it
    import torch
except Exception:  # pragma: no cover
    torch = None

FWD_OUTPUTS = ("inter", "rgba", "alpha", "cumprod", "weights", "color", "loss")
GRAD_OUTPUTS = ("d_ws", "d_bs", "d_X", "d_target", "d_dists", "d_color", "d_inter")
PATHS = {"f32": L.PATH_F32, "tc": L.PATH_TC, "f32_layerwise": L.PATH_F32_LAYERWISE}


class LnbError(RuntimeError):
    pass


def make_camera(c2w, normalized_K, width, height=None, near=2.0, far=6.0, first_pixel=0, pixels=None, stratified=False, seed=0):
    """lnb_camera from what the reference hosts hold: the 4x4 (or 3x4) camera-to-world pose and the normalised intrinsics
    K = [[f, 0, .5], [0, f, .5], [0, 0, 1]] (train_nerf.py:265-271), near / far (train_nerf.py:201-202).  `pixels` (int32
    array, on the device for device calls) picks arbitrary pixels; otherwise ray r is pixel first_pixel + r."""
    c = L.LnbCamera()
    m = np.asarray(c2w, np.float64)[:3, :4]
    for i in range(12):
        c.c2w[i] = float(m.flat[i])
    K = np.asarray(normalized_K, np.float64)
    c.fx, c.fy, c.cx, c.cy = float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])
    c.width, c.height = int(width), int(width if height is None else height)
    c.first_pixel, c.near, c.far = int(first_pixel), float(near), float(far)
    c.stratified, c.seed = int(bool(stratified)), int(seed) & 0xFFFFFFFFFFFFFFFF
    if pixels is not None:
        assert str(pixels.dtype).endswith("int32")
        c.pixels = pixels.data_ptr() if torch is not None and isinstance(pixels, torch.Tensor) else pixels.ctypes.data
        c._keep = pixels
    return c


def make_mlp(dims, ws_shape, head):
    m = L.LnbMlp()
    dims = [int(d) for d in dims]
    if len(dims) - 1 > L.LNB_MAX_LAYERS:
        raise ValueError("too many layers")
    m.n_layers = len(dims) - 1
    for i, d in enumerate(dims):
        m.dims[i] = d
    m.max_in, m.max_out = int(ws_shape[1]), int(ws_shape[2])
    m.head = head
    return m


def _is_cuda(x):
    return torch is not None and isinstance(x, torch.Tensor) and x.is_cuda


def _ptr(x):
    if x is None:
        return None
    if torch is not None and isinstance(x, torch.Tensor):
        assert x.is_contiguous() and x.dtype in (torch.float32, torch.float64), (x.dtype, x.is_contiguous())
        return x.data_ptr()
    assert x.flags.c_contiguous
    return x.ctypes.data


class Context:
    """One lnb_ctx: a device, a stream and a grow-only workspace.

    Stream rule: the library launches on ONE stream per context.  With torch importable that stream
    defaults to torch's current stream of the device at construction time, so torch-side work on the
    tensors passed in (zero fills, NCCL collectives, allocator reuse) is ordered with the library's
    kernels without further ado.  `set_stream` moves it; buffers this class allocates are then
    filled on that stream, and caller tensors are recorded on it (`Tensor.record_stream`) so the
    caching allocator does not hand their memory out while a kernel still reads it."""

    def __init__(self, device=None, stream=None):
        self.lib = L.load()
        if self.lib.lnb_device_count() <= 0:
            raise LnbError("no CUDA device: libloma_nerf_b200 has no CPU fallback")
        if device is None:
            device = torch.cuda.current_device() if torch is not None and torch.cuda.is_available() else 0
        self.device = int(device)
        h = ctypes.c_void_p()
        rc = self.lib.lnb_create(ctypes.byref(h), self.device)
        if rc != L.LNB_OK:
            raise LnbError("lnb_create failed (%d)" % rc)
        self.h = h
        self._stream_raw = None          # None: the context's own non-blocking stream
        if stream is not None:
            self.set_stream(stream)
        elif torch is not None and torch.cuda.is_available():
            self.use_current_torch_stream()

    def close(self):
        if getattr(self, "h", None):
            self.lib.lnb_destroy(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream):
        """stream: a torch.cuda.Stream, a raw cudaStream_t int (0 = the legacy default stream), or
        None (the context's own non-blocking stream)."""
        raw = ctypes.c_void_p(-1) if stream is None else ctypes.c_void_p(getattr(stream, "cuda_stream", stream))
        self._check(self.lib.lnb_set_stream(self.h, raw))
        self._stream_raw = None if stream is None else int(getattr(stream, "cuda_stream", stream))

    def torch_stream(self):
        """The context's stream as a torch stream object (for events / wait_stream), or None when the
        context runs on its own private stream (which torch cannot see: synchronize() orders it)."""
        if torch is None or self._stream_raw is None:
            return None
        cur = torch.cuda.current_stream(self.device)
        if cur.cuda_stream == self._stream_raw:
            return cur
        return torch.cuda.ExternalStream(self._stream_raw, device=torch.device("cuda", self.device))

    def _on_torch_current(self):
        return torch is not None and self._stream_raw is not None and \
            torch.cuda.current_stream(self.device).cuda_stream == self._stream_raw

    def order_after_torch(self):
        """Make the context's stream wait for everything queued on torch's current stream so far."""
        if torch is None or self._on_torch_current():
            return
        ts = self.torch_stream()
        if ts is None:                      # private stream: the only ordering tool is the host
            torch.cuda.current_stream(self.device).synchronize()
        else:
            ts.wait_stream(torch.cuda.current_stream(self.device))

    def order_torch_after(self):
        """Make torch's current stream wait for everything the context has launched so far."""
        if torch is None or self._on_torch_current():
            return
        ts = self.torch_stream()
        if ts is None:
            self.synchronize()
        else:
            torch.cuda.current_stream(self.device).wait_stream(ts)

    def _hold(self, *tensors):
        """Caller tensors read by kernels on a stream that is not torch's current one."""
        if torch is None or self._on_torch_current():
            return
        ts = self.torch_stream()
        for x in tensors:
            if ts is not None and _is_cuda(x):
                x.record_stream(ts)

    def use_current_torch_stream(self):
        self.set_stream(torch.cuda.current_stream(self.device))

    def synchronize(self):
        self._check(self.lib.lnb_synchronize(self.h))

    @property
    def launches(self):
        return int(self.lib.lnb_launch_count(self.h))

    def profile_dominant(self, fn):
        """Run fn() with CUDA events around every launch of the dominant (fused) kernel.
        Returns dict(kernel, launches, ms_per_launch) or None when that kernel never ran."""
        self._check(self.lib.lnb_profile(self.h, 1))
        try:
            fn()
            ms, n = ctypes.c_double(), ctypes.c_longlong()
            name = ctypes.create_string_buffer(128)
            self._check(self.lib.lnb_profile_read(self.h, ctypes.byref(ms), ctypes.byref(n), name, 128))
        finally:
            self.lib.lnb_profile(self.h, 0)
        if n.value == 0:
            return None
        return dict(kernel=name.value.decode(), launches=int(n.value), ms_per_launch=ms.value / n.value)

    def _check(self, rc):
        if rc != L.LNB_OK:
            raise LnbError("libloma_nerf_b200 error %d: %s" % (rc, self.lib.lnb_last_error(self.h).decode()))

    # -------------------------------------------------------------------------------------------
    def _step(self, nerf, dims, X, ws, bs, target, dists, R, S, grad, seed, outputs, out, rows,
              path, inter_accumulate, color_accumulate, head, rays=None, pe_bands=0, camera=None):
        dev = _is_cuda(ws)
        mlp = make_mlp(dims, ws.shape, head)
        Ln, mi, mo = mlp.n_layers, mlp.max_in, mlp.max_out
        N = int(X.shape[0]) if X is not None else int(R) * int(S)
        M = max(int(rows or 0), N)
        Wt = int(target.shape[1]) if target is not None else 3
        a = L.LnbStepArgs()
        a.R, a.S, a.n_rows, a.rows, a.target_w = int(R), int(S), N, M, Wt
        a.X, a.ws, a.bs = _ptr(X), _ptr(ws), _ptr(bs)
        a.target, a.dists = _ptr(target), _ptr(dists)
        if rays is not None:
            ro, rd, tv = rays
            f64 = str(ro.dtype).endswith("float64")
            assert str(rd.dtype) == str(ro.dtype) == str(tv.dtype), "rays_o, rays_d, t must share one dtype"
            assert tuple(ro.shape) == (R, 3) and tuple(rd.shape) == (R, 3) and tuple(tv.shape) == (R, S)
            a.rays_o, a.rays_d, a.t = _ptr(ro), _ptr(rd), _ptr(tv)
            a.ray_dtype, a.pe_bands = (L.RAY_F64 if f64 else L.RAY_F32), int(pe_bands)
        if camera is not None:
            a.cam, a.pe_bands = ctypes.pointer(camera), int(pe_bands)
        a.inter_rows, a.inter_ld = M, mo
        a.inter_accumulate, a.color_accumulate = int(inter_accumulate), int(color_accumulate)
        a.want_grad = int(bool(grad))
        if isinstance(seed, str):
            assert seed == "loss"
            a.seed_mode, a.seed = L.SEED_LOSS, 1.0
        else:
            a.seed_mode, a.seed = L.SEED_VALUE, float(seed)
        a.path = PATHS[path]
        shapes = dict(inter=(Ln, M, mo), rgba=(R, S, 4), alpha=(R, S), cumprod=(R, S),
                      weights=(R, S), color=(R, 3), loss=(1,), d_ws=(Ln, mi, mo), d_bs=(Ln, mo),
                      d_X=(N, int(dims[0])), d_target=(R, Wt), d_dists=(R, S),
                      d_color=(R, Wt), d_inter=(Ln, M, mo))
        res = {}
        out = out or {}
        want = list(outputs)
        if grad:
            want += [k for k in ("d_ws", "d_bs") if k not in want]
        for k in want:
            if k not in shapes:
                raise ValueError("unknown output " + k)
            if not nerf and k in ("rgba", "alpha", "cumprod", "weights", "color", "d_dists"):
                continue
            buf = out.get(k)
            if buf is None:
                if dev:
                    buf = torch.zeros(shapes[k], dtype=torch.float32, device=ws.device)
                else:
                    buf = np.zeros(shapes[k], np.float32)
            else:
                assert tuple(buf.shape) == tuple(shapes[k]), (k, tuple(buf.shape), shapes[k])
            res[k] = buf
            setattr(a, k, _ptr(buf))
        fn = {(True, True): self.lib.lnb_nerf_step, (True, False): self.lib.lnb_nerf_step_host,
              (False, True): self.lib.lnb_fit_step, (False, False): self.lib.lnb_fit_step_host}[(nerf, dev)]
        if dev:
            # the zero fills above and whatever produced the inputs ran on torch's current stream
            self.order_after_torch()
            self._hold(X, ws, bs, target, dists, *(rays or ()), *res.values())
        self._check(fn(self.h, ctypes.byref(mlp), ctypes.byref(a)))
        return res

    def nerf_step(self, dims, X, ws, bs, dists, target=None, R=None, S=None, grad=False, seed=1.0,
                  outputs=("color", "loss"), out=None, rows=None, path="f32",
                  inter_accumulate=False, color_accumulate=False):
        """MLP -> compositing -> SSE loss [-> gradients] (scripts/nerf.py:67-306).
        X [R*S][C_in] features, dists [R][S], target [R][3]; returns {name: buffer}.  Gradient
        outputs ACCUMULATE into buffers passed through `out` (the reference's += semantics)."""
        R = int(dists.shape[0]) if R is None else R
        S = int(dists.shape[1]) if S is None else S
        return self._step(True, dims, X, ws, bs, target, dists, R, S, grad, seed, outputs, out,
                          rows, path, inter_accumulate, color_accumulate, L.HEAD_NERF)

    def nerf_step_rays(self, dims, rays_o, rays_d, t, pe_bands, ws, bs, target=None, grad=False, seed=1.0,
                       outputs=("color", "loss"), out=None, path="f32"):
        """The same step starting from rays: the sample positions o + d*t, their positional encoding
        and the dists are computed on the device (train_nerf.py:289-311, pos_encoding.py:38-70) --
        fused into the first layer on the tensor-core path.  rays_o, rays_d [R][3], t [R][S], all
        float64 (the reference's dtype) or all float32."""
        R, S = int(t.shape[0]), int(t.shape[1])
        return self._step(True, dims, None, ws, bs, target, None, R, S, grad, seed, outputs, out, None, path,
                          False, False, L.HEAD_NERF, rays=(rays_o, rays_d, t), pe_bands=pe_bands)

    def nerf_step_camera(self, dims, camera, R, S, pe_bands, ws, bs, target=None, grad=False, seed=1.0,
                         outputs=("color", "loss"), out=None, path="f32"):
        """The same step with rays AND sample depths generated on the device from a pose (`make_camera`): the
        device-side get_rays (train_nerf.py:23-62) and linspace / stratified depths (train_nerf.py:289-311).
        Nothing per ray or per sample is read: R rays of S samples through pixels first_pixel .. (or `pixels`)."""
        return self._step(True, dims, None, ws, bs, target, None, int(R), int(S), grad, seed, outputs, out, None, path,
                          False, False, L.HEAD_NERF, pe_bands=pe_bands, camera=camera)

    def camera_rays(self, camera, R, S):
        """(rays_o, rays_d, t) float64 cuda tensors camera mode generates (for checks / hosts that want them)."""
        dev = torch.device("cuda", self.device)
        o = torch.empty((R, 3), dtype=torch.float64, device=dev)
        d = torch.empty((R, 3), dtype=torch.float64, device=dev)
        t = torch.empty((R, S), dtype=torch.float64, device=dev)
        self.order_after_torch()
        self._check(self.lib.lnb_camera_rays(self.h, ctypes.byref(camera), int(R), int(S), o.data_ptr(), d.data_ptr(), t.data_ptr()))
        return o, d, t

    def color_to_u8(self, rgb, out=None):
        """uint8 image bytes of float colours, clamp(rgb, 0, 1) * 255 rounded (train_nerf.py:686-700), on the device."""
        assert _is_cuda(rgb) and rgb.dtype == torch.float32 and rgb.is_contiguous()
        if out is None:
            out = torch.empty(rgb.shape, dtype=torch.uint8, device=rgb.device)
        self.order_after_torch()
        self._check(self.lib.lnb_color_to_u8(self.h, rgb.data_ptr(), rgb.numel(), out.data_ptr()))
        return out

    def fit_step(self, dims, X, ws, bs, target, grad=False, seed=1.0, outputs=("loss",), out=None,
                 rows=None, path="f32", inter_accumulate=False):
        """MLP (sigmoid head) -> SSE loss vs target [-> gradients] (scripts/mlp_fit.py:39-174)."""
        return self._step(False, dims, X, ws, bs, target, None, int(target.shape[0]), 1, grad, seed,
                          outputs, out, rows, path, inter_accumulate, False, L.HEAD_SIGMOID)

    # -------------------------------------------------------------------------------------------
    def pos_encoding(self, x, E):
        """pos_encoding.py:4-70 on the device: x (..., F) float64 cuda -> (..., F*(1+2E)) float32."""
        assert _is_cuda(x) and x.dtype == torch.float64 and x.is_contiguous()
        F = int(x.shape[-1])
        n = x.numel() // F
        o = torch.empty(tuple(x.shape[:-1]) + (F * (1 + 2 * E),), dtype=torch.float32, device=x.device)
        self._check(self.lib.lnb_pos_encoding(self.h, x.data_ptr(), n, F, int(E), o.data_ptr()))
        return o

    def sample_encode(self, rays_o, rays_d, t, E):
        """train_nerf.py:289-311 + pos_encoding_3d: returns X [R*S][3+6E] f32 and dists [R][S]."""
        assert all(_is_cuda(v) and v.dtype == torch.float64 and v.is_contiguous() for v in (rays_o, rays_d, t))
        R, S = int(t.shape[0]), int(t.shape[1])
        X = torch.empty((R * S, 3 + 6 * E), dtype=torch.float32, device=t.device)
        dists = torch.empty((R, S), dtype=torch.float32, device=t.device)
        self._check(self.lib.lnb_sample_encode(self.h, rays_o.data_ptr(), rays_d.data_ptr(),
                                               t.data_ptr(), R, S, int(E), X.data_ptr(), dists.data_ptr()))
        return X, dists

    def mult_a_b(self, a, b, c):
        """c += a @ b on the device (scripts/mlp_fit.py:150-172)."""
        self._check(self.lib.lnb_mult_a_b(self.h, a.data_ptr(), int(a.shape[0]), int(a.shape[1]),
                                          b.data_ptr(), int(b.shape[1]), c.data_ptr()))
        return c

    def adam_step(self, param, grad, m, v, t, lr=5e-4, beta1=0.9, beta2=0.999, eps=1e-8):
        """AdamOptimizer.update of train_nerf.py:133-161 (double bias correction kept), in place."""
        self._check(self.lib.lnb_adam_step(self.h, param.data_ptr(), grad.data_ptr(), m.data_ptr(),
                                           v.data_ptr(), param.numel(), int(t), lr, beta1, beta2, eps))

    def adam_step_dev(self, param, grad, m, v, t_dev, lr=5e-4, beta1=0.9, beta2=0.999, eps=1e-8):
        """adam_step with the step counter in device memory (int32 tensor): graph-capturable."""
        self._check(self.lib.lnb_adam_step_dev(self.h, param.data_ptr(), grad.data_ptr(), m.data_ptr(),
                                               v.data_ptr(), param.numel(), t_dev.data_ptr(), lr, beta1, beta2, eps))

    def sgd_step(self, param, grad, lr):
        """ws -= lr * d_ws (fit_img.py:512-513), in place."""
        self._check(self.lib.lnb_sgd_step(self.h, param.data_ptr(), grad.data_ptr(), param.numel(), lr))


class Trainer:
    """Device-resident weights + optimiser state (lnb_trainer): the reference hosts' per-chunk
    "grad call, then Adam / SGD on the padded arrays" loop (train_nerf.py:395-499,
    fit_img.py:468-513) without host round trips.  Batches are torch CUDA tensors."""

    def __init__(self, ctx, dims, ws, bs, head=L.HEAD_NERF, optimizer="adam", lr=5e-4, beta1=0.9, beta2=0.999,
                 eps=1e-8):
        self.ctx, self.lib = ctx, ctx.lib
        ws = np.ascontiguousarray(ws, np.float32)
        bs = np.ascontiguousarray(bs, np.float32)
        self.dims, self.ws_shape, self.bs_shape = [int(d) for d in dims], ws.shape, bs.shape
        self.mlp = make_mlp(dims, ws.shape, head)
        self.head = head
        h = ctypes.c_void_p()
        opt = {"adam": L.OPT_ADAM, "sgd": L.OPT_SGD}[optimizer]
        ctx._check(self.lib.lnb_trainer_create(ctx.h, ctypes.byref(self.mlp), ws.ctypes.data, bs.ctypes.data, opt,
                                               lr, beta1, beta2, eps, ctypes.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.lnb_trainer_destroy(self.h)
            self.h = None

    def _batch(self, X=None, dists=None, target=None, rays=None, pe_bands=0, R=None, S=None, path="tc", seed=1.0, camera=None):
        a = L.LnbStepArgs()
        nerf = self.head == L.HEAD_NERF
        if camera is not None:
            R, S = int(target.shape[0]) if R is None else int(R), int(S)
            a.cam, a.pe_bands = ctypes.pointer(camera), int(pe_bands)
            self._cam_keep = camera
        elif rays is not None:
            ro, rd, tv = rays
            R, S = int(tv.shape[0]), int(tv.shape[1])
            a.rays_o, a.rays_d, a.t = _ptr(ro), _ptr(rd), _ptr(tv)
            a.ray_dtype = L.RAY_F64 if str(ro.dtype).endswith("float64") else L.RAY_F32
            a.pe_bands = int(pe_bands)
        else:
            a.X = _ptr(X)
            if nerf:
                R = int(dists.shape[0]) if R is None else R
                S = int(dists.shape[1]) if S is None else S
                a.dists = _ptr(dists)
            else:
                R, S = int(target.shape[0]), 1
        a.R, a.S, a.target_w = R, S, int(target.shape[1])
        a.n_rows = R * S
        a.target = _ptr(target)
        a.path = PATHS[path]
        if isinstance(seed, str):
            a.seed_mode, a.seed = L.SEED_LOSS, 1.0
        else:
            a.seed_mode, a.seed = L.SEED_VALUE, float(seed)
        return a, int(nerf)

    def prepare(self, **batch):
        """A batch marshalled once (the lnb_step_args struct and the buffers it points at), to be handed to step(),
        step_host() or submit_host() any number of times: a hosts' loop that cycles through pre-built batches pays one ctypes
        call per step instead of rebuilding the struct in Python."""
        a, nerf = self._batch(**batch)
        return PreparedBatch(a, nerf, batch)

    def _order(self, batch):
        self.ctx.order_after_torch()
        self.ctx._hold(*[v for v in batch.values() if _is_cuda(v)], *[x for x in (batch.get("rays") or ()) if _is_cuda(x)])

    def step(self, prepared=None, **batch):
        """forward + backward + optimiser update on one batch (keyword arguments, or a prepare()d batch)."""
        if prepared is not None:
            a, nerf, batch = prepared.args, prepared.nerf, prepared.batch
        else:
            a, nerf = self._batch(**batch)
        self._order(batch)
        self.ctx._check(self.lib.lnb_trainer_step(self.h, ctypes.byref(a), nerf))

    def step_host(self, prepared=None, **batch):
        """step() with the batch in host memory (numpy arrays / pinned CPU tensors): the C library
        stages it to the device, steps, and returns the loss.  Synchronous."""
        a, nerf = (prepared.args, prepared.nerf) if prepared is not None else self._batch(**batch)
        loss = ctypes.c_float()
        self.ctx._check(self.lib.lnb_trainer_step_host(self.h, ctypes.byref(a), nerf, ctypes.byref(loss)))
        return float(loss.value)

    def submit_host(self, prepared=None, **batch):
        """step_host() without the wait: the batch is staged on a copy stream while the previous step still runs
        (lnb_trainer_submit_host).  Pinned host buffers must stay unchanged until the submission after next."""
        if prepared is not None:
            a, nerf, batch = prepared.args, prepared.nerf, prepared
        else:
            a, nerf = self._batch(**batch)
        self._keep = (self.__dict__.get("_keep", []) + [batch])[-3:]
        self.ctx._check(self.lib.lnb_trainer_submit_host(self.h, ctypes.byref(a), nerf))

    def wait(self, max_losses=4096):
        """Block until every submitted step is done; losses of the steps submitted since the last wait (oldest first)."""
        buf = (ctypes.c_float * max_losses)()
        n = ctypes.c_int()
        self.ctx._check(self.lib.lnb_trainer_wait(self.h, buf, max_losses, ctypes.byref(n)))
        return [float(buf[i]) for i in range(n.value)]

    def grad(self, **batch):
        """forward + backward only: gradients (and loss) land in grad_buffer()."""
        a, nerf = self._batch(**batch)
        self._order(batch)
        self.ctx._check(self.lib.lnb_trainer_grad(self.h, ctypes.byref(a), nerf))

    def apply(self):
        self.ctx._check(self.lib.lnb_trainer_apply(self.h))

    def enable_peer_allreduce(self, group=None):
        """Fuse the gradient all-reduce into the step (tensor-core path, GPUs of one box): exchanges
        CUDA IPC handles of the per-rank exchange buffers through torch.distributed and attaches
        them.  Afterwards step()/step_host() sum gradients and loss over all ranks inside the
        reduction kernel (NVLink peer loads) and every rank must step in lockstep."""
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        h = ctypes.create_string_buffer(64)
        self.ctx._check(self.lib.lnb_trainer_comm_export(self.h, h))
        handles = [None] * world
        dist.all_gather_object(handles, bytes(h.raw), group=group)
        blob = ctypes.create_string_buffer(b"".join(handles), 64 * world)
        self.ctx._check(self.lib.lnb_trainer_comm_attach(self.h, rank, world, blob))
        dist.barrier(group)
        return True

    def comm_status(self):
        """0 = fine, 1 = a peer timed out and that step was poisoned with NaN.  Synchronises."""
        return int(self.lib.lnb_trainer_comm_status(self.h))

    def check_comm(self):
        if self.comm_status() != 0:
            raise LnbError("peer all-reduce: a rank did not deliver its gradients in time; the step was poisoned (NaN)")

    def grad_buffer(self):
        """torch view of the flat [d_ws | d_bs | loss] device buffer (for dist.all_reduce).  The
        library writes it on the context's stream: use ctx.order_torch_after() before torch reads it
        and ctx.order_after_torch() before apply() (sharding.data_parallel_step does both)."""
        n = ctypes.c_longlong()
        ptr = self.lib.lnb_trainer_grad_buffer(self.h, ctypes.byref(n))
        return _torch_view(ptr, n.value, self.ctx.device)

    def read(self):
        """(ws, bs, loss) as numpy; synchronises."""
        ws = np.empty(self.ws_shape, np.float32)
        bs = np.empty(self.bs_shape, np.float32)
        loss = np.empty(1, np.float32)
        self.ctx._check(self.lib.lnb_trainer_read(self.h, ws.ctypes.data, bs.ctypes.data, loss.ctypes.data))
        return ws, bs, float(loss[0])


class PreparedBatch:
    """Trainer.prepare(): the marshalled lnb_step_args of one batch plus references that keep its buffers alive."""
    __slots__ = ("args", "nerf", "batch")

    def __init__(self, args, nerf, batch):
        self.args, self.nerf, self.batch = args, nerf, batch


def _torch_view(ptr, n_floats, device):
    """float32 torch tensor aliasing n_floats of device memory at ptr (no copy, no ownership)."""
    class _Arr:
        __cuda_array_interface__ = {"shape": (int(n_floats),), "typestr": "<f4", "data": (int(ptr), False), "version": 3,
                                    "strides": None}
    return torch.as_tensor(_Arr(), device=torch.device("cuda", device))
