"""Synthetic inputs of the shapes BASELINE.json names: Blender-lego-like rays as
get_rays makes them (/root/reference/train_nerf.py:23-62), stratified sample positions, He-normal
MLP weights in the reference's padded layout (/root/reference/mlp_utils.py:166-204,272-313).
The reference's data (data/lego, data/warren.jpeg) is not shipped, so benchmarks use these."""
import numpy as np

NEAR, FAR = 2.0, 6.0                      # train_nerf.py:201-202
FOCAL = 0.5 / np.tan(0.5 * 0.6911)        # dataloader.py: focal = 0.5 / tan(0.5 * camera_angle_x)


def mlp_dims(c_in, width, n_layers, c_out):
    return [c_in] + [width] * (n_layers - 1) + [c_out]


def init_mlp(rng, dims, sigma_bias_shift=1.0):
    """weights [in][out] ~ N(0, 2/in), biases ~ N(0, 0.5), padded to (L,max_in,max_out)/(L,max_out).
    The density-head bias is shifted so that sigma is not identically zero at initialisation (else
    every gradient vanishes: the reference's own 'gradients are zero' branch, train_nerf.py:492)."""
    L = len(dims) - 1
    ws = np.zeros((L, max(dims[:-1]), max(dims[1:])), np.float32)
    bs = np.zeros((L, max(dims[1:])), np.float32)
    for l in range(L):
        i, o = dims[l], dims[l + 1]
        ws[l, :i, :o] = rng.normal(0.0, (2.0 / i) ** 0.5, size=(i, o)).astype(np.float32)
        bs[l, :o] = rng.normal(0.0, 0.5, size=o).astype(np.float32)
    if sigma_bias_shift and dims[-1] == 4:
        bs[L - 1, 3] += np.float32(sigma_bias_shift)
    return ws, bs


def camera(rng):
    th = rng.uniform(0, 2 * np.pi)
    ph = rng.uniform(np.deg2rad(10), np.deg2rad(60))
    cam = 4.0 * np.array([np.cos(th) * np.cos(ph), np.sin(th) * np.cos(ph), np.sin(ph)])
    fwd = -cam / np.linalg.norm(cam)
    right = np.cross(fwd, np.array([0.0, 0.0, 1.0]))
    right /= np.linalg.norm(right)
    up = np.cross(right, fwd)
    return cam, np.stack([right, up, -fwd], axis=1)


def rays_for_pixels(cam, rot, i, j):
    """dirs = ((i-.5)/f, -(j-.5)/f, -1) @ R^T, not normalised (train_nerf.py:46-58); float64."""
    dirs = np.stack([(i - 0.5) / FOCAL, -(j - 0.5) / FOCAL, -np.ones_like(i)], -1)
    return np.broadcast_to(cam, dirs.shape).copy(), dirs @ rot.T


def random_rays(rng, n_rays, n_views=1):
    o_all, d_all = [], []
    per = [n_rays // n_views + (1 if v < n_rays % n_views else 0) for v in range(n_views)]
    for v in range(n_views):
        cam, rot = camera(rng)
        o, d = rays_for_pixels(cam, rot, rng.uniform(0, 1, per[v]), rng.uniform(0, 1, per[v]))
        o_all.append(o)
        d_all.append(d)
    return np.concatenate(o_all), np.concatenate(d_all)


def frame_rays(rng, height, width):
    cam, rot = camera(rng)
    jj, ii = np.meshgrid((np.arange(height) + 0.5) / height, (np.arange(width) + 0.5) / width, indexing="ij")
    return rays_for_pixels(cam, rot, ii.ravel(), jj.ravel())


def stratified_t(rng, n_rays, n_samples, near=NEAR, far=FAR):
    u = rng.uniform(0, 1, (n_rays, n_samples))
    return near + (np.arange(n_samples)[None, :] + u) * (far - near) / n_samples
