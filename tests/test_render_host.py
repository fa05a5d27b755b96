"""CPU checks of the host helpers around the render path (no CUDA calls)."""
import os

import numpy as np

from loma_nerf_b200 import render


def test_get_rays_matches_the_reference_construction():
    K = np.array([[1.2, 0, 0.5], [0, 1.2, 0.5], [0, 0, 1.0]])
    c2w = np.eye(4)
    c2w[:3, :3] = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1.0]])
    c2w[:3, 3] = [1.0, 2.0, 3.0]
    o, d = render.get_rays(4, 4, K, c2w)
    assert o.shape == (16, 3) and d.shape == (16, 3) and o.dtype == np.float64
    assert np.all(o == [1.0, 2.0, 3.0])
    # pixel (i=1/3, j=0): direction ((i-.5)/f, -(j-.5)/f, -1) @ R^T   (train_nerf.py:40-58)
    cam = np.array([(1 / 3 - 0.5) / 1.2, 0.5 / 1.2, -1.0])
    assert np.allclose(d[1], cam @ c2w[:3, :3].T)


def test_psnr_and_weight_files(tmp_path):
    a = np.full((8, 8, 3), 0.5)
    assert np.isclose(render.compute_psnr(a, a + 0.1), 20.0)          # mse 0.01 -> 20 dB
    ws = np.random.default_rng(0).normal(size=(3, 33, 30)).astype(np.float32)
    bs = np.zeros((3, 30), np.float32)
    render.save_weights(str(tmp_path) + os.sep, ws, bs)
    w2, b2 = render.load_weights(str(tmp_path) + os.sep)
    assert np.array_equal(w2, ws) and np.array_equal(b2, bs)


def test_shipped_reference_model_layout_is_the_padded_layout():
    """models/weights.npy (3,16,16) and biases.npy (3,16) in the reference tree use this layout;
    the shapes are recorded in SURVEY.md 2 (#7) -- a synthetic pair with those shapes must load."""
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        render.save_weights(d + os.sep, np.zeros((3, 16, 16)), np.zeros((3, 16)))
        ws, bs = render.load_weights(d + os.sep)
        assert ws.shape == (3, 16, 16) and bs.shape == (3, 16) and ws.dtype == np.float32


def test_blender_scene_reader_follows_the_reference_conventions(tmp_path):
    import json
    from PIL import Image
    from loma_nerf_b200.dataset import BlenderScene
    os.makedirs(tmp_path / "train")
    rng = np.random.default_rng(1)
    frames = []
    for i in range(3):
        Image.fromarray(rng.integers(0, 255, (8, 8, 4), dtype=np.uint8), "RGBA").save(tmp_path / "train" / ("r_%d.png" % i))
        pose = np.eye(4); pose[:3, 3] = [i, 0, 4.0]
        frames.append({"file_path": "./train/r_%d" % i, "transform_matrix": pose.tolist()})
    json.dump({"camera_angle_x": 0.6911, "frames": frames}, open(tmp_path / "transforms_train.json", "w"))
    sc = BlenderScene(str(tmp_path), img_size=4, phase="train")
    assert len(sc) == 3
    s = sc[1]
    assert s["image"].shape == (4, 4, 3) and s["image"].dtype == np.float32 and 0.0 <= s["image"].min() and s["image"].max() <= 1.0
    assert np.isclose(s["focal_length"], 0.5 / np.tan(0.5 * 0.6911))      # dataloader.py:54
    assert np.array_equal(s["pose"][:3, 3], [1, 0, 4.0])
    assert np.isclose(sc.normalized_K[0, 0], s["focal_length"]) and sc.normalized_K[0, 2] == 0.5


def test_pose_spherical_is_a_rigid_camera_looking_at_the_origin():
    from loma_nerf_b200 import render
    for th, ph, r in [(0.0, -30.0, 4.0), (123.0, -10.0, 2.5), (-77.0, -60.0, 6.0)]:
        c2w = render.pose_spherical(th, ph, r)
        R, T = c2w[:3, :3], c2w[:3, 3]
        assert np.allclose(R.T @ R, np.eye(3), atol=1e-12) and np.isclose(np.linalg.det(R), 1.0)
        assert np.isclose(np.linalg.norm(T), r)
        view = R @ np.array([0.0, 0.0, -1.0])                 # get_rays looks down -z
        assert np.allclose(view, -T / r, atol=1e-12)          # ... straight at the origin


def test_render_video_writes_frames_gif_and_psnr(tmp_path):
    from PIL import Image
    from loma_nerf_b200 import render
    H = 12
    gt = [np.full((H, H, 3), 0.25 * (i + 1), np.float32) for i in range(3)]
    poses = [render.pose_spherical(40.0 * i, -30.0, 4.0) for i in range(3)]
    frames, psnr = render.render_video(None, None, None, None, H, H, None, poses, 8, 5, out_dir=str(tmp_path), ground_truth=gt,
                                       frame_fn=lambda i, pose: gt[i] + (0.1 if i == 1 else 0.0))
    assert frames.shape == (3, H, H, 3) and frames.dtype == np.uint8
    assert np.isinf(psnr[0]) and np.isclose(psnr[1], 20.0, atol=1e-4) and np.isinf(psnr[2])
    for i in range(3):
        assert np.array(Image.open(tmp_path / ("frame_%04d.png" % i))).shape == (H, H, 3)
    gif = Image.open(tmp_path / "orbit.gif")
    assert getattr(gif, "n_frames", 1) == 3
    # the Motion-JPEG AVI: walk the RIFF structure and decode every frame chunk
    import io
    import struct
    raw = open(tmp_path / "orbit.avi", "rb").read()
    assert raw[:4] == b"RIFF" and raw[8:12] == b"AVI " and struct.unpack("<I", raw[4:8])[0] == len(raw) - 8
    pos, got, n_idx = 12, [], 0
    while pos < len(raw):
        tag, size = raw[pos:pos + 4], struct.unpack("<I", raw[pos + 4:pos + 8])[0]
        if tag == b"LIST" and raw[pos + 8:pos + 12] == b"hdrl":
            avih = raw[pos + 20:pos + 20 + 56]
            assert raw[pos + 12:pos + 16] == b"avih" and struct.unpack("<I", avih[16:20])[0] == 3          # total frames
            assert struct.unpack("<II", avih[32:40]) == (H, H)
        if tag == b"LIST" and raw[pos + 8:pos + 12] == b"movi":
            q = pos + 12
            while q < pos + 8 + size:
                ctag, csize = raw[q:q + 4], struct.unpack("<I", raw[q + 4:q + 8])[0]
                assert ctag == b"00dc"
                got.append(np.array(Image.open(io.BytesIO(raw[q + 8:q + 8 + csize]))))
                q += 8 + csize + (csize & 1)
        if tag == b"idx1":
            n_idx = size // 16
        pos += 8 + size + (size & 1)
    assert len(got) == 3 and n_idx == 3
    for i in range(3):
        assert got[i].shape == (H, H, 3) and np.abs(got[i].astype(int) - frames[i].astype(int)).max() <= 6   # JPEG
