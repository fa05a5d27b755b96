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
