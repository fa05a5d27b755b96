"""The drop-in boundary as the reference hosts use it: `import compiler` as a TOP-LEVEL module
(train_nerf.py:4-12, fit_img.py:4-9 append loma_public/ to sys.path and import it by that name),
then the hosts' own call sequence (tests/host_replay.py) against the golden numbers the real
reference library produced for the same calls (tests/golden/host_replay.npz)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, rel_err

PKG = os.path.join(ROOT, "loma_nerf_b200")
GOLD = os.path.join(ROOT, "tests", "golden", "host_replay.npz")


@pytest.mark.parametrize("path", [os.path.join(PKG, "dropin"), PKG])
def test_compiler_imports_as_a_top_level_module(path):
    """PYTHONPATH=<...> python train_nerf.py must reach our compile()."""
    from loma_nerf_b200 import build
    build.build_library()
    env = dict(os.environ, PYTHONPATH=path)
    code = ("import compiler, ctypes\n"
            "src = 'def mlp_fit(a):\\n    pass\\n\\ndef mult_a_b(a):\\n    pass\\n\\ngrad_mlp_fit = rev_diff(mlp_fit)\\n'\n"
            "structs, lib = compiler.compile(src, target='c', output_filename='_code/mlp_fit')\n"
            "assert structs == {} and len(lib.grad_mlp_fit.argtypes) == 29 and lib.mlp_fit.restype is ctypes.c_float\n"
            "print('ok', compiler.__file__)\n")
    r = subprocess.run([sys.executable, "-c", code], env=env, cwd="/tmp", capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "ok " + path in r.stdout


def test_host_marshaller_restatement_round_trips():
    import host_replay as H
    a = np.arange(24, dtype=np.float32).reshape(2, 3, 4)
    assert np.array_equal(H.from_ctypes(H.to_ctypes(a), a.shape), a)
    i = np.arange(6, dtype=np.int32).reshape(3, 2)
    p = H.to_ctypes(i)
    assert p[2][1] == 5 and type(p[0][0]) is int
    # float64 scratch becomes c_float, as in the reference (python floats -> c_float)
    assert H.to_ctypes(np.zeros((2, 2)))._type_._type_._type_ == "f"


def test_host_adam_restatement_is_the_double_corrected_update():
    import host_replay as H
    opt = H.HostAdam(learning_rate=5e-4)
    p = [np.ones(3, np.float32)]
    g = [np.array([0.5, -1.0, 2.0], np.float32)]
    opt.update(p, g)
    c1, c2 = 1 - 0.9, 1 - 0.999
    lr_t = 5e-4 * np.sqrt(c2) / c1
    m_hat, v_hat = (0.1 * g[0]) / c1, (0.001 * g[0] ** 2) / c2
    assert np.allclose(p[0], 1 - lr_t * m_hat / (np.sqrt(v_hat) + 1e-8), rtol=1e-6)


@pytest.mark.gpu
def test_hosts_call_sequence_through_the_top_level_compiler_matches_the_reference():
    """3 chunks of train_nerf.py's loop (4 rays x 30 samples, forward call, grad call seeded with the
    loss, numpy Adam) and 3 chunks of fit_img.py's (mult_a_b assert, grad_mlp_fit, SGD, mlp_fit)."""
    import host_replay as H
    sys.modules.pop("compiler", None)
    sys.path.insert(0, os.path.join(PKG, "dropin"))
    try:
        import compiler
        assert os.path.dirname(compiler.__file__) == os.path.join(PKG, "dropin")
        nerf = H.run_nerf_host(compiler, n_chunks=3)
        fit = H.run_fit_host(compiler, n_chunks=3)
    finally:
        sys.path.remove(os.path.join(PKG, "dropin"))
        sys.modules.pop("compiler", None)
    g = np.load(GOLD)
    tol = 1e-5
    for k in ("loss", "color", "d_ws", "d_bs"):
        for c in range(3):   # per chunk: the weights of chunk c come out of c Adam steps on our gradients
            assert rel_err(nerf[k][c], g["nerf_" + k][c]) <= (tol if c == 0 else 2e-4), (k, c)
    # Adam divides by sqrt(v): the first steps move every weight by ~lr whatever the gradient's size, so
    # compare the trajectories by the size of the update
    for k in ("ws", "bs"):
        assert np.abs(nerf[k] - g["nerf_" + k]).max() <= 2e-5, k
    assert np.array_equal(fit["kat"], g["fit_kat"])
    for k in ("loss", "d_ws", "d_bs", "ws", "bs"):
        assert rel_err(fit[k], g["fit_" + k]) <= tol, k
