"""world_size-2 tests of the distributed path.  CPU: gloo + oracle backend
(host logic).  GPU: NCCL + CUDA backend, needs >= 2 devices."""
import os
import subprocess
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(nproc, mode, order=2, port=29611, case='c4', transport=None):
    env = dict(os.environ)
    if transport:
        env["CX_TRANSPORT"] = transport
    env["OMP_NUM_THREADS"] = "2"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py"),
           mode, str(order), case]
    r = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "ok=True" in r.stdout
    return r.stdout


def test_world2_gloo_oracle_backend():
    _run(2, "gloo-oracle", 2, 29611)


def test_world3_gloo_oracle_backend_order3():
    _run(3, "gloo-oracle", 3, 29612)


def test_world2_gloo_oracle_backend_multibody_nodes():
    """bench_c2n.py's workload shape: node-located receptors, donor tree is the receptor tree."""
    _run(2, "gloo-oracle", 2, 29614, case='c2n')


def test_world2_gloo_oracle_backend_cartesian_base():
    """Zones of a base named 'CARTESIAN' are InterpCart donors on every rank, also when replicated."""
    _run(2, "gloo-oracle", 2, 29621, case='cartbase')


@pytest.mark.gpu
def test_world2_nccl_device_backend_cartesian_base():
    import ctypes
    from cassiopee_b200 import _lib
    n = ctypes.c_int(0)
    _lib.lib().cx_device_count(ctypes.byref(n))
    if n.value < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, "nccl", 2, 29622, case='cartbase', transport="p2p")


@pytest.mark.gpu
def test_world2_nccl_device_backend_multibody_nodes():
    import ctypes
    from cassiopee_b200 import _lib
    n = ctypes.c_int(0)
    _lib.lib().cx_device_count(ctypes.byref(n))
    if n.value < 2:
        pytest.skip("needs 2 GPUs")
    for k, transport in enumerate(("p2p", "nccl")):
        out = _run(2, "nccl", 2, 29615 + k, case='c2n', transport=transport)
        assert ("transport=" + transport) in out


@pytest.mark.gpu
def test_world2_nccl_device_backend():
    import ctypes
    from cassiopee_b200 import _lib
    n = ctypes.c_int(0)
    _lib.lib().cx_device_count(ctypes.byref(n))
    if n.value < 2:
        pytest.skip("needs 2 GPUs")
    for k, transport in enumerate(("p2p", "nccl")):
        out = _run(2, "nccl", 2, 29617 + k, transport=transport)
        assert ("transport=" + transport) in out
