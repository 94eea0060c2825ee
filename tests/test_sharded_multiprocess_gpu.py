"""The multi-GPU path as it really runs: one PROCESS per rank, peers opened with cudaIpcOpenMemHandle, flags spun on in the
step kernels (tests/mp_sharded_worker.py holds the rank code and the parity assertions).  With a single GPU the ranks share
cuda:0; with several each rank takes its own."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def gpu_count():
    import torch

    return torch.cuda.device_count()


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def launch(world, *args, env=None, timeout=420):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(ROOT, "tests", "mp_sharded_worker.py"), *args]
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=e, cwd=ROOT)
    return out.returncode, out.stdout + out.stderr


@pytest.mark.parametrize("world", [2, 4, 8])
def test_direct_exchange_between_processes(world):
    if world > 2 and gpu_count() < world:
        pytest.skip("more than two ranks only with one GPU per rank")
    rc, log = launch(world, "--exchange", "direct", "--steps", "3")
    assert rc == 0 and log.count("MP_SHARDED_OK") == world, log[-4000:]


def test_direct_exchange_between_processes_dense():
    rc, log = launch(2, "--exchange", "direct", "--steps", "2", "--scene", "dense", "--hitboxes", "4000")
    assert rc == 0 and log.count("MP_SHARDED_OK") == 2, log[-4000:]


@pytest.mark.parametrize("world", [2, 4])
def test_owner_migration_between_processes(world):
    """Rows change rank between the steps (ShardedEngine.migrate) while the peers keep writing halo records into the mapped
    receive areas: residents start unrelated to position, every step installs fast new velocities."""
    if world > 2 and gpu_count() < world:
        pytest.skip("more than two ranks only with one GPU per rank")
    rc, log = launch(world, "--exchange", "direct", "--steps", "4", "--migrate")
    assert rc == 0 and log.count("MP_SHARDED_OK") == world, log[-4000:]


@pytest.mark.parametrize("world", [2, 8])
def test_nccl_exchange_between_processes(world):
    if gpu_count() < world:
        pytest.skip("NCCL needs one GPU per rank")
    rc, log = launch(world, "--exchange", "nccl", "--steps", "3")
    assert rc == 0 and log.count("MP_SHARDED_OK") == world, log[-4000:]


@pytest.mark.parametrize("scenario", ["abort", "timeout"])
def test_a_lost_peer_fails_the_step_instead_of_hanging(scenario):
    rc, log = launch(2, "--exchange", "direct", "--scenario", scenario, env={"CB200_PEER_TIMEOUT_MS": "3000"})
    assert rc == 0 and "step failed as it must" in log, log[-4000:]
