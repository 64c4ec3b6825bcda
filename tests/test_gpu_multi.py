"""Multi-GPU parity (needs >= 2 GPUs; run with `gpurun --gpus N -- python -m pytest tests/test_gpu_multi.py -m gpu`, N = 2, 4, 8):
partitioned apply and PCG with the NCCL shared-dof exchange against the serial oracle.  Worlds of 4 and 8 ranks have
neighbours that share only an edge or a corner (up to 7 per rank on the 2 x 2 x 2 brick grid)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world,n,p,perturbed", [(2, 4, 2, False), (2, 3, 2, True), (2, 3, 1, True), (2, 2, 3, True),
                                                 (4, 4, 2, False), (4, 3, 1, True), (8, 4, 2, False), (8, 2, 2, True), (8, 3, 1, False)])
def test_multi_gpu_apply_and_pcg(tmp_path, lib_built, world, n, p, perturbed):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    script = os.path.join(ROOT, "tests", "gpu_multi_worker.py")
    procs = [subprocess.Popen([sys.executable, script, str(r), str(world), str(tmp_path), str(n), str(p), "1" if perturbed else "0"],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = []
    for pr in procs:
        try:
            outs.append(pr.communicate(timeout=600)[0])
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
    for pr, o in zip(procs, outs):
        assert pr.returncode == 0, o
    assert all("OK" in o for o in outs), outs


@pytest.mark.parametrize("world,name,p", [(2, "team7", 1), (2, "cylinder_hex_q2", 2), (2, "beam_tet", 3), (4, "team7", 1), (8, "cylinder_hex_q2", 2)])
def test_multi_gpu_reference_meshes_general_partition(tmp_path, lib_built, world, name, p):
    """the reference's own meshes split across 2 GPUs by a host-supplied element partition (TEAM7 runs with mpirun -n 4 in
    the reference, BASELINE.json configs[0]): apply and Jacobi-PCG against the serial oracle"""
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    script = os.path.join(ROOT, "tests", "gpu_multi_worker.py")
    procs = [subprocess.Popen([sys.executable, script, str(r), str(world), str(tmp_path), "general", name, str(p)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = []
    for pr in procs:
        try:
            outs.append(pr.communicate(timeout=900)[0])
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
    for pr, o in zip(procs, outs):
        assert pr.returncode == 0, o
    assert all("OK" in o for o in outs), outs
