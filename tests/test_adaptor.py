"""The C++ adaptor (adaptor/heph_gpu_operator.hpp: HephGpuOperator / HephGpuPCG, the classes a
Hephaestus maintainer would plug in) compiled against adaptor/mock_mfem.hpp and driven end to end:
mock ParFiniteElementSpace -> hcurl_space_create -> Mult / FormRHS / PCG, checked against the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

from hephaestus_b200 import _lib as L

from helpers import make_pair, rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def driver(lib_built, tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("adaptor") / "adaptor_driver")
    libdir = os.path.dirname(L.LIB_PATH)
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "adaptor", "adaptor_driver.cpp"), "-o", exe, "-L" + libdir, "-lhcurl_b200",
           "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def _arr(f, a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype).ravel()
    f.write(struct.pack("q", a.size))
    f.write(a.tobytes())


def _dump(path, kind, p, q, m, sp, alpha_attr, beta_attr, ess, x, xbc, b):
    with open(path, "wb") as f:
        _arr(f, [5 if kind == "hex" else 4, p, sp.nd, sp.ndof, q], np.int64)
        _arr(f, m.verts, np.float64)
        _arr(f, m.elems, np.int32)
        _arr(f, m.attr, np.int32)
        _arr(f, sp.el_faces if kind == "tet" else [], np.int32)
        _arr(f, sp.face_verts if kind == "tet" else [], np.int32)
        _arr(f, sp.signed_elem_dofs(), np.int32)
        _arr(f, alpha_attr, np.float64)
        _arr(f, beta_attr, np.float64)
        _arr(f, ess, np.int32)
        _arr(f, x, np.float64)
        _arr(f, xbc, np.float64)
        _arr(f, b, np.float64)


def _read(path):
    out = []
    with open(path, "rb") as f:
        while True:
            h = f.read(8)
            if not h:
                break
            n = struct.unpack("q", h)[0]
            out.append(np.frombuffer(f.read(8 * n), dtype=np.float64))
    return out


def _case(kind, n, p):
    m, sp, mo, so = make_pair(kind, n, p, True)
    alpha_attr, beta_attr = np.array([0.05, 0.1]), np.array([1.0, 1.0e3])
    ess = sp.boundary_dofs()
    rng = np.random.default_rng(11)
    x, b = rng.standard_normal(sp.ndof), rng.standard_normal(sp.ndof)
    xbc = np.zeros(sp.ndof)
    xbc[ess] = rng.standard_normal(len(ess))
    return m, sp, mo, so, alpha_attr, beta_attr, ess, x, xbc, b


def test_adaptor_without_gpu_aborts_loudly(driver, tmp_path):
    """No CPU fallback: on a box without CUDA the adaptor's context creation is an MFEM_ABORT."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m, sp, mo, so, aa, ba, ess, x, xbc, b = _case("hex", 2, 1)
    _dump(tmp_path / "in.bin", "hex", 1, 0, m, sp, aa, ba, ess, x, xbc, b)
    r = subprocess.run([driver, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode == 1
    assert "MFEM abort" in r.stderr and "hcurl_ctx_create failed" in r.stderr


@pytest.mark.parametrize("p", [1, 2, 3])
def test_adaptor_tensor_dofs_composition(driver, tmp_path, p):
    """HephGpuDiscreteInterpolator::TensorDofs (GetElementDofs composed with TensorBasisElement::GetDofMap, signs
    included) against the library's own composition for ND spaces (hcurl_host_hex_slot_codes) -- host only."""
    from oracle.nd_hex import NDHex
    m, sp, mo, so = make_pair("hex", (3, 2, 2), p, False, scramble=True)
    fe = NDHex(p)
    dm = np.where(np.asarray(fe.sgn) < 0, -1 - np.asarray(fe.loc), np.asarray(fe.loc))     # MFEM's dof_map encoding
    sd = np.ascontiguousarray(sp.signed_elem_dofs(), dtype=np.int32)
    with open(tmp_path / "in.bin", "wb") as f:
        _arr(f, [sd.shape[1], sd.shape[0]], np.int64)
        _arr(f, dm, np.int32)
        _arr(f, sd, np.int32)
    r = subprocess.run([driver, "--tensor-dofs", str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = _read(tmp_path / "out.bin")[0].astype(np.int64).reshape(sd.shape)
    code = np.zeros(sd.shape, np.uint32)
    L.check(L.lib().hcurl_host_hex_slot_codes(p, sd.shape[0], L.ptr(sd), L.ptr(code)), "slot codes")
    idx, neg = (code & 0x7FFFFFFF).astype(np.int64), (code >> 31).astype(bool)
    assert np.array_equal(got, np.where(neg, -1 - idx, idx))


@pytest.mark.gpu
@pytest.mark.parametrize("kind,n,p,q", [("hex", 4, 2, 0), ("hex", 3, 1, 0), ("hex", 3, 3, 4), ("tet", 3, 1, 0), ("tet", 2, 2, 0)])
def test_adaptor_end_to_end(driver, tmp_path, kind, n, p, q):
    from oracle import assemble as oa, pcg as opcg
    m, sp, mo, so, aa, ba, ess, x, xbc, b = _case(kind, n, p)
    _dump(tmp_path / "in.bin", kind, p, q, m, sp, aa, ba, ess, x, xbc, b)
    r = subprocess.run([driver, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    y, rhs, sol, info = _read(tmp_path / "out.bin")
    alpha, beta = aa[mo.attr - 1], ba[mo.attr - 1]
    A = oa.assemble(so, alpha, beta, q_override=q if q else None)
    assert rel_err(y, oa.constrained_apply(A, ess, x)) < 1e-12
    Ac, B = oa.eliminate(A, ess, xbc, b)
    assert rel_err(rhs, B) < 1e-12
    xo, it_o, _, conv_o, _ = opcg.pcg(Ac, B, diag=Ac.diagonal(), rtol=float(np.float32(1e-9)), atol=float(np.float32(1e-16)),
                                       max_iter=2000)
    assert info[2] == 1.0 and conv_o
    assert abs(info[0] - it_o) <= 2
    assert rel_err(sol, xo) < 1e-6


# ---- the GPU equation system (adaptor/gpu_curlcurl_equation_system.hpp) through the reference's per-step call sequence --------
@pytest.fixture(scope="module")
def timestep_driver(lib_built, tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("adaptor_ts") / "timestep_driver")
    libdir = os.path.dirname(L.LIB_PATH)
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "adaptor", "timestep_driver.cpp"), "-o", exe, "-L" + libdir, "-lhcurl_b200", "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def _timestep_case(kind, n, p, nsteps, dt):
    m, sp, mo, so = make_pair(kind, n, p, False)
    nu_attr, sigma_attr = np.array([1.0, 2.0]), np.array([1.0, 1.0e3])
    ess = sp.boundary_dofs()
    rng = np.random.default_rng(21)
    loads = rng.standard_normal((nsteps, sp.ndof))
    kbc = np.zeros((nsteps, sp.ndof))
    kbc[:, ess] = rng.standard_normal((nsteps, len(ess)))
    return m, sp, mo, so, nu_attr, sigma_attr, ess, loads, kbc


def _dump_timestep(path, kind, p, m, sp, nu_attr, sigma_attr, ess, loads, kbc, dt):
    with open(path, "wb") as f:
        _arr(f, [5 if kind == "hex" else 4, p, sp.nd, sp.ndof, loads.shape[0]], np.int64)
        _arr(f, [dt], np.float64)
        _arr(f, m.verts, np.float64)
        _arr(f, m.elems, np.int32)
        _arr(f, m.attr, np.int32)
        _arr(f, sp.el_faces if kind == "tet" else [], np.int32)
        _arr(f, sp.face_verts if kind == "tet" else [], np.int32)
        _arr(f, sp.signed_elem_dofs(), np.int32)
        _arr(f, nu_attr, np.float64)
        _arr(f, sigma_attr, np.float64)
        _arr(f, ess, np.int32)
        _arr(f, loads, np.float64)
        _arr(f, kbc, np.float64)


def test_gpu_equation_system_compiles_and_fails_loudly_without_a_gpu(timestep_driver, tmp_path):
    """the glue a maintainer adds compiles (-Wall -Werror) against the mock of the reference's EquationSystem hierarchy;
    on a box without CUDA its first device call is an MFEM_ABORT -- no CPU fallback"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m, sp, mo, so, nu, sg, ess, loads, kbc = _timestep_case("hex", 2, 1, 1, 0.05)
    _dump_timestep(tmp_path / "in.bin", "hex", 1, m, sp, nu, sg, ess, loads, kbc, 0.05)
    r = subprocess.run([timestep_driver, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode == 1 and "MFEM abort" in r.stderr and "hcurl_ctx_create failed" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("kind,n,p", [("hex", 4, 2), ("hex", 5, 1), ("tet", 3, 1), ("tet", 2, 2)])
def test_time_steps_through_the_gpu_equation_system(timestep_driver, tmp_path, kind, n, p):
    """three backward-Euler steps of the A-formulation driven exactly as
    TimeDomainEquationSystemProblemOperator::ImplicitSolve drives them (SetTimeStep, UpdateEquationSystem, BuildJacobian,
    NewtonSolver with the GPU PCG, RecoverFEMSolution) against the oracle: b_n = load_n - K(nu) u_n, A = M(sigma) + K(dt nu),
    essential data on k = du/dt, u += dt k.  Affine hexes take the patch kernels (HCURL_QUAD_AUTO)."""
    from oracle import assemble as oa, pcg as opcg
    dt, nsteps = 0.05, 3
    m, sp, mo, so, nu_attr, sg_attr, ess, loads, kbc = _timestep_case(kind, n, p, nsteps, dt)
    _dump_timestep(tmp_path / "in.bin", kind, p, m, sp, nu_attr, sg_attr, ess, loads, kbc, dt)
    r = subprocess.run([timestep_driver, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    u_gpu, k_gpu, iters = _read(tmp_path / "out.bin")
    nu, sg = nu_attr[mo.attr - 1], sg_attr[mo.attr - 1]
    A = oa.assemble(so, dt * nu, sg)
    K = oa.assemble(so, nu, None)
    u = np.zeros(sp.ndof)
    tol = float(np.float32(1e-12))
    for s in range(nsteps):
        Ac, B = oa.eliminate(A, ess, kbc[s], loads[s] - K @ u)
        k, it_o, _, conv, _ = opcg.pcg(Ac, B, diag=Ac.diagonal(), rtol=tol, atol=0.0, max_iter=5000)
        assert conv and abs(iters[s] - it_o) <= 2, (s, iters[s], it_o)
        u += dt * k
    assert rel_err(k_gpu, k) < 1e-8 and rel_err(u_gpu, u) < 1e-8
