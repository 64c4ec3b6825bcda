"""Host-side handles over the C ABI: context, space, operator, PCG.

Naming follows the reference's plug points:
``HephGpuOperator`` plays the role of the ``mfem::Operator`` that
``EquationSystem::FormLinearSystem`` hands out
(/root/reference/src/equation_systems/equation_system.cpp:132-210) and
``HephGpuPCG`` the role of ``mfem::HyprePCG`` + ``HypreDiagScale``
(/root/reference/src/solvers/hephaestus_solvers.hpp:34-57).  Method names keep
MFEM's (`Mult`, `SetOperator`, `Height`).
"""
import ctypes as C

import numpy as np

from . import _lib as L


class Context:
    """One per rank / GPU.  `unique_id` (128 bytes) + rank/nranks join a NCCL
    communicator; omit for a single GPU."""

    def __init__(self, device=0, unique_id=None, rank=0, nranks=1):
        lib = L.lib()
        uid = None
        if unique_id is not None:
            self._uid = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
            uid = C.cast(self._uid, C.c_void_p)
        self.h = L.check_ptr(lib.hcurl_ctx_create(device, uid, rank, nranks), "hcurl_ctx_create")
        self.device, self.rank, self.nranks = device, rank, nranks

    @staticmethod
    def unique_id():
        buf = (C.c_char * 128)()
        L.check(L.lib().hcurl_comm_unique_id(C.cast(buf, C.c_void_p)), "hcurl_comm_unique_id")
        return bytes(buf)

    def synchronize(self):
        L.check(L.lib().hcurl_ctx_synchronize(self.h), "hcurl_ctx_synchronize")

    def alloc(self, n):
        return DeviceVector(self, n)

    def mark(self, which):
        L.check(L.lib().hcurl_ctx_mark(self.h, which), "hcurl_ctx_mark")

    def elapsed_ms(self):
        return L.lib().hcurl_ctx_elapsed_ms(self.h)

    def launch_count(self):
        return L.lib().hcurl_ctx_launch_count(self.h)

    def fp64_peak_tflops(self):
        return L.lib().hcurl_fp64_peak_tflops(self.h)

    def close(self):
        if self.h:
            L.lib().hcurl_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceVector:
    """FP64 device vector owned by the library (plain cudaMalloc)."""

    def __init__(self, ctx, n):
        self.ctx, self.n = ctx, int(n)
        self.ptr = L.check_ptr(L.lib().hcurl_device_alloc(ctx.h, self.n), "hcurl_device_alloc")

    def upload(self, a):
        a = L.as_f64(a)
        assert a.size == self.n
        L.check(L.lib().hcurl_copy_to_device(self.ctx.h, C.c_void_p(self.ptr), L.ptr(a), self.n), "copy_to_device")
        return self

    def download(self):
        out = np.empty(self.n)
        L.check(L.lib().hcurl_copy_to_host(self.ctx.h, L.ptr(out), C.c_void_p(self.ptr), self.n), "copy_to_host")
        return out

    def free(self):
        if self.ptr:
            L.lib().hcurl_device_free(self.ctx.h, C.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def pinned_array(n):
    """numpy view over page-locked host memory (freed when the array is collected)."""
    p = L.check_ptr(L.lib().hcurl_host_alloc_pinned(int(n)), "hcurl_host_alloc_pinned")
    buf = (C.c_double * int(n)).from_address(p)
    arr = np.frombuffer(buf, dtype=np.float64)
    _PINNED[arr.ctypes.data] = p
    return arr


_PINNED = {}


class GpuNDSpace:
    """Device-resident ND space.  Build from a serial `fespace.NDSpace`, or from
    a `partition.LocalSpace` (one rank's share) for multi-GPU."""

    def __init__(self, ctx, space, local=None):
        lib = L.lib()
        self.ctx = ctx
        mesh = space.mesh
        geom = L.HEX if mesh.kind == "hex" else L.TET
        gorder = 2 if mesh.geom_nodes is not None else 1
        if local is None:
            xyz = L.as_f64(mesh.elem_xyz())
            ldof = L.as_i32(space.signed_elem_dofs())
            ev = L.as_i32(mesh.elems) if geom == L.TET else None
            fb = space.elem_face_base if geom == L.TET and space.p >= 2 else None
            n_elem, n_ldof = mesh.n_elem, space.ndof
            args = (None, n_ldof, 0, None, None, None, None, None, None)
            self._keep = (xyz, ldof, ev, fb)
        else:
            xyz = L.as_f64(local.elem_xyz)
            ldof = L.as_i32(local.elem_ldof)
            ev = L.as_i32(local.elem_verts) if geom == L.TET else None
            fb = L.as_i32(local.elem_face_base) if (geom == L.TET and space.p >= 2) else None
            n_elem, n_ldof = len(local.elem_ldof), local.n_ldof
            lt = L.as_i32(local.ldof_ltdof)
            nbr = L.as_i32(local.nbr_rank)
            oo, ol = L.as_i32(local.own_off), L.as_i32(local.own_ldof)
            go, gl = L.as_i32(local.ghost_off), L.as_i32(local.ghost_ldof)
            gs = np.ascontiguousarray(local.ghost_sign, dtype=np.int8) if local.ghost_sign is not None else None
            args = (L.ptr(lt), local.n_tdof, len(nbr), L.ptr(nbr), L.ptr(oo), L.ptr(ol), L.ptr(go), L.ptr(gl), L.ptr(gs))
            self._keep = (xyz, ldof, ev, fb, lt, nbr, oo, ol, go, gl, gs)
        self.h = L.check_ptr(
            lib.hcurl_space_create(ctx.h, geom, gorder, space.p, n_elem, L.ptr(xyz), L.ptr(ldof), L.ptr(ev), L.ptr(fb),
                                   n_ldof, args[0], args[1], args[2], *args[3:]),
            "hcurl_space_create")
        self._keep = None
        self.n_elem = n_elem
        self.n_tdof = lib.hcurl_space_ndof_true(self.h)
        self.n_ldof = lib.hcurl_space_ndof_local(self.h)
        self.nd = lib.hcurl_space_elem_ndof(self.h)
        self.order = space.p
        self.kind = mesh.kind

    def close(self):
        if self.h:
            L.lib().hcurl_space_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class HephGpuOperator:
    """A = K(alpha) + M(beta), essential rows/cols -> identity; an
    ``mfem::Operator`` on T-vectors (Mult), cf. EquationSystem::Mult
    (/root/reference/src/equation_systems/equation_system.cpp:200-204)."""

    def __init__(self, ctx, gspace, quad_pts_per_dim=0):
        self.ctx, self.space = ctx, gspace
        self.h = L.check_ptr(L.lib().hcurl_op_create(ctx.h, gspace.h, quad_pts_per_dim), "hcurl_op_create")
        self.qpts_curl = L.lib().hcurl_op_qpts_curl(self.h)
        self.qpts_mass = L.lib().hcurl_op_qpts_mass(self.h)

    def Height(self):
        return self.space.n_tdof

    Width = Height

    def SetTetRules(self, pts_curl, w_curl, pts_mass, w_mass):
        """Tets, coefficients varying inside an element: the quadrature rules MFEM's integrators use (IntRules of order
        2p-2 for the curl-curl term, 2p for the mass term), after which SetCoefficients(layout=PER_QPT) takes values per
        rule point."""
        pc, wc, pm, wm = (L.as_f64(np.ascontiguousarray(v)) for v in (pts_curl, w_curl, pts_mass, w_mass))
        L.check(L.lib().hcurl_op_set_tet_rules(self.h, len(wc), L.ptr(pc), L.ptr(wc), len(wm), L.ptr(pm), L.ptr(wm)), "hcurl_op_set_tet_rules")
        self.qpts_curl, self.qpts_mass = len(wc), len(wm)

    def SetCoefficients(self, alpha=None, beta=None, layout=L.PER_ELEM):
        a = L.as_f64(alpha) if alpha is not None else None
        b = L.as_f64(beta) if beta is not None else None
        ne = self.space.n_elem
        for v, nq, nm in ((a, self.qpts_curl, "alpha"), (b, self.qpts_mass, "beta")):
            if v is not None and v.size != ne * (nq if layout == L.PER_QPT else 1):
                raise L.HcurlError("%s has %d values, expected %d" % (nm, v.size, ne * (nq if layout == L.PER_QPT else 1)))
        L.check(L.lib().hcurl_op_set_coefficients(self.h, layout, L.ptr(a), L.ptr(b), L.HOST), "hcurl_op_set_coefficients")

    def SetKernel(self, which):
        """0 = automatic (fast path where available), 1 = generic kernel."""
        L.check(L.lib().hcurl_op_set_kernel(self.h, int(which)), "hcurl_op_set_kernel")

    def LastKernel(self):
        """Which kernel the most recent apply launched: 1 generic, 2 element-wise fast, 3 patch (0 = none yet)."""
        return int(L.lib().hcurl_op_last_kernel(self.h))

    def SetEssentialTrueDofs(self, ess):
        e = L.as_i32(ess)
        L.check(L.lib().hcurl_op_set_essential_tdofs(self.h, L.ptr(e), e.size), "hcurl_op_set_essential_tdofs")

    def Mult(self, x, y=None):
        """y = A x.  numpy in -> numpy out (host path, copies inside);
        DeviceVector in -> DeviceVector out (device-resident path)."""
        if isinstance(x, DeviceVector):
            assert isinstance(y, DeviceVector)
            # vectors with room for the ghosts are used in place (no staging copies on ranks that hold ghosts)
            n_ext = int(L.lib().hcurl_space_ext_size(self.space.h))
            where = L.DEVICE_EXT if min(x.n, y.n) >= n_ext else L.DEVICE
            L.check(L.lib().hcurl_op_apply(self.h, C.c_void_p(x.ptr), C.c_void_p(y.ptr), where), "hcurl_op_apply")
            return y
        x = L.as_f64(x)
        if x.size != self.space.n_tdof:
            raise L.HcurlError("Mult: x has %d entries, operator width is %d" % (x.size, self.space.n_tdof))
        if y is None:
            y = np.empty_like(x)
        L.check(L.lib().hcurl_op_apply(self.h, L.ptr(x), L.ptr(y), L.HOST), "hcurl_op_apply")
        return y

    def ArrayMult(self, xs, ys=None):
        """y_j = A x_j for a list of host vectors (mfem::Operator::ArrayMult): uploads, applies and downloads of
        consecutive vectors overlap (hcurl_op_apply_batch).  Pinned buffers (Context.pinned) make the copies asynchronous."""
        xs = [L.as_f64(x) for x in xs]
        ys = [np.empty_like(x) for x in xs] if ys is None else ys
        n = len(xs)
        px = (C.c_void_p * n)(*[x.ctypes.data for x in xs])
        py = (C.c_void_p * n)(*[y.ctypes.data for y in ys])
        L.check(L.lib().hcurl_op_apply_batch(self.h, n, px, py, L.HOST), "hcurl_op_apply_batch")
        return ys

    def MultUnconstrained(self, x, y=None, scale=1.0, accumulate=False):
        x = L.as_f64(x)
        if y is None:
            y = np.zeros_like(x)
        L.check(L.lib().hcurl_op_apply_unconstrained(self.h, L.ptr(x), L.ptr(y), scale, int(accumulate), L.HOST),
                "hcurl_op_apply_unconstrained")
        return y

    def Diagonal(self):
        d = np.empty(self.space.n_tdof)
        L.check(L.lib().hcurl_op_diagonal(self.h, L.ptr(d), L.HOST), "hcurl_op_diagonal")
        return d

    def FormRHS(self, x_bc, b):
        """FormLinearSystem's right-hand side: B = b - A x_bc|ess, B[ess] = x_bc[ess]."""
        x_bc, b = L.as_f64(x_bc), L.as_f64(b)
        B = np.empty_like(b)
        L.check(L.lib().hcurl_op_form_rhs(self.h, L.ptr(x_bc), L.ptr(b), L.ptr(B), L.HOST), "hcurl_op_form_rhs")
        return B

    def last_kernel_ms(self):
        return L.lib().hcurl_op_last_kernel_ms(self.h)

    def kernel_stats_reset(self):
        L.lib().hcurl_op_kernel_stats_reset(self.h)

    def kernel_stats_ms(self):
        n = C.c_int(0)
        ms = L.lib().hcurl_op_kernel_stats_ms(self.h, C.byref(n))
        return ms, n.value

    def close(self):
        if self.h:
            L.lib().hcurl_op_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def complex_mult(op_re, op_im, xr, xi):
    """[[Ar,-Ai],[Ai,Ar]] block apply (HERMITIAN convention),
    /root/reference/src/formulations/ComplexMaxwell/complex_maxwell_formulation.cpp:181-217."""
    xr, xi = L.as_f64(xr), L.as_f64(xi)
    yr, yi = np.empty_like(xr), np.empty_like(xi)
    L.check(L.lib().hcurl_op_apply_complex(op_re.h, op_im.h, L.ptr(xr), L.ptr(xi), L.ptr(yr), L.ptr(yi), L.HOST),
            "hcurl_op_apply_complex")
    return yr, yi


class HephGpuInterpolator:
    """Discrete gradient (H1 -> ND) / curl (ND -> RT) on hexes; mirrors mfem::ParDiscreteLinearOperator with a
    GradientInterpolator / CurlInterpolator as the reference uses them on grid functions
    (/root/reference/src/auxsolvers/curl_aux.cpp:13-30, /root/reference/src/sources/scalar_potential_source.cpp:89-97)."""
    GRAD, CURL = 0, 1

    def __init__(self, ctx, kind, order, in_dofs, out_dofs, n_in, n_out, geom=L.HEX):
        """hexes: element dofs in tensor order, orders 1..4; geom=L.TET: order 1, MFEM's local vertex / edge / face order"""
        self.ctx = ctx
        i, o = L.as_i32(in_dofs), L.as_i32(out_dofs)
        self.n_in, self.n_out = int(n_in), int(n_out)
        self.h = L.check_ptr(L.lib().hcurl_interp_create_geom(ctx.h, geom, kind, order, i.shape[0], L.ptr(i), L.ptr(o), self.n_in, self.n_out),
                             "hcurl_interp_create_geom")

    def Mult(self, x, y=None):
        if isinstance(x, DeviceVector):
            assert isinstance(y, DeviceVector) and x.n >= self.n_in and y.n >= self.n_out
            L.check(L.lib().hcurl_interp_apply(self.h, C.c_void_p(x.ptr), C.c_void_p(y.ptr), L.DEVICE), "hcurl_interp_apply")
            return y
        x = L.as_f64(x)
        if x.size != self.n_in:
            raise L.HcurlError("Mult: x has %d entries, interpolator width is %d" % (x.size, self.n_in))
        y = np.empty(self.n_out) if y is None else y
        L.check(L.lib().hcurl_interp_apply(self.h, L.ptr(x), L.ptr(y), L.HOST), "hcurl_interp_apply")
        return y

    def close(self):
        if self.h:
            L.lib().hcurl_interp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class HephGpuPCG:
    """hypre-semantics PCG (+ Jacobi), an ``mfem::Solver``: SetOperator / Mult.
    Defaults are DefaultJacobiPCGSolver's
    (/root/reference/src/solvers/hephaestus_solvers.hpp:39-42); tolerances go
    through C `float` like the reference's InputParameters."""

    def __init__(self, ctx, tol=1.0e-9, abstol=1e-16, max_iter=1000, print_level=0, jacobi=True, chebyshev_order=0):
        """chebyshev_order > 0: Jacobi-scaled Chebyshev polynomial preconditioner of that order (HCURL_PC_CHEBYSHEV)"""
        self.ctx = ctx
        self.chebyshev_order = int(chebyshev_order)
        self.tol = float(np.float32(tol))
        self.abstol = float(np.float32(abstol))
        self.max_iter, self.print_level = int(max_iter), int(print_level)
        self.precond = L.PC_CHEBYSHEV if self.chebyshev_order > 0 else (L.PC_JACOBI if jacobi else L.PC_NONE)
        self.iterative_mode = False  # NewtonSolver sets this (zero initial guess)
        self.h = None
        self.op = None
        self.num_iterations, self.final_rel_res, self.converged = 0, 0.0, False

    def SetOperator(self, op):
        if not isinstance(op, HephGpuOperator):
            raise L.HcurlError("HephGpuPCG::SetOperator needs a HephGpuOperator (cf. HyprePCG needing a HypreParMatrix)")
        if self.h and self.op is op:
            L.check(L.lib().hcurl_pcg_update(self.h), "hcurl_pcg_update")
            return
        self.close()
        self.op = op
        self.h = L.check_ptr(L.lib().hcurl_pcg_create(self.ctx.h, op.h, self.precond), "hcurl_pcg_create")
        L.check(L.lib().hcurl_pcg_set(self.h, self.tol, self.abstol, self.max_iter, self.print_level), "hcurl_pcg_set")
        if self.chebyshev_order > 0:
            L.check(L.lib().hcurl_pcg_set_chebyshev(self.h, self.chebyshev_order, 10), "hcurl_pcg_set_chebyshev")

    def Mult(self, b, x=None):
        if self.h is None:
            raise L.HcurlError("HephGpuPCG::Mult called before SetOperator")
        it, rr, cv = C.c_int(0), C.c_double(0.0), C.c_int(0)
        if isinstance(b, DeviceVector):
            assert isinstance(x, DeviceVector)
            L.check(L.lib().hcurl_pcg_solve(self.h, C.c_void_p(b.ptr), C.c_void_p(x.ptr), int(not self.iterative_mode), L.DEVICE,
                                            C.byref(it), C.byref(rr), C.byref(cv)), "hcurl_pcg_solve")
        else:
            b = L.as_f64(b)
            if x is None:
                x = np.zeros_like(b)
            assert x.dtype == np.float64 and x.flags.c_contiguous
            L.check(L.lib().hcurl_pcg_solve(self.h, L.ptr(b), L.ptr(x), int(not self.iterative_mode), L.HOST,
                                            C.byref(it), C.byref(rr), C.byref(cv)), "hcurl_pcg_solve")
        self.num_iterations, self.final_rel_res, self.converged = it.value, rr.value, bool(cv.value)
        return x

    def last_loop_ms(self):
        return L.lib().hcurl_pcg_last_loop_ms(self.h)

    def last_used_graph(self):
        return bool(L.lib().hcurl_pcg_last_used_graph(self.h))

    def lambda_max(self):
        return L.lib().hcurl_pcg_lambda_max(self.h)

    def close(self):
        if self.h:
            L.lib().hcurl_pcg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
