"""First-run harness for GPU code written after the round-1 GPU budget was spent (runs last, cannot mask other tests).

The order-1 patch kernel (hex_patch_p1.cu), the split-blob order-2 variant (hex_patch_split_p2_kernel) and the discrete
curl / gradient kernel (interp.cu) are verified on the host (tests/test_patch_p1_emu.py, tests/test_interp.py) or
compile-checked only (split variant), but have never executed on a device.  Their GPU parity tests are opt-in
(HCURL_TEST_PATCH_P1 / HCURL_TEST_SPLIT / HCURL_TEST_INTERP).  Each test below runs one of those groups in a CHILD process
with the switch set and a hard time limit -- a kernel that hangs or faults takes the child down, not this run -- and is
marked xfail(strict=False): XPASS in the summary = the kernel is parity-green on this GPU, XFAIL = it is not yet (the
child's output is printed).  Neither outcome turns the suite red; the default paths never use these kernels."""
import os
import subprocess
import sys
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


# the whole harness is bounded too: once the children have used BUDGET_S seconds together, the remaining groups are
# skipped (reported as XFAIL) -- a series of hanging first runs must not eat the time limit of the GPU test run itself
BUDGET_S = float(os.environ.get("HCURL_FIRST_RUN_BUDGET_S", "420"))
_spent = [0.0]


def _child(env_switch, args, limit=150):
    left = BUDGET_S - _spent[0]
    if left < 20:
        pytest.fail("first-run budget of %.0f s used up by the earlier groups; %s not run" % (BUDGET_S, env_switch))
    limit = min(limit, left)
    t0 = time.monotonic()
    try:
        _run_child(env_switch, args, limit)
    finally:
        _spent[0] += time.monotonic() - t0


def _run_child(env_switch, args, limit):
    env = dict(os.environ)
    env[env_switch] = "1"
    cmd = [sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-p", "no:cacheprovider"] + args
    try:
        r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=limit)
    except subprocess.TimeoutExpired as e:
        out = (e.stdout or b"").decode(errors="replace") if isinstance(e.stdout, bytes) else (e.stdout or "")
        pytest.fail("child exceeded %d s (hang?)\n%s" % (limit, out[-3000:]))
    print(r.stdout[-3000:])
    print(r.stderr[-2000:])
    if r.returncode != 0:
        key = [ln.strip() for ln in r.stdout.splitlines() if ln.startswith(("E  ", "FAILED", "ERROR")) or "Error" in ln][:6]
        pytest.fail("child failed:\n" + r.stdout[-3000:] + "\nKEY LINES:\n" + "\n".join(ln[:200] for ln in key), pytrace=False)
    assert " passed" in r.stdout and "skipped" not in r.stdout.splitlines()[-1], r.stdout[-500:]


@pytest.mark.xfail(strict=False, reason="order-2 patch kernel built from the Env program (plain + split-blob): first execution on a GPU")
def test_first_run_order2_env_program():
    _child("HCURL_TEST_PATCH_ENV", ["tests/test_gpu_apply.py", "-k", "env_program"])


@pytest.mark.xfail(strict=False, reason="two-threads-per-component order-2 patch kernel (96 / 80 registers): first execution on a GPU")
def test_first_run_order2_half_program():
    _child("HCURL_TEST_PATCH_HALF", ["tests/test_gpu_apply.py", "-k", "half_program"])


@pytest.mark.xfail(strict=False, reason="split-blob order-2 patch kernel variant: first execution on a GPU")
def test_first_run_split_blob_variant():
    _child("HCURL_TEST_SPLIT", ["tests/test_gpu_apply.py", "-k", "split"])


@pytest.mark.xfail(strict=False, reason="order-1 patch kernel: first execution on a GPU")
def test_first_run_order1_patch_kernel():
    _child("HCURL_TEST_PATCH_P1", ["tests/test_gpu_apply.py", "tests/test_gpu_pcg.py", "-k", "p1_patch or order1_patch"])


@pytest.mark.xfail(strict=False, reason="patch kernel for single-term operators (statics, mass solves): first execution on a GPU")
def test_first_run_single_term_patch_path():
    _child("HCURL_TEST_SINGLE_TERM", ["tests/test_gpu_apply.py", "-k", "single_term"])


@pytest.mark.xfail(strict=False, reason="HCURL_QUAD_AUTO: first execution on a GPU")
def test_first_run_quad_auto():
    _child("HCURL_TEST_QAUTO", ["tests/test_gpu_apply.py", "-k", "quad_auto"])


@pytest.mark.xfail(strict=False, reason="Chebyshev-preconditioned PCG: first execution on a GPU")
def test_first_run_chebyshev_pcg():
    _child("HCURL_TEST_CHEBYSHEV", ["tests/test_gpu_pcg.py", "-k", "chebyshev"])


@pytest.mark.xfail(strict=False, reason="discrete curl / gradient kernel: first execution on a GPU")
def test_first_run_interpolators():
    _child("HCURL_TEST_INTERP", ["tests/test_interp.py"])


@pytest.mark.xfail(strict=False, reason="GPU operator on the iris-waveguide KAT mesh: first execution on a GPU")
def test_first_run_iris_kat_operator():
    _child("HCURL_TEST_IRIS", ["tests/test_oracle_iris_kat.py"])
