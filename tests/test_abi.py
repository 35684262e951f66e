"""CPU: the C-ABI library loads and exports exactly what include/satfill.h declares (no compute calls)."""
from __future__ import annotations

import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "satfill.h")


def header_symbols() -> list[str]:
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sa_[a-z0-9_]+)\s*\(", text)))


def test_header_matches_binding_list():
    from satellite_approximation_b200 import _capi

    assert header_symbols() == sorted(_capi.EXPORTS)


def test_library_exports_every_declared_symbol():
    from satellite_approximation_b200 import _capi

    if not os.path.exists(_capi.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    lib = ctypes.CDLL(_capi.LIB_PATH)
    for name in header_symbols():
        assert hasattr(lib, name), f"{name} declared in include/satfill.h but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", _capi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    assert set(header_symbols()) <= exported
    assert lib.sa_abi_version() == 5


def test_struct_layouts_match_header():
    from satellite_approximation_b200 import _capi

    assert ctypes.sizeof(_capi.Options) == 8 + 8 + 4 * 4 + 16
    assert ctypes.sizeof(_capi.Stats) == 8 * 3 + 8 * 4 + 4 * 2 + 3 * 8 * 8
    lib = _capi.load()
    o = _capi.Options()
    lib.sa_default_options(ctypes.byref(o), _capi.SA_POISSON)
    assert o.tolerance == 1e-6 and o.max_iterations == 0 and o.precond == _capi.SA_PRECOND_MULTIGRID  # the drop-in call takes the fast path
    lib.sa_default_options(ctypes.byref(o), _capi.SA_LAPLACE)
    assert o.tolerance == 2.220446049250313e-16  # Eigen default (IterativeSolverBase.h:367-368)


def test_no_device_fails_loudly():
    """Without a GPU the product must refuse to run rather than fall back to anything on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import satellite_approximation_b200 as sab

    with pytest.raises(sab.SatfillError):
        sab.Context(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "satellite_approximation_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "import oracle" not in text and "liboracle" not in text and "libref_eigen" not in text, fn


def _build_c_example(tmp_path, name="fill_c_abi"):
    from satellite_approximation_b200 import _capi

    libdir = os.path.dirname(_capi.LIB_PATH)
    exe = str(tmp_path / name)
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", name + ".c"), "-L", libdir, "-lsatfill", f"-Wl,-rpath,{libdir}", "-lm", "-o", exe]  # fmt: skip
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_header_is_plain_c_and_the_example_fails_loudly_without_a_gpu(tmp_path):
    """include/satfill.h compiles as pedantic C99 (and as C++), and a plain-C caller of the C-ABI gets a clean error --
    not a CPU fallback -- on a machine without a device."""
    import torch

    exe = _build_c_example(tmp_path)
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-x", "c++", HEADER], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: test_c_example_on_gpu runs it")
    for exe in (exe, _build_c_example(tmp_path, "blend_c_abi")):
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 2 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_c_example_on_gpu(tmp_path):
    """examples/fill_c_abi.c: sa_laplace_fill from plain C on a column-major image reproduces a discrete-harmonic field."""
    exe = _build_c_example(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "max error" in r.stdout


@pytest.mark.gpu
def test_c_blend_example_on_gpu(tmp_path):
    """examples/blend_c_abi.c: sa_unknown_numbering, sa_poisson_blend (replacement = input + constant returns the input)
    and sa_label_components (the reference's own test case, tests/approximation.h:55-75) from plain C."""
    exe = _build_c_example(tmp_path, "blend_c_abi")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "components: 2 labels, sizes 4 and 8" in r.stdout


def test_c_examples_logic_on_the_cpu(tmp_path):
    """examples/*.c run to completion on the CPU against tests/fake_satfill.c (the C-ABI answered by the ORACLE, put in
    front of the real library through LD_LIBRARY_PATH): their own arithmetic and known answers are right, independently
    of the device run (test_c_example_on_gpu, test_c_blend_example_on_gpu)."""
    import oracle

    oracle.port()
    odir = os.path.join(ROOT, "oracle", "_build")
    fake = tmp_path / "fake"
    fake.mkdir()
    r = subprocess.run(["gcc", "-O2", "-std=c99", "-Wall", "-Wextra", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "fake_satfill.c"), "-o", str(fake / "libsatfill.so"), "-L", odir,
                        "-loracle", f"-Wl,-rpath,{odir}"], capture_output=True, text=True)  # fmt: skip
    assert r.returncode == 0, r.stderr
    env = dict(os.environ, LD_LIBRARY_PATH=str(fake) + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([_build_c_example(tmp_path)], capture_output=True, text=True, env=env)
    assert r.returncode == 0 and "max error" in r.stdout, r.stdout + r.stderr
    r = subprocess.run([_build_c_example(tmp_path, "blend_c_abi")], capture_output=True, text=True, env=env)
    assert r.returncode == 0 and "components: 2 labels, sizes 4 and 8" in r.stdout, r.stdout + r.stderr
