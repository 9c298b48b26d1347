"""CPU: the C-ABI shared library loads without a GPU and exports every symbol include/gen_b200.h
declares; calls that need a device fail loudly instead of falling back to a CPU path."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import gen_b200
from gen_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if build.needs_build():
        build.build()
    return gen_b200.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gen_b200.h")).read()
    return sorted(set(re.findall(r"GSMC_API\s+[\w\s\*]+?\b(gsmc_\w+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(lib):
    names = declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), "libgensmc.so does not export %s" % name
        assert name in _lib.SIGNATURES, "gen_b200/_lib.py does not bind %s" % name
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_config_layout(lib):
    assert b"sm_100a" in lib.gsmc_version()
    assert C.sizeof(_lib.Config) == 56          # matches struct gsmc_config in include/gen_b200.h


def test_only_sm100a_code_is_built():
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="GPU present: covered by the gpu tests")
def test_no_cpu_fallback(lib):
    """Without a device the product refuses to compute (it must never route through a CPU path)."""
    model = gen_b200.LinearGaussianSSM()
    with pytest.raises(gen_b200.GsmcError) as ei:
        gen_b200.initialize_particle_filter(model, (1,), gen_b200.choicemap(("y_init", 0.1)), 128)
    assert ei.value.code == _lib.E_CUDA
    with pytest.raises(gen_b200.GsmcError):
        gen_b200.importance_sampling(gen_b200.NormalNormal(), (), gen_b200.choicemap(("y", 2.0)), 4)


def test_argument_validation_happens_before_the_device(lib):
    cfg = _lib.Config()
    cfg.struct_size = C.sizeof(_lib.Config)
    cfg.model_id = 99
    cfg.num_particles = 10
    h = C.c_void_p()
    p = np.zeros(3)
    assert lib.gsmc_create(C.byref(cfg), _lib.dptr(p), 3, C.byref(h)) == _lib.E_UNSUPPORTED
    cfg.model_id = _lib.MODEL_LGSSM
    assert lib.gsmc_create(C.byref(cfg), _lib.dptr(p), 3, C.byref(h)) == _lib.E_BADARG      # needs 7 params
    assert b"LGSSM" in lib.gsmc_last_error(None)
    cfg.struct_size = 8
    assert lib.gsmc_create(C.byref(cfg), _lib.dptr(p), 3, C.byref(h)) == _lib.E_BADARG
    assert lib.gsmc_step(None, None, 0, 0, None, 0) == _lib.E_BADARG


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under gen_b200/ may import, include, link or call it."""
    bad = re.compile(r"(^\s*(import|from)\s+oracle\b)|(#\s*include\s*[\"<][^\">]*oracle)|liboracle|\borc_\w+\s*\(", re.M)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gen_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not bad.search(text), "%s uses the oracle" % f


def _build_c_smoke(tmp_path):
    import subprocess
    exe = str(tmp_path / "c_abi_smoke")
    libdir = os.path.dirname(_lib.LIB_PATH)
    cmd = ["gcc", "-std=c11", "-Wall", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c_abi_smoke.c"),
           "-o", exe, "-L", libdir, "-l:libgensmc.so", "-Wl,-rpath," + libdir]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def test_c_caller_compiles_and_links_against_the_header(lib, tmp_path):
    """A plain-C translation unit includes include/gen_b200.h (no C++-isms in the header) and links every entry point it uses."""
    _build_c_smoke(tmp_path)


@pytest.mark.gpu
def test_c_caller_runs_the_filter(lib, tmp_path):
    """The same loop from C and through the ctypes mirror: identical bits (struct layout, argument order, call order)."""
    import subprocess
    exe = _build_c_smoke(tmp_path)
    N, T = 10000, 12
    res = subprocess.run([exe, str(N), str(tmp_path / "c.ckpt")], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    m = re.search(r"resamples=(\d+) log_ml=(\S+) lw_sum=(\S+) x_sum=(\S+) anc_sum=(-?\d+)", res.stdout)
    assert m, res.stdout
    ys = [0.3 * t - 1.0 + ((t * 7) % 5) * 0.21 for t in range(T)]
    st = gen_b200.ParticleFilterState(gen_b200.LinearGaussianSSM(0.0, 1.0, 0.9, 0.0, 1.0, 1.0, 1.0), N, seed=42, keep_history=True, history_capacity=T)
    st.init([ys[0]])
    n_res = 0
    for t in range(1, T):
        n_res += st.maybe_resample(0.5 * N)
        st.step([ys[t]])
    assert n_res == int(m.group(1)) and n_res > 0
    assert st.log_ml_estimate() == float(m.group(2))
    lw, x, anc = st.log_weights(), st.state()[0], st.ancestors()
    s1 = s2 = 0.0
    for i in range(N):          # same summation order as the C loop
        s1 += lw[i]
        s2 += x[i]
    assert s1 == float(m.group(3)) and s2 == float(m.group(4)) and int(anc.sum()) == int(m.group(5))
    st.close()
