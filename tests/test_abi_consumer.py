"""The C ABI exercised by a NON-Python consumer: tests/abi_consumer.cc includes include/gaccum.h, is built by plain
g++ -std=c++17 (no nvcc, no torch) against csrc/libgaccum.so, and drives plan -> gaccum_step on a non-default stream
with cudaMalloc'd scattered tensors exactly as the TensorFlow op of INTEGRATION.md would.  Its output is compared
with the golden fixtures produced by the reference's own code (tests/golden/)."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

from golden_util import Golden, RecipeGolden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _gxx():
    for c in ("g++", "c++"):
        p = shutil.which(c)
        if p:
            return p
    pytest.skip("no g++ in this image")


@pytest.fixture(scope="module")
def consumer(tmp_path_factory):
    import gaccum_b200 as g
    lib = g.lib_path()
    out = str(tmp_path_factory.mktemp("abi") / "abi_consumer")
    cmd = [_gxx(), "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"),
           os.path.join(ROOT, "tests", "abi_consumer.cc"), "-o", out,
           lib, "-L", os.path.join(CUDA, "lib64"), "-lcudart", "-Wl,-rpath," + os.path.dirname(lib), "-Wl,-rpath," + os.path.join(CUDA, "lib64")]
    env = dict(os.environ); env.pop("CC", None); env.pop("CXX", None)
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    return out


def test_header_is_plain_c(tmp_path):
    """include/gaccum.h must be consumable from C (the TF C API / cgo / JNI side): gcc -std=c99 -pedantic."""
    gcc = shutil.which("gcc") or pytest.skip("no gcc")
    src = tmp_path / "c.c"
    src.write_text('#include "gaccum.h"\nint main(void) { gaccum_hparams hp; gaccum_step_args a; (void)hp; (void)a; return GACCUM_VERSION > 0 ? 0 : 1; }\n')
    env = dict(os.environ); env.pop("CC", None)
    r = subprocess.run([gcc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                       capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr


def test_cpp_consumer_links_and_host_logic_runs_without_a_gpu(consumer):
    r = subprocess.run([consumer, "--host-only"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host-only ok" in r.stdout


def _export(path, names, shapes, N, steps, variant, init_lr, train_steps, warmup_steps, clip, init, grads_fn):
    with open(path, "wb") as f:
        f.write(struct.pack("<iiiidqqd", len(names), N, steps, variant, init_lr, train_steps, warmup_steps, clip))
        for n, shp in zip(names, shapes):
            nb = n.encode()
            f.write(struct.pack("<qi", int(np.prod(shp)), len(nb))); f.write(nb)
        for a in init:
            f.write(np.ascontiguousarray(a, np.float32).tobytes())
        for s in range(steps):
            for a in grads_fn(s):
                f.write(np.ascontiguousarray(a, np.float32).tobytes())


def _read(path, shapes, steps):
    raw = np.fromfile(path, dtype=np.float32)
    o, out = 0, []
    for s in range(steps):
        hdr = raw[o:o + 4]; o += 4
        st = []
        for shp in shapes:
            n = int(np.prod(shp))
            st.append([raw[o + k * n:o + (k + 1) * n].reshape(shp) for k in range(4)]); o += 4 * n
        out.append((hdr, st))
    assert o == raw.size
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["warmup_unclipped", "n4_clipped", "n3_warmup"])
def test_cpp_consumer_reproduces_reference_fixture_variant_a(consumer, tmp_path, case):
    gd = Golden(case)
    shapes = [gd.z[f"init/{n}"].shape for n in gd.names]
    inp, outp = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    _export(inp, gd.names, shapes, gd.N, gd.steps, 0, gd.init_lr, gd.num_train_steps, gd.num_warmup_steps, 1.0, gd.init(), gd.grads)
    r = subprocess.run([consumer, inp, outp], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    exact = True
    for s, (hdr, st) in enumerate(_read(outp, shapes, gd.steps)):
        if hdr[3] and hdr[1] != 1.0:
            exact = False                     # the clip scale's norm is summed in a different order than the fixture's
        for i, n in enumerate(gd.names):
            for k, kind in enumerate(("param", "accum", "m", "v")):
                exp = gd.z[f"{kind}/{s}/{n}"]
                if exact or kind == "accum":
                    assert np.array_equal(st[i][k], exp), f"{case} step {s} {kind} {n}"
                else:
                    assert np.allclose(st[i][k], exp, rtol=1e-5, atol=1e-8), f"{case} step {s} {kind} {n}"
    if case == "warmup_unclipped":
        assert exact


@pytest.mark.gpu
def test_cpp_consumer_reproduces_reference_recipe_variant_b(consumer, tmp_path):
    gd = RecipeGolden("another_example_n3")
    inp, outp = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    _export(inp, gd.names, gd.shapes, gd.N, gd.steps, 1, gd.lr, 1, 0, 0.0, gd.init(), gd.grads)
    r = subprocess.run([consumer, inp, outp], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    for s, (hdr, st) in enumerate(_read(outp, gd.shapes, gd.steps)):
        assert bool(hdr[3]) == (s % gd.N == 0)
        for i, n in enumerate(gd.names):
            for k, kind in enumerate(("param", "accum", "m", "v")):
                gd.check(f"{kind}/{s}/{n}", st[i][k])        # variant B has no reduction: bit-identical
