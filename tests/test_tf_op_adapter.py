"""tf_shim/gaccum_tf_op.cc -- the C++ half of the TensorFlow binding (the boundary a maintainer of the
reference adds) -- compiled UNMODIFIED with plain g++ against tests/tf_mock (a mock of the slice of TensorFlow's op-kernel
API it uses: TensorFlow's headers are not in this image) and linked with csrc/libgaccum.so.  tests/tf_mock/tf_op_driver.cc
plays the executor: kernel looked up in the registry the adapter's REGISTER_* statements filled, inputs in the order of the
adapter's own OpDef, device memory unless the registration says HostMemory, one Compute() per micro-step.

CPU: the adapter compiles and links, its registration matches what tf_shim/optimization.py passes, its error paths work and it
fails loudly without a CUDA device.  GPU: both registrations (ref variables / resource variables) reproduce the
reference-produced fixtures through Compute() -> gaccum_step -> the CUDA kernels.
What a mock cannot prove is said in tests/tf_mock/tensorflow/core/framework/op_kernel.h."""
import os
import re
import shutil
import subprocess
import sys

import numpy as np
import pytest

from golden_util import Golden, RecipeGolden
from test_abi_consumer import _export, _read

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")
SHIM = os.path.join(ROOT, "gradient-accumulation-tf-estimator_b200", "tf_shim")


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    import gaccum_b200 as g
    gxx = shutil.which("g++") or shutil.which("c++") or pytest.skip("no g++ in this image")
    lib = g.lib_path()
    out = str(tmp_path_factory.mktemp("tfop") / "tf_op_driver")
    cmd = [gxx, "-std=c++17", "-O1", "-Wall", "-Wno-comment", "-Werror",
           "-I", os.path.join(ROOT, "tests", "tf_mock"), "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"),
           os.path.join(SHIM, "gaccum_tf_op.cc"), os.path.join(ROOT, "tests", "tf_mock", "tf_op_driver.cc"), "-o", out,
           lib, "-L", os.path.join(CUDA, "lib64"), "-lcudart", "-Wl,-rpath," + os.path.dirname(lib), "-Wl,-rpath," + os.path.join(CUDA, "lib64")]
    env = dict(os.environ); env.pop("CC", None); env.pop("CXX", None)
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    return out


def _registry(driver):
    r = subprocess.run([driver, "--registry"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    ops, cur = {}, None
    for line in r.stdout.splitlines():
        if line.startswith("op "):
            cur = ops.setdefault(line.split()[1], {"inputs": [], "attrs": [], "host": [], "stateful": "stateful=1" in line})
        elif line.startswith("  input "):
            cur["inputs"].append(line[len("  input "):])
        elif line.startswith("  attr "):
            cur["attrs"].append(line[len("  attr "):])
        elif line.startswith("  kernel "):
            assert "device=GPU" in line
            cur["host"] = [h for h in line.split("host_memory=")[1].split(",") if h]
    return ops


def test_adapter_compiles_and_registers_both_ops(driver):
    ops = _registry(driver)
    assert set(ops) == {"GaccumStep", "GaccumStepV2"}
    for name, op in ops.items():
        assert op["stateful"], "an op that mutates variables and has no outputs must be stateful or grappler prunes it"
        names = [i.split(":")[0] for i in op["inputs"]]
        assert names == ["params", "accum", "m", "v", "grads", "global_step", "lr", "beta_powers"]
        # the accumulate / apply decision is taken on the host: no D2H of the step counter
        assert {"global_step", "lr", "beta_powers"} <= set(op["host"])
    assert all("Ref(" in i for i in ops["GaccumStep"]["inputs"][:4])
    assert all(i.endswith("resource") for i in ops["GaccumStepV2"]["inputs"][:4])
    # resource HANDLES live in host memory (as for TF's own ResourceApplyAdam GPU kernel)
    assert {"params", "accum", "m", "v"} <= set(ops["GaccumStepV2"]["host"])


def test_python_shim_passes_exactly_what_the_cc_registration_declares(driver):
    """The keyword arguments tf_shim/optimization.py hands the generated op wrapper == the adapter's inputs + attrs
    (``N`` is inferred by TensorFlow from the list length), and hyper-parameters are declared as string attrs."""
    ops = _registry(driver)
    stub = os.path.join(ROOT, "oracle", "tf_stub")
    code = f"""
import importlib.util, sys, json
sys.path.insert(0, {stub!r})
import tensorflow as tf, numpy as np
spec = importlib.util.spec_from_file_location('shim_opt', {os.path.join(SHIM, 'optimization.py')!r})
mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
seen = []
lib = mod._load()
real = type(lib).gaccum_step
def spy(**kw):
    seen.append({{k: type(v).__name__ for k, v in kw.items()}})
    return real(**kw)
type(lib).gaccum_step = staticmethod(spy)
tf.get_variable('w/kernel', shape=[4, 3], dtype=tf.float32, initializer=np.ones((4, 3), np.float32))
mod.create_optimizer(tf.constant(0.0), 1e-3, 100, 10, False)
print(json.dumps(seen))
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    import json
    seen = json.loads(r.stdout.strip().splitlines()[-1])
    assert len(seen) == 1
    op = ops["GaccumStep"]
    declared = [i.split(":")[0] for i in op["inputs"]] + [a.split(":")[0] for a in op["attrs"]]
    assert set(seen[0]) == set(declared) - {"N"}
    for a in op["attrs"]:
        n, t = [x.strip() for x in a.split(":", 1)]
        if n in ("beta1", "beta2", "epsilon", "weight_decay_rate", "clip_norm"):
            assert t.startswith("string"), a                    # a float attr would be fp32 in the GraphDef
            assert seen[0][n] == "str"
            default = re.search(r"'([^']*)'", t).group(1)
            assert repr(float(default)) == default               # defaults are reprs too


def test_adapter_error_paths_and_no_cpu_fallback(driver):
    r = subprocess.run([driver, "--errors"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "error paths ok" in r.stdout
    import torch
    if not torch.cuda.is_available():
        assert "no CPU fallback" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("op", ["GaccumStep", "GaccumStepV2"])
@pytest.mark.parametrize("case", ["warmup_unclipped", "n4_clipped", "n3_warmup"])
def test_adapter_compute_reproduces_reference_fixture_variant_a(driver, tmp_path, op, case):
    gd = Golden(case)
    shapes = [gd.z[f"init/{n}"].shape for n in gd.names]
    inp, outp = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    _export(inp, gd.names, shapes, gd.N, gd.steps, 0, gd.init_lr, gd.num_train_steps, gd.num_warmup_steps, 1.0, gd.init(), gd.grads)
    r = subprocess.run([driver, op, inp, outp], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "warning" not in r.stderr
    exact = case == "warmup_unclipped"            # otherwise the clip scale's norm is summed in a different order than the fixture's
    for s, (hdr, st) in enumerate(_read(outp, shapes, gd.steps)):
        for i, n in enumerate(gd.names):
            for k, kind in enumerate(("param", "accum", "m", "v")):
                exp = gd.z[f"{kind}/{s}/{n}"]
                if exact or kind == "accum":
                    assert np.array_equal(st[i][k], exp), f"{op} {case} step {s} {kind} {n}"
                else:
                    assert np.allclose(st[i][k], exp, rtol=1e-5, atol=1e-8), f"{op} {case} step {s} {kind} {n}"


@pytest.mark.gpu
@pytest.mark.parametrize("op", ["GaccumStep", "GaccumStepV2"])
def test_adapter_compute_reproduces_reference_recipe_variant_b(driver, tmp_path, op):
    gd = RecipeGolden("another_example_n3")
    inp, outp = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    _export(inp, gd.names, gd.shapes, gd.N, gd.steps, 1, gd.lr, 1, 0, 0.0, gd.init(), gd.grads)
    r = subprocess.run([driver, op, inp, outp], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    for s, (hdr, st) in enumerate(_read(outp, gd.shapes, gd.steps)):
        assert bool(hdr[3]) == (s % gd.N == 0)
        for i, n in enumerate(gd.names):
            for k, kind in enumerate(("param", "accum", "m", "v")):
                gd.check(f"{kind}/{s}/{n}", st[i][k])        # variant B has no reduction: bit-identical
