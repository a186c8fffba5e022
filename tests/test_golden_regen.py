"""The committed golden fixtures are exactly what tests/golden/make_golden.py produces from the
reference's own optimization.py (only checkable where /root/reference exists: the authoring container)."""
import importlib.util
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = "/root/reference/optimization.py"


@pytest.mark.skipif(not os.path.exists(REFERENCE), reason="/root/reference is not present (GPU box)")
def test_fixtures_regenerate_bit_identically(tmp_path, monkeypatch):
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    monkeypatch.setattr(mg, "HERE", str(tmp_path))
    saved_path, saved_mods = list(sys.path), {k: sys.modules.get(k) for k in ("tensorflow", "optimization")}
    try:
        for name, cfg in mg.CASES.items():
            path, steps = mg.run_case(name, *cfg)
            new = np.load(path)
            old = np.load(os.path.join(HERE, "golden", os.path.basename(path)))
            assert set(new.files) == set(old.files)
            for k in new.files:
                assert np.array_equal(new[k], old[k]), f"{name}: {k} differs from the committed fixture"
    finally:
        sys.path[:] = saved_path
        for k, v in saved_mods.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
