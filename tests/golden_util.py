"""Loader for tests/golden/ref_optimization_*.npz (made by tests/golden/make_golden.py)."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def cases():
    return sorted(os.path.basename(p)[len("ref_optimization_"):-4]
                  for p in glob.glob(os.path.join(GOLDEN_DIR, "ref_optimization_*.npz")))


class Golden:
    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN_DIR, f"ref_optimization_{name}.npz"))
        self.names = [str(n) for n in self.z["names"]]
        self.N = int(self.z["N"])
        self.steps = int(self.z["steps"])
        self.init_lr = float(self.z["init_lr"])
        self.num_train_steps = int(self.z["num_train_steps"])
        self.num_warmup_steps = int(self.z["num_warmup_steps"])

    def init(self):
        return [self.z[f"init/{n}"].copy() for n in self.names]

    def grads(self, s):
        return [self.z[f"grad/{s}/{n}"] for n in self.names]

    def state(self, s, kind):
        """kind in param|accum|m|v; m/v are None before the first apply created them (step 0 applies)."""
        return [self.z[f"{kind}/{s}/{n}"] for n in self.names]

    def global_step(self, s):
        return int(self.z[f"global_step/{s}"])
