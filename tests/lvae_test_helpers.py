"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import json
import os

import numpy as np
import torch

from oracle import lvae_oracle as O
from oracle.make_golden import make_inputs, cases, GOLDEN_DIR  # noqa: F401


def load_golden(name):
    blob = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    meta = json.loads(str(blob["meta"]))
    cfg = O.LVAEConfig(**meta["cfg"])
    return cfg, meta, blob


def rel_err(a, b, floor=0.0):
    a = torch.as_tensor(a.detach() if isinstance(a, torch.Tensor) else a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b.detach() if isinstance(b, torch.Tensor) else b, dtype=torch.float64).cpu()
    return float((a - b).abs().max() / (b.abs().max() + floor + 1e-300))


def golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz"))
