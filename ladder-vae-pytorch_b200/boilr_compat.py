"""The five boilr symbols the model needs (models/lvae.py:3-4 of the reference).

If the real ``boilr`` package is importable it is used (so ``main.py`` / ``evaluate.py`` get
boilr's own ``BaseGenerativeModel`` with checkpointing); otherwise these stand-ins provide the
same behaviour.  PARITY UNPINNED: boilr==0.7.4 is not vendored in the reference tree, the
stand-ins restate its documented behaviour.  pad/crop/Interpolate/free-bits are
re-implemented on our kernels either way (ops.pad_image / ops.crop / ops.upsample2x).
"""
from __future__ import annotations

import torch
from torch import nn

try:  # pragma: no cover - boilr is not installed in the build image
    from boilr.models import BaseGenerativeModel  # type: ignore
    HAVE_BOILR = True
except Exception:  # noqa: BLE001
    HAVE_BOILR = False

    class BaseGenerativeModel(nn.Module):
        """global_step bookkeeping + checkpoint/load, as boilr's base class offers."""

        def __init__(self):
            super().__init__()
            self.global_step = 0

        def increment_global_step(self):
            self.global_step += 1

        def get_device(self):
            return next(self.parameters()).device

        def checkpoint(self, ckpt_folder, max_ckpt=None):
            import os
            os.makedirs(ckpt_folder, exist_ok=True)
            path = os.path.join(ckpt_folder, "model_{}.pt".format(self.global_step))
            torch.save(self.state_dict(), path)
            if max_ckpt:
                files = sorted((f for f in os.listdir(ckpt_folder) if f.startswith("model_") and f.endswith(".pt")),
                               key=lambda f: int(f[6:-3]))
                for f in files[:-max_ckpt]:
                    os.remove(os.path.join(ckpt_folder, f))
            return path

        def load(self, ckpt_folder, device=None, step=None):
            import os
            files = sorted((f for f in os.listdir(ckpt_folder) if f.startswith("model_") and f.endswith(".pt")),
                           key=lambda f: int(f[6:-3]))
            name = "model_{}.pt".format(step) if step is not None else files[-1]
            self.load_state_dict(torch.load(os.path.join(ckpt_folder, name), map_location=device))
            self.global_step = int(name[6:-3])


def free_bits_kl(kl, free_bits, batch_average=False, eps=1e-6):
    """boilr.nn.free_bits_kl: kl is (batch, layers); returns (layers,)."""
    assert kl.dim() == 2
    if free_bits < eps:
        return kl.mean(0)
    if batch_average:
        return kl.mean(0).clamp(min=free_bits)
    return kl.clamp(min=free_bits).mean(0)
