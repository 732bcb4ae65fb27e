"""TEST INFRASTRUCTURE ONLY -- imports the *unmodified* reference from /root/reference.

The reference (addtt/ladder-vae-pytorch) depends on two packages that are not
installed here and cannot be (no network): ``boilr==0.7.4`` (requirements.txt:6)
and ``matplotlib`` (lib/likelihoods.py:3, used only under ``__main__``).  This
module puts minimal stand-ins for the five boilr symbols the model imports
(models/lvae.py:3-4) into ``sys.modules`` and then imports ``models.lvae``
from /root/reference unchanged.

PARITY UNPINNED for the boilr pieces: boilr's source is not under
/root/reference, so ``free_bits_kl``, ``pad_img_tensor``, ``crop_img_tensor``
and ``Interpolate`` below are restated from the package's documented
behaviour, not checked against its code.

/root/reference does not exist on the GPU box: only ``oracle/make_golden.py``
and CPU tests guarded by ``reference_available()`` may call into this file.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("LVAE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "lvae.py"))


def _boilr_stubs():
    import torch
    from torch import nn
    import torch.nn.functional as F

    class BaseGenerativeModel(nn.Module):
        """boilr.models.BaseGenerativeModel stand-in: global_step bookkeeping."""

        def __init__(self):
            super().__init__()
            self.global_step = 0

        def increment_global_step(self):
            self.global_step += 1

        def get_device(self):
            return next(self.parameters()).device

    def _pad_crop(x, size, mode):
        assert x.dim() == 4 and len(size) == 2
        size = tuple(int(s) for s in size)
        cur = tuple(x.shape[2:4])
        dr, dc = abs(cur[0] - size[0]), abs(cur[1] - size[1])
        dr1, dr2 = dr // 2, dr - dr // 2
        dc1, dc2 = dc // 2, dc - dc // 2
        if mode == "pad":
            if cur[0] > size[0] or cur[1] > size[1]:
                raise ValueError("trying to pad to a smaller size")
            return F.pad(x, [dc1, dc2, dr1, dr2, 0, 0, 0, 0])
        if cur[0] < size[0] or cur[1] < size[1]:
            raise ValueError("trying to crop to a larger size")
        return x[:, :, dr1:cur[0] - dr2, dc1:cur[1] - dc2]

    def pad_img_tensor(x, size):
        return _pad_crop(x, size, "pad")

    def crop_img_tensor(x, size):
        return _pad_crop(x, size, "crop")

    class Interpolate(nn.Module):
        def __init__(self, size=None, scale=None, mode="bilinear", align_corners=False):
            super().__init__()
            assert (size is None) == (scale is not None)
            self.size, self.scale = size, scale
            self.mode, self.align_corners = mode, align_corners

        def forward(self, x):
            return F.interpolate(x, size=self.size, scale_factor=self.scale,
                                 mode=self.mode, align_corners=self.align_corners)

    def free_bits_kl(kl, free_bits, batch_average=False, eps=1e-6):
        assert kl.dim() == 2
        if free_bits < eps:
            return kl.mean(0)
        if batch_average:
            return kl.mean(0).clamp(min=free_bits)
        return kl.clamp(min=free_bits).mean(0)

    boilr = types.ModuleType("boilr")
    boilr.__path__ = []
    models = types.ModuleType("boilr.models")
    models.BaseGenerativeModel = BaseGenerativeModel
    bnn = types.ModuleType("boilr.nn")
    bnn.crop_img_tensor = crop_img_tensor
    bnn.pad_img_tensor = pad_img_tensor
    bnn.Interpolate = Interpolate
    bnn.free_bits_kl = free_bits_kl
    boilr.models, boilr.nn = models, bnn
    return {"boilr": boilr, "boilr.models": models, "boilr.nn": bnn}


def load_reference():
    """Return the reference's modules: dict(lvae=, lvae_layers=, nn=, stochastic=, likelihoods=)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if "_lvae_reference_modules" in sys.modules:
        return sys.modules["_lvae_reference_modules"].mods
    saved = {k: sys.modules.get(k) for k in
             ("boilr", "boilr.models", "boilr.nn", "matplotlib", "matplotlib.pyplot",
              "models", "models.lvae", "models.lvae_layers", "lib", "lib.nn",
              "lib.stochastic", "lib.likelihoods")}
    for k in saved:
        sys.modules.pop(k, None)
    sys.modules.update(_boilr_stubs())
    mpl = types.ModuleType("matplotlib")
    mpl.__path__ = []
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        mods = dict(
            lvae=importlib.import_module("models.lvae"),
            lvae_layers=importlib.import_module("models.lvae_layers"),
            nn=importlib.import_module("lib.nn"),
            stochastic=importlib.import_module("lib.stochastic"),
            likelihoods=importlib.import_module("lib.likelihoods"),
        )
    finally:
        sys.path.remove(REFERENCE_ROOT)
        # Un-register the reference's top-level names so that our own package's
        # ``models`` / ``lib`` mirrors can be imported in the same process.
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    holder = types.ModuleType("_lvae_reference_modules")
    holder.mods = mods
    sys.modules["_lvae_reference_modules"] = holder
    return mods


class EpsQueue:
    """Context manager: make ``Normal.rsample`` consume pre-drawn eps tensors.

    torch.distributions.Normal.rsample (used at lib/stochastic.py:65) draws
    ``_standard_normal(shape, dtype, device)``; replacing that function with a
    FIFO of fixed tensors gives ``z = mu + sigma * eps`` with our eps.
    """

    def __init__(self, eps_list):
        self.eps = list(eps_list)

    def __enter__(self):
        import torch.distributions.normal as tdn
        self._mod = tdn
        self._orig = tdn._standard_normal
        queue = self.eps

        def fake(shape, dtype, device):
            e = queue.pop(0)
            assert tuple(e.shape) == tuple(shape), (e.shape, shape)
            return e.to(dtype=dtype, device=device)

        tdn._standard_normal = fake
        return self

    def __exit__(self, *a):
        self._mod._standard_normal = self._orig
        return False


class DropoutMaskQueue:
    """Context manager: make ``F.dropout2d`` (nn.Dropout2d, lib/nn.py:62,76,89)
    multiply by pre-drawn keep masks (B,C,1,1) already scaled by 1/(1-p)."""

    def __init__(self, masks):
        self.masks = list(masks)

    def __enter__(self):
        import torch.nn.functional as F
        self._F = F
        self._orig = F.dropout2d
        queue = self.masks

        def fake(input, p=0.5, training=True, inplace=False):
            if not training or p == 0.0:
                return input
            m = queue.pop(0)
            return input * m.to(input.dtype)

        F.dropout2d = fake
        return self

    def __exit__(self, *a):
        self._F.dropout2d = self._orig
        return False
