"""Simple data-dependent initialisation (call site: experiment/experiment_manager.py:62-72): every Conv2d /
ConvTranspose2d / Linear found through model.modules() is re-initialised (Kaiming normal weights, zero bias) and, during
ONE forward pass on a batch, a forward hook rescales its weight and bias -- and the output it is about to hand on -- so
that each output channel has zero mean and unit standard deviation on that batch."""
import torch
from torch import nn


def data_dependent_init(model, model_input_dict, eps=1e-6):
    handles = []
    affine = (nn.Conv2d, nn.ConvTranspose2d, nn.Linear)

    def hook(module, inputs, output):
        with torch.no_grad():
            dims = [d for d in range(output.dim()) if d != 1]
            mean = output.float().mean(dim=dims)
            std = output.float().std(dim=dims)
            scale = 1.0 / (std + eps)
            # output channel axis of the weight: 0 for Conv2d / Linear, 1 for ConvTranspose2d
            if isinstance(module, nn.ConvTranspose2d):
                module.weight.data.mul_(scale.view(1, -1, 1, 1).to(module.weight.dtype))
            else:
                module.weight.data.mul_(scale.view(-1, *([1] * (module.weight.dim() - 1))).to(module.weight.dtype))
            if module.bias is not None:
                module.bias.data.copy_(((module.bias.data.float() - mean) * scale).to(module.bias.dtype))
            shape = [1, -1] + [1] * (output.dim() - 2)
            return ((output.float() - mean.view(shape)) * scale.view(shape)).to(output.dtype)

    for m in model.modules():
        if isinstance(m, affine):
            nn.init.kaiming_normal_(m.weight.data)
            if m.bias is not None:
                m.bias.data.zero_()
            handles.append(m.register_forward_hook(hook))
    was_training = model.training
    model.train()
    with torch.no_grad():
        model(**model_input_dict)
    model.train(was_training)
    for h in handles:
        h.remove()
