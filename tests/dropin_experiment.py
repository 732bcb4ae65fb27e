"""Restatement (test infrastructure) of the three pieces of the reference's experiment layer the GPU drop-in test needs,
for the GPU box where /root/reference does not exist.  tests/test_dropin_cpu.py pins `forward_pass` against the real
LVAEExperiment.forward_pass where the reference is present."""
import torch


def linear_anneal(x, start, end, steps):
    if x >= steps:
        return end
    return start + (end - start) * x / steps


def forward_pass(model, x, device, beta_anneal=0):
    """experiment/experiment_manager.py:322-367: ELBO terms, loss with beta, L2 norm of the parameters."""
    x = x.to(device, non_blocking=True)
    model_out = model(x)
    recons_sep = -model_out["ll"]
    kl_sep = model_out["kl_sep"]
    elbo_sep = -(recons_sep + kl_sep)
    beta = 1.0
    if beta_anneal != 0:
        beta = linear_anneal(model.global_step, 0.0, 1.0, beta_anneal)
    recons = recons_sep.mean()
    loss = recons + model_out["kl_loss"] * beta
    l2 = 0.0
    for p in model.parameters():
        l2 = l2 + torch.sum(p ** 2)
    l2 = l2.sqrt()
    output = {"loss": loss, "elbo": elbo_sep.mean(), "elbo_sep": elbo_sep, "kl": model_out["kl"], "l2": l2, "recons": recons,
              "out_mean": model_out["out_mean"], "out_mode": model_out["out_mode"], "out_sample": model_out["out_sample"],
              "likelihood_params": model_out["likelihood_params"]}
    if "kl_avg_layerwise" in model_out:
        output["kl_avg_layerwise"] = model_out["kl_avg_layerwise"]
    return output
