"""Stand-in for boilr 0.7.x (test infrastructure; see ../README.md)."""
import argparse

import torch

from . import data, models, nn, utils  # noqa: F401

_options = {}


def set_options(**kw):
    _options.update(kw)


class BaseExperimentManager:
    """args -> dataloaders, model, optimizer; subclasses provide _make_* and forward_pass."""

    def __init__(self, args=None):
        self._config = None
        if args is None:
            args = self._parse_args([])
        self.args = args
        self.run_description = self._make_run_description(args)
        self.device = None
        self.dataloaders = None
        self.model = None
        self.optimizer = None

    # ---- arguments ----
    @classmethod
    def _define_args_defaults(cls):
        return dict(batch_size=64, test_batch_size=1000, lr=1e-3, seed=54321, train_log_every=1000, test_log_every=1000,
                    checkpoint_every=1000, keep_checkpoint_max=3, resume="", max_steps=10 ** 10, max_epochs=10 ** 7,
                    additional_descr="", nocuda=False, dry_run=False)

    def _add_args(self, parser):
        d = self._define_args_defaults()
        parser.add_argument("--batch-size", type=int, default=d["batch_size"], dest="batch_size")
        parser.add_argument("--test-batch-size", type=int, default=d["test_batch_size"], dest="test_batch_size")
        parser.add_argument("--lr", type=float, default=d["lr"])
        parser.add_argument("--seed", type=int, default=d["seed"])
        parser.add_argument("--tr-log-every", type=int, default=d["train_log_every"], dest="train_log_every")
        parser.add_argument("--ts-log-every", type=int, default=d["test_log_every"], dest="test_log_every")
        parser.add_argument("--checkpoint-every", type=int, default=d["checkpoint_every"])
        parser.add_argument("--keep-checkpoint-max", type=int, default=d["keep_checkpoint_max"])
        parser.add_argument("--max-steps", type=int, default=d["max_steps"])
        parser.add_argument("--max-epochs", type=int, default=d["max_epochs"])
        parser.add_argument("--nocuda", action="store_true")
        parser.add_argument("--descr", type=str, default=d["additional_descr"], dest="additional_descr")
        parser.add_argument("--dry-run", action="store_true")
        parser.add_argument("--resume", type=str, default=d["resume"])

    @classmethod
    def _check_args(cls, args):
        return args

    def _parse_args(self, argv=None):
        parser = argparse.ArgumentParser(allow_abbrev=False)
        self._add_args(parser)
        return self._check_args(parser.parse_args(argv))

    @staticmethod
    def _make_run_description(args):
        return "run"

    # ---- set-up ----
    def setup(self, device, checkpoint_folder=None):
        self.device = torch.device(device)
        torch.manual_seed(self.args.seed)
        self.dataloaders = self._make_datamanager()
        self.model = self._make_model()
        self.optimizer = self._make_optimizer()
        return self


class VAEExperimentManager(BaseExperimentManager):

    @classmethod
    def _define_args_defaults(cls):
        d = super()._define_args_defaults()
        d.update(loglikelihood_every=50000, loglikelihood_samples=100)
        return d

    def _add_args(self, parser):
        super()._add_args(parser)
        d = self._define_args_defaults()
        parser.add_argument("--ll-every", type=int, default=d["loglikelihood_every"], dest="loglikelihood_every")
        parser.add_argument("--ll-samples", type=int, default=d["loglikelihood_samples"], dest="loglikelihood_samples")

    def test_procedure(self, iw_samples=None):
        """ELBO and, with iw_samples = K, the importance-weighted bound log mean_k exp(elbo_k) over the test set
        (one full forward_pass per sample, as boilr does it; call site evaluate.py:30)."""
        import math
        elbos, iws, n = 0.0, 0.0, 0
        for x, _ in self.dataloaders.test:
            out = self.forward_pass(x)
            b = x.shape[0]
            elbos += float(out["elbo_sep"].sum())
            if iw_samples:
                all_k = [out["elbo_sep"].detach()]
                for _ in range(iw_samples - 1):
                    all_k.append(self.forward_pass(x)["elbo_sep"].detach())
                iw = torch.logsumexp(torch.stack(all_k, 1), dim=1) - math.log(iw_samples)
                iws += float(iw.sum())
            n += b
        res = {"elbo/elbo": elbos / n}
        if iw_samples:
            res["elbo/elbo_IW_{}".format(iw_samples)] = iws / n
        return res
