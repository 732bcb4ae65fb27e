class BaseOfflineEvaluator:
    """evaluate.py subclasses this; only import-compatibility is exercised here."""

    def __init__(self, experiment_class=None):
        self._experiment_class = experiment_class

    def _add_args(self, parser):
        pass
