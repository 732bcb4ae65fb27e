from . import viz  # noqa: F401


def linear_anneal(x, start, end, steps):
    """start -> end over `steps` steps, constant afterwards (call site: experiment_manager.py:341)."""
    assert x >= 0 and steps > 0
    if x >= steps:
        return end
    return start + (end - start) * x / steps
