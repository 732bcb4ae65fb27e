from torch import nn


class BaseGenerativeModel(nn.Module):
    def __init__(self):
        super().__init__()
        self.global_step = 0

    def increment_global_step(self):
        self.global_step += 1

    def get_device(self):
        return next(self.parameters()).device
