"""Learning-rate schedulers: closed-form host scalars (okl.py:2708-2761).  The value reaches the update kernels as a scalar
argument, or through a device scalar when the train step is replayed from a CUDA graph."""
import math


class LearningRateScheduler:
    def __init__(self, initial_lr: float):
        self.initial_lr = initial_lr
        self.current_lr = initial_lr

    def step(self, epoch: int) -> float:
        return self.current_lr                                   # constant (okl.py:2713-2715)

    def get_lr(self) -> float:
        return self.current_lr


class StepLRScheduler(LearningRateScheduler):
    def __init__(self, initial_lr: float, step_size: int, gamma: float):
        super().__init__(initial_lr)
        self.step_size, self.gamma = step_size, gamma

    def step(self, epoch: int) -> float:
        self.current_lr = self.initial_lr * (self.gamma ** (epoch // self.step_size))      # okl.py:2727-2729
        return self.current_lr


class TimeBasedDecayScheduler(LearningRateScheduler):
    def __init__(self, initial_lr: float, decay_rate: float):
        super().__init__(initial_lr)
        self.decay_rate = decay_rate

    def step(self, epoch: int) -> float:
        self.current_lr = self.initial_lr / (1 + self.decay_rate * epoch)                  # okl.py:2737-2739
        return self.current_lr


class ExponentialDecayScheduler(LearningRateScheduler):
    def __init__(self, initial_lr: float, decay_rate: float):
        super().__init__(initial_lr)
        self.decay_rate = decay_rate

    def step(self, epoch: int) -> float:
        self.current_lr = self.initial_lr * math.exp(-self.decay_rate * epoch)             # okl.py:2747-2749
        return self.current_lr


class LinearDecayScheduler(LearningRateScheduler):
    def __init__(self, initial_lr: float, final_lr: float, total_epochs: int):
        super().__init__(initial_lr)
        self.final_lr, self.total_epochs = final_lr, total_epochs
        self.decay_rate = (initial_lr - final_lr) / total_epochs

    def step(self, epoch: int) -> float:
        self.current_lr = max(self.final_lr, self.initial_lr - self.decay_rate * epoch)     # okl.py:2759-2761
        return self.current_lr
