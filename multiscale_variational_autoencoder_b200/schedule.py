"""Step-decay learning-rate schedule (mvae/schedule.py:7-21): lr = initial_lr * decay_factor ** floor(epoch / step_size).

The reference wraps the function in a Keras LearningRateScheduler; here `train()` calls it once per epoch."""
import numpy as np


def step_decay_schedule(initial_lr, decay_factor=0.5, step_size=1):
    def schedule(epoch):
        return initial_lr * (decay_factor ** np.floor(epoch / step_size))

    return schedule
