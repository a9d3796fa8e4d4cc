"""First-order low-pass filter on the output joint vector (host mirror).

Same behaviour as the reference filter (src/dex_retargeting/optimizer_utils.py:1-17): the first
sample initialises the state, later samples move it by `alpha * (x - y)`.  The batched sequence
kernel applies the identical recurrence on device (csrc/dexr.cu, dexr_sequences_kernel).
"""
import numpy as np


class LPFilter:
    def __init__(self, alpha):
        self.alpha = alpha
        self.reset()

    def next(self, x):
        x = np.asarray(x)
        if self.is_init:
            self.y = self.y + self.alpha * (x - self.y)
        else:
            self.y, self.is_init = x, True
        return self.y.copy()

    def reset(self):
        self.y = None
        self.is_init = False
