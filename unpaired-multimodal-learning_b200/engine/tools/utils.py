"""Small host utilities used by the entry points (reference: engine/tools/utils.py)."""
import os
import random

import numpy as np
import torch


class Tee:
    """File-like object that duplicates writes (stdout + log.txt, reference finetune.py:475-476)."""

    def __init__(self, *streams):
        self.streams = streams

    def write(self, text):
        for s in self.streams:
            s.write(text)

    def flush(self):
        for s in self.streams:
            s.flush()


def set_random_seed(seed):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def makedirs(path):
    if path and not os.path.exists(path):
        os.makedirs(path, exist_ok=True)
