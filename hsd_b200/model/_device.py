"""Device scoping for the model classes.

The C library launches on the CURRENT CUDA device and torch allocates on whatever device a
tensor names; a model built with ``device=`` different from the current device would otherwise run
kernels on one GPU against another GPU's pointers.  Every public method that reaches a kernel is
wrapped so that it runs under ``torch.cuda.device(model._device())``."""
import functools

import torch


def on_model_device(fn):
    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        with torch.cuda.device(self._device()):
            return fn(self, *args, **kwargs)
    return wrapper
