"""SummaryProfile, ProgressBar, gradient_check (mirror of xs/utils/toolkit.py:8-194)."""
import os
import time
import warnings
from collections import OrderedDict

import numpy as np

from .common import no_grad


def format_time(seconds):
    seconds = int(seconds)
    h, rem = divmod(seconds, 3600)
    m, s = divmod(rem, 60)
    if h:
        return "%dh%dm%ds" % (h, m, s)
    if m:
        return "%dm%ds" % (m, s)
    return "%ds" % s


class SummaryProfile:
    """Wall time + host RSS around a training run, plus named metric series (toolkit.py:8-57)."""

    def __init__(self, *names, **kwargs):
        self._start = None
        self._psutil = None
        self._data = OrderedDict((n, []) for n in names)
        self._print_gap = kwargs.pop("print_gap", None)
        self._verbose = kwargs.pop("verbose", 1)
        self._steps = 0

    def __enter__(self):
        try:
            self._psutil = __import__("psutil")
        except ImportError:
            warnings.warn("Please install psutil module!")
        self._start = time.time()
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        if self._verbose:
            print("\033[1;31mTime cost: %.3fs\033[0m" % (time.time() - self._start))
            if self._psutil is not None:
                rss = self._psutil.Process(os.getpid()).memory_info().rss / 1024 ** 3
                print("\033[1;31mMemory cost: %.4f GB\033[0m" % rss)

    def step_all(self, *datas):
        for i, k in enumerate(self._data.keys()):
            self._data[k].append(datas[i])
        if self._print_gap is not None and self._steps % self._print_gap == 0:
            print("### Step", self._steps, "###", ", ".join("%s: %s" % (k, d) for k, d in zip(self._data, datas)))
        self._steps += 1

    def step(self, name, value):
        self._data[name].append(value)

    def visualize(self):
        import matplotlib.pyplot as plt
        for name, value in self._data.items():
            plt.figure()
            plt.plot(value)
            plt.title(name)
        plt.show()

    @property
    def data(self):
        return self._data


class ProgressBar:
    def __init__(self, max_iter=100, verbose=1, bar_nums=20):
        self.max_iter, self.verbose, self.bar_nums = max_iter, verbose, bar_nums
        self.iter = 0

    def update(self, n_iter=1):
        self.iter += n_iter

    def console(self, verbose=1, time_used=0.0, loss=0.0, acc=0.0, v_loss=0.0, v_acc=0.0):
        if verbose == 0:
            return
        done = int(self.bar_nums * min(self.iter, self.max_iter) / max(self.max_iter, 1))
        bar = "=" * done + ">" + "." * (self.bar_nums - done)
        print("\r%d/%d [%s] - %s - loss: %.4f - acc: %.4f" % (self.iter, self.max_iter, bar, format_time(time_used),
                                                               loss, acc), end="", flush=True)


def gradient_check(inputs, target, layer, criterion, epsilon=1e-4):
    """Central-difference gradients of criterion(layer(inputs), target) w.r.t. every parameter
    (toolkit.py:160-194). Each probe writes the perturbed parameter through `param.eval = ...`."""
    variables = list(layer.parameters())
    numeric = []
    with no_grad():
        for var in variables:
            base = var.eval.astype(np.float64)
            flat = base.reshape(-1).copy()
            grad = np.zeros(flat.size)
            for j in range(flat.size):
                keep = flat[j]
                flat[j] = keep + epsilon
                var.eval = flat.reshape(base.shape)
                loss1 = float(criterion(layer(inputs), target).eval.reshape(-1)[0])
                flat[j] = keep - epsilon
                var.eval = flat.reshape(base.shape)
                loss2 = float(criterion(layer(inputs), target).eval.reshape(-1)[0])
                flat[j] = keep
                grad[j] = (loss1 - loss2) / (2 * epsilon)
            var.eval = base
            numeric.append(grad.reshape(base.shape))
    return numeric
