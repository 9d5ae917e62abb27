"""DataSet / DataLoader (mirror of xs/utils/data/dataset.py:1-12, dataloader.py:7-65).

Batches are pre-sliced on the host exactly like the reference (including the re-seeding of the
GLOBAL NumPy RNG with 0,1,2,... per epoch, Q11) and each one is uploaded to the device when it is
yielded -- the reference's per-batch `x.to('cuda')` copy (tests/resnet/resnet18_dg.py:94).
"""
import numpy as np

from ...core.tensor import Tensor


class DataSet:
    def __init__(self, *datas):
        self.datas = list(datas)

    def __len__(self):
        return len(self.datas[0])

    def __getitem__(self, item):
        return [d[item] for d in self.datas]


class DataLoader:
    def __init__(self, dataset, batch_size=None, shuffle=False, seed=None):
        self.dataset = dataset
        self.batch_size = batch_size
        self.shuffle = shuffle
        self.seed = 0 if seed is None else seed
        self.mini_batches = self.make_batches(self.dataset.datas, self.batch_size, self.seed, self.shuffle)
        self.sp = -1

    def __iter__(self):
        return self

    def __next__(self):
        self.sp += 1
        if self.sp >= len(self.mini_batches):
            self.seed += 1
            self.mini_batches = self.make_batches(self.dataset.datas, self.batch_size, self.seed, self.shuffle)
            self.sp = -1
            raise StopIteration
        out = [Tensor(np.asarray(d)) for d in self.mini_batches[self.sp]]
        return out[0] if len(out) == 1 else out

    def __len__(self):
        return len(self.mini_batches)

    @staticmethod
    def make_batches(datas, batch_size, seed, shuffle=False):
        if batch_size is None:
            return [datas]
        np.random.seed(seed)
        m = datas[0].shape[0]
        if shuffle:
            perm = np.random.permutation(m)
            datas = [np.asarray(d)[perm] for d in datas]
        full = m // batch_size
        batches = [[d[batch_size * i:batch_size * (i + 1)] for d in datas] for i in range(full)]
        if m % batch_size != 0:
            batches.append([d[batch_size * full:] for d in datas])
        return batches
