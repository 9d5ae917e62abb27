"""DataSet / DataLoader (mirror of xs/utils/data/dataset.py:1-12, dataloader.py:7-65).

Batches are pre-sliced on the host exactly like the reference (including the re-seeding of the GLOBAL NumPy
RNG with 0,1,2,... per epoch, Q11). SURVEY 8f-2: instead of the reference's synchronous per-batch
`x.to('cuda')` copy (tests/resnet/resnet18_dg.py:94) the loader stages the NEXT batch in pinned host memory and
uploads it on a side CUDA stream while the current step computes (double buffering), and it can shard every
batch across data-parallel ranks (`rank`, `world`: contiguous slices).
"""
import numpy as np

from ...core.device import rt, torch
from ...core.tensor import Tensor


class DataSet:
    def __init__(self, *datas):
        self.datas = list(datas)

    def __len__(self):
        return len(self.datas[0])

    def __getitem__(self, item):
        return [d[item] for d in self.datas]


class DataLoader:
    def __init__(self, dataset, batch_size=None, shuffle=False, seed=None, prefetch=True, rank=0, world=1):
        self.dataset = dataset
        self.batch_size = batch_size
        self.shuffle = shuffle
        self.seed = 0 if seed is None else seed
        self.prefetch = prefetch
        self.rank, self.world = rank, world
        self.mini_batches = self.make_batches(self.dataset.datas, self.batch_size, self.seed, self.shuffle)
        self.sp = -1
        self._staged = None      # (index, tensors, event) of the batch uploaded ahead of time
        self._side = None

    def __iter__(self):
        return self

    # ---- upload machinery
    def _shard(self, arrays):
        if self.world <= 1:
            return arrays
        out = []
        for a in arrays:
            per = a.shape[0] // self.world
            out.append(a[self.rank * per:(self.rank + 1) * per])
        return out

    def _upload(self, index):
        arrays = [np.asarray(d) for d in self._shard(self.mini_batches[index])]
        if not (self.prefetch and rt.ready and rt.device is not None and rt.device.type == "cuda"):
            return [Tensor(a) for a in arrays], None
        t = torch()
        if self._side is None:
            self._side = t.cuda.Stream(device=rt.device)
        main = t.cuda.current_stream()
        with t.cuda.stream(self._side):
            tensors = []
            for a in arrays:
                dt = np.int64 if a.dtype.kind in "iu" else np.float32
                pinned = t.from_numpy(np.ascontiguousarray(a, dtype=dt)).pin_memory()
                ten = Tensor(pinned.numpy())          # H2D + NCHW -> channels-last on the side stream
                ten._arr.buf.record_stream(main)
                ten._pinned = pinned                  # keep the staging buffer alive until the copy ran
                tensors.append(ten)
            ev = t.cuda.Event()
            ev.record(self._side)
        return tensors, ev

    def __next__(self):
        self.sp += 1
        if self.sp >= len(self.mini_batches):
            self.seed += 1
            self.mini_batches = self.make_batches(self.dataset.datas, self.batch_size, self.seed, self.shuffle)
            self.sp = -1
            self._staged = None
            raise StopIteration
        rt.ensure()
        if self._staged is not None and self._staged[0] == self.sp:
            _, out, ev = self._staged
        else:
            out, ev = self._upload(self.sp)
        if ev is not None:
            torch().cuda.current_stream().wait_event(ev)
        self._staged = None
        if self.prefetch and self.sp + 1 < len(self.mini_batches):
            nxt, nev = self._upload(self.sp + 1)       # overlaps with the step the caller is about to run
            self._staged = (self.sp + 1, nxt, nev)
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
