"""Device plumbing under the Tensor: runtime singleton (device, stream, scratch) and DevArray,
the owning handle of one fp32 / int64 buffer in HBM.

Data layout in HBM: every 4-D array is stored channels-last ([N,H,W,C]; filters
[OC,KH,KW,C]) while its *logical* shape stays the reference's NCHW / [OC,C,KH,KW]
(xs/core/base.py:69-75: `.eval` is the universal NCHW read/write port). All other ranks are plain
row-major. PyTorch is used only as the caching device allocator, for streams and for host<->device
copies; every computation goes through libxsb200.
"""
import os

import numpy as np

from .._lib import MATH_FP32, MATH_TF32, lib

_torch = None


def torch():
    global _torch
    if _torch is None:
        import torch as _t
        _torch = _t
    return _torch


class Runtime:
    """One process drives one GPU (LOCAL_RANK) from one host thread."""

    WS_SMALL_FLOATS = 4 << 20     # BN / colsum partials + ticket counters (zero-initialised once)
    WS_BIG_FLOATS = 80 << 20      # split-K partials (FP32 implicit GEMM); flipped / per-class sub-filters (tcgen05 dgrad)
    WS_STATS_FLOATS = 8 << 20     # BatchNorm (count, mean, M2) partials left by a conv epilogue (xsb_conv2d_fwd_stats)

    def __init__(self):
        self.ready = False
        self.device = None
        self.math_mode = MATH_FP32
        self._ws_small = None
        self._ws_big = None
        self._ws_stats = None
        self._label_flag = None
        self.launches = 0          # kernel-launching C-ABI calls issued (bench.py reports it)
        # BatchNorm -> in-place ReLU -> MaxPool as one kernel (xsb_bn_relu_maxpool_fwd): True = the 3x3 / stride 2 / pad 1
        # pool of the ResNet stem (direct kernel: 165 us against 135 + 90 us for the two bandwidth kernels it replaces on
        # the R224 stem), "all" = any geometry (shared-memory tile kernel, measured no faster: 288 us), False = never
        self.fuse_bn_relu_pool = {"0": False, "all": "all"}.get(os.environ.get("XSB_FUSE_BN_RELU_POOL", "1"), True)
        # ... and its backward: the BatchNorm backward gathers the 3x3/2 pool's gradient on the fly (xsb_bn_bwd_pool_dxsum)
        # instead of reading a scattered copy twice. The gather (~350 instructions per 2x2 quad) is paid twice and the
        # kernels are instruction-bound: 396 us against 88 + 311 us back to back on L2-warm tensors, but -42 us inside the
        # training step (5.419 -> 5.377 ms), where the 411 MB scattered copy is cold
        self.fuse_pool_bn_bwd = os.environ.get("XSB_FUSE_POOL_BN_BWD", "1") != "0"
        # BatchNormalization -> add -> ReLU(inplace), the tail of a residual block, as one kernel (xsb_bn_add_relu): the
        # BatchNorm output(s) are never written and read back
        self.fuse_bn_add_relu = os.environ.get("XSB_FUSE_BN_ADD_RELU", "1") != "0"
        # ... and, backward, the identity shortcut's masked gradient written by the BatchNorm backward of the other branch
        # (xsb_bn_bwd_dxsum_side) instead of by a relu_bwd_mask pass of its own
        self.fuse_add_bwd_side = os.environ.get("XSB_FUSE_ADD_BWD_SIDE", "1") != "0"

    def ensure(self):
        if self.ready:
            return self
        t = torch()
        if not t.cuda.is_available():
            raise RuntimeError("xshinnosuke_b200 needs a CUDA device (sm_100a): there is no CPU fallback")
        dev = int(os.environ.get("LOCAL_RANK", "0"))
        t.cuda.set_device(dev)
        lib.xsb_init(dev)
        self.device = t.device("cuda", dev)
        self.ready = True
        return self

    def stream(self):
        return torch().cuda.current_stream().cuda_stream

    def empty(self, n, dtype="float32"):
        t = torch()
        return t.empty(int(n), dtype=t.float32 if dtype == "float32" else t.int64 if dtype == "int64" else t.uint8,
                       device=self.device)

    @property
    def ws_small(self):
        if self._ws_small is None:
            self._ws_small = torch().zeros(self.WS_SMALL_FLOATS, dtype=torch().float32, device=self.device)
        return self._ws_small

    @property
    def ws_big(self):
        if self._ws_big is None:
            self._ws_big = torch().empty(self.WS_BIG_FLOATS, dtype=torch().float32, device=self.device)
        return self._ws_big

    @property
    def ws_stats(self):
        if self._ws_stats is None:
            self._ws_stats = torch().empty(self.WS_STATS_FLOATS, dtype=torch().float32, device=self.device)
        return self._ws_stats

    @property
    def label_flag(self):
        """Device int the cross-entropy kernel sets when it meets a label outside [0, classes)."""
        if self._label_flag is None:
            self._label_flag = torch().zeros(1, dtype=torch().int32, device=self.device)
        return self._label_flag

    def take_label_error(self):
        """-> True (and clears the flag) if a cross-entropy since the last call saw an out-of-range label."""
        if self._label_flag is None or int(self._label_flag.item()) == 0:
            return False
        self._label_flag.zero_()
        return True

    def synchronize(self):
        torch().cuda.current_stream().synchronize()


rt = Runtime()


MATH_FP32X3 = 2   # host-level composition of XSB_MATH_TF32 calls (nn/x3.py); never passed to the C ABI


def set_math_mode(mode):
    """'fp32' (CUDA-core FFMA, <=1e-5 vs the reference), 'tf32' (tcgen05 tensor cores, <=2e-3) or 'fp32x3' (tcgen05 with
    the 3xTF32 operand split, <=1e-5: nn/x3.py)."""
    m = {"fp32": MATH_FP32, "tf32": MATH_TF32, "fp32x3": MATH_FP32X3, MATH_FP32: MATH_FP32, MATH_TF32: MATH_TF32,
         MATH_FP32X3: MATH_FP32X3}[mode]
    rt.math_mode = m


def get_math_mode():
    return {MATH_FP32: "fp32", MATH_TF32: "tf32", MATH_FP32X3: "fp32x3"}[rt.math_mode]


def fill_floats(buf, value, n=None):
    """buf[:n] = value through libxsb200 (xsb_fill: a memset node for 0, a fill kernel otherwise) -- no torch kernel runs."""
    n = buf.numel() if n is None else int(n)
    if n > 0:
        lib.xsb_fill(buf.data_ptr(), float(value), n, rt.stream())
        rt.launches += 1


def _prod(shape):
    n = 1
    for s in shape:
        n *= int(s)
    return n


class DevArray:
    """A device buffer + logical shape. `cl` marks a 4-D array physically stored channels-last.

    `pending_zero`: the contents are logically zero but have not been written. A kernel that
    accumulates into the array asks `take_acc()` -- 0 ("=", and the flag clears) the first time,
    1 ("+=") afterwards -- which is how `zero_grad` costs no memset and no read-modify-write
    (the reference allocates and fills zeros: xs/utils/common.py:29-32, xs/core/base.py:312-316).
    """
    __slots__ = ("buf", "shape", "dtype", "cl", "_pz", "_root", "lazy_mask", "side_out")

    def __init__(self, buf, shape, dtype="float32", cl=None, pending_zero=False, _share=None):
        self.buf = buf
        self.shape = tuple(int(s) for s in shape)
        self.dtype = dtype
        self.cl = (len(self.shape) == 4 and dtype == "float32") if cl is None else cl
        # lazy_mask: `buf` ALIASES another gradient and the logical contents are buf with the elements whose bit is set
        # in this ReLU bit mask zeroed (AddBackward hands dZ to a BatchNorm branch this way; the BN backward kernels apply
        # the mask on the fly, anyone else materialises a masked copy first)
        self.lazy_mask = None
        # side_out (set on a lazy-masked alias only): (gradient DevArray, parked closure) of the identity shortcut that wants the
        # same masked dZ written out -- the BatchNorm backward that applies the mask writes it on the way (AddBackward)
        self.side_out = None
        if _share is None:
            # one cell shared by an array and all its views: [pending_zero, pending_op, deferred_add]
            self._pz = [bool(pending_zero), None, None]
            self._root = None
        else:
            self._pz = _share._pz
            self._root = _share if _share._root is None else _share._root

    @property
    def pending_zero(self):
        return self._pz[0]

    @pending_zero.setter
    def pending_zero(self, v):
        self._pz[0] = bool(v)

    # pending_op: a deferred producer kernel (BN apply, residual add). It runs the first time the contents are
    # needed -- unless the consumer is an in-place ReLU, which runs the fused producer+ReLU(+mask) kernel instead
    # (layers/act.py). Called as op(mask_or_None).
    @property
    def pending_op(self):
        return self._pz[1]

    @pending_op.setter
    def pending_op(self, op):
        self._pz[1] = op

    def _run_pending_op(self):
        op = self._pz[1]
        if op is not None:
            self._pz[1] = None
            op(None)

    def flush(self):
        self._run_pending_op()
        add = self._pz[2]
        if add is not None:
            self._pz[2] = None
            add(self.take_acc())

    # deferred_add: an accumulating kernel that is cheap as "+=" and expensive as the first "=" write of a gradient (a
    # strided 1x1 dgrad reaches a quarter of the pixels: as "=" it must zero-fill the rest). It is postponed until the
    # gradient is READ; a later writer then finds the array still logically zero, writes "=" without a read-modify-write,
    # and the postponed kernel only adds where it has something to add. Called as add(accumulate_flag).
    def defer_add(self, add):
        """-> True if `add` was taken (the array is logically zero and holds no other postponed kernel)."""
        if self._pz[0] and self._pz[2] is None and self.lazy_mask is None:
            self._pz[2] = add
            return True
        return False

    def _materialize_zero(self):
        root = self if self._root is None else self._root
        if root.dtype == "float32":
            fill_floats(root.buf, 0.0)
        else:
            root.buf.zero_()
        self._pz[0] = False

    def _drop_deferred(self):
        """Everything parked on this array is void once its contents are overwritten directly: the deferred producer, the
        postponed add, the lazily masked alias (an alias is never written through: it gets its own buffer)."""
        if self.lazy_mask is not None:
            self.lazy_mask = None
            self.buf = rt.empty(max(self.size, 1))
        self._pz[1] = None
        self._pz[2] = None
        self.pending_zero = False

    # ---- construction
    @staticmethod
    def empty(shape, dtype="float32", pending_zero=False):
        rt.ensure()
        shape = tuple(int(s) for s in shape)
        return DevArray(rt.empty(max(_prod(shape), 1), dtype), shape, dtype, pending_zero=pending_zero)

    @staticmethod
    def zeros(shape, dtype="float32"):
        return DevArray.full(shape, 0.0, dtype)

    @staticmethod
    def full(shape, value, dtype="float32"):
        a = DevArray.empty(shape, dtype)
        if dtype == "float32":
            fill_floats(a.buf, value)
        else:
            a.buf.fill_(value)
        return a

    def fill(self, value):
        """Direct overwrite with a constant (Tensor.zero_ / fill_)."""
        self._drop_deferred()
        if self.dtype == "float32":
            fill_floats(self.buf, value, max(self.size, 1))
        else:
            self.buf[: max(self.size, 1)].fill_(value)

    @staticmethod
    def from_numpy(a):
        rt.ensure()
        a = np.asarray(a)
        if a.dtype.kind in "iu":
            dtype, host = "int64", np.ascontiguousarray(a, dtype=np.int64)
        else:
            dtype, host = "float32", np.ascontiguousarray(a, dtype=np.float32)
        if not host.flags.writeable:
            host = host.copy()   # torch.from_numpy refuses read-only arrays (e.g. views of an .npz)
        shape = host.shape if host.ndim else (1,)  # 0-d scalars behave as shape (1,) (xs/core/base.py:82-84)
        t = torch()
        dev = t.from_numpy(host.reshape(-1)).to(rt.device, non_blocking=True)   # async when `host` is pinned
        out = DevArray(dev, shape, dtype)
        if out.cl and out.size > 0:
            n, c, h, w = out.shape
            if c > 1 and h * w > 1:
                dst = rt.empty(out.size)
                lib.xsb_nchw_to_nhwc(dev.data_ptr(), dst.data_ptr(), n, c, h, w, 0, rt.stream())
                rt.launches += 1
                out.buf = dst
        return out

    # ---- properties
    @property
    def size(self):
        return _prod(self.shape)

    @property
    def ndim(self):
        return len(self.shape)

    def _materialize_mask(self):
        mask, src = self.lazy_mask, self.buf
        self.lazy_mask = None
        self.buf = rt.empty(max(self.size, 1))
        lib.xsb_relu_bwd_mask(src.data_ptr(), mask.data_ptr(), self.buf.data_ptr(), self.size, 0, rt.stream())
        rt.launches += 1

    @property
    def ptr(self):
        """Pointer for a kernel that will READ (or read-modify-write) the contents."""
        if self.lazy_mask is not None:
            self._materialize_mask()
        if self._pz[1] is not None or self._pz[2] is not None:
            self.flush()
        if self.pending_zero:
            self._materialize_zero()
        return self.buf.data_ptr()

    @property
    def wptr(self):
        """Pointer for a kernel that fully OVERWRITES the contents."""
        self._drop_deferred()
        return self.buf.data_ptr()

    def take_acc(self):
        """-> accumulate flag for a kernel that adds into this array (see class doc)."""
        if self.pending_zero:
            self.pending_zero = False
            return 0
        return 1

    def peek_ptr(self):
        """Pointer for a kernel that accumulates (paired with take_acc): a postponed add stays postponed."""
        if self.lazy_mask is not None:
            self._materialize_mask()
        self._run_pending_op()
        return self.buf.data_ptr()

    @property
    def phys_shape(self):
        if self.cl:
            n, c, h, w = self.shape
            return (n, h, w, c)
        return self.shape

    # ---- host transfer
    def to_numpy(self):
        rt.ensure()
        if self.lazy_mask is not None:
            self._materialize_mask()
        self.flush()
        if self.pending_zero:
            return np.zeros(self.shape, dtype=np.float32 if self.dtype == "float32" else np.int64)
        n = self.size
        src = self.buf[:n] if self.buf.numel() != n else self.buf
        if self.cl and n > 0:
            nn_, c, h, w = self.shape
            if c > 1 and h * w > 1:
                tmp = rt.empty(n)
                lib.xsb_nhwc_to_nchw(src.data_ptr(), tmp.data_ptr(), nn_, c, h, w, 0, rt.stream())
                rt.launches += 1
                src = tmp
        host = src.cpu().numpy().reshape(self.shape)
        # `.eval` hands out a COPY: when the buffer already lives on the host (CPU host-logic tests) .cpu() is the live buffer itself
        return host.copy() if src.device.type == "cpu" else host

    def assign_numpy(self, a):
        """Overwrite the contents from a host array of the same logical shape (keeps the buffer:
        parameter storage may be a view into the flat optimizer buffer)."""
        a = np.asarray(a)
        if tuple(a.shape) != tuple(self.shape):
            a = np.broadcast_to(a, self.shape)       # (read-only view: from_numpy copies it)
        new = DevArray.from_numpy(a)
        if new.dtype != self.dtype:
            raise TypeError(f"cannot assign {new.dtype} data to a {self.dtype} tensor")
        if new.cl != self.cl:                        # a 4-D array whose layout flag was forced: match it
            new = new.as_cl() if self.cl else new.logical_rowmajor()
        self._drop_deferred()
        self.buf[: max(self.size, 1)].copy_(new.buf[: max(self.size, 1)])

    # ---- views / copies
    def prefix(self, n):
        """First n entries along axis 0 (contiguous in both layouts)."""
        if n == self.shape[0]:
            return self
        per = self.size // self.shape[0] if self.shape[0] else 0
        return DevArray(self.buf[: max(n * per, 1)], (n,) + self.shape[1:], self.dtype, self.cl, _share=self)

    def clone(self):
        if self.lazy_mask is not None:
            self._materialize_mask()
        self.flush()
        if self.pending_zero:
            return DevArray.empty(self.shape, self.dtype, pending_zero=True)
        return DevArray(self.buf.clone(), self.shape, self.dtype, self.cl)

    def logical_rowmajor(self):
        """-> a DevArray whose physical order equals the logical (NCHW) row-major order."""
        if not self.cl:
            return self
        n, c, h, w = self.shape
        if c == 1 or h * w == 1:
            return DevArray(self.buf, self.shape, self.dtype, cl=False, _share=self)
        dst = rt.empty(self.size)
        lib.xsb_nhwc_to_nchw(self.ptr, dst.data_ptr(), n, c, h, w, 0, rt.stream())
        rt.launches += 1
        return DevArray(dst, self.shape, self.dtype, cl=False)

    def as_cl(self):
        """4-D row-major (NCHW) -> channels-last physical order; no-op for everything else."""
        if self.cl or len(self.shape) != 4 or self.dtype != "float32":
            return self
        n, c, h, w = self.shape
        if c == 1 or h * w == 1:
            return DevArray(self.buf, self.shape, self.dtype, cl=True, _share=self)
        dst = rt.empty(self.size)
        lib.xsb_nchw_to_nhwc(self.ptr, dst.data_ptr(), n, c, h, w, 0, rt.stream())
        rt.launches += 1
        return DevArray(dst, self.shape, self.dtype, cl=True)

    def reshaped(self, shape):
        """Reinterpret a row-major (non channels-last) buffer with a new logical shape (zero copy)."""
        assert not self.cl
        shape = tuple(int(s) for s in shape)
        assert _prod(shape) == self.size, (shape, self.shape)
        return DevArray(self.buf, shape, self.dtype, cl=False, _share=self)
