"""rustfft_b200 -- B200-native batched complex FFT behind RustFFT's FftPlanner / Fft interface.

Host-side mirror (Python, because the image has no Rust toolchain) of the reference's public
API for the hot path; the names, argument meaning and error behaviour follow the reference
(paths relative to the RustFFT 6.4.1 tree):

    FftPlanner.plan_fft / plan_fft_forward / plan_fft_inverse      src/plan.rs:67-126
    Fft.process / process_with_scratch                             src/lib.rs:195-211
    Fft.process_outofplace_with_scratch                            src/lib.rs:231-236
    Fft.process_immutable_with_scratch                             src/lib.rs:250-255
    Fft.get_{inplace,outofplace,immutable}_scratch_len             src/lib.rs:262-277
    Fft.len / Fft.fft_direction                                    src/lib.rs:140-181
    FftDirection                                                   src/lib.rs:147-171

Everything numeric happens in rustfft_b200/libb200fft.so (hand-written sm_100a CUDA behind the
C ABI of include/b200fft.h).  There is no CPU fallback: without the library or without a B200
the planner raises.  PyTorch is only used for device memory / streams by callers (tests, bench).
"""
from __future__ import annotations

import ctypes
import enum
import os
import threading
from typing import Dict, Optional, Tuple

import numpy as np

__all__ = ["FftDirection", "FftPlanner", "Fft", "Library", "FftError", "Recipe", "RealFftPlanner", "RealFft", "Fft2d", "default_library", "shard_range"]

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200FFT_LIB: load another build of the same C ABI (A/B measurements of kernel variants; tools/ab_two_pass.py)
DEFAULT_LIB_PATH = os.environ.get("B200FFT_LIB") or os.path.join(_HERE, "libb200fft.so")

F32, F64 = 0, 1


class FftError(RuntimeError):
    """Raised where the reference panics (src/common.rs:13-104) or where CUDA fails."""

    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code


class FftDirection(enum.IntEnum):
    Forward = 0
    Inverse = 1

    def opposite_direction(self) -> "FftDirection":  # src/lib.rs:156-161
        return FftDirection.Inverse if self == FftDirection.Forward else FftDirection.Forward


class _RecipeNode(ctypes.Structure):  # b200fft_recipe_node (include/b200fft.h)
    _fields_ = [("kind", ctypes.c_uint32), ("child", ctypes.c_uint32), ("len", ctypes.c_uint64), ("a", ctypes.c_uint64), ("b", ctypes.c_uint64)]


class Recipe:
    """A decomposition chosen by the host (the reference's Recipe enum, src/plan.rs:134-226) handed to the library as data.

    Built with the class methods; `inner` is the Recipe of the inner FFT of a Rader / Bluestein node."""

    AUTO, POW2, SMOOTH, MIXED_RADIX, GOOD_THOMAS, RADER, BLUESTEIN, CLUSTER = range(8)

    def __init__(self, kind: int, len: int, a: int = 0, b: int = 0, inner: Optional["Recipe"] = None):
        self.kind, self.len, self.a, self.b, self.inner = kind, int(len), int(a), int(b), inner

    @classmethod
    def pow2(cls, n):
        return cls(cls.POW2, n)

    @classmethod
    def cluster(cls, n, half_tiles=False):
        return cls(cls.CLUSTER, n, 1 if half_tiles else 0)

    def to_dict(self):
        """JSON-friendly form (plan serialisation): Recipe.from_dict(json.loads(json.dumps(r.to_dict()))) rebuilds the same plan."""
        d = {"kind": self.kind, "len": self.len, "a": self.a, "b": self.b}
        if self.inner is not None:
            d["inner"] = self.inner.to_dict()
        return d

    @classmethod
    def from_dict(cls, d):
        return cls(d["kind"], d["len"], d.get("a", 0), d.get("b", 0), cls.from_dict(d["inner"]) if d.get("inner") else None)

    @classmethod
    def smooth(cls, n):
        return cls(cls.SMOOTH, n)

    @classmethod
    def mixed_radix(cls, a, b):
        return cls(cls.MIXED_RADIX, a * b, a, b)

    @classmethod
    def good_thomas(cls, a, b):
        return cls(cls.GOOD_THOMAS, a * b, a, b)

    @classmethod
    def rader(cls, n, outer_radix=1, inner: Optional["Recipe"] = None):
        return cls(cls.RADER, n, outer_radix, 0, inner)

    @classmethod
    def bluestein(cls, n, inner: Optional["Recipe"] = None):
        return cls(cls.BLUESTEIN, n, 0, 0, inner)

    def flatten(self):
        nodes, r = [], self
        while r is not None:
            nodes.append(r)
            r = r.inner
        arr = (_RecipeNode * len(nodes))()
        for i, r in enumerate(nodes):
            arr[i] = _RecipeNode(r.kind, i + 1 if r.inner is not None else 0, r.len, r.a, r.b)
        return arr


class Library:
    """A loaded C-ABI library (include/b200fft.h)."""

    SYMBOLS = [
        "b200fft_device_count", "b200fft_plan_create", "b200fft_plan_create_from_recipe", "b200fft_plan_recipe", "b200fft_plan_destroy",
        "b200fft_plan_len",
        "b200fft_plan_direction", "b200fft_plan_precision", "b200fft_plan_scratch_len", "b200fft_plan_describe",
        "b200fft_plan_launches", "b200fft_exec_host_inplace", "b200fft_exec_host_outofplace", "b200fft_exec_device",
        "b200fft_workspace_bytes", "b200fft_exec_device_ws", "b200fft_last_error", "b200fft_version",
        "b200fft_real_plan_create", "b200fft_real_plan_destroy", "b200fft_real_workspace_bytes", "b200fft_real_forward_device",
        "b200fft_real_inverse_device", "b200fft_real_forward_host", "b200fft_real_inverse_host",
        "b200fft_plan2d_create", "b200fft_plan2d_destroy", "b200fft_exec2d_device", "b200fft_exec2d_host",
    ]

    def __init__(self, path: str = DEFAULT_LIB_PATH):
        if not os.path.exists(path):
            raise FftError(-2, f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
        self.path = path
        self.c = ctypes.CDLL(path)
        c, u64, i32, vp = self.c, ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p
        c.b200fft_device_count.argtypes = [ctypes.POINTER(i32)]
        c.b200fft_plan_create.argtypes = [ctypes.POINTER(vp), u64, i32, i32, i32]
        c.b200fft_plan_create_from_recipe.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(_RecipeNode), ctypes.c_uint32, i32, i32, i32]
        c.b200fft_plan_recipe.argtypes = [vp, ctypes.POINTER(_RecipeNode), ctypes.c_uint32]
        c.b200fft_plan_destroy.argtypes = [vp]
        c.b200fft_plan_len.argtypes = [vp]
        c.b200fft_plan_len.restype = u64
        c.b200fft_plan_direction.argtypes = [vp]
        c.b200fft_plan_precision.argtypes = [vp]
        c.b200fft_plan_scratch_len.argtypes = [vp, i32]
        c.b200fft_plan_scratch_len.restype = u64
        c.b200fft_plan_describe.argtypes = [vp, ctypes.c_char_p, u64]
        c.b200fft_plan_launches.argtypes = [vp, u64]
        c.b200fft_plan_launches.restype = u64
        c.b200fft_exec_host_inplace.argtypes = [vp, vp, u64]
        c.b200fft_exec_host_outofplace.argtypes = [vp, vp, vp, u64]
        c.b200fft_exec_device.argtypes = [vp, vp, vp, u64, vp]
        c.b200fft_workspace_bytes.argtypes = [vp, u64]
        c.b200fft_workspace_bytes.restype = u64
        c.b200fft_exec_device_ws.argtypes = [vp, vp, vp, u64, vp, vp, u64]
        c.b200fft_last_error.restype = ctypes.c_char_p
        c.b200fft_version.restype = ctypes.c_char_p
        c.b200fft_real_plan_create.argtypes = [ctypes.POINTER(vp), u64, i32, i32]
        c.b200fft_real_plan_destroy.argtypes = [vp]
        c.b200fft_real_workspace_bytes.argtypes = [vp, u64]
        c.b200fft_real_workspace_bytes.restype = u64
        c.b200fft_real_forward_device.argtypes = [vp, vp, vp, u64, vp]
        c.b200fft_real_inverse_device.argtypes = [vp, vp, vp, u64, vp]
        c.b200fft_real_forward_host.argtypes = [vp, vp, vp, u64]
        c.b200fft_real_inverse_host.argtypes = [vp, vp, vp, u64]
        c.b200fft_plan2d_create.argtypes = [ctypes.POINTER(vp), u64, u64, i32, i32, i32]
        c.b200fft_plan2d_destroy.argtypes = [vp]
        c.b200fft_exec2d_device.argtypes = [vp, vp, vp, u64, vp]
        c.b200fft_exec2d_host.argtypes = [vp, vp, vp, u64]

    def device_count(self) -> int:
        n = ctypes.c_int(0)
        self.check(self.c.b200fft_device_count(ctypes.byref(n)))
        return n.value

    def version(self) -> str:
        return self.c.b200fft_version().decode()

    def check(self, rc: int) -> None:
        if rc != 0:
            raise FftError(rc, self.c.b200fft_last_error().decode())


_default: Optional[Library] = None
_default_lock = threading.Lock()


def default_library() -> Library:
    global _default
    with _default_lock:
        if _default is None:
            _default = Library(DEFAULT_LIB_PATH)
        return _default


_DTYPES = {np.dtype(np.complex64): F32, np.dtype(np.complex128): F64}


class Fft:
    """One planned transform: the Arc<dyn Fft<T>> of the reference (src/lib.rs:184-278).

    Immutable after construction and safe to call from many threads (examples/concurrency.rs)."""

    def __init__(self, lib: Library, length: int, direction: FftDirection, precision: int, device: int, recipe: Optional[Recipe] = None):
        self._lib = lib
        self._h = ctypes.c_void_p()
        if recipe is not None:
            nodes = recipe.flatten()
            lib.check(lib.c.b200fft_plan_create_from_recipe(ctypes.byref(self._h), nodes, len(nodes), int(direction), precision, device))
        else:
            lib.check(lib.c.b200fft_plan_create(ctypes.byref(self._h), length, int(direction), precision, device))
        self._len = length
        self._direction = FftDirection(direction)
        self._precision = precision
        self.device = device

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self._lib.c.b200fft_plan_destroy(h)
            except Exception:
                pass

    # ---- Length / Direction ------------------------------------------------------------
    def len(self) -> int:
        return self._len

    def __len__(self) -> int:
        return self._len

    def fft_direction(self) -> FftDirection:
        return self._direction

    @property
    def dtype(self):
        return np.complex64 if self._precision == F32 else np.complex128

    def describe(self) -> str:
        buf = ctypes.create_string_buffer(512)
        rc = self._lib.c.b200fft_plan_describe(self._h, buf, len(buf))
        if rc < 0:
            self._lib.check(rc)
        return buf.value.decode()

    def recipe(self) -> "Recipe":
        """The decomposition this plan was built as (b200fft_plan_recipe): feed it to FftPlanner.plan_fft_with_recipe to rebuild it."""
        arr = (_RecipeNode * 8)()
        n = self._lib.c.b200fft_plan_recipe(self._h, arr, 8)
        if n < 0:
            self._lib.check(n)
        rc = None
        for i in range(n - 1, -1, -1):
            rc = Recipe(arr[i].kind, arr[i].len, arr[i].a, arr[i].b, rc if arr[i].child else None)
        return rc

    def launches(self, batch: int) -> int:
        return int(self._lib.c.b200fft_plan_launches(self._h, batch))

    # ---- scratch getters: this backend needs no caller scratch (allowed, src/lib.rs:259-261)
    def get_inplace_scratch_len(self) -> int:
        return int(self._lib.c.b200fft_plan_scratch_len(self._h, 0))

    def get_outofplace_scratch_len(self) -> int:
        return int(self._lib.c.b200fft_plan_scratch_len(self._h, 1))

    def get_immutable_scratch_len(self) -> int:
        return int(self._lib.c.b200fft_plan_scratch_len(self._h, 2))

    # ---- host-slice trait methods ------------------------------------------------------
    def _host(self, a, name: str, writable: bool) -> np.ndarray:
        if not isinstance(a, np.ndarray) or a.dtype != np.dtype(self.dtype) or a.ndim != 1 or not a.flags.c_contiguous:
            raise TypeError(f"{name} must be a contiguous 1-D numpy array of {np.dtype(self.dtype).name}")
        if writable and not a.flags.writeable:
            raise TypeError(f"{name} must be writable")
        return a

    def _check_scratch(self, scratch, need: int) -> None:
        # "Not enough scratch space was provided..." (src/common.rs:32-37); need is 0 here, so any
        # scratch (even dirty, src/test_utils.rs:131-141) is accepted and ignored.
        if scratch is not None and len(scratch) < need:
            raise FftError(-9, f"Not enough scratch space was provided. Expected scratch len >= {need}, "
                               f"got scratch len = {len(scratch)}")

    def process(self, buffer: np.ndarray) -> None:
        """In place over every contiguous chunk of len() elements (src/lib.rs:195-198)."""
        self.process_with_scratch(buffer, None)

    def process_with_scratch(self, buffer: np.ndarray, scratch=None) -> None:
        buffer = self._host(buffer, "buffer", True)
        self._check_scratch(scratch, self.get_inplace_scratch_len())
        self._lib.check(self._lib.c.b200fft_exec_host_inplace(self._h, buffer.ctypes.data, buffer.size))

    def process_outofplace_with_scratch(self, input: np.ndarray, output: np.ndarray, scratch=None) -> None:
        """input may be used as scratch by the reference (src/lib.rs:213-236); here it is left intact."""
        input = self._host(input, "input", True)
        output = self._host(output, "output", True)
        self._check_scratch(scratch, self.get_outofplace_scratch_len())
        if input.size != output.size:
            raise FftError(-6, "Provided FFT input buffer and output buffer must have the same length. "
                               f"Got input.len() = {input.size}, output.len() = {output.size}")
        self._lib.check(self._lib.c.b200fft_exec_host_outofplace(self._h, input.ctypes.data, output.ctypes.data, input.size))

    def process_immutable_with_scratch(self, input: np.ndarray, output: np.ndarray, scratch=None) -> None:
        input = self._host(input, "input", False)
        output = self._host(output, "output", True)
        self._check_scratch(scratch, self.get_immutable_scratch_len())
        if input.size != output.size:
            raise FftError(-6, "Provided FFT input buffer and output buffer must have the same length. "
                               f"Got input.len() = {input.size}, output.len() = {output.size}")
        self._lib.check(self._lib.c.b200fft_exec_host_outofplace(self._h, input.ctypes.data, output.ctypes.data, input.size))

    # ---- device-resident path (no reference equivalent; the measured one) --------------
    def workspace_bytes(self, batch: int) -> int:
        return int(self._lib.c.b200fft_workspace_bytes(self._h, batch))

    def process_device_ptr(self, d_in: int, d_out: int, batch: int, stream: int = 0,
                           workspace: int = 0, workspace_bytes: int = 0) -> None:
        """Raw pointers on the plan's device, batch*len() elements each, async on `stream`."""
        if workspace:
            rc = self._lib.c.b200fft_exec_device_ws(self._h, d_in, d_out, batch, stream, workspace, workspace_bytes)
        else:
            rc = self._lib.c.b200fft_exec_device(self._h, d_in, d_out, batch, stream)
        self._lib.check(rc)

    def process_device(self, x, out=None, workspace=None):
        """x: torch complex tensor on the plan's device holding batch*len() elements (any shape,
        contiguous).  In place when `out` is None.  Asynchronous on torch's current stream."""
        import torch

        want = torch.complex64 if self._precision == F32 else torch.complex128
        if x.dtype != want or not x.is_cuda or not x.is_contiguous():
            raise TypeError(f"process_device wants a contiguous CUDA tensor of {want}")
        if x.device.index != self.device:
            raise FftError(-1, f"tensor is on cuda:{x.device.index}, plan is on cuda:{self.device}")
        dst = x if out is None else out
        if dst.dtype != want or dst.numel() != x.numel() or not dst.is_contiguous() or dst.device != x.device:
            raise FftError(-6, "Provided FFT input buffer and output buffer must have the same length. "
                               f"Got input.len() = {x.numel()}, output.len() = {dst.numel()}")
        n = x.numel()
        if self._len == 0 or n == 0:
            return dst
        if n < self._len:  # (an empty buffer is zero chunks and validates, src/array_utils.rs:151-177)
            raise FftError(-4, f"Provided FFT buffer was too small. Expected len = {self._len}, got len = {n}")
        if n % self._len:
            raise FftError(-5, "Input FFT buffer must be a multiple of FFT length. "
                               f"Expected multiple of {self._len}, got len = {n}")
        stream = torch.cuda.current_stream(x.device).cuda_stream
        ws_ptr, ws_bytes = 0, 0
        if workspace is not None:
            ws_ptr, ws_bytes = workspace.data_ptr(), workspace.numel() * workspace.element_size()
        self.process_device_ptr(x.data_ptr(), dst.data_ptr(), n // self._len, stream, ws_ptr, ws_bytes)
        return dst


class FftPlanner:
    """FftPlanner<T> of the reference (src/plan.rs:67-126) for this backend.

    Like FftPlannerAvx::new() (src/avx/avx_planner.rs:121-164) construction fails when the hardware
    is absent -- the reference would then fall through to its next backend; here it raises, because
    this package has no other backend.  Plans are cached per (len, direction) exactly like
    FftCache (src/fft_cache.rs:5-38): planning the same transform twice returns the same object."""

    def __init__(self, dtype=np.complex64, device: int = 0, lib: Optional[Library] = None):
        dt = np.dtype(dtype)
        if dt == np.dtype(np.float32):
            dt = np.dtype(np.complex64)
        if dt == np.dtype(np.float64):
            dt = np.dtype(np.complex128)
        if dt not in _DTYPES:
            raise TypeError("FftPlanner accelerates f32 and f64 only (src/avx/avx_planner.rs:149-163)")
        self._precision = _DTYPES[dt]
        self._lib = lib if lib is not None else default_library()
        if self._lib.device_count() <= 0:
            raise FftError(-2, "no sm_100 CUDA device is visible (there is no CPU fallback)")
        self.device = device
        self._cache: Dict[Tuple[int, int], Fft] = {}
        self._lock = threading.Lock()

    def plan_fft(self, len: int, direction: FftDirection) -> Fft:
        key = (int(len), int(direction))
        with self._lock:
            fft = self._cache.get(key)
            if fft is None:
                fft = Fft(self._lib, int(len), FftDirection(direction), self._precision, self.device)
                self._cache[key] = fft
            return fft

    def plan_fft_with_recipe(self, recipe: Recipe, direction: FftDirection) -> Fft:
        """Planning owned by the caller: build exactly the decomposition `recipe` names (not cached)."""
        return Fft(self._lib, recipe.len, FftDirection(direction), self._precision, self.device, recipe=recipe)

    def plan_fft_2d(self, height: int, width: int, direction: FftDirection = FftDirection.Forward) -> Fft2d:
        """2-D transform of [height][width] images: the width-point plan over the rows, one strided pass down the columns."""
        return Fft2d(self._lib, height, width, direction, self._precision, self.device)

    def plan_fft_forward(self, len: int) -> Fft:
        return self.plan_fft(len, FftDirection.Forward)

    def plan_fft_inverse(self, len: int) -> Fft:
        return self.plan_fft(len, FftDirection.Inverse)


class Fft2d:
    """2-D complex transform of row-major [height][width] images (a batch of them, contiguous): unnormalised, forward sign as in 1-D.
    numpy arrays go through the synchronous host entry point (in place), torch CUDA tensors through the device one (in place or into
    `out`, asynchronous on torch's current stream)."""

    def __init__(self, lib: Library, height: int, width: int, direction: FftDirection, precision: int, device: int):
        self._lib, self.height, self.width, self._precision, self.device = lib, int(height), int(width), precision, device
        self._direction = FftDirection(direction)
        self._h = ctypes.c_void_p()
        lib.check(lib.c.b200fft_plan2d_create(ctypes.byref(self._h), self.height, self.width, int(direction), precision, device))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self._lib.c.b200fft_plan2d_destroy(h)
            except Exception:
                pass

    def fft_direction(self) -> FftDirection:
        return self._direction

    def process(self, buffer: np.ndarray) -> None:
        want = np.complex64 if self._precision == F32 else np.complex128
        if buffer.dtype != want or not buffer.flags.c_contiguous or not buffer.flags.writeable:
            raise TypeError(f"Fft2d.process wants a contiguous writable {np.dtype(want)} array")
        per = self.height * self.width
        if buffer.size % per:
            raise FftError(-5, f"Input FFT buffer must be a multiple of FFT length. Expected multiple of {per}, got len = {buffer.size}")
        self._lib.check(self._lib.c.b200fft_exec2d_host(self._h, buffer.ctypes.data, buffer.ctypes.data, buffer.size // per))

    def process_device(self, x, out=None):
        import torch

        want = torch.complex64 if self._precision == F32 else torch.complex128
        dst = x if out is None else out
        if x.dtype != want or dst.dtype != want or not x.is_cuda or not x.is_contiguous() or not dst.is_contiguous() or dst.numel() != x.numel():
            raise TypeError(f"Fft2d.process_device wants contiguous CUDA tensors of {want} with equal sizes")
        per = self.height * self.width
        if x.numel() % per:
            raise FftError(-5, f"Input FFT buffer must be a multiple of FFT length. Expected multiple of {per}, got len = {x.numel()}")
        self._lib.check(self._lib.c.b200fft_exec2d_device(self._h, x.data_ptr(), dst.data_ptr(), x.numel() // per,
                                                          torch.cuda.current_stream(x.device).cuda_stream))
        return dst


class RealFft:
    """Real-to-complex / complex-to-real transforms of one even length (the shape of the `realfft` crate's RealToComplex /
    ComplexToReal on top of RustFFT's Fft).  forward: batch * len reals -> batch * (len/2 + 1) complex;
    inverse: the reverse, unnormalised (inverse(forward(x)) == len * x).  numpy arrays go through the synchronous host entry
    points, torch CUDA tensors through the device ones (asynchronous on torch's current stream)."""

    def __init__(self, lib: Library, length: int, precision: int, device: int):
        self._lib, self._len, self._precision, self.device = lib, int(length), precision, device
        self._h = ctypes.c_void_p()
        lib.check(lib.c.b200fft_real_plan_create(ctypes.byref(self._h), self._len, precision, device))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self._lib.c.b200fft_real_plan_destroy(h)
            except Exception:
                pass

    def len(self) -> int:
        return self._len

    def complex_len(self) -> int:
        return self._len // 2 + 1

    def _dtypes(self):
        return (np.float32, np.complex64) if self._precision == F32 else (np.float64, np.complex128)

    def _run(self, inverse: bool, src, dst):
        rdt, cdt = self._dtypes()
        n, h = self._len, self._len // 2 + 1
        sdt, ddt, sper, dper = (cdt, rdt, h, n) if inverse else (rdt, cdt, n, h)
        if isinstance(src, np.ndarray):
            if src.dtype != sdt or dst.dtype != ddt or not src.flags.c_contiguous or not dst.flags.c_contiguous or not dst.flags.writeable:
                raise TypeError(f"RealFft wants contiguous {np.dtype(sdt)} input and writable {np.dtype(ddt)} output")
            if src.size % sper or dst.size != src.size // sper * dper:
                raise FftError(-6, f"RealFft: input holds {src.size} elements, output {dst.size}: expected batch * {sper} and batch * {dper}")
            fn = self._lib.c.b200fft_real_inverse_host if inverse else self._lib.c.b200fft_real_forward_host
            self._lib.check(fn(self._h, src.ctypes.data, dst.ctypes.data, src.size // sper))
            return dst
        import torch

        tmap = {np.float32: torch.float32, np.float64: torch.float64, np.complex64: torch.complex64, np.complex128: torch.complex128}
        if src.dtype != tmap[sdt] or dst.dtype != tmap[ddt] or not src.is_cuda or not dst.is_cuda or not src.is_contiguous() or not dst.is_contiguous():
            raise TypeError("RealFft wants contiguous CUDA tensors of the plan's real / complex dtypes")
        if src.numel() % sper or dst.numel() != src.numel() // sper * dper:
            raise FftError(-6, f"RealFft: input holds {src.numel()} elements, output {dst.numel()}: expected batch * {sper} and batch * {dper}")
        fn = self._lib.c.b200fft_real_inverse_device if inverse else self._lib.c.b200fft_real_forward_device
        self._lib.check(fn(self._h, src.data_ptr(), dst.data_ptr(), src.numel() // sper, torch.cuda.current_stream(src.device).cuda_stream))
        return dst

    def forward(self, real_in, complex_out):
        return self._run(False, real_in, complex_out)

    def inverse(self, complex_in, real_out):
        return self._run(True, complex_in, real_out)


class RealFftPlanner:
    """Plans RealFft instances (cached per length), like realfft::RealFftPlanner over rustfft::FftPlanner."""

    def __init__(self, dtype=np.float32, device: int = 0, lib: Optional[Library] = None):
        dt = np.dtype(dtype)
        if dt in (np.dtype(np.float32), np.dtype(np.complex64)):
            self._precision = F32
        elif dt in (np.dtype(np.float64), np.dtype(np.complex128)):
            self._precision = F64
        else:
            raise TypeError("RealFftPlanner accelerates f32 and f64 only")
        self._lib = lib if lib is not None else default_library()
        if self._lib.device_count() <= 0:
            raise FftError(-2, "no sm_100 CUDA device is visible (there is no CPU fallback)")
        self.device = device
        self._cache: Dict[int, RealFft] = {}
        self._lock = threading.Lock()

    def plan_fft(self, len: int) -> RealFft:
        with self._lock:
            f = self._cache.get(int(len))
            if f is None:
                f = self._cache[int(len)] = RealFft(self._lib, int(len), self._precision, self.device)
            return f


def shard_range(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous batch shard [lo, hi) of `rank` (remainder to the low ranks): transforms in a
    batch are independent (src/array_utils.rs:164-170 is a plain loop), so a batch shards across
    GPUs with no data-path collective."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
