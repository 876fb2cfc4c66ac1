"""ctypes binding of libdeepsphere_b200.so (include/deepsphere_b200.h).

PyTorch is only the carrier of device memory and streams: every call hands raw device
pointers (``tensor.data_ptr()``) and the current CUDA stream to the C-ABI.  There is no
CPU fallback — if the library is missing, or no CUDA device is visible, the compute
entry points raise.
"""

import ctypes
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG_ROOT = os.path.dirname(_HERE)
# DEEPSPHERE_LIB: alternative build of the same C-ABI (A/B measurements of kernel variants on one GPU box)
LIB_PATH = os.environ.get("DEEPSPHERE_LIB") or os.path.join(_PKG_ROOT, "lib", "libdeepsphere_b200.so")

MODE_FP32, MODE_TF32, MODE_TF32X3 = 0, 1, 2
MODES = {"fp32": MODE_FP32, "tf32": MODE_TF32, "tf32x3": MODE_TF32X3}
RECURSION_CHEBYSHEV, RECURSION_MONOMIAL = 0, 1
ACT_LINEAR, ACT_RELU, ACT_ELU, ACT_SIGMOID, ACT_TANH, ACT_SOFTPLUS = range(6)
POOL_MAX, POOL_AVG = 0, 1

_i32, _i64, _f32, _ptr = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p

# name -> (restype, argtypes); must list every symbol declared in include/deepsphere_b200.h
SIGNATURES = {
    "ds_abi_version": (ctypes.c_int, []),
    "ds_last_error": (ctypes.c_char_p, []),
    "ds_device_count": (ctypes.c_int, []),
    "ds_launch_count": (_i64, []),
    "ds_plan_create_coo": (ctypes.c_int, [_i64, _i64, _ptr, _ptr, _i32, ctypes.POINTER(_ptr)]),
    "ds_plan_destroy": (ctypes.c_int, [_ptr]),
    "ds_plan_info": (ctypes.c_int, [_ptr, _i32, ctypes.POINTER(_i64)]),
    "ds_plan_attach_lattice": (
        ctypes.c_int, [_ptr, _i32, _i32, _i32, _i32, _ptr, _ptr, _ptr, _i64, _ptr, _i64, _ptr]
    ),
    "ds_plan_attach_patches": (ctypes.c_int, [_ptr, _i32, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "ds_spmm": (ctypes.c_int, [_ptr, _i32, _i64, _i64, _ptr, _f32, _ptr, _f32, _ptr, _f32, _ptr, _ptr]),
    "ds_graph_conv_basis_elems": (_i64, [_i64, _i64, _i64, _i32]),
    "ds_graph_conv_forward_writes_basis": (_i32, [_ptr, _i32, _i64, _i64, _i64, _i32]),
    "ds_graph_conv_basis": (ctypes.c_int, [_ptr, _i32, _i32, _i64, _i64, _ptr, _ptr, _i32, _ptr]),
    "ds_graph_conv_forward": (
        ctypes.c_int,
        [_ptr, _i32, _i32, _i64, _i64, _i64, _ptr, _ptr, _ptr, _i32, _ptr, _ptr, _i32, _ptr],
    ),
    "ds_graph_conv_backward_workspace_elems": (_i64, [_i64, _i64, _i64, _i64, _i32, _i32, _i32]),
    "ds_graph_conv_backward": (
        ctypes.c_int,
        [_ptr, _i32, _i32, _i64, _i64, _i64, _ptr, _ptr, _ptr, _ptr, _i32, _ptr, _ptr, _ptr, _ptr, _ptr, _i32, _ptr],
    ),
    "ds_bias_act_forward": (ctypes.c_int, [_i64, _i64, _ptr, _ptr, _i32, _ptr, _ptr]),
    "ds_bias_act_backward": (ctypes.c_int, [_i64, _i64, _ptr, _ptr, _i32, _ptr, _ptr, _ptr, _ptr]),
    "ds_pool_forward": (ctypes.c_int, [_i64, _i64, _i64, _i32, _i32, _ptr, _ptr, _ptr]),
    "ds_pool_backward": (ctypes.c_int, [_i64, _i64, _i64, _i32, _i32, _ptr, _ptr, _ptr, _ptr]),
    "ds_pconv_forward": (ctypes.c_int, [_i64, _i64, _i64, _i64, _i32, _ptr, _ptr, _ptr, _i32, _ptr, _i32, _ptr]),
    "ds_pconv_backward_workspace_elems": (_i64, [_i64, _i64, _i64, _i64, _i32, _i32]),
    "ds_pconv_backward": (
        ctypes.c_int,
        [_i64, _i64, _i64, _i64, _i32, _ptr, _ptr, _ptr, _ptr, _i32, _ptr, _ptr, _ptr, _ptr, _i32, _ptr],
    ),
    "ds_bn_workspace_doubles": (_i64, [_i64, _i64, _i64]),
    "ds_bn_stats": (ctypes.c_int, [_i64, _i64, _i64, _i64, _i64, _ptr, _ptr, _ptr, _ptr]),
    "ds_bn_bias_act_forward": (
        ctypes.c_int,
        [_i64, _i64, _i64, _ptr, _ptr, ctypes.c_double, _ptr, _f32, _f32, _i32, _ptr, _ptr, _ptr, _i32, _ptr, _ptr, _ptr,
         _ptr],
    ),
    "ds_bn_backward_stats": (ctypes.c_int, [_i64, _i64, _i64, _i64, _i64, _ptr, _ptr, _ptr, _ptr, _i32, _ptr, _ptr, _ptr]),
    "ds_bn_backward_apply": (
        ctypes.c_int,
        [_i64, _i64, _i64, _i64, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, ctypes.c_double, _ptr, _i32, _i32, _ptr, _ptr, _ptr],
    ),
    "ds_halo_pack": (ctypes.c_int, [_i64, _i64, _i64, _i64, _ptr, _ptr, _ptr, _ptr]),
    "ds_halo_assemble": (ctypes.c_int, [_i64, _i64, _i64, _i64, _i64, _i64, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "ds_halo_reduce": (ctypes.c_int, [_i64, _i64, _i64, _i64, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "ds_comm_load_nccl": (ctypes.c_int, [ctypes.c_char_p]),
    "ds_comm_unique_id": (ctypes.c_int, [_ptr]),
    "ds_comm_create": (ctypes.c_int, [_i32, _i32, _ptr, ctypes.POINTER(_ptr)]),
    "ds_comm_destroy": (ctypes.c_int, [_ptr]),
    "ds_comm_allreduce_sum": (ctypes.c_int, [_ptr, _ptr, _i64, _i32, _ptr]),
    "ds_comm_alltoallv": (ctypes.c_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "ds_halo_exchange": (ctypes.c_int, [_ptr, _i64, _i64, _i64, _i64, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "ds_halo_exchange_backward": (
        ctypes.c_int, [_ptr, _i64, _i64, _i64, _i64, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr]
    ),
    "ds_pconvT_forward": (ctypes.c_int, [_i64, _i64, _i64, _i64, _i32, _ptr, _ptr, _ptr, _i32, _ptr, _i32, _ptr]),
    "ds_pconvT_backward_workspace_elems": (_i64, [_i64, _i64, _i64, _i64, _i32, _i32]),
    "ds_pconvT_backward": (
        ctypes.c_int,
        [_i64, _i64, _i64, _i64, _i32, _ptr, _ptr, _ptr, _ptr, _i32, _ptr, _ptr, _ptr, _ptr, _i32, _ptr],
    ),
}

_lib = None
_lock = threading.Lock()


class NativeError(RuntimeError):
    """An entry point of libdeepsphere_b200.so reported a failure."""


def lib():
    """Load (once) and return the C-ABI library; fails loudly if it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise NativeError(
                        f"{LIB_PATH} not found: build it with `python {os.path.join(_PKG_ROOT, 'build.py')}` "
                        "(there is no CPU / PyTorch fallback for the hot path)"
                    )
                handle = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(handle, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = handle
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().ds_last_error().decode("utf-8", "replace")
        raise NativeError(f"{what}: {msg}" if what else msg)


def require_cuda():
    if lib().ds_device_count() <= 0:
        raise NativeError("no CUDA device visible: the deepsphere B200 hot path has no CPU fallback")


def launch_count():
    return int(lib().ds_launch_count())


def ptr(t):
    """Raw device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def current_stream():
    import torch

    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class GraphPlan:
    """Device-resident ELL(+tail) form of a rescaled Laplacian and its transpose.

    Built from the same triple the reference stores as tf.constants
    (gnn_layers.py:68-72): COO ``indices`` int64 [nnz, 2], ``values`` float32 [nnz],
    ``shape``.  One plan per device, created lazily on first use."""

    def __init__(self, indices, values, shape, ell_width=0, lattice_builder=None):
        self.indices = np.ascontiguousarray(indices, dtype=np.int64)
        self.values = np.ascontiguousarray(values, dtype=np.float32)
        self.shape = (int(shape[0]), int(shape[1]))
        if self.shape[0] != self.shape[1]:
            raise ValueError(f"the graph Laplacian must be square, got {self.shape}")
        self.ell_width = int(ell_width)
        # optional callable -> dict(n_tiles, LW, H, T, pix, w, sub=(indices, values, n), closure_rows, own_sub)
        # enabling the fused HEALPix lattice kernel (deepsphere/lattice.py); evaluated once, lazily
        self.lattice_builder = lattice_builder
        self._lattice_payload = None
        self._handles = {}

    @property
    def M(self):
        return self.shape[0]

    def handle(self, device_index):
        h = self._handles.get(device_index)
        if h is None:
            import torch

            require_cuda()
            out = ctypes.c_void_p()
            with torch.cuda.device(device_index):
                check(
                    lib().ds_plan_create_coo(
                        self.shape[0],
                        len(self.values),
                        self.indices.ctypes.data_as(ctypes.c_void_p),
                        self.values.ctypes.data_as(ctypes.c_void_p),
                        self.ell_width,
                        ctypes.byref(out),
                    ),
                    "ds_plan_create_coo",
                )
            h = self._handles[device_index] = out
            self._attach_lattice(h, device_index)
        return h

    def _attach_lattice(self, h, device_index):
        if self.lattice_builder is None or os.environ.get("DEEPSPHERE_LATTICE", "1") == "0":
            return
        if self._lattice_payload is None:
            self._lattice_payload = self.lattice_builder() or {}
        p = self._lattice_payload
        if not p:
            return
        try:
            self._attach_lattice_payload(h, device_index, p)
        except NativeError as exc:  # the optional fast path must never break the layer: generic kernels serve it
            import logging

            logging.getLogger("deepsphere").warning(f"lattice plan not attached ({exc}); using the generic kernels")
            self._lattice_payload = {}

    def _attach_lattice_payload(self, h, device_index, p):
        import torch

        def cptr(a):
            return a.ctypes.data_as(ctypes.c_void_p)

        with torch.cuda.device(device_index):
            sub = ctypes.c_void_p()
            n_own = len(p["own_sub"])
            if n_own > 0:
                si, sv, sn = p["sub"]
                check(lib().ds_plan_create_coo(sn, len(sv), cptr(si), cptr(sv), 0, ctypes.byref(sub)),
                      "ds_plan_create_coo (lattice sub-problem)")
            check(
                lib().ds_plan_attach_lattice(
                    h, p["n_tiles"], p["LW"], p["H"], p["T"], cptr(p["pix"]), cptr(p["w"]), sub,
                    len(p["closure_rows"]), cptr(p["closure_rows"]), n_own, cptr(p["own_sub"]),
                ),
                "ds_plan_attach_lattice",
            )
            pt = p.get("patches")
            if pt is not None and os.environ.get("DEEPSPHERE_PATCH", "1") != "0":
                check(
                    lib().ds_plan_attach_patches(
                        h, pt["n_patches"], cptr(pt["row_ptr"]), cptr(pt["rows"]), cptr(pt["ell_col"]),
                        cptr(pt["ell_val"]), cptr(pt["own_ptr"]), cptr(pt["own_local"]),
                    ),
                    "ds_plan_attach_patches",
                )

    def info(self, device_index=0):
        names = ["M", "nnz", "ell_width", "tail_rows", "tail_nnz", "ell_width_T", "tail_rows_T", "device_bytes",
                 "symmetric", "lattice"]
        h = self.handle(device_index)
        out = {}
        for i, n in enumerate(names):
            v = ctypes.c_int64()
            check(lib().ds_plan_info(h, i, ctypes.byref(v)), "ds_plan_info")
            out[n] = int(v.value)
        return out

    def __del__(self):
        try:
            for h in self._handles.values():
                _lib.ds_plan_destroy(h)
        except Exception:
            pass
