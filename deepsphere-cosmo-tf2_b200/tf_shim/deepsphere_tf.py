"""Python side of the TensorFlow custom-op shim (deepsphere_tf_ops.cc): the replacement of
``Chebyshev.call`` lines 113-150 of the reference's src/deepsphere/gnn_layers.py.  Import fails without TensorFlow
(not installable in this repository's image): source-level recipe, exercised by tests only when TF exists.

    from deepsphere_tf import make_plan, graph_conv
    self._plan = make_plan(self._L_indices, self._L_values, self._L_shape)        # once, in Chebyshev.__init__
    x = graph_conv(input_tensor, self.kernel, self._plan, recursion=0, K=self.K)  # in Chebyshev.call
"""
import ctypes
import os

import numpy as np
import tensorflow as tf  # noqa: F401  (ImportError here = the environment has no TensorFlow)

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = ctypes.CDLL(os.path.join(_HERE, "..", "lib", "libdeepsphere_b200.so"))
_ops = tf.load_op_library(os.path.join(_HERE, "deepsphere_tf_ops.so"))

DS_MODE = {"fp32": 0, "tf32": 1, "tf32x3": 2}


def make_plan(indices, values, shape):
    """ds_plan_create_coo from the COO triple of gnn_layers.py:68-72 (int64 [nnz, 2], float32 [nnz], shape)."""
    indices = np.ascontiguousarray(np.asarray(indices), dtype=np.int64)
    values = np.ascontiguousarray(np.asarray(values), dtype=np.float32)
    plan = ctypes.c_void_p()
    _lib.ds_plan_create_coo.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32,
                                        ctypes.POINTER(ctypes.c_void_p)]
    rc = _lib.ds_plan_create_coo(int(shape[0]), len(values), indices.ctypes.data, values.ctypes.data, 0, ctypes.byref(plan))
    if rc != 0:
        _lib.ds_last_error.restype = ctypes.c_char_p
        raise RuntimeError(_lib.ds_last_error().decode())
    return int(plan.value)


def graph_conv(x, kernel, plan, recursion, K, bias=None, act=0, mode="fp32"):
    """y = act(sum_k T_k(L~) x W_k + bias) with the gradient wired through DsGraphConvBackward."""
    b = tf.zeros([0], tf.float32) if bias is None else tf.reshape(bias, [-1])
    attrs = dict(plan=plan, recursion=recursion, k=K, act=act, mode=DS_MODE[mode])

    @tf.custom_gradient
    def _fn(x, kernel, b):
        y, basis = _ops.ds_graph_conv_forward(x, kernel, b, **attrs)

        def grad(dy):
            dx, dk, db = _ops.ds_graph_conv_backward(x, kernel, y, dy, basis, has_bias=bias is not None, **attrs)
            return dx, dk, (tf.reshape(db, tf.shape(b)) if bias is not None else tf.zeros_like(b))

        return y, grad

    return _fn(x, kernel, b)
