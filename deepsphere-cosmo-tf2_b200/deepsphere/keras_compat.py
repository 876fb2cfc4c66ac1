"""A Keras-shaped surface on torch.nn.Module.

The reference's layers are ``tf.keras.Model`` subclasses driven through the Keras protocol
(``build(input_shape)``, ``call(x, training=False)``, ``add_weight``, auto-generated
snake_case names, ``Sequential``).  Keras/TensorFlow are not installable in this image,
so the host side mirrors that protocol on ``torch.nn.Module``: same constructor
arguments, weight names / shapes / initialisers, error types and layer names, so that
reference user code and reference tests read the same.  Only what the hot path's callers
use is provided.
"""

import math
import re
from collections import OrderedDict

import numpy as np
import torch

from . import _native as nat

floatx = torch.float32

# --------------------------------------------------------------------------------------
# names
# --------------------------------------------------------------------------------------

_name_counts = {}


def to_snake_case(name):
    """Keras' layer auto-naming ("GCNN_ResidualLayer" -> "gcnn__residual_layer")."""
    intermediate = re.sub("(.)([A-Z][a-z0-9]+)", r"\1_\2", name)
    insecure = re.sub("([a-z])([A-Z])", r"\1_\2", intermediate).lower()
    return "private" + insecure if insecure[0] == "_" else insecure


def unique_name(base, scope=None):
    counts = _name_counts if scope is None else scope
    n = counts.get(base, 0)
    counts[base] = n + 1
    return base if n == 0 else f"{base}_{n}"


def reset_name_counts():
    _name_counts.clear()


# --------------------------------------------------------------------------------------
# initializers (tf.keras.initializers semantics)
# --------------------------------------------------------------------------------------


def _fans(shape):
    shape = tuple(int(s) for s in shape)
    if len(shape) < 1:
        return 1, 1
    if len(shape) == 1:
        return shape[0], shape[0]
    if len(shape) == 2:
        return shape[0], shape[1]
    rf = int(np.prod(shape[:-2]))
    return shape[-2] * rf, shape[-1] * rf


class Initializer:
    def __call__(self, shape, dtype=None):
        raise NotImplementedError


class TruncatedNormal(Initializer):
    def __init__(self, mean=0.0, stddev=0.05, seed=None):
        self.mean, self.stddev, self.seed = mean, stddev, seed

    def __call__(self, shape, dtype=None):
        g = None if self.seed is None else torch.Generator().manual_seed(int(self.seed))
        t = torch.empty(tuple(shape), dtype=floatx)
        torch.nn.init.trunc_normal_(
            t, mean=self.mean, std=self.stddev, a=self.mean - 2 * self.stddev, b=self.mean + 2 * self.stddev,
            generator=g,
        )
        return t


class RandomNormal(Initializer):
    def __init__(self, mean=0.0, stddev=0.05, seed=None):
        self.mean, self.stddev, self.seed = mean, stddev, seed

    def __call__(self, shape, dtype=None):
        g = None if self.seed is None else torch.Generator().manual_seed(int(self.seed))
        return torch.randn(tuple(shape), generator=g, dtype=floatx) * self.stddev + self.mean


class GlorotUniform(Initializer):
    def __init__(self, seed=None):
        self.seed = seed

    def __call__(self, shape, dtype=None):
        fan_in, fan_out = _fans(shape)
        limit = math.sqrt(6.0 / (fan_in + fan_out))
        g = None if self.seed is None else torch.Generator().manual_seed(int(self.seed))
        return (torch.rand(tuple(shape), generator=g, dtype=floatx) * 2 - 1) * limit


class Zeros(Initializer):
    def __call__(self, shape, dtype=None):
        return torch.zeros(tuple(shape), dtype=floatx)


class Ones(Initializer):
    def __call__(self, shape, dtype=None):
        return torch.ones(tuple(shape), dtype=floatx)


class Constant(Initializer):
    def __init__(self, value=0.0):
        self.value = value

    def __call__(self, shape, dtype=None):
        return torch.as_tensor(np.broadcast_to(np.asarray(self.value, dtype=np.float32), tuple(shape)).copy())


_INITIALIZERS = {
    "glorot_uniform": GlorotUniform,
    "zeros": Zeros,
    "ones": Ones,
    "truncated_normal": TruncatedNormal,
    "random_normal": RandomNormal,
}


def get_initializer(spec):
    """Keras ``initializers.get``: None -> glorot_uniform (the add_weight default)."""
    if spec is None:
        return GlorotUniform()
    if isinstance(spec, str):
        if spec not in _INITIALIZERS:
            raise ValueError(f"Unknown initializer: {spec}")
        return _INITIALIZERS[spec]()
    if callable(spec):
        return spec
    raise ValueError(f"Could not interpret initializer: {spec}")


# --------------------------------------------------------------------------------------
# activations (tf.keras.activations names; gnn_layers.py:55-60)
# --------------------------------------------------------------------------------------


def _linear(x):
    return x


linear = _linear

# name -> (fused C-ABI activation id or None, torch callable)
ACTIVATIONS = {
    "linear": (nat.ACT_LINEAR, _linear),
    "relu": (nat.ACT_RELU, torch.relu),
    "elu": (nat.ACT_ELU, torch.nn.functional.elu),
    "sigmoid": (nat.ACT_SIGMOID, torch.sigmoid),
    "tanh": (nat.ACT_TANH, torch.tanh),
    "softplus": (nat.ACT_SOFTPLUS, torch.nn.functional.softplus),
    "selu": (None, torch.nn.functional.selu),
    "softsign": (None, torch.nn.functional.softsign),
    "swish": (None, torch.nn.functional.silu),
    "silu": (None, torch.nn.functional.silu),
    "gelu": (None, torch.nn.functional.gelu),
    "exponential": (None, torch.exp),
    "relu6": (None, torch.nn.functional.relu6),
    "leaky_relu": (None, lambda x: torch.nn.functional.leaky_relu(x, 0.2)),
    "softmax": (None, lambda x: torch.softmax(x, dim=-1)),
    "hard_sigmoid": (None, lambda x: torch.clamp(x / 6.0 + 0.5, 0.0, 1.0)),
    "mish": (None, torch.nn.functional.mish),
}


def resolve_activation(activation):
    """Returns (fused_id or None, callable or None) following gnn_layers.py:55-60:
    None or a callable is taken as is; a string must name a tf.keras activation, else
    ValueError."""
    if activation is None:
        return nat.ACT_LINEAR, None
    if callable(activation):
        for fused, fn in ACTIVATIONS.values():
            if fn is activation and fused is not None:
                return fused, fn
        return None, activation
    if isinstance(activation, str) and activation in ACTIVATIONS:
        return ACTIVATIONS[activation]
    raise ValueError(f"Could not find activation <{activation}> in tf.keras.activations...")


# --------------------------------------------------------------------------------------
# Model base
# --------------------------------------------------------------------------------------


def as_tensor(x, device=None):
    """numpy / list / torch -> float32 torch tensor on the compute device."""
    if isinstance(x, torch.Tensor):
        t = x if x.dtype == floatx else x.to(floatx)
    else:
        t = torch.as_tensor(np.asarray(x), dtype=floatx)
    if device is not None and t.device != device:
        t = t.to(device)
    return t


def default_device():
    if torch.cuda.is_available():
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


class Model(torch.nn.Module):
    """tf.keras.Model-shaped base: lazy ``build``, ``call(x, training=False)``, ``add_weight``."""

    def __init__(self, name=None):
        super().__init__()
        self.built = False
        self._auto_name = name
        self._regularizers = []

    # -- naming (assigned when the layer is inserted into a Sequential, like Keras) --
    @property
    def name(self):
        if self._auto_name is None:
            self._auto_name = unique_name(to_snake_case(type(self).__name__))
        return self._auto_name

    def add_weight(self, name, shape, initializer=None, trainable=True, regularizer=None, constraint=None,
                   dtype=None, **unused):
        init = get_initializer(initializer)
        value = init(tuple(int(s) for s in shape), dtype=dtype)
        value = as_tensor(value).reshape(tuple(int(s) for s in shape)).clone()
        param = torch.nn.Parameter(value.to(default_device()), requires_grad=bool(trainable))
        self.register_parameter(name, param)
        if regularizer is not None:
            self._regularizers.append((regularizer, param))
        return param

    @property
    def losses(self):
        """Regularisation terms of this layer AND of every layer below it (Keras aggregates the `losses` of nested
        layers: a Sequential / HealpyGCNN / residual layer built from layers with `regularizer=` kwargs reports them
        all), one entry per (regularizer, weight) pair."""
        out, seen = [], set()
        for m in self.modules():
            for reg, p in getattr(m, "_regularizers", ()):
                key = (id(reg), id(p))
                if key not in seen:
                    seen.add(key)
                    out.append(reg(p))
        return out

    def build(self, input_shape):
        pass

    def call(self, inputs, *args, **kwargs):
        raise NotImplementedError

    def compute_output_shape(self, input_shape):
        """Static shape inference (no device work); default: shape-preserving layer."""
        return tuple(input_shape)

    def build_from_shape(self, input_shape):
        """Create the weights for `input_shape` without running the layer (Keras' build step)."""
        if not self.built:
            self.build(tuple(input_shape))
            self.built = True
        return self.compute_output_shape(tuple(input_shape))

    def _maybe_build(self, x):
        if not self.built:
            self.build(tuple(x.shape))
            self.built = True

    def forward(self, inputs, *args, **kwargs):
        x = as_tensor(inputs, default_device() if not isinstance(inputs, torch.Tensor) else None)
        self._maybe_build(x)
        return self.call(x, *args, **kwargs)

    # -- Keras-style accessors --
    @property
    def trainable_variables(self):
        return [p for p in self.parameters() if p.requires_grad]

    @property
    def weights(self):
        return list(self.parameters()) + list(self.buffers())

    def count_params(self):
        return int(sum(p.numel() for p in self.parameters()) + sum(b.numel() for b in self.buffers()))

    def get_weights(self):
        return [w.detach().cpu().numpy() for w in self.weights]

    def set_weights(self, values):
        ws = self.weights
        if len(ws) != len(values):
            raise ValueError(f"expected {len(ws)} arrays, got {len(values)}")
        with torch.no_grad():
            for w, v in zip(ws, values):
                w.copy_(torch.as_tensor(np.asarray(v)).reshape(w.shape))


def _accepts_training(layer):
    import inspect

    try:
        fn = layer.call if isinstance(layer, Model) else layer.forward if isinstance(layer, torch.nn.Module) else layer
        return "training" in inspect.signature(fn).parameters
    except (TypeError, ValueError):
        return False


# --------------------------------------------------------------------------------------
# the few stock Keras layers the reference's example networks put around the hot path
# --------------------------------------------------------------------------------------


class Lambda(Model):
    def __init__(self, function, name=None):
        super().__init__(name=name)
        self.function = function

    def call(self, x):
        return self.function(x)

    def compute_output_shape(self, input_shape):
        with torch.no_grad():
            return tuple(self.function(torch.zeros(tuple(input_shape), dtype=floatx)).shape)


class Flatten(Model):
    def call(self, x):
        return x.reshape(x.shape[0], -1)

    def compute_output_shape(self, input_shape):
        return (input_shape[0], int(np.prod(input_shape[1:])))


class Dense(Model):
    """tf.keras.layers.Dense: kernel [Fin, units] glorot_uniform, bias zeros."""

    def __init__(self, units, activation=None, use_bias=True, kernel_initializer=None, name=None):
        super().__init__(name=name)
        self.units = int(units)
        self.use_bias = use_bias
        self.kernel_initializer = kernel_initializer
        _, self._act = resolve_activation(activation)

    def build(self, input_shape):
        self.kernel = self.add_weight("kernel", [int(input_shape[-1]), self.units], self.kernel_initializer)
        if self.use_bias:
            self.bias = self.add_weight("bias", [self.units], "zeros")

    def call(self, x):
        y = x @ self.kernel
        if self.use_bias:
            y = y + self.bias
        return y if self._act is None else self._act(y)

    def compute_output_shape(self, input_shape):
        return tuple(input_shape[:-1]) + (self.units,)


class BatchNormalization(Model):
    """tf.keras.layers.BatchNormalization over the last axis (statistics over all other
    axes), moving statistics updated with `momentum`; gnn_layers.py:53 uses
    (momentum=0.9, epsilon=1e-5, center=False, scale=False)."""

    def __init__(self, axis=-1, momentum=0.99, epsilon=1e-3, center=True, scale=True, name=None):
        super().__init__(name=name)
        if axis != -1:
            raise NotImplementedError("BatchNormalization shim supports axis=-1 only")
        self.momentum, self.epsilon, self.center, self.scale = momentum, epsilon, center, scale
        # statistics of the GLOBAL batch when a process group is initialised (True), of this process only (False), or
        # over a given group
        self.sync_group = True

    def build(self, input_shape):
        F = int(input_shape[-1])
        if self.scale:
            self.gamma = self.add_weight("gamma", [F], "ones")
        if self.center:
            self.beta = self.add_weight("beta", [F], "zeros")
        self.register_buffer("moving_mean", torch.zeros(F, device=default_device()))
        self.register_buffer("moving_variance", torch.ones(F, device=default_device()))

    def call(self, x, training=False):
        dims = tuple(range(x.dim() - 1))
        if training:
            # batch sharded over ranks: statistics of the GLOBAL batch (one small all-reduce), see distributed.py
            from .distributed import batch_statistics

            mean, var = batch_statistics(x, dims, group=None if self.sync_group in (True, False) else self.sync_group,
                                         sync=self.sync_group is not False)
            with torch.no_grad():
                self.moving_mean.mul_(self.momentum).add_(mean.detach() * (1 - self.momentum))
                self.moving_variance.mul_(self.momentum).add_(var.detach() * (1 - self.momentum))
        else:
            mean, var = self.moving_mean, self.moving_variance
        y = (x - mean) * torch.rsqrt(var + self.epsilon)
        if self.scale:
            y = y * self.gamma
        if self.center:
            y = y + self.beta
        return y


class LayerNormalization(Model):
    """tf.keras.layers.LayerNormalization(axis=...), gamma/beta shaped like the normalised axes."""

    def __init__(self, axis=-1, epsilon=1e-3, center=True, scale=True, name=None):
        super().__init__(name=name)
        self.axis = (axis,) if isinstance(axis, int) else tuple(axis)
        self.epsilon, self.center, self.scale = epsilon, center, scale

    def build(self, input_shape):
        nd = len(input_shape)
        self._axes = tuple(sorted(a % nd for a in self.axis))
        pshape = [int(input_shape[a]) for a in self._axes]
        self._bshape = [int(input_shape[a]) if a in self._axes else 1 for a in range(nd)]
        if self.scale:
            self.gamma = self.add_weight("gamma", pshape, "ones")
        if self.center:
            self.beta = self.add_weight("beta", pshape, "zeros")

    def call(self, x):
        mean = x.mean(dim=self._axes, keepdim=True)
        var = x.var(dim=self._axes, unbiased=False, keepdim=True)
        y = (x - mean) * torch.rsqrt(var + self.epsilon)
        if self.scale:
            y = y * self.gamma.reshape(self._bshape)
        if self.center:
            y = y + self.beta.reshape(self._bshape)
        return y


# --------------------------------------------------------------------------------------
# Sequential
# --------------------------------------------------------------------------------------


class Sequential(Model):
    """tf.keras.Sequential-shaped container: ``layers``, ``get_layer``, ``summary``,
    ``save_weights`` / ``load_weights``; ``training`` is forwarded to layers whose
    ``call`` takes it."""

    def __init__(self, layers=None, name=None):
        super().__init__(name=name if name is not None else "sequential")
        self._layers = torch.nn.ModuleList()
        self._callables = []
        self._name_scope = {}
        self._summary_shapes = None
        for layer in layers or []:
            self.add(layer)

    def add(self, layer):
        if isinstance(layer, torch.nn.Module):
            if isinstance(layer, Model):
                if layer._auto_name is None:
                    layer._auto_name = unique_name(to_snake_case(type(layer).__name__), self._name_scope)
            self._layers.append(layer)
            self._callables.append(layer)
        elif callable(layer):
            wrapped = Lambda(layer)
            wrapped._auto_name = unique_name("lambda", self._name_scope)
            self._layers.append(wrapped)
            self._callables.append(wrapped)
        else:
            raise TypeError(f"cannot add {layer!r} to a Sequential model")

    @property
    def layers(self):
        return list(self._callables)

    def get_layer(self, name=None, index=None):
        if index is not None:
            if name is not None:
                raise ValueError("Provide only a layer name or a layer index.")
            if index >= len(self._callables):
                raise ValueError(f"Was asked to retrieve layer at index {index} but model only has "
                                 f"{len(self._callables)} layers.")
            return self._callables[index]
        for layer in self._callables:
            if getattr(layer, "name", None) == name:
                return layer
        raise ValueError(f"No such layer: {name}.")

    def build(self, input_shape):
        """Keras' ``model.build(input_shape=(None, M, F))``: creates every layer's weights by
        static shape inference (batch None -> 1); no device work."""
        shape = tuple(1 if s is None else int(s) for s in input_shape)
        shapes = []
        for layer in self._callables:
            if isinstance(layer, Model):
                shape = layer.build_from_shape(shape)
            else:  # a plain torch module: infer by running it on zeros (CPU)
                with torch.no_grad():
                    shape = tuple(layer(torch.zeros(shape, dtype=floatx)).shape)
            shapes.append(shape)
        self._summary_shapes = shapes
        self.built = True

    def _run(self, x, training, record=False):
        shapes = []
        for layer in self._callables:
            if _accepts_training(layer):
                x = layer(x, training=training)
            else:
                x = layer(x)
            if record:
                shapes.append(tuple(x.shape) if hasattr(x, "shape") else None)
        if record:
            self._summary_shapes = shapes
        return x

    def _maybe_build(self, x):
        self.built = True

    def call(self, x, training=False):
        return self._run(x, training)

    def predict(self, x, batch_size=32):
        outs = []
        with torch.no_grad():
            for i in range(0, len(x), batch_size):
                outs.append(self(x[i : i + batch_size], training=False).cpu().numpy())
        return np.concatenate(outs, axis=0)

    def summary(self, print_fn=print):
        lines = [f'Model: "{self.name}"', "_" * 65, f"{'Layer (type)':<34}{'Output Shape':<20}{'Param #':>11}",
                 "=" * 65]
        total = 0
        for i, layer in enumerate(self._callables):
            n = layer.count_params() if hasattr(layer, "count_params") else sum(p.numel() for p in layer.parameters())
            total += n
            shp = "?"
            if self._summary_shapes is not None and self._summary_shapes[i] is not None:
                shp = str((None,) + tuple(self._summary_shapes[i][1:]))
            lines.append(f"{getattr(layer, 'name', type(layer).__name__) + ' (' + type(layer).__name__ + ')':<34}"
                         f"{shp:<20}{n:>11}")
        lines += ["=" * 65, f"Total params: {total}", "_" * 65]
        for ln in lines:
            print_fn(ln)
        return total

    # -- checkpointing: Keras' save_weights/load_weights contract (names, shapes, order),
    #    stored as .npz because h5py is not available in this image --
    def _named_weights(self):
        out = OrderedDict()
        for layer in self._callables:
            lname = getattr(layer, "name", type(layer).__name__)
            for wname, w in list(layer.named_parameters()) + list(layer.named_buffers()):
                out[f"{lname}/{wname}"] = w
        return out

    def save_weights(self, path):
        arrays = {k: v.detach().cpu().numpy() for k, v in self._named_weights().items()}
        with open(path, "wb") as f:
            np.savez(f, **arrays)

    def load_weights(self, path):
        with np.load(path) as data:
            named = self._named_weights()
            missing = [k for k in named if k not in data]
            if missing:
                raise ValueError(f"weights file lacks {missing}")
            with torch.no_grad():
                for k, w in named.items():
                    w.copy_(torch.as_tensor(data[k]).reshape(w.shape))
