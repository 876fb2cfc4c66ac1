"""TEST INFRASTRUCTURE (same rules as deepsphere_oracle.py: only tests/, bench.py's CPU legs and smoke() may import it).

Reads a model built with the product package (deepsphere-cosmo-tf2_b200/deepsphere) by its public attributes and returns
the layer specs of oracle.torch_cpu_network with float64 CPU copies of the weights (requires_grad), plus the list of
(product parameter, oracle tensor) pairs for gradient comparisons.  Nothing here is called by the product."""
import numpy as np
from scipy import sparse

_ACT_NAMES = ("relu", "elu", "sigmoid", "tanh", "softplus")


def _act_name(layer):
    """Name of the layer's activation as the oracle knows it (None = linear)."""
    fn = getattr(layer, "activation", None)
    if fn is None:
        return None
    from deepsphere import keras_compat as kc

    for name in _ACT_NAMES:
        if kc.ACTIVATIONS.get(name, (None, None))[1] is fn:
            return name
    if kc.ACTIVATIONS.get("linear", (None, None))[1] is fn:
        return None
    return fn  # an arbitrary torch callable works on CPU tensors too


def _Lt(layer):
    M = int(layer._L_shape[0])
    return sparse.csr_matrix((layer._L_values.astype(np.float64), (layer._L_indices[:, 0], layer._L_indices[:, 1])),
                             shape=(M, M))


def _w(t, pairs):
    c = t.detach().double().cpu().clone().requires_grad_(True)
    pairs.append((t, c))
    return c


def _conv_spec(layer, pairs):
    from deepsphere import gnn_layers

    rec = "monomial" if isinstance(layer, gnn_layers.Monomial) else "chebyshev"
    p = dict(Lt=_Lt(layer), K=layer.K, recursion=rec, kernel=_w(layer.kernel, pairs),
             bias=_w(layer.bias, pairs) if layer.use_bias else None, activation=_act_name(layer), use_bn=layer.use_bn)
    if layer.use_bn:
        p["moving_mean"] = layer.bn.moving_mean.detach().double().cpu().clone()
        p["moving_var"] = layer.bn.moving_variance.detach().double().cpu().clone()
    return p


def specs_from_layers(layers):
    """layers: the built layers of a HealpyGCNN (model.layers) or any list of product layers / marked Lambdas
    (`lambda_layer._oracle_kind = "mean" | "mean_softmax"`)."""
    from deepsphere import gnn_layers, healpy_layers as hl, keras_compat as kc

    specs, pairs = [], []
    for layer in layers:
        if isinstance(layer, gnn_layers.Bernstein):
            raise NotImplementedError("oracle network: Bernstein layers are checked by their own tests")
        if isinstance(layer, (gnn_layers.Chebyshev, gnn_layers.Monomial)):
            specs.append(("conv", _conv_spec(layer, pairs)))
        elif isinstance(layer, gnn_layers.GCNN_ResidualLayer):
            l1, l2 = layer.layer1, layer.layer2
            s1, s2 = _conv_spec(l1, pairs), _conv_spec(l2, pairs)
            specs.append(("residual", dict(Lt=s1["Lt"], K=l1.K, recursion=s1["recursion"], kernels=[s1["kernel"], s2["kernel"]],
                                           biases=(s1["bias"], s2["bias"]), layer_activation=s1["activation"],
                                           layer_use_bn=l1.use_bn, activation=_act_name(layer), act_before=layer.act_before,
                                           use_bn=layer.use_bn, norm_type=layer.norm_type, alpha=layer.alpha,
                                           layer_moving=((s1.get("moving_mean"), s1.get("moving_var")),
                                                         (s2.get("moving_mean"), s2.get("moving_var"))))))
        elif isinstance(layer, hl.HealpyPool):
            specs.append(("pool", dict(p=layer.p, pool_type=layer.pool_type)))
        elif isinstance(layer, hl.HealpyPseudoConv_Transpose):
            specs.append(("pconvT", dict(kernel=_w(layer.kernel, pairs), bias=_w(layer.bias, pairs), activation=_act_name(layer))))
        elif isinstance(layer, hl.HealpyPseudoConv):
            specs.append(("pconv", dict(kernel=_w(layer.kernel, pairs), bias=_w(layer.bias, pairs), activation=_act_name(layer))))
        elif isinstance(layer, kc.LayerNormalization):
            specs.append(("layernorm", dict(axis=layer.axis, eps=layer.epsilon,
                                            gamma=_w(layer.gamma, pairs) if layer.scale else None,
                                            beta=_w(layer.beta, pairs) if layer.center else None)))
        elif isinstance(layer, kc.Dense):
            specs.append(("dense", dict(kernel=_w(layer.kernel, pairs), bias=_w(layer.bias, pairs) if layer.use_bias else None)))
        elif getattr(layer, "_oracle_kind", None) in ("mean", "mean_softmax"):
            specs.append((layer._oracle_kind, {}))
        else:
            raise NotImplementedError(f"oracle network: no restatement for layer {type(layer).__name__}")
    return specs, pairs
