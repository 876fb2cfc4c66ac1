"""The networks of the reference's example notebooks as layer lists for HealpyGCNN (same layers, same arguments):

* quick_start_layers        examples/quick_start.ipynb:118-127   (classifier; nside 64, n_neighbors 20, batch 16)
* advanced_tutorial_layers  examples/advanced_tutorial.ipynb:309-325 (masked survey; Chebyshev, Monomial, residual layer)
* autoencoder_layers        examples/generative_models.ipynb:185-213 (encoder / decoder with pseudo-convolutions)
* regression_layers         the full-sphere regression pattern of SURVEY 8d C5 (no notebook: BASELINE.json configs[4])

BASELINE.json scales these to nside 64 / 512 / 128 / 1024; bench.py and the GPU tests build
them through these functions.  The Lambda heads carry `_oracle_kind` so that the test oracle can restate them."""
from . import healpy_layers as hp_layer
from . import keras_compat as kc


def _mean_softmax():
    import torch

    head = kc.Lambda(lambda x: torch.softmax(x.mean(dim=1), dim=-1))
    head._oracle_kind = "mean_softmax"
    return head


def quick_start_layers(mode=None):
    kw = dict(K=10, Fout=5, use_bias=True, use_bn=True, activation="relu")
    if mode is not None:
        kw["mode"] = mode
    layers = []
    for _ in range(3):
        layers += [hp_layer.HealpyChebyshev(**kw), hp_layer.HealpyPool(p=1)]
    layers += [hp_layer.HealpyChebyshev(K=10, Fout=2, **({"mode": mode} if mode is not None else {})), _mean_softmax()]
    return layers


def advanced_tutorial_layers(mode=None):
    m = {"mode": mode} if mode is not None else {}
    return [hp_layer.HealpyChebyshev(K=10, Fout=5, use_bias=True, use_bn=True, activation="relu", **m),
            hp_layer.HealpyPool(p=1, pool_type="MAX"),
            hp_layer.HealpyMonomial(K=10, Fout=5, use_bias=True, use_bn=True, activation="relu", **m),
            hp_layer.HealpyPool(p=1, pool_type="AVG"),
            hp_layer.Healpy_ResidualLayer(layer_type="CHEBY",
                                          layer_kwargs={"K": 10, "activation": "relu", "use_bn": True, "use_bias": True, **m},
                                          use_bn=False, activation="relu", alpha=0.1),
            hp_layer.HealpyPseudoConv(Fout=2, p=1, activation="relu"),
            _mean_softmax()]


def autoencoder_layers(K=5, mode=None):
    """(encoder_layers, decoder_layers) of generative_models.ipynb:185-213."""
    m = {"mode": mode} if mode is not None else {}

    def cheb(act="elu", Fout=16):
        return hp_layer.HealpyChebyshev(K=K, Fout=Fout, use_bias=True, use_bn=False, activation=act, **m)

    def ln():
        return kc.LayerNormalization(axis=1)

    enc = [hp_layer.HealpyPseudoConv(p=1, Fout=4, activation="elu"), hp_layer.HealpyPseudoConv(p=1, Fout=8, activation="elu"),
           hp_layer.HealpyPseudoConv(p=1, Fout=16, activation="elu"), cheb(), ln(), cheb(), ln(), cheb("linear")]
    dec = [cheb(), ln(), cheb(), ln(), cheb(), ln(),
           hp_layer.HealpyPseudoConv_Transpose(p=1, Fout=16, activation="elu"), cheb(),
           hp_layer.HealpyPseudoConv_Transpose(p=1, Fout=16, activation="elu"), cheb(),
           hp_layer.HealpyPseudoConv_Transpose(p=1, Fout=1, activation="linear")]
    return enc, dec
