#!/usr/bin/env python
"""bench.py — HealpyChebyshev layer microbench (BASELINE.json configs[1]): nside 256, K 5,
Fin = Fout = 64, batch 32 per GPU, forward + backward.

Metric: algorithmic GB/s of the layer (fwd+bwd touches 4*B*M*(3*Fin + 2*Fout)
bytes = 32.21 GB per GPU per step) — `value` with inputs resident in HBM, `e2e` through the
public layer API with pinned HOST buffers and the host<->device copies inside the timed region.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference ...                     # the reference's op sequence on the host
                                                           # CPUs (oracle port, bounded sample)
Under torchrun (N > 1) every rank runs the same per-GPU workload (weak scaling, batch
sharding); the per-step exchange is the all-reduce of the kernel gradient.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "deepsphere-cosmo-tf2_b200"), ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

# the package logger mirrors the reference's (INFO lines to stdout while a HealpyGCNN is built): bench.py's stdout is ONE
# JSON line, so the level goes to WARNING here unless the caller set one
os.environ.setdefault("DEEPSPHERE_LOG_LEVEL", "3")

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "HealpyChebyshev fwd+bwd algorithmic GB/s (nside 256, K 5, Fin=Fout=64, batch 32/GPU)"
# dram__bytes_read.sum + dram__bytes_write.sum of lattice_conv2_kernel from the round-2 `ncu --set full` capture of the
# default build (profiles/r2h_prof_lattice_conv2_raw.csv: batch 8, 1.9007 GB read + 1.5576 GB written per launch), scaled
# to the bench's batch 32 (the kernel's work and traffic are linear in the batch): 13.83 GB vs 12.88 GB algorithmic
TRAFFIC_DEFAULT = int(4 * (1.900713e9 + 1.557593e9))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    # contraction arithmetic (the recursion is always fp32 FMA): north_star names "3xTF32 or TF32" with a stated
    # <= 1e-3 for TF32; tf32 is the mode the register-resident fused kernel (ds_lattice_conv2.cu) serves
    ap.add_argument("--mode", default=os.environ.get("DEEPSPHERE_MODE", "tf32"), choices=["fp32", "tf32", "tf32x3"])
    ap.add_argument("--no-other-modes", action="store_true")
    ap.add_argument("--no-model", action="store_true")
    ap.add_argument("--model-only", action="store_true",
                    help="only the HealpyGCNN training-step measurement; prints its JSON object (used for the "
                         "experimental-kernel child process)")
    ap.add_argument("--no-experimental", action="store_true")
    ap.add_argument("--model-graph", action="store_true",
                    help="with --model-only: capture the training step in a CUDA graph and time graph replays")
    ap.add_argument("--no-f-sweep", action="store_true", help="skip the fused forward at Fin = Fout = 16 / 32")
    ap.add_argument("--no-configs", action="store_true",
                    help="skip the named configurations C1 (quick_start), C3 (masked survey nside 512), C4 (autoencoder)")
    ap.add_argument("--experimental", action="store_true",
                    help="also time the HealpyGCNN step with the opt-in switches (child processes)")
    ap.add_argument("--no-graph", action="store_true", help="do not try CUDA-graph replays of the model training steps")
    ap.add_argument("--no-partitioned", action="store_true",
                    help="skip the nside-1024 sphere-partitioned HealpyGCNN (strong scaling over the ranks)")
    ap.add_argument("--part-nside", type=int, default=1024)
    ap.add_argument("--part-batch", type=int, default=16,
                    help="global batch of the sphere-partitioned nside-1024 network (halved until it fits one GPU's memory)")
    ap.add_argument("--adam", default="fused", choices=["fused", "foreach"],
                    help="torch.optim.Adam implementation of the training-step benches")
    ap.add_argument("--model-nside", type=int, default=256)
    ap.add_argument("--model-batch", type=int, default=16)
    ap.add_argument("--nside", type=int, default=256)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--features", type=int, default=64)
    ap.add_argument("--K", type=int, default=5)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-batch", type=int, default=1)
    return ap.parse_args()


def algorithmic_bytes(B, M, Fin, Fout):
    """fwd: read x, write y; bwd: read x, read dy, write dx (basis recomputed on chip)."""
    return 4 * B * M * (3 * Fin + 2 * Fout)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """Samples SM clock and throttle reasons during the timed region (pynvml)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPUs that are local to GPU `index` (NVML's ideal CPU affinity) BEFORE any pinned host
    buffer is allocated: first-touch then places the buffers on the GPU's own NUMA node.  Round 1's 8-rank e2e leg
    (host -> device -> host of 19 GB per rank and step) ran 4x slower per rank than at N = 1 with every rank's buffers
    wherever the kernel put them.  Returns the CPU list (or None when NVML / the affinity call is unavailable)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_cpu = os.cpu_count() or 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        pass
    return None


def build_layer(args, mode):
    from deepsphere import gnn_layers
    from deepsphere.graph import SphereHealpix

    g = SphereHealpix(args.nside, k=8)
    layer = gnn_layers.Chebyshev(L=g.L, K=args.K, Fout=args.features, mode=mode)
    return g, layer


PARITY_TOL = {"fp32": 1e-5, "tf32": 1e-3, "tf32x3": 2e-5}  # north_star: rel <= 1e-5 fp32 path, stated <= 1e-3 for TF32


def cpu_reference_time(L, args, batch, steps, warmup, Lt=None, sample=None):
    """The reference's op sequence (gnn_layers.py:131-150 + autodiff) with torch CPU ops on all
    host threads — oracle/deepsphere_oracle.py:torch_cpu_graph_conv — fwd + bwd.

    `sample` = dict(x, kernel, dy, y, dx, dkernel) of CPU tensors taken from the GPU layer at the bench configuration
    (the first `batch` maps of the workload): the CPU run then uses those inputs and the third return value is the
    parity of the GPU results against it, max|err| / max|ref| per tensor."""
    from oracle import deepsphere_oracle as orc

    if Lt is None:
        Lt, _ = orc.prepare_laplacian(L, 0.75)
    M = Lt.shape[0]
    F = args.features
    g = torch.Generator().manual_seed(0)
    if sample is not None:
        x = sample["x"].clone().requires_grad_(True)
        w = sample["kernel"].clone().requires_grad_(True)
        dy = sample["dy"]
        batch = x.shape[0]
    else:
        x = torch.randn(batch, M, F, generator=g, requires_grad=True)
        w = (torch.randn(args.K * F, F, generator=g) * 0.075).requires_grad_(True)
        dy = torch.randn(batch, M, F, generator=g)
    times = []
    parity = None
    for i in range(warmup + steps):
        x.grad = w.grad = None
        t0 = time.perf_counter()
        y = orc.torch_cpu_graph_conv(x, Lt, w, args.K, "chebyshev")
        y.backward(dy)
        t1 = time.perf_counter()
        if i >= warmup:
            times.append(t1 - t0)
        if sample is not None and parity is None:
            def rel(a, b):
                return float((a - b).abs().max() / b.abs().max())

            parity = {"y": rel(sample["y"], y.detach()), "dx": rel(sample["dx"], x.grad),
                      "dkernel": rel(sample["dkernel"], w.grad)}
    return float(np.mean(times)), algorithmic_bytes(batch, M, F, F), parity


def _adam(params, args):
    """Adam for the training-step benches: the fused multi-tensor implementation (ONE kernel per step for all parameters
    instead of ~13 foreach launches + one step-counter update per parameter), capturable into the step's CUDA graph."""
    return torch.optim.Adam(params, lr=1e-3, capturable=True, fused=getattr(args, "adam", "fused") == "fused")


def model_train_bench(args, mode, device, world):
    """Second half of BASELINE.json's metric: HealpyGCNN training throughput (maps/s).  The regression network of
    SURVEY 8d config C5 (PseudoConv p=1 F16 -> [Chebyshev K5 F32 + MAX pool] x 3 -> Chebyshev K5 F64 -> pool ->
    mean over pixels -> Dense(2)) on synthetic full-sphere maps, batch sharded over the ranks, MSE loss, Adam,
    one flat gradient all-reduce per step.  nside 256 here, batch sharded; C5's nside 1024 runs on the
    sphere-partitioned path (deepsphere/partition.py, tools/bench_partition.py)."""
    import deepsphere
    from deepsphere import distributed as dsd
    from deepsphere import healpy_layers as hl, keras_compat as kc

    nside, Bm = args.model_nside, args.model_batch
    npix = 12 * nside * nside
    kw = dict(use_bias=True, activation="relu", mode=mode)
    layers = [hl.HealpyPseudoConv(p=1, Fout=16, activation="relu")]
    for _ in range(3):
        layers += [hl.HealpyChebyshev(K=5, Fout=32, **kw), hl.HealpyPool(p=1, pool_type="MAX")]
    layers += [hl.HealpyChebyshev(K=5, Fout=64, **kw), hl.HealpyPool(p=1, pool_type="AVG"),
               kc.Lambda(lambda t: t.mean(dim=1)), kc.Dense(2)]
    torch.manual_seed(11)
    model = deepsphere.HealpyGCNN(nside=nside, indices=np.arange(npix), layers=layers)
    model.build(input_shape=(None, npix, 1))
    dsd.broadcast_parameters(model)
    params = model.trainable_variables
    use_graph = bool(getattr(args, "model_graph", False))
    opt = _adam(params, args)
    gen = torch.Generator(device=device).manual_seed(11 + int(os.environ.get("RANK", "0")))
    x = torch.randn(Bm, npix, 1, device=device, generator=gen)
    t = torch.randn(Bm, 2, device=device, generator=gen)

    def train_step():
        opt.zero_grad(set_to_none=True)
        loss = ((model(x, training=True) - t) ** 2).mean()
        loss.backward()
        dsd.allreduce_gradients(params)
        opt.step()
        return loss

    # fingerprint of the very first step (same seeds on every run): loss and per-parameter gradient norms before any
    # update - what an alternative kernel build must reproduce (bench.py --model-only, experimental_model_run)
    def first():
        opt.zero_grad(set_to_none=True)
        loss0 = ((model(x, training=True) - t) ** 2).mean()
        loss0.backward()
        return {"loss": float(loss0.detach()),
                "grad_norms": [float(p.grad.norm()) if p.grad is not None else 0.0 for p in params]}

    graph = None
    if use_graph:
        # whole-step capture (forward, backward, Adam): the C-ABI calls only enqueue work on the current stream and take
        # their scratch memory from cudaMallocAsync, so the step is capturable; replays remove the host launch path.
        # EVERY backward before the capture runs on a side stream: autograd ties a parameter's AccumulateGrad node to the
        # stream of its first backward, and a node tied to the legacy default stream invalidates the capture (round 1's
        # "cudaErrorStreamCaptureImplicit" - the fingerprint step ran on the default stream)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            first_step = first()
            for _ in range(3):
                train_step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(graph):
            static_loss = train_step()
        run_step = lambda: (graph.replay(), static_loss)[1]  # noqa: E731
    else:
        first_step = first()
        run_step = train_step
        for _ in range(3):
            train_step()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
        torch.cuda.synchronize()
    n = 10
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        loss = run_step()
    b.record()
    torch.cuda.synchronize()
    ms = dsd.allreduce_max(a.elapsed_time(b) / n, device)
    n_params = int(sum(p.numel() for p in params))
    final_loss = float(loss.detach())
    del loss
    # the same step as a CUDA-graph replay (counts only when validated against eager, see graph_replay_timing)
    graph_t = None
    eager_ms = ms
    if not use_graph and not getattr(args, "no_graph", False):
        graph_t = graph_replay_timing(train_step, params, opt, n, device, world)
        if graph_t.get("validated"):
            ms = graph_t["ms_per_step"]
    del graph, model, opt, x, t
    torch.cuda.empty_cache()
    return {"metric": "HealpyGCNN train maps/s", "value": world * Bm / (ms * 1e-3), "unit": "maps/s",
            "ms_per_step": ms, "batch_per_gpu": Bm, "n_gpus": world, "parameters": n_params, "final_loss": final_loss,
            "execution": "CUDA graph replay of the whole step (validated against eager)" if ms != eager_ms else "eager",
            "eager_ms_per_step": eager_ms, "cuda_graph": graph_t,
            "first_step": first_step,
            "config": f"nside {nside} full sphere ({npix} px): PseudoConv p1 F16 -> [Chebyshev K5 F32 + MAX pool] x3 -> "
                      f"Chebyshev K5 F64 -> AVG pool -> mean -> Dense(2); MSE, Adam, fwd+bwd+all-reduce+step, mode {mode}"}


def graph_replay_timing(train_step, params, opt, n, device, world):
    """Capture one whole training step (forward, backward, gradient collectives, Adam) in a CUDA graph and time replays.

    The C-ABI calls only enqueue work on the current stream (scratch memory from cudaMallocAsync) and NCCL collectives are
    capturable, so the step has no host-side dependency; a replay removes the Python / launch path that bounds the small
    levels of a HealpyGCNN.  The number only counts when VALIDATED: from the same weights and optimizer state, 3 eager
    steps and 3 replays must produce the same losses.  Returns a dict (never raises)."""
    from deepsphere import distributed as dsd

    try:
        def snapshot():
            st = []
            for p in params:
                s_ = opt.state.get(p, {})
                st.append((p.detach().clone(), {k: v.detach().clone() for k, v in s_.items() if isinstance(v, torch.Tensor)}))
            return st

        def restore(st):
            with torch.no_grad():
                for p, (w, s_) in zip(params, st):
                    p.copy_(w)
                    for k, v in s_.items():
                        opt.state[p][k].copy_(v)

        torch.cuda.synchronize()
        torch.cuda.empty_cache()   # the capture allocates from its own pool: give the eager pool's cached blocks back first
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):   # warm-up on a side stream: no autograd node may be tied to the legacy stream
            for _ in range(3):
                train_step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        state = snapshot()
        eager_losses = [float(train_step().detach()) for _ in range(3)]
        restore(state)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(graph):
            static_loss = train_step()
        restore(state)   # the capture itself does not run the step, but keep the comparison exact
        replay_losses = []
        for _ in range(3):
            graph.replay()
            replay_losses.append(float(static_loss.detach()))
        rel = max(abs(a - b) / max(abs(a), 1e-12) for a, b in zip(eager_losses, replay_losses))
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            graph.replay()
        b.record()
        torch.cuda.synchronize()
        ms = dsd.allreduce_max(a.elapsed_time(b) / n, device)
        ok = dsd.allreduce_max(0.0 if rel <= 1e-4 else 1.0, device) == 0.0
        out = {"ms_per_step": ms, "validated": bool(ok), "loss_rel_diff_vs_eager": rel, "eager_losses": eager_losses,
               "replay_losses": replay_losses, "final_loss": float(static_loss.detach())}
        del graph
        return out
    except Exception as exc:  # a failed capture must not take the eager measurement with it
        return {"error": str(exc)[:300], "validated": False}


def _c5_layers(mode, mean_layer):
    """PseudoConv p1 F16 -> [Chebyshev K5 F32 + MAX pool] x 4 -> Chebyshev K5 F64 -> AVG pool ->
    mean over pixels -> Dense(2)."""
    from deepsphere import healpy_layers as hl, keras_compat as kc

    kw = dict(use_bias=True, activation="relu", mode=mode)
    layers = [hl.HealpyPseudoConv(p=1, Fout=16, activation="relu")]
    for _ in range(4):
        layers += [hl.HealpyChebyshev(K=5, Fout=32, **kw), hl.HealpyPool(p=1, pool_type="MAX")]
    layers += [hl.HealpyChebyshev(K=5, Fout=64, **kw), hl.HealpyPool(p=1, pool_type="AVG"), mean_layer, kc.Dense(2)]
    return layers


def _step_bytes(model, x):
    """Algorithmic bytes of one training step of a layer stack: every layer reads its input and writes its output in the
    forward (in + out) and reads input + output gradient and writes the input gradient in the backward (2 in + out)."""
    sizes = []
    hooks = [l.register_forward_hook(lambda m, i, o: sizes.append((i[0].numel(), o.numel()))) for l in model.layers]
    with torch.no_grad():
        model(x, training=False)
    for h in hooks:
        h.remove()
    return 4 * sum(3 * a + 2 * b for a, b in sizes)


def named_config_bench(name, args, device, rank, world, hbm_peak):
    """Training throughput of one of BASELINE.json's named configurations, batch sharded over the
    ranks, next to the oracle's torch-CPU restatement of the same network on the host cores (bounded sample, forward +
    backward) and the forward parity of the two on that sample.  Modes: the library default (fp32 contraction) - these
    networks have 1 - 16 channels, their layers run on the generic ELL + tail / fp32 paths."""
    import deepsphere
    from deepsphere import distributed as dsd, example_networks as nets, healpix as hpx, utils
    from oracle import bridge, deepsphere_oracle as orc

    torch.manual_seed(11)
    if name == "C1_quick_start":
        nside, B, k = 64, 16, 20
        idx = np.arange(12 * nside**2)
        model = deepsphere.HealpyGCNN(nside=nside, indices=idx, layers=nets.quick_start_layers(), n_neighbors=k)
        desc = "examples/quick_start.ipynb:118-127,197 as shipped: nside 64, K 10, F 5, BatchNorm, n_neighbors 20, batch 16/GPU"
        loss_fn = lambda y, t: -(torch.log(y.clamp_min(1e-12)) * t).sum(dim=1).mean()  # sparse categorical cross-entropy
        n_out = 2
    elif name == "C3_masked_survey":
        nside, B, k = 512, 8, 20
        idx = utils.extend_indices(hpx.query_disc(nside, [1, 0, 0], 1.5), nside, 64)
        model = deepsphere.HealpyGCNN(nside=nside, indices=idx, layers=nets.advanced_tutorial_layers(), n_neighbors=k)
        desc = ("examples/advanced_tutorial.ipynb:137,211,309-325 scaled to nside 512 (BASELINE.json configs[2]): 1.48 M masked "
                "pixels (query_disc 1.5 rad + extend_indices), Chebyshev / Monomial / residual K 10 F 5 with BatchNorm, "
                "n_neighbors 20 (ELL + CSR tail path), batch 8/GPU")
        loss_fn = lambda y, t: -(torch.log(y.clamp_min(1e-12)) * t).sum(dim=1).mean()
        n_out = 2
    elif name == "C4_autoencoder":
        nside, B, k = 128, 5, 20
        idx = np.arange(12 * nside**2)
        enc_l, dec_l = nets.autoencoder_layers()
        enc = deepsphere.HealpyGCNN(nside=nside, indices=idx, layers=enc_l, n_neighbors=k)
        dec = deepsphere.HealpyGCNN(nside=nside // 8, indices=np.arange(12 * (nside // 8) ** 2), layers=dec_l, n_neighbors=k)
        model = None
        desc = ("examples/generative_models.ipynb:185-213,456 scaled to nside 128 (BASELINE.json configs[3]): pseudo-conv "
                "encoder, Chebyshev K 5 F 16 + LayerNormalization(axis=1) + elu, transposed pseudo-conv decoder, "
                "n_neighbors 20, MAE loss, batch 5/GPU")
        loss_fn = None
        n_out = None
    else:
        raise ValueError(name)
    M = len(idx)
    gen = torch.Generator(device=device).manual_seed(11 + rank)
    x = torch.randn(B, M, 1, device=device, generator=gen)
    if name == "C4_autoencoder":
        enc.build(input_shape=(None, M, 1))
        dec.build(input_shape=(None, M // 64, 16))
        modules = [enc, dec]
        fwd = lambda v, training: dec(enc(v, training=training), training=training)
        step_loss = lambda: (fwd(x, True) - x).abs().mean()
        layers_all = list(enc.layers) + list(dec.layers)
    else:
        model.build(input_shape=(None, M, 1))
        modules = [model]
        labels = torch.nn.functional.one_hot(torch.randint(0, n_out, (B,), device=device, generator=gen), n_out).float()
        fwd = lambda v, training: model(v, training=training)
        step_loss = lambda: loss_fn(fwd(x, True), labels)
        layers_all = list(model.layers)
    for m in modules:
        m.to(device)
        dsd.broadcast_parameters(m)
    params = [p for m in modules for p in m.trainable_variables]
    opt = _adam(params, args)

    def train_step():
        opt.zero_grad(set_to_none=True)
        loss = step_loss()
        loss.backward()
        dsd.allreduce_gradients(params)
        opt.step()
        return loss

    for _ in range(3):
        train_step()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
        torch.cuda.synchronize()
    n = max(3, min(args.steps, 10))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        loss = train_step()
    b.record()
    torch.cuda.synchronize()
    ms = dsd.allreduce_max(a.elapsed_time(b) / n, device)
    nbytes = sum(_step_bytes(m, xi) for m, xi in zip(modules, [x] if len(modules) == 1 else
                                                      [x, torch.zeros(B, M // 64, 16, device=device)]))
    final_loss = float(loss.detach())
    del loss
    eager_ms = ms
    graph_t = None if args.no_graph else graph_replay_timing(train_step, params, opt, n, device, world)
    if graph_t is not None and graph_t.get("validated"):
        ms = graph_t["ms_per_step"]
    out = {"metric": "HealpyGCNN train maps/s", "value": world * B / (ms * 1e-3), "unit": "maps/s", "ms_per_step": ms,
           "batch_per_gpu": B, "n_gpus": world, "parameters": int(sum(p.numel() for p in params)),
           "final_loss": final_loss, "config": desc,
           "execution": "CUDA graph replay of the whole step (validated against eager)" if ms != eager_ms else "eager",
           "eager_ms_per_step": eager_ms, "cuda_graph": graph_t,
           "roofline": {"bound": "hbm", "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": nbytes / (ms * 1e-3) / 1e9 / hbm_peak,
                        "note": "algorithmic bytes of one step (every layer: forward in + out, backward 2 in + out) / step "
                                "time; these 1 - 16 channel networks are launch- and latency-bound, not HBM-bound"}}
    # CPU restatement on the host cores (rank 0, N = 1 only): one sample, forward + backward; parity of the forward
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        specs, _ = bridge.specs_from_layers(layers_all)
        def f32(v):   # float32 weights: the reference's floatx arithmetic
            if isinstance(v, torch.Tensor):
                return v.detach().float().requires_grad_(v.requires_grad)
            if isinstance(v, (list, tuple)):
                return type(v)(f32(u) for u in v)
            return v

        for kind, p in specs:
            for key in list(p.keys()):
                p[key] = f32(p[key])
        xs = x[:1].detach().cpu()
        with torch.no_grad():
            y_gpu = fwd(x[:1], False).detach().cpu()
        t0 = time.perf_counter()
        y_cpu = orc.torch_cpu_network(xs.clone().requires_grad_(True), specs, training=False)
        y_cpu.sum().backward()
        t_cpu = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": 1.0 / t_cpu, "unit": "maps/s", "cores": torch.get_num_threads(), "kind": "port",
                               "sample": "1 map, forward + backward, torch-CPU restatement of the same layer list "
                                         "(oracle.torch_cpu_network), inference-mode BatchNorm", "ms_per_step": t_cpu * 1e3}
        err = float((y_gpu - y_cpu.detach()).abs().max() / y_cpu.detach().abs().max().clamp_min(1e-30))
        out["parity"] = {"rel_err": {"y": err}, "tolerance": 1e-4, "ok": bool(err <= 1e-4),
                         "what": "GPU forward of one map vs the CPU restatement (fp32), max|err| / max|ref|"}
    for m in modules:
        del m
    del opt, x
    torch.cuda.empty_cache()
    return out


def partition_parity_check(mode, device, rank, world, nside=64, batch=2):
    """N ranks == 1 rank, on hardware and visible to the driver: the C5-pattern network on ONE nside-64
    sphere partitioned over the ranks of this run, against the whole-sphere HealpyGCNN on rank 0 with the same weights
    and input: output and every weight gradient (max|err| / max|ref|)."""
    import deepsphere
    from deepsphere import distributed as dsd, keras_compat as kc, partition

    npix = 12 * nside * nside
    torch.manual_seed(5)
    part = partition.PartitionedHealpyGCNN(nside, np.arange(npix), _c5_layers(mode, partition.PartitionedMean()),
                                           rank=rank, world=world)
    gen = torch.Generator(device=device).manual_seed(7)
    x = torch.randn(batch, npix, 1, device=device, generator=gen)
    t = torch.randn(batch, 2, device=device, generator=gen)
    b0, e0 = part.own_range
    y = part(x[:, b0:e0].contiguous(), training=True)   # builds the lazily-shaped weights
    dsd.broadcast_parameters(part)
    params = [p for p in part.parameters() if p.requires_grad]
    for p in params:
        p.grad = None
    y = part(x[:, b0:e0].contiguous(), training=True)
    ((y - t) ** 2).mean().backward()
    part.allreduce_gradients()   # sums the row-local partial sums; the head behind the mean is replicated
    torch.cuda.synchronize()
    out = None
    if rank == 0:
        torch.manual_seed(5)
        whole = deepsphere.HealpyGCNN(nside=nside, indices=np.arange(npix),
                                      layers=_c5_layers(mode, kc.Lambda(lambda v: v.mean(dim=1))))
        whole.build(input_shape=(None, npix, 1))
        wparams = whole.trainable_variables
        assert len(wparams) == len(params), (len(wparams), len(params))
        with torch.no_grad():
            for a, b in zip(wparams, params):
                a.copy_(b)
        yw = whole(x, training=True)
        ((yw - t) ** 2).mean().backward()

        def rel(a, b):
            return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))

        errs = {"y": rel(y.detach(), yw.detach()), "grad_max": max(rel(b.grad, a.grad) for a, b in zip(wparams, params))}
        tol = 2e-3 if mode == "tf32" else 1e-4
        out = {"rel_err": errs, "tolerance": tol, "ok": bool(max(errs.values()) <= tol), "nside": nside, "batch": batch,
               "ranks": world, "what": "sphere-partitioned C5-pattern network vs the whole-sphere HealpyGCNN on rank 0: "
                                       "output and all weight gradients"}
    return out


def model_train_partitioned_bench(args, mode, device, rank, world):
    """North-star scaling configuration (BASELINE.json configs[4]): the full-sphere regression
    HealpyGCNN at nside 1024 (12.6 M pixels), ONE sphere per sample partitioned over the N ranks by quarter-face blocks
    with a (K-1)-ring halo exchange per graph layer — strong scaling: the same global batch at every N, N = 1 is the
    whole sphere on one GPU.  Step = forward + backward + weight-gradient all-reduce (sum of the ranks' partial sums) +
    Adam.  Also reports the device time inside the halo exchanges and the all-reduce, and the halo volume."""
    from deepsphere import distributed as dsd, partition

    nside, Bm = args.part_nside, args.part_batch
    npix = 12 * nside * nside
    # the whole batch must fit ONE GPU (N = 1 is the whole sphere on one GPU): measured peak 4.9 GB per nside-1024 sample
    # (156 GB at batch 32: activations kept for the backward + the top layer's basis workspace); the same decision on
    # every rank and at every N, so that the scaling curve is one global batch throughout
    per_sample = 5.0e9 * (nside / 1024.0) ** 2
    cap = 0.8 * torch.cuda.get_device_properties(device).total_memory
    while Bm > 1 and Bm * per_sample > cap:
        Bm //= 2
    torch.cuda.reset_peak_memory_stats(device)
    t0 = time.time()
    torch.manual_seed(11)
    model = partition.PartitionedHealpyGCNN(nside, np.arange(npix), _c5_layers(mode, partition.PartitionedMean()),
                                            rank=rank, world=world)
    b0, e0 = model.own_range
    gen = torch.Generator(device=device).manual_seed(11 + rank)
    x = torch.randn(Bm, e0 - b0, 1, device=device, generator=gen)
    gt = torch.Generator(device=device).manual_seed(3)
    t = torch.randn(Bm, 2, device=device, generator=gt)
    model(x, training=True)  # builds the weights and the device plans
    dsd.broadcast_parameters(model)
    params = [p for p in model.parameters() if p.requires_grad]
    opt = _adam(params, args)
    prep_s = time.time() - t0
    ar_events = []

    def train_step(timed=False):
        opt.zero_grad(set_to_none=True)
        loss = ((model(x, training=True) - t) ** 2).mean()
        loss.backward()
        if timed:
            a = torch.cuda.Event(enable_timing=True); a.record()
        model.allreduce_gradients()
        if timed:
            b = torch.cuda.Event(enable_timing=True); b.record()
            ar_events.append((a, b))
        opt.step()
        return loss

    for _ in range(3):
        train_step()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
        torch.cuda.synchronize()
    n = max(3, min(args.steps, 10))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        loss = train_step()
    b.record()
    torch.cuda.synchronize()
    ms = dsd.allreduce_max(a.elapsed_time(b) / n, device)
    # time split (separate, instrumented steps: the event pairs serialise nothing but add host work)
    partition._Timing.start()
    for _ in range(2):
        train_step(timed=True)
    torch.cuda.synchronize()
    ex_ms, ex_n = partition._Timing.stop()
    ar_ms = sum(u.elapsed_time(v) for u, v in ar_events) / 2
    ex_ms = dsd.allreduce_max(ex_ms / 2, device)
    ar_ms = dsd.allreduce_max(ar_ms, device)
    del loss
    # a captured step allocates from its own memory pool, next to the eager pool's working set: only when both fit
    fits_twice = 2.2 * torch.cuda.max_memory_allocated(device) < cap
    graph_t = graph_replay_timing(train_step, params, opt, n, device, world) if (fits_twice and not args.no_graph) else \
        {"skipped": "the step's working set does not fit twice (eager pool + graph pool)", "validated": False}
    eager_ms = ms
    if graph_t is not None and graph_t.get("validated"):
        ms = graph_t["ms_per_step"]
    convs = [l for l in model.layers_use if isinstance(l, partition.PartitionedGraphConv)]
    halo = [{"level_rows_own": int(c.plan.n_own), "halo_rows": int(c.plan.halo_rows), "hops": int(c.plan.n_hops),
             "channels_in": int(c.layer.kernel.shape[0] // c.layer._n_terms),
             "lattice": int(c.layer._plan.info(device.index or 0)["lattice"])} for c in convs]
    # bytes this rank RECEIVES per step: every layer's halo rows once in the forward (inputs) and once in the backward
    # (input gradients travel the other way, same volume)
    halo_bytes = int(sum(2 * h["halo_rows"] * Bm * h["channels_in"] * 4 for h in halo))
    n_params = int(sum(p.numel() for p in params))
    out = {"metric": "HealpyGCNN train maps/s, nside %d, one sphere partitioned over the ranks" % nside,
           "value": Bm / (ms * 1e-3), "unit": "maps/s", "ms_per_step": ms, "global_batch": Bm, "n_gpus": world,
           "scaling": "strong", "parameters": n_params, "peak_memory_GB": torch.cuda.max_memory_allocated(device) / 1e9,
           "execution": "CUDA graph replay of the whole step (validated against eager)" if ms != eager_ms else "eager",
           "eager_ms_per_step": eager_ms, "cuda_graph": graph_t,
           "time_split_ms": {"halo_exchanges": ex_ms, "halo_exchanges_per_step": ex_n // 2, "grad_allreduce": ar_ms,
                             "rest (kernels, optimizer, host launch gaps)": eager_ms - ex_ms - ar_ms,
                             "note": "device time between event pairs in two instrumented EAGER steps"},
           "limiting_collective": None if world == 1 else ("halo all_to_all" if ex_ms >= ar_ms else "gradient all-reduce"),
           "halo": halo, "halo_bytes_per_step_per_rank": halo_bytes, "host_prep_s": prep_s,
           "config": f"nside {nside} ({npix} px): PseudoConv p1 F16 -> [Chebyshev K5 F32 + MAX pool] x4 -> Chebyshev K5 F64 "
                     f"-> AVG pool -> mean -> Dense(2); MSE, Adam; mode {mode}; rank r owns 48/{world} quarter-face blocks"}
    del model, opt, x
    torch.cuda.empty_cache()
    return out


def experimental_model_run(args, baseline, switch_env=None, extra_args=()):
    """`bench.py --model-only` in a child process (bounded: 4 minutes) with an opt-in switch: `switch_env` (e.g.
    DEEPSPHERE_SKINNY=1) and / or extra flags (--model-graph).  The child's number only counts as validated if its
    first training step reproduces the default path's loss and per-parameter gradient norms."""
    import subprocess

    switch_env = dict(switch_env or {})
    label = " ".join([f"{k}={v}" for k, v in switch_env.items()] + list(extra_args))
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "0"), **switch_env)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    cmd = [sys.executable, os.path.abspath(__file__), "--model-only", "--mode", args.mode,
           "--model-nside", str(args.model_nside), "--model-batch", str(args.model_batch), *extra_args]
    try:
        res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
        lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
        if res.returncode != 0 or not lines:
            return {"switch": label, "error": (res.stderr or res.stdout)[-300:]}
        out = json.loads(lines[-1])
        out["switch"] = label
        a, b = baseline["first_step"], out["first_step"]
        rel = [abs(a["loss"] - b["loss"]) / max(abs(a["loss"]), 1e-12)]
        rel += [abs(u - v) / max(abs(u), 1e-12) for u, v in zip(a["grad_norms"], b["grad_norms"])]
        out["first_step_max_rel_diff"] = max(rel)
        # same forward, same gradients (fp32 summation order differs in the fused weight-gradient sweep); a graph
        # replay must also land on the same loss after the same number of steps
        ok = max(rel) <= 1e-4 and len(a["grad_norms"]) == len(b["grad_norms"])
        if "--model-graph" in extra_args:
            ok = ok and abs(out["final_loss"] - baseline["final_loss"]) <= 1e-3 * max(abs(baseline["final_loss"]), 1e-12)
        out["validated"] = bool(ok)
        out["speedup_vs_default"] = baseline["ms_per_step"] / out["ms_per_step"]
        return out
    except Exception as exc:
        return {"switch": label, "error": str(exc)[:300]}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        from deepsphere.graph import SphereHealpix

        # torchrun exports OMP_NUM_THREADS=1 for its workers; the reference arm is rank 0 alone on the box's host
        # cores and may use all of them
        try:
            torch.set_num_threads(max(torch.get_num_threads(), len(os.sched_getaffinity(0))))
        except (AttributeError, RuntimeError):
            pass
        g = SphereHealpix(args.nside, k=8)
        steps, warmup = max(1, min(args.steps, 3)), max(0, min(args.warmup, 1))
        t, nbytes, _ = cpu_reference_time(g.L, args, args.cpu_sample_batch, steps, warmup)
        val = nbytes / t / 1e9
        sample = (f"batch {args.cpu_sample_batch} of the workload's {args.batch} (same nside/K/F), fwd+bwd, "
                  f"{steps} timed steps after {warmup} warm-up")
        line = {
            "impl": "reference", "metric": METRIC, "value": val, "unit": "GB/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": f"HealpyChebyshev nside {args.nside} K {args.K} Fin=Fout={args.features} fwd+bwd",
                       "parallelism": "host CPU threads"},
            "cpu_baseline": {"value": val, "unit": "GB/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line))
        return

    from deepsphere import _native as nat
    from deepsphere import distributed as dsd

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    dsd.init_from_env()
    mode = args.mode
    if args.model_only:
        out = model_train_bench(args, mode, device, world)
        if rank == 0:
            print(json.dumps(out))
        return
    g, layer = build_layer(args, mode)
    M, F, B, K = g.L.shape[0], args.features, args.batch, args.K
    layer.build_from_shape((B, M, F))
    dsd.broadcast_parameters(layer)
    gen = torch.Generator(device=device).manual_seed(1234 + rank)
    x = torch.randn(B, M, F, device=device, generator=gen).requires_grad_(True)
    dy = torch.randn(B, M, F, device=device, generator=gen)
    nbytes = algorithmic_bytes(B, M, F, F)

    def step():
        x.grad = None
        layer.kernel.grad = None
        y = layer(x)
        y.backward(dy)
        dsd.allreduce_gradients([layer.kernel])
        return y

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sync_all()
    launches0 = nat.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        sync_all()
    ms = e0.elapsed_time(e1) / args.steps
    launches = (nat.launch_count() - launches0)

    # ---- parity sample at THIS configuration: the first cpu_sample_batch maps of the workload through the same layer
    # (forward, dx, dkernel); compared further down with the CPU restatement of the reference on the same inputs
    parity_sample = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        nb = max(1, min(args.cpu_sample_batch, B))
        xs = x.detach()[:nb].clone().requires_grad_(True)
        layer.kernel.grad = None
        ys = layer(xs)
        ys.backward(dy[:nb])
        torch.cuda.synchronize()
        parity_sample = {"x": xs.detach().cpu(), "dy": dy[:nb].cpu(), "kernel": layer.kernel.detach().cpu().clone(),
                         "y": ys.detach().cpu(), "dx": xs.grad.cpu(), "dkernel": layer.kernel.grad.detach().cpu().clone()}
        del xs, ys
        layer.kernel.grad = None
    ms = dsd.allreduce_max(ms, device)
    value = world * nbytes / (ms * 1e-3) / 1e9
    hbm_peak, bf16_peak, peak_kind = peaks()

    # ---- per-kernel timing of the path's kernels, in isolation, on the launching stream ----------
    from deepsphere import _ops

    def time_fn(fn, n=5):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    A_bytes = 4 * B * M * F
    fused = layer._plan.info(local_rank)["lattice"] == 1
    fused_conv = fused and mode == "tf32" and F % 8 == 0 and F <= 80 and K <= 5
    with torch.no_grad():
        xd = x.detach()
        fwd_ms = time_fn(lambda: layer(xd))                              # recursion + contraction
        t1 = _ops.spmm(layer._plan, xd)
        hop_ms = time_fn(lambda: _ops.spmm(layer._plan, t1, 2.0, xd, -1.0))  # one generic streaming hop
        basis_ms = None if fused_conv else time_fn(lambda: _ops.basis(layer._plan, xd, K))
        del t1
    gemm_flops = 2.0 * B * M * K * F * F
    nnz_hops = 9.0 * (K - 1)
    fma_flops = 2.0 * nnz_hops * B * M * F            # useful stencil FLOPs of one recursion pass
    kernels = {
        "generic_hop": {"kernel": "spmm_tile_kernel", "ms": hop_ms, "algorithmic_bytes": 3 * A_bytes,
                        "GBps": 3 * A_bytes / hop_ms / 1e6},
        "forward": {"ms": fwd_ms, "algorithmic_bytes": 2 * A_bytes, "GBps": 2 * A_bytes / fwd_ms / 1e6},
        "backward": {"ms": ms - fwd_ms, "algorithmic_bytes": 3 * A_bytes, "GBps": 3 * A_bytes / max(ms - fwd_ms, 1e-6) / 1e6,
                     "note": "step minus forward: dz recursion + dx contraction (+ basis U_k to HBM) and the dW kernel"},
    }
    if fused_conv:
        # one launch = recursion (fp32 FFMA2, register resident) + tcgen05 contraction + bias/activation epilogue;
        # HBM sees x (+ halo, from L2) and y only: algorithmic bytes per launch = 2 A
        dom = {"kernel": "lattice_conv2_kernel<Chebyshev> (all %d hops + contraction fused)" % (K - 1), "ms": fwd_ms,
               "algorithmic_bytes": 2 * A_bytes, "GBps": 2 * A_bytes / fwd_ms / 1e6,
               "stencil_TFLOPs_fp32": fma_flops / fwd_ms / 1e9, "contraction_TFLOPs_tf32": gemm_flops / fwd_ms / 1e9}
        kernels["fused_forward"] = dom
        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel, one `ncu --set full` capture (see TRAFFIC_DEFAULT);
        # only quoted for the default config
        traffic = TRAFFIC_DEFAULT if (args.nside, B, F, K) == (256, 32, 64, 5) else None
    else:
        contraction_ms = max(fwd_ms - basis_ms, 1e-6)
        rec_name = ("lattice_recursion_kernel<%d,16,8> (all %d hops fused)" % (max(K - 1, 4), K - 1)) if fused \
            else "spmm_tile_kernel x %d" % (K - 1)
        kernels["recursion"] = {"kernel": rec_name, "ms": basis_ms, "algorithmic_bytes": K * A_bytes,
                                "GBps": K * A_bytes / basis_ms / 1e6}
        kernels["contraction"] = {"kernel": "umma_gemm_kernel" if mode != "fp32" else "gemm_nn_kernel",
                                  "ms": contraction_ms, "algorithmic_bytes": (K + 1) * A_bytes,
                                  "GBps": (K + 1) * A_bytes / contraction_ms / 1e6,
                                  "TFLOPs": gemm_flops / contraction_ms / 1e9}
        dom = kernels["recursion"] if 2 * basis_ms >= 2.5 * contraction_ms else kernels["contraction"]
        traffic = None
    roofline = {"kernel": dom["kernel"], "bound": "hbm", "achieved": dom["GBps"], "peak": hbm_peak, "unit": "GB/s",
                "frac": dom["GBps"] / hbm_peak, "traffic": traffic, "peak_source": peak_kind,
                "note": "algorithmic bytes of one launch / its CUDA-event duration; DESIGN.md section 3 explains why the "
                        "fused kernel is bound on chip (shared-memory exchange + fp32 FMA issue), not by HBM"}
    layer_roofline = {"bound": "hbm", "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                      "frac": nbytes / (ms * 1e-3) / 1e9 / hbm_peak,
                      "note": "whole layer fwd+bwd, algorithmic bytes / step time, per GPU"}

    # ---- the same step in the other contraction modes (short runs, reported next to the headline) ------
    other_modes = {}
    if not args.no_other_modes and rank == 0 and world == 1:
        for om in ("tf32x3", "tf32", "fp32"):
            if om == mode:
                continue
            try:
                layer.mode = om
                step()
                torch.cuda.synchronize()
                a0, b0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                for _ in range(2):
                    step()
                b0.record()
                torch.cuda.synchronize()
                oms = a0.elapsed_time(b0) / 2
                other_modes[om] = {"ms_per_step": oms, "GBps": nbytes / (oms * 1e-3) / 1e9}
            except Exception as exc:  # e.g. out of memory for the unfused fp32 workspace
                other_modes[om] = {"error": str(exc)[:120]}
            finally:
                layer.mode = mode
                torch.cuda.empty_cache()

    # ---- the fused forward at narrower layers (the arithmetic intensity falls with F) -------------------
    f_sweep = None
    if not args.no_f_sweep and rank == 0 and world == 1 and fused_conv:
        from deepsphere import gnn_layers

        f_sweep = {}
        for Fs in (16, 32):
            try:
                ls = gnn_layers.Chebyshev(L=g.L, K=K, Fout=Fs, mode=mode, lmax=layer.lmax)
                xs = torch.randn(B, M, Fs, device=device)
                with torch.no_grad():
                    t_ms = time_fn(lambda: ls(xs))
                f_sweep[f"F{Fs}"] = {"ms": t_ms, "GBps": 2 * 4 * B * M * Fs / t_ms / 1e6,
                                     "frac_of_hbm_peak": 2 * 4 * B * M * Fs / t_ms / 1e6 / hbm_peak}
                del ls, xs
            except Exception as exc:
                f_sweep[f"F{Fs}"] = {"error": str(exc)[:120]}
            torch.cuda.empty_cache()

    # ---- HealpyGCNN training throughput (the second half of the metric) ------------------------------
    model_train = None
    if not args.no_model:
        try:
            model_train = model_train_bench(args, mode, device, world)
        except Exception as exc:
            model_train = {"error": str(exc)[:200]}

    # ---- the north-star scaling configuration: nside 1024, one sphere partitioned over the ranks (strong scaling), and
    # the N-rank == 1-rank parity of that path on a small sphere
    model_train_partitioned = partition_parity = None
    if not args.no_partitioned and not args.no_model:
        try:
            if world > 1:
                partition_parity = partition_parity_check(mode, device, rank, world)
            model_train_partitioned = model_train_partitioned_bench(args, mode, device, rank, world)
        except Exception as exc:
            import traceback

            model_train_partitioned = {"error": str(exc)[:300], "trace": traceback.format_exc()[-600:]}

    # ---- BASELINE.json's other named configurations, each with its CPU restatement beside it
    named_configs = None
    if not args.no_configs and not args.no_model:
        named_configs = {}
        for cname in ("C1_quick_start", "C4_autoencoder", "C3_masked_survey"):
            try:
                named_configs[cname] = named_config_bench(cname, args, device, rank, world, hbm_peak)
            except Exception as exc:
                import traceback

                named_configs[cname] = {"error": str(exc)[:300], "trace": traceback.format_exc()[-500:]}
            torch.cuda.empty_cache()

    # ---- the same training step with the opt-in streaming pseudo-convolution kernels (ds_skinny.cu, written after
    # this round's GPU budget was spent: host-emulated only, hence not the default).  Separate process so that a fault
    # there cannot touch this measurement; "validated" = its first-step loss and per-parameter gradient norms (same
    # seeds, before any update) equal the default path's to 1e-4.
    model_train_experimental = None
    if model_train is not None and "error" not in model_train and args.experimental and not args.no_experimental \
            and world == 1:
        model_train_experimental = {}
        # (round 2: the streaming pseudo-convolution kernels became the default; the child runs the OTHER setting)
        for key, env_sw, flags in (("tiled_pseudo_conv_kernels", {"DEEPSPHERE_SKINNY": "0"}, ("--no-graph",)),):
            model_train_experimental[key] = experimental_model_run(args, model_train, env_sw, flags)
            if "timed out" in str(model_train_experimental[key].get("error", "")):
                break  # do not spend more of the bench's minutes on a child that hangs

    # ---- e2e: public layer API, pinned host buffers, H2D + D2H inside the timed region ------------
    e2e = None
    if not args.no_e2e:
        try:
            import psutil

            free_host = psutil.virtual_memory().available
        except Exception:
            free_host = 0
        # every rank of the node pins its own buffers at the same time: budget 40 % of the host memory over
        # the local ranks (8 ranks x 19.3 GB would not fit a 196 GB box)
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
        Be = B
        while Be > 1 and 3 * 4 * Be * M * F > 0.4 * free_host / max(local_world, 1):
            Be //= 2
        xh = torch.empty((Be, M, F), dtype=torch.float32).pin_memory()
        dyh = torch.empty((Be, M, F), dtype=torch.float32).pin_memory()
        dxh = torch.empty((Be, M, F), dtype=torch.float32).pin_memory()
        dkh = torch.empty((K * F, F), dtype=torch.float32).pin_memory()
        xh.normal_()
        dyh.normal_()
        del x, dy
        torch.cuda.empty_cache()

        # The reference-facing call with HOST buffers: batch chunks are pipelined over three streams (H2D of chunk
        # i+1, layer fwd+bwd of chunk i, D2H of chunk i-1), all through the public layer API; the kernel gradient
        # accumulates over the chunks exactly like a gradient-accumulation step.
        n_chunks = max(1, min(8, Be))
        while Be % n_chunks:
            n_chunks -= 1
        Bc = Be // n_chunks
        s_in, s_cmp, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
        x_d = [torch.empty((Bc, M, F), device=device) for _ in range(2)]
        dy_d = [torch.empty((Bc, M, F), device=device) for _ in range(2)]

        def e2e_step():
            layer.kernel.grad = None
            cur = torch.cuda.current_stream()
            for st_ in (s_in, s_cmp, s_out):
                st_.wait_stream(cur)
            ev_in, ev_cmp, ev_out = [], [], []
            for i in range(n_chunks):
                sl = slice(i * Bc, (i + 1) * Bc)
                with torch.cuda.stream(s_in):
                    if i >= 2:
                        s_in.wait_event(ev_cmp[i - 2])  # the device buffers of chunk i-2 are free again
                    x_d[i & 1].copy_(xh[sl], non_blocking=True)
                    dy_d[i & 1].copy_(dyh[sl], non_blocking=True)
                    ev_in.append(torch.cuda.Event()); ev_in[-1].record(s_in)
                with torch.cuda.stream(s_cmp):
                    s_cmp.wait_event(ev_in[i])
                    xg = x_d[i & 1].detach().requires_grad_(True)
                    y = layer(xg)
                    y.backward(dy_d[i & 1])
                    gx = xg.grad
                    ev_cmp.append(torch.cuda.Event()); ev_cmp[-1].record(s_cmp)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_cmp[i])
                    gx.record_stream(s_out)
                    dxh[sl].copy_(gx, non_blocking=True)
            with torch.cuda.stream(s_cmp):
                dsd.allreduce_gradients([layer.kernel])
                dkh.copy_(layer.kernel.grad, non_blocking=True)
            for st_ in (s_in, s_cmp, s_out):
                cur.wait_stream(st_)

        e2e_step()
        sync_all()
        n_e2e = max(2, min(args.steps, 4))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n_e2e):
            e2e_step()
        b.record()
        sync_all()
        e2e_ms = dsd.allreduce_max(a.elapsed_time(b) / n_e2e, device)
        e2e = {"value": world * algorithmic_bytes(Be, M, F, F) / (e2e_ms * 1e-3) / 1e9, "unit": "GB/s",
               "h2d_bytes_per_step": 2 * 4 * Be * M * F, "d2h_bytes_per_step": 4 * Be * M * F + 4 * K * F * F,
               "ms_per_step": e2e_ms, "batch": Be, "steps": n_e2e,
               "pipeline": f"{n_chunks} batch chunks over 3 streams (H2D | fwd+bwd | D2H), pinned host buffers",
               "host_numa_binding": None if numa_cpus is None else f"{len(numa_cpus)} CPUs local to GPU {local_rank} (NVML)"}

    # ---- CPU baseline (rank 0, N = 1 only): the oracle port on a bounded sample -------------------
    cpu_baseline = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from scipy import sparse

        Lt = sparse.csr_matrix((layer._L_values.astype(np.float64), (layer._L_indices[:, 0], layer._L_indices[:, 1])),
                               shape=g.L.shape)  # the layer's own prepared L~ (skips a second ARPACK run)
        t, nb, par = cpu_reference_time(g.L, args, args.cpu_sample_batch, 2, 1, Lt=Lt, sample=parity_sample)
        cpu_baseline = {"value": nb / t / 1e9, "unit": "GB/s", "cores": torch.get_num_threads(), "kind": "port",
                        "sample": f"batch {args.cpu_sample_batch} of {B} (same nside/K/F), fwd+bwd, torch-CPU "
                                  f"restatement of gnn_layers.py:131-150, 2 timed steps", "ms_per_step": t * 1e3}
        if par is not None:
            tol = PARITY_TOL[mode]
            parity = {"rel_err": par, "tolerance": tol, "ok": bool(max(par.values()) <= tol),
                      "what": f"GPU y / dx / dkernel of the first {args.cpu_sample_batch} map(s) of this workload vs the "
                              f"CPU restatement on the same inputs (max|err| / max|ref|), mode {mode}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"fp32": "fp32", "tf32": "fp32 recursion (FFMA2) + TF32 tensor-core contraction, fp32 accumulate (rel err <= 1e-3, measured 5e-4)",
                                           "tf32x3": "fp32 (recursion fp32 FMA; contraction 3xTF32 error-compensated on tcgen05, rel err <= 2e-5)"}[mode],
            "data": "synthetic",
            "config": {"workload": f"HealpyChebyshev layer nside {args.nside} (M={M}) K {K} Fin=Fout={F} "
                                   f"batch {B}/GPU fwd+bwd, 8-neighbour HEALPix graph",
                       "mode": mode, "parallelism": f"batch-sharded x{world}", "l2": "inputs (6.4 GB/tensor) >> L2"},
            "roofline": roofline, "layer_roofline": layer_roofline, "kernels": kernels, "fused_forward_narrow_layers": f_sweep,
            "other_modes": other_modes, "model_train": model_train,
            "model_train_experimental": model_train_experimental,
            "model_train_partitioned": model_train_partitioned, "partition_parity": partition_parity,
            "named_configs": named_configs,
            "cpu_baseline": cpu_baseline, "parity": parity, "e2e": e2e, "gpu_launches": launches, "clocks": clocks.summary(),
        }
        print(json.dumps(line))
        if parity is not None and not parity["ok"]:
            sys.stderr.write(f"bench.py: PARITY FAILED at the bench configuration: {parity}\n")
            sys.exit(3)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
