"""Per-layer forward / backward device time of the sphere-partitioned C5 network (nside 1024), under torchrun:
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/part_layer_times.py [batch]
Prints rank 0's table: which layers scale with 1/N and which do not."""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deepsphere-cosmo-tf2_b200")); sys.path.insert(0, ROOT)
from deepsphere import distributed as dsd, partition
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nside = int(os.environ.get("NSIDE", "1024"))
rank, world, local = dsd.init_from_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
npix = 12 * nside * nside
torch.manual_seed(11)
model = partition.PartitionedHealpyGCNN(nside, np.arange(npix), bench._c5_layers("tf32", partition.PartitionedMean()), rank=rank, world=world)
b0, e0 = model.own_range
x = torch.randn(B, e0 - b0, 1, device=dev)
model(x, training=True)
dsd.broadcast_parameters(model)
names = [type(l.layer).__name__ + "(part)" if isinstance(l, partition.PartitionedGraphConv) else type(l).__name__ for l in model.layers_use]
from deepsphere.keras_compat import _accepts_training
def run(n_rep=3):
    fw = np.zeros(len(names)); bw = np.zeros(len(names))
    for rep in range(n_rep + 1):
        hs, evs = [x.clone().requires_grad_(True)], []
        for layer in model.layers_use:
            inner = layer.layer if isinstance(layer, partition.PartitionedGraphConv) else layer
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record()
            h = layer(hs[-1], training=True) if _accepts_training(inner) else layer(hs[-1])
            b.record(); evs.append((a, b))
            hs.append(h)
        # backward layer by layer: detach chain
        loss = hs[-1].pow(2).mean()
        torch.cuda.synchronize()
        if rep: fw += [a.elapsed_time(b) for a, b in evs]
    return fw / n_rep
# per-layer backward: rebuild with detached inputs
def run_bwd(n_rep=3):
    bw = np.zeros(len(names))
    for rep in range(n_rep + 1):
        inp = x.clone().requires_grad_(True)
        outs = []
        h = inp
        ins = []
        for layer in model.layers_use:
            inner = layer.layer if isinstance(layer, partition.PartitionedGraphConv) else layer
            hin = h.detach().requires_grad_(True)
            ins.append(hin)
            h = layer(hin, training=True) if _accepts_training(inner) else layer(hin)
            outs.append(h)
        g = torch.ones_like(outs[-1]) / outs[-1].numel()
        for i in reversed(range(len(names))):
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record()
            outs[i].backward(g)
            b.record()
            torch.cuda.synchronize()
            if rep: bw[i] += a.elapsed_time(b)
            g = ins[i].grad
            if g is None: break
    return bw / n_rep
fw = run(); bw = run_bwd()
if rank == 0:
    print(json.dumps({"n_gpus": world, "batch": B, "layers": [{"layer": n, "fwd_ms": round(float(f), 3), "bwd_ms": round(float(b), 3)} for n, f, b in zip(names, fw, bw)],
                      "fwd_total": float(fw.sum()), "bwd_total": float(bw.sum())}))
if world > 1:
    torch.distributed.destroy_process_group()
