"""Driver for ncu launch lists of the sphere-partitioned nside-1024 training step of bench.py (model_train_partitioned) at
N = 1 (the whole sphere on one GPU): BATCH (default 4) keeps the kernel replay cheap."""
import argparse, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deepsphere-cosmo-tf2_b200"))
import bench
a = argparse.Namespace(part_nside=int(os.environ.get("NSIDE", "1024")), part_batch=int(os.environ.get("BATCH", "4")), steps=3,
                       no_graph=True)
torch.cuda.set_device(0)
out = bench.model_train_partitioned_bench(a, os.environ.get("MODE", "tf32"), torch.device("cuda", 0), 0, 1)
print({k: out[k] for k in ("ms_per_step", "global_batch", "time_split_ms", "halo_bytes_per_step_per_rank")})
