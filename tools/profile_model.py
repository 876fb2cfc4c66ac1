"""Driver for ncu launch lists of the HealpyGCNN training step used by bench.py (model_train)."""
import os, sys, argparse
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deepsphere-cosmo-tf2_b200"))
import bench
a = argparse.Namespace(model_nside=int(os.environ.get("NSIDE", "256")), model_batch=int(os.environ.get("BATCH", "16")))
torch.cuda.set_device(0)
print(bench.model_train_bench(a, os.environ.get("MODE", "tf32"), torch.device("cuda", 0), 1))
