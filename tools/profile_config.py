"""Driver for ncu launch lists of a named configuration of bench.py (C1_quick_start / C3_masked_survey / C4_autoencoder)."""
import argparse, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deepsphere-cosmo-tf2_b200"))
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "C3_masked_survey"
a = argparse.Namespace(steps=3, no_graph=True, no_cpu_baseline=True)
torch.cuda.set_device(0)
out = bench.named_config_bench(name, a, torch.device("cuda", 0), 0, 1, 6551.0)
print({k: out[k] for k in ("ms_per_step", "value")})
