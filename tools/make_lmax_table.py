"""Writes deepsphere/_lmax_table.json: largest eigenvalue (utils.largest_eigenvalue, cache bypassed) of the normalised
Laplacians of graph.SphereHealpix on the full sphere, keyed by utils.matrix_fingerprint of the matrix content.
  python tools/make_lmax_table.py [max_nside_k8] [max_nside_k20]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deepsphere-cosmo-tf2_b200"))
os.environ["DEEPSPHERE_LMAX"] = "nocache"
from scipy import sparse
import numpy as np
from deepsphere import utils
from deepsphere.graph import SphereHealpix

max8 = int(sys.argv[1]) if len(sys.argv) > 1 else 512
max20 = int(sys.argv[2]) if len(sys.argv) > 2 else 128
path = os.path.join(ROOT, "deepsphere-cosmo-tf2_b200", "deepsphere", "_lmax_table.json")
table = json.load(open(path)) if os.path.exists(path) else {}
for k, top in ((8, max8), (20, max20)):
    nside = 32
    while nside <= top:
        t = time.time()
        L = sparse.csr_matrix(SphereHealpix(nside, k=k).L, dtype=np.float64)
        key = utils.matrix_fingerprint(L)
        if key not in table:
            table[key] = utils.largest_eigenvalue(L)
            json.dump(table, open(path, "w"), indent=0, sort_keys=True)
        print(k, nside, key, table[key], f"{time.time() - t:.1f}s", flush=True)
        nside *= 2

# the masked survey of SURVEY 8d C3 (examples/advanced_tutorial.ipynb:137,211 scaled to nside 512): pixels within 1.5 rad of
# [1, 0, 0] padded with extend_indices to nside_out 64, and its two pooled levels - what bench.py's C3 entry builds
from deepsphere import healpix as hpx
ext = utils.extend_indices(hpx.query_disc(512, [1, 0, 0], 1.5), 512, 64)
for k in (20, 8):
    idx, nside = ext, 512
    for level in range(3):
        t = time.time()
        L = sparse.csr_matrix(SphereHealpix(nside, indexes=idx, k=k).L, dtype=np.float64)
        key = utils.matrix_fingerprint(L)
        if key not in table:
            table[key] = utils.largest_eigenvalue(L)
            json.dump(table, open(path, "w"), indent=0, sort_keys=True)
        print("C3", k, nside, len(idx), key, table[key], f"{time.time() - t:.1f}s", flush=True)
        idx, nside = hpx.coarsen_indices(idx, 1), nside // 2
