"""TEST TOOLING (not collected by pytest, not a product path): runs `-m gpu` test files on a machine WITHOUT a GPU with the
C-ABI ops of deepsphere._ops replaced by float64 torch / scipy stand-ins built on the oracle.  It cannot say anything about
the kernels; it catches Python-level breakage of the host code the GPU tests drive (layer classes, HealpyGCNN bookkeeping,
autograd glue, the tests' own reference arithmetic) before GPU minutes are spent.

    python tests/mocked_gpu_run.py tests/test_gpu_parity.py tests/test_gpu_model.py tests/test_gpu_zz_next.py

Tests that create CUDA tensors / generators directly, count launches of the real library or assert bit-exact summation
orders fail here by construction.
"""
import sys
import os
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [os.path.join(ROOT, 'deepsphere-cosmo-tf2_b200'), ROOT, HERE]
import numpy as np, torch, pytest
from scipy import sparse
from deepsphere import _ops, _native as nat
from oracle import deepsphere_oracle as orc

ACT = {0: lambda v: v, 1: torch.relu, 2: torch.nn.functional.elu, 3: torch.sigmoid, 4: torch.tanh, 5: torch.nn.functional.softplus}
def plan_csr(plan):
    return sparse.csr_matrix((plan.values.astype(np.float64), (plan.indices[:, 0], plan.indices[:, 1])), shape=plan.shape)
cnt = [0]
def graph_conv(x, kernel, bias, plan, recursion, K, act=0, mode=0):
    cnt[0] += 1
    z = orc.torch_cpu_graph_conv(x.double(), plan_csr(plan), kernel.double(), K, "chebyshev" if recursion == 0 else "monomial")
    if bias is not None: z = z + bias.double().reshape(1, 1, -1)
    return ACT[act](z).float()
def bias_act(z, bias, act):
    cnt[0] += 1
    if bias is not None: z = z + bias.reshape(1, 1, -1) if z.dim() == 3 else z + bias
    return ACT[act](z)
def pool(x, p, pool_type):
    cnt[0] += 1
    B, M, F = x.shape; r = 4 ** p
    v = x.reshape(B, M // r, r, F)
    return v.max(dim=2).values if pool_type == nat.POOL_MAX else v.mean(dim=2)
def pseudo_conv(x, w, bias, p, Fout, act=0, mode=0, transpose=False):
    cnt[0] += 1
    B, M, Fin = x.shape; r = 4 ** p
    if not transpose:
        z = x.reshape(B, M // r, r * Fin) @ w.reshape(r * Fin, Fout)
    else:
        z = torch.einsum("bmf,cof->bmco", x, w.reshape(r, Fout, Fin)).reshape(B, M * r, Fout)
    if bias is not None: z = z + bias
    return ACT[act](z)
class Spmm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, plan, T):
        ctx.plan, ctx.T = plan, T
        A = plan_csr(plan); A = A.T if T else A
        return torch.tensor(np.stack([A @ x[b].detach().double().numpy() for b in range(x.shape[0])])).float()

    @staticmethod
    def backward(ctx, dy):
        A = plan_csr(ctx.plan); A = A if ctx.T else A.T
        return torch.tensor(np.stack([A @ dy[b].double().numpy() for b in range(dy.shape[0])])).float(), None, None
def spmm(plan, x, alpha=1.0, prev=None, beta=0.0, add=None, gamma=0.0, transpose=False):
    cnt[0] += 1
    out = alpha * Spmm.apply(x, plan, transpose)
    if prev is not None: out = out + beta * prev
    if add is not None: out = out + gamma * add
    return out
def basis(plan, x, K, recursion=0, transpose=False):
    cnt[0] += 1
    A = plan_csr(plan); A = A.T if transpose else A
    X = orc._basis_stack(x.double().numpy(), A, K, "chebyshev" if recursion == 0 else "monomial", np.float64)
    B, M, F = x.shape
    return torch.tensor(X.reshape(B, M, F, K)).permute(3, 0, 1, 2)[1:].float().contiguous()
_ops.graph_conv, _ops.bias_act, _ops.pool, _ops.pseudo_conv, _ops.spmm, _ops.basis = graph_conv, bias_act, pool, pseudo_conv, spmm, basis
def sparse_matmul(plan, x):
    cnt[0] += 1
    return Spmm.apply(x, plan, False)
_ops.sparse_matmul = sparse_matmul
nat.launch_count = lambda: cnt[0]
nat.GraphPlan.info = lambda self, d=0: {"lattice": 1, "tail_rows": 2, "symmetric": 0, "nnz": len(self.values), "M": self.shape[0]}
torch.Tensor.cuda = lambda self, *a, **k: self
torch.cuda.synchronize = lambda *a, **k: None
if __name__ == "__main__":
    sys.exit(pytest.main(sys.argv[1:] + ["-m", "gpu", "-q", "-p", "no:cacheprovider"]))
