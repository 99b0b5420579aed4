"""Direct (fast-diagonalisation) Helmholtz solve at C2: hand-written DGEMM vs cuBLAS."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-topt_b200"))
from sktopt.filters._fastdiag import FastDiagHelmholtz
axes = (np.linspace(0, 8, 140), np.linspace(0, 6, 105), np.linspace(0, 4, 71))
fd = FastDiagHelmholtz(axes); fd.set_radius(0.01)
b = torch.randn(140 * 105 * 71, dtype=torch.float64, device="cuda"); out = torch.empty_like(b)
def t(reps=20):
    for _ in range(3): fd.solve(b, out=out)
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fd.solve(b, out=out)
    e.record(); torch.cuda.synchronize()
    return a.elapsed_time(e) / reps
print("dgemm.cu  %.4f ms" % t())
os.environ["SKTOPT_B200_FD_TORCH"] = "1"
print("cuBLAS    %.4f ms" % t())
