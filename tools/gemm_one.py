import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uninext_b200 import _cabi
lib = _cabi.load()
m, n, k = 44646, int(sys.argv[1]) if len(sys.argv) > 1 else 256, 256
a = torch.randn(m, k, device="cuda"); w = torch.randn(n, k, device="cuda") * 0.1; b = torch.randn(n, device="cuda")
c = torch.empty((m, n), dtype=torch.float32, device="cuda")
for _ in range(4):
    lib.msda_linear_tf32(a.data_ptr(), w.data_ptr(), b.data_ptr(), m, n, k, c.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
