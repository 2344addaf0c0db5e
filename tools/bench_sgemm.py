"""Time the exact sequential-k SGEMM at coarse-quantizer shapes:  python tools/bench_sgemm.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from densephrases_b200 import _lib as L
for n, m in [(64, 4096), (64, 768), (128, 8192), (256, 16384), (512, 32768), (1024, 65536)]:
    X = torch.randn(n, 768, device="cuda"); W = torch.randn(m, 768, device="cuda"); out = torch.empty(n, m, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(3): L.check(L.lib().dph_sgemm_nt_seq(X.data_ptr(), n, W.data_ptr(), m, 768, out.data_ptr(), st))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): L.lib().dph_sgemm_nt_seq(X.data_ptr(), n, W.data_ptr(), m, 768, out.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    ref = (X.double() @ W.double().T)
    print(f"n={n:5d} m={m:6d}: {ms*1000:8.1f} us  {2*n*m*768/ms/1e9:6.1f} TFLOP/s  max|err| {(out.double()-ref).abs().max().item():.2e}")
