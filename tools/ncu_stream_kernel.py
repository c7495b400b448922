"""One launch of xr_gemm_stream at a hermitian cfg1 precontraction shape (rho[ij, a, b, c, d, e] contracted over (a,b,d,e) with V:
88 x 18 rows, K = 324 x 324, N = 1; 1.33 GB of density read once), for ncu --set full:
    ncu --set full --clock-control none --import-source on -k regex:gemm_tma_stream -c 1 -o gpurun_out/r02g_stream python tools/ncu_stream_kernel.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qodeapplications_b200.device import Device
dev = Device(0)
n, P = 18, 88
rho = torch.randn((P, n * n, n, n * n), dtype=torch.float64, device=dev.torch_device)
V = torch.randn((1, n ** 4), dtype=torch.float64, device=dev.torch_device)
out = dev.empty((P * n, 1))
for _ in range(2):
    ok = dev.ctx.gemm_stream(P, n ** 5, n, n * n, n * n, n ** 3, n * n, 1, 1.0, rho, V, n ** 4, out, None, 1, None, False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
dev.ctx.gemm_stream(P, n ** 5, n, n * n, n * n, n ** 3, n * n, 1, 1.0, rho, V, n ** 4, out, None, 1, None, False)
e1.record()
torch.cuda.synchronize()
ref = torch.einsum("pacd,ad->pc", rho.reshape(P, n * n, n, n * n), V.reshape(n * n, n * n)).reshape(-1)
print("ok", ok, "ms", e0.elapsed_time(e1), "GB/s", rho.numel() * 8 / e0.elapsed_time(e1) / 1e6, "max err", float((out.reshape(-1) - ref).abs().max() / ref.abs().max()))
