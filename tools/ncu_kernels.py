"""One launch of each hot kernel at the shapes the benchmarks use, for `ncu --set full` (see profiles/README or DESIGN 4):

    ncu --set full --clock-control none --import-source on -k regex:'trimer_stream|gemm_tma' -o gpurun_out/r01j_kernels \
        python tools/ncu_kernels.py

1. trimer_stream_kernel<4,2,2>   cfg4 class shape, a slab of 384 W rows (n=18, Pb=Pc=9984)
2. gemm_tma_scatter_kernel<false> d=0 dimer class of cfg4 (15272^2, K=326) with the 4-index offset-table epilogue
3. gemm_tma_scatter_kernel<false> d=+-1 class (9984^2, K=36): HBM-write bound
4. gemm_tma_scatter_kernel<true>  streamed class (xr_gemm_reduce), configs[4] K=2306, 15272 x 60000
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
from qodeapplications_b200.device import Device
from qodeapplications_b200 import lib as xr

dev = Device(0)
rand = lambda *shape: torch.randn(shape, dtype=torch.float64, device=dev.torch_device)

n, Pa, Pb, Pc = 18, 384, 9984, 9984
W, B, G, mom = rand(Pa, n * n), rand(Pb, n), rand(Pc, n), dev.zeros((2,))
dev.ctx.trimer_stream(n, Pa, Pb, Pc, 1.0, W, n * n, B, n, G, n, 0, Pa, xr.TRIMER_REDUCE, mom)

for M, K in ((15272, 326), (9984, 36)):
    A, Bm = rand(M, K), rand(M, K)
    n1 = int(M ** 0.5) + 1
    D = n1 * n1
    C = dev.empty((D * D,))
    offM = ((numpy.arange(M) // n1) * n1 * D + (numpy.arange(M) % n1) * n1).astype(numpy.int64)
    offN = ((numpy.arange(M) // n1) * D + (numpy.arange(M) % n1)).astype(numpy.int64)
    dev.ctx.gemm_scatter(M, M, K, 1.0, A, K, Bm, K, C, dev.upload(offM, numpy.int64), 0, dev.upload(offN, numpy.int64), False)
    del A, Bm, C

A, Bm = rand(15272, 2306), rand(60000, 2306)
dev.ctx.gemm_reduce(15272, 60000, 2306, 1.0, A, 2306, Bm, 2306, mom)
torch.cuda.synchronize()
print("done")
