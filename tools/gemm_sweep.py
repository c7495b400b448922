"""Times xr_gemm_scatter alone at dimer-class sizes (development tool)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
from qodeapplications_b200.device import Device
dev = Device(0)
rng = numpy.random.default_rng(0)
for (M, N, K, scatter) in [(15272, 15272, 326, False), (15272, 15272, 326, True), (9984, 9984, 36, True), (8192, 8192, 2304, False),
                           (15272, 324, 324, False), (9984, 18, 5832, False)]:
    ld = K + (K & 1)
    A, B = dev.upload(rng.standard_normal((M, ld))), dev.upload(rng.standard_normal((N, ld)))
    if scatter:   # dimer-like layout: rows -> (i,j) pairs scattered in a (sqrt-ish) 4-index matrix
        n1 = int(M ** 0.5) + 1; n2 = int(N ** 0.5) + 1
        D = n1 * n2
        C = dev.empty((D * D,))
        offM = ((numpy.arange(M) // n1) * n2 * D + (numpy.arange(M) % n1) * n2).astype(numpy.int64)
        offN = ((numpy.arange(N) // n2) * D + (numpy.arange(N) % n2)).astype(numpy.int64)
        oM, oN = dev.upload(offM, numpy.int64), dev.upload(offN, numpy.int64)
        run = lambda: dev.ctx.gemm_scatter(M, N, K, 1.0, A, ld, B, ld, C, oM, 0, oN, False)
    else:
        C = dev.empty((M, N))
        run = lambda: dev.ctx.gemm_scatter(M, N, K, 1.0, A, ld, B, ld, C, None, N, None, False)
    run(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(json.dumps({"M": M, "N": N, "K": K, "scatter": scatter, "ms": best, "tflops": 2.0 * M * N * K / best / 1e9,
                      "write_GBs": 8.0 * M * N / best / 1e6}))
    del A, B, C

# xr_gemm_reduce (streamed dimer classes; nothing is written) at configs[4] per-rank shapes
for (M, N, K) in [(47483, 379864, 2306), (31190, 249516, 96), (7569, 60552, 2304), (15272, 15272, 326)]:
    ld = K + (K & 1)
    A = torch.randn((M, ld), dtype=torch.float64, device=dev.torch_device)
    B = torch.randn((N, ld), dtype=torch.float64, device=dev.torch_device)
    mom = dev.zeros((2,))
    run = lambda: dev.ctx.gemm_reduce(M, N, K, 1.0, A, ld, B, ld, mom)
    run(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(json.dumps({"op": "gemm_reduce", "M": M, "N": N, "K": K, "ms": best, "tflops": 2.0 * M * N * K / best / 1e9}))
    del A, B
