"""Small launches of the kernels added in round 2, for compute-sanitizer (memcheck + racecheck are run on this script):
persistent TMA GEMM (several tiles per CTA, K tails), trimer sample / threshold consumers, the FP64 probe, the device
inverse (Newton-Schulz + double-double polish) and a recorded/replayed get_xr_H."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
from qodeapplications_b200 import synth
from qodeapplications_b200.device import Device

dev = Device(0)
rng = numpy.random.default_rng(0)
for M, N, K in ((700, 900, 36), (300, 520, 326), (130, 70, 37)):
    A, B = dev.upload(rng.standard_normal((M, K + K % 2))), dev.upload(rng.standard_normal((N, K + K % 2)))
    C = dev.zeros((M, N))
    dev.ctx.gemm_scatter(M, N, K, 1.0, A, K + K % 2, B, K + K % 2, C, None, N, None, False)
    ref = dev.download(A)[:, :K] @ dev.download(B)[:, :K].T
    assert numpy.abs(dev.download(C) - ref).max() <= 1e-11 * numpy.abs(ref).max()
n, Pa, Pb, Pc = 18, 40, 50, 300
W, beta, gamma = rng.standard_normal((Pa, n * n)), rng.standard_normal((Pb, n)), rng.standard_normal((Pc, n))
dW, dB, dG = dev.upload(W), dev.upload(beta), dev.upload(gamma)
abc = numpy.stack([rng.integers(Pa, size=64), rng.integers(Pb, size=64), rng.integers(Pc, size=64)], axis=1).astype(numpy.int64)
out = dev.empty((64,))
dev.ctx.trimer_sample(n, Pa, Pb, Pc, 1.0, dW, n * n, dB, n, dG, n, numpy.ascontiguousarray(abc), out)
tables = [dev.upload(numpy.arange(Pa, dtype=numpy.int64) * Pb * Pc, numpy.int64), dev.upload(numpy.arange(Pb, dtype=numpy.int64) * Pc, numpy.int64),
          dev.upload(numpy.arange(Pc, dtype=numpy.int64), numpy.int64)]
cnt = dev.zeros((1,), dtype=torch.int64)
idx, val = dev.empty((1000,), dtype=torch.int64), dev.empty((1000,))
dev.ctx.trimer_threshold(n, Pa, Pb, Pc, 1.0, dW, n * n, dB, n, dG, n, 0, Pa, 12.0, tables[0], tables[1], tables[2], 1000, idx, val, cnt)
print("kept", int(cnt.cpu()[0]), "probe", dev.ctx.probe_fp64(0.01))
from qodeapplications_b200.hermitian.plan import plan
system = synth.make_system("toy", ops=synth.OPS_ORDER1, with_bior=True)
ch = system["charges"]
build = plan((system["symm"], system["bior"], system["nuc"]), system["densities"][:2], 1, [ch, ch], device=dev)
H1, H2 = build()
H1, H2 = build(system["densities"][:2])
torch.cuda.synchronize()
print("plan ok", build.graph is not None, build.launches)
