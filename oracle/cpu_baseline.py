"""Timed CPU evaluation of the reference's per-element H build on a bounded sample.

TEST/BENCH INFRASTRUCTURE (bench.py's cpu_baseline leg and `--impl reference` arm only).
What is timed is the reference's algorithm as the reference runs it: one Python call per matrix
element (the control flow of general-XRCC/build_H.py:42-188, restated in
oracle/general_oracle.element_oracle) into the reference's own C kernels compiled -O2
(oracle/_ref/libH_contractions_ref.so, kind "reference"; falls back to our C restatement
oracle/libxr_oracle.so, kind "port", if that file is absent), optionally fanned out over a
multiprocessing.Pool like general-XRCC/test_H.py:131.

The full workload cannot be run on a CPU (cfg4 is ~1e6 core-hours), so a fixed random sample of
charge-allowed elements of every class is timed and the whole-workload time is extrapolated with
the exact per-class element counts.

`factored_numpy` is the second, best-effort CPU figure: the factored algorithm of the GPU path in NumPy/BLAS on all cores.
"""
import itertools
import os
import time
import multiprocessing
import numpy

from . import general_oracle as go

_STATE = {}


def _pick_c_library():
    try:
        return go.c_contractions("ref")
    except (FileNotFoundError, OSError):
        return go.c_contractions("port")


def _allowed_dimer_tuples(fragments, m1, m2, d1, count, rng):
    f1, f2 = fragments[m1], fragments[m2]
    n1, n2 = go._n_states(f1), go._n_states(f2)
    combos = [(c1, c2) for c1 in n1 for c2 in n2 if (c1 - d1) in n1 and (c2 + d1) in n2]
    out = []
    for _ in range(count):
        c1, c2 = combos[rng.integers(len(combos))]
        k1, k2 = c1 - d1, c2 + d1
        I = ((c1, int(rng.integers(n1[c1]))), (c2, int(rng.integers(n2[c2]))))
        J = ((k1, int(rng.integers(n1[k1]))), (k2, int(rng.integers(n2[k2]))))
        out.append(("dimer", (m1, m2), I, J))
    return out


_TRIMER_PATTERNS = {
    "2min": [p for p in set(itertools.permutations((-2, 1, 1)))],
    "2pls": [p for p in set(itertools.permutations((2, -1, -1)))],
    "ex": [p for p in set(itertools.permutations((0, -1, 1)))],
}


def _allowed_trimer_tuples(fragments, ms, kind, count, rng):
    ns = [go._n_states(fragments[m]) for m in ms]
    combos = []
    for pattern in sorted(_TRIMER_PATTERNS[kind]):
        for bra in itertools.product(*ns):
            ket = tuple(b - d for b, d in zip(bra, pattern))
            if all(k in n for k, n in zip(ket, ns)):
                combos.append((bra, ket))
    out = []
    for _ in range(count):
        bra, ket = combos[rng.integers(len(combos))]
        I = tuple((c, int(rng.integers(ns[x][c]))) for x, c in enumerate(bra))
        J = tuple((c, int(rng.integers(ns[x][c]))) for x, c in enumerate(ket))
        out.append(("trimer", ms, I, J))
    return out


def make_sample(system, per_class, seed=0):
    """{class key: [task, ...]}; class keys match build_matrix_elements.element_counts()."""
    rng = numpy.random.default_rng(seed)
    frags = system["fragments"]
    F = len(frags)
    sample = {}
    for d1 in (-2, -1, 0, 1, 2):
        try:
            sample[("dimer", d1)] = _allowed_dimer_tuples(frags, 0, 1, d1, per_class, rng)
        except ValueError:
            pass
    if F >= 3:
        for kind in ("2min", "2pls", "ex"):
            sample[("trimer", kind)] = _allowed_trimer_tuples(frags, (0, 1, 2), kind, per_class, rng)
    return sample


def _evaluate(task):
    eo = _STATE["oracle"]
    what, frags, I, J = task
    return eo.dimer(frags, I, J) if what == "dimer" else eo.trimer(frags, I, J)


def _evaluate_chunk(tasks):
    return [_evaluate(t) for t in tasks]


def prepare(system):
    _STATE["oracle"] = go.element_oracle(system["fragments"], system["symm"], system["nuc"], _pick_c_library())
    return _STATE["oracle"].c.kind


def time_sample(sample, cores=1, pool=None):
    """seconds per class for one pass over the sample (wall clock, `cores` worker processes)"""
    seconds = {}
    for key, tasks in sample.items():
        t0 = time.perf_counter()
        if cores == 1:
            _evaluate_chunk(tasks)
        else:
            chunk = max(1, len(tasks) // (cores * 4))
            pool.map(_evaluate_chunk, [tasks[i:i + chunk] for i in range(0, len(tasks), chunk)])
        seconds[key] = time.perf_counter() - t0
    return seconds


def extrapolate(seconds, sample, counts):
    """whole-workload seconds from per-class sample timings and exact element counts"""
    total = 0.0
    for key, count in counts.items():
        if key in seconds and count:
            total += count * seconds[key] / len(sample[key])
    return total


def make_pool(cores):
    ctx = multiprocessing.get_context("fork")      # workers inherit the prepared oracle (no CUDA in this process)
    return ctx.Pool(cores)


def factored_numpy(n_states, n_orb, seconds=5.0, seed=0):
    """Best-effort CPU figure beside the reference-faithful one: the FACTORED trimer-class algorithm the GPU
    path uses -- X = beta . W[a] (GEMM), T = X . gamma^T (GEMM), sum of T^2 -- in NumPy/BLAS on all host cores, on factors
    of the workload's own class shape (Pb = P(-1), Pc = P(+1) pair counts; random values: timing only), for about `seconds`
    of wall clock.  Returns (algorithmic TFLOP/s = 2 n Pb Pc per row of W, rows done, seconds)."""
    P = lambda d: sum(n_states[c] * n_states[c - d] for c in n_states if (c - d) in n_states)
    Pb, Pc = P(-1), P(+1)
    rng = numpy.random.default_rng(seed)
    W = rng.standard_normal((8, n_orb, n_orb))
    beta, gamma_t = rng.standard_normal((Pb, n_orb)), numpy.ascontiguousarray(rng.standard_normal((Pc, n_orb)).T)
    chunk = 2048
    T = numpy.empty((chunk, Pc))
    rows, moment, t0 = 0, 0.0, time.perf_counter()
    while True:
        X = beta @ W[rows % 8]
        for b0 in range(0, Pb, chunk):
            t = numpy.matmul(X[b0:b0 + chunk], gamma_t, out=T[:min(chunk, Pb - b0)])
            moment += float(numpy.vdot(t, t))
        rows += 1
        elapsed = time.perf_counter() - t0
        if elapsed >= seconds:
            break
    return 2.0 * n_orb * Pb * Pc * rows / elapsed / 1e12, rows, elapsed
