// The eleven legacy scalar entry points of general-XRCC/H_contractions.c on the GPU.
//
// Same C ABI as the reference (int64 sizes, borrowed host double buffers, double by value).  Each
// call stages its operands into one pinned buffer, uploads them with a single cudaMemcpyAsync, runs
// one single-block reduction kernel (fixed summation tree: bit-reproducible) and downloads the
// scalar.  This exists so the reference's per-element call pattern keeps working unchanged and as
// the parity surface for it; the block-level path (xr_gemm_scatter / xr_trimer_stream) is the one
// that is fast.  There is no CPU fallback: without a device the functions return NaN.
#include "xr_common.cuh"
#include <cmath>
#include <cstring>
#include <mutex>

namespace {

enum ScalarOp {
    OP_DOT2 = 0,          // sum h[x] R[x]                                    (monomer_1e, monomer_extPot)
    OP_MONOMER_2E,        // sum V[p,q,r,s] R[p,q,s,r]
    OP_DIMER_2MIN2PLS,    // sum V[p1,q1,r2,s2] Rcc1[p1,q1] Raa2[s2,r2]
    OP_DIMER_1E,          // sum h[p1,q2] Rc1[p1] Ra2[q2]
    OP_DIMER_2E_A,        // sum V1112[p1,q1,r1,s2] Rcca1[q1,p1,r1] Ra2[s2]
    OP_DIMER_2E_B,        // sum V1222[p1,q2,r2,s2] Rc1[p1] Rcaa2[q2,s2,r2]
    OP_DIMER_EXEX,        // sum V[p1,q2,r1,s2] Rca1[p1,r1] Rca2[q2,s2]
    OP_TRIMER_2MIN,       // sum V[p1,q1,r2,s3] Rcc1[q1,p1] Ra2[r2] Ra3[s3]
    OP_TRIMER_2PLS,       // sum V[r2,s3,p1,q1] Rc2[r2] Rc3[s3] Raa1[q1,p1]
    OP_TRIMER_EX          // sum V[p1,r2,q1,s3] Rca1[p1,q1] Rc2[r2] Ra3[s3]
};

struct ScalarArgs {
    int op;
    int64_t n1, n2, n3;
    const double* V;     // the integral block (or h)
    const double* R1;
    const double* R2;
    const double* R3;
    int64_t total;       // number of integral elements
    double* out;
};

__device__ __forceinline__ double scalar_term(const ScalarArgs& a, int64_t x) {
    const int64_t n1 = a.n1, n2 = a.n2, n3 = a.n3;
    const double v = a.V[x];
    switch (a.op) {
        case OP_DOT2:
            return v * a.R1[x];
        case OP_MONOMER_2E: {
            int64_t s = x % n1, r = (x / n1) % n1, pq = x / (n1 * n1);
            return v * a.R1[(pq * n1 + s) * n1 + r];
        }
        case OP_DIMER_2MIN2PLS: {
            int64_t s = x % n2, r = (x / n2) % n2, pq = x / (n2 * n2);
            return v * a.R1[pq] * a.R2[s * n2 + r];
        }
        case OP_DIMER_1E: {
            int64_t q = x % n2, p = x / n2;
            return v * a.R1[p] * a.R2[q];
        }
        case OP_DIMER_2E_A: {
            int64_t s = x % n2, r = (x / n2) % n1, q = (x / (n2 * n1)) % n1, p = x / (n2 * n1 * n1);
            return v * a.R1[(q * n1 + p) * n1 + r] * a.R2[s];
        }
        case OP_DIMER_2E_B: {
            int64_t s = x % n2, r = (x / n2) % n2, q = (x / (n2 * n2)) % n2, p = x / (n2 * n2 * n2);
            return v * a.R1[p] * a.R2[(q * n2 + s) * n2 + r];
        }
        case OP_DIMER_EXEX: {
            int64_t s = x % n2, r = (x / n2) % n1, q = (x / (n2 * n1)) % n2, p = x / (n2 * n1 * n2);
            return v * a.R1[p * n1 + r] * a.R2[q * n2 + s];
        }
        case OP_TRIMER_2MIN: {
            int64_t s = x % n3, r = (x / n3) % n2, q = (x / (n3 * n2)) % n1, p = x / (n3 * n2 * n1);
            return v * a.R1[q * n1 + p] * a.R2[r] * a.R3[s];
        }
        case OP_TRIMER_2PLS: {
            int64_t q = x % n1, p = (x / n1) % n1, s = (x / (n1 * n1)) % n3, r = x / (n1 * n1 * n3);
            return v * a.R2[r] * a.R3[s] * a.R1[q * n1 + p];
        }
        default: {  // OP_TRIMER_EX
            int64_t s = x % n3, q = (x / n3) % n1, r = (x / (n3 * n1)) % n2, p = x / (n3 * n1 * n2);
            return v * a.R1[p * n1 + q] * a.R2[r] * a.R3[s];
        }
    }
}

constexpr int SCALAR_THREADS = 1024;

__global__ void __launch_bounds__(SCALAR_THREADS, 1) scalar_contract_kernel(const ScalarArgs a, double scale, int accumulate) {
    __shared__ double red[SCALAR_THREADS / 32];
    double acc = 0.0;
    for (int64_t x = threadIdx.x; x < a.total; x += SCALAR_THREADS) acc += scalar_term(a, x);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < SCALAR_THREADS / 32; ++w) tot += red[w];
        tot *= scale;
        *a.out = accumulate ? *a.out + tot : tot;
    }
}

std::mutex g_mutex;
xr_ctx* g_ctx = nullptr;   // default context of the legacy ABI (device 0), created on first use

struct Operand {
    const double* host;
    int64_t count;
};

// Stages operands into pinned memory, uploads, runs `n_ops` kernels accumulating into one scalar.
struct Job {
    int op;
    double scale;
    int v, r1, r2, r3;   // operand slots (-1 = unused)
    int64_t n1, n2, n3;
};

double run_jobs(const Operand* ops, int n_operands, const Job* jobs, int n_jobs) {
    std::lock_guard<std::mutex> lock(g_mutex);
    if (!g_ctx) {
        if (xr_ctx_create(0, nullptr, 1, &g_ctx) != XR_OK) return NAN;
    }
    xr_ctx* ctx = g_ctx;
    int64_t offsets[8];
    int64_t total = 2;   // slot 0..1: result (+pad)
    for (int i = 0; i < n_operands; ++i) {
        if (!ops[i].host) {
            xr_set_error("legacy scalar call: null operand %d", i);
            return NAN;
        }
        offsets[i] = total;
        total += (ops[i].count + 1) & ~int64_t(1);
    }
    size_t bytes = (size_t)total * sizeof(double);
    if (bytes > ctx->pinned_bytes) {
        if (ctx->pinned) cudaFreeHost(ctx->pinned);
        ctx->pinned = nullptr;
        ctx->pinned_bytes = 0;
        if (cudaMallocHost(&ctx->pinned, bytes * 2) != cudaSuccess) {
            xr_set_error("cudaMallocHost(%zu) failed", bytes * 2);
            return NAN;
        }
        ctx->pinned_bytes = bytes * 2;
    }
    if (xr_ensure_scratch(ctx, bytes) != XR_OK) return NAN;
    double* stage = static_cast<double*>(ctx->pinned);
    double* dev = static_cast<double*>(ctx->scratch);
    stage[0] = stage[1] = 0.0;
    for (int i = 0; i < n_operands; ++i) memcpy(stage + offsets[i], ops[i].host, (size_t)ops[i].count * sizeof(double));
    if (cudaMemcpyAsync(dev, stage, bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
        xr_set_error("legacy scalar call: upload failed");
        return NAN;
    }
    for (int j = 0; j < n_jobs; ++j) {
        ScalarArgs a{};
        a.op = jobs[j].op;
        a.n1 = jobs[j].n1;
        a.n2 = jobs[j].n2;
        a.n3 = jobs[j].n3;
        a.V = dev + offsets[jobs[j].v];
        a.R1 = jobs[j].r1 >= 0 ? dev + offsets[jobs[j].r1] : nullptr;
        a.R2 = jobs[j].r2 >= 0 ? dev + offsets[jobs[j].r2] : nullptr;
        a.R3 = jobs[j].r3 >= 0 ? dev + offsets[jobs[j].r3] : nullptr;
        a.total = ops[jobs[j].v].count;
        a.out = dev;
        scalar_contract_kernel<<<1, SCALAR_THREADS, 0, ctx->stream>>>(a, jobs[j].scale, j > 0);
        ctx->launches++;
    }
    if (cudaMemcpyAsync(stage, dev, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        xr_set_error("legacy scalar call: kernel or download failed: %s", cudaGetErrorString(cudaGetLastError()));
        return NAN;
    }
    return stage[0];
}

}  // namespace

extern "C" {

PyFloat monomer_1e(PyInt n, Double* Rca, Double* h) {
    Operand ops[] = {{h, n * n}, {Rca, n * n}};
    Job jobs[] = {{OP_DOT2, 1.0, 0, 1, -1, -1, n, n, n}};
    return run_jobs(ops, 2, jobs, 1);
}

PyFloat monomer_extPot(PyInt n, Double* Rca, Double* h) { return monomer_1e(n, Rca, h); }

PyFloat monomer_2e(PyInt n, Double* Rccaa, Double* V) {
    Operand ops[] = {{V, n * n * n * n}, {Rccaa, n * n * n * n}};
    Job jobs[] = {{OP_MONOMER_2E, 1.0, 0, 1, -1, -1, n, n, n}};
    return run_jobs(ops, 2, jobs, 1);
}

PyFloat monomer(PyInt n, Double* Rca, Double* Rccaa, Double* h, Double* V) {
    Operand ops[] = {{V, n * n * n * n}, {Rccaa, n * n * n * n}, {h, n * n}, {Rca, n * n}};
    Job jobs[] = {{OP_MONOMER_2E, 1.0, 0, 1, -1, -1, n, n, n}, {OP_DOT2, 1.0, 2, 3, -1, -1, n, n, n}};
    return run_jobs(ops, 4, jobs, 2);
}

PyFloat dimer_2min2pls(PyInt n1, PyInt n2, Double* Rcc1, Double* Raa2, Double* V) {
    Operand ops[] = {{V, n1 * n1 * n2 * n2}, {Rcc1, n1 * n1}, {Raa2, n2 * n2}};
    Job jobs[] = {{OP_DIMER_2MIN2PLS, 1.0, 0, 1, 2, -1, n1, n2, 0}};
    return run_jobs(ops, 3, jobs, 1);
}

PyFloat dimer_1min1pls_1e(PyInt n1, PyInt n2, Double* Rc1, Double* Ra2, Double* h) {
    Operand ops[] = {{h, n1 * n2}, {Rc1, n1}, {Ra2, n2}};
    Job jobs[] = {{OP_DIMER_1E, 1.0, 0, 1, 2, -1, n1, n2, 0}};
    return run_jobs(ops, 3, jobs, 1);
}

PyFloat dimer_1min1pls_2e(PyInt n1, PyInt n2, Double* Rc1, Double* Rcca1, Double* Ra2, Double* Rcaa2, Double* V1112,
                          Double* V1222) {
    Operand ops[] = {{V1112, n1 * n1 * n1 * n2}, {V1222, n1 * n2 * n2 * n2}, {Rc1, n1}, {Rcca1, n1 * n1 * n1},
                     {Ra2, n2},                  {Rcaa2, n2 * n2 * n2}};
    Job jobs[] = {{OP_DIMER_2E_A, 2.0, 0, 3, 4, -1, n1, n2, 0}, {OP_DIMER_2E_B, 2.0, 1, 2, 5, -1, n1, n2, 0}};
    return run_jobs(ops, 6, jobs, 2);
}

PyFloat dimer_ExEx(PyInt n1, PyInt n2, Double* Rca1, Double* Rca2, Double* V) {
    Operand ops[] = {{V, n1 * n2 * n1 * n2}, {Rca1, n1 * n1}, {Rca2, n2 * n2}};
    Job jobs[] = {{OP_DIMER_EXEX, 4.0, 0, 1, 2, -1, n1, n2, 0}};
    return run_jobs(ops, 3, jobs, 1);
}

PyFloat trimer_2min1pls1pls(PyInt n1, PyInt n2, PyInt n3, Double* Rcc1, Double* Ra2, Double* Ra3, Double* V) {
    Operand ops[] = {{V, n1 * n1 * n2 * n3}, {Rcc1, n1 * n1}, {Ra2, n2}, {Ra3, n3}};
    Job jobs[] = {{OP_TRIMER_2MIN, 2.0, 0, 1, 2, 3, n1, n2, n3}};
    return run_jobs(ops, 4, jobs, 1);
}

PyFloat trimer_2pls1min1min(PyInt n1, PyInt n2, PyInt n3, Double* Raa1, Double* Rc2, Double* Rc3, Double* V) {
    Operand ops[] = {{V, n2 * n3 * n1 * n1}, {Raa1, n1 * n1}, {Rc2, n2}, {Rc3, n3}};
    Job jobs[] = {{OP_TRIMER_2PLS, 2.0, 0, 1, 2, 3, n1, n2, n3}};
    return run_jobs(ops, 4, jobs, 1);
}

PyFloat trimer_Ex1min1pls(PyInt n1, PyInt n2, PyInt n3, Double* Rca1, Double* Rc2, Double* Ra3, Double* V) {
    Operand ops[] = {{V, n1 * n2 * n1 * n3}, {Rca1, n1 * n1}, {Rc2, n2}, {Ra3, n3}};
    Job jobs[] = {{OP_TRIMER_EX, 4.0, 0, 1, 2, 3, n1, n2, n3}};
    return run_jobs(ops, 4, jobs, 1);
}

}  // extern "C"
