// Transition-density tensors from CI vectors -- the step before the H build (general-XRCC/density_tensors.c:142-556,
// driven by build_density_tensors.py:70-157):
//
//   rho[I,J,(i_0..i_{k-1})] = sum_Q  parity * z_bra[I, P] * z_ket[J, Q],   |P> = +- op_0(i_0) ... op_{k-1}(i_{k-1}) |Q>
//
// The reference walks ket configurations and scatters; here every thread OWNS output elements (one tensor index, one bra
// state, a tile of ket states), walks the ket configurations in the reference's order and gathers -- no atomics, and the
// floating-point summation order (and so every bit of the result) is the reference's.  Configurations are 64-bit
// occupation masks; the rank of the bra configuration (find_config_index, density_tensors.c:29-64) comes from a prefix
// table of binomials.  Integer/bit work + FP64 adds, HBM-write bound on the output (N_bra N_ket dim^k doubles).
#include "xr_common.cuh"
#include <mutex>
#include <vector>

namespace {

constexpr int JT = 8;          // ket states per thread (register accumulators)
constexpr int MAX_OPS = 4;

struct DensityParams {
    int k;                     // number of field operators
    int create[MAX_OPS];       // 1 = creation, 0 = annihilation, in string order (applied right to left)
    int64_t dim, T;            // spin orbitals, dim^k
    int64_t n_bra, n_ket, ncfg_bra, ncfg_ket;
    int n_orbs, n_core, n_val_elec_bra, S;
    unsigned long long core_mask;
    const double* z_bra;
    const double* z_ket;
    const unsigned long long* ket_masks;
    const long long* G;        // [n_val_elec_bra][S+1]: G[i][m] = sum_{n=1..m} C(S-n, n_val_elec_bra-i-1)
    double* rho;
};

__device__ __forceinline__ long long config_rank(unsigned long long mask, const DensityParams& p) {
    const int nv = p.n_orbs - p.n_core;                                  // valence orbitals per spin
    const unsigned long long lo = (mask >> p.n_core) & ((1ull << nv) - 1ull);
    const unsigned long long hi = (mask >> (p.n_orbs + p.n_core)) & ((1ull << nv) - 1ull);
    unsigned long long val = lo | (hi << nv);
    long long index = 0;
    int prev = -1;
    for (int i = 0; i < p.n_val_elec_bra; ++i) {
        const int c = __ffsll((long long)val) - 1;
        val &= val - 1ull;
        const long long* g = p.G + (size_t)i * (p.S + 1);
        index += g[c] - g[prev + 1];
        prev = c;
    }
    return index;
}

__global__ void __launch_bounds__(256) density_kernel(const DensityParams p) {
    const int64_t jtiles = (p.n_ket + JT - 1) / JT;
    const int64_t total = p.T * p.n_bra * jtiles;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t idx = t % p.T;
        const int64_t rest = t / p.T;
        const int64_t I = rest % p.n_bra, J0 = (rest / p.n_bra) * JT;
        int orb[MAX_OPS];
        {
            int64_t rem = idx;
            for (int o = p.k - 1; o >= 0; --o) {
                orb[o] = (int)(rem % p.dim);
                rem /= p.dim;
            }
        }
        double acc[JT];
#pragma unroll
        for (int j = 0; j < JT; ++j) acc[j] = J0 + j < p.n_ket ? p.rho[((I * p.n_ket) + J0 + j) * p.T + idx] : 0.0;
        for (int64_t Q = 0; Q < p.ncfg_ket; ++Q) {
            unsigned long long mask = p.ket_masks[Q];
            int flips = 0;
            bool alive = true;
            for (int o = p.k - 1; o >= 0 && alive; --o) {
                const unsigned long long bit = 1ull << orb[o];
                const int below = __popcll(mask & (bit - 1ull));
                const int n = __popcll(mask);
                if (p.create[o]) {
                    alive = !(mask & bit);
                    flips += n - below;                 // density_tensors.c:104-109: shifts to insert at position `below`
                    mask |= bit;
                } else {
                    alive = (mask & bit) != 0ull;
                    flips += n - 1 - below;             // density_tensors.c:88-92: shifts to move it to the end
                    mask &= ~bit;
                }
            }
            if (!alive || (mask & p.core_mask) != p.core_mask) continue;
            const long long P = config_rank(mask, p);
            const double zb = p.z_bra[I * p.ncfg_bra + P];
            const double left = (flips & 1) ? -zb : zb;                  // parity * zI_P (exact)
#pragma unroll
            for (int j = 0; j < JT; ++j)
                if (J0 + j < p.n_ket)                                       // (parity*zI_P)*zJ_Q rounded, then added: no FMA, as the C
                    acc[j] = __dadd_rn(acc[j], __dmul_rn(left, p.z_ket[(J0 + j) * p.ncfg_ket + Q]));
        }
#pragma unroll
        for (int j = 0; j < JT; ++j)
            if (J0 + j < p.n_ket) p.rho[((I * p.n_ket) + J0 + j) * p.T + idx] = acc[j];
    }
}

long long binomial(int n, int k) {
    if (k < 0 || k > n) return 0;
    long long r = 1;
    for (int i = 1; i <= k; ++i) r = r * (n - k + i) / i;
    return r;
}

}  // namespace

extern "C" int xr_density_tensor(xr_ctx* ctx, const char* ops, double* rho, int64_t n_bra_states, int64_t n_ket_states,
                                 const double* z_bra, int64_t n_configs_bra, const double* z_ket, int64_t n_configs_ket,
                                 const uint64_t* ket_masks, int64_t n_elec_bra, int64_t n_elec_ket, int64_t n_orbs, int64_t n_core) {
    XR_REQUIRE(ctx && ops, "xr_density_tensor: null ctx or operator string");
    DensityParams p{};
    int dchg = 0;
    for (p.k = 0; ops[p.k]; ++p.k) {
        XR_REQUIRE(p.k < MAX_OPS, "xr_density_tensor: more than %d operators in '%s'", MAX_OPS, ops);
        XR_REQUIRE(ops[p.k] == 'c' || ops[p.k] == 'a', "xr_density_tensor: operator string '%s' must consist of c and a", ops);
        p.create[p.k] = ops[p.k] == 'c';
        dchg += p.create[p.k] ? 1 : -1;
    }
    XR_REQUIRE(p.k >= 1, "xr_density_tensor: empty operator string");
    XR_REQUIRE(n_orbs >= 1 && 2 * n_orbs <= 64 && n_core >= 0 && n_core <= n_orbs, "xr_density_tensor: need 1 <= 2*n_orbs <= 64, 0 <= n_core <= n_orbs");
    XR_REQUIRE(n_elec_ket + dchg == n_elec_bra, "xr_density_tensor: '%s' does not connect %lld to %lld electrons", ops,
               (long long)n_elec_ket, (long long)n_elec_bra);
    XR_REQUIRE(n_elec_bra >= 2 * n_core && n_elec_bra <= 2 * n_orbs, "xr_density_tensor: bra electron count out of range");
    if (n_bra_states <= 0 || n_ket_states <= 0 || n_configs_ket <= 0) return XR_OK;
    XR_REQUIRE(rho && z_bra && z_ket && ket_masks, "xr_density_tensor: null pointer");
    p.dim = 2 * n_orbs;
    p.T = 1;
    for (int o = 0; o < p.k; ++o) p.T *= p.dim;
    p.n_bra = n_bra_states; p.n_ket = n_ket_states; p.ncfg_bra = n_configs_bra; p.ncfg_ket = n_configs_ket;
    p.n_orbs = (int)n_orbs; p.n_core = (int)n_core;
    p.n_val_elec_bra = (int)(n_elec_bra - 2 * n_core);
    p.S = (int)(2 * (n_orbs - n_core));
    XR_REQUIRE(binomial(p.S, p.n_val_elec_bra) == n_configs_bra, "xr_density_tensor: n_configs_bra=%lld is not C(%d,%d): the bra "
               "coefficients must span every valence configuration in find_config_index order", (long long)n_configs_bra, p.S, p.n_val_elec_bra);
    p.core_mask = 0;
    for (int i = 0; i < n_core; ++i) p.core_mask |= (1ull << i) | (1ull << (n_orbs + i));
    p.z_bra = z_bra; p.z_ket = z_ket; p.ket_masks = reinterpret_cast<const unsigned long long*>(ket_masks); p.rho = rho;
    // prefix table of the ranking binomials (density_tensors.c:34-45): G[i][m] = sum_{n=1..m} C(S-n, e-i-1)
    std::vector<long long> G((size_t)(p.n_val_elec_bra > 0 ? p.n_val_elec_bra : 1) * (p.S + 1), 0);
    for (int i = 0; i < p.n_val_elec_bra; ++i)
        for (int m = 1; m <= p.S; ++m) G[(size_t)i * (p.S + 1) + m] = G[(size_t)i * (p.S + 1) + m - 1] + binomial(p.S - m, p.n_val_elec_bra - i - 1);
    int rc = xr_ensure_scratch(ctx, G.size() * sizeof(long long));
    if (rc != XR_OK) return rc;
    XR_CUDA(cudaMemcpyAsync(ctx->scratch, G.data(), G.size() * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
    XR_CUDA(cudaStreamSynchronize(ctx->stream));      // G lives on this stack frame
    p.G = static_cast<const long long*>(ctx->scratch);
    const int64_t total = p.T * p.n_bra * ((p.n_ket + JT - 1) / JT);
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)ctx->sm_count * 64;
    if (blocks > cap) blocks = cap;
    density_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(p);
    XR_CUDA(cudaGetLastError());
    ctx->launches++;
    return XR_OK;
}

// ------------------------------------------------------------------------------------------------------------------------
// The eight entry points of general-XRCC/density_tensors.c with the reference's own C signature (host pointers; PyInt =
// BigInt = int64, Double = double), so that build_density_tensors.py:23 `import_C("density_tensors")` can bind this
// library instead.  Each call uploads the two CI-vector blocks and the ket configurations (as occupation masks), runs
// density_kernel, and adds into `storage` exactly as the reference's `tensor[index] +=` does.  `combinatorics` is accepted
// for signature compatibility (the ranking table is rebuilt from n_orbs/n_core); n_threads is ignored, as in the reference.
namespace {

std::mutex g_density_mutex;
xr_ctx* g_density_ctx = nullptr;

void legacy_density(const char* ops, double* storage, int64_t bra, int64_t ket, const int64_t* n_elec, const int64_t* n_states,
                    double* const* z_list, const int64_t* n_configs, int64_t* const* configs, int64_t n_orbs, int64_t n_core) {
    std::lock_guard<std::mutex> lock(g_density_mutex);
    if (!g_density_ctx && xr_ctx_create(0, nullptr, 1, &g_density_ctx) != XR_OK) return;
    xr_ctx* ctx = g_density_ctx;
    if (!storage || !n_elec || !n_states || !z_list || !n_configs || !configs) {
        xr_set_error("%s_tensor: null argument", ops);
        return;
    }
    int k = 0;
    while (ops[k]) ++k;
    int64_t T = 1;
    for (int o = 0; o < k; ++o) T *= 2 * n_orbs;
    const int64_t nb = n_states[bra], nk = n_states[ket], cb = n_configs[bra], ck = n_configs[ket], ne = n_elec[ket];
    if (nb <= 0 || nk <= 0 || ck <= 0) return;
    std::vector<unsigned long long> masks((size_t)ck, 0ull);
    for (int64_t Q = 0; Q < ck; ++Q)
        for (int64_t e = 0; e < ne; ++e) {
            const int64_t orb = configs[ket][Q * ne + e];
            if (orb < 0 || orb >= 64) {
                xr_set_error("%s_tensor: orbital index %lld outside 0..63", ops, (long long)orb);
                return;
            }
            masks[(size_t)Q] |= 1ull << orb;
        }
    const size_t bytes_rho = (size_t)(nb * nk * T) * sizeof(double), bytes_zb = (size_t)(nb * cb) * sizeof(double),
                 bytes_zk = (size_t)(nk * ck) * sizeof(double), bytes_m = (size_t)ck * sizeof(unsigned long long);
    double *d_rho = nullptr, *d_zb = nullptr, *d_zk = nullptr;
    unsigned long long* d_m = nullptr;
    bool ok = cudaSetDevice(ctx->device) == cudaSuccess && cudaMalloc(&d_rho, bytes_rho) == cudaSuccess &&
              cudaMalloc(&d_zb, bytes_zb) == cudaSuccess && cudaMalloc(&d_zk, bytes_zk) == cudaSuccess && cudaMalloc(&d_m, bytes_m) == cudaSuccess;
    ok = ok && cudaMemcpyAsync(d_rho, storage, bytes_rho, cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess &&
         cudaMemcpyAsync(d_zb, z_list[bra], bytes_zb, cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess &&
         cudaMemcpyAsync(d_zk, z_list[ket], bytes_zk, cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess &&
         cudaMemcpyAsync(d_m, masks.data(), bytes_m, cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess;
    if (!ok) xr_set_error("%s_tensor: device allocation or upload failed: %s", ops, cudaGetErrorString(cudaGetLastError()));
    if (ok && xr_density_tensor(ctx, ops, d_rho, nb, nk, d_zb, cb, d_zk, ck, reinterpret_cast<const uint64_t*>(d_m), n_elec[bra], ne,
                                n_orbs, n_core) == XR_OK) {
        if (cudaMemcpyAsync(storage, d_rho, bytes_rho, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess)
            xr_set_error("%s_tensor: download failed: %s", ops, cudaGetErrorString(cudaGetLastError()));
    }
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_rho); cudaFree(d_zb); cudaFree(d_zk); cudaFree(d_m);
}

}  // namespace

#define XR_LEGACY_DENSITY(name, ops)                                                                                          \
    extern "C" void name(double* storage, int64_t bra_chg_idx, int64_t ket_chg_idx, int64_t* n_elec, int64_t* n_states,        \
                         double** z_list, int64_t* n_configs, int64_t** configs, int64_t n_orbs, int64_t n_core,               \
                         int64_t** combinatorics, int64_t n_threads) {                                                         \
        (void)combinatorics; (void)n_threads;                                                                                  \
        legacy_density(ops, storage, bra_chg_idx, ket_chg_idx, n_elec, n_states, z_list, n_configs, configs, n_orbs, n_core);  \
    }

XR_LEGACY_DENSITY(a_tensor, "a")
XR_LEGACY_DENSITY(c_tensor, "c")
XR_LEGACY_DENSITY(aa_tensor, "aa")
XR_LEGACY_DENSITY(cc_tensor, "cc")
XR_LEGACY_DENSITY(ca_tensor, "ca")
XR_LEGACY_DENSITY(caa_tensor, "caa")
XR_LEGACY_DENSITY(cca_tensor, "cca")
XR_LEGACY_DENSITY(ccaa_tensor, "ccaa")
