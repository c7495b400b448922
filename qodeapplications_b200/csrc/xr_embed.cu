// Consumer side of the H build: expansion of fragment blocks (H1, H2, H3) into the supersystem matrix.
// Replaces the interpreter loops of general-XRCC/hamiltonian.py:21-84 (braket_loops) and the H1 (x) 1 einsums of
// hermitian-XRCC/mains/workflow.py:216-226.  HBM-bound: every source element is read once per spectator configuration and
// added to one element of the big matrix; the three offset tables carry all of the layout.
#include "xr_common.cuh"

namespace {

struct EmbedParams {
    int64_t R, Cn, S, ld;
    int k;
    int64_t dims[4];
    int min_transitions;
    double alpha;
};

__global__ void __launch_bounds__(256)
embed_add_kernel(double* __restrict__ H, const double* __restrict__ src, const int64_t* __restrict__ offR,
                 const int64_t* __restrict__ offC, const int64_t* __restrict__ offS, const EmbedParams p) {
    const int64_t total = p.R * p.S * p.Cn;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = t % p.Cn;
        const int64_t rs = t / p.Cn;
        const int64_t s = rs % p.S, r = rs / p.S;
        if (p.min_transitions > 0) {      // count the sub-fragments whose bra and ket state differ (mixed-radix digits of r, c)
            int64_t rr = r, cc = c;
            int changed = 0;
#pragma unroll
            for (int d = 3; d >= 0; --d) {
                if (d < p.k) {
                    changed += (rr % p.dims[d]) != (cc % p.dims[d]);
                    rr /= p.dims[d];
                    cc /= p.dims[d];
                }
            }
            if (changed < p.min_transitions) continue;
        }
        const int64_t at = offR[r] + offC[c] + (offS ? offS[s] : 0);
        H[at] += p.alpha * src[r * p.ld + c];
    }
}

}  // namespace

extern "C" int xr_embed_add(xr_ctx* ctx, double* H, const double* src, int64_t ld, int64_t R, int64_t Cn, int64_t S,
                            const int64_t* offR, const int64_t* offC, const int64_t* offS, int k, const int64_t* dims_sub,
                            int min_transitions, double alpha) {
    XR_REQUIRE(ctx, "xr_embed_add: null ctx");
    if (R <= 0 || Cn <= 0 || S <= 0) return XR_OK;
    XR_REQUIRE(H && src && offR && offC, "xr_embed_add: null pointer");
    XR_REQUIRE(ld >= Cn, "xr_embed_add: ld smaller than the column count");
    XR_REQUIRE(S == 1 || offS, "xr_embed_add: S > 1 needs a spectator offset table");
    XR_REQUIRE(min_transitions == 0 || (k >= 1 && k <= 4 && dims_sub), "xr_embed_add: transition mask needs 1..4 sub-fragment dims");
    EmbedParams p{R, Cn, S, ld, k, {1, 1, 1, 1}, min_transitions, alpha};
    if (min_transitions > 0) {
        int64_t prod = 1;
        for (int d = 0; d < k; ++d) {
            XR_REQUIRE(dims_sub[d] >= 1, "xr_embed_add: non-positive sub-fragment dimension");
            p.dims[d] = dims_sub[d];
            prod *= dims_sub[d];
        }
        XR_REQUIRE(prod == R && prod == Cn, "xr_embed_add: sub-fragment dims do not multiply to the block dimension");
    }
    const int64_t total = R * Cn * S;
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)ctx->sm_count * 32;
    if (blocks > cap) blocks = cap;
    embed_add_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(H, src, offR, offC, offS, p);
    XR_CUDA(cudaGetLastError());
    ctx->launches++;
    return XR_OK;
}
