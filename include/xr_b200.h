/* xr_b200.h -- C ABI of libxr_b200.so, the B200 (sm_100a) implementation of the
 * excitonic-renormalization Hamiltonian-build hot path of adutoi/QodeApplications.
 *
 * Two groups of entry points:
 *
 *  (A) The eleven legacy scalar symbols of the reference's general-XRCC/H_contractions.c,
 *      with the identical C ABI the reference binds through qode.util.PyC.import_C
 *      (general-XRCC/build_H.py:20-29, build_density_tensors.py:23-24,132):
 *      PyInt = int64_t sizes, Double* = borrowed C-contiguous HOST buffers, PyFloat result
 *      returned by value, no error channel.  Here each call copies its operands to the
 *      GPU, runs one reduction kernel and returns the scalar -- a drop-in for the
 *      per-element call pattern (and the parity surface for it), not the fast path.
 *      On a CUDA failure they return NaN and xr_last_error() says why.
 *
 *  (B) Block-level entry points (xr_*): what replaces the reference's per-element Python
 *      loops (general-XRCC/test_H.py:90-142) and per-diagram tensornet einsums
 *      (hermitian-XRCC/diagrams/S[TUV]_?mer_?.py via XRbase/XR_tensor.py:57).  They work on DEVICE
 *      pointers, are asynchronous on the context's stream, and return 0 on success or a
 *      negative xr_status (message from xr_last_error()).  One context per host thread;
 *      a context must be created in the process that uses it (CUDA does not survive the
 *      fork that general-XRCC/test_H.py:131 performs).
 *
 * All data are FP64, all offsets int64, nothing is complex.  No torch types appear here.
 */
#ifndef XR_B200_H
#define XR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int64_t PyInt;      /* Qode's PyC_types.h names, as used by H_contractions.c */
typedef double  PyFloat;
typedef double  Double;

/* ---------------------------------------------------------------- (A) legacy scalar ABI */
/* replaces general-XRCC/H_contractions.c:22  */ PyFloat monomer(PyInt n_orb, Double* Rca, Double* Rccaa, Double* h, Double* V);
/* replaces general-XRCC/H_contractions.c:48  */ PyFloat monomer_1e(PyInt n_orb, Double* Rca, Double* h);
/* replaces general-XRCC/H_contractions.c:61  */ PyFloat monomer_2e(PyInt n_orb, Double* Rccaa, Double* V);
/* replaces general-XRCC/H_contractions.c:80  */ PyFloat monomer_extPot(PyInt n_orb, Double* Rca, Double* h);
/* replaces general-XRCC/H_contractions.c:95  */ PyFloat dimer_2min2pls(PyInt n_orb1, PyInt n_orb2, Double* Rcc1, Double* Raa2, Double* V);
/* replaces general-XRCC/H_contractions.c:116 */ PyFloat dimer_1min1pls_1e(PyInt n_orb1, PyInt n_orb2, Double* Rc1, Double* Ra2, Double* h);
/* replaces general-XRCC/H_contractions.c:129 */ PyFloat dimer_1min1pls_2e(PyInt n_orb1, PyInt n_orb2, Double* Rc1, Double* Rcca1, Double* Ra2, Double* Rcaa2, Double* V1112, Double* V1222);
/* replaces general-XRCC/H_contractions.c:163 */ PyFloat dimer_ExEx(PyInt n_orb1, PyInt n_orb2, Double* Rca1, Double* Rca2, Double* V);
/* replaces general-XRCC/H_contractions.c:184 */ PyFloat trimer_2min1pls1pls(PyInt n_orb1, PyInt n_orb2, PyInt n_orb3, Double* Rcc1, Double* Ra2, Double* Ra3, Double* V);
/* replaces general-XRCC/H_contractions.c:208 */ PyFloat trimer_2pls1min1min(PyInt n_orb1, PyInt n_orb2, PyInt n_orb3, Double* Raa1, Double* Rc2, Double* Rc3, Double* V);
/* replaces general-XRCC/H_contractions.c:232 */ PyFloat trimer_Ex1min1pls(PyInt n_orb1, PyInt n_orb2, PyInt n_orb3, Double* Rca1, Double* Rc2, Double* Ra3, Double* V);

/* (A') The eight entry points of the reference's general-XRCC/density_tensors.c (the step before the H build, bound by
 * build_density_tensors.py:23 and called at :85-153), identical C signature: HOST pointers, z_list / configs /
 * combinatorics are arrays of per-charge-sector pointers, `storage` [n_states[bra]*n_states[ket]*(2 n_orbs)^k] is added
 * into.  `combinatorics` and `n_threads` are accepted and unused.  The reference's signature has no error channel: on any
 * failure (no sm_100 device, 2*n_orbs > 64, CUDA errors) the message is printed to stderr, xr_last_error() is set and the
 * requested block of `storage` is filled with NaN -- never left as the zeros the caller allocated. */
typedef int64_t BigInt;
#define XR_DENSITY_ARGS Double storage[], PyInt bra_chg_idx, PyInt ket_chg_idx, BigInt n_elec[], BigInt n_states[], \
                        Double* z_list[], BigInt n_configs[], BigInt* configs[], PyInt n_orbs, PyInt n_core,         \
                        BigInt* combinatorics[], PyInt n_threads
/* replaces general-XRCC/density_tensors.c:162 */ void a_tensor(XR_DENSITY_ARGS);
/* replaces general-XRCC/density_tensors.c:199 */ void c_tensor(XR_DENSITY_ARGS);
/* replaces general-XRCC/density_tensors.c:236 */ void ca_tensor(XR_DENSITY_ARGS);
/* replaces general-XRCC/density_tensors.c:283 */ void aa_tensor(XR_DENSITY_ARGS);
/* replaces general-XRCC/density_tensors.c:330 */ void cc_tensor(XR_DENSITY_ARGS);
/* replaces general-XRCC/density_tensors.c:377 */ void caa_tensor(XR_DENSITY_ARGS);
/* replaces general-XRCC/density_tensors.c:437 */ void cca_tensor(XR_DENSITY_ARGS);
/* replaces general-XRCC/density_tensors.c:493 */ void ccaa_tensor(XR_DENSITY_ARGS);

/* ------------------------------------------------------------------ (B) block-level ABI */
typedef struct xr_ctx xr_ctx;

enum xr_status {
    XR_OK = 0,
    XR_ERR_CUDA = -1,         /* a CUDA runtime call failed */
    XR_ERR_ARG = -2,          /* bad argument (null pointer, unsupported size, misalignment) */
    XR_ERR_NO_DEVICE = -3,    /* no usable sm_100 device: there is NO CPU fallback */
    XR_ERR_UNSUPPORTED = -4
};

/* Conventions of every block-level call below: returns an xr_status; asynchronous on the context's stream; a call whose
 * output has a zero extent (an empty charge sector: M, N, rows, count, Pa.. = 0, or an empty a range) is a no-op that
 * returns XR_OK without launching anything, and its buffers may then be NULL. */

/* Thread-local message of the last failure in the calling thread. */
const char* xr_last_error(void);

/* Library identification: "xr_b200 <version> sm_100a". */
const char* xr_version(void);

/* Create a context on `device`.  With own_stream != 0 the context creates (and later destroys) its
 * own non-blocking stream and `stream` is ignored; otherwise it borrows `stream`, a cudaStream_t
 * (e.g. torch's current stream; NULL is the CUDA default stream). */
int xr_ctx_create(int device, void* stream, int own_stream, xr_ctx** out);
int xr_ctx_destroy(xr_ctx* ctx);
int xr_ctx_set_stream(xr_ctx* ctx, void* stream);
int xr_sync(xr_ctx* ctx);
/* Number of xr kernels launched through this context since creation (bench.py's gpu_launches). */
int xr_launch_count(xr_ctx* ctx, int64_t* count);
int xr_device_info(xr_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, size_t* free_bytes, size_t* total_bytes);

/* Sustained FP64 tensor-pipe rate of this device, measured now: independent DMMA.8x8x4 chains from 16 warps per SM for
 * about `seconds` (0 < seconds <= 10), timed with CUDA events on the context's stream; synchronous.  This is the roofline
 * denominator bench.py quotes for the tensor-bound kernels (tcgen05 has no f64 kind; MEASURED_PEAKS.json has no FP64 entry). */
int xr_probe_fp64(xr_ctx* ctx, double seconds, double* dmma_tflops);

/* Raw device memory for hosts that do not bring their own allocator. */
int xr_malloc(xr_ctx* ctx, size_t bytes, void** dptr);
int xr_free(xr_ctx* ctx, void* dptr);
int xr_memset_zero(xr_ctx* ctx, void* dptr, size_t bytes);
int xr_upload(xr_ctx* ctx, void* dst_device, const void* src_host, size_t bytes);      /* async on ctx stream */
int xr_download(xr_ctx* ctx, void* dst_host, const void* src_device, size_t bytes);    /* async on ctx stream */

/* The contraction engine.  Every pairwise rho x integral or factor x factor contraction of the
 * hot path is one call of this (replaces tensornet/opt_einsum/BLAS below XRbase/XR_tensor.py:57
 * and the n^4 loops of H_contractions.c):
 *
 *     C[ offM(m) + offN(n) ]  (=  or  +=)   alpha * sum_{k<K} A[m*lda + k] * B[n*ldb + k]
 *
 * with offM(m) = offM[m] if offM else m*ldc, offN(n) = offN[n] if offN else n.  The offset
 * tables (device int64) let the epilogue write straight into the final Hamiltonian layout
 * ([i0,i1,j0,j1] blocks, transposed permutations, charge-blocked or state_indices ordering)
 * so no transpose/packing pass exists.  FP64 DMMA (mma.sync m8n8k4); operand tiles reach shared
 * memory by tensor-map TMA (cp.async.bulk.tensor.2d, 128-byte swizzle, mbarrier ring) in persistent
 * CTAs, K/M/N tails zero-filled by the TMA unit; products with few output tiles and a long K are split
 * over K with a fixed-order second pass.  Requirements: A, B, C device pointers to doubles.  Rows must
 * be 16-byte aligned (even lda/ldb, 16-byte-aligned bases) for the TMA path; anything else takes a
 * cp.async-staged kernel with 8-byte copies (same results). */
int xr_gemm_scatter(xr_ctx* ctx, int64_t M, int64_t N, int64_t K, double alpha,
                    const double* A, int64_t lda, const double* B, int64_t ldb,
                    double* C, const int64_t* offM, int64_t ldc, const int64_t* offN, int accumulate);

/* The same product for a tall-skinny output with a long contraction (N <= 32; the rho x integral precontractions of
 * hermitian-XRCC, hermitian-XRCC/precontract.py:26-94), with A addressed where it lies instead of as a matrix:
 *
 *     row   m = r1*E2 + r2     at element offset  r1*s1 + r2*s2          (r1 < E1, r2 < E2)
 *     index k = k1*EK2 + k2    at element offset  k1*sk1 + k2            (k1 < EK1, k2 < EK2: the contiguous index)
 *     C[ offM(m) + offN(n) ]  (= or +=)  alpha * sum_k A[row(m) + col(k)] * B[n*ldb + k]
 *
 * so a density rho[ij, a, b, c, d, e] contracted over (a,b,d,e) is E1 = #ij, E2 = #c, EK1 = #(a,b), EK2 = #(d,e) -- no
 * re-ordering copy of the density.  E2 = EK1 = 1 is a plain matrix with lda = s1.  The contraction is always split over
 * the whole GPU and summed in a fixed order (bit-reproducible); the roofline is one read of A from HBM.  Requirements:
 * 16-byte aligned A and B, even s1 / s2 / sk1 / ldb (else XR_ERR_UNSUPPORTED: use xr_permute_copy + xr_gemm_scatter). */
int xr_gemm_stream(xr_ctx* ctx, int64_t E1, int64_t s1, int64_t E2, int64_t s2, int64_t EK1, int64_t sk1, int64_t EK2,
                   int64_t N, double alpha, const double* A, const double* B, int64_t ldb,
                   double* C, const int64_t* offM, int64_t ldc, const int64_t* offN, int accumulate);

/* The same product handed to the streaming consumer instead of memory:
 *     moments[0] += alpha * sum_{m,n} C[m,n],   moments[1] += alpha^2 * sum_{m,n} C[m,n]^2,   C = A . B^T
 * (device doubles, caller zeroes; bit-reproducible).  For dimer blocks that cannot be stored (the 1e12-element H2 of
 * the 1000-states/fragment stress configuration).  Operands must be 16-byte aligned with even leading dimensions. */
int xr_gemm_reduce(xr_ctx* ctx, int64_t M, int64_t N, int64_t K, double alpha,
                   const double* A, int64_t lda, const double* B, int64_t ldb, double* moments);

/* dst[r*dst_ld + c] = alpha * src[r*src_ld + c]  (rows x cols, device to device).  Used to lay
 * densities into zero-padded, sign-folded factor matrices. */
int xr_copy2d_scaled(xr_ctx* ctx, double* dst, int64_t dst_ld, const double* src, int64_t src_ld,
                     int64_t rows, int64_t cols, double alpha);

/* dst (contiguous, row-major over `shape`) <- alpha * src viewed with arbitrary element strides:
 *   dst[i0,...,i_{nd-1}] = alpha * src[ sum_d i_d * src_strides[d] ]      (nd <= 12; host arrays of length nd)
 * The index-permutation step of a tensor contraction (hermitian-XRCC operands whose contracted
 * indices are not trailing, e.g. "ccaa0pXsr_Vp1rs" in hermitian-XRCC/diagrams/SV_2mer_1.py:30-31). */
int xr_permute_copy(xr_ctx* ctx, double* dst, const double* src, int nd, const int64_t* shape,
                    const int64_t* src_strides, double alpha);

/* C[idx[t]] (= or +=) value for t < count  (Kronecker-delta terms; idx is a device int64 table). */
int xr_scatter_const(xr_ctx* ctx, double* C, const int64_t* idx, int64_t count, double value, int accumulate);

/* Expansion of a fragment block into the supersystem matrix (the consumer of the build): replaces the interpreter loops
 * of general-XRCC/hamiltonian.py:21-84 (braket_loops; hermitian-XRCC/hamiltonian.py is the two-body subset) and the
 * H1 (x) 1 einsums of hermitian-XRCC/mains/workflow.py:216-226.
 *
 *     H[offR[r] + offC[c] + offS[s]] += alpha * src[r*ld + c]        r < R, c < Cn, s < S
 *
 * offR/offC place the bra/ket states of the block's k fragments, offS the common state of the spectator fragments
 * (device int64 tables; offS may be NULL when S == 1).  With min_transitions > 0 an element is skipped unless at least
 * that many of the k sub-fragments change state between r and c, the digits of r and c over dims_sub[0..k) (a HOST
 * array, k <= 4, last fastest): hamiltonian.py:44-56 reads trimer couplings only where >= 2 fragments change. */
int xr_embed_add(xr_ctx* ctx, double* H, const double* src, int64_t ld, int64_t R, int64_t Cn, int64_t S,
                 const int64_t* offR, const int64_t* offC, const int64_t* offS, int k, const int64_t* dims_sub,
                 int min_transitions, double alpha);

/* out[i*ldo + j] = C0[i*ldc0 + j] (delta_ij when C0 is NULL) + sign * sum_k A[i*lda + k] * B[k*ldb + j],  sign = +-1,
 * with every product and the running sum carried in double-double (~106 bits).  The two halves of the Newton polish
 * X <- X + X (I - M X) of the overlap inverse: hermitian-XRCC/get_xr_result.py:165,202,246,285 call
 * qode.math.precise_numpy_inverse (an extended-precision refinement of numpy's inverse). out must not alias A or B. */
int xr_gemm_dd(xr_ctx* ctx, int64_t M, int64_t N, int64_t K, const double* A, int64_t lda, const double* B, int64_t ldb,
               const double* C0, int64_t ldc0, double sign, double* out, int64_t ldo);

/* Transition-density tensor of one operator string between two charge sectors (general-XRCC/density_tensors.c:142-556):
 *
 *     rho[(I*n_ket_states + J)*dim^k + index] (+)= sum_Q parity * z_bra[I*n_configs_bra + P] * z_ket[J*n_configs_ket + Q]
 *
 * for every ket configuration Q and orbital indices (i_0..i_{k-1}) (index row-major, dim = 2*n_orbs <= 64) for which
 * |P> = parity * op_0(i_0) ... op_{k-1}(i_{k-1}) |Q> keeps the n_core core orbitals of both spins occupied; ops is a
 * string of 'c'/'a' (k <= 4: "a","c","aa","cc","ca","caa","cca","ccaa").  ket_masks[Q] (device) is the occupation bit
 * mask of ket configuration Q; the bra coefficients must span ALL C(2(n_orbs-n_core), n_elec_bra-2 n_core) valence
 * configurations in find_config_index order (density_tensors.c:29-64).  accumulate = 0 overwrites rho (no need to clear
 * it), 1 adds to what is there (the reference's `+=`).  Gather formulation, no atomics: the summation
 * order over Q -- and therefore every bit of the result -- is the reference's.  All pointers are device pointers. */
int xr_density_tensor(xr_ctx* ctx, const char* ops, double* rho, int64_t n_bra_states, int64_t n_ket_states,
                      const double* z_bra, int64_t n_configs_bra, const double* z_ket, int64_t n_configs_ket,
                      const uint64_t* ket_masks, int64_t n_elec_bra, int64_t n_elec_ket, int64_t n_orbs, int64_t n_core,
                      int accumulate);

/* The same tensor contracted on the fly with weights[dim^k] (device), never stored:
 *     out[I*n_ket_states + J] (+)= sum_index weights[index] * rho[I,J,index]
 * general-XRCC/build_density_tensors.py:125-133 forms every ccaa tensor only to reduce it at once with the two-electron
 * integrals (monomer_2e: weights[p,q,r,s] = V[p,q,s,r]); fused, the largest tensor of the density build never touches HBM.
 * Fixed-order reductions (bit-reproducible); the summation order differs from monomer_2e's, the values agree to rounding. */
int xr_density_contracted(xr_ctx* ctx, const char* ops, double* out, const double* weights, int64_t n_bra_states,
                          int64_t n_ket_states, const double* z_bra, int64_t n_configs_bra, const double* z_ket,
                          int64_t n_configs_ket, const uint64_t* ket_masks, int64_t n_elec_bra, int64_t n_elec_ket, int64_t n_orbs,
                          int64_t n_core, int accumulate);

/* Streamed three-factor contraction, the trimer classes of general-XRCC/build_H.py:103-188
 * after the rho x V precontraction:
 *
 *     T[a,b,c] = alpha * sum_{r,s<n} W[a*ldw + r*n + s] * beta[b*ldbeta + r] * gamma[c*ldgamma + s]
 *
 * for a in [a_begin, a_end), b < Pb, c < Pc.  The Pa*Pb*Pc elements are formed tile by tile in
 * registers by FP64 DMMA and handed to a consumer, because at the benchmark sizes they cannot
 * be stored (1e13 elements per trimer):
 *   XR_TRIMER_REDUCE      moments[0] += sum T, moments[1] += sum T^2   (device doubles, caller zeroes).  The second moment
 *                         is accumulated element by element from the streamed tiles; the first is linear in T and is taken
 *                         from the factor sums (sum_a sum_rs W[a,rs] (sum_b beta[b,r]) (sum_c gamma[c,s])), same value to rounding
 *   XR_TRIMER_MATERIALIZE C[offA[a] + offB[b] + offC[c]] = T[a,b,c]    (device int64 tables)
 * n <= 48.
 *
 * Tile-consumer contract.  A consumer sees every 128 x 128 tile of T once, in DMMA accumulator registers: a lane holds
 * T[row 32*wm + 8i + g][column 64*wn + 8j + 2t + e] (i < 4, j < 8, e < 2) of the work item (8 values of a x 16 of b) and
 * gamma tile (128 values of c) being streamed; it may fold its elements into per-thread state (REDUCE), store them
 * (MATERIALIZE), or emit a subset (xr_trimer_threshold, xr_trimer_sample below).  It must not block other warps: a tile's
 * shared-memory slot is released before the consumer runs.  The reference's consumer of H3 is
 * general-XRCC/hamiltonian.py:44-56 (excitonic.fci reads trimer couplings where >= 2 fragments change state). */
enum xr_trimer_mode { XR_TRIMER_REDUCE = 0, XR_TRIMER_MATERIALIZE = 1 };
int xr_trimer_stream(xr_ctx* ctx, int n, int64_t Pa, int64_t Pb, int64_t Pc, double alpha,
                     const double* W, int64_t ldw, const double* beta, int64_t ldbeta,
                     const double* gamma, int64_t ldgamma, int64_t a_begin, int64_t a_end, int mode,
                     double* moments, double* C, const int64_t* offA, const int64_t* offB, const int64_t* offC);

/* Compaction consumer of the same stream (screened, sparse H3): every element with |T[a,b,c]| > tau (alpha included) is
 * appended to a COO list as (idx_out[s], val_out[s]) = (offA[a] + offB[b] + offC[c], T[a,b,c]).  *count (a DEVICE int64,
 * zeroed by the call) receives the number of such elements; when it exceeds `capacity` only `capacity` of them were stored
 * and the caller re-runs with a longer list.  The CONTENT of the list is reproducible, its order is not (one atomic
 * reservation per warp and tile): sort by idx for a canonical form. */
int xr_trimer_threshold(xr_ctx* ctx, int n, int64_t Pa, int64_t Pb, int64_t Pc, double alpha,
                        const double* W, int64_t ldw, const double* beta, int64_t ldbeta,
                        const double* gamma, int64_t ldgamma, int64_t a_begin, int64_t a_end, double tau,
                        const int64_t* offA, const int64_t* offB, const int64_t* offC,
                        int64_t capacity, int64_t* idx_out, double* val_out, int64_t* count);

/* Sampled-element consumer: out[s] (device) = T[abc[3s], abc[3s+1], abc[3s+2]] for a caller-given HOST list of `count`
 * (a, b, c) triples.  Only the work items that hold a requested element are streamed, through the same tile code as every
 * other consumer -- an element-level view into blocks that cannot be stored (the parity tests compare it with the
 * reference's trimer_* C functions at the benchmark size). */
int xr_trimer_sample(xr_ctx* ctx, int n, int64_t Pa, int64_t Pb, int64_t Pc, double alpha,
                     const double* W, int64_t ldw, const double* beta, int64_t ldbeta,
                     const double* gamma, int64_t ldgamma, int64_t count, const int64_t* abc_host, double* out);

#ifdef __cplusplus
}
#endif
#endif /* XR_B200_H */
