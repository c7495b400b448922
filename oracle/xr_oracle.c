/* Plain-C restatement of the eleven scalar contractions of the reference's
 * general-XRCC/H_contractions.c, same C ABI (int64 sizes, borrowed row-major double
 * buffers, result by value).  TEST INFRASTRUCTURE: the CPU checker and the "port" CPU
 * baseline; never linked or called by the product path.
 *
 * Pinned against the reference's own compiled C (oracle/_ref/libH_contractions_ref.so) in
 * tests/test_oracle.py on random inputs: the two must agree to rounding.
 *
 * Every function is written as "fold the densities into one weight per integral element,
 * then dot with the integral block", which is the same sum the reference accumulates with
 * nested loops (file:line cited per function).
 */
#include <stdint.h>
#include <stddef.h>

typedef int64_t i64;

/* sum_{ab} M[a*nb+b] * W[a*nb+b] */
static double dot2(i64 na, i64 nb, const double *M, const double *W)
{
    double acc = 0.0;
    for (i64 x = 0; x < na * nb; ++x) acc += M[x] * W[x];
    return acc;
}

/* H_contractions.c:48-59 (monomer_1e) and :80-91 (monomer_extPot): sum_pq h[p,q] Rca[p,q] */
double monomer_1e(i64 n, const double *Rca, const double *h)     { return dot2(n, n, h, Rca); }
double monomer_extPot(i64 n, const double *Rca, const double *h) { return dot2(n, n, h, Rca); }

/* H_contractions.c:61-76: sum_pqrs V[p,q,r,s] Rccaa[p,q,s,r] */
double monomer_2e(i64 n, const double *Rccaa, const double *V)
{
    double acc = 0.0;
    for (i64 pq = 0; pq < n * n; ++pq) {
        const double *Vpq = V + pq * n * n, *Rpq = Rccaa + pq * n * n;
        for (i64 r = 0; r < n; ++r)
            for (i64 s = 0; s < n; ++s)
                acc += Vpq[r * n + s] * Rpq[s * n + r];
    }
    return acc;
}

/* H_contractions.c:22-46: two-electron part as monomer_2e, then the one-electron part */
double monomer(i64 n, const double *Rca, const double *Rccaa, const double *h, const double *V)
{
    return monomer_2e(n, Rccaa, V) + dot2(n, n, h, Rca);
}

/* H_contractions.c:95-112: sum V[p1,q1,r2,s2] Rcc1[p1,q1] Raa2[s2,r2] */
double dimer_2min2pls(i64 n1, i64 n2, const double *Rcc1, const double *Raa2, const double *V)
{
    double acc = 0.0;
    for (i64 pq = 0; pq < n1 * n1; ++pq) {
        const double *Vpq = V + pq * n2 * n2;
        double inner = 0.0;
        for (i64 r = 0; r < n2; ++r)
            for (i64 s = 0; s < n2; ++s)
                inner += Vpq[r * n2 + s] * Raa2[s * n2 + r];
        acc += Rcc1[pq] * inner;
    }
    return acc;
}

/* H_contractions.c:116-127: sum h[p1,q2] Rc1[p1] Ra2[q2] */
double dimer_1min1pls_1e(i64 n1, i64 n2, const double *Rc1, const double *Ra2, const double *h)
{
    double acc = 0.0;
    for (i64 p = 0; p < n1; ++p) {
        double inner = 0.0;
        for (i64 q = 0; q < n2; ++q) inner += h[p * n2 + q] * Ra2[q];
        acc += Rc1[p] * inner;
    }
    return acc;
}

/* H_contractions.c:129-159:
 * 2*( sum V1112[p1,q1,r1,s2] Rcca1[q1,p1,r1] Ra2[s2] + sum V1222[p1,q2,r2,s2] Rc1[p1] Rcaa2[q2,s2,r2] ) */
double dimer_1min1pls_2e(i64 n1, i64 n2, const double *Rc1, const double *Rcca1, const double *Ra2,
                         const double *Rcaa2, const double *V1112, const double *V1222)
{
    double first = 0.0, second = 0.0;
    for (i64 p = 0; p < n1; ++p)
        for (i64 q = 0; q < n1; ++q)
            for (i64 r = 0; r < n1; ++r) {
                const double *Vpqr = V1112 + ((p * n1 + q) * n1 + r) * n2;
                double inner = 0.0;
                for (i64 s = 0; s < n2; ++s) inner += Vpqr[s] * Ra2[s];
                first += Rcca1[(q * n1 + p) * n1 + r] * inner;
            }
    for (i64 p = 0; p < n1; ++p) {
        const double *Vp = V1222 + p * n2 * n2 * n2;
        double inner = 0.0;
        for (i64 q = 0; q < n2; ++q)
            for (i64 r = 0; r < n2; ++r)
                for (i64 s = 0; s < n2; ++s)
                    inner += Vp[(q * n2 + r) * n2 + s] * Rcaa2[(q * n2 + s) * n2 + r];
        second += Rc1[p] * inner;
    }
    return 2.0 * (first + second);
}

/* H_contractions.c:163-180: 4 * sum V[p1,q2,r1,s2] Rca1[p1,r1] Rca2[q2,s2] */
double dimer_ExEx(i64 n1, i64 n2, const double *Rca1, const double *Rca2, const double *V)
{
    double acc = 0.0;
    for (i64 p = 0; p < n1; ++p)
        for (i64 q = 0; q < n2; ++q)
            for (i64 r = 0; r < n1; ++r) {
                const double *Vpqr = V + ((p * n2 + q) * n1 + r) * n2;
                double inner = 0.0;
                for (i64 s = 0; s < n2; ++s) inner += Vpqr[s] * Rca2[q * n2 + s];
                acc += Rca1[p * n1 + r] * inner;
            }
    return 4.0 * acc;
}

/* H_contractions.c:184-204: 2 * sum V[p1,q1,r2,s3] Rcc1[q1,p1] Ra2[r2] Ra3[s3] */
double trimer_2min1pls1pls(i64 n1, i64 n2, i64 n3, const double *Rcc1, const double *Ra2, const double *Ra3,
                           const double *V)
{
    double acc = 0.0;
    for (i64 p = 0; p < n1; ++p)
        for (i64 q = 0; q < n1; ++q) {
            const double *Vpq = V + (p * n1 + q) * n2 * n3;
            double inner = 0.0;
            for (i64 r = 0; r < n2; ++r) {
                double row = 0.0;
                for (i64 s = 0; s < n3; ++s) row += Vpq[r * n3 + s] * Ra3[s];
                inner += Ra2[r] * row;
            }
            acc += Rcc1[q * n1 + p] * inner;
        }
    return 2.0 * acc;
}

/* H_contractions.c:208-228: 2 * sum V[r2,s3,p1,q1] Rc2[r2] Rc3[s3] Raa1[q1,p1] */
double trimer_2pls1min1min(i64 n1, i64 n2, i64 n3, const double *Raa1, const double *Rc2, const double *Rc3,
                           const double *V)
{
    double acc = 0.0;
    for (i64 r = 0; r < n2; ++r)
        for (i64 s = 0; s < n3; ++s) {
            const double *Vrs = V + (r * n3 + s) * n1 * n1;
            double inner = 0.0;
            for (i64 p = 0; p < n1; ++p)
                for (i64 q = 0; q < n1; ++q)
                    inner += Vrs[p * n1 + q] * Raa1[q * n1 + p];
            acc += Rc2[r] * Rc3[s] * inner;
        }
    return 2.0 * acc;
}

/* H_contractions.c:232-252: 4 * sum V[p1,r2,q1,s3] Rca1[p1,q1] Rc2[r2] Ra3[s3] */
double trimer_Ex1min1pls(i64 n1, i64 n2, i64 n3, const double *Rca1, const double *Rc2, const double *Ra3,
                         const double *V)
{
    double acc = 0.0;
    for (i64 p = 0; p < n1; ++p)
        for (i64 r = 0; r < n2; ++r)
            for (i64 q = 0; q < n1; ++q) {
                const double *Vprq = V + ((p * n2 + r) * n1 + q) * n3;
                double inner = 0.0;
                for (i64 s = 0; s < n3; ++s) inner += Vprq[s] * Ra3[s];
                acc += Rca1[p * n1 + q] * Rc2[r] * inner;
            }
    return 4.0 * acc;
}
