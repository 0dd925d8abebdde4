// Generic (any angular momentum) kernels: one thread per contracted *function* quartet.
//
// This is the device counterpart of the reference's own evaluation granularity -- one
// primitive Cartesian component quartet at a time (computePGTOrbTwoBodyRepulsion!,
// src/Integration/Engines/GaussianOrbitals.jl:594-663) contracted by four nested loops
// (getOrbLayoutIntegralCore!, src/Integration/Framework.jl:526-554).  It serves
//   * qbx_eri_quartets  (elecRepulsion on arbitrary functions, golden vectors up to l = 10),
//   * bases that do not factor into s/p/d shells (the class kernels' eligibility guard),
//   * the dense-tensor mode of qbx_eri_store (small N) and its getGcore kernel,
//   * the one-electron matrices (overlap / kinetic / nuclear attraction).
// The s/p/d production path is eri_class.cuh.
//
// Per axis the auxiliary-order vector g[0..nUp] is reduced in place: for descending order n
// a column c[i] = [i,0|0,0]^(n) is pushed from the previous column (vertical recurrence,
// cf. vertTransfer :386-391), and for n <= nUp - ioSum the completed column goes through
// the electron-transfer recurrence (cf. modeTransfer :529-538) and the two horizontal
// recurrences (cf. horiTransfer :394-396) to a single number that overwrites g[n].
#include "qbx_internal.h"

namespace {

struct Prim { double c[3]; double a; int l[3]; };

__device__ inline Prim load_prim(const DevFlat &f, int64_t p)
{
    Prim r;
    r.c[0] = f.cen[3 * p]; r.c[1] = f.cen[3 * p + 1]; r.c[2] = f.cen[3 * p + 2];
    r.a = f.xpn[p];
    r.l[0] = f.ang[3 * p]; r.l[1] = f.ang[3 * p + 1]; r.l[2] = f.ang[3 * p + 2];
    return r;
}

#define GMAX (QBX_GEN_MAXAX + 1)

// column c[0..ioSum] = [i,0|0,0] -> [iL,iR|oL,oR] on one axis
__device__ double axis_finish(const double *c, int iL, int iR, int oL, int oR, double AB, double CD, double b,
                              double d, double zeta, double eta, double (*M)[GMAX], double *h)
{
    const int iSum = iL + iR, oSum = oL + oR, ioSum = iSum + oSum;
    for (int i = 0; i <= ioSum; ++i) M[0][i] = c[i];
    const double i2e = 0.5 / eta, ie = 1.0 / eta, k0 = b * AB + d * CD;
    for (int o = 1; o <= oSum; ++o)
        for (int i = 0; i <= ioSum - o; ++i) {
            double t = -(k0 * M[o - 1][i] + zeta * M[o - 1][i + 1]) * ie;
            if (i > 0) t += i * i2e * M[o - 1][i - 1];
            if (o > 1) t += (o - 1) * i2e * M[o - 2][i];
            M[o][i] = t;
        }
    // horizontal recurrence on electron 2 for every i <= iSum: (oL+k, oR-k) -> (oL, oR)
    for (int i = 0; i <= iSum; ++i) {
        for (int k = 0; k <= oR; ++k) h[k] = M[oL + k][i];
        for (int y = 1; y <= oR; ++y)
            for (int x = 0; x <= oR - y; ++x) h[x] = h[x + 1] + CD * h[x];
        M[0][i] = h[0];
    }
    for (int k = 0; k <= iR; ++k) h[k] = M[0][iL + k];
    for (int y = 1; y <= iR; ++y)
        for (int x = 0; x <= iR - y; ++x) h[x] = h[x + 1] + AB * h[x];
    return h[0];
}

// (p1 p2|p3 p4) over unnormalised primitive Cartesian Gaussians; NaN if out of range
__device__ double prim_eri_generic(const Prim &A, const Prim &B, const Prim &C, const Prim &D)
{
    int L = 0;
    for (int x = 0; x < 3; ++x) {
        int s = A.l[x] + B.l[x] + C.l[x] + D.l[x];
        if (s > QBX_GEN_MAXAX) return nan("");
        L += s;
    }
    if (L > QBX_GEN_MAXL) return nan("");
    const double zeta = A.a + B.a, eta = C.a + D.a;
    const double rho = zeta * eta / (zeta + eta), fz = rho / zeta;
    double P[3], Q[3], ab2 = 0, cd2 = 0, pq2 = 0;
    for (int x = 0; x < 3; ++x) {
        P[x] = (A.a * A.c[x] + B.a * B.c[x]) / zeta;
        Q[x] = (C.a * C.c[x] + D.a * D.c[x]) / eta;
        const double ab = A.c[x] - B.c[x], cd = C.c[x] - D.c[x], pq = P[x] - Q[x];
        ab2 += ab * ab; cd2 += cd * cd; pq2 += pq * pq;
    }
    const double pref = 34.986836655249725693 /* 2 pi^(5/2) */ * exp(-A.a * B.a / zeta * ab2 - C.a * D.a / eta * cd2) /
                        (zeta * eta * sqrt(zeta + eta));
    double g[QBX_GEN_MAXL + 1], col[GMAX], h[GMAX];
    double M[GMAX][GMAX];
    boys_generic(rho * pq2, L, g);
    int nUp = L;
    for (int x = 0; x < 3; ++x) {
        const int iL = A.l[x], iR = B.l[x], oL = C.l[x], oR = D.l[x];
        const int ioSum = iL + iR + oL + oR;
        if (ioSum == 0) continue;
        const double PA = P[x] - A.c[x], fPQ = fz * (P[x] - Q[x]), i2z = 0.5 / zeta;
        const double AB = A.c[x] - B.c[x], CD = C.c[x] - D.c[x];
        const int nrem = nUp - ioSum;
        for (int n = nUp; n >= 0; --n) {
            const int imax = min(ioSum, nUp - n);
            double o2 = 0.0, o1 = col[0], n2 = 0.0, n1 = g[n];
            col[0] = n1;
            for (int i = 1; i <= imax; ++i) {
                const double oi = col[i];
                double nw = PA * n1 - fPQ * o1;
                if (i > 1) nw += (i - 1) * i2z * (n2 - fz * o2);
                col[i] = nw;
                o2 = o1; o1 = oi; n2 = n1; n1 = nw;
            }
            if (n <= nrem) g[n] = axis_finish(col, iL, iR, oL, oR, AB, CD, B.a, D.a, zeta, eta, M, h);
        }
        nUp = nrem;
    }
    return pref * g[0];
}

__device__ inline int fn_l(const DevFlat &f, int64_t fn)
{
    const int64_t p = f.bf_prim[f.bf_off[fn]];
    return f.ang[3 * p] + f.ang[3 * p + 1] + f.ang[3 * p + 2];
}

// The quartet is first brought into the l-canonical orientation (higher-l pair = electron 1,
// higher-l function first in each pair): (ij|kl) is invariant, but the recurrences build all
// angular momentum on the first function's centre and then transfer it to electron 2, which loses
// up to 11 digits when that centre is far from a tight partner and the angular momentum sits on
// the other electron (the reference's index-order evaluation has exactly this problem; DESIGN.md
// section 2).
__device__ double contracted_eri_generic(const DevFlat &f, int64_t i, int64_t j, int64_t k, int64_t l)
{
    {
        int li = fn_l(f, i), lj = fn_l(f, j), lk = fn_l(f, k), ll = fn_l(f, l);
        int64_t t; int tl;
        if (lj > li) { t = i; i = j; j = t; tl = li; li = lj; lj = tl; }
        if (ll > lk) { t = k; k = l; l = t; tl = lk; lk = ll; ll = tl; }
        if (lk + ll > li + lj || (lk + ll == li + lj && lk > li)) { t = i; i = k; k = t; t = j; j = l; l = t; }
    }
    double res = 0.0;
    for (int64_t s = f.bf_off[l]; s < f.bf_off[l + 1]; ++s) {
        const Prim D = load_prim(f, f.bf_prim[s]);
        for (int64_t r = f.bf_off[k]; r < f.bf_off[k + 1]; ++r) {
            const Prim C = load_prim(f, f.bf_prim[r]);
            const double w2 = f.bf_w[r] * f.bf_w[s];
            for (int64_t q = f.bf_off[j]; q < f.bf_off[j + 1]; ++q) {
                const Prim B = load_prim(f, f.bf_prim[q]);
                for (int64_t p = f.bf_off[i]; p < f.bf_off[i + 1]; ++p) {
                    const Prim A = load_prim(f, f.bf_prim[p]);
                    res += prim_eri_generic(A, B, C, D) * (f.bf_w[p] * f.bf_w[q] * w2);
                }
            }
        }
    }
    return res;
}

__global__ void __launch_bounds__(64) k_generic_quartets(DevFlat f, int64_t n, const int64_t *ijkl, double *out)
{
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n) return;
    out[t] = contracted_eri_generic(f, ijkl[4 * t], ijkl[4 * t + 1], ijkl[4 * t + 2], ijkl[4 * t + 3]);
}

__device__ inline void tri_decode(int64_t n, int64_t &i, int64_t &j)
{   // n -> (i <= j) with n = j(j+1)/2 + i   (convertIndex1DtoTri2D, src/Iteration.jl:7-12, 0-based)
    int64_t jj = (int64_t)((sqrt(8.0 * (double)n + 1.0) - 1.0) * 0.5);
    while ((jj + 1) * (jj + 2) / 2 <= n) ++jj;
    while (jj * (jj + 1) / 2 > n) --jj;
    j = jj; i = n - jj * (jj + 1) / 2;
}

// dense N^4 tensor from the M(M+1)/2 unique function quartets and their 8 images
// (getOrbVectorIntegralCore!, Framework.jl:651-665)
__global__ void __launch_bounds__(64) k_generic_tensor(DevFlat f, int64_t nuniq, double *T)
{
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nuniq) return;
    int64_t p, q, i, j, k, l;
    tri_decode(t, p, q);
    tri_decode(p, i, j);
    tri_decode(q, k, l);
    const double v = contracted_eri_generic(f, i, j, k, l);
    const int64_t N = f.nbf;
#define AT(a, b, c, d) T[(a) + N * ((b) + N * ((c) + N * (d)))]
    AT(i, j, k, l) = v; AT(j, i, k, l) = v; AT(i, j, l, k) = v; AT(j, i, l, k) = v;
    AT(k, l, i, j) = v; AT(k, l, j, i) = v; AT(l, k, i, j) = v; AT(l, k, j, i) = v;
#undef AT
}

// ---------------------------------------------------------------------------------------
// one-electron primitives (overlap / kinetic: Obara-Saika 1D tables; nuclear attraction:
// the one-centre version of the axis reduction above).
// Replaces computePGTOrbOverlap! (:149-180), computePGTOrbCoordDiff! (:322-363, degree 2,
// direction -1/2) and computePGTOrbOneBodyRepulsion! (:478-522).
// ---------------------------------------------------------------------------------------
#define OB_MAX 20
// s[i][j] = <x_A^i | x_B^j> 1D without the sqrt(pi/p) exp() prefactor, i <= iL, j <= jR+2
__device__ void ob_axis(double (*s)[OB_MAX + 3], int iL, int jmax, double PA, double PB, double i2p)
{
    s[0][0] = 1.0;
    for (int i = 0; i < iL; ++i) s[i + 1][0] = PA * s[i][0] + (i > 0 ? i * i2p * s[i - 1][0] : 0.0);
    for (int j = 0; j < jmax; ++j)
        for (int i = 0; i <= iL; ++i) {
            double t = PB * s[i][j];
            if (j > 0) t += j * i2p * s[i][j - 1];
            if (i > 0) t += i * i2p * s[i - 1][j];
            s[i][j + 1] = t;
        }
}

__device__ double prim_overlap_kinetic(const Prim &A, const Prim &B, int kinetic)
{
    const double p = A.a + B.a, i2p = 0.5 / p;
    double ab2 = 0.0, ov[3], kd[3];
    double s[OB_MAX + 1][OB_MAX + 3];
    for (int x = 0; x < 3; ++x) {
        if (A.l[x] > OB_MAX || B.l[x] > OB_MAX) return nan("");
        const double Px = (A.a * A.c[x] + B.a * B.c[x]) / p, ab = A.c[x] - B.c[x];
        ab2 += ab * ab;
        const int j = B.l[x];
        ob_axis(s, A.l[x], j + 2, Px - A.c[x], Px - B.c[x], i2p);
        ov[x] = s[A.l[x]][j];
        // d^2/dx^2 acting on the right function
        kd[x] = 4.0 * B.a * B.a * s[A.l[x]][j + 2] - 2.0 * B.a * (2 * j + 1) * s[A.l[x]][j] +
                (j > 1 ? j * (j - 1) * s[A.l[x]][j - 2] : 0.0);
    }
    const double pre = pow(3.14159265358979323846 / p, 1.5) * exp(-A.a * B.a / p * ab2);
    if (!kinetic) return pre * ov[0] * ov[1] * ov[2];
    return -0.5 * pre * (kd[0] * ov[1] * ov[2] + ov[0] * kd[1] * ov[2] + ov[0] * ov[1] * kd[2]);
}

__device__ double prim_nuclear(const Prim &A, const Prim &B, const double *Cn)
{
    int L = 0;
    for (int x = 0; x < 3; ++x) {
        if (A.l[x] + B.l[x] > QBX_GEN_MAXAX) return nan("");
        L += A.l[x] + B.l[x];
    }
    if (L > QBX_GEN_MAXL) return nan("");
    const double p = A.a + B.a;
    double P[3], ab2 = 0, pc2 = 0;
    for (int x = 0; x < 3; ++x) {
        P[x] = (A.a * A.c[x] + B.a * B.c[x]) / p;
        const double ab = A.c[x] - B.c[x], pc = P[x] - Cn[x];
        ab2 += ab * ab; pc2 += pc * pc;
    }
    const double pref = 2.0 * 3.14159265358979323846 / p * exp(-A.a * B.a / p * ab2);
    double g[QBX_GEN_MAXL + 1], col[GMAX], h[GMAX];
    boys_generic(p * pc2, L, g);
    int nUp = L;
    for (int x = 0; x < 3; ++x) {
        const int iL = A.l[x], iR = B.l[x], iSum = iL + iR;
        if (iSum == 0) continue;
        const double PA = P[x] - A.c[x], PC = P[x] - Cn[x], i2p = 0.5 / p, AB = A.c[x] - B.c[x];
        const int nrem = nUp - iSum;
        for (int n = nUp; n >= 0; --n) {
            const int imax = min(iSum, nUp - n);
            double o2 = 0.0, o1 = col[0], n2 = 0.0, n1 = g[n];
            col[0] = n1;
            for (int i = 1; i <= imax; ++i) {
                const double oi = col[i];
                double nw = PA * n1 - PC * o1;
                if (i > 1) nw += (i - 1) * i2p * (n2 - o2);
                col[i] = nw;
                o2 = o1; o1 = oi; n2 = n1; n1 = nw;
            }
            if (n <= nrem) {
                for (int k = 0; k <= iR; ++k) h[k] = col[iL + k];
                for (int y = 1; y <= iR; ++y)
                    for (int xx = 0; xx <= iR - y; ++xx) h[xx] = h[xx + 1] + AB * h[xx];
                g[n] = h[0];
            }
        }
        nUp = nrem;
    }
    return pref * g[0];
}

// one thread per (i <= j) function pair; two-index contraction (Framework.jl:498-523)
__global__ void __launch_bounds__(64) k_one_body(DevFlat f, int kind, int64_t nnuc, const double *Z, const double *R,
                                                 double *out)
{
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t N = f.nbf;
    if (t >= N * (N + 1) / 2) return;
    int64_t i, j;
    tri_decode(t, i, j);
    double res = 0.0;
    for (int64_t q = f.bf_off[j]; q < f.bf_off[j + 1]; ++q) {
        const Prim B = load_prim(f, f.bf_prim[q]);
        for (int64_t p = f.bf_off[i]; p < f.bf_off[i + 1]; ++p) {
            const Prim A = load_prim(f, f.bf_prim[p]);
            double v = 0.0;
            if (kind < 2) v = prim_overlap_kinetic(A, B, kind);
            else {
                for (int64_t c = 0; c < nnuc; ++c) v -= Z[c] * prim_nuclear(A, B, R + 3 * c);
            }
            res += v * f.bf_w[p] * f.bf_w[q];
        }
    }
    out[i + N * j] = res;
    out[j + N * i] = res;
}

// ---------------------------------------------------------------------------------------
// getGcore on a dense tensor (src/HartreeFock.jl:305-319): one block per (mu <= nu),
// G[mu,nu] = sum_{lm,sg} DJ[sg,lm] H[mu,nu,lm,sg] - DK[lm,sg] H[mu,lm,sg,nu].
// Small-N mode only; the production Fock build digests packed shell-quartet blocks.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dense_gcore(int64_t N, const double *H, int nmat, const double *DJ,
                                                     const double *DK, double *G)
{
    int64_t mu, nu;
    tri_decode(blockIdx.x, mu, nu);
    __shared__ double red[256];
    double accJ = 0.0;
    const int64_t N2 = N * N;
    for (int64_t e = threadIdx.x; e < N2; e += blockDim.x) {
        const int64_t lm = e % N, sg = e / N;
        accJ += DJ[sg + N * lm] * H[mu + N * (nu + N * e)];
    }
    for (int m = 0; m < nmat; ++m) {
        double acc = accJ;
        const double *dk = DK + (int64_t)m * N2;
        for (int64_t e = threadIdx.x; e < N2; e += blockDim.x)
            acc -= dk[e] * H[mu + N * (e + N2 * nu)];      // e = lm + N*sg
        red[threadIdx.x] = acc;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
            if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            G[(int64_t)m * N2 + mu + N * nu] = red[0];
            G[(int64_t)m * N2 + nu + N * mu] = red[0];
        }
        __syncthreads();
    }
}

__global__ void k_boys(int64_t n, const double *T, int mmax, int table, BoysTable tb, double *out)
{
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n) return;
    if (table) {
        double F[9];
        boys_table_global<8>(tb, T[t], 1.0, F);
        for (int m = 0; m <= mmax; ++m) out[(int64_t)(mmax + 1) * t + m] = F[m];
    } else {
        double F[129];
        boys_generic(T[t], mmax, F);
        for (int m = 0; m <= mmax; ++m) out[(int64_t)(mmax + 1) * t + m] = F[m];
    }
}

}  // namespace

int qbx_launch_generic_quartets(const DevFlat &f, int64_t n, const int64_t *d_ijkl, double *d_out, cudaStream_t s)
{
    if (n == 0) return QBX_OK;
    k_generic_quartets<<<(unsigned)((n + 63) / 64), 64, 0, s>>>(f, n, d_ijkl, d_out);
    QBX_CUDA(cudaGetLastError());
    return QBX_OK;
}

int qbx_launch_generic_tensor(const DevFlat &f, double *d_tensor, cudaStream_t s)
{
    const int64_t M = f.nbf * (f.nbf + 1) / 2, U = M * (M + 1) / 2;
    if (U == 0) return QBX_OK;
    k_generic_tensor<<<(unsigned)((U + 63) / 64), 64, 0, s>>>(f, U, d_tensor);
    QBX_CUDA(cudaGetLastError());
    return QBX_OK;
}

int qbx_launch_one_body(const DevFlat &f, int kind, int64_t nnuc, const double *dZ, const double *dR, double *d_out,
                        cudaStream_t s)
{
    const int64_t M = f.nbf * (f.nbf + 1) / 2;
    if (M == 0) return QBX_OK;
    k_one_body<<<(unsigned)((M + 63) / 64), 64, 0, s>>>(f, kind, nnuc, dZ, dR, d_out);
    QBX_CUDA(cudaGetLastError());
    return QBX_OK;
}

int qbx_launch_dense_gcore(int64_t n, const double *dH, int nmat, const double *dDJ, const double *dDK, double *dG,
                           cudaStream_t s)
{
    const int64_t M = n * (n + 1) / 2;
    if (M == 0) return QBX_OK;
    k_dense_gcore<<<(unsigned)M, 256, 0, s>>>(n, dH, nmat, dDJ, dDK, dG);
    QBX_CUDA(cudaGetLastError());
    return QBX_OK;
}

int qbx_launch_boys(int64_t n, const double *dT, int mmax, int table, double *d_out, cudaStream_t s)
{
    if (n == 0) return QBX_OK;
    k_boys<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(n, dT, mmax, table, qbx_boys_table(), d_out);
    QBX_CUDA(cudaGetLastError());
    return QBX_OK;
}
