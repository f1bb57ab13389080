// K2 (tiled version): y = sum_p (MF_p (x) K_p) x for one work unit = (bra-block tile, state tile).
//
// Thread mapping: one thread owns one row m1 of the bra block for TWO states and all (<= 16)
// columns k1 of the tile; accumulators live in registers.  Per block product p
//   * the ket rows the tile needs are staged in shared memory with cp.async (double buffered,
//     coalesced 16-byte copies; row stride padded to an odd number of 16-byte words so that the
//     row-strided reads below are bank-conflict free),
//   * z = sum_j MF_p[m1, j] * X[col_j, k2] is formed on the fly in registers (zero MF diagonals are
//     skipped: after the field contraction most of the (2w+1) diagonals vanish for polarised fields),
//   * acc[k1] += K_p[k1, k2] * z with K_p^T broadcast from shared memory.
// H(t) itself is never materialised: the kernel only sees the MF and K factors.
#pragma once
#include "rmb_internal.h"

namespace rmb {

constexpr int MV2_THREADS = 128;
constexpr int MV2_NCMAX = 16;     // columns (k1) per thread
constexpr int MV2_NDMAX = 5;      // max ELL width handled by the tiled kernel (rank <= 2)
constexpr int MV2_R = 2;          // states per thread

struct Item2D {
    long long bra_off;
    int dk1;
    int r0, nrows;       // rows (m1) of the tile, nrows <= MV2_THREADS
    int c0, nc;          // columns (k1) of the tile, nc <= MV2_NCMAX
    int p_begin, p_end;
    int pairs;           // state pairs per CTA (states per CTA = 2 * pairs)
    int xr_off;          // offset into the per-(item, product) ket row ranges
    int kt_total;        // doubles of K^T staged in shared memory for this item
};

struct XRange { int c_lo, nr; };   // ket rows [c_lo, c_lo + nr) needed by (item, product)
struct Unit2D { int item, s0; };

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ int xrs_of(int dk2) { return dk2 | 1; }

// stage the ket rows of product p for all states of the CTA (coalesced 16-byte cp.async copies)
__device__ __forceinline__ void mv2_stage(double2* xb, const ProdD& pr, const XRange xr, const double2* X,
                                          long long ldx, int s0, int nsl, const int* active, int nstates) {
    const int xrs = xrs_of(pr.dk2);
    const int per_state = xr.nr * pr.dk2;
    const float inv = 1.0f / (float)pr.dk2;
    for (int s = 0; s < nsl; ++s) {
        const int st = s0 + s;
        if (st >= nstates || (active != nullptr && !active[st])) continue;
        const double2* src = X + (long long)st * ldx + pr.ket_off + (long long)xr.c_lo * pr.dk2;
        double2* dst = xb + (long long)s * xr.nr * xrs;
        if (xrs == pr.dk2) {
            for (int e = threadIdx.x; e < per_state; e += MV2_THREADS) cp_async16(dst + e, src + e);
        } else {
            for (int e = threadIdx.x; e < per_state; e += MV2_THREADS) {
                const int rl = __float2int_rz(((float)e + 0.5f) * inv);   // exact for e < 2^21
                cp_async16(dst + e + rl, src + e);                        // rl * xrs + k2 = e + rl
            }
        }
    }
}

template <int NC, int NNZ, bool KC>
__device__ __forceinline__ void mv2_inner(const double2* __restrict__ xa, const double2* __restrict__ xb,
                                          const double2 (&mf)[MV2_NDMAX], const int (&xo)[MV2_NDMAX],
                                          const double* __restrict__ ktp, int dk2, double2 (&accA)[NC],
                                          double2 (&accB)[NC]) {
    constexpr int NCP = NC;   // NC is even (or 1) for real K: rows of K^T are padded by the caller
#pragma unroll 2
    for (int k2 = 0; k2 < dk2; ++k2) {
        double2 zA = make_double2(0.0, 0.0), zB = make_double2(0.0, 0.0);
#pragma unroll
        for (int q = 0; q < NNZ; ++q) {
            const double2 a = xa[xo[q] + k2], b = xb[xo[q] + k2];
            zA.x += mf[q].x * a.x - mf[q].y * a.y;
            zA.y += mf[q].x * a.y + mf[q].y * a.x;
            zB.x += mf[q].x * b.x - mf[q].y * b.y;
            zB.y += mf[q].x * b.y + mf[q].y * b.x;
        }
        if (KC) {
            const double2* krow = reinterpret_cast<const double2*>(ktp) + k2 * NCP;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const double2 kv = krow[c];
                accA[c].x += kv.x * zA.x - kv.y * zA.y;
                accA[c].y += kv.x * zA.y + kv.y * zA.x;
                accB[c].x += kv.x * zB.x - kv.y * zB.y;
                accB[c].y += kv.x * zB.y + kv.y * zB.x;
            }
        } else if (NC == 1) {
            const double kv = ktp[k2];
            accA[0].x += kv * zA.x;
            accA[0].y += kv * zA.y;
            accB[0].x += kv * zB.x;
            accB[0].y += kv * zB.y;
        } else {
            const double2* krow = reinterpret_cast<const double2*>(ktp + k2 * NCP);
#pragma unroll
            for (int c2 = 0; c2 < NC / 2; ++c2) {
                const double2 kv = krow[c2];
                accA[2 * c2].x += kv.x * zA.x;
                accA[2 * c2].y += kv.x * zA.y;
                accB[2 * c2].x += kv.x * zB.x;
                accB[2 * c2].y += kv.x * zB.y;
                accA[2 * c2 + 1].x += kv.y * zA.x;
                accA[2 * c2 + 1].y += kv.y * zA.y;
                accB[2 * c2 + 1].x += kv.y * zB.x;
                accB[2 * c2 + 1].y += kv.y * zB.y;
            }
        }
    }
}

// NC = number of register columns (>= it.nc; 1 or even), surplus columns are zero-padded in K^T
template <int NC, bool KC>
__device__ __forceinline__ void mv2_body(const Item2D& it, const ProdD* __restrict__ prods,
                                         const XRange* __restrict__ xrs_tab, const int* __restrict__ ent_col,
                                         const double2* __restrict__ ent_val, const unsigned* __restrict__ tab_mask,
                                         const double* __restrict__ kpool, const double2* __restrict__ X,
                                         double2* __restrict__ Y, long long ldx, long long ldy, int nstates,
                                         int s0, const int* __restrict__ active, double* kt, double2* xbuf0,
                                         double2* xbuf1) {
    constexpr int KW = KC ? 2 : 1;                       // doubles per K element
    const int nsl = 2 * it.pairs;
    const int pair = threadIdx.x / it.nrows;
    const int rl = threadIdx.x - pair * it.nrows;
    const int sA = s0 + 2 * pair, sB = sA + 1;
    const bool vA = pair < it.pairs && sA < nstates && (active == nullptr || active[sA]);
    const bool vB = pair < it.pairs && sB < nstates && (active == nullptr || active[sB]);
    const bool work = vA || vB;
    const int np = it.p_end - it.p_begin;

    if (np > 0) mv2_stage(xbuf0, prods[it.p_begin], xrs_tab[it.xr_off], X, ldx, s0, nsl, active, nstates);
    cp_async_commit();
    // ---- K^T of every product of the item -> shared memory: kt[p][k2][NC]
    {
        int base = 0;
        for (int p = it.p_begin; p < it.p_end; ++p) {
            const ProdD pr = prods[p];
            const int n = pr.dk2 * NC;
            for (int idx = threadIdx.x; idx < n; idx += MV2_THREADS) {
                const int k2 = idx / NC, c = idx - k2 * NC;
                if (KC) {
                    double2 v = make_double2(0.0, 0.0);
                    if (c < it.nc) v = reinterpret_cast<const double2*>(kpool)[pr.koff + (long long)(it.c0 + c) * pr.dk2 + k2];
                    reinterpret_cast<double2*>(kt + base)[idx] = v;
                } else {
                    kt[base + idx] = (c < it.nc) ? kpool[pr.koff + (long long)(it.c0 + c) * pr.dk2 + k2] : 0.0;
                }
            }
            base += n * KW;
        }
    }
    double2 accA[NC], accB[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) accA[c] = accB[c] = make_double2(0.0, 0.0);

    int ktbase = 0;
    for (int ip = 0; ip < np; ++ip) {
        const ProdD pr = prods[it.p_begin + ip];
        const XRange xr = xrs_tab[it.xr_off + ip];
        double2* xcur = (ip & 1) ? xbuf1 : xbuf0;
        double2* xnext = (ip & 1) ? xbuf0 : xbuf1;
        if (ip + 1 < np)
            mv2_stage(xnext, prods[it.p_begin + ip + 1], xrs_tab[it.xr_off + ip + 1], X, ldx, s0, nsl, active, nstates);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const unsigned mask = tab_mask[pr.tab];     // non-zero diagonals after the field contraction
        if (work && mask != 0u) {
            double2 mf[MV2_NDMAX];
            int xo[MV2_NDMAX];
#pragma unroll
            for (int q = 0; q < MV2_NDMAX; ++q) { mf[q] = make_double2(0.0, 0.0); xo[q] = 0; }
            const long long eb = pr.ent_off + (long long)(it.r0 + rl) * pr.nd;
            const int xrs = xrs_of(pr.dk2);
            int nnz = 0;
#pragma unroll
            for (int j = 0; j < MV2_NDMAX; ++j) {
                if ((mask >> j) & 1u) {
                    const int col = ent_col[eb + j];
                    const double2 v = ent_val[eb + j];
#pragma unroll
                    for (int q = 0; q < MV2_NDMAX; ++q)
                        if (q == nnz && col >= 0) { mf[q] = v; xo[q] = (col - xr.c_lo) * xrs; }
                    ++nnz;
                }
            }
            const double2* xa = xcur + (long long)(2 * pair) * xr.nr * xrs;
            const double2* xb = xa + (long long)xr.nr * xrs;
            const double* ktp = kt + ktbase;
            switch (nnz) {
                case 1: mv2_inner<NC, 1, KC>(xa, xb, mf, xo, ktp, pr.dk2, accA, accB); break;
                case 2: mv2_inner<NC, 2, KC>(xa, xb, mf, xo, ktp, pr.dk2, accA, accB); break;
                case 3: mv2_inner<NC, 3, KC>(xa, xb, mf, xo, ktp, pr.dk2, accA, accB); break;
                case 4: mv2_inner<NC, 4, KC>(xa, xb, mf, xo, ktp, pr.dk2, accA, accB); break;
                default: mv2_inner<NC, 5, KC>(xa, xb, mf, xo, ktp, pr.dk2, accA, accB); break;
            }
        }
        ktbase += pr.dk2 * NC * KW;
        __syncthreads();   // everyone is done with xcur before it is refilled
    }
    cp_async_wait<0>();
    if (vA) {
        double2* y = Y + (long long)sA * ldy + it.bra_off + (long long)(it.r0 + rl) * it.dk1 + it.c0;
#pragma unroll
        for (int c = 0; c < NC; ++c)
            if (c < it.nc) y[c] = accA[c];
    }
    if (vB) {
        double2* y = Y + (long long)sB * ldy + it.bra_off + (long long)(it.r0 + rl) * it.dk1 + it.c0;
#pragma unroll
        for (int c = 0; c < NC; ++c)
            if (c < it.nc) y[c] = accB[c];
    }
}

template <bool KC>
__global__ void __launch_bounds__(MV2_THREADS, 2)
k_matvec_tiled(const Unit2D* __restrict__ units, const Item2D* __restrict__ items,
               const ProdD* __restrict__ prods, const XRange* __restrict__ xrs_tab,
               const int* __restrict__ ent_col, const double2* __restrict__ ent_val,
               const unsigned* __restrict__ tab_mask, const double* __restrict__ kpool,
               const double2* __restrict__ X, double2* __restrict__ Y, long long ldx, long long ldy,
               int nstates, const int* __restrict__ active, int kt_doubles, int xbuf_elems) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* xbuf0 = reinterpret_cast<double2*>(smem_raw);
    double2* xbuf1 = xbuf0 + xbuf_elems;
    double* kt = reinterpret_cast<double*>(xbuf1 + xbuf_elems);
    const Unit2D u = units[blockIdx.x];
    const Item2D it = items[u.item];
#define RMB_CASE(N)                                                                                        \
    case N:                                                                                                \
        mv2_body<N, KC>(it, prods, xrs_tab, ent_col, ent_val, tab_mask, kpool, X, Y, ldx, ldy, nstates,    \
                        u.s0, active, kt, xbuf0, xbuf1);                                                   \
        break;
    switch (it.nc == 1 ? 1 : (it.nc + 1) & ~1) {
        RMB_CASE(1) RMB_CASE(2) RMB_CASE(4) RMB_CASE(6) RMB_CASE(8)
        RMB_CASE(10) RMB_CASE(12) RMB_CASE(14) RMB_CASE(16)
        default: break;
    }
#undef RMB_CASE
}

}  // namespace rmb
