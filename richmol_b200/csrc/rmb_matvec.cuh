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

constexpr int MV2_THREADS = 256;
constexpr int MV2_NCMAX = 12;     // columns (k1) per thread
constexpr int MV2_NDMAX = 5;      // max ELL width handled by the tiled kernel (rank <= 2)
constexpr int MV2_SMAX = 32;      // max states per CTA

struct Item2D {
    long long bra_off;
    int dk1, dm1;
    int r0, nrows;       // rows (m1) of the tile, nrows <= MV2_THREADS
    int c0, nc;          // columns (k1) of the tile, nc <= MV2_NCMAX
    int p_begin, p_end;
    int nst;             // states per CTA (threads = nst * nrows <= MV2_THREADS)
    int xr_off;          // offset into the per-(item, product) ket row ranges
    int kt_total;        // doubles of K^T staged in shared memory for this item
};

struct XRange { int c_lo, nr; };   // ket rows [c_lo, c_lo + nr) needed by (item, product)
struct Unit2D { int item, s0; };

// per-product descriptor staged in shared memory at CTA start
struct ProdS {
    long long ket_off;   // + c_lo * dk2 already added
    long long ent_off;   // first entry of the MF table
    int dk2, nnz, c_lo, nr, xrs, pad0, pad1, pad2;
};

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Stage one product for all states of the CTA:
//  * ket rows X[c_lo .. c_lo+nr) of every active state with coalesced 16-byte cp.async copies; rows are
//    padded to an odd number of 16-byte words (bank-conflict-free row-strided reads),
//  * the surviving MF diagonals of the tile rows (values + shared-memory offsets of their ket rows).
__device__ __forceinline__ void mv2_stage(double2* xb, double2* mfs, int* xos, const ProdS& pr,
                                          const Item2D& it, const double2* __restrict__ X,
                                          const long long* __restrict__ sbase, const double2* __restrict__ cval,
                                          const int* __restrict__ ccol) {
    // the internal vectors store rows of (dim_k | 1) elements, so the ket rows are one contiguous run
    const int per_state = pr.nr * pr.xrs;
    const unsigned dst0 = (unsigned)__cvta_generic_to_shared(xb);
    const unsigned sstride = (unsigned)per_state * 16u;
    const double2* src0 = X + pr.ket_off;
    for (int e = threadIdx.x; e < per_state; e += MV2_THREADS) {
        const double2* src = src0 + e;
#pragma unroll 4
        for (int s = 0; s < it.nst; ++s) {
            const long long sb = sbase[s];
            if (sb >= 0)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst0 + s * sstride + e * 16u), "l"(src + sb));
        }
    }
    // MF diagonals: mfs[q][rl], xos[q][rl]
    const int n = pr.nnz * it.nrows;
    for (int idx = threadIdx.x; idx < n; idx += MV2_THREADS) {
        const int q = idx / it.nrows, r = idx - q * it.nrows;
        const long long e = pr.ent_off + (long long)q * it.dm1 + it.r0 + r;
        const int col = ccol[e];
        mfs[q * it.nrows + r] = cval[e];
        xos[q * it.nrows + r] = col >= 0 ? (col - pr.c_lo) * pr.xrs : 0;   // value is 0 off the block edge
    }
}

// one block product for one thread: acc[k1] += sum_k2 K[k1,k2] * (sum_q MF[q] * X[row_q, k2])
template <int NC, int NNZ, bool KC>
__device__ __forceinline__ void mv2_inner(const double2* __restrict__ xa, const double2* __restrict__ mfs,
                                          const int* __restrict__ xos, int nrows,
                                          const double* __restrict__ ktp, int dk2, double2 (&acc)[NC]) {
    double2 mf[NNZ];
    int xo[NNZ];
#pragma unroll
    for (int q = 0; q < NNZ; ++q) {
        mf[q] = mfs[q * nrows];
        xo[q] = xos[q * nrows];
    }
#pragma unroll 2
    for (int k2 = 0; k2 < dk2; ++k2) {
        double2 z = make_double2(0.0, 0.0);
#pragma unroll
        for (int q = 0; q < NNZ; ++q) {
            const double2 a = xa[xo[q] + k2];
            z.x = fma(mf[q].x, a.x, z.x);
            z.y = fma(mf[q].x, a.y, z.y);
            z.x = fma(-mf[q].y, a.y, z.x);
            z.y = fma(mf[q].y, a.x, z.y);
        }
        if (KC) {
            const double2* krow = reinterpret_cast<const double2*>(ktp) + k2 * NC;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const double2 kv = krow[c];
                acc[c].x = fma(kv.x, z.x, acc[c].x);
                acc[c].y = fma(kv.x, z.y, acc[c].y);
                acc[c].x = fma(-kv.y, z.y, acc[c].x);
                acc[c].y = fma(kv.y, z.x, acc[c].y);
            }
        } else if (NC == 1) {
            const double kv = ktp[k2];
            acc[0].x = fma(kv, z.x, acc[0].x);
            acc[0].y = fma(kv, z.y, acc[0].y);
        } else {
            const double2* krow = reinterpret_cast<const double2*>(ktp + k2 * NC);
#pragma unroll
            for (int c2 = 0; c2 < NC / 2; ++c2) {
                const double2 kv = krow[c2];
                acc[2 * c2].x = fma(kv.x, z.x, acc[2 * c2].x);
                acc[2 * c2].y = fma(kv.x, z.y, acc[2 * c2].y);
                acc[2 * c2 + 1].x = fma(kv.y, z.x, acc[2 * c2 + 1].x);
                acc[2 * c2 + 1].y = fma(kv.y, z.y, acc[2 * c2 + 1].y);
            }
        }
    }
}

struct Mv2Smem {
    double2* xbuf[2];
    double2* mfs[2];
    int* xos[2];
    double* kt;
    ProdS* sp;
    long long* sbase;
};

// NC = number of register columns (>= it.nc; 1 or even), surplus columns are zero-padded in K^T
template <int NC, bool KC>
__device__ __forceinline__ void mv2_body(const Item2D& it, const ProdD* __restrict__ prods,
                                         const XRange* __restrict__ xrs_tab, const int* __restrict__ ccol,
                                         const double2* __restrict__ cval, const unsigned* __restrict__ tab_mask,
                                         const double* __restrict__ kpool, const double2* __restrict__ X,
                                         double2* __restrict__ Y, long long ldx, long long ldy, int nstates,
                                         int s0, const int* __restrict__ active, const Mv2Smem& sm,
                                         const double* __restrict__ scale, int scale_stride,
                                         double2* __restrict__ pdot, int npart, int item_index) {
    constexpr int KW = KC ? 2 : 1;                       // doubles per K element
    const int sl = threadIdx.x / it.nrows;
    const int rl = threadIdx.x - sl * it.nrows;
    const int st = s0 + sl;
    const bool work = sl < it.nst && st < nstates && (active == nullptr || active[st]);
    const int np = it.p_end - it.p_begin;
    if (__syncthreads_or(work) == 0) return;             // every state of the tile has converged

    // ---- product descriptors and state base offsets -> shared memory
    for (int ip = threadIdx.x; ip < np; ip += MV2_THREADS) {
        const ProdD pr = prods[it.p_begin + ip];
        const XRange xr = xrs_tab[it.xr_off + ip];
        ProdS d;
        d.ket_off = pr.ket_off + (long long)xr.c_lo * (pr.dk2 | 1);
        d.ent_off = pr.ent_off;
        d.dk2 = pr.dk2;
        d.nnz = min(__popc(tab_mask[pr.tab]), MV2_NDMAX);   // diagonals that survived the field contraction
        d.c_lo = xr.c_lo;
        d.nr = xr.nr;
        d.xrs = pr.dk2 | 1;
        d.pad0 = d.pad1 = d.pad2 = 0;
        sm.sp[ip] = d;
    }
    if (threadIdx.x < it.nst) {
        const int s = s0 + threadIdx.x;
        sm.sbase[threadIdx.x] = (s < nstates && (active == nullptr || active[s])) ? (long long)s * ldx : -1;
    }
    __syncthreads();
    if (np > 0) mv2_stage(sm.xbuf[0], sm.mfs[0], sm.xos[0], sm.sp[0], it, X, sm.sbase, cval, ccol);
    cp_async_commit();
    // ---- K^T of every product of the item -> shared memory: kt[p][k2][NC]
    {
        int base = 0;
        for (int ip = 0; ip < np; ++ip) {
            const ProdD pr = prods[it.p_begin + ip];
            const int n = pr.dk2 * NC;
            for (int idx = threadIdx.x; idx < n; idx += MV2_THREADS) {
                const int k2 = idx / NC, c = idx - k2 * NC;
                if (KC) {
                    double2 v = make_double2(0.0, 0.0);
                    if (c < it.nc) v = reinterpret_cast<const double2*>(kpool)[pr.koff + (long long)(it.c0 + c) * pr.dk2 + k2];
                    reinterpret_cast<double2*>(sm.kt + base)[idx] = v;
                } else {
                    sm.kt[base + idx] = (c < it.nc) ? kpool[pr.koff + (long long)(it.c0 + c) * pr.dk2 + k2] : 0.0;
                }
            }
            base += n * KW;
        }
    }
    double2 acc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[c] = make_double2(0.0, 0.0);

    int ktbase = 0;
    for (int ip = 0; ip < np; ++ip) {
        const int cur = ip & 1;
        cp_async_wait<0>();
        __syncthreads();       // product ip has landed; everyone is done with product ip-1
        if (ip + 1 < np)
            mv2_stage(sm.xbuf[cur ^ 1], sm.mfs[cur ^ 1], sm.xos[cur ^ 1], sm.sp[ip + 1], it, X, sm.sbase, cval, ccol);
        cp_async_commit();
        const int dk2 = sm.sp[ip].dk2, nnz = sm.sp[ip].nnz;
        if (work) {
            const double2* xa = sm.xbuf[cur] + (long long)sl * sm.sp[ip].nr * sm.sp[ip].xrs;
            const double2* mfs = sm.mfs[cur] + rl;
            const int* xos = sm.xos[cur] + rl;
            const double* ktp = sm.kt + ktbase;
            switch (nnz) {
                case 0: break;
                case 1: mv2_inner<NC, 1, KC>(xa, mfs, xos, it.nrows, ktp, dk2, acc); break;
                case 2: mv2_inner<NC, 2, KC>(xa, mfs, xos, it.nrows, ktp, dk2, acc); break;
                case 3: mv2_inner<NC, 3, KC>(xa, mfs, xos, it.nrows, ktp, dk2, acc); break;
                case 4: mv2_inner<NC, 4, KC>(xa, mfs, xos, it.nrows, ktp, dk2, acc); break;
                default: mv2_inner<NC, 5, KC>(xa, mfs, xos, it.nrows, ktp, dk2, acc); break;
            }
        }
        ktbase += dk2 * NC * KW;
    }
    cp_async_wait<0>();
    // ---- epilogue: optional per-state scale (w = rinv_k * H slab_k), store, fused partial dot
    //      sum conj(w) * x over the rows of this tile (alpha of the Lanczos recurrence, tdse.py:468)
    double pre = 0.0, pim = 0.0;
    if (work) {
        const long long row_off = it.bra_off + (long long)(it.r0 + rl) * (it.dk1 | 1) + it.c0;
        if (scale != nullptr) {
            const double sc = scale[(long long)st * scale_stride];
#pragma unroll
            for (int c = 0; c < NC; ++c) { acc[c].x *= sc; acc[c].y *= sc; }
        }
        if (Y != nullptr) {
            double2* y = Y + (long long)st * ldy + row_off;
#pragma unroll
            for (int c = 0; c < NC; ++c)
                if (c < it.nc) y[c] = acc[c];
        }
        if (pdot != nullptr) {
            const double2* x = X + (long long)st * ldx + row_off;
#pragma unroll
            for (int c = 0; c < NC; ++c)
                if (c < it.nc) {
                    const double2 v = x[c];
                    pre += acc[c].x * v.x + acc[c].y * v.y;
                    pim += acc[c].x * v.y - acc[c].y * v.x;
                }
        }
    }
    if (pdot != nullptr) {
        __syncthreads();                        // shared buffers are free again
        double* red = reinterpret_cast<double*>(sm.xbuf[0]);
        red[threadIdx.x] = pre;
        red[MV2_THREADS + threadIdx.x] = pim;
        __syncthreads();
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int s = warp; s < it.nst; s += MV2_THREADS / 32) {
            const int sg = s0 + s;
            if (sg >= nstates || (active != nullptr && !active[sg])) continue;
            double a = 0.0, b = 0.0;
            for (int r = lane; r < it.nrows; r += 32) {
                a += red[s * it.nrows + r];
                b += red[MV2_THREADS + s * it.nrows + r];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a += __shfl_down_sync(0xffffffffu, a, o);
                b += __shfl_down_sync(0xffffffffu, b, o);
            }
            if (lane == 0) pdot[(long long)sg * npart + item_index] = make_double2(a, b);
        }
    }
}

template <bool KC>
__global__ void __launch_bounds__(MV2_THREADS, 2)
k_matvec_tiled(const Unit2D* __restrict__ units, const Item2D* __restrict__ items,
               const ProdD* __restrict__ prods, const XRange* __restrict__ xrs_tab,
               const int* __restrict__ ccol, const double2* __restrict__ cval,
               const unsigned* __restrict__ tab_mask, const double* __restrict__ kpool,
               const double2* __restrict__ X, double2* __restrict__ Y, long long ldx, long long ldy,
               int nstates, const int* __restrict__ active, int np_max, int kt_doubles, int xbuf_elems,
               int mf_elems, const double* __restrict__ scale, int scale_stride, double2* __restrict__ pdot,
               int npart) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Mv2Smem sm;
    sm.xbuf[0] = reinterpret_cast<double2*>(smem_raw);
    sm.xbuf[1] = sm.xbuf[0] + xbuf_elems;
    sm.mfs[0] = sm.xbuf[1] + xbuf_elems;
    sm.mfs[1] = sm.mfs[0] + mf_elems;
    sm.kt = reinterpret_cast<double*>(sm.mfs[1] + mf_elems);
    sm.sp = reinterpret_cast<ProdS*>(sm.kt + kt_doubles);
    sm.sbase = reinterpret_cast<long long*>(sm.sp + np_max);
    sm.xos[0] = reinterpret_cast<int*>(sm.sbase + MV2_SMAX);
    sm.xos[1] = sm.xos[0] + mf_elems;
    const Unit2D u = units[blockIdx.x];
    const Item2D it = items[u.item];
#define RMB_CASE(N)                                                                                        \
    case N:                                                                                                \
        mv2_body<N, KC>(it, prods, xrs_tab, ccol, cval, tab_mask, kpool, X, Y, ldx, ldy, nstates,          \
                        u.s0, active, sm, scale, scale_stride, pdot, npart, u.item);                       \
        break;
    switch (it.nc == 1 ? 1 : (it.nc + 1) & ~1) {
        RMB_CASE(1) RMB_CASE(2) RMB_CASE(4) RMB_CASE(6) RMB_CASE(8)
        RMB_CASE(10) RMB_CASE(12)
        default: break;
    }
#undef RMB_CASE
}

}  // namespace rmb
