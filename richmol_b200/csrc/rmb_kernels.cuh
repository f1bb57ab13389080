// Device kernels of the TDSE hot path (sm_100a).  See DESIGN.md for the data layout and the
// per-kernel rooflines.  All arithmetic is IEEE double / complex128.
#pragma once
#include "rmb_internal.h"

namespace rmb {

#ifndef RMB_VEC_MINB
#define RMB_VEC_MINB 6       // CTAs per SM requested for k_recur_conv / k_combine (40 registers, a few spills): measured faster and
                             // steadier than 4 CTAs at 64 registers (OCS, 120 steps: 5.5 vs 6.5-7.7 ms per step)
#endif
constexpr int VEC_THREADS = 256;
constexpr int VEC_PER_THREAD = 4;
constexpr int VEC_CHUNK = VEC_THREADS * VEC_PER_THREAD;   // complex elements per CTA of the vector kernels
constexpr int MV_THREADS = 256;
constexpr int MV_ACC = 4;                                  // outputs per thread of the scalar matvec
constexpr int MAX_ORDER_SMEM = 128;                        // small-exponential scratch (maxorder <= 128)

__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ void cfma(cplx& acc, cplx a, cplx b) {   // acc += a*b
    acc.x += a.x * b.x - a.y * b.y;
    acc.y += a.x * b.y + a.y * b.x;
}
__device__ __forceinline__ double cabs2(cplx a) { return a.x * a.x + a.y * a.y; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Deterministic block sum (fixed tree): result valid in thread 0.
template <int NT>
__device__ __forceinline__ double block_sum(double v, double* sm /* NT/32 */) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sm[wid] = v;
    __syncthreads();
    double r = 0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NT / 32; ++i) r += sm[i];
    }
    __syncthreads();
    return r;
}

// Sum of n partials by one warp in a fixed order; result in all lanes.
__device__ __forceinline__ double warp_reduce_partials(const double* p, int n, int stride) {
    const int lane = threadIdx.x & 31;
    double v = 0;
    for (int i = lane; i < n; i += 32) v += p[(long long)i * stride];
    v = warp_sum(v);
    return __shfl_sync(0xffffffffu, v, 0);
}

// ------------------------------------------------------------------------------------------
// K1: field contraction  MF[e] = sum_c f[c] * M_c[e], element threshold (field.py:1122-1139)
// ------------------------------------------------------------------------------------------
struct FieldProds { double f[16]; };   // products of field components by value (ncart <= 9 for rank <= 2)

__global__ void k_field_contract(long long nent, int ncart, const cplx* __restrict__ coef,
                                 const double* __restrict__ fprod_dev, const FieldProds fp, double thresh, int all_dropped,
                                 cplx* __restrict__ val, int* __restrict__ nz_flag,
                                 const int* __restrict__ ent_tab, const int* __restrict__ tab_off,
                                 const int* __restrict__ tab_nd, unsigned* __restrict__ tab_mask,
                                 unsigned* __restrict__ tab_cplx, long long ent_begin) {
    long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= nent) return;
    cplx acc = make_double2(0.0, 0.0);
    if (!all_dropped) {
        for (int c = 0; c < ncart; ++c) {
            const double f = fprod_dev ? fprod_dev[c] : fp.f[c];
            if (f != 0.0) {
                const cplx m = coef[(long long)c * nent + e];
                acc.x += f * m.x;
                acc.y += f * m.y;
            }
        }
        if (thresh > 0.0 && hypot(acc.x, acc.y) < thresh) acc = make_double2(0.0, 0.0);
    }
    val[e] = acc;
    if (acc.x != 0.0 || acc.y != 0.0) {
        *nz_flag = 1;
        // diagonal slot of this entry inside its (diagonal-aligned) ELL table
        const int t = ent_tab[ent_begin + e];
        const int slot = (int)((ent_begin + e - tab_off[t]) % tab_nd[t]);
        if (slot < 32) atomicOr(&tab_mask[t], 1u << slot);
        if (acc.y != 0.0 && tab_cplx[t] == 0u) tab_cplx[t] = 1u;   // benign race: every writer stores 1
    }
}

// K1b: compaction of the non-zero diagonals.  After k_field_contract the mask of table t tells which
// diagonal slots carry any non-zero MF entry; the surviving diagonals are packed diagonal-major
// (cval/ccol[tab_off + q*dm1 + row], q < popc(mask)), so the matvec only touches diagonals the field
// actually couples and can stage them with contiguous copies.
__global__ void k_compact_tables(long long nent, long long ent_begin, const cplx* __restrict__ val,
                                 const int* __restrict__ col, const int* __restrict__ ent_tab,
                                 const int* __restrict__ tab_off, const int* __restrict__ tab_nd,
                                 const unsigned* __restrict__ tab_mask, double* __restrict__ cent) {
    const long long e = ent_begin + blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= ent_begin + nent) return;
    const int t = ent_tab[e];
    const int nd = tab_nd[t];
    const int rel = (int)(e - tab_off[t]);
    const int r = rel / nd, j = rel - r * nd;
    const unsigned mask = tab_mask[t];
    if (j < 32 && ((mask >> j) & 1u)) {
        const int q = __popc(mask & ((1u << j) - 1u));
        const int c = col[e];
        // rows of the table from its extent (tab_off has ntab + 1 entries)
        const int dm1 = (tab_off[t + 1] - tab_off[t]) / nd;
        const long long dst = (long long)tab_off[t] + (long long)q * dm1 + r;
        // 32-byte entry {re, im, col, pad} (struct MfEntry of the tiled matvec)
        const cplx v = c >= 0 ? val[e] : make_double2(0.0, 0.0);
        double4 out;
        out.x = v.x;
        out.y = v.y;
        out.z = __hiloint2double(0, c);
        out.w = 0.0;
        reinterpret_cast<double4*>(cent)[dst] = out;
    }
}

// ------------------------------------------------------------------------------------------
// K2 (scalar version): y = sum_p (MF_p (x) K_p) x   for one work item and a tile of states.
// Z = MF_p * X_ket is staged in shared memory, then contracted with K_p.
// ------------------------------------------------------------------------------------------
template <bool KC>
__global__ void __launch_bounds__(MV_THREADS)
k_matvec_scalar(const ItemD* __restrict__ items, const ProdD* __restrict__ prods,
                const int* __restrict__ ent_col, const cplx* __restrict__ ent_val,
                const double* __restrict__ kpool, const cplx* __restrict__ X, cplx* __restrict__ Y,
                long long ldx, long long ldy, int nstates, int S, const int* __restrict__ active,
                int zstride) {
    extern __shared__ double2 zs[];   // [S][zstride]
    const ItemD it = items[blockIdx.x];
    const int s0 = blockIdx.y * S;
    const int ns = min(S, nstates - s0);
    const int nout = it.nrows * it.ncols;
    const int total = ns * nout;

    cplx acc[MV_ACC];
    int o_s[MV_ACC], o_r[MV_ACC], o_c[MV_ACC];
#pragma unroll
    for (int i = 0; i < MV_ACC; ++i) {
        acc[i] = make_double2(0.0, 0.0);
        const int o = threadIdx.x + i * MV_THREADS;
        const int oo = o < total ? o : 0;
        o_s[i] = oo / nout;
        const int rem = oo - o_s[i] * nout;
        o_r[i] = rem / it.ncols;
        o_c[i] = rem - o_r[i] * it.ncols;
    }

    for (int p = it.p_begin; p < it.p_end; ++p) {
        const ProdD pr = prods[p];
        // --- Z[s][r][k2] = sum_j MF[r0+r][j] * X[s][ket_off + col*dk2 + k2]
        const int zn = it.nrows * pr.dk2;
        for (int idx = threadIdx.x; idx < ns * zn; idx += MV_THREADS) {
            const int s = idx / zn;
            const int rem = idx - s * zn;
            const int r = rem / pr.dk2;
            const int k2 = rem - r * pr.dk2;
            cplx z = make_double2(0.0, 0.0);
            const int st = s0 + s;
            if (active == nullptr || active[st]) {
                const long long eb = pr.ent_off + (long long)(it.r0 + r) * pr.nd;
                const cplx* xs = X + (long long)st * ldx + pr.ket_off + k2;
                for (int j = 0; j < pr.nd; ++j) {
                    const int col = ent_col[eb + j];
                    if (col >= 0) cfma(z, ent_val[eb + j], xs[(long long)col * (pr.dk2 | 1)]);
                }
            }
            zs[s * zstride + rem] = z;
        }
        __syncthreads();
        // --- acc[s][r][c] += sum_k2 K[c0+c][k2] * Z[s][r][k2]
#pragma unroll
        for (int i = 0; i < MV_ACC; ++i) {
            if (threadIdx.x + i * MV_THREADS < total) {
                const double2* zrow = zs + o_s[i] * zstride + o_r[i] * pr.dk2;
                if (KC) {
                    const cplx* krow = reinterpret_cast<const cplx*>(kpool) + pr.koff +
                                       (long long)(it.c0 + o_c[i]) * pr.dk2;
                    for (int k2 = 0; k2 < pr.dk2; ++k2) cfma(acc[i], krow[k2], zrow[k2]);
                } else {
                    const double* krow = kpool + pr.koff + (long long)(it.c0 + o_c[i]) * pr.dk2;
                    for (int k2 = 0; k2 < pr.dk2; ++k2) {
                        const double kv = krow[k2];
                        acc[i].x += kv * zrow[k2].x;
                        acc[i].y += kv * zrow[k2].y;
                    }
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < MV_ACC; ++i) {
        if (threadIdx.x + i * MV_THREADS < total) {
            const int st = s0 + o_s[i];
            if (active == nullptr || active[st])
                Y[(long long)st * ldy + it.bra_off + (long long)(it.r0 + o_r[i]) * (it.dk1 | 1) + it.c0 + o_c[i]] = acc[i];
        }
    }
}

// ------------------------------------------------------------------------------------------
// Lanczos vector kernels.  Grid (nchunk, nstates); state-major vectors with leading dim ldv.
//
// Storage convention: slab 0 holds V_0 = psi*ph; slab k (k >= 1) holds the *unnormalised* residual
// W_{k-1}, and the Krylov vector of the reference is V_k = slab_k * rinv_k with rinv_k = 1/beta_k
// (tdse.py:455-456).  Every consumer applies rinv_k on the fly, which removes the separate
// normalisation pass over the vector.
// ------------------------------------------------------------------------------------------

// V0 = psi * ph  (tdse.py:375), or a plain copy when ph == nullptr; scatters into the padded layout
// (pad elements of every internal vector are zero at all times)
__global__ void __launch_bounds__(VEC_THREADS)
k_phase_init(const cplx* __restrict__ psi, long long ld, const cplx* __restrict__ ph,
             cplx* __restrict__ V0, long long ldv, long long n, const int* __restrict__ pmap) {
    const long long base = (long long)blockIdx.x * VEC_CHUNK;
    const long long s = blockIdx.y;
#pragma unroll
    for (int i = 0; i < VEC_PER_THREAD; ++i) {
        const long long x = base + threadIdx.x + i * VEC_THREADS;
        if (x < n) {
            cplx v = psi[s * ld + x];
            if (ph) v = cmul(v, ph[x]);
            V0[s * ldv + pmap[x]] = v;
        }
    }
}

// padded -> user layout
__global__ void __launch_bounds__(VEC_THREADS)
k_unpad(const cplx* __restrict__ src, long long ldv, cplx* __restrict__ dst, long long ld, long long n,
        const int* __restrict__ pmap) {
    const long long base = (long long)blockIdx.x * VEC_CHUNK;
    const long long s = blockIdx.y;
#pragma unroll
    for (int i = 0; i < VEC_PER_THREAD; ++i) {
        const long long x = base + threadIdx.x + i * VEC_THREADS;
        if (x < n) dst[s * ld + x] = src[s * ldv + pmap[x]];
    }
}

// in-place psi *= ph (used when the Krylov part is skipped)
__global__ void __launch_bounds__(VEC_THREADS)
k_phase_mul2(cplx* __restrict__ psi, long long ld, const cplx* __restrict__ ph, long long n) {
    const long long base = (long long)blockIdx.x * VEC_CHUNK;
    const long long s = blockIdx.y;
#pragma unroll
    for (int i = 0; i < VEC_PER_THREAD; ++i) {
        const long long x = base + threadIdx.x + i * VEC_THREADS;
        if (x < n) {
            // two successive multiplications, as the reference does (tdse.py:375,394)
            cplx v = cmul(psi[s * ld + x], ph[x]);
            psi[s * ld + x] = cmul(v, ph[x]);
        }
    }
}

// w *= rinv_k per state (only when the matvec could not apply the scale in its epilogue)
__global__ void __launch_bounds__(VEC_THREADS)
k_scale_rows(cplx* __restrict__ w, long long ldv, long long n, const double* __restrict__ scale, int stride,
             const int* __restrict__ active) {
    const long long s = blockIdx.y;
    if (!active[s]) return;
    const double sc = scale[s * stride];
    const long long base = (long long)blockIdx.x * VEC_CHUNK;
#pragma unroll
    for (int i = 0; i < VEC_PER_THREAD; ++i) {
        const long long x = base + threadIdx.x + i * VEC_THREADS;
        if (x < n) {
            cplx v = w[s * ldv + x];
            w[s * ldv + x] = make_double2(v.x * sc, v.y * sc);
        }
    }
}

// partial sum conj(w) * slab_k per chunk (used when the matvec did not fuse the dot product)
__global__ void __launch_bounds__(VEC_THREADS)
k_dot(const cplx* __restrict__ w, const cplx* __restrict__ Vk, long long ldv, long long n,
      cplx* __restrict__ pdot, int npart, const int* __restrict__ active) {
    __shared__ double sm[VEC_THREADS / 32];
    const long long s = blockIdx.y;
    if (active && !active[s]) return;
    const long long base = (long long)blockIdx.x * VEC_CHUNK;
    double re = 0, im = 0;
#pragma unroll
    for (int i = 0; i < VEC_PER_THREAD; ++i) {
        const long long x = base + threadIdx.x + i * VEC_THREADS;
        if (x < n) {
            const cplx a = w[s * ldv + x], b = Vk[s * ldv + x];
            re += a.x * b.x + a.y * b.y;
            im += a.x * b.y - a.y * b.x;
        }
    }
    re = block_sum<VEC_THREADS>(re, sm);
    im = block_sum<VEC_THREADS>(im, sm);
    if (threadIdx.x == 0) pdot[s * npart + blockIdx.x] = make_double2(re, im);
}

// First column of exp(fac * T), T symmetric tridiagonal (k+1 x k+1) with diagonal alpha[0..k] and
// off-diagonal beta[1..k]; scaled Taylor series applied to e_0 (replaces scipy.sparse.linalg.expm,
// tdse.py:474).  One warp; y/term/tmp are shared-memory arrays of >= n entries.
// Register path of warp_expm_col0 for n <= 32: lane i owns row i of the tridiagonal T (alpha_i and its two couplings) and
// element i of the running term / sum; the neighbours' term elements come by shuffle, the scale factors 1 / (j nsub) are
// computed once per call (lane L: j = L + 1 and L + 33) and broadcast, and the stop test runs every fourth term -- the same
// series, sub-stepping and stop criterion as the shared-memory path below, which needed ~600 cycles per term (two
// __syncwarp, two warp_max, a division, six shared-memory round trips) and made the small exponential 85 % of a fused
// single-state step at high Lanczos order (one warp works, fifteen wait).
__device__ __forceinline__ void warp_expm_col0_reg(int n, const cplx* __restrict__ alpha, const double* __restrict__ beta,
                                                   cplx fac, int nsub, double mu, cplx phase, cplx* y_out) {
    const int lane = threadIdx.x & 31;
    const bool in = lane < n;
    const cplx a = in ? make_double2(alpha[lane].x - mu, alpha[lane].y) : make_double2(0.0, 0.0);
    const double bl = (in && lane > 0) ? beta[lane] : 0.0;
    const double bu = (in && lane + 1 < n) ? beta[lane + 1] : 0.0;
    const double sc_lo = 1.0 / ((double)(lane + 1) * (double)nsub);
    const double sc_hi = 1.0 / ((double)(lane + 33) * (double)nsub);
    cplx y = make_double2(lane == 0 ? 1.0 : 0.0, 0.0);
    for (int sub = 0; sub < nsub; ++sub) {
        cplx term = y;
        for (int j = 1; j <= 60; ++j) {
            const double sc = __shfl_sync(0xffffffffu, j <= 32 ? sc_lo : sc_hi, (j - 1) & 31);
            const cplx f = make_double2(fac.x * sc, fac.y * sc);
            const double tmx = __shfl_up_sync(0xffffffffu, term.x, 1), tmy = __shfl_up_sync(0xffffffffu, term.y, 1);
            const double tpx = __shfl_down_sync(0xffffffffu, term.x, 1), tpy = __shfl_down_sync(0xffffffffu, term.y, 1);
            cplx t = cmul(a, term);
            if (lane > 0) { t.x += bl * tmx; t.y += bl * tmy; }
            if (lane + 1 < n) { t.x += bu * tpx; t.y += bu * tpy; }
            t = cmul(f, t);
            y = cadd(y, t);
            term = t;
            if ((j & 3) == 0 || j == 60) {
                const double tmax = warp_max(fmax(fabs(t.x), fabs(t.y)));
                const double ymax = warp_max(fmax(fabs(y.x), fabs(y.y)));
                if (!(tmax > 1e-19 * ymax)) break;
            }
        }
    }
    if (in) y_out[lane] = cmul(phase, y);
    __syncwarp();
}

// c = expm(fac T)[:, 0] for the Hermitian tridiagonal T (diagonal alpha, off-diagonal beta) by one warp: Taylor series of
// nsub sub-steps.  The common part mu of the diagonal (the centre of T's Gershgorin interval: the isotropic shift of the
// interaction, which dominates ||T|| for polarisability Hamiltonians) is split off as the scalar exp(fac mu), so that the
// series runs on T - mu with a spectral bound of half the interval: fewer sub-steps AND less cancellation; sub-steps of norm
// <= EXPM_THETA.
constexpr double EXPM_THETA = 2.0;
__device__ void warp_expm_col0(int n, const cplx* __restrict__ alpha, const double* __restrict__ beta,
                               cplx fac, cplx* y, cplx* term, cplx* tmp) {
    const int lane = threadIdx.x & 31;
    double lo = 1e300, hi = -1e300, im = 0.0;
    for (int i = lane; i < n; i += 32) {
        double r = 0.0;
        if (i > 0) r += fabs(beta[i]);
        if (i + 1 < n) r += fabs(beta[i + 1]);
        lo = fmin(lo, alpha[i].x - r);
        hi = fmax(hi, alpha[i].x + r);
        im = fmax(im, fabs(alpha[i].y));
    }
    lo = -warp_max(-lo);
    hi = warp_max(hi);
    im = warp_max(im);
    const double mu = 0.5 * (lo + hi);
    const double nrm = (0.5 * (hi - lo) + im) * hypot(fac.x, fac.y);
    int nsub = 1;
    if (nrm > EXPM_THETA && nrm < 1e7) nsub = (int)ceil(nrm / EXPM_THETA);
    // exp(fac mu)
    const double pm = exp(fac.x * mu);
    double ps, pc;
    sincos(fac.y * mu, &ps, &pc);
    const cplx phase = make_double2(pm * pc, pm * ps);
    if (n <= 32) {
        warp_expm_col0_reg(n, alpha, beta, fac, nsub, mu, phase, y);
        return;
    }
    for (int i = lane; i < n; i += 32) y[i] = make_double2(i == 0 ? 1.0 : 0.0, 0.0);
    __syncwarp();
    for (int sub = 0; sub < nsub; ++sub) {
        for (int i = lane; i < n; i += 32) term[i] = y[i];
        __syncwarp();
        for (int j = 1; j <= 60; ++j) {
            const double sc = 1.0 / ((double)j * (double)nsub);
            const cplx f = make_double2(fac.x * sc, fac.y * sc);
            double tmax = 0, ymax = 0;
            for (int i = lane; i < n; i += 32) {
                cplx t = cmul(make_double2(alpha[i].x - mu, alpha[i].y), term[i]);
                if (i > 0) { t.x += beta[i] * term[i - 1].x; t.y += beta[i] * term[i - 1].y; }
                if (i + 1 < n) { t.x += beta[i + 1] * term[i + 1].x; t.y += beta[i + 1] * term[i + 1].y; }
                t = cmul(f, t);
                tmp[i] = t;
                const cplx yy = cadd(y[i], t);
                y[i] = yy;
                tmax = fmax(tmax, fmax(fabs(t.x), fabs(t.y)));
                ymax = fmax(ymax, fmax(fabs(yy.x), fabs(yy.y)));
            }
            __syncwarp();
            for (int i = lane; i < n; i += 32) term[i] = tmp[i];
            __syncwarp();
            tmax = warp_max(tmax);
            ymax = warp_max(ymax);
            if (!(tmax > 1e-19 * ymax)) break;
        }
    }
    for (int i = lane; i < n; i += 32) y[i] = cmul(phase, y[i]);
    __syncwarp();
}

// rmb_small_expm: one warp per matrix
__global__ void __launch_bounds__(32)
k_small_expm(int n, const cplx* __restrict__ alpha, const double* __restrict__ beta, cplx fac, cplx* __restrict__ out) {
    __shared__ double2 y[MAX_ORDER_SMEM], term[MAX_ORDER_SMEM], tmp[MAX_ORDER_SMEM];
    __shared__ double2 sa[MAX_ORDER_SMEM];
    __shared__ double sb[MAX_ORDER_SMEM];
    const long long mat = blockIdx.x;
    for (int i = threadIdx.x; i < n; i += 32) {
        sa[i] = alpha[mat * n + i];
        sb[i] = beta[mat * n + i];
    }
    __syncwarp();
    warp_expm_col0(n, sa, sb, fac, y, term, tmp);
    __syncwarp();
    for (int i = threadIdx.x; i < n; i += 32) out[mat * n + i] = y[i];
}

// per state (one warp), after the matvec of iteration k:
//   alpha_k = <w, V_k> (tdse.py:445,468) from the partial sums conj(w)*slab_k (times rinv_k),
//   c^k = expm(fac T_k)[:,0] and dc = c^k - c^{k-1} (tdse.py:474-475); the coefficients are stored
//   pre-multiplied by rinv_i so that u = sum_i ceff_i * slab_i.
__global__ void __launch_bounds__(32)
k_small_a(const cplx* __restrict__ pdot, int npart, cplx* __restrict__ alpha, const double* __restrict__ beta,
          const double* __restrict__ rinv, int tstride, int bstride, int k, cplx fac,
          cplx* __restrict__ ccur, cplx* __restrict__ ceff, cplx* __restrict__ dceff,
          const int* __restrict__ active) {
    __shared__ double2 y[MAX_ORDER_SMEM], term[MAX_ORDER_SMEM], tmp[MAX_ORDER_SMEM];
    const long long s = blockIdx.x;
    if (!active[s]) return;
    const int lane = threadIdx.x;
    const double* p = reinterpret_cast<const double*>(pdot + s * npart);
    const double rk = rinv[s * bstride + k];
    const double re = warp_reduce_partials(p, npart, 2) * rk;
    const double im = warp_reduce_partials(p + 1, npart, 2) * rk;
    if (lane == 0) alpha[s * tstride + k] = make_double2(re, im);
    if (k == 0) {
        if (lane == 0) {
            ccur[s * tstride] = make_double2(1.0, 0.0);   // u_0 = V_0
            ceff[s * tstride] = make_double2(1.0, 0.0);
        }
        return;
    }
    __syncwarp();
    warp_expm_col0(k + 1, alpha + s * tstride, beta + s * bstride, fac, y, term, tmp);
    for (int i = lane; i <= k; i += 32) {
        const cplx prev = (i < k) ? ccur[s * tstride + i] : make_double2(0.0, 0.0);
        const double ri = rinv[s * bstride + i];
        const cplx d = csub(y[i], prev);
        dceff[s * tstride + i] = make_double2(d.x * ri, d.y * ri);
        ceff[s * tstride + i] = make_double2(y[i].x * ri, y[i].y * ri);
        ccur[s * tstride + i] = y[i];
    }
}

// W_k = w - alpha_k V_k - beta_k V_{k-1}  (tdse.py:446,469-470) written to slab k+1, partial |W_k|^2,
// and partial | sum_i dc_i V_i |^2  (u_k - u_{k-1}, tdse.py:475-476) in the same pass
__global__ void __launch_bounds__(VEC_THREADS, RMB_VEC_MINB)
k_recur_conv(const cplx* __restrict__ w, cplx* const* __restrict__ slabs, long long ldv, long long n,
             const cplx* __restrict__ alpha, const double* __restrict__ beta, const double* __restrict__ rinv,
             const cplx* __restrict__ dceff, int tstride, int bstride, int k, double* __restrict__ pnrm,
             double* __restrict__ pconv, int nchunk, const int* __restrict__ active) {
    __shared__ double sm[VEC_THREADS / 32];
    __shared__ double2 sdc[MAX_ORDER_SMEM];
    const long long s = blockIdx.y;
    if (!active[s]) return;
    for (int i = threadIdx.x; i <= k; i += VEC_THREADS) sdc[i] = dceff[s * tstride + i];
    __syncthreads();
    const cplx a = alpha[s * tstride + k];
    const double rk = rinv[s * bstride + k];
    const double b = (k > 0) ? beta[s * bstride + k] : 0.0;
    const double rkm1 = (k > 0) ? rinv[s * bstride + k - 1] : 0.0;
    const cplx* sk = slabs[k] + s * ldv;
    const cplx* skm1 = (k > 0) ? slabs[k - 1] + s * ldv : nullptr;
    cplx* out = slabs[k + 1] + s * ldv;
    const long long base = (long long)blockIdx.x * VEC_CHUNK;
    double nr = 0, cv = 0;
    cplx d[VEC_PER_THREAD];
#pragma unroll
    for (int i = 0; i < VEC_PER_THREAD; ++i) {
        const long long x = base + threadIdx.x + i * VEC_THREADS;
        d[i] = make_double2(0.0, 0.0);
        if (x < n) {
            const cplx raw = sk[x];
            const cplx vk = make_double2(raw.x * rk, raw.y * rk);
            cplx r = csub(w[s * ldv + x], cmul(a, vk));
            if (k > 0) {
                const cplx rawm = skm1[x];
                r.x -= b * (rawm.x * rkm1);
                r.y -= b * (rawm.y * rkm1);
                cfma(d[i], sdc[k], raw);
                cfma(d[i], sdc[k - 1], rawm);
            }
            out[x] = r;
            nr += cabs2(r);
        }
    }
    // older Krylov vectors, two per trip: all 2 x VEC_PER_THREAD loads of a trip are issued before the first use
    // (the kernel is latency-bound otherwise: 52 % of the HBM peak with one vector per trip)
    int j = 0;
    for (; j + 3 <= k; j += 2) {
        const cplx* v0 = slabs[j] + s * ldv;
        const cplx* v1 = slabs[j + 1] + s * ldv;
        const cplx c0 = sdc[j], c1 = sdc[j + 1];
        cplx t0[VEC_PER_THREAD], t1[VEC_PER_THREAD];
#pragma unroll
        for (int i = 0; i < VEC_PER_THREAD; ++i) {
            const long long x = base + threadIdx.x + i * VEC_THREADS;
            t0[i] = t1[i] = make_double2(0.0, 0.0);
            if (x < n) { t0[i] = v0[x]; t1[i] = v1[x]; }
        }
#pragma unroll
        for (int i = 0; i < VEC_PER_THREAD; ++i) { cfma(d[i], c0, t0[i]); cfma(d[i], c1, t1[i]); }
    }
    for (; j + 2 <= k; ++j) {
        const cplx* vj = slabs[j] + s * ldv;
        const cplx c = sdc[j];
#pragma unroll
        for (int i = 0; i < VEC_PER_THREAD; ++i) {
            const long long x = base + threadIdx.x + i * VEC_THREADS;
            if (x < n) cfma(d[i], c, vj[x]);
        }
    }
#pragma unroll
    for (int i = 0; i < VEC_PER_THREAD; ++i) cv += cabs2(d[i]);
    nr = block_sum<VEC_THREADS>(nr, sm);
    cv = block_sum<VEC_THREADS>(cv, sm);
    if (threadIdx.x == 0) {
        pnrm[s * nchunk + blockIdx.x] = nr;
        pconv[s * nchunk + blockIdx.x] = cv;
    }
}

// per state (one CTA), end of iteration k:
//   beta_{k+1} = sqrt(sum |W_k|^2) (tdse.py:453), rinv_{k+1};
//   stop rule (tdse.py:450,478-484): a state leaves the loop when !(conv > tol); reaching
//   k == maxorder-1 raises even if that iteration converged;
//   zero-beta fallback (tdse.py:459-465): V_{k+1} = normalised Gram-Schmidt of the all-ones vector
//   against V_0..V_k, written to slab k+1 with rinv = 1.
__global__ void __launch_bounds__(VEC_THREADS)
k_small_b(const double* __restrict__ pnrm, const double* __restrict__ pconv, int nchunk,
          double* __restrict__ beta, double* __restrict__ rinv, int bstride, int k, double tol, int maxorder,
          int* __restrict__ active, int* __restrict__ order, int* __restrict__ ctrl,
          cplx* const* __restrict__ slabs, long long ldv, long long n, const int* __restrict__ pmap) {
    __shared__ double sm[VEC_THREADS / 32];
    __shared__ double2 bc;
    __shared__ int s_fallback;
    const long long s = blockIdx.x;
    if (!active[s]) return;
    if (threadIdx.x < 32) {
        const double nr = warp_reduce_partials(pnrm + s * nchunk, nchunk, 1);
        double conv = 1.0;
        if (k > 0) conv = warp_reduce_partials(pconv + s * nchunk, nchunk, 1);
        if (threadIdx.x == 0) {
            const double b = sqrt(nr);
            beta[s * bstride + k + 1] = b;
            rinv[s * bstride + k + 1] = (b != 0.0) ? 1.0 / b : 1.0;
            int still = 1;
            if (k > 0) {
                order[s] = k;
                if (k == maxorder - 1) {
                    still = 0;
                    atomicExch(&ctrl[4 * k + 1], 1);
                } else if (!(conv > tol)) {
                    still = 0;
                }
            } else if (maxorder <= 1) {
                still = 0;                               // `while k < maxorder` never entered
                atomicExch(&ctrl[4 * k + 1], 1);
            }
            if (still) atomicAdd(&ctrl[4 * k], 1); else active[s] = 0;
            s_fallback = (still && b == 0.0) ? 1 : 0;
        }
    }
    __syncthreads();
    if (!s_fallback) return;
    // ---- Gram-Schmidt of ones against V_0..V_k (V_j = slab_j * rinv_j); pad elements stay zero
    cplx* v = slabs[k + 1] + s * ldv;
    for (long long x = threadIdx.x; x < n; x += VEC_THREADS) v[pmap[x]] = make_double2(1.0, 0.0);
    __syncthreads();
    for (int j = 0; j <= k; ++j) {
        const cplx* vj = slabs[j] + s * ldv;
        const double rj = rinv[s * bstride + j];
        double pr = 0, pi = 0;   // proj = vdot(V_j, v) = sum conj(V_j) * v
        for (long long x = threadIdx.x; x < n; x += VEC_THREADS) {
            const int xp = pmap[x];
            const cplx a = make_double2(vj[xp].x * rj, vj[xp].y * rj), b = v[xp];
            pr += a.x * b.x + a.y * b.y;
            pi += a.x * b.y - a.y * b.x;
        }
        pr = block_sum<VEC_THREADS>(pr, sm);
        pi = block_sum<VEC_THREADS>(pi, sm);
        if (threadIdx.x == 0) bc = make_double2(pr, pi);
        __syncthreads();
        const cplx proj = bc;
        for (long long x = threadIdx.x; x < n; x += VEC_THREADS) {
            const int xp = pmap[x];
            const cplx a = make_double2(vj[xp].x * rj, vj[xp].y * rj);
            v[xp] = csub(v[xp], cmul(proj, a));
        }
        __syncthreads();
    }
    double nr = 0;
    for (long long x = threadIdx.x; x < n; x += VEC_THREADS) nr += cabs2(v[pmap[x]]);
    nr = block_sum<VEC_THREADS>(nr, sm);
    if (threadIdx.x == 0) bc = make_double2(sqrt(nr), 0.0);
    __syncthreads();
    const double nv = bc.x;
    for (long long x = threadIdx.x; x < n; x += VEC_THREADS) {
        const int xp = pmap[x];
        v[xp] = make_double2(v[xp].x / nv, v[xp].y / nv);
    }
}

// psi = ph * sum_{i<=order} c_i V_i   (tdse.py:475,394), c_i V_i = ceff_i * slab_i
__global__ void __launch_bounds__(VEC_THREADS, RMB_VEC_MINB)
k_combine(cplx* const* __restrict__ slabs, long long ldv, long long n, const cplx* __restrict__ ceff,
          int tstride, const int* __restrict__ order, const cplx* __restrict__ ph,
          cplx* __restrict__ psi, long long ld, const int* __restrict__ pmap) {
    __shared__ double2 sc[MAX_ORDER_SMEM];
    const long long s = blockIdx.y;
    const int k = order[s];
    for (int i = threadIdx.x; i <= k; i += VEC_THREADS) sc[i] = ceff[s * tstride + i];
    __syncthreads();
    const long long base = (long long)blockIdx.x * VEC_CHUNK;
    cplx u[VEC_PER_THREAD];
#pragma unroll
    for (int i = 0; i < VEC_PER_THREAD; ++i) u[i] = make_double2(0.0, 0.0);
    int xp[VEC_PER_THREAD];
#pragma unroll
    for (int i = 0; i < VEC_PER_THREAD; ++i) {
        const long long x = base + threadIdx.x + i * VEC_THREADS;
        xp[i] = x < n ? pmap[x] : 0;
    }
    int j = 0;
    for (; j + 1 <= k; j += 2) {                  // two Krylov vectors per trip (all loads issued before the first use)
        const cplx* v0 = slabs[j] + s * ldv;
        const cplx* v1 = slabs[j + 1] + s * ldv;
        const cplx c0 = sc[j], c1 = sc[j + 1];
        cplx t0[VEC_PER_THREAD], t1[VEC_PER_THREAD];
#pragma unroll
        for (int i = 0; i < VEC_PER_THREAD; ++i) {
            const long long x = base + threadIdx.x + i * VEC_THREADS;
            t0[i] = t1[i] = make_double2(0.0, 0.0);
            if (x < n) { t0[i] = v0[xp[i]]; t1[i] = v1[xp[i]]; }
        }
#pragma unroll
        for (int i = 0; i < VEC_PER_THREAD; ++i) { cfma(u[i], c0, t0[i]); cfma(u[i], c1, t1[i]); }
    }
    for (; j <= k; ++j) {
        const cplx* vj = slabs[j] + s * ldv;
        const cplx c = sc[j];
#pragma unroll
        for (int i = 0; i < VEC_PER_THREAD; ++i) {
            const long long x = base + threadIdx.x + i * VEC_THREADS;
            if (x < n) cfma(u[i], c, vj[xp[i]]);
        }
    }
#pragma unroll
    for (int i = 0; i < VEC_PER_THREAD; ++i) {
        const long long x = base + threadIdx.x + i * VEC_THREADS;
        if (x < n) psi[s * ld + x] = ph ? cmul(u[i], ph[x]) : u[i];
    }
}

// per-state initialisation of a Lanczos batch
__global__ void k_init_states(int* __restrict__ active, int* __restrict__ order, double* __restrict__ rinv,
                              double* __restrict__ beta, int bstride, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        active[i] = 1;
        order[i] = 0;
        rinv[(long long)i * bstride] = 1.0;
        beta[(long long)i * bstride] = 0.0;
    }
}

// copies n ints, e.g. the control words of a Lanczos iteration into the pinned, mapped host mirror (system-scope fence:
// visible to the host once the event recorded behind this kernel has completed) -- a kernel instead of a D2H memcpy so
// that nothing on the compute stream ever queues behind a large download on the copy engine
__global__ void k_publish(const int* __restrict__ src, int* __restrict__ dst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
    __threadfence_system();
}

__global__ void k_fill_int(int* p, int v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// K5: partial <a|b> = sum conj(a) * b with separate leading dimensions
__global__ void __launch_bounds__(VEC_THREADS)
k_dot2(const cplx* __restrict__ a, long long lda, const cplx* __restrict__ b, long long ldb, long long n,
       cplx* __restrict__ pdot, int nchunk) {
    __shared__ double sm[VEC_THREADS / 32];
    const long long s = blockIdx.y;
    const long long base = (long long)blockIdx.x * VEC_CHUNK;
    double re = 0, im = 0;
#pragma unroll
    for (int i = 0; i < VEC_PER_THREAD; ++i) {
        const long long x = base + threadIdx.x + i * VEC_THREADS;
        if (x < n) {
            const cplx u = a[s * lda + x], v = b[s * ldb + x];
            re += u.x * v.x + u.y * v.y;
            im += u.x * v.y - u.y * v.x;
        }
    }
    re = block_sum<VEC_THREADS>(re, sm);
    im = block_sum<VEC_THREADS>(im, sm);
    if (threadIdx.x == 0) pdot[s * nchunk + blockIdx.x] = make_double2(re, im);
}

// final reduction of per-chunk partials to expval[s]
__global__ void __launch_bounds__(32)
k_reduce_dot(const cplx* __restrict__ pdot, int nchunk, cplx* __restrict__ out, double imsign) {
    const long long s = blockIdx.x;
    const double* p = reinterpret_cast<const double*>(pdot + s * nchunk);
    const double re = warp_reduce_partials(p, nchunk, 2);
    const double im = warp_reduce_partials(p + 1, nchunk, 2);
    if (threadIdx.x == 0) out[s] = make_double2(re, imsign * im);
}

// pop[i] = sum_s |psi_s[i]|^2
__global__ void k_populations(const cplx* __restrict__ psi, long long nstates, long long n,
                              long long ld, double* __restrict__ pop) {
    const long long x = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (x >= n) return;
    double acc = 0;
    for (long long s = 0; s < nstates; ++s) acc += cabs2(psi[s * ld + x]);
    pop[x] = acc;
}

// ---- FP64 roofline denominators (rmb_fp64_peak): register-resident DFMA and DMMA loops
__global__ void k_peak_dfma(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void k_peak_dmma(double* out, int iters) {
    double c0[4][2] = {{0}};
    const double a = threadIdx.x * 1e-9, b = 1.0000001;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[j][0]), "+d"(c0[j][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int j = 0; j < 4; ++j) s += c0[j][0] + c0[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- laboratory-frame tensor generator (richmol/rot/labtens.py:482-523) -------------------------
// Wigner 3j symbol (j1 j2 j3; m1 m2 m3) for integer arguments by the Racah sum with log-factorials; with the small
// tensor rank j2 <= 2 used here the sum has at most 2 j2 + 1 terms and no cancellation to speak of.
__device__ __forceinline__ double lfact(int n) { return lgamma((double)(n > 0 ? n : 0) + 1.0); }
__device__ double wigner3j_int(int j1, int j2, int j3, int m1, int m2, int m3) {
    if (m1 + m2 + m3 != 0 || abs(m1) > j1 || abs(m2) > j2 || abs(m3) > j3 || j3 < abs(j1 - j2) || j3 > j1 + j2) return 0.0;
    const double lnpref = 0.5 * (lfact(j1 + j2 - j3) + lfact(j1 - j2 + j3) + lfact(-j1 + j2 + j3) - lfact(j1 + j2 + j3 + 1) +
                                 lfact(j1 + m1) + lfact(j1 - m1) + lfact(j2 + m2) + lfact(j2 - m2) + lfact(j3 + m3) +
                                 lfact(j3 - m3));
    const int tmin = max(0, max(j2 - j3 - m1, j1 - j3 + m2));
    const int tmax = min(j1 + j2 - j3, min(j1 - m1, j2 + m2));
    double res = 0.0;
    for (int t = tmin; t <= tmax; ++t) {
        const double lnden = lfact(t) + lfact(j3 - j2 + t + m1) + lfact(j3 - j1 + t - m2) + lfact(j1 + j2 - j3 - t) +
                             lfact(j1 - t - m1) + lfact(j2 - t + m2);
        const double term = exp(lnpref - lnden);
        res += (t & 1) ? -term : term;
    }
    return (abs(j1 - j2 - m3) & 1) ? -res : res;
}

// out[c][a][b] = pref * (-1)^|qa| * sum_sigma coef[c][sigma + omega] * 3j(j2 omega j1; qb sigma -qa),  qa = a - j1, qb = b - j2
// (a, b: positions of the projection quantum numbers -j..j).  With coef = Ux[cart,(omega,sigma)] and
// pref = sqrt((2 j1 + 1)(2 j2 + 1)) this is the M tensor (labtens.py:504-523); with coef = (Us T)_{omega sigma} and
// pref = 1 the primitive K tensor over |J,k> (labtens.py:482-502).  Only b = a - j1 + j2 - sigma contributes.
__global__ void k_threej_band(int j1, int j2, int omega, int ncoef, const cplx* __restrict__ coef, double pref,
                              cplx* __restrict__ out) {
    const int d1 = 2 * j1 + 1, d2 = 2 * j2 + 1;
    const long long total = (long long)ncoef * d1 * d2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i % d2);
        const int a = (int)((i / d2) % d1);
        const int c = (int)(i / ((long long)d1 * d2));
        const int qa = a - j1, qb = b - j2;
        const int sigma = qa - qb;
        cplx v = make_double2(0.0, 0.0);
        if (abs(sigma) <= omega) {
            const double tj = wigner3j_int(j2, omega, j1, qb, sigma, -qa) * pref * ((abs(qa) & 1) ? -1.0 : 1.0);
            const cplx cf = coef[c * (2 * omega + 1) + sigma + omega];
            v = make_double2(cf.x * tj, cf.y * tj);
        }
        out[i] = v;
    }
}

}  // namespace rmb
