// Device kernels of the TDSE hot path (sm_100a).  See DESIGN.md for the data layout and the
// per-kernel rooflines.  All arithmetic is IEEE double / complex128.
#pragma once
#include "rmb_internal.h"

namespace rmb {

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
__global__ void k_field_contract(long long nent, int ncart, const cplx* __restrict__ coef,
                                 const double* __restrict__ fprod, double thresh, int all_dropped,
                                 cplx* __restrict__ val, int* __restrict__ nz_flag,
                                 const int* __restrict__ ent_tab, const int* __restrict__ tab_off,
                                 const int* __restrict__ tab_nd, unsigned* __restrict__ tab_mask,
                                 long long ent_begin) {
    long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= nent) return;
    cplx acc = make_double2(0.0, 0.0);
    if (!all_dropped) {
        for (int c = 0; c < ncart; ++c) {
            const double f = fprod[c];
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
                    if (col >= 0) cfma(z, ent_val[eb + j], xs[(long long)col * pr.dk2]);
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
                Y[(long long)st * ldy + it.bra_off + (long long)(it.r0 + o_r[i]) * it.dk1 + it.c0 + o_c[i]] = acc[i];
        }
    }
}

// ------------------------------------------------------------------------------------------
// Lanczos vector kernels.  Grid (nchunk, nstates); state-major vectors with leading dim ldv.
// ------------------------------------------------------------------------------------------

// V0 = psi * ph  (tdse.py:375), or a plain copy when ph == nullptr
__global__ void __launch_bounds__(VEC_THREADS)
k_phase_init(const cplx* __restrict__ psi, long long ld, const cplx* __restrict__ ph,
             cplx* __restrict__ V0, long long ldv, long long n) {
    const long long base = (long long)blockIdx.x * VEC_CHUNK;
    const long long s = blockIdx.y;
#pragma unroll
    for (int i = 0; i < VEC_PER_THREAD; ++i) {
        const long long x = base + threadIdx.x + i * VEC_THREADS;
        if (x < n) {
            cplx v = psi[s * ld + x];
            if (ph) v = cmul(v, ph[x]);
            V0[s * ldv + x] = v;
        }
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

// V_k = W_{k-1} / beta_k   (tdse.py:455-456); states with beta == 0 are left to k_fallback_ones
__global__ void __launch_bounds__(VEC_THREADS)
k_scale(const cplx* __restrict__ W, cplx* __restrict__ Vk, long long ldv, long long n,
        const double* __restrict__ beta, int bstride, int k, const int* __restrict__ active) {
    const long long s = blockIdx.y;
    if (!active[s]) return;
    const double b = beta[s * bstride + k];
    if (b == 0.0) return;
    const long long base = (long long)blockIdx.x * VEC_CHUNK;
#pragma unroll
    for (int i = 0; i < VEC_PER_THREAD; ++i) {
        const long long x = base + threadIdx.x + i * VEC_THREADS;
        if (x < n) {
            const cplx w = W[s * ldv + x];
            Vk[s * ldv + x] = make_double2(w.x / b, w.y / b);
        }
    }
}

// zero-beta fallback (tdse.py:459-465): V_k = normalised Gram-Schmidt of the all-ones vector
// against V_0..V_{k-1}.  One CTA per state; only runs for states with beta_k == 0.
__global__ void __launch_bounds__(VEC_THREADS)
k_fallback_ones(cplx* const* __restrict__ slabs, long long ldv, long long n,
                const double* __restrict__ beta, int bstride, int k, const int* __restrict__ active) {
    __shared__ double sm[VEC_THREADS / 32];
    __shared__ double2 bc;
    const long long s = blockIdx.x;
    if (!active[s] || beta[s * bstride + k] != 0.0) return;
    cplx* v = slabs[k] + s * ldv;
    for (long long x = threadIdx.x; x < n; x += VEC_THREADS) v[x] = make_double2(1.0, 0.0);
    __syncthreads();
    for (int j = 0; j < k; ++j) {
        const cplx* vj = slabs[j] + s * ldv;
        double pr = 0, pi = 0;   // proj = vdot(V_j, v) = sum conj(V_j) * v
        for (long long x = threadIdx.x; x < n; x += VEC_THREADS) {
            const cplx a = vj[x], b = v[x];
            pr += a.x * b.x + a.y * b.y;
            pi += a.x * b.y - a.y * b.x;
        }
        pr = block_sum<VEC_THREADS>(pr, sm);
        pi = block_sum<VEC_THREADS>(pi, sm);
        if (threadIdx.x == 0) bc = make_double2(pr, pi);
        __syncthreads();
        const cplx proj = bc;
        for (long long x = threadIdx.x; x < n; x += VEC_THREADS) v[x] = csub(v[x], cmul(proj, vj[x]));
        __syncthreads();
    }
    double nr = 0;
    for (long long x = threadIdx.x; x < n; x += VEC_THREADS) nr += cabs2(v[x]);
    nr = block_sum<VEC_THREADS>(nr, sm);
    if (threadIdx.x == 0) bc = make_double2(sqrt(nr), 0.0);
    __syncthreads();
    const double nv = bc.x;
    for (long long x = threadIdx.x; x < n; x += VEC_THREADS) v[x] = make_double2(v[x].x / nv, v[x].y / nv);
}

// partial alpha: pdot[s][chunk] = sum conj(w) * V_k   (np.vdot(w, V[k]), tdse.py:445,468)
__global__ void __launch_bounds__(VEC_THREADS)
k_dot(const cplx* __restrict__ w, const cplx* __restrict__ Vk, long long ldv, long long n,
      cplx* __restrict__ pdot, int nchunk, const int* __restrict__ active) {
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
    if (threadIdx.x == 0) pdot[s * nchunk + blockIdx.x] = make_double2(re, im);
}

// W_k = w - alpha_k V_k - beta_k V_{k-1}   (tdse.py:446,469-470) + partial |W_k|^2
__global__ void __launch_bounds__(VEC_THREADS)
k_recur(const cplx* __restrict__ w, const cplx* __restrict__ Vk, const cplx* __restrict__ Vkm1,
        cplx* __restrict__ W, long long ldv, long long n, const cplx* __restrict__ pdot, int nchunk,
        cplx* __restrict__ alpha, const double* __restrict__ beta, int tstride, int bstride, int k,
        double* __restrict__ pnrm, const int* __restrict__ active) {
    __shared__ double sm[VEC_THREADS / 32];
    __shared__ double2 s_alpha;
    const long long s = blockIdx.y;
    if (!active[s]) return;
    if (threadIdx.x < 32) {
        const double* p = reinterpret_cast<const double*>(pdot + s * nchunk);
        const double re = warp_reduce_partials(p, nchunk, 2);
        const double im = warp_reduce_partials(p + 1, nchunk, 2);
        if (threadIdx.x == 0) {
            s_alpha = make_double2(re, im);
            if (blockIdx.x == 0) alpha[s * tstride + k] = make_double2(re, im);
        }
    }
    __syncthreads();
    const cplx a = s_alpha;
    const double b = (k > 0) ? beta[s * bstride + k] : 0.0;
    const long long base = (long long)blockIdx.x * VEC_CHUNK;
    double nr = 0;
#pragma unroll
    for (int i = 0; i < VEC_PER_THREAD; ++i) {
        const long long x = base + threadIdx.x + i * VEC_THREADS;
        if (x < n) {
            cplx r = csub(w[s * ldv + x], cmul(a, Vk[s * ldv + x]));
            if (k > 0) {
                const cplx vm = Vkm1[s * ldv + x];
                r.x -= b * vm.x;
                r.y -= b * vm.y;
            }
            W[s * ldv + x] = r;
            nr += cabs2(r);
        }
    }
    nr = block_sum<VEC_THREADS>(nr, sm);
    if (threadIdx.x == 0) pnrm[s * nchunk + blockIdx.x] = nr;
}

// First column of exp(fac * T), T symmetric tridiagonal (k+1 x k+1) with diagonal alpha[0..k] and
// off-diagonal beta[1..k]; scaled Taylor series applied to e_0 (replaces scipy.sparse.linalg.expm,
// tdse.py:474).  One warp; y/term/tmp are shared-memory arrays of >= n entries.
__device__ void warp_expm_col0(int n, const cplx* __restrict__ alpha, const double* __restrict__ beta,
                               cplx fac, cplx* y, cplx* term, cplx* tmp) {
    const int lane = threadIdx.x & 31;
    double loc = 0;
    for (int i = lane; i < n; i += 32) {
        double cs = hypot(alpha[i].x, alpha[i].y);
        if (i > 0) cs += fabs(beta[i]);
        if (i + 1 < n) cs += fabs(beta[i + 1]);
        loc = fmax(loc, cs);
    }
    const double nrm = warp_max(loc) * hypot(fac.x, fac.y);
    int nsub = 1;
    if (nrm > 1.0 && nrm < 1e7) nsub = (int)ceil(nrm);
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
                cplx t = cmul(alpha[i], term[i]);
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
}

// per state (one warp): beta_{k+1} = sqrt(sum |W_k|^2) (tdse.py:453), coefficients
// c^k = expm(fac T_k)[:,0] and dc = c^k - c^{k-1} (tdse.py:474-475)
__global__ void __launch_bounds__(32)
k_small(const double* __restrict__ pnrm, int nchunk, const cplx* __restrict__ alpha,
        double* __restrict__ beta, int tstride, int bstride, int k, cplx fac,
        cplx* __restrict__ ccur, cplx* __restrict__ dc, const int* __restrict__ active) {
    __shared__ double2 y[MAX_ORDER_SMEM], term[MAX_ORDER_SMEM], tmp[MAX_ORDER_SMEM];
    const long long s = blockIdx.x;
    if (!active[s]) return;
    const int lane = threadIdx.x;
    const double nr = warp_reduce_partials(pnrm + s * nchunk, nchunk, 1);
    if (lane == 0) beta[s * bstride + k + 1] = sqrt(nr);
    if (k == 0) {
        if (lane == 0) ccur[s * tstride] = make_double2(1.0, 0.0);   // u_0 = V_0
        return;
    }
    warp_expm_col0(k + 1, alpha + s * tstride, beta + s * bstride, fac, y, term, tmp);
    for (int i = lane; i <= k; i += 32) {
        const cplx prev = (i < k) ? ccur[s * tstride + i] : make_double2(0.0, 0.0);
        dc[s * tstride + i] = csub(y[i], prev);
        ccur[s * tstride + i] = y[i];
    }
}

// partial conv: pconv[s][chunk] = sum | sum_i dc_i V_i |^2  (u_k - u_{k-1}, tdse.py:475-476)
__global__ void __launch_bounds__(VEC_THREADS)
k_conv(cplx* const* __restrict__ slabs, long long ldv, long long n, const cplx* __restrict__ dc,
       int tstride, int k, double* __restrict__ pconv, int nchunk, const int* __restrict__ active) {
    __shared__ double sm[VEC_THREADS / 32];
    __shared__ double2 sdc[MAX_ORDER_SMEM];
    const long long s = blockIdx.y;
    if (!active[s]) return;
    for (int i = threadIdx.x; i <= k; i += VEC_THREADS) sdc[i] = dc[s * tstride + i];
    __syncthreads();
    const long long base = (long long)blockIdx.x * VEC_CHUNK;
    cplx d[VEC_PER_THREAD];
#pragma unroll
    for (int i = 0; i < VEC_PER_THREAD; ++i) d[i] = make_double2(0.0, 0.0);
    for (int j = 0; j <= k; ++j) {
        const cplx* vj = slabs[j] + s * ldv;
        const cplx c = sdc[j];
#pragma unroll
        for (int i = 0; i < VEC_PER_THREAD; ++i) {
            const long long x = base + threadIdx.x + i * VEC_THREADS;
            if (x < n) cfma(d[i], c, vj[x]);
        }
    }
    double nr = 0;
#pragma unroll
    for (int i = 0; i < VEC_PER_THREAD; ++i) nr += cabs2(d[i]);
    nr = block_sum<VEC_THREADS>(nr, sm);
    if (threadIdx.x == 0) pconv[s * nchunk + blockIdx.x] = nr;
}

// stop rule (tdse.py:450,478-484): a state leaves the loop when !(conv > tol); reaching
// k == maxorder-1 raises even if that iteration converged.
__global__ void __launch_bounds__(32)
k_decide(const double* __restrict__ pconv, int nchunk, double tol, int k, int maxorder,
         int* __restrict__ active, int* __restrict__ order, int* __restrict__ ctrl) {
    const long long s = blockIdx.x;
    if (!active[s]) return;
    const double conv = warp_reduce_partials(pconv + s * nchunk, nchunk, 1);
    if (threadIdx.x == 0) {
        order[s] = k;
        if (k == maxorder - 1) {
            active[s] = 0;
            atomicExch(&ctrl[1], 1);
        } else if (!(conv > tol)) {
            active[s] = 0;
        } else {
            atomicAdd(&ctrl[0], 1);
        }
    }
}

// count states whose next beta is exactly zero (they need the Gram-Schmidt fallback)
__global__ void k_count_zero_beta(const double* __restrict__ beta, int bstride, int k,
                                  const int* __restrict__ active, int nstates, int* __restrict__ ctrl) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < nstates && active[s] && beta[(long long)s * bstride + k] == 0.0) atomicAdd(&ctrl[2], 1);
}

// psi = ph * sum_{i<=order} c_i V_i   (tdse.py:475,394)
__global__ void __launch_bounds__(VEC_THREADS)
k_combine(cplx* const* __restrict__ slabs, long long ldv, long long n, const cplx* __restrict__ ccur,
          int tstride, const int* __restrict__ order, const cplx* __restrict__ ph,
          cplx* __restrict__ psi, long long ld) {
    __shared__ double2 sc[MAX_ORDER_SMEM];
    const long long s = blockIdx.y;
    const int k = order[s];
    for (int i = threadIdx.x; i <= k; i += VEC_THREADS) sc[i] = ccur[s * tstride + i];
    __syncthreads();
    const long long base = (long long)blockIdx.x * VEC_CHUNK;
    cplx u[VEC_PER_THREAD];
#pragma unroll
    for (int i = 0; i < VEC_PER_THREAD; ++i) u[i] = make_double2(0.0, 0.0);
    for (int j = 0; j <= k; ++j) {
        const cplx* vj = slabs[j] + s * ldv;
        const cplx c = sc[j];
#pragma unroll
        for (int i = 0; i < VEC_PER_THREAD; ++i) {
            const long long x = base + threadIdx.x + i * VEC_THREADS;
            if (x < n) cfma(u[i], c, vj[x]);
        }
    }
#pragma unroll
    for (int i = 0; i < VEC_PER_THREAD; ++i) {
        const long long x = base + threadIdx.x + i * VEC_THREADS;
        if (x < n) psi[s * ld + x] = ph ? cmul(u[i], ph[x]) : u[i];
    }
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
k_reduce_dot(const cplx* __restrict__ pdot, int nchunk, cplx* __restrict__ out) {
    const long long s = blockIdx.x;
    const double* p = reinterpret_cast<const double*>(pdot + s * nchunk);
    const double re = warp_reduce_partials(p, nchunk, 2);
    const double im = warp_reduce_partials(p + 1, nchunk, 2);
    if (threadIdx.x == 0) out[s] = make_double2(re, im);
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

}  // namespace rmb
