// Lanczos vector kernels, second generation (tdse.py:417-486 semantics unchanged):
//   * sliced grids: a CTA walks `cps` consecutive 1024-element chunks of one state (fewer, longer CTAs; one block
//     reduction per CTA instead of one per chunk; tens instead of hundreds of partials per state);
//   * k_recur_gram: ONE launch per Lanczos iteration after the matvec.  Every CTA reduces the <w, V_k> partials of
//     the matvec epilogue itself (alpha_k, bit-identical in all CTAs of a state), forms W_k = w - alpha_k V_k -
//     beta_k V_{k-1}, and the LAST CTA of each state (atomic ticket) runs the per-state scalar part: beta_{k+1},
//     expm(fac T_k) e_0, the convergence metric and the stop rule (what k_small_a / k_small_b did in two extra
//     launches), plus the zero-beta fallback;
//   * the convergence metric sum |u_k - u_{k-1}|^2 = || sum_i dc_i V_i ||^2 (tdse.py:475-476) is evaluated through
//     the diagonal of the Gram matrix of the Krylov vectors, sum_i |dc_i|^2 <V_i, V_i>, with <V_0,V_0> accumulated in
//     the k = 0 pass and <V_i,V_i> = |W_{i-1}|^2 rinv_i^2: O(k) per iteration instead of re-reading the whole Krylov
//     history (O(k N)).  The neglected terms are the off-diagonal Gram entries 2 Re conj(dc_i) dc_j <V_i,V_j>, i.e. the
//     loss of orthogonality of the three-term recurrence (<= 1e-13 relative for the first ~15 vectors); beyond
//     RMB_GRAM_KMAX iterations the driver switches back to the explicit evaluation (k_small_a / k_recur_conv /
//     k_small_b), so strongly driven cases with tens of vectors keep the reference's stop decisions.
#pragma once
#include "rmb_kernels.cuh"

namespace rmb {

constexpr int RMB_GRAM_KMAX = 16;

// deterministic block-wide sum of `npart` complex partials (stride 1): thread t adds partials t, t+NT, ... in order,
// then the fixed tree of block_sum.  Result valid in thread 0.
__device__ __forceinline__ cplx block_reduce_partials(const cplx* __restrict__ p, int npart, double* sm) {
    double re = 0.0, im = 0.0;
    for (int i = threadIdx.x; i < npart; i += VEC_THREADS) {
        const cplx v = p[i];
        re += v.x;
        im += v.y;
    }
    re = block_sum<VEC_THREADS>(re, sm);
    im = block_sum<VEC_THREADS>(im, sm);
    return make_double2(re, im);
}

__global__ void __launch_bounds__(VEC_THREADS, 4)
k_recur_gram(const cplx* __restrict__ w, cplx* const* __restrict__ slabs, long long ldv, long long n,
             const cplx* __restrict__ pdot, int npart, cplx* __restrict__ alpha, double* __restrict__ beta,
             double* __restrict__ rinv, double* __restrict__ gdiag, int tstride, int bstride, int k, cplx fac,
             cplx* __restrict__ ccur, cplx* __restrict__ ceff, double* __restrict__ pnrm, double* __restrict__ pg0,
             int nsl, int cps, int nchunk, unsigned* __restrict__ ticket, double tol, int maxorder,
             int* __restrict__ active, int* __restrict__ order, int* __restrict__ ctrl,
             const int* __restrict__ pmap, long long nuser) {
    __shared__ double sm[VEC_THREADS / 32];
    __shared__ double2 s_alpha;
    __shared__ int s_last, s_fallback;
    __shared__ double2 ey[MAX_ORDER_SMEM], eterm[MAX_ORDER_SMEM], etmp[MAX_ORDER_SMEM];
    __shared__ double2 bc;
    const long long s = blockIdx.y;
    if (!active[s]) return;
    // ---- alpha_k = <w, V_k> (tdse.py:445,468): partial sums conj(w) * slab_k of the matvec epilogue, times rinv_k
    const double rk = rinv[s * bstride + k];
    {
        const cplx a = block_reduce_partials(pdot + s * npart, npart, sm);
        if (threadIdx.x == 0) s_alpha = make_double2(a.x * rk, a.y * rk);
    }
    __syncthreads();
    const cplx a = s_alpha;
    const double b = (k > 0) ? beta[s * bstride + k] : 0.0;
    const double rkm1 = (k > 0) ? rinv[s * bstride + k - 1] : 0.0;
    const cplx* sk = slabs[k] + s * ldv;
    const cplx* skm1 = (k > 0) ? slabs[k - 1] + s * ldv : nullptr;
    const cplx* ws = w + s * ldv;
    cplx* out = slabs[k + 1] + s * ldv;
    // ---- W_k = w - alpha_k V_k - beta_k V_{k-1} (tdse.py:446,469-470) over this CTA's chunks
    double nr = 0.0, g0 = 0.0;
    const int c_end = min(nchunk, ((int)blockIdx.x + 1) * cps);
    for (int c = blockIdx.x * cps; c < c_end; ++c) {
        const long long base = (long long)c * VEC_CHUNK;
        cplx tw[VEC_PER_THREAD], tk[VEC_PER_THREAD], tm[VEC_PER_THREAD];
#pragma unroll
        for (int i = 0; i < VEC_PER_THREAD; ++i) {
            const long long x = base + threadIdx.x + i * VEC_THREADS;
            tw[i] = tk[i] = tm[i] = make_double2(0.0, 0.0);
            if (x < n) {
                tw[i] = ws[x];
                tk[i] = sk[x];
                if (k > 0) tm[i] = skm1[x];
            }
        }
#pragma unroll
        for (int i = 0; i < VEC_PER_THREAD; ++i) {
            const long long x = base + threadIdx.x + i * VEC_THREADS;
            if (x < n) {
                const cplx vk = make_double2(tk[i].x * rk, tk[i].y * rk);
                cplx r = csub(tw[i], cmul(a, vk));
                if (k > 0) {
                    r.x -= b * (tm[i].x * rkm1);
                    r.y -= b * (tm[i].y * rkm1);
                } else {
                    g0 += cabs2(tk[i]);
                }
                out[x] = r;
                nr += cabs2(r);
            }
        }
    }
    nr = block_sum<VEC_THREADS>(nr, sm);
    if (k == 0) g0 = block_sum<VEC_THREADS>(g0, sm);
    if (threadIdx.x == 0) {
        pnrm[s * nsl + blockIdx.x] = nr;
        if (k == 0) pg0[s * nsl + blockIdx.x] = g0;
        __threadfence();
        const unsigned old = atomicAdd(&ticket[s], 1u);
        s_last = (old == (unsigned)nsl - 1u) ? 1 : 0;
        s_fallback = 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // ---- the last CTA of the state: per-state scalar part of iteration k
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        const double nrs = warp_reduce_partials(pnrm + s * nsl, nsl, 1);
        double g00 = 0.0;
        if (k == 0) g00 = warp_reduce_partials(pg0 + s * nsl, nsl, 1);
        const double bn = sqrt(nrs);
        const double rn = (bn != 0.0) ? 1.0 / bn : 1.0;
        if (lane == 0) {
            alpha[s * tstride + k] = a;
            beta[s * bstride + k + 1] = bn;
            rinv[s * bstride + k + 1] = rn;
            gdiag[s * bstride + k + 1] = (bn != 0.0) ? nrs * rn * rn : 1.0;     // <V_{k+1}, V_{k+1}>
            if (k == 0) gdiag[s * bstride] = g00;                                 // <V_0, V_0>
            ticket[s] = 0u;
        }
        __syncwarp();
        double conv = 1.0;
        if (k == 0) {
            if (lane == 0) {
                ccur[s * tstride] = make_double2(1.0, 0.0);   // u_0 = V_0
                ceff[s * tstride] = make_double2(1.0, 0.0);
            }
        } else {
            // c^k = expm(fac T_k)[:,0], dc = c^k - c^{k-1} (tdse.py:474-475)
            warp_expm_col0(k + 1, alpha + s * tstride, beta + s * bstride, fac, ey, eterm, etmp);
            double cv = 0.0;
            for (int i = lane; i <= k; i += 32) {
                const cplx prev = (i < k) ? ccur[s * tstride + i] : make_double2(0.0, 0.0);
                const double ri = rinv[s * bstride + i];
                const cplx d = csub(ey[i], prev);
                cv += cabs2(d) * gdiag[s * bstride + i];
                ceff[s * tstride + i] = make_double2(ey[i].x * ri, ey[i].y * ri);
                ccur[s * tstride + i] = ey[i];
            }
            cv = warp_sum(cv);
            conv = __shfl_sync(0xffffffffu, cv, 0);
        }
        if (lane == 0) {
            // stop rule (tdse.py:450,478-484): a state leaves the loop when !(conv > tol); reaching k == maxorder-1
            // raises even if that iteration converged
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
            s_fallback = (still && bn == 0.0) ? 1 : 0;
        }
    }
    __syncthreads();
    if (!s_fallback) return;
    // ---- zero-beta fallback (tdse.py:459-465): V_{k+1} = normalised Gram-Schmidt of the all-ones vector against
    //      V_0..V_k (V_j = slab_j * rinv_j), written to slab k+1 with rinv = 1; pad elements stay zero
    cplx* v = slabs[k + 1] + s * ldv;
    for (long long x = threadIdx.x; x < nuser; x += VEC_THREADS) v[pmap[x]] = make_double2(1.0, 0.0);
    __syncthreads();
    for (int j = 0; j <= k; ++j) {
        const cplx* vj = slabs[j] + s * ldv;
        const double rj = rinv[s * bstride + j];
        double pr = 0, pi = 0;   // proj = vdot(V_j, v) = sum conj(V_j) * v
        for (long long x = threadIdx.x; x < nuser; x += VEC_THREADS) {
            const int xp = pmap[x];
            const cplx aa = make_double2(vj[xp].x * rj, vj[xp].y * rj), bb = v[xp];
            pr += aa.x * bb.x + aa.y * bb.y;
            pi += aa.x * bb.y - aa.y * bb.x;
        }
        pr = block_sum<VEC_THREADS>(pr, sm);
        pi = block_sum<VEC_THREADS>(pi, sm);
        if (threadIdx.x == 0) bc = make_double2(pr, pi);
        __syncthreads();
        const cplx proj = bc;
        for (long long x = threadIdx.x; x < nuser; x += VEC_THREADS) {
            const int xp = pmap[x];
            const cplx aa = make_double2(vj[xp].x * rj, vj[xp].y * rj);
            v[xp] = csub(v[xp], cmul(proj, aa));
        }
        __syncthreads();
    }
    double nv2 = 0;
    for (long long x = threadIdx.x; x < nuser; x += VEC_THREADS) nv2 += cabs2(v[pmap[x]]);
    nv2 = block_sum<VEC_THREADS>(nv2, sm);
    if (threadIdx.x == 0) bc = make_double2(sqrt(nv2), 0.0);
    __syncthreads();
    const double nv = bc.x;
    for (long long x = threadIdx.x; x < nuser; x += VEC_THREADS) {
        const int xp = pmap[x];
        v[xp] = make_double2(v[xp].x / nv, v[xp].y / nv);
    }
}

// V0 = psi * ph  (tdse.py:375), or a plain copy when ph == nullptr; scatters into the padded layout.  Sliced grid.
__global__ void __launch_bounds__(VEC_THREADS)
k_phase_init_s(const cplx* __restrict__ psi, long long ld, const cplx* __restrict__ ph, cplx* __restrict__ V0,
               long long ldv, long long n, const int* __restrict__ pmap, int cps, int nchunk) {
    const long long s = blockIdx.y;
    const int c_end = min(nchunk, ((int)blockIdx.x + 1) * cps);
    for (int c = blockIdx.x * cps; c < c_end; ++c) {
        const long long base = (long long)c * VEC_CHUNK;
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
}

// psi = ph * sum_{i<=order} c_i V_i   (tdse.py:475,394), c_i V_i = ceff_i * slab_i.  Sliced grid.
__global__ void __launch_bounds__(VEC_THREADS, 4)
k_combine_s(cplx* const* __restrict__ slabs, long long ldv, long long n, const cplx* __restrict__ ceff, int tstride,
            const int* __restrict__ order, const cplx* __restrict__ ph, cplx* __restrict__ psi, long long ld,
            const int* __restrict__ pmap, int cps, int nchunk) {
    __shared__ double2 sc[MAX_ORDER_SMEM];
    const long long s = blockIdx.y;
    const int k = order[s];
    for (int i = threadIdx.x; i <= k; i += VEC_THREADS) sc[i] = ceff[s * tstride + i];
    __syncthreads();
    const int c_end = min(nchunk, ((int)blockIdx.x + 1) * cps);
    for (int c = blockIdx.x * cps; c < c_end; ++c) {
        const long long base = (long long)c * VEC_CHUNK;
        cplx u[VEC_PER_THREAD];
        int xp[VEC_PER_THREAD];
#pragma unroll
        for (int i = 0; i < VEC_PER_THREAD; ++i) {
            const long long x = base + threadIdx.x + i * VEC_THREADS;
            u[i] = make_double2(0.0, 0.0);
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
            const cplx cc = sc[j];
#pragma unroll
            for (int i = 0; i < VEC_PER_THREAD; ++i) {
                const long long x = base + threadIdx.x + i * VEC_THREADS;
                if (x < n) cfma(u[i], cc, vj[xp[i]]);
            }
        }
#pragma unroll
        for (int i = 0; i < VEC_PER_THREAD; ++i) {
            const long long x = base + threadIdx.x + i * VEC_THREADS;
            if (x < n) psi[s * ld + x] = ph ? cmul(u[i], ph[x]) : u[i];
        }
    }
}

}  // namespace rmb
