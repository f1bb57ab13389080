// K3 (fused, small Hilbert spaces): one whole split-operator Lanczos step per launch.
//
// One CTA per ensemble member.  The three live vectors of the recurrence (V_{k-1}, V_k, w) stay in shared
// memory for the whole step, the Krylov history V_0..V_k goes to a global (L2-resident) slab only for the
// convergence metric and the final combination, the small exponential exp(fac*T_k) e_0 is evaluated by one
// warp in shared memory, and every inner product is a fixed-order block reduction.  No host round trip
// per Lanczos iteration: this is the path for the reference's latency-bound cases (OCS alignment with a
// single state, 1 K ensembles with tens of states; examples/ocs_alignment.py, ocs_mixed_field.py).
// Restricted to operators whose (J,sym) blocks all have dim_k = 1 (linear rotors) and N*48 B of shared
// memory; everything else takes the batched path (rmb.cu: lanczos_batch).
//
// Arithmetic follows tdse.py:417-486 literally, including V_k = W_{k-1} / beta_k by division.
#pragma once
#include "rmb_kernels.cuh"
#include "rmb_matvec.cuh"
#include "rmb_matvec_lin.cuh"

namespace rmb {

constexpr int FUSED_THREADS = 512;

struct FusedArgs {
    long long n;                 // Hilbert-space dimension (= padded length: dim_k = 1 everywhere)
    const int* row_blk;          // [n] (J,sym) block of the row
    const int* blk_begin;        // [nblocks + 1] products of the block (sorted by bra)
    const long long* blk_off;    // [nblocks] offset of the block
    const int* blk_dm;           // [nblocks]
    const ProdD* prods;
    const unsigned* tab_mask;
    const MfEntry* cent;
    const double* kpool;
    int k_complex;
    int nprod;
    cplx* const* slabs;          // [maxorder + 1] history, each nstates * n
    cplx fac;
    double tol;
    int maxorder;
    const cplx* ph;              // H0 phase (may be null)
    cplx* psi;                   // in/out, state-major with leading dimension ld
    long long ld;
    int* order;                  // [nstates]
    int* ctrl;                   // [1] maxorder flag
    // merged entry lists of the sliding-window matvec (k_lin_entries), when the operator has them: per bra block the
    // surviving (ket block, diagonal) pairs with K * MF folded into one value per row
    const LinBlk* lin_blk;       // null: walk the products (fused_matvec)
    const LinEnt* lin_flat;      // [nblocks][ML_FLAT]; record 0 = header (count), LinEnt::pad = ket block
    const cplx* lin_val;
    int gram_kmax;               // iterations whose convergence metric uses the Gram diagonal (0: always explicit)
    int nblocks;
    int lin_lcap;                // > 0: the entry lists (<= lin_lcap entries per block) and the per-block table are staged in
                                 // shared memory behind the three vectors (fused_lin_table_bytes)
};

// entry of a staged list: offset of the ket block in the state vector, diagonal offset (col - row), rows of the ket block
struct __align__(16) FusedEnt { int koff, doff, dm2, pad; };
__host__ __device__ inline size_t fused_lin_table_bytes(int nblocks, int lcap) {
    return (size_t)nblocks * ((size_t)lcap * sizeof(FusedEnt) + 16);
}

template <int NT>
__device__ __forceinline__ double block_sum_all(double v, double* sm) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sm[wid] = v;
    __syncthreads();
    double r = 0;
#pragma unroll
    for (int i = 0; i < NT / 32; ++i) r += sm[i];
    return r;                     // same value in every thread, fixed order
}

// compact per-product descriptor kept in shared memory for the whole step
struct FusedProd {
    int ket_off;         // offset of the ket block in the state vector
    int ent_off;         // first compacted MF entry of the table
    int nnz;             // diagonals that survived the field contraction
    int pad;
    double kre, kim;     // the 1 x 1 K factor
};

// y = H x for one state, x and y in shared memory (dim_k = 1: H = sum_p k_p MF_p).  The <= 5 entries of a
// product are fetched with independent loads (memory-level parallelism: the entries stream from L2).
__device__ __forceinline__ void fused_matvec(const FusedArgs& a, const FusedProd* __restrict__ sp,
                                             const cplx* __restrict__ x, cplx* __restrict__ y) {
    for (long long i = threadIdx.x; i < a.n; i += FUSED_THREADS) {
        const int b = a.row_blk[i];
        const int m1 = (int)(i - a.blk_off[b]);
        const int dm1 = a.blk_dm[b];
        cplx acc = make_double2(0.0, 0.0);
        const int p1 = a.blk_begin[b + 1];
        for (int p = a.blk_begin[b]; p < p1; ++p) {
            const FusedProd pr = sp[p];
            const MfEntry* ep = a.cent + pr.ent_off + m1;
            MfEntry e[MV2_NDMAX];
#pragma unroll
            for (int q = 0; q < MV2_NDMAX; ++q)
                if (q < pr.nnz) e[q] = ep[(long long)q * dm1];
            cplx z = make_double2(0.0, 0.0);
#pragma unroll
            for (int q = 0; q < MV2_NDMAX; ++q)
                if (q < pr.nnz && e[q].col >= 0) {
                    const cplx v = x[pr.ket_off + e[q].col];
                    z.x = fma(e[q].re, v.x, z.x);
                    z.y = fma(e[q].re, v.y, z.y);
                    z.x = fma(-e[q].im, v.y, z.x);
                    z.y = fma(e[q].im, v.x, z.y);
                }
            for (int q = MV2_NDMAX; q < pr.nnz; ++q) {      // rank > 2 tensors: plain loop
                const MfEntry eq = ep[(long long)q * dm1];
                if (eq.col >= 0) cfma(z, make_double2(eq.re, eq.im), x[pr.ket_off + eq.col]);
            }
            acc.x = fma(pr.kre, z.x, acc.x);
            acc.y = fma(pr.kre, z.y, acc.y);
            acc.x = fma(-pr.kim, z.y, acc.x);
            acc.y = fma(pr.kim, z.x, acc.y);
        }
        y[i] = acc;
    }
}

constexpr int FUSED_ROWS = 9;      // rows per thread: n <= 210 KB / 48 B = 4375 <= 9 * 512

// block sum of two values at once (one pair of barriers); same value in every thread, fixed order
__device__ __forceinline__ double2 block_sum2_all(double x, double y, double2* sm /* FUSED_THREADS / 32 */) {
    x = warp_sum(x);
    y = warp_sum(y);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sm[wid] = make_double2(x, y);
    __syncthreads();
    double2 r = make_double2(0.0, 0.0);
#pragma unroll
    for (int i = 0; i < FUSED_THREADS / 32; ++i) { r.x += sm[i].x; r.y += sm[i].y; }
    return r;
}

// Second generation of the single-launch step.  Against the first one: the matvec walks the MERGED entry lists of
// k_lin_entries (one value per (entry, row) with K folded in, all loads of a row independent: one L2 latency per row
// instead of one per product), row -> block lookups are hoisted out of the iteration loop, alpha rides on the matvec
// pass, the convergence metric uses the Gram diagonal (as the batched path: rmb_lanczos.cuh) for the first
// `gram_kmax` iterations, so no pass over the Krylov history, and warp 0 evaluates the small exponential while the
// other warps already normalise V_{k+1}.  Five block reductions and three passes per iteration become two and three.
__global__ void __launch_bounds__(FUSED_THREADS, 1)
k_lanczos_fused(const FusedArgs a, long long nstates) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cplx* vk = reinterpret_cast<cplx*>(smem_raw);       // V_k
    cplx* vkm1 = vk + a.n;                              // V_{k-1}
    cplx* w = vkm1 + a.n;                               // H V_k, then W_k
    FusedProd* sp = reinterpret_cast<FusedProd*>(w + a.n);   // [nprod] (only without entry lists)
    __shared__ double2 red2[FUSED_THREADS / 32];
    __shared__ double red[FUSED_THREADS / 32];
    __shared__ double2 s_alpha[MAX_ORDER_SMEM];
    __shared__ double s_beta[MAX_ORDER_SMEM + 1], s_g[MAX_ORDER_SMEM + 1];
    __shared__ double2 s_c[MAX_ORDER_SMEM], s_dc[MAX_ORDER_SMEM];
    __shared__ double2 s_y[MAX_ORDER_SMEM], s_t1[MAX_ORDER_SMEM], s_t2[MAX_ORDER_SMEM];
    __shared__ double s_conv;
    const long long s = blockIdx.x;
    cplx* psi = a.psi + s * a.ld;
    const long long n = a.n;
    const bool lin = a.lin_blk != nullptr;
    // -DRMB_FUSED_TRACE: thread 0 of CTA 0 prints the cycles of every phase of the step (tools/ab_build.sh, RMB_LIB)
#ifdef RMB_FUSED_TRACE
    long long tr_t0 = clock64(), tr_mv = 0, tr_up = 0, tr_ex = 0, tr_hist = 0, tr_rot = 0, tr_last;
#define TR(acc) { const long long tr_now = clock64(); acc += tr_now - tr_last; tr_last = tr_now; }
#else
#define TR(acc)
#endif

    // rows of this thread: block and position inside it (fixed for the whole step)
    int rb[FUSED_ROWS], rm[FUSED_ROWS];
#pragma unroll
    for (int r = 0; r < FUSED_ROWS; ++r) {
        const long long i = threadIdx.x + (long long)r * FUSED_THREADS;
        rb[r] = 0;
        rm[r] = 0;
        if (i < n) {
            rb[r] = a.row_blk[i];
            rm[r] = (int)(i - a.blk_off[rb[r]]);
        }
    }
    // V_0 = psi * ph (not normalised, tdse.py:443 / 375); <V_0, V_0> for the Gram diagonal
    double g0 = 0.0;
    for (long long i = threadIdx.x; i < n; i += FUSED_THREADS) {
        cplx v = psi[i];
        if (a.ph) v = cmul(v, a.ph[i]);
        vk[i] = v;
        a.slabs[0][s * n + i] = v;
        g0 += cabs2(v);
    }
    const bool staged = lin && a.lin_lcap > 0;
    FusedEnt* s_ent = reinterpret_cast<FusedEnt*>(w + a.n);                          // [nblocks][lin_lcap]
    long long* s_voff = reinterpret_cast<long long*>(s_ent + (size_t)a.nblocks * a.lin_lcap);   // [nblocks]
    int* s_bdm = reinterpret_cast<int*>(s_voff + a.nblocks);                         // [nblocks]
    int* s_bL = s_bdm + a.nblocks;                                                   // [nblocks]
    if (staged) {
        for (int idx = threadIdx.x; idx < a.nblocks * a.lin_lcap; idx += FUSED_THREADS) {
            const int b = idx / a.lin_lcap, j = idx - b * a.lin_lcap;
            const LinEnt* fl = a.lin_flat + (size_t)b * ML_FLAT;
            FusedEnt f = {0, 0, 1, 0};
            if (j < (int)fl[0].xbyte) {
                const LinEnt g = fl[1 + j];
                f.koff = (int)a.blk_off[g.pad];
                f.doff = g.doff;
                f.dm2 = g.dm2;
            }
            s_ent[idx] = f;
        }
        for (int b = threadIdx.x; b < a.nblocks; b += FUSED_THREADS) {
            const LinBlk bt = a.lin_blk[b];
            s_voff[b] = bt.val_off;
            s_bdm[b] = bt.dm;
            s_bL[b] = min((int)a.lin_flat[(size_t)b * ML_FLAT].xbyte, a.lin_lcap);
        }
    }
    if (!lin)
        for (int p = threadIdx.x; p < a.nprod; p += FUSED_THREADS) {
            const ProdD pr = a.prods[p];
            FusedProd f;
            f.ket_off = (int)pr.ket_off;
            f.ent_off = (int)pr.ent_off;
            f.nnz = __popc(a.tab_mask[pr.tab]);
            f.pad = 0;
            if (a.k_complex) {
                const cplx kv = reinterpret_cast<const cplx*>(a.kpool)[pr.koff];
                f.kre = kv.x;
                f.kim = kv.y;
            } else {
                f.kre = a.kpool[pr.koff];
                f.kim = 0.0;
            }
            sp[p] = f;
        }
    g0 = block_sum_all<FUSED_THREADS>(g0, red);
    if (threadIdx.x == 0) { s_c[0] = make_double2(1.0, 0.0); s_beta[0] = 0.0; s_g[0] = g0; }
    __syncthreads();

    int k = 0, last = 0;
    bool hit_max = (a.maxorder <= 1);
#ifdef RMB_FUSED_TRACE
    const long long tr_init = clock64() - tr_t0;
    tr_last = clock64();
#endif
    for (;; ++k) {
        // ---- w = H V_k and alpha_k = vdot(w, V_k) in one pass
        double re = 0, im = 0;
        if (lin && staged) {
            // entry lists and block table in shared memory: the only global loads of a row are its entry values, eight
            // independent ones per trip (before: block record -> list header -> entry -> ket offset, four dependent L2 round
            // trips per trip of four entries, ~5.6 k cycles per row)
#pragma unroll
            for (int r = 0; r < FUSED_ROWS; ++r) {
                const long long i = threadIdx.x + (long long)r * FUSED_THREADS;
                if (i < n) {
                    const int b = rb[r], m = rm[r];
                    const int L = s_bL[b], dm = s_bdm[b];
                    const FusedEnt* ent = s_ent + b * a.lin_lcap;
                    const cplx* ev = a.lin_val + s_voff[b] + m;
                    cplx acc = make_double2(0.0, 0.0);
                    for (int j0 = 0; j0 < L; j0 += 8) {
                        cplx e[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            e[u] = make_double2(0.0, 0.0);
                            if (j0 + u < L) e[u] = ev[(long long)(j0 + u) * dm];
                        }
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            if (j0 + u < L) {
                                const FusedEnt f = ent[j0 + u];
                                const int col = min(max(m + f.doff, 0), f.dm2 - 1);   // outside the ket block e is zero
                                const cplx v = vk[f.koff + col];
                                acc.x = fma(e[u].x, v.x, acc.x);
                                acc.y = fma(e[u].x, v.y, acc.y);
                                acc.x = fma(-e[u].y, v.y, acc.x);
                                acc.y = fma(e[u].y, v.x, acc.y);
                            }
                        }
                    }
                    w[i] = acc;
                    const cplx y = vk[i];
                    re += acc.x * y.x + acc.y * y.y;
                    im += acc.x * y.y - acc.y * y.x;
                }
            }
        } else if (lin) {
#pragma unroll
            for (int r = 0; r < FUSED_ROWS; ++r) {
                const long long i = threadIdx.x + (long long)r * FUSED_THREADS;
                if (i < n) {
                    const int b = rb[r], m = rm[r];
                    const LinBlk bt = a.lin_blk[b];
                    const LinEnt* fl = a.lin_flat + (size_t)b * ML_FLAT;
                    const int L = (int)fl[0].xbyte;
                    const cplx* ev = a.lin_val + bt.val_off + m;
                    cplx acc = make_double2(0.0, 0.0);
                    for (int j0 = 0; j0 < L; j0 += 4) {                 // four independent (entry, ket) load pairs per trip
                        cplx e[4], v[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            e[u] = v[u] = make_double2(0.0, 0.0);
                            if (j0 + u < L) {
                                const LinEnt f = fl[1 + j0 + u];
                                e[u] = ev[(long long)(j0 + u) * bt.dm];
                                const int col = min(max(m + f.doff, 0), f.dm2 - 1);   // outside the ket block e is zero
                                v[u] = vk[a.blk_off[f.pad] + col];
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            acc.x = fma(e[u].x, v[u].x, acc.x);
                            acc.y = fma(e[u].x, v[u].y, acc.y);
                            acc.x = fma(-e[u].y, v[u].y, acc.x);
                            acc.y = fma(e[u].y, v[u].x, acc.y);
                        }
                    }
                    w[i] = acc;
                    const cplx y = vk[i];
                    re += acc.x * y.x + acc.y * y.y;
                    im += acc.x * y.y - acc.y * y.x;
                }
            }
        } else {
            fused_matvec(a, sp, vk, w);
            __syncthreads();
            for (long long i = threadIdx.x; i < n; i += FUSED_THREADS) {
                const cplx x = w[i], y = vk[i];
                re += x.x * y.x + x.y * y.y;
                im += x.x * y.y - x.y * y.x;
            }
        }
        const double2 al = block_sum2_all(re, im, red2);
        TR(tr_mv)
        const cplx alpha = make_double2(al.x, al.y);
        const double beta = s_beta[k];
        if (threadIdx.x == 0) s_alpha[k] = alpha;
        // ---- W_k = w - alpha V_k - beta V_{k-1}; norm
        double nr = 0;
        for (long long i = threadIdx.x; i < n; i += FUSED_THREADS) {
            cplx r = csub(w[i], cmul(alpha, vk[i]));
            if (k > 0) { r.x -= beta * vkm1[i].x; r.y -= beta * vkm1[i].y; }
            w[i] = r;
            nr += cabs2(r);
        }
        nr = block_sum_all<FUSED_THREADS>(nr, red);
        const double beta_next = sqrt(nr);
        if (threadIdx.x == 0) {
            s_beta[k + 1] = beta_next;
            s_g[k + 1] = (beta_next != 0.0) ? nr / (beta_next * beta_next) : 1.0;      // <V_{k+1}, V_{k+1}>
        }
        __syncthreads();
        TR(tr_up)
        const bool use_gram = k < a.gram_kmax;
        bool done = false;
        if (k > 0) {
            // c^k = expm(fac T_k) e_0 ; conv = sum |sum_i (c^k_i - c^{k-1}_i) V_i|^2 (tdse.py:474-476)
            if (threadIdx.x < 32) {
                warp_expm_col0(k + 1, s_alpha, s_beta, a.fac, s_y, s_t1, s_t2);
                double cv = 0.0;
                for (int i = threadIdx.x; i <= k; i += 32) {
                    const cplx prev = (i < k) ? s_c[i] : make_double2(0.0, 0.0);
                    const cplx d = csub(s_y[i], prev);
                    s_dc[i] = d;
                    s_c[i] = s_y[i];
                    cv += cabs2(d) * s_g[i];
                }
                cv = warp_sum(cv);
                if (threadIdx.x == 0) s_conv = cv;                  // Gram-diagonal form of the metric
                TR(tr_ex)
            }
            if (!use_gram) {
                // explicit evaluation over the Krylov history (many vectors: orthogonality is no longer a given)
                __syncthreads();
                double cv = 0;
                for (long long i = threadIdx.x; i < n; i += FUSED_THREADS) {
                    cplx d = make_double2(0.0, 0.0);
                    for (int j = 0; j + 2 <= k; ++j) cfma(d, s_dc[j], a.slabs[j][s * n + i]);
                    cfma(d, s_dc[k - 1], vkm1[i]);
                    cfma(d, s_dc[k], vk[i]);
                    cv += cabs2(d);
                }
                cv = block_sum_all<FUSED_THREADS>(cv, red);
                if (threadIdx.x == 0) s_conv = cv;
            }
        }
        // ---- history: V_{k+1} = W_k / beta_{k+1} (tdse.py:455-456).  No barrier since the exponential started: the other
        //      warps write the slab while warp 0 is still busy with it.  Written even if this turns out to be the last
        //      iteration (nobody reads it then).
        const bool fallback = beta_next == 0.0;
        if (!fallback)
            for (long long i = threadIdx.x; i < n; i += FUSED_THREADS) {
                const cplx r = w[i];
                const cplx q = make_double2(r.x / beta_next, r.y / beta_next);
                a.slabs[k + 1][s * n + i] = q;
                w[i] = q;                                  // becomes V_{k+1} by the pointer rotation below
            }
        __syncthreads();
        TR(tr_hist)
        if (k > 0) {
            last = k;
            if (k == a.maxorder - 1) { hit_max = true; done = true; }
            else if (!(s_conv > a.tol)) done = true;
        } else if (a.maxorder <= 1) {
            done = true;
        }
        if (done) break;
        // rotate: V_{k-1} <- V_k, V_k <- V_{k+1}
        if (!fallback) {
            cplx* const t = vkm1;                          // no copies: the buffer of V_{k-1} takes the next H V
            vkm1 = vk;
            vk = w;
            w = t;
        } else {
            // zero-beta fallback (tdse.py:459-465): Gram-Schmidt of the all-ones vector against V_0..V_k
            for (long long i = threadIdx.x; i < n; i += FUSED_THREADS) {
                vkm1[i] = vk[i];
                w[i] = make_double2(1.0, 0.0);
            }
            __syncthreads();
            for (int j = 0; j <= k; ++j) {
                double pr = 0, pi = 0;
                for (long long i = threadIdx.x; i < n; i += FUSED_THREADS) {
                    const cplx x = (j == k) ? vkm1[i] : a.slabs[j][s * n + i], y = w[i];
                    pr += x.x * y.x + x.y * y.y;
                    pi += x.x * y.y - x.y * y.x;
                }
                pr = block_sum_all<FUSED_THREADS>(pr, red);
                pi = block_sum_all<FUSED_THREADS>(pi, red);
                const cplx proj = make_double2(pr, pi);
                for (long long i = threadIdx.x; i < n; i += FUSED_THREADS) {
                    const cplx x = (j == k) ? vkm1[i] : a.slabs[j][s * n + i];
                    w[i] = csub(w[i], cmul(proj, x));
                }
                __syncthreads();
            }
            double nv = 0;
            for (long long i = threadIdx.x; i < n; i += FUSED_THREADS) nv += cabs2(w[i]);
            nv = sqrt(block_sum_all<FUSED_THREADS>(nv, red));
            for (long long i = threadIdx.x; i < n; i += FUSED_THREADS) {
                const cplx v = make_double2(w[i].x / nv, w[i].y / nv);
                vk[i] = v;
                a.slabs[k + 1][s * n + i] = v;
            }
        }
        __syncthreads();
        TR(tr_rot)
    }
    // psi = ph * u_k,  u_k = sum_i c_i V_i  (u_0 = V_0 when the loop never ran)
    __syncthreads();
    for (long long i = threadIdx.x; i < n; i += FUSED_THREADS) {
        cplx u = make_double2(0.0, 0.0);
        if (last == 0) {
            u = a.slabs[0][s * n + i];
        } else {
            for (int j = 0; j + 2 <= last; ++j) cfma(u, s_c[j], a.slabs[j][s * n + i]);
            cfma(u, s_c[last - 1], vkm1[i]);
            cfma(u, s_c[last], vk[i]);
        }
        psi[i] = a.ph ? cmul(u, a.ph[i]) : u;
    }
    if (threadIdx.x == 0) {
        a.order[s] = last;
        if (hit_max) atomicExch(a.ctrl, 1);
    }
#ifdef RMB_FUSED_TRACE
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const long long tr_end = clock64();
        printf("[fused] n %lld order %d cycles: total %lld init %lld matvec+alpha %lld update+norm %lld expm(warp0) %lld hist+wait %lld "
               "rotate %lld final %lld\n", n, last, tr_end - tr_t0, tr_init, tr_mv, tr_up, tr_ex, tr_hist, tr_rot, tr_end - tr_last);
    }
#endif
}

}  // namespace rmb
