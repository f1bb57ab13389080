// K2 (wide K blocks): y = sum_p (MF_p (x) K_p) x with the K contraction on the FP64 tensor pipe (DMMA).
//
// For asymmetric tops at high J the K factors are dense dim_k x dim_k blocks with dim_k of 25-50
// (AI 20-50 flop/B) and the contraction  Y[(s,m1), k1] += sum_k2 Z_p[(s,m1), k2] K_p[k1, k2]
// is a genuine GEMM: M = states x rows of the tile (64), N = dim_k of the bra block (<= 64), K = dim_k of
// the ket block, accumulated over the ~10-16 products of the bra block without leaving registers.
//
//  * Z_p = MF_p X (only the diagonals that survived the field contraction) is built ONCE per product and
//    k2-chunk in shared memory (no recomputation per column chunk as in the FMA kernel),
//  * mma.sync.aligned.m8n8k4.f64: each of the 8 warps owns one 8-row m-tile and all n-tiles; real and
//    imaginary parts of Z are two A operands against the same real B = K^T fragment,
//  * ket rows and MF diagonals arrive by TMA bulk copies (cp.async.bulk + mbarrier), K^T chunks are
//    host-built images in the bank-conflict-free fragment layout.
#pragma once
#include "rmb_matvec.cuh"

namespace rmb {

constexpr int MG_THREADS = 256;
constexpr int MG_M = 64;          // rows (state, m1) per CTA
constexpr int MG_KCH = 16;        // k2 chunk
constexpr int MG_LDZ = 20;        // leading dimension of the Z planes (== 4 mod 16: conflict-free A loads)
constexpr int MG_NTMAX = 8;       // n-tiles of 8 columns -> dim_k <= 64 per item

struct ItemG {
    long long bra_off;
    long long kt_off;        // offset (doubles) of the first product's K^T image
    int dk1, dm1;
    int r0, nrows;           // rows (m1) of the tile; nst * nrows <= MG_M
    int c0, nc;              // columns of the tile (nc <= 64)
    int nt;                  // n-tiles = ceil(nc / 8)
    int ldk;                 // leading dimension of the K^T images (== 4 mod 16)
    int p_begin, p_end;
    int nst;                 // states per CTA
    int desc_off;            // first ProdS descriptor
    int xbuf_elems;          // elements of the ket-row staging buffer
    int pad;
};

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NT>
__device__ __forceinline__ void mg_body(const ItemG& it, const ProdS* __restrict__ gdesc,
                                        const MfEntry* __restrict__ cent, const unsigned* __restrict__ tab_mask,
                                        const double* __restrict__ ktpool, const double2* __restrict__ X,
                                        double2* __restrict__ Y, long long ldx, long long ldy, int nstates, int s0,
                                        const int* __restrict__ active, const double* __restrict__ scale,
                                        int scale_stride, double2* __restrict__ pdot, int npart, int item_index,
                                        unsigned char* smem_raw) {
    // ---- shared memory carve-up
    double2* xbuf = reinterpret_cast<double2*>(smem_raw);
    MfEntry* mfe = reinterpret_cast<MfEntry*>(xbuf + it.xbuf_elems);
    double* zbuf = reinterpret_cast<double*>(mfe + MV2_NDMAX * it.nrows);  // [2][re|im][MG_M][MG_LDZ]
    double* kbuf = zbuf + 4 * MG_M * MG_LDZ;                               // [2][MG_KCH][ldk]
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(kbuf + 2 * MG_KCH * it.ldk);
    __shared__ int s_nnz;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int np = it.p_end - it.p_begin;
    const int mrows = it.nst * it.nrows;                                  // valid rows of the M tile
    // state validity of this CTA
    bool any = false;
    for (int s = 0; s < it.nst; ++s) {
        const int sg = s0 + s;
        any = any || (sg < nstates && (active == nullptr || active[sg]));
    }
    if (!any) return;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    double cre[NT][2], cim[NT][2];
#pragma unroll
    for (int n = 0; n < NT; ++n) cre[n][0] = cre[n][1] = cim[n][0] = cim[n][1] = 0.0;

    // Z build: a quarter warp (8 lanes) reads 8 consecutive k2 of ONE row (128 contiguous bytes: no bank
    // conflicts); each thread handles rows zrow0 and zrow0 + 32 and k2 offsets zk and zk + 8
    const int zrow0 = threadIdx.x >> 3, zk = threadIdx.x & 7;

    long long ktoff = it.kt_off;
    unsigned phase = 0;
    for (int ip = 0; ip < np; ++ip) {
        const ProdS d = gdesc[it.desc_off + ip];
        // ---- stage the ket rows of every state of the tile and the surviving MF diagonals (TMA)
        __syncthreads();                                   // previous product fully consumed
        if (threadIdx.x == 0) {
            const int nnz = min(__popc(tab_mask[d.tab]), MV2_NDMAX);
            s_nnz = nnz;
            const unsigned xbytes = (unsigned)(d.nr * d.xrs) * 16u;
            const unsigned mbytes = (unsigned)it.nrows * (unsigned)sizeof(MfEntry);
            int nact = 0;
            for (int s = 0; s < it.nst; ++s) {
                const int sg = s0 + s;
                nact += (sg < nstates && (active == nullptr || active[sg])) ? 1 : 0;
            }
            mbar_arrive_expect_tx(bar, (unsigned)nact * xbytes + (unsigned)nnz * mbytes);
            for (int s = 0; s < it.nst; ++s) {
                const int sg = s0 + s;
                if (xbytes && sg < nstates && (active == nullptr || active[sg]))
                    tma_load_1d(xbuf + (long long)s * d.nr * d.xrs, X + (long long)sg * ldx + d.ket_off, xbytes, bar);
            }
            for (int q = 0; q < nnz; ++q)
                tma_load_1d(mfe + q * it.nrows, cent + d.ent_off + (long long)q * it.dm1 + it.r0, mbytes, bar);
        }
        __syncthreads();                                   // s_nnz visible
        const int nnz = s_nnz;
        mbar_wait(bar, phase);
        phase ^= 1u;
        // MF rows of this thread's two Z rows
        double2 mf[2][MV2_NDMAX];
        int xo[2][MV2_NDMAX];
        const double2* xs[2];
        bool zvalid[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int zrow = zrow0 + 32 * h;
            const int zs = zrow / it.nrows, zm = zrow - zs * it.nrows;
            zvalid[h] = zrow < mrows;
            xs[h] = xbuf + (long long)zs * d.nr * d.xrs;
#pragma unroll
            for (int q = 0; q < MV2_NDMAX; ++q) {
                mf[h][q] = make_double2(0.0, 0.0);
                xo[h][q] = 0;
                if (q < nnz && zvalid[h]) {
                    const MfEntry e = mfe[q * it.nrows + zm];
                    if (e.col >= 0) {
                        mf[h][q] = make_double2(e.re, e.im);
                        xo[h][q] = (e.col - d.c_lo) * d.xrs;
                    }
                }
            }
        }
        const int nchunks = nnz > 0 ? (d.dk2 + MG_KCH - 1) / MG_KCH : 0;
        // stage chunk c: K^T chunk by cp.async (host image [dk2 padded to 16][ldk]), Z chunk by FMA
        auto stage_chunk = [&](int c) {
            double* kc_ = kbuf + (c & 1) * MG_KCH * it.ldk;
            const double* src = ktpool + ktoff + (long long)c * MG_KCH * it.ldk;
            const int n16 = MG_KCH * it.ldk / 2;                 // 16-byte pieces
            for (int i = threadIdx.x; i < n16; i += MG_THREADS)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(kc_ + 2 * i)), "l"(src + 2 * i));
            asm volatile("cp.async.commit_group;\n" ::);
            double* zr = zbuf + (c & 1) * 2 * MG_M * MG_LDZ;
            double* zi = zr + MG_M * MG_LDZ;
            const int k0 = c * MG_KCH;
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int k2 = k0 + zk + 8 * j;
                    double2 z = make_double2(0.0, 0.0);
                    if (zvalid[h] && k2 < d.dk2) {
#pragma unroll
                        for (int q = 0; q < MV2_NDMAX; ++q)
                            if (q < nnz) {
                                const double2 a = xs[h][xo[h][q] + k2];
                                z.x = fma(mf[h][q].x, a.x, z.x);
                                z.y = fma(mf[h][q].x, a.y, z.y);
                                z.x = fma(-mf[h][q].y, a.y, z.x);
                                z.y = fma(mf[h][q].y, a.x, z.y);
                            }
                    }
                    zr[(zrow0 + 32 * h) * MG_LDZ + zk + 8 * j] = z.x;
                    zi[(zrow0 + 32 * h) * MG_LDZ + zk + 8 * j] = z.y;
                }
        };
        if (nchunks > 0) {
            stage_chunk(0);
            asm volatile("cp.async.wait_group 0;\n" ::);
            __syncthreads();
        }
        for (int c = 0; c < nchunks; ++c) {
            // overlap: stage chunk c+1 (other buffers) while the tensor pipe works on chunk c
            if (c + 1 < nchunks) stage_chunk(c + 1);
            {
                const double* zr = zbuf + (c & 1) * 2 * MG_M * MG_LDZ;
                const double* are = zr + (warp * 8 + (lane >> 2)) * MG_LDZ + (lane & 3);
                const double* aim = are + MG_M * MG_LDZ;
                const double* bb = kbuf + (c & 1) * MG_KCH * it.ldk + (lane & 3) * it.ldk + (lane >> 2);
#pragma unroll
                for (int kk = 0; kk < MG_KCH / 4; ++kk) {
                    const double ar = are[kk * 4], ai = aim[kk * 4];
#pragma unroll
                    for (int n = 0; n < NT; ++n) {
                        const double b = bb[kk * 4 * it.ldk + n * 8];
                        dmma_m8n8k4(cre[n][0], cre[n][1], ar, b);
                        dmma_m8n8k4(cim[n][0], cim[n][1], ai, b);
                    }
                }
            }
            asm volatile("cp.async.wait_group 0;\n" ::);
            __syncthreads();       // chunk c consumed by every warp, chunk c+1 complete
        }
        ktoff += (long long)((d.dk2 + MG_KCH - 1) / MG_KCH) * MG_KCH * it.ldk;
    }
    // ---- epilogue: thread holds C[row = warp*8 + lane/4][cols 2*(lane%4), +1] of every n-tile
    const int row = warp * 8 + (lane >> 2);
    const int rs = row / it.nrows, rm = row - rs * it.nrows;
    const int sg = s0 + rs;
    const bool valid = row < mrows && sg < nstates && (active == nullptr || active[sg]);
    double pre = 0.0, pim = 0.0;
    if (valid) {
        const double sc = scale ? scale[(long long)sg * scale_stride] : 1.0;
        const long long row_off = it.bra_off + (long long)(it.r0 + rm) * (it.dk1 | 1) + it.c0;
#pragma unroll
        for (int n = 0; n < NT; ++n) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int c = n * 8 + 2 * (lane & 3) + j;
                if (c < it.nc) {
                    const double2 v = make_double2(cre[n][j] * sc, cim[n][j] * sc);
                    if (Y != nullptr) Y[(long long)sg * ldy + row_off + c] = v;
                    if (pdot != nullptr) {
                        const double2 x = X[(long long)sg * ldx + row_off + c];
                        pre += v.x * x.x + v.y * x.y;
                        pim += v.x * x.y - v.y * x.x;
                    }
                }
            }
        }
    }
    if (pdot != nullptr) {
        // reduce over the 4 lanes that share a row, then over the rows of each state (fixed order)
        pre += __shfl_xor_sync(0xffffffffu, pre, 1);
        pim += __shfl_xor_sync(0xffffffffu, pim, 1);
        pre += __shfl_xor_sync(0xffffffffu, pre, 2);
        pim += __shfl_xor_sync(0xffffffffu, pim, 2);
        __syncthreads();
        double* red = zbuf;                                  // [2][MG_M]
        if ((lane & 3) == 0) {
            red[row] = valid ? pre : 0.0;
            red[MG_M + row] = valid ? pim : 0.0;
        }
        __syncthreads();
        for (int s = warp; s < it.nst; s += MG_THREADS / 32) {
            const int sgs = s0 + s;
            if (sgs >= nstates || (active != nullptr && !active[sgs])) continue;
            double a = 0.0, b = 0.0;
            for (int r = lane; r < it.nrows; r += 32) {
                a += red[s * it.nrows + r];
                b += red[MG_M + s * it.nrows + r];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a += __shfl_down_sync(0xffffffffu, a, o);
                b += __shfl_down_sync(0xffffffffu, b, o);
            }
            if (lane == 0) pdot[(long long)sgs * npart + item_index] = make_double2(a, b);
        }
    }
}

__global__ void __launch_bounds__(MG_THREADS, 2)
k_matvec_gemm(const Unit2D* __restrict__ units, const ItemG* __restrict__ items,
              const ProdS* __restrict__ gdesc, const MfEntry* __restrict__ cent,
              const unsigned* __restrict__ tab_mask, const double* __restrict__ ktpool,
              const double2* __restrict__ X, double2* __restrict__ Y, long long ldx, long long ldy,
              int nstates, const int* __restrict__ active, const double* __restrict__ scale,
              int scale_stride, double2* __restrict__ pdot, int npart, int item_base) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const Unit2D u = units[blockIdx.x];
    const ItemG it = items[u.item];
#define RMB_GCASE(N)                                                                                      \
    case N:                                                                                               \
        mg_body<N>(it, gdesc, cent, tab_mask, ktpool, X, Y, ldx, ldy, nstates, u.s0, active, scale,       \
                   scale_stride, pdot, npart, item_base + u.item, smem_raw);                              \
        break;
    switch (it.nt) {
        RMB_GCASE(1) RMB_GCASE(2) RMB_GCASE(3) RMB_GCASE(4) RMB_GCASE(5) RMB_GCASE(6) RMB_GCASE(7) RMB_GCASE(8)
        default: break;
    }
#undef RMB_GCASE
}

}  // namespace rmb
