// K2 (wide K blocks), second generation: y = sum_p (MF_p (x) K_p) x with the K contraction on the FP64 tensor
// pipe (DMMA, mma.sync.aligned.m8n8k4.f64) and NO shared-memory round trip for Z = MF_p X.
//
//   Y[(s,m1), k1] += sum_k2 Z_p[(s,m1), k2] K_p[k1, k2],      Z_p[(s,m1), k2] = sum_q MF_p[m1,q] X_p[(s, m1+off_q), k2]
//
// Work unit = (bra block, row tile of 8*mt rows, column tile of <= 64 columns) x (nst states), nst*mt <= 12:
// every one of the 12 MMA warps owns ONE 8-row m-tile of one state and all n-tiles, with the accumulators in
// registers across all products of the bra block.  What makes it fast:
//   * the A fragment is built in registers: lane (row f = lane/4, k = lane%4) of an m8n8k4 tile needs exactly one
//     element of Z; it forms it from the staged ket rows with nnz LDS.128 + 4 nnz DFMA (nnz = surviving MF diagonals,
//     coefficients and row offsets in registers for the whole product).  Z is never written anywhere; the MMA warps
//     never synchronise with each other inside the product loop.
//   * conflict-free ket-row loads for any (odd) row stride: fragment rows (2i, 2i+1) are mapped to tile rows
//     (i, i+4), so the two rows a quarter-warp touches are 4 rows = 4*stride = 4 (mod 8) 16-byte slots apart.
//   * a producer warp stages product p+1 (ket rows of every state, MF diagonals, the host-built K^T image: one bulk
//     copy each, cp.async.bulk + mbarrier complete_tx) while the MMA warps work on product p (full/empty ring); a CTA
//     walks several state tiles of its item, so the ring also runs across tiles (no start-up bubble per tile).
//   * K^T images are padded to multiples of 4 in k2 (not 16) and 8 in k1; leading dimension == 4 (mod 16) makes the B
//     fragment loads conflict-free.
#pragma once
#include "rmb_matvec.cuh"

namespace rmb {

constexpr int MD_MMA_WARPS = 12;
constexpr int MD_THREADS = (MD_MMA_WARPS + 1) * 32;   // + one producer warp
constexpr int MD_STAGES = 2;                          // products in flight: minimum ...
constexpr int MD_STAGES_MAX = 4;                      // ... and maximum per item (ItemD2::nstages: as many as fit the shared memory)
constexpr int MD_NTMAX = 8;                           // n-tiles of 8 columns -> <= 64 columns per item
constexpr size_t MD_SMEM_MAX = 226 * 1024;
constexpr size_t MD_SMEM_DEEP = 222 * 1024;           // budget for stages beyond MD_STAGES (227 KB per CTA minus the static arrays)

struct ItemD2 {
    long long bra_off;
    long long kt_off;        // offset (doubles) of the first product's K^T image
    int dk1, dm1;
    int r0, nrows;           // rows (m1) of the tile: nrows <= 8 * mt
    int c0, nc;              // columns of the tile (nc <= 64)
    int nt;                  // n-tiles = ceil(nc / 8)
    int ldk;                 // leading dimension of the K^T images (== 4 mod 16)
    int p_begin, p_end;
    int nst;                 // states per CTA
    int mt;                  // m-tiles per state (nst * mt <= MD_MMA_WARPS)
    int desc_off;            // first ProdS descriptor
    int x_elems;             // double2 elements of the ket-row area of one stage
    int kt_doubles;          // doubles of the K^T area of one stage
    int nstages;             // pipeline stages of this item (MD_STAGES .. MD_STAGES_MAX)
    int pslot;               // first <w,v> partial slot of the item: one slot per m-tile (slots pslot .. pslot + mt - 1 of a state)
    int pad2;
};

// stage layout: [X: x_elems double2][MF: MV2_NDMAX * 8*mt MfEntry][K^T: kt_doubles double]
__host__ __device__ inline size_t md_stage_bytes(int x_elems, int mt, int kt_doubles) {
    return (size_t)x_elems * 16 + (size_t)MV2_NDMAX * 8 * mt * sizeof(MfEntry) + (size_t)kt_doubles * 8;
}
__host__ __device__ inline size_t md_smem_bytes(int x_elems, int mt, int kt_doubles, int nprod, int nstages = MD_STAGES) {
    return nstages * md_stage_bytes(x_elems, mt, kt_doubles) + (size_t)nprod * sizeof(ProdS) +
           (size_t)4 * MD_MMA_WARPS * 8 + 2 * MD_STAGES_MAX * 8 + 128;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// k2 loop of one product for one m-tile: NNZ surviving MF diagonals (compile-time: no predicated-off loads or DFMAs).
// The A fragment of k-step kk+1 is formed BETWEEN the DMMAs of step kk, one diagonal per n-tile, so that the dependent
// DFMA chain hides behind the tensor work of the same warp (DMMA and DFMA share the FP64 datapath).
template <int NT, int NNZ>
__device__ __forceinline__ void md_kloop(const double2* __restrict__ xs, const int (&xo)[MV2_NDMAX],
                                         const double2 (&mf)[MV2_NDMAX], const double* __restrict__ bb, int ldk,
                                         int dk2, int kq, double (&cre)[NT][2], double (&cim)[NT][2]) {
    const int nkk = (dk2 + 3) >> 2;
    const int kmax = dk2 - 1;
    double ar = 0.0, ai = 0.0;
    {
        const int k2 = min(kq, kmax);
#pragma unroll
        for (int q = 0; q < NNZ; ++q) {
            const double2 a = xs[xo[q] + k2];
            ar = fma(mf[q].x, a.x, ar);
            ai = fma(mf[q].x, a.y, ai);
            ar = fma(-mf[q].y, a.y, ar);
            ai = fma(mf[q].y, a.x, ai);
        }
    }
    for (int kk = 0; kk < nkk; ++kk) {
        // ket elements of the next k-step (rows k2 >= dk2 of K^T are zero: the clamp only keeps the address inside the
        // staged rows)
        const int k2n = min(kk * 4 + 4 + kq, kmax);
        double2 an[NNZ];
#pragma unroll
        for (int q = 0; q < NNZ; ++q) an[q] = xs[xo[q] + k2n];
        const double* bk = bb + kk * 4 * ldk;
        double zr = 0.0, zi = 0.0;
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            const double b = bk[n * 8];
            dmma884(cre[n][0], cre[n][1], ar, b);
            dmma884(cim[n][0], cim[n][1], ai, b);
            if (n < NNZ) {
                zr = fma(mf[n].x, an[n].x, zr);
                zi = fma(mf[n].x, an[n].y, zi);
                zr = fma(-mf[n].y, an[n].y, zr);
                zi = fma(mf[n].y, an[n].x, zi);
            }
        }
#pragma unroll
        for (int q = NT; q < NNZ; ++q) {
            zr = fma(mf[q].x, an[q].x, zr);
            zi = fma(mf[q].x, an[q].y, zi);
            zr = fma(-mf[q].y, an[q].y, zr);
            zi = fma(mf[q].y, an[q].x, zi);
        }
        ar = zr;
        ai = zi;
    }
}

template <int NT>
__device__ __forceinline__ void md_body(const ItemD2& it, const ProdS* __restrict__ gdesc,
                                        const MfEntry* __restrict__ cent, const double* __restrict__ ktpool,
                                        const double2* __restrict__ X, double2* __restrict__ Y, long long ldx,
                                        long long ldy, int nstates, int s_first, int ntiles,
                                        const int* __restrict__ active, const double* __restrict__ scale,
                                        int scale_stride, double2* __restrict__ pdot, int npart, int item_index) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int np = it.p_end - it.p_begin;
    const size_t stage_bytes = md_stage_bytes(it.x_elems, it.mt, it.kt_doubles);
    unsigned char* stage0 = rmb_dsmem;
    const int NSTG = it.nstages;
    ProdS* sp = reinterpret_cast<ProdS*>(rmb_dsmem + NSTG * stage_bytes);
    double* red = reinterpret_cast<double*>(sp + np);                                 // [2 (tile parity)][2][MD_MMA_WARPS]
    unsigned long long* full = reinterpret_cast<unsigned long long*>(red + 4 * MD_MMA_WARPS);
    unsigned long long* empty = full + MD_STAGES_MAX;
    __shared__ int s_valid[MV2_TILES_MAX][MD_MMA_WARPS];      // state s of tile t is present and active
    __shared__ int s_nact[MV2_TILES_MAX];

    // A CTA walks `ntiles` consecutive state tiles of one item: the descriptors, barriers and the producer's
    // pipeline are set up once, and the first products of tile t+1 are staged while tile t is being finished.
    for (int i = threadIdx.x; i < ntiles * MD_MMA_WARPS; i += MD_THREADS) {
        const int t = i / MD_MMA_WARPS, s = i - t * MD_MMA_WARPS;
        const int sg = s_first + t * it.nst + s;
        s_valid[t][s] = (s < it.nst && sg < nstates && (active == nullptr || active[sg])) ? 1 : 0;
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTG; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], MD_MMA_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    for (int i = threadIdx.x; i < np * (int)(sizeof(ProdS) / 16); i += MD_THREADS)
        reinterpret_cast<int4*>(sp)[i] = reinterpret_cast<const int4*>(gdesc + it.desc_off)[i];
    __syncthreads();
    if (threadIdx.x < ntiles) {
        int n = 0;
        for (int s = 0; s < it.nst; ++s) n += s_valid[threadIdx.x][s];
        s_nact[threadIdx.x] = n;
    }
    __syncthreads();

    if (warp == MD_MMA_WARPS) {
        // ================= producer warp =================
        int stage = 0, ph = 0, fill = 0;
        for (int t = 0; t < ntiles; ++t) {
            const int nact = s_nact[t];
            if (nact == 0) continue;
            const int s0 = s_first + t * it.nst;
            long long ktoff = it.kt_off;
            for (int ip = 0; ip < np; ++ip) {
                const ProdS d = sp[ip];
                const int k2p = (d.dk2 + 3) & ~3;
                const unsigned ktbytes = (unsigned)(k2p * it.ldk) * 8u;
                if (d.nnz > 0 && d.nr > 0) {
                    if (fill >= NSTG) mbar_wait(&empty[stage], (unsigned)(ph ^ 1));   // consumers released the stage
                    unsigned char* st = stage0 + (size_t)stage * stage_bytes;
                    double2* xst = reinterpret_cast<double2*>(st);
                    MfEntry* mfe = reinterpret_cast<MfEntry*>(st + (size_t)it.x_elems * 16);
                    double* kst = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(mfe) +
                                                            (size_t)MV2_NDMAX * 8 * it.mt * sizeof(MfEntry));
                    const unsigned xbytes = (unsigned)(d.nr * d.xrs) * 16u;
                    const unsigned mbytes = (unsigned)it.nrows * (unsigned)sizeof(MfEntry);
                    if (lane == 0)
                        mbar_arrive_expect_tx(&full[stage], (unsigned)nact * xbytes + (unsigned)d.nnz * mbytes + ktbytes);
                    __syncwarp();
                    if (lane < it.nst) {
                        if (s_valid[t][lane])
                            tma_load_1d(xst + (size_t)lane * d.nr * d.xrs, X + (long long)(s0 + lane) * ldx + d.ket_off,
                                        xbytes, &full[stage]);
                    } else if (lane < it.nst + d.nnz) {
                        const int q = lane - it.nst;
                        tma_load_1d(mfe + q * 8 * it.mt, cent + d.ent_off + (long long)q * it.dm1 + it.r0, mbytes,
                                    &full[stage]);
                    } else if (lane == 31) {
                        tma_load_1d(kst, ktpool + ktoff, ktbytes, &full[stage]);
                    }
                    ++fill;
                    if (++stage == NSTG) { stage = 0; ph ^= 1; }
                }
                ktoff += (long long)k2p * it.ldk;
            }
        }
        return;
    }

    // ================= MMA warps =================
    const int ls = warp / it.mt;                      // state of this warp inside the tile
    const int rb = (warp - ls * it.mt) * 8;           // first tile row of its m-tile
    const int f = lane >> 2, kq = lane & 3;
    const int arow = rb + (f >> 1) + 4 * (f & 1);     // tile row held by this lane's fragment row (see header)
    int stage = 0, ph = 0;
    for (int t = 0; t < ntiles; ++t) {
        if (s_nact[t] == 0) continue;
        const int s0 = s_first + t * it.nst;
        const bool wvalid = ls < it.nst && s_valid[t][ls < it.nst ? ls : 0];
        const bool rvalid = wvalid && arow < it.nrows;

        double cre[NT][2], cim[NT][2];
#pragma unroll
        for (int n = 0; n < NT; ++n) cre[n][0] = cre[n][1] = cim[n][0] = cim[n][1] = 0.0;

        for (int ip = 0; ip < np; ++ip) {
            const ProdS d = sp[ip];
            if (!(d.nnz > 0 && d.nr > 0)) continue;
            mbar_wait(&full[stage], (unsigned)ph);
            if (wvalid) {
                const unsigned char* st = stage0 + (size_t)stage * stage_bytes;
                const double2* xs = reinterpret_cast<const double2*>(st) + (size_t)ls * d.nr * d.xrs;
                const MfEntry* mfe = reinterpret_cast<const MfEntry*>(st + (size_t)it.x_elems * 16);
                const double* kst = reinterpret_cast<const double*>(reinterpret_cast<const unsigned char*>(mfe) +
                                                                    (size_t)MV2_NDMAX * 8 * it.mt * sizeof(MfEntry));
                // MF coefficients and ket-row offsets of this lane's row, for the whole product
                double2 mf[MV2_NDMAX];
                int xo[MV2_NDMAX];
#pragma unroll
                for (int q = 0; q < MV2_NDMAX; ++q) {
                    mf[q] = make_double2(0.0, 0.0);
                    xo[q] = 0;
                    if (q < d.nnz && rvalid) {
                        const MfEntry e = mfe[q * 8 * it.mt + arow];
                        if (e.col >= 0) {
                            mf[q] = make_double2(e.re, e.im);
                            xo[q] = (e.col - d.c_lo) * d.xrs;
                        }
                    }
                }
                const double* bb = kst + kq * it.ldk + f;
                switch (d.nnz) {
                    case 1: md_kloop<NT, 1>(xs, xo, mf, bb, it.ldk, d.dk2, kq, cre, cim); break;
                    case 2: md_kloop<NT, 2>(xs, xo, mf, bb, it.ldk, d.dk2, kq, cre, cim); break;
                    case 3: md_kloop<NT, 3>(xs, xo, mf, bb, it.ldk, d.dk2, kq, cre, cim); break;
                    case 4: md_kloop<NT, 4>(xs, xo, mf, bb, it.ldk, d.dk2, kq, cre, cim); break;
                    default: md_kloop<NT, 5>(xs, xo, mf, bb, it.ldk, d.dk2, kq, cre, cim); break;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            if (++stage == NSTG) { stage = 0; ph ^= 1; }
        }

        // ---- epilogue: lane holds C[tile row arow][cols 2*kq, 2*kq+1] of every n-tile
        const int sg = s0 + ls;
        double pre = 0.0, pim = 0.0;
        if (rvalid) {
            const double sc = scale ? scale[(long long)sg * scale_stride] : 1.0;
            const long long row_off = it.bra_off + (long long)(it.r0 + arow) * (it.dk1 | 1) + it.c0;
#pragma unroll
            for (int n = 0; n < NT; ++n) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int c = n * 8 + 2 * kq + j;
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
            // fixed-order reduction over the lanes of the m-tile; every MMA warp writes the partial of its own (state, m-tile)
            // -- no barrier between the warps of a CTA at the end of a tile (it stalled the product ring once per tile:
            // the launch with this epilogue took 8.2 ms against 6.0 ms without on the dim_k = 25 workload)
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                pre += __shfl_xor_sync(0xffffffffu, pre, o);
                pim += __shfl_xor_sync(0xffffffffu, pim, o);
            }
            if (lane == 0 && wvalid) pdot[(long long)sg * npart + it.pslot + (warp - ls * it.mt)] = make_double2(pre, pim);
        }
    }
}

__global__ void __launch_bounds__(MD_THREADS, 1)
k_matvec_dmma(const Unit2D* __restrict__ units, const ItemD2* __restrict__ items,
              const ProdS* __restrict__ gdesc, const MfEntry* __restrict__ cent,
              const double* __restrict__ ktpool, const double2* __restrict__ X, double2* __restrict__ Y,
              long long ldx, long long ldy, int nstates, const int* __restrict__ active,
              const double* __restrict__ scale, int scale_stride, double2* __restrict__ pdot, int npart,
              int item_base) {
    const Unit2D u = units[blockIdx.x];
    const ItemD2 it = items[u.item];
#define RMB_DCASE(N)                                                                                      \
    case N:                                                                                               \
        md_body<N>(it, gdesc, cent, ktpool, X, Y, ldx, ldy, nstates, u.s0, u.ntiles, active, scale,        \
                   scale_stride, pdot, npart, item_base + u.item);                                        \
        break;
    switch (it.nt) {
        RMB_DCASE(1) RMB_DCASE(2) RMB_DCASE(3) RMB_DCASE(4) RMB_DCASE(5) RMB_DCASE(6) RMB_DCASE(7) RMB_DCASE(8)
        default: break;
    }
#undef RMB_DCASE
}

}  // namespace rmb
