// K2 (tiled version): y = sum_p (MF_p (x) K_p) x for one work unit = (bra-block tile, one or more state tiles).
//
// The CTA is two warpgroups.  Consumers: one thread owns one row m1 of the bra block for TWO states and all
// (<= 12) columns k1 of the tile; accumulators live in registers.  Per block product p
//   * the producer warpgroup stages the ket rows the tile needs and the surviving MF diagonals in shared memory
//     with TMA bulk copies (cp.async.bulk) through a full / empty mbarrier ring of 2-6 stages; the internal
//     vectors store rows of (dim_k | 1) elements, so a tile's ket rows are one contiguous run per state and the
//     row-strided reads below are bank-conflict free,
//   * z = sum_j MF_p[m1, j] * X[col_j, k2] is formed on the fly in registers (zero MF diagonals are
//     skipped: after the field contraction most of the (2w+1) diagonals vanish for polarised fields; for real
//     MF, i.e. fields in the XZ plane, half of the products are skipped as well),
//   * acc[k1] += K_p[k1, k2] * z with the K_p^T row loaded once from shared memory for both states.
// H(t) itself is never materialised: the kernel only sees the MF and K factors.
#pragma once
#include "rmb_internal.h"

namespace rmb {

constexpr int MV2_CONSUMERS = 128;               // compute threads: one row of TWO states each
constexpr int MV2_PRODUCERS = 4;                 // producer warps issuing the TMA bulk copies (state s -> warp s % 4)
constexpr int MV2_THREADS = MV2_CONSUMERS + 32 * MV2_PRODUCERS;   // two warpgroups: consumers, producers
// Register split between the warpgroups (setmaxnreg): the CTA is launched with 128 registers per thread
// (256 threads, two CTAs per SM); the producers keep 40 and the consumers grow to 208, which holds the
// 2 x 12 complex accumulators, a K^T row and both sets of ket elements without spilling.
constexpr int MV2_REGS_PRODUCER = 40;
constexpr int MV2_REGS_CONSUMER = 208;
constexpr int MV2_NCMAX = 12;     // columns (k1) per thread
constexpr int MV2_NDMAX = 5;      // max ELL width handled by the tiled kernel (rank <= 2)
constexpr int MV2_SMAX = 32;      // max states per CTA
constexpr int MV2_STAGES = 2;     // products in flight: minimum (sizes the state tile) ...
constexpr int MV2_STAGES_MAX = 6; // ... and maximum per item (Item2D::nstages: as many as fit the shared-memory budget)

struct Item2D {
    long long bra_off;
    int dk1, dm1;
    int r0, nrows;       // rows (m1) of the tile, nrows <= MV2_CONSUMERS
    int c0, nc;          // columns (k1) of the tile, nc <= MV2_NCMAX
    int p_begin, p_end;
    int nst;             // states per CTA, even (compute threads = nst/2 * nrows <= MV2_CONSUMERS)
    int xr_off;          // offset into the per-(item, product) ket row ranges
    int kt_total;        // doubles of K^T staged in shared memory for this item (even)
    int xbuf_elems;      // elements of one staging buffer: max over products of nst * nr * (dk2 | 1)
    int desc_off;        // first static descriptor (ProdS) of the item in the global descriptor table
    int nstages;         // pipeline stages of this item (MV2_STAGES .. MV2_STAGES_MAX)
    long long kt_off;    // offset (doubles) of the item's K^T image in the global K^T pool
};

struct XRange { int c_lo, nr; };   // ket rows [c_lo, c_lo + nr) needed by (item, product)
struct Unit2D { int item, s0, ntiles, pad; };   // `ntiles` consecutive state tiles (tiled kernel; 1 for the DMMA kernel)
constexpr int MV2_TILES_MAX = 8;       // state tiles one CTA of the tiled kernel may walk
constexpr int MV2_RED_BYTES = 4 * MV2_CONSUMERS * 8 + MV2_TILES_MAX * 4;   // <w,V_k> partials [2][nst][nrows] + tile flags

// compacted MF entry (written by k_compact_tables): value and ket m index of one surviving diagonal
struct __align__(32) MfEntry { double re, im; int col; int pad[3]; };

// per-product descriptor staged in shared memory at CTA start
// (built on the host; `nnz` and `mreal` are refreshed on the device after every field update: k_fill_nnz)
struct ProdS {
    long long ket_off;   // padded offset of the first staged ket row (c_lo already added)
    long long ent_off;   // first compacted entry of the MF table
    int dk2, nnz, c_lo, nr, xrs, tab, mreal, pad2;   // nnz, mreal (all surviving MF entries real): k_fill_nnz
};

// After a field update: surviving diagonals per (item, product) descriptor, so that the descriptors a CTA copies to
// shared memory are complete (no dependent gather from the diagonal masks at CTA start)
__global__ void k_fill_nnz(int n, ProdS* __restrict__ gdesc, const unsigned* __restrict__ tab_mask,
                           const unsigned* __restrict__ tab_cplx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int t = gdesc[i].tab;
        gdesc[i].nnz = min(__popc(tab_mask[t]), MV2_NDMAX);
        gdesc[i].mreal = tab_cplx[t] == 0u ? 1 : 0;
    }
}

// ---- mbarrier / TMA (cp.async.bulk) helpers -----------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy (TMA), completion signalled on an mbarrier with the byte count
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// dynamic shared memory of the matvec kernels.  Staged data is addressed as `rmb_dsmem + byte offset` so that the
// compiler sees the shared address space (LDS instead of generic loads).
extern __shared__ __align__(128) unsigned char rmb_dsmem[];

#ifndef RMB_MV2_PREFETCH_K
#define RMB_MV2_PREFETCH_K 1      // K^T rows one k2 ahead in registers when they fit (the ket elements: see PX)
#endif
#ifndef RMB_MV2_REGS
#define RMB_MV2_REGS 180          // registers the inner loop may plan with (the rest: addresses, loop state)
#endif

// acc[k1] += K[k1,k2] * z for the two states of a thread; `krow` is row k2 of the K^T image in registers
template <int NC, bool KC>
__device__ __forceinline__ void mv2_kstage(const double (&krow)[KC ? 2 * NC : NC], const double2& zA, const double2& zB,
                                           double2 (&accA)[NC], double2 (&accB)[NC]) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        if (KC) {
            const double kx = krow[2 * c], ky = krow[2 * c + 1];
            accA[c].x = fma(kx, zA.x, accA[c].x);
            accA[c].y = fma(kx, zA.y, accA[c].y);
            accB[c].x = fma(kx, zB.x, accB[c].x);
            accB[c].y = fma(kx, zB.y, accB[c].y);
            accA[c].x = fma(-ky, zA.y, accA[c].x);
            accA[c].y = fma(ky, zA.x, accA[c].y);
            accB[c].x = fma(-ky, zB.y, accB[c].x);
            accB[c].y = fma(ky, zB.x, accB[c].y);
        } else {
            const double kv = krow[c];
            accA[c].x = fma(kv, zA.x, accA[c].x);
            accA[c].y = fma(kv, zA.y, accA[c].y);
            accB[c].x = fma(kv, zB.x, accB[c].x);
            accB[c].y = fma(kv, zB.y, accB[c].y);
        }
    }
}

// row k2 of the K^T image ([k2][NC] doubles, or [k2][NC] double2 for complex K) -> registers
template <int NC, bool KC>
__device__ __forceinline__ void mv2_load_krow(unsigned kt_byte, int k2, double (&krow)[KC ? 2 * NC : NC]) {
    constexpr int ND = KC ? 2 * NC : NC;
    if (ND == 1) {
        krow[0] = *reinterpret_cast<const double*>(rmb_dsmem + kt_byte + (unsigned)k2 * 8u);
    } else {
        const double2* src = reinterpret_cast<const double2*>(rmb_dsmem + kt_byte + (unsigned)(k2 * ND) * 8u);
#pragma unroll
        for (int c = 0; c < ND / 2; ++c) {
            const double2 v = src[c];
            krow[2 * c] = v.x;
            krow[2 * c + 1] = v.y;
        }
    }
}

// one block product for one thread (one row m1 of two states A, B):
//   acc[k1] += sum_k2 K[k1,k2] * (sum_q MF[q] * X[row_q, k2])
// K^T values are loaded once per k2 and used for both states (halves the shared-memory traffic per DFMA).
// Software pipeline over k2: the ket elements of column k2 + 1 (and, for few diagonals, row k2 + 1 of K^T) are
// loaded into a second register set before the DFMAs of column k2, so the LDS latency of one column hides behind
// the 4 * NC + 8 * NNZ DFMAs of the previous one (two consumer warps per scheduler are not enough to hide it).
template <int NC, int NNZ, bool KC, bool MR>
__device__ __forceinline__ void mv2_inner(unsigned xa_byte, unsigned xb_byte, unsigned mfe_byte,
                                          int nrows, int c_lo, int xrs, unsigned kt_byte, int dk2,
                                          double2 (&accA)[NC], double2 (&accB)[NC]) {
    constexpr int ND = KC ? 2 * NC : NC;
    // register budget (168 per thread at two CTAs of 192 threads per SM): accumulators 8 NC, a K^T row 2 ND,
    // ket elements, MF values and offsets 13 NNZ (+ 8 NNZ for the second set)
    constexpr bool PK = RMB_MV2_PREFETCH_K && 8 * NC + 4 * ND + 21 * NNZ <= RMB_MV2_REGS;
    constexpr bool PX = 8 * NC + (PK ? 4 : 2) * ND + 21 * NNZ <= RMB_MV2_REGS;
    double2 mf[NNZ];
    unsigned xo[NNZ];
#pragma unroll
    for (int q = 0; q < NNZ; ++q) {
        const unsigned char* ep = rmb_dsmem + mfe_byte + (unsigned)(q * nrows) * (unsigned)sizeof(MfEntry);
        mf[q] = *reinterpret_cast<const double2*>(ep);
        const int col = *reinterpret_cast<const int*>(ep + 16);
        xo[q] = col >= 0 ? (unsigned)((col - c_lo) * xrs) * 16u : 0u;   // value is 0 when the diagonal leaves the block
    }
    double2 a[NNZ], b[NNZ];
    double krow[ND];
#pragma unroll
    for (int q = 0; q < NNZ; ++q) {
        a[q] = *reinterpret_cast<const double2*>(rmb_dsmem + xa_byte + xo[q]);
        b[q] = *reinterpret_cast<const double2*>(rmb_dsmem + xb_byte + xo[q]);
    }
    if (PK) mv2_load_krow<NC, KC>(kt_byte, 0, krow);
#pragma unroll 2
    for (int k2 = 0; k2 < dk2; ++k2) {
        const int kn = min(k2 + 1, dk2 - 1);                 // the last column re-loads itself (no branch)
        double2 an[NNZ], bn[NNZ];
        double krown[ND];
        if (!PK) mv2_load_krow<NC, KC>(kt_byte, k2, krow);
        if (PX) {
#pragma unroll
            for (int q = 0; q < NNZ; ++q) {
                an[q] = *reinterpret_cast<const double2*>(rmb_dsmem + xa_byte + xo[q] + (unsigned)kn * 16u);
                bn[q] = *reinterpret_cast<const double2*>(rmb_dsmem + xb_byte + xo[q] + (unsigned)kn * 16u);
            }
        }
        if (PK) mv2_load_krow<NC, KC>(kt_byte, kn, krown);
        double2 zA = make_double2(0.0, 0.0), zB = make_double2(0.0, 0.0);
#pragma unroll
        for (int q = 0; q < NNZ; ++q) {
            zA.x = fma(mf[q].x, a[q].x, zA.x);
            zA.y = fma(mf[q].x, a[q].y, zA.y);
            zB.x = fma(mf[q].x, b[q].x, zB.x);
            zB.y = fma(mf[q].x, b[q].y, zB.y);
            if (!MR) {                                       // MR: every surviving MF entry of the product is real
                zA.x = fma(-mf[q].y, a[q].y, zA.x);
                zA.y = fma(mf[q].y, a[q].x, zA.y);
                zB.x = fma(-mf[q].y, b[q].y, zB.x);
                zB.y = fma(mf[q].y, b[q].x, zB.y);
            }
        }
        mv2_kstage<NC, KC>(krow, zA, zB, accA, accB);
        if (PX) {
#pragma unroll
            for (int q = 0; q < NNZ; ++q) { a[q] = an[q]; b[q] = bn[q]; }
        } else {
#pragma unroll
            for (int q = 0; q < NNZ; ++q) {
                a[q] = *reinterpret_cast<const double2*>(rmb_dsmem + xa_byte + xo[q] + (unsigned)kn * 16u);
                b[q] = *reinterpret_cast<const double2*>(rmb_dsmem + xb_byte + xo[q] + (unsigned)kn * 16u);
            }
        }
        if (PK) {
#pragma unroll
            for (int c = 0; c < ND; ++c) krow[c] = krown[c];
        }
    }
}

struct Mv2Smem {
    unsigned xbuf_b, xstride;    // ket staging buffers: byte offset of stage 0 in rmb_dsmem, bytes per stage
    unsigned mfe_b, mstride;     // MF diagonal staging buffers
    unsigned kt_b, sp_b, red_b;  // K^T image, descriptors; red: [2][nst][nrows] partial dots, then int flags[MV2_TILES_MAX]
    int nstages;
    double* kt;
    ProdS* sp;
    unsigned long long* full;    // [nstages] data of a product has landed (TMA transaction bytes)
    unsigned long long* empty;   // [nstages] every consumer warp is done with the stage
    unsigned long long* setup;   // K^T image + descriptors have landed
};

// Consumer warpgroup (4 warps, 208 registers per thread after setmaxnreg): walks the state tiles of the unit; per
// tile it zeroes the accumulators, consumes the products of the item in pipeline order (full wait -> mv2_inner ->
// empty arrive) and runs the epilogue (scale, store, fused <w,V_k> partials through a consumer-only named barrier).
template <int NC, bool KC>
__device__ __forceinline__ void mv2_consumer(const Item2D& it, const double2* __restrict__ X,
                                             double2* __restrict__ Y, long long ldx, long long ldy, int nstates,
                                             int s0, int ntiles, const int* __restrict__ active, const Mv2Smem& sm,
                                             const double* __restrict__ scale, int scale_stride,
                                             double2* __restrict__ pdot, int npart, int item_index,
                                             int sl, int rl, bool inb) {
    constexpr int KW = KC ? 2 : 1;                       // doubles per K element
    const int np = it.p_end - it.p_begin, nrows = it.nrows, nst = it.nst, nc = it.nc;
    const long long row_off = it.bra_off + (long long)(it.r0 + rl) * (it.dk1 | 1) + it.c0;
    const ProdS* sp = reinterpret_cast<const ProdS*>(rmb_dsmem + sm.sp_b);
    const int* tflag = reinterpret_cast<const int*>(rmb_dsmem + sm.red_b + 4 * MV2_CONSUMERS * 8);
    mbar_wait(sm.setup, 0);
    int stage = 0, phase = 0;                            // pipeline position of the next product
    for (int t = 0; t < ntiles; ++t, s0 += nst) {
        if (!tflag[t]) continue;                         // every state of the tile has converged
        const int stA = s0 + 2 * sl, stB = stA + 1;
        const bool vA = inb && stA < nstates && (active == nullptr || active[stA]);
        const bool vB = inb && stB < nstates && (active == nullptr || active[stB]);
        const bool work = vA || vB;
        double2 accA[NC], accB[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) accA[c] = accB[c] = make_double2(0.0, 0.0);
        int ktbase = 0;
        for (int ip = 0; ip < np; ++ip) {
            mbar_wait(&sm.full[stage], (unsigned)phase);
            const int dk2 = sp[ip].dk2, nnz = sp[ip].nnz;
            if (work) {
                const int nr = sp[ip].nr, xrs = sp[ip].xrs, c_lo = sp[ip].c_lo;
                const unsigned xa = sm.xbuf_b + (unsigned)stage * sm.xstride + (unsigned)((2 * sl) * nr * xrs) * 16u;
                const unsigned xb = xa + (unsigned)(nr * xrs) * 16u;   // an inactive partner reads stale data: never stored
                const unsigned mfe = sm.mfe_b + (unsigned)stage * sm.mstride + (unsigned)rl * (unsigned)sizeof(MfEntry);
                const unsigned ktp = sm.kt_b + (unsigned)ktbase * 8u;
#define RMB_NNZ(MRV)                                                                                              \
                switch (nnz) {                                                                                     \
                    case 0: break;                                                                                 \
                    case 1: mv2_inner<NC, 1, KC, MRV>(xa, xb, mfe, nrows, c_lo, xrs, ktp, dk2, accA, accB); break; \
                    case 2: mv2_inner<NC, 2, KC, MRV>(xa, xb, mfe, nrows, c_lo, xrs, ktp, dk2, accA, accB); break; \
                    case 3: mv2_inner<NC, 3, KC, MRV>(xa, xb, mfe, nrows, c_lo, xrs, ktp, dk2, accA, accB); break; \
                    case 4: mv2_inner<NC, 4, KC, MRV>(xa, xb, mfe, nrows, c_lo, xrs, ktp, dk2, accA, accB); break; \
                    default: mv2_inner<NC, 5, KC, MRV>(xa, xb, mfe, nrows, c_lo, xrs, ktp, dk2, accA, accB); break; \
                }
                if (sp[ip].mreal) { RMB_NNZ(true) } else { RMB_NNZ(false) }
#undef RMB_NNZ
            }
            ktbase += dk2 * NC * KW;
            __syncwarp();
            if ((threadIdx.x & 31) == 0) mbar_arrive(&sm.empty[stage]);   // this warp is done with the stage
            if (++stage == sm.nstages) { stage = 0; phase ^= 1; }
        }
        // ---- epilogue: optional per-state scale (w = rinv_k * H slab_k), store, fused partial dot
        //      sum conj(w) * x over the rows of this tile (alpha of the Lanczos recurrence, tdse.py:468)
        double preA = 0.0, pimA = 0.0, preB = 0.0, pimB = 0.0;
        auto finish = [&](double2 (&acc)[NC], int st, double& pre, double& pim) {
            if (scale != nullptr) {
                const double sc = scale[(long long)st * scale_stride];
#pragma unroll
                for (int c = 0; c < NC; ++c) { acc[c].x *= sc; acc[c].y *= sc; }
            }
            if (Y != nullptr) {
                double2* y = Y + (long long)st * ldy + row_off;
#pragma unroll
                for (int c = 0; c < NC; ++c)
                    if (c < nc) y[c] = acc[c];
            }
            if (pdot != nullptr) {
                const double2* x = X + (long long)st * ldx + row_off;
#pragma unroll
                for (int c = 0; c < NC; ++c)
                    if (c < nc) {
                        const double2 v = x[c];
                        pre += acc[c].x * v.x + acc[c].y * v.y;
                        pim += acc[c].x * v.y - acc[c].y * v.x;
                    }
            }
        };
        if (vA) finish(accA, stA, preA, pimA);
        if (vB) finish(accB, stB, preB, pimB);
        if (pdot != nullptr) {
            // consumer warps only (named barrier 1; the producers are already staging the next tile).  The first
            // barrier keeps a fast warp from overwriting the partials of the previous tile while they are read
            asm volatile("bar.sync 1, %0;\n" ::"n"(MV2_CONSUMERS) : "memory");
            double* red = reinterpret_cast<double*>(rmb_dsmem + sm.red_b);   // [2][nst][nrows]
            const int half = nst * nrows;
            if (inb) {
                red[(2 * sl) * nrows + rl] = preA;
                red[(2 * sl + 1) * nrows + rl] = preB;
                red[half + (2 * sl) * nrows + rl] = pimA;
                red[half + (2 * sl + 1) * nrows + rl] = pimB;
            }
            asm volatile("bar.sync 1, %0;\n" ::"n"(MV2_CONSUMERS) : "memory");
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
            for (int s = warp; s < nst; s += MV2_CONSUMERS / 32) {
                const int sg = s0 + s;
                if (sg >= nstates || (active != nullptr && !active[sg])) continue;
                double a = 0.0, b = 0.0;
                for (int r = lane; r < nrows; r += 32) {
                    a += red[s * nrows + r];
                    b += red[half + s * nrows + r];
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
}

// producer warpgroup: TMA bulk copies, MV2_STAGES products ahead of the consumers
__device__ __forceinline__ void mv2_producer(const Item2D& it, const ProdS* __restrict__ gdesc,
                                             const MfEntry* __restrict__ cent, const unsigned* __restrict__ tab_mask,
                                             const double* __restrict__ ktpool, const double2* __restrict__ X,
                                             long long ldx, int nstates, int s0, int ntiles,
                                             const int* __restrict__ active, const Mv2Smem& sm) {
    const int np = it.p_end - it.p_begin;
    const int pw = (threadIdx.x - MV2_CONSUMERS) >> 5;    // producer warp: copies the states s % MV2_PRODUCERS == pw
    const int lane = threadIdx.x & 31;
    const ProdS* sp = reinterpret_cast<const ProdS*>(rmb_dsmem + sm.sp_b);
    if (pw == 0) {
        // the K^T image of all products and the static descriptors were laid out on the host exactly as
        // they sit in shared memory: two bulk copies
        if (lane == 0) {
            const unsigned kbytes = (unsigned)it.kt_total * 8u, dbytes = (unsigned)np * (unsigned)sizeof(ProdS);
            mbar_arrive_expect_tx(sm.setup, kbytes + dbytes);
            if (kbytes) tma_load_1d(sm.kt, ktpool + it.kt_off, kbytes, sm.setup);
            if (dbytes) tma_load_1d(sm.sp, gdesc + it.desc_off, dbytes, sm.setup);
        }
    }
    const int* tflag = reinterpret_cast<const int*>(rmb_dsmem + sm.red_b + 4 * MV2_CONSUMERS * 8);
    const int sidx = lane * MV2_PRODUCERS + pw;           // state of the tile handled by this lane
    mbar_wait(sm.setup, 0);
    int stage = 0, phase = 1;                             // next stage to fill; parity of its `empty` barrier
    for (int t = 0; t < ntiles; ++t, s0 += it.nst) {
        if (!tflag[t]) continue;
        long long sb = -1;
        if (sidx < it.nst) {
            const int s = s0 + sidx;
            if (s < nstates && (active == nullptr || active[s])) sb = (long long)s * ldx;
        }
        const int nact = __popc(__ballot_sync(0xffffffffu, sb >= 0));
        for (int ip = 0; ip < np; ++ip) {
            // (a fresh barrier passes a wait on parity 1: the first round over the stages does not block)
            mbar_wait(&sm.empty[stage], (unsigned)phase);
            const ProdS d = sp[ip];
            const int nnz = pw == 0 ? d.nnz : 0;             // warp 0 also brings the MF diagonals (k_fill_nnz)
            const unsigned xbytes = (unsigned)(d.nr * d.xrs) * 16u;
            const unsigned mbytes = (unsigned)it.nrows * (unsigned)sizeof(MfEntry);
            if (lane == 0) mbar_arrive_expect_tx(&sm.full[stage], (unsigned)nact * xbytes + (unsigned)nnz * mbytes);
            __syncwarp();
            if (xbytes > 0 && sb >= 0)
                tma_load_1d(rmb_dsmem + sm.xbuf_b + (unsigned)stage * sm.xstride + (unsigned)(sidx * d.nr * d.xrs) * 16u,
                            X + sb + d.ket_off, xbytes, &sm.full[stage]);
            if (it.nrows == it.dm1) {
                // the tile covers every row of the bra block: the surviving diagonals are one contiguous run
                if (lane == 0 && nnz > 0)
                    tma_load_1d(rmb_dsmem + sm.mfe_b + (unsigned)stage * sm.mstride, cent + d.ent_off, (unsigned)nnz * mbytes,
                                &sm.full[stage]);
            } else if (lane < nnz) {
                tma_load_1d(rmb_dsmem + sm.mfe_b + (unsigned)stage * sm.mstride + (unsigned)(lane * it.nrows) * (unsigned)sizeof(MfEntry),
                            cent + d.ent_off + (long long)lane * it.dm1 + it.r0, mbytes, &sm.full[stage]);
            }
            if (++stage == sm.nstages) { stage = 0; phase ^= 1; }
        }
    }
}

template <bool KC>
__global__ void __launch_bounds__(MV2_THREADS, 2)
k_matvec_tiled(const Unit2D* __restrict__ units, const Item2D* __restrict__ items,
               const ProdS* __restrict__ gdesc, const MfEntry* __restrict__ cent,
               const unsigned* __restrict__ tab_mask, const double* __restrict__ ktpool,
               const double2* __restrict__ X, double2* __restrict__ Y, long long ldx, long long ldy,
               int nstates, const int* __restrict__ active, const double* __restrict__ scale,
               int scale_stride, double2* __restrict__ pdot, int npart) {
    unsigned char* const smem_raw = rmb_dsmem;
    const Unit2D u = units[blockIdx.x];
    const Item2D it = items[u.item];
    const int np = it.p_end - it.p_begin;
    // shared memory is carved with the sizes of this item (the launch reserves the maximum over items)
    Mv2Smem sm;
    unsigned char* p = smem_raw;
    sm.nstages = it.nstages;
    sm.xbuf_b = 0;
    sm.xstride = (unsigned)it.xbuf_elems * 16u;
    p += (size_t)it.nstages * sm.xstride;
    sm.mfe_b = (unsigned)(p - smem_raw);
    sm.mstride = (unsigned)(MV2_NDMAX * it.nrows) * (unsigned)sizeof(MfEntry);
    p += (size_t)it.nstages * sm.mstride;
    sm.kt = reinterpret_cast<double*>(p); sm.kt_b = (unsigned)(p - smem_raw); p += (size_t)it.kt_total * 8;
    sm.sp = reinterpret_cast<ProdS*>(p); sm.sp_b = (unsigned)(p - smem_raw); p += (size_t)np * sizeof(ProdS);
    sm.full = reinterpret_cast<unsigned long long*>(p); p += it.nstages * 8;
    sm.empty = reinterpret_cast<unsigned long long*>(p); p += it.nstages * 8;
    sm.setup = reinterpret_cast<unsigned long long*>(p);
    p += 16;
    sm.red_b = (unsigned)(p - smem_raw);

    const bool producer = threadIdx.x >= MV2_CONSUMERS;
    const int sl = threadIdx.x / it.nrows;               // state pair of this thread
    const int rl = threadIdx.x - sl * it.nrows;
    const bool inb = !producer && 2 * sl < it.nst;
    // which of the unit's state tiles still have work (converged states are skipped at tile granularity)
    int* tflag = reinterpret_cast<int*>(smem_raw + sm.red_b + 4 * MV2_CONSUMERS * 8);
    if (threadIdx.x < MV2_TILES_MAX) tflag[threadIdx.x] = 0;
    __syncthreads();
    bool mine = false;
    {
        const int t = threadIdx.x / it.nst;              // nst <= 32, ntiles <= 8: one thread per (tile, state)
        const int s = u.s0 + threadIdx.x;
        if (t < u.ntiles && s < nstates && (active == nullptr || active[s])) { tflag[t] = 1; mine = true; }
    }
    if (__syncthreads_or(mine) == 0) return;             // every state of the unit has converged

    if (threadIdx.x == 0) {
        for (int i = 0; i < it.nstages; ++i) {
            mbar_init(&sm.full[i], MV2_PRODUCERS);           // lane 0 of every producer warp (arrive.expect_tx) + tx bytes
            mbar_init(&sm.empty[i], MV2_CONSUMERS / 32);     // one arrival per consumer warp
        }
        mbar_init(sm.setup, 1);                              // the expect_tx arrival of the K^T / descriptor copies
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    if (producer) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(MV2_REGS_PRODUCER));
        mv2_producer(it, gdesc, cent, tab_mask, ktpool, X, ldx, nstates, u.s0, u.ntiles, active, sm);
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(MV2_REGS_CONSUMER));
#define RMB_CASE(N)                                                                                        \
    case N:                                                                                                \
        mv2_consumer<N, KC>(it, X, Y, ldx, ldy, nstates, u.s0, u.ntiles, active, sm, scale, scale_stride,  \
                            pdot, npart, u.item, sl, rl, inb);                                             \
        break;
    switch (it.nc == 1 ? 1 : (it.nc + 1) & ~1) {
        RMB_CASE(1) RMB_CASE(2) RMB_CASE(4) RMB_CASE(6) RMB_CASE(8)
        RMB_CASE(10) RMB_CASE(12)
        default: break;
    }
#undef RMB_CASE
}

}  // namespace rmb
