// K2l, second generation ("m-walk"): H.Psi of a linear rotor without shared memory, every ket element read once into a
// REGISTER window.
//
// For dim_k = 1 the operator couples (J, m) to (J + dJ, m + dm) with |dJ| <= 2 (block distance) and, for fields in a
// plane containing Z, |dm| <= 1.  k_matvec_lin keeps a ring of ket blocks in shared memory with lanes over the rows of a
// block: every complex FMA then needs its ket element from shared memory, and the shared-memory pipe (1.46 wavefronts per
// warp-DFMA) caps the kernel at roughly half of the HBM roofline.  Here a thread owns ONE m (one row position) of ONE
// state and walks the blocks in J order with the 5 x 3 ket elements x[b-2..b+2][m-1..m+1] in registers: per block step
// it loads the three elements of block b+3 that enter the window (one step ahead of their first use) and the surviving
// entry values of its row, and issues up to 15 complex FMAs.  A warp = 8 states x 4 consecutive m (64 contiguous bytes
// per state and load; entry values are read by address, broadcast to the 8 states); the 8 warps of a CTA take the SAME
// m-group for 8 different state tiles, so entry values hit L1 after the first warp and the warps of a CTA do equal work.
// CTAs are persistent and draw (m-group, 64-state super tile) items, longest first, from an atomic counter.
//
// Row bookkeeping uses block metadata only (the C ABI carries no m quanta): blocks are ordered by J with symmetric,
// contiguous m ranges, so a thread's row in block b is  r_b = R - (cshift[last] - cshift[b]),  cshift[b+1] - cshift[b] =
// (dm[b+1] - dm[b]) / 2, and an entry of diagonal `doff` (col - row) between bra b and ket b' has
// dm = doff - (cshift[b'] - cshift[b]).  The host enables the kernel only when this holds and every surviving diagonal
// has |dm| <= 1 (rmb.cu: lin_update_bound); k_lin_entries then fills, per bra block, a 15-slot map (dJ, dm) -> entry.
// Epilogue as in the other matvec kernels: per-state scale, store, partial sums conj(y).x per (state, m-group).
#pragma once
#include "rmb_matvec_lin.cuh"

namespace rmb {

constexpr int MW_WARPS = 8;                    // warps per CTA = state tiles of 8 states per item
constexpr int MW_THREADS = MW_WARPS * 32;
constexpr int MW_DB = 2;                       // |block distance| <= MW_DB
constexpr int MW_DM = 1;                       // |dm| <= MW_DM
constexpr int MW_NB = 2 * MW_DB + 1, MW_NM = 2 * MW_DM + 1;
constexpr int MW_NC = MW_NB * MW_NM;           // 15 (dJ, dm) combinations

struct MwItem { int group; int tile0; int b_first; int pad; };   // m-group, first state of the 64-state super tile, first active block

struct MwArgs {
    int nblocks;
    int nitems;
    int dm_last;                         // rows of the last (largest) block
    const long long* blk_off;            // [nblocks] first element of block b in a state vector
    const int* blk_dm;                   // [nblocks]
    const int* cshift;                   // [nblocks] prefix of the row shifts between consecutive blocks
    const long long* val_off;            // [nblocks + 1] first entry value of block b (k_lin_entries)
    const unsigned char* cmap;           // [nblocks][16] (dJ, dm) -> merged entry index, 0xff = absent
    const double2* val;                  // entry values [U_b][dm_b] per block
    const MwItem* items;                 // sorted by length, longest first
    int* counter;                        // work queue head (zeroed before the launch)
};

// shared-memory layout: offsets, value offsets (8 B each), rows, shifts (4 B each), then the 16-byte slot maps
__host__ __device__ inline size_t mw_meta_bytes(int nblocks) { return ((size_t)24 * nblocks + 15) & ~(size_t)15; }
__host__ __device__ inline size_t mw_smem_bytes(int nblocks) { return mw_meta_bytes(nblocks) + (size_t)16 * nblocks; }

__global__ void __launch_bounds__(MW_THREADS, 2)
k_matvec_mw(const MwArgs a, const double2* __restrict__ X, double2* __restrict__ Y, long long ldx, long long ldy,
            int nstates, const int* __restrict__ active, const double* __restrict__ scale, int scale_stride,
            double2* __restrict__ pdot, int npart) {
    extern __shared__ __align__(16) unsigned char mw_smem[];
    // per-block metadata in shared memory: uniform, conflict-free reads inside the walk
    long long* s_off = reinterpret_cast<long long*>(mw_smem);                    // [nblocks]
    long long* s_voff = s_off + a.nblocks;                                       // [nblocks]
    int* s_dm = reinterpret_cast<int*>(s_voff + a.nblocks);                      // [nblocks]
    int* s_cs = s_dm + a.nblocks;                                                // [nblocks]
    uint4* s_cmap = reinterpret_cast<uint4*>(mw_smem + mw_meta_bytes(a.nblocks));                    // [nblocks]
    __shared__ int s_item;
    for (int b = threadIdx.x; b < a.nblocks; b += MW_THREADS) {
        s_off[b] = a.blk_off[b];
        s_voff[b] = a.val_off[b];
        s_dm[b] = a.blk_dm[b];
        s_cs[b] = a.cshift[b];
        s_cmap[b] = reinterpret_cast<const uint4*>(a.cmap)[b];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ls = lane >> 2, q = lane & 3;             // state of the tile, m slot of the group
    const int nb = a.nblocks;
    const int cs_last = s_cs[nb - 1];

    for (;;) {
        if (threadIdx.x == 0) s_item = atomicAdd(a.counter, 1);
        __syncthreads();
        const int item = s_item;
        __syncthreads();
        if (item >= a.nitems) break;
        const MwItem it = a.items[item];
        const int sg = it.tile0 + warp * 8 + ls;                                 // this lane's state
        const bool sv = sg < nstates && (active == nullptr || active[sg]);
        if (!__any_sync(0xffffffffu, sv)) continue;                              // (warp-uniform; the CTA barriers above are outside)
        const int R = it.group * 4 + q;                                          // row position in the last block
        const bool rv = R < a.dm_last;
        const double2* xs = X + (long long)(sv ? sg : 0) * ldx;
        double2* ys = (Y != nullptr) ? Y + (long long)(sv ? sg : 0) * ldy : nullptr;
        const double sc = (sv && scale != nullptr) ? scale[(long long)sg * scale_stride] : 1.0;
        const bool lv = sv && rv;                                                // lane does real work

        // x[b'][r_b' + d] or 0 outside the block / state tile
        auto ldx_ = [&](int bp, int d) -> double2 {
            double2 v = make_double2(0.0, 0.0);
            if (lv && bp >= 0 && bp < nb) {
                const int r = R - (cs_last - s_cs[bp]) + d;
                if (r >= 0 && r < s_dm[bp]) v = xs[s_off[bp] + r];
            }
            return v;
        };
        // register window w[db + 2][d + 1] = x[b + db][m + d]; `inc` = the row of block b + 3 (enters after this step)
        double2 w[MW_NB][MW_NM], inc[MW_NM];
        const int b0 = it.b_first;
#pragma unroll
        for (int i = 0; i < MW_NB; ++i)
#pragma unroll
            for (int d = 0; d < MW_NM; ++d) w[i][d] = ldx_(b0 + i - MW_DB, d - MW_DM);
        double pre = 0.0, pim = 0.0;
        for (int b = b0; b < nb; ++b) {
#pragma unroll
            for (int d = 0; d < MW_NM; ++d) inc[d] = ldx_(b + MW_DB + 1, d - MW_DM);     // one step ahead of its first use
            const int r = R - (cs_last - s_cs[b]);
            const int dmb = s_dm[b];
            const bool rowv = lv && r >= 0 && r < dmb;
            const uint4 cm = s_cmap[b];
            const unsigned cmw[4] = {cm.x, cm.y, cm.z, cm.w};
            const double2* ev = a.val + s_voff[b] + (rowv ? r : 0);
            // entry values of this row: all loads of a window row (3 m offsets) are issued before the first FMA, absent
            // slots predicated off (a branch per slot would serialise one L2 latency per entry); two accumulators halve
            // the dependent DFMA chain
            double2 acc = make_double2(0.0, 0.0), acc2 = make_double2(0.0, 0.0);
#pragma unroll
            for (int i = 0; i < MW_NB; ++i) {
                double2 e[MW_NM];
#pragma unroll
                for (int d = 0; d < MW_NM; ++d) {
                    const int c = i * MW_NM + d;
                    const unsigned u = (cmw[c >> 2] >> (8 * (c & 3))) & 0xffu;   // warp-uniform
                    e[d] = make_double2(0.0, 0.0);
                    if (u != 0xffu && rowv) e[d] = ev[(long long)u * dmb];
                }
#pragma unroll
                for (int d = 0; d < MW_NM; ++d) {
                    const double2 v = w[i][d];
                    if (d & 1) {
                        acc2.x = fma(e[d].x, v.x, acc2.x);
                        acc2.y = fma(e[d].x, v.y, acc2.y);
                        acc2.x = fma(-e[d].y, v.y, acc2.x);
                        acc2.y = fma(e[d].y, v.x, acc2.y);
                    } else {
                        acc.x = fma(e[d].x, v.x, acc.x);
                        acc.y = fma(e[d].x, v.y, acc.y);
                        acc.x = fma(-e[d].y, v.y, acc.x);
                        acc.y = fma(e[d].y, v.x, acc.y);
                    }
                }
            }
            acc.x += acc2.x;
            acc.y += acc2.y;
            if (rowv) {
                const double2 y = make_double2(acc.x * sc, acc.y * sc);
                if (ys != nullptr) ys[s_off[b] + r] = y;
                const double2 xb = w[MW_DB][MW_DM];                              // x[b][m]
                pre += y.x * xb.x + y.y * xb.y;
                pim += y.x * xb.y - y.y * xb.x;
            }
            // slide the window
#pragma unroll
            for (int i = 0; i + 1 < MW_NB; ++i)
#pragma unroll
                for (int d = 0; d < MW_NM; ++d) w[i][d] = w[i + 1][d];
#pragma unroll
            for (int d = 0; d < MW_NM; ++d) w[MW_NB - 1][d] = inc[d];
        }
        if (pdot != nullptr) {
            // the 4 m slots of a state are adjacent lanes: fixed-order reduction
            pre += __shfl_xor_sync(0xffffffffu, pre, 1);
            pim += __shfl_xor_sync(0xffffffffu, pim, 1);
            pre += __shfl_xor_sync(0xffffffffu, pre, 2);
            pim += __shfl_xor_sync(0xffffffffu, pim, 2);
            if (q == 0 && sv) pdot[(long long)sg * npart + it.group] = make_double2(pre, pim);
        }
    }
}

}  // namespace rmb
