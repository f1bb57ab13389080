// K2 (linear rotors, dim_k = 1 everywhere): HBM-bound H.Psi for large ensembles.
//
// For dim_k = 1 the operator data (10-25 MF entries per row) outweighs the state data (16 B per row and
// state), and the J +- 2 band makes every row tile re-read its ket halo.  This kernel removes both:
//  * a persistent CTA owns a tile of T states and walks the (J,sym) blocks in order, keeping a ring of
//    >= 2W+2 ket blocks in shared memory (W = block bandwidth of the operator): every element of Psi is
//    read from HBM exactly once (TMA bulk copies issued by a producer warp, one per state and block,
//    completing on per-slot mbarriers) and every element of the product is written once;
//  * the operator is streamed the same way: after every field update k_lin_entries folds the 1 x 1 K factors
//    into the surviving MF diagonals and lays them out per bra block as [entry][row]; the producer warp
//    brings the block's entries in with ONE bulk copy, two blocks ahead of their use;
//  * lanes run over the rows of a block (entry rows and ket rows are then consecutive shared-memory words:
//    bank-conflict free); ring tiles are [state][row].
// The epilogue applies the per-state scale, stores the product with coalesced rows and produces the partial
// sums conj(y).x of the Lanczos recurrence per (state, 32-row chunk).
#pragma once
#include "rmb_matvec.cuh"

namespace rmb {

constexpr int ML_CWARPS = 16;                        // compute warps (a power of two: chunk -> warp by masking)
constexpr int ML_XPROD = 1;               // producer warps for the ket blocks (block j -> warp j % ML_XPROD)
constexpr int ML_THREADS = (ML_CWARPS + ML_XPROD + 1) * 32;   // + producer warps (TMA bulk copies: ket blocks, entries)
constexpr int ML_TS = 4;                             // states per thread for the 8-state tile (T / 2 in general)
constexpr int ML_LMAX = 40;                          // max (product, diagonal) entries per bra block
constexpr int ML_NBMAX = 4;                          // entry buffers: 2 .. 4, as many as fit next to the ring

// one (product, surviving diagonal) of a bra block: byte offset of the ket block's ring slot, diagonal offset
// (col - row) and rows of the ket block.  Record 0 of every per-block list is a header: xbyte = number of
// entries that follow.
struct __align__(16) LinEnt { unsigned xbyte; int doff; int dm2; int pad; };   // pad: ket block of the entry
constexpr int ML_FLAT = ML_LMAX + 2;                 // LinEnt records per bra block: header, entries, (block distance, dm) slot map

// static per-block data, copied to shared memory at kernel start
struct __align__(16) LinBlk {
    long long off;        // first element of the block in a (padded) state vector
    long long val_off;    // first element of the block's entries in `val`
    int dm;               // rows
    int chunk0;           // index of the block's first 32-row chunk (partial-dot slot)
    int ubase;            // (number of units of all earlier blocks) % ML_CWARPS, per state-group count G
    int L;                // filled in shared memory from the entry-list header
};

struct LinArgs {
    int nblocks;
    int W;                               // |ket block - bra block| <= W for every product
    int dms;                             // row stride of a state inside a ring slot (odd, >= max dim_m)
    int NS;                              // ring slots (>= 2W + 2)
    int NB;                              // entry buffers (2 .. ML_NBMAX)
    int ebuf_elems;                      // elements of one entry buffer (bound for the field currently applied)
    const LinBlk* blk;                   // [nblocks] static per-block table
    const LinEnt* flat;                  // [nblocks][ML_FLAT]
    const double2* val;                  // K * MF per (block, entry, row): [L_b][dm_b] per block
};

// After every field update: per bra block, the list of surviving (product, diagonal) pairs and their values
// K_p * MF_p[row] (zero where the diagonal leaves the ket block).  One CTA per block.
__global__ void __launch_bounds__(128)
k_lin_entries(int nblocks, int NS, unsigned slot_bytes, const int* __restrict__ blk_begin,
              const int* __restrict__ blk_dm, const int* __restrict__ prod_ket, const ProdD* __restrict__ prods,
              const unsigned* __restrict__ tab_mask, const MfEntry* __restrict__ cent,
              const double* __restrict__ kpool, int k_complex, const long long* __restrict__ val_off,
              LinEnt* __restrict__ flat, double2* __restrict__ val, const int* __restrict__ cshift,
              unsigned char* __restrict__ cmap) {
    __shared__ long long s_ent[ML_LMAX];      // first compacted MF entry of the (product, diagonal) pair
    __shared__ double2 s_k[ML_LMAX];          // its 1 x 1 K factor
    __shared__ LinEnt s_le[ML_LMAX];          // its descriptor
    __shared__ int s_map[ML_LMAX];            // pair -> merged entry
    __shared__ int s_L, s_U;
    const int b = blockIdx.x, lane = threadIdx.x & 31;
    const int p0 = blk_begin[b], p1 = blk_begin[b + 1], dm1 = blk_dm[b];
    LinEnt* out = flat + (size_t)b * ML_FLAT;
    if (threadIdx.x < 32) {
        int L = 0;
        for (int pb = p0; pb < p1; pb += 32) {
            const int p = pb + lane;
            int nnz = 0;
            ProdD pr;
            if (p < p1) {
                pr = prods[p];
                nnz = min(__popc(tab_mask[pr.tab]), MV2_NDMAX);
            }
            int incl = nnz;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            const int base = L + incl - nnz;
            if (p < p1 && nnz > 0) {
                double2 kv = make_double2(0.0, 0.0);
                if (k_complex) kv = reinterpret_cast<const double2*>(kpool)[pr.koff];
                else kv.x = kpool[pr.koff];
                const int ket = prod_ket[p];
                for (int q = 0; q < nnz; ++q)
                    if (base + q < ML_LMAX) {
                        const long long e0 = pr.ent_off + (long long)q * dm1;
                        LinEnt e;
                        e.xbyte = (unsigned)(ket % NS) * slot_bytes;
                        e.doff = 0;                        // from the first row whose entry lies inside the ket block
                        for (int r = 0; r < dm1; ++r) {
                            const int c = cent[e0 + r].col;
                            if (c >= 0) { e.doff = c - r; break; }
                        }
                        e.dm2 = blk_dm[ket];
                        e.pad = ket;                       // ket block (two blocks never share a ring slot within W)
                        s_le[base + q] = e;
                        s_ent[base + q] = e0;
                        s_k[base + q] = kv;
                    }
            }
            L += __shfl_sync(0xffffffffu, incl, 31);
        }
        __syncwarp();
        if (lane == 0) {
            // pairs that read the same ket block along the same diagonal (the rank-0 and rank-2 parts of a
            // polarisability, operands of a sum that couple the same blocks) are merged into one entry: their
            // values are added once per field update instead of once per state and launch
            L = min(L, ML_LMAX);
            int U = 0;
            for (int j = 0; j < L; ++j) {
                int u = 0;
                for (; u < U; ++u)
                    if (s_le[u].pad == s_le[j].pad && s_le[u].doff == s_le[j].doff) break;   // s_le[u], u < U <= j: already final
                if (u == U) {
                    const LinEnt e = s_le[j];
                    s_le[U] = e;
                    ++U;
                }
                s_map[j] = u;
            }
            LinEnt h;
            h.xbyte = (unsigned)U;
            h.doff = h.dm2 = h.pad = 0;
            out[0] = h;
            s_L = L;
            s_U = U;
            if (cshift != nullptr) {
                // (block distance, dm) -> merged entry, for the register-window kernels (k_matvec_linw below,
                // rmb_matvec_mw.cuh): 5 x 3 slots, 0xff = absent; dm = doff - (row shift between the two blocks).  Last
                // record of the block's list (travels with the bulk copy of the entries) and, if given, the separate map.
                unsigned char cmb[16];
                for (int c = 0; c < 16; ++c) cmb[c] = 0xff;
                for (int u = 0; u < U; ++u) {
                    const int ket = s_le[u].pad;
                    const int db = ket - b, dmq = s_le[u].doff - (cshift[ket] - cshift[b]);
                    if (db >= -2 && db <= 2 && dmq >= -1 && dmq <= 1) cmb[(db + 2) * 3 + (dmq + 1)] = (unsigned char)u;
                }
                unsigned char* rec = reinterpret_cast<unsigned char*>(out + (ML_FLAT - 1));
                for (int c = 0; c < 16; ++c) rec[c] = cmb[c];
                if (cmap != nullptr)
                    for (int c = 0; c < 16; ++c) cmap[(size_t)b * 16 + c] = cmb[c];
            }
        }
    }
    __syncthreads();
    const int L = s_L, U = s_U;
    for (int u = threadIdx.x; u < U; u += blockDim.x) {
        out[1 + u] = s_le[u];                  // LinEnt::pad = ket block (read by the single-launch step, rmb_fused.cuh)
    }
    double2* vout = val + val_off[b];
    for (int i = threadIdx.x; i < U * dm1; i += blockDim.x) {
        const int u = i / dm1, r = i - u * dm1;
        double2 acc = make_double2(0.0, 0.0);
        for (int j = u; j < L; ++j) {                        // members in list order (s_map[j] <= j)
            if (s_map[j] != u) continue;
            const MfEntry e = cent[s_ent[j] + r];
            if (e.col < 0) continue;
            const double2 k = s_k[j];
            acc.x += k.x * e.re - k.y * e.im;
            acc.y += k.x * e.im + k.y * e.re;
        }
        vout[i] = acc;
    }
}

// Work unit = (bra block b, 32-row chunk c, state group g): one warp, one lane per row m1, ML_TS states per
// thread; per (entry, state) the inner loop is one LDS.128 and four DFMA.  Units are dealt round-robin to the
// compute warps across blocks; warps only synchronise through mbarriers (`full`: ket block landed, `efull`:
// entries landed, `done`: bra block consumed by a warp).
// G1 = true: the single-state-group instantiation (one warp per 32-row chunk and all T states), built for the
// soak / sanitizer runs of tools/lin_soak.py only (RMB_LIN_G1=1); the product path uses two state groups.
template <int T, bool G1 = false>
__global__ void __launch_bounds__(ML_THREADS, 1)
k_matvec_lin(const LinArgs a, const double2* __restrict__ X, double2* __restrict__ Y, long long ldx,
             long long ldy, int nstates, const int* __restrict__ active, const double* __restrict__ scale,
             int scale_stride, double2* __restrict__ pdot, int npart) {
    constexpr int TS = G1 ? T : (T <= ML_TS ? T / 2 : ML_TS); // states per thread (two state groups unless G1)
    constexpr int G = T / TS;                                 // state groups
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int NS = a.NS;
    const int slot_elems = T * a.dms;                         // [state][row]
    double2* ring = reinterpret_cast<double2*>(smem_raw);                     // [NS][T][dms]
    double2* ebuf = ring + (size_t)NS * slot_elems;                           // [ML_NB][ebuf_elems]
    const int NB = a.NB;
    LinEnt* flat = reinterpret_cast<LinEnt*>(ebuf + (size_t)NB * a.ebuf_elems);      // [NB][ML_FLAT]
    unsigned long long* full = reinterpret_cast<unsigned long long*>(flat + NB * ML_FLAT);   // [NS]
    unsigned long long* done = full + NS;                                     // [NS]
    unsigned long long* efull = done + NS;                                    // [NB]
    LinBlk* blk = reinterpret_cast<LinBlk*>(efull + NB + ((2 * NS + NB) & 1));        // [nblocks], 16-byte aligned
    __shared__ long long s_sb[T];
    __shared__ double s_sc[T];
    __shared__ int s_nact;

    const int s0 = blockIdx.x * T;
    if (threadIdx.x < T) {
        const int s = s0 + threadIdx.x;
        const bool ok = s < nstates && (active == nullptr || active[s]);
        s_sb[threadIdx.x] = ok ? (long long)s : -1;
        s_sc[threadIdx.x] = (ok && scale != nullptr) ? scale[(long long)s * scale_stride] : 1.0;
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < NS; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&done[i], ML_CWARPS);
        }
        for (int i = 0; i < NB; ++i) mbar_init(&efull[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    // per-block table: no global metadata loads inside the block loops
    for (int b = threadIdx.x; b < a.nblocks; b += ML_THREADS) {
        LinBlk t = a.blk[b];
        t.L = (int)a.flat[(size_t)b * ML_FLAT].xbyte;
        blk[b] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int n = 0;
        for (int i = 0; i < T; ++i) n += s_sb[i] >= 0 ? 1 : 0;
        s_nact = n;
    }
    __syncthreads();
    const int nact = s_nact;
    if (nact == 0) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp >= ML_CWARPS) {
        // ================= producer warps =================
        // both observe `done` strictly in block order with their own counters
        int ds = 0, dph = 0, dcnt = 0;       // slot / phase of the next `done` barrier to observe, its block
        auto observe = [&](int need) {
            while (dcnt <= need) {
                mbar_wait(&done[ds], (unsigned)dph);
                ++dcnt;
                if (++ds == NS) { ds = 0; dph ^= 1; }
            }
        };
        if (warp < ML_CWARPS + ML_XPROD) {
            // ket blocks: lane t copies the block of state t; ring slot of block j is free when every warp
            // is finished with bra blocks <= j - NS + W
            const long long sb = lane < T ? s_sb[lane] : -1;
            const double2* xrow = X + (sb >= 0 ? sb : 0) * ldx;
            double2* dst = ring + (size_t)lane * a.dms;
            const int j0 = warp - ML_CWARPS;
            int xs = j0 % NS;
            for (int j = j0; j < a.nblocks; j += ML_XPROD) {
                observe(j - NS + a.W);
                const LinBlk t = blk[j];
                unsigned long long* bar = &full[xs];
                if (lane == 0) mbar_arrive_expect_tx(bar, (unsigned)nact * (unsigned)t.dm * 16u);
                __syncwarp();
                if (sb >= 0) tma_load_1d(dst + (size_t)xs * slot_elems, xrow + t.off, (unsigned)t.dm * 16u, bar);
                xs += ML_XPROD;
                if (xs >= NS) xs -= NS;
            }
        } else if (lane == 0) {
            // entries of bra block eb (descriptor list + K * MF values): buffer free when every warp is
            // finished with bra block eb - NB
            int es = 0;
            for (int eb = 0; eb < a.nblocks; ++eb) {
                observe(eb - NB);
                const LinBlk t = blk[eb];
                const unsigned vbytes = (unsigned)t.L * (unsigned)t.dm * 16u;
                unsigned long long* bar = &efull[es];
                mbar_arrive_expect_tx(bar, vbytes + (unsigned)(ML_FLAT * sizeof(LinEnt)));
                tma_load_1d(flat + es * ML_FLAT, a.flat + (size_t)eb * ML_FLAT, (unsigned)(ML_FLAT * sizeof(LinEnt)), bar);
                if (vbytes) tma_load_1d(ebuf + (size_t)es * a.ebuf_elems, a.val + t.val_off, vbytes, bar);
                if (++es == NB) es = 0;
            }
        }
        return;
    }

    // ================= compute warps =================
    // Warp w serves state group g = w % G for the whole kernel and, of every bra block, the 32-row chunks c with
    // (chunk0 + c) % WPG == w / G.  <w,v> is accumulated per thread over all its units and reduced once per state at the end
    // (one partial per state and warp of the group): the per-unit shuffle reductions were 20 % of the kernel's instructions.
    constexpr int WPG = ML_CWARPS / G;               // warps per state group (a power of two)
    static_assert(WPG * G == ML_CWARPS && (WPG & (WPG - 1)) == 0, "chunk -> warp mapping masks with WPG - 1");
    const int g = warp % G, wi = warp / G;
    const int tb = g * TS;
    const char* ring_b = reinterpret_cast<const char*>(ring);
    const unsigned st_bytes = (unsigned)a.dms * 16u;
    const unsigned tb_bytes = (unsigned)tb * st_bytes;
    double px[TS], py[TS];
#pragma unroll
    for (int t = 0; t < TS; ++t) px[t] = py[t] = 0.0;
    int ws = 0, wph = 0;                             // slot / phase of the next `full` barrier to observe
    int bs = 0;                                      // b % NS
    int es = 0, eph = 0;                             // b % NB and the phase of its `efull` barrier
    // Every warp observes every phase of every barrier in order, also for blocks in which it owns no unit: a parity wait is
    // only defined for the current or the immediately preceding phase, and the waits are what keeps a warp without work from
    // running ahead of the data.  Ket blocks 0 .. W-1 here, block b + W at the top of iteration b.
    for (int j = 0; j < a.W && j < a.nblocks; ++j) {
        mbar_wait(&full[ws], (unsigned)wph);
        if (++ws == NS) { ws = 0; wph ^= 1; }
    }
    for (int b = 0; b < a.nblocks; ++b) {
        if (b + a.W < a.nblocks) {
            mbar_wait(&full[ws], (unsigned)wph);
            if (++ws == NS) { ws = 0; wph ^= 1; }
        }
        mbar_wait(&efull[es], (unsigned)eph);
        const int4 bm = reinterpret_cast<const int4*>(blk + b)[1];          // dm, chunk0, -, L
        const int dm1 = bm.x;
        const int nch = (dm1 + 31) >> 5;
        int c = (wi - bm.y) & (WPG - 1);             // first chunk of the block owned by this warp
        if (c < nch) {
            const long long boff = blk[b].off;
            const LinEnt* fl = flat + es * ML_FLAT;
            const double2* ev = ebuf + (size_t)es * a.ebuf_elems;
            const int L = bm.w;
            const double2* xbra = ring + (size_t)bs * slot_elems + (size_t)tb * a.dms;
            for (; c < nch; c += WPG) {
                const int r = c * 32 + lane;
                const bool rv = r < dm1;
                const int rr = rv ? r : dm1 - 1;              // idle lanes repeat the last row (never stored)
                double2 acc[TS];
#pragma unroll
                for (int t = 0; t < TS; ++t) acc[t] = make_double2(0.0, 0.0);
#pragma unroll 2
                for (int j = 0; j < L; ++j) {
                    const LinEnt f = fl[1 + j];
                    const double2 e = ev[j * dm1 + rr];
                    const int col = min(max(rr + f.doff, 0), f.dm2 - 1);   // outside the ket block e is zero
                    const char* xp = ring_b + f.xbyte + tb_bytes + (unsigned)col * 16u;
                    double2 v[TS];
#pragma unroll
                    for (int t = 0; t < TS; ++t) v[t] = *reinterpret_cast<const double2*>(xp + t * st_bytes);
#pragma unroll
                    for (int t = 0; t < TS; ++t) {
                        acc[t].x = fma(e.x, v[t].x, acc[t].x);
                        acc[t].y = fma(e.x, v[t].y, acc[t].y);
                        acc[t].x = fma(-e.y, v[t].y, acc[t].x);
                        acc[t].y = fma(e.y, v[t].x, acc[t].y);
                    }
                }
                if (rv) {
#pragma unroll
                    for (int t = 0; t < TS; ++t) {
                        const long long sg = s_sb[tb + t];
                        if (sg < 0) continue;                     // warp-uniform
                        const double sc = s_sc[tb + t];
                        const double2 y = make_double2(acc[t].x * sc, acc[t].y * sc);
                        if (Y != nullptr) Y[sg * ldy + boff + r] = y;       // lanes = consecutive rows
                        if (pdot != nullptr) {
                            const double2 v = xbra[(size_t)t * a.dms + r];
                            px[t] = fma(y.x, v.x, fma(y.y, v.y, px[t]));
                            py[t] = fma(y.x, v.y, fma(-y.y, v.x, py[t]));
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&done[bs]);                // this warp is finished with bra block b
        if (++bs == NS) bs = 0;
        if (++es == NB) { es = 0; eph ^= 1; }
    }
    if (pdot != nullptr) {
        // fixed-order reduction: the lanes of the warp; the host sums the WPG partials of a state in index order
#pragma unroll
        for (int t = 0; t < TS; ++t) {
            double a0 = px[t], a1 = py[t];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a0 += __shfl_down_sync(0xffffffffu, a0, o);
                a1 += __shfl_down_sync(0xffffffffu, a1, o);
            }
            const long long sg = s_sb[tb + t];
            if (lane == 0 && sg >= 0) pdot[sg * npart + wi] = make_double2(a0, a1);
        }
    }
}

// k_matvec_linw: the same ring / producer protocol as k_matvec_lin for a tile of 4 states, but the compute warps keep
// the ket elements in a REGISTER WINDOW.  Warp g owns the 8 row positions R = 8g .. 8g+7 (counted in the last, largest
// block: blocks are ordered by J with symmetric m ranges, so the row of the same m in block b is
// r_b = R - (dm_last - dm_b) / 2) of all 4 states for the whole walk: lane = (state, row position).  At bra block b it loads
// the three elements x[b+2][m-1..m+1] that enter its 5 x 3 window from the ring (conflict-free LDS.128: 8 consecutive rows
// of a state are 128 contiguous bytes and the state stride is odd) and the surviving entry values of its row (8 distinct
// 16-byte words per load, broadcast to the 4 states): 3 + L loads of which L cost one wavefront, for 4 L DFMA -- 21
// wavefronts per 36 warp-DFMA at L = 9 against 53 in k_matvec_lin, where every complex FMA fetches its ket element from
// shared memory.  Requires |block distance| <= 2, |dm| <= 1 for every surviving diagonal (fields in a plane containing Z)
// and dm_last <= 128; the host checks (rmb.cu: lin_update_bound).  Partial sums conj(y).x per (state, row group).
constexpr int LW_CWARPS = 16;
constexpr int LW_THREADS = (LW_CWARPS + 2) * 32;
constexpr int LW_T = 4;

__global__ void __launch_bounds__(LW_THREADS, 1)
k_matvec_linw(const LinArgs a, int dm_last, int ngroups, const double2* __restrict__ X, double2* __restrict__ Y,
              long long ldx, long long ldy, int nstates, const int* __restrict__ active,
              const double* __restrict__ scale, int scale_stride, double2* __restrict__ pdot, int npart) {
    constexpr int T = LW_T;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int NS = a.NS;
    const int slot_elems = T * a.dms;                         // [state][row]
    double2* ring = reinterpret_cast<double2*>(smem_raw);                     // [NS][T][dms]
    double2* ebuf = ring + (size_t)NS * slot_elems;                           // [NB][ebuf_elems]
    const int NB = a.NB;
    LinEnt* flat = reinterpret_cast<LinEnt*>(ebuf + (size_t)NB * a.ebuf_elems);      // [NB][ML_FLAT]
    unsigned long long* full = reinterpret_cast<unsigned long long*>(flat + NB * ML_FLAT);   // [NS]
    unsigned long long* done = full + NS;                                     // [NS]
    unsigned long long* efull = done + NS;                                    // [NB]
    LinBlk* blk = reinterpret_cast<LinBlk*>(efull + NB + ((2 * NS + NB) & 1));        // [nblocks], 16-byte aligned
    __shared__ long long s_sb[T];
    __shared__ double s_sc[T];
    __shared__ int s_nact;

    const int s0 = blockIdx.x * T;
    if (threadIdx.x < T) {
        const int s = s0 + threadIdx.x;
        const bool ok = s < nstates && (active == nullptr || active[s]);
        s_sb[threadIdx.x] = ok ? (long long)s : -1;
        s_sc[threadIdx.x] = (ok && scale != nullptr) ? scale[(long long)s * scale_stride] : 1.0;
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < NS; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&done[i], LW_CWARPS);
        }
        for (int i = 0; i < NB; ++i) mbar_init(&efull[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    for (int b = threadIdx.x; b < a.nblocks; b += LW_THREADS) {
        LinBlk t = a.blk[b];
        t.L = (int)a.flat[(size_t)b * ML_FLAT].xbyte;
        blk[b] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int n = 0;
        for (int i = 0; i < T; ++i) n += s_sb[i] >= 0 ? 1 : 0;
        s_nact = n;
    }
    __syncthreads();
    const int nact = s_nact;
    if (nact == 0) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp >= LW_CWARPS) {
        // ================= producer warps (as in k_matvec_lin) =================
        int ds = 0, dph = 0, dcnt = 0;
        auto observe = [&](int need) {
            while (dcnt <= need) {
                mbar_wait(&done[ds], (unsigned)dph);
                ++dcnt;
                if (++ds == NS) { ds = 0; dph ^= 1; }
            }
        };
        if (warp == LW_CWARPS) {
            const long long sb = lane < T ? s_sb[lane] : -1;
            const double2* xrow = X + (sb >= 0 ? sb : 0) * ldx;
            double2* dst = ring + (size_t)lane * a.dms;
            int xs = 0;
            for (int j = 0; j < a.nblocks; ++j) {
                observe(j - NS + a.W);
                const LinBlk t = blk[j];
                unsigned long long* bar = &full[xs];
                if (lane == 0) mbar_arrive_expect_tx(bar, (unsigned)nact * (unsigned)t.dm * 16u);
                __syncwarp();
                if (sb >= 0) tma_load_1d(dst + (size_t)xs * slot_elems, xrow + t.off, (unsigned)t.dm * 16u, bar);
                if (++xs == NS) xs = 0;
            }
        } else if (lane == 0) {
            int es = 0;
            for (int eb = 0; eb < a.nblocks; ++eb) {
                observe(eb - NB);
                const LinBlk t = blk[eb];
                const unsigned vbytes = (unsigned)t.L * (unsigned)t.dm * 16u;
                unsigned long long* bar = &efull[es];
                mbar_arrive_expect_tx(bar, vbytes + (unsigned)(ML_FLAT * sizeof(LinEnt)));
                tma_load_1d(flat + es * ML_FLAT, a.flat + (size_t)eb * ML_FLAT, (unsigned)(ML_FLAT * sizeof(LinEnt)), bar);
                if (vbytes) tma_load_1d(ebuf + (size_t)es * a.ebuf_elems, a.val + t.val_off, vbytes, bar);
                if (++es == NB) es = 0;
            }
        }
        return;
    }

    // ================= compute warps =================
    const int ls = lane >> 3, q = lane & 7;                 // state of the tile, row position inside the group
    const int R = warp * 8 + q;
    const long long sg = s_sb[ls];
    const bool lv = sg >= 0 && R < dm_last;
    const double sc = s_sc[ls];
    const int nb = a.nblocks;
    // x[bp][r_bp + d] from the ring (0 outside the block)
    auto ring_x = [&](int bp, int slot, int d) -> double2 {
        double2 v = make_double2(0.0, 0.0);
        if (lv && bp < nb) {
            const int dmp = blk[bp].dm;
            const int r = R - ((dm_last - dmp) >> 1) + d;
            if (r >= 0 && r < dmp) v = ring[(size_t)slot * slot_elems + (size_t)ls * a.dms + r];
        }
        return v;
    };
    double2 w[5][3];
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int d = 0; d < 3; ++d) w[i][d] = make_double2(0.0, 0.0);
    int waited = -1, ws = 0, wph = 0;
    int bs = 0, es = 0, eph = 0;
    double pre = 0.0, pim = 0.0;
    for (int b = 0; b < nb; ++b) {
        // (the window reads block b+2 whatever the band width W <= 2 of the operator: the producer can always deliver it
        //  without waiting for this block, NS >= W + 3)
        const int newest = min(b + 2, nb - 1);
        while (waited < newest) {
            ++waited;
            mbar_wait(&full[ws], (unsigned)wph);
            if (++ws == NS) { ws = 0; wph ^= 1; }
        }
        mbar_wait(&efull[es], (unsigned)eph);
        // ---- window: blocks b-2 .. b+2; block b+2 enters now (blocks 0 .. 2 at the first step)
        if (b == 0) {
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int d = 0; d < 3; ++d) w[2 + i][d] = ring_x(i, i % NS, d - 1);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int d = 0; d < 3; ++d) w[i][d] = w[i + 1][d];
            const int slot2 = (b + 2) % NS;
#pragma unroll
            for (int d = 0; d < 3; ++d) w[4][d] = ring_x(b + 2, slot2, d - 1);
        }
        const LinBlk bt = blk[b];
        const int r = R - ((dm_last - bt.dm) >> 1);
        const bool rowv = lv && r >= 0 && r < bt.dm;
        const uint4 cm = *reinterpret_cast<const uint4*>(flat + es * ML_FLAT + (ML_FLAT - 1));
        const unsigned cmw[4] = {cm.x, cm.y, cm.z, cm.w};
        const double2* ev = ebuf + (size_t)es * a.ebuf_elems + (rowv ? r : 0);
        double2 acc = make_double2(0.0, 0.0), acc2 = make_double2(0.0, 0.0);
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            double2 e[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const int c = i * 3 + d;
                const unsigned u = (cmw[c >> 2] >> (8 * (c & 3))) & 0xffu;   // warp-uniform
                e[d] = make_double2(0.0, 0.0);
                if (u != 0xffu && rowv) e[d] = ev[u * bt.dm];
            }
#pragma unroll
            for (int d = 0; d < 3; ++d) {
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
        if (rowv) {
            const double2 y = make_double2((acc.x + acc2.x) * sc, (acc.y + acc2.y) * sc);
            if (Y != nullptr) Y[sg * ldy + bt.off + r] = y;
            const double2 xb = w[2][1];
            pre += y.x * xb.x + y.y * xb.y;
            pim += y.x * xb.y - y.y * xb.x;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&done[bs]);                // this warp is finished with bra block b
        if (++bs == NS) bs = 0;
        if (++es == NB) { es = 0; eph ^= 1; }
    }
    if (pdot != nullptr && warp < ngroups) {
        // the 8 row positions of a state are adjacent lanes: fixed-order reduction
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            pre += __shfl_xor_sync(0xffffffffu, pre, o);
            pim += __shfl_xor_sync(0xffffffffu, pim, o);
        }
        if (q == 0 && sg >= 0) pdot[sg * npart + warp] = make_double2(pre, pim);
    }
}

}  // namespace rmb
