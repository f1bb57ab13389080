// Host side of librichmol_b200.so: the C ABI declared in include/richmol_b200.h.
// Operator upload + work decomposition, kernel launches, the Lanczos driver loop.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <map>
#include <numeric>

#include "rmb_kernels.cuh"
#include "rmb_lanczos.cuh"
#include "rmb_matvec.cuh"
#include "rmb_fused.cuh"
#include "rmb_matvec_dmma.cuh"
#include "rmb_matvec_lin.cuh"
#include "rmb_matvec_mw.cuh"

namespace rmb {

static thread_local std::string g_error;

void set_error(const std::string& msg) { g_error = msg; }

int cuda_fail(cudaError_t e, const char* what) {
    g_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
    return RMB_ERR_CUDA;
}

// dynamic shared memory k_lanczos_fused has been allowed so far (13 KB of static shared memory come on top)
static size_t g_fused_smem = 16 * 1024;

template <typename T>
static int upload(T** dptr, const T* src, size_t count) {
    *dptr = nullptr;
    if (count == 0) count = 1;   // keep pointers valid
    RMB_CUDA(cudaMalloc((void**)dptr, count * sizeof(T)));
    if (src) RMB_CUDA(cudaMemcpy(*dptr, src, count * sizeof(T), cudaMemcpyHostToDevice));
    return RMB_OK;
}

static inline int nchunks(long long n) { return (int)((n + VEC_CHUNK - 1) / VEC_CHUNK); }

}  // namespace rmb

using namespace rmb;

extern "C" {

int32_t rmb_abi_version(void) { return RMB_ABI_VERSION; }

const char* rmb_last_error(void) { return g_error.c_str(); }

int32_t rmb_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cuda_fail(e, "cudaGetDeviceCount");
        return RMB_ERR_CUDA;
    }
    return n;
}

void rmb_operator_destroy(rmb_operator* op) {
    if (!op) return;
    cudaFree(op->d_prods);
    cudaFree(op->d_items);
    cudaFree(op->d_pmap);
    cudaFree(op->d_row_blk);
    cudaFree(op->d_blk_begin);
    cudaFree(op->d_blk_off);
    cudaFree(op->d_blk_dm);
    cudaFree(op->d_prod_ket);
    cudaFree(op->d_lin_blk);
    cudaFree(op->d_lin_flat);
    cudaFree(op->d_lin_val);
    cudaFree(op->d_lin_val_off);
    cudaFree(op->d_cshift);
    cudaFree(op->d_cmap);
    cudaFree(op->d_mw_counter);
    for (auto& kv : op->mw_items_cache) cudaFree(kv.second.first);
    cudaFree(op->d_items2);
    cudaFree(op->d_itemsG);
    for (auto& kv : op->unitsG_cache) cudaFree(kv.second.first);
    cudaFree(op->d_gdesc);
    cudaFree(op->d_ktpool);
    for (auto& kv : op->units_cache) cudaFree(kv.second.first);
    cudaFree(op->d_ent_col);
    cudaFree(op->d_ent_val);
    cudaFree(op->d_ent_cent);
    cudaFree(op->d_ent_tab);
    cudaFree(op->d_tab_off);
    cudaFree(op->d_tab_nd);
    cudaFree(op->d_tab_mask);
    cudaFree(op->d_kpool);
    cudaFree(op->d_flags);
    for (auto& p : op->parts) {
        cudaFree(p.d_coef);
        cudaFree(p.d_fprod);
    }
    for (auto& w : op->wsp) {
        for (auto* s : w.slabs) cudaFree(s);
        cudaFree(w.d_w);
        cudaFree(w.d_slab_ptrs);
        cudaFree(w.d_alpha);
        cudaFree(w.d_beta);
        cudaFree(w.d_ccur);
        cudaFree(w.d_dc);
        cudaFree(w.d_rinv);
        cudaFree(w.d_ceff);
        for (auto e : w.it_events) cudaEventDestroy(e);
        cudaFree(w.d_active);
        cudaFree(w.d_order);
        cudaFree(w.d_pdot);
        cudaFree(w.d_pnrm);
        cudaFree(w.d_pconv);
        cudaFree(w.d_pg0);
        cudaFree(w.d_gdiag);
        cudaFree(w.d_ticket);
        cudaFree(w.d_ctrl);
        if (w.h_ctrl) cudaFreeHost(w.h_ctrl);
    }
    cudaFree(op->d_stage);
    cudaFree(op->d_pipe_orders);
    cudaFree(op->d_expv);
    if (op->s_in) cudaStreamDestroy(op->s_in);
    if (op->s_out) cudaStreamDestroy(op->s_out);
    for (auto e : op->pipe_events) cudaEventDestroy(e);
    cudaFree(op->d_phase);
    for (auto& e : op->mv_events) {
        cudaEventDestroy(e.first);
        cudaEventDestroy(e.second);
    }
    for (auto& e : op->mv_event_pool) {
        cudaEventDestroy(e.first);
        cudaEventDestroy(e.second);
    }
    delete op;
}

// Linear-rotor kernel: bound of the entries per bra block for the fields currently applied (a diagonal can
// only survive the contraction if a Cartesian component with a non-zero field product has a coefficient on
// it), and the number of entry buffers that fit next to the ring.
static constexpr size_t LIN_SMEM_MAX = 226 * 1024;
static void lin_update_bound(rmb_operator* op) {
    if (!op->lin_ok) return;
    const int ntab = (int)op->h_tab_part.size();
    std::vector<int> alive(ntab, 0);
    for (int t = 0; t < ntab; ++t) {
        const PartH& ph = op->parts[op->h_tab_part[t]];
        const unsigned nz = ph.has_field ? (ph.all_dropped ? 0u : ph.nzmask) : 0xffffffffu;
        int n = 0;
        for (int j = op->h_diag_off[t]; j < op->h_diag_off[t + 1]; ++j) n += (op->h_diag_cart[j] & nz) ? 1 : 0;
        alive[t] = std::min(n, MV2_NDMAX);
    }
    long long ebuf = 1;
    for (int b = 0; b < op->nblocks; ++b) {
        long long l = 0;
        for (int p = op->h_bra_begin[b]; p < op->h_bra_begin[b + 1]; ++p) l += alive[op->h_prods[p].tab];
        ebuf = std::max(ebuf, std::min<long long>(l, ML_LMAX) * op->h_blk_dm[b]);
    }
    // register-window kernels (k_matvec_linw, rmb_matvec_mw.cuh): every diagonal that can survive the fields has |dm| <= 1
    op->mw_cur = op->sym_ok;
    for (size_t p = 0; p < op->h_prods.size() && op->mw_cur; ++p) {
        const int t = op->h_prods[p].tab;
        const PartH& ph = op->parts[op->h_tab_part[t]];
        const unsigned nz = ph.has_field ? (ph.all_dropped ? 0u : ph.nzmask) : 0xffffffffu;
        for (int j = 0; j < op->h_prods[p].nd; ++j)
            if ((op->h_diag_cart[op->h_diag_off[t] + j] & nz) && std::abs((int)op->h_prod_dm[op->h_prod_dm_off[p] + j]) > MW_DM)
                op->mw_cur = false;
    }
    op->lw_cur = op->mw_cur && op->lw_static;
    op->mw_cur = op->mw_cur && op->mw_static;
    op->lin_ebuf_cur = (int)std::min<long long>(ebuf, op->lin_ebuf);
    const size_t per = (size_t)op->lin_ebuf_cur * 16 + ML_FLAT * sizeof(LinEnt);
    {
        // shared memory of k_matvec_linw: the same layout with a ring of 4-state tiles
        // (deep rings: a block step of a 4-state tile is a few hundred cycles of work, a bulk copy takes a couple of
        //  thousand to land, so the producers must run many blocks ahead)
        const size_t slotb = (size_t)LW_T * op->lin_dm_max * 16;
        const size_t misc = (size_t)(2 * 16 + 8 + 1) * 8 + (size_t)op->nblocks * sizeof(LinBlk) + 128;
        int nb4 = (int)std::min<size_t>(8, (size_t)(0.55 * LIN_SMEM_MAX) / per);
        nb4 = std::max(2, nb4);
        size_t left = LIN_SMEM_MAX > misc + (size_t)nb4 * per ? LIN_SMEM_MAX - misc - (size_t)nb4 * per : 0;
        int ns4 = (int)std::min<size_t>(16, left / slotb);
        op->lw_NB = nb4;
        op->lw_NS = ns4;
        op->lw_smem = misc + (size_t)ns4 * slotb + (size_t)nb4 * per;
        if (ns4 < op->lin_W + 3 || ns4 < 5 || op->lw_smem > LIN_SMEM_MAX) op->lw_cur = false;
    }
    int nb = (int)((LIN_SMEM_MAX - op->lin_smem_fixed) / per);
    op->lin_NB = std::max(2, std::min(nb, ML_NBMAX));
    op->lin_smem = op->lin_smem_fixed + (size_t)op->lin_NB * per;
}

int32_t rmb_operator_create(const rmb_operator_desc* d, rmb_operator** out) {
    if (!d || !out) {
        set_error("null descriptor");
        return RMB_ERR_INVALID;
    }
    *out = nullptr;
    if (d->nblocks <= 0 || d->nparts < 0) {
        set_error("operator needs at least one (J,sym) block");
        return RMB_ERR_INVALID;
    }
    int ndev = 0;
    RMB_CUDA(cudaGetDeviceCount(&ndev));
    rmb_operator* op = new rmb_operator();
    struct Guard {
        rmb_operator* p;
        ~Guard() { if (p) rmb_operator_destroy(p); }
    } guard{op};
    RMB_CUDA(cudaGetDevice(&op->device));
    RMB_CUDA(cudaDeviceGetAttribute(&op->num_sms, cudaDevAttrMultiProcessorCount, op->device));
    op->nblocks = d->nblocks;
    op->n = d->blk_off[d->nblocks];
    // internal (padded) layout: every (J,sym) block stores rows of (dim_k | 1) elements, so that rows
    // staged into shared memory are contiguous in HBM and row-strided reads are bank-conflict free
    std::vector<long long> poff(d->nblocks + 1, 0);
    for (int b = 0; b < d->nblocks; ++b)
        poff[b + 1] = poff[b] + (long long)d->blk_dm[b] * (d->blk_dk[b] | 1);
    op->np = poff[d->nblocks];
    if (op->np >= (1LL << 31)) {
        set_error("Hilbert space too large for 32-bit internal indices");
        return RMB_ERR_INVALID;
    }
    {
        std::vector<int> pmap((size_t)op->n);
        for (int b = 0; b < d->nblocks; ++b) {
            const int dk = d->blk_dk[b], dkp = dk | 1;
            for (int r = 0; r < d->blk_dm[b]; ++r)
                for (int c = 0; c < dk; ++c)
                    pmap[(size_t)(d->blk_off[b] + (long long)r * dk + c)] = (int)(poff[b] + (long long)r * dkp + c);
        }
        int rc0 = upload(&op->d_pmap, pmap.data(), pmap.size());
        if (rc0) return rc0;
    }
    for (int b = 0; b < d->nblocks; ++b) {
        if (d->blk_off[b + 1] - d->blk_off[b] != (long long)d->blk_dm[b] * d->blk_dk[b]) {
            set_error("block offsets inconsistent with dim_m*dim_k");
            return RMB_ERR_INVALID;
        }
        op->dk_max = std::max(op->dk_max, d->blk_dk[b]);
    }
    // ---- merge parts: global entry pool, global K pool (complex if any part is complex)
    bool kc = false;
    for (int q = 0; q < d->nparts; ++q) kc = kc || d->parts[q].k_is_complex;
    op->k_complex = kc;
    std::vector<int> ent_col;
    std::vector<double> kpool;
    struct HProd { int bra, ket; long long koff, ent_off; int nd, tab; };
    std::vector<int> ent_tab, tab_off, tab_nd;
    std::vector<HProd> hp;
    op->parts.resize(d->nparts);
    for (int q = 0; q < d->nparts; ++q) {
        const rmb_part_desc& pd = d->parts[q];
        PartH& ph = op->parts[q];
        ph.ncart = pd.ncart;
        const long long nent = pd.tb_off ? pd.tb_off[pd.ntables] : 0;
        ph.ent_begin = (long long)ent_col.size();
        ph.ent_end = ph.ent_begin + nent;
        ph.tab_begin = (int)tab_off.size();
        ph.tab_end = ph.tab_begin + pd.ntables;
        for (int t = 0; t < pd.ntables; ++t) {
            tab_off.push_back((int)(ph.ent_begin + pd.tb_off[t]));
            tab_nd.push_back(pd.tb_nd[t]);
            for (long long e = pd.tb_off[t]; e < pd.tb_off[t + 1]; ++e) ent_tab.push_back(ph.tab_begin + t);
        }
        ent_col.insert(ent_col.end(), pd.ent_col, pd.ent_col + nent);
        const long long kbase = (long long)(kc ? kpool.size() / 2 : kpool.size());
        if (kc && !pd.k_is_complex) {
            for (long long i = 0; i < pd.kpool_len; ++i) {
                kpool.push_back(pd.kpool[i]);
                kpool.push_back(0.0);
            }
        } else {
            kpool.insert(kpool.end(), pd.kpool, pd.kpool + pd.kpool_len * (pd.k_is_complex ? 2 : 1));
        }
        for (int t = 0; t < pd.ntables; ++t) op->nd_max = std::max(op->nd_max, pd.tb_nd[t]);
        // per (table, diagonal): which Cartesian components can make it non-zero (host-side upper bound of the
        // device's diagonal masks, used to size the entry buffers of the linear-rotor kernel)
        for (int t = 0; t < pd.ntables; ++t) {
            const int nd = pd.tb_nd[t];
            if (op->h_diag_off.empty()) op->h_diag_off.push_back(0);
            const size_t base = op->h_diag_cart.size();
            op->h_diag_cart.resize(base + nd, 0u);
            for (long long e = pd.tb_off[t]; e < pd.tb_off[t + 1]; ++e) {
                if (pd.ent_col[e] < 0) continue;
                const int j = (int)((e - pd.tb_off[t]) % nd);
                for (int c = 0; c < pd.ncart && c < 32; ++c) {
                    const double* v = pd.ent_coef + 2 * ((size_t)c * nent + e);
                    if (v[0] != 0.0 || v[1] != 0.0) op->h_diag_cart[base + j] |= 1u << c;
                }
            }
            op->h_diag_off.push_back((int)op->h_diag_cart.size());
            op->h_tab_part.push_back(q);
        }
        for (int p = 0; p < pd.nprod; ++p) {
            const int b1 = pd.pr_bra[p], b2 = pd.pr_ket[p], t = pd.pr_table[p];
            if (b1 < 0 || b1 >= d->nblocks || b2 < 0 || b2 >= d->nblocks || t < 0 || t >= pd.ntables) {
                set_error("product references a block or table out of range");
                return RMB_ERR_INVALID;
            }
            if (pd.tb_dm1[t] != d->blk_dm[b1] || pd.tb_dm2[t] != d->blk_dm[b2]) {
                set_error("M table shape does not match the (J,sym) block dim_m");
                return RMB_ERR_INVALID;
            }
            if (pd.pr_koff[p] < 0 ||
                pd.pr_koff[p] + (long long)d->blk_dk[b1] * d->blk_dk[b2] > pd.kpool_len) {
                set_error("K block out of range of kpool");
                return RMB_ERR_INVALID;
            }
            hp.push_back({b1, b2, kbase + pd.pr_koff[p], ph.ent_begin + pd.tb_off[t], pd.tb_nd[t], ph.tab_begin + t});
        }
        for (long long e = 0; e < nent; ++e) {
            // column bounds are checked per table below
            (void)e;
        }
        for (int t = 0; t < pd.ntables; ++t)
            for (long long e = pd.tb_off[t]; e < pd.tb_off[t + 1]; ++e)
                if (pd.ent_col[e] >= pd.tb_dm2[t]) {
                    set_error("M table column index out of range");
                    return RMB_ERR_INVALID;
                }
        int rc = upload(&ph.d_coef, reinterpret_cast<const cplx*>(pd.ent_coef), (size_t)pd.ncart * nent);
        if (rc) return rc;
        rc = upload<double>(&ph.d_fprod, nullptr, (size_t)std::max(pd.ncart, 1));
        if (rc) return rc;
    }
    op->nent = (long long)ent_col.size();
    tab_off.push_back((int)ent_col.size());   // extent of the last table (tab_off has ntab + 1 entries)
    op->nprod = (int)hp.size();
    // ---- sort products by bra block (stable: keeps part / input order inside a block)
    std::vector<int> perm(hp.size());
    std::iota(perm.begin(), perm.end(), 0);
    std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return hp[a].bra < hp[b].bra; });
    op->h_prods.resize(hp.size());
    std::vector<int> h_prod_ket, h_prod_bra;
    std::vector<int> bra_begin(d->nblocks + 1, 0);
    for (size_t i = 0; i < perm.size(); ++i) {
        const HProd& h = hp[perm[i]];
        ProdD& pr = op->h_prods[i];
        pr.ket_off = poff[h.ket];
        pr.koff = h.koff;
        pr.ent_off = h.ent_off;
        pr.dk2 = d->blk_dk[h.ket];
        pr.dm2 = d->blk_dm[h.ket];
        pr.nd = h.nd;
        pr.tab = h.tab;
        h_prod_ket.push_back(h.ket);
        h_prod_bra.push_back(h.bra);
        op->h_prod_dm1.push_back(d->blk_dm[h.bra]);
        op->h_prod_dk1.push_back(d->blk_dk[h.bra]);
        bra_begin[h.bra + 1]++;
    }
    for (int b = 0; b < d->nblocks; ++b) bra_begin[b + 1] += bra_begin[b];
    // ---- work items: row/column tiles of every bra block (also blocks without products: zeros)
    const int S = 4;
    op->matvec_S = S;
    const int max_out = MV_THREADS * MV_ACC / S;   // outputs per state and CTA
    int zstride = 1;
    double flops = 0, opbytes = 0;
    // diagonal span (max - min of col - row) of every product's MF table: rows staged = tile rows + span
    std::vector<int> h_span(op->h_prods.size(), 0);
    for (size_t p = 0; p < op->h_prods.size(); ++p) {
        const ProdD& q = op->h_prods[p];
        const long long nrow = (long long)(tab_off[q.tab + 1] - tab_off[q.tab]) / q.nd;
        int lo = 1 << 30, hi = -(1 << 30);
        for (long long r = 0; r < nrow; ++r)
            for (int j = 0; j < q.nd; ++j) {
                const int col = ent_col[q.ent_off + r * q.nd + j];
                if (col >= 0) { lo = std::min(lo, col - (int)r); hi = std::max(hi, col - (int)r); }
            }
        h_span[p] = hi >= lo ? hi - lo : 0;
    }
    auto item_smem_st = [](size_t xbuf_elems, int nrows, size_t kt_doubles, int nprod, int stages) {
        return (size_t)stages * xbuf_elems * 16 + (size_t)stages * MV2_NDMAX * nrows * sizeof(MfEntry) +
               kt_doubles * 8 + (size_t)nprod * sizeof(ProdS) + (size_t)(2 * stages) * 8 + 16 + MV2_RED_BYTES + 128;
    };
    auto item_smem = [&](size_t xbuf_elems, int nrows, size_t kt_doubles, int nprod) {
        return item_smem_st(xbuf_elems, nrows, kt_doubles, nprod, MV2_STAGES);
    };
    const char* force = getenv("RMB_MATVEC");
    const bool force_scalar = force && strcmp(force, "scalar") == 0;
    std::vector<Item2D> items2;
    std::vector<ItemD2> itemsG;
    const bool force_nogemm = force && strcmp(force, "nogemm") == 0;
    int gemm_min_dk = MV2_NCMAX;              // bra blocks with more columns go to the DMMA kernel
    if (const char* e = getenv("RMB_GEMM_MIN_DK")) gemm_min_dk = atoi(e);
    std::vector<XRange> xranges;
    std::vector<ProdS> gdesc;                 // static per-(item, product) descriptors, shared-memory layout
    std::vector<double> ktpool;               // K^T images per (bra block, column chunk), shared-memory layout
    std::map<std::pair<int, int>, long long> kt_index;
    const size_t smem_budget = 112 * 1024;   // two CTAs per SM (228 KB per SM, 1 KB reserved per CTA)
    for (int b = 0; b < d->nblocks; ++b) {
        const int dk1 = d->blk_dk[b], dm1 = d->blk_dm[b];
        if (dk1 == 0 || dm1 == 0) continue;
        int dk2max = 1, ndmax = 1;
        for (int p = bra_begin[b]; p < bra_begin[b + 1]; ++p) {
            const ProdD& pr = op->h_prods[p];
            dk2max = std::max(dk2max, pr.dk2);
            ndmax = std::max(ndmax, pr.nd);
            // algorithmic work: K contraction + banded M contraction
            flops += (kc ? 8.0 : 4.0) * dm1 * (double)dk1 * pr.dk2 + 8.0 * (double)pr.nd * dm1 * pr.dk2;
            opbytes += (kc ? 16.0 : 8.0) * dk1 * pr.dk2;
        }
        // ---- DMMA kernel for wide real K blocks (rmb_matvec_dmma.cuh): 12 MMA warps = nst states x mt m-tiles of 8
        //      rows, all columns (<= 64 per item) in registers, A fragments formed in registers from the staged ket rows
        if (!force_scalar && !force_nogemm && !kc && dk1 > gemm_min_dk && ndmax <= MV2_NDMAX) {
            const int nc_max = std::min(dk1, 8 * MD_NTMAX);
            int ldk_max = ((nc_max + 7) / 8) * 8 + 4;
            if (ldk_max % 16 != 4) ldk_max += 8;
            int kt_max = 2;
            for (int p = bra_begin[b]; p < bra_begin[b + 1]; ++p)
                kt_max = std::max(kt_max, ((op->h_prods[p].dk2 + 3) & ~3) * ldk_max);
            const int nprod_b = bra_begin[b + 1] - bra_begin[b];
            int best_mt = 0, best_nst = 0;
            double best_util = -1;
            static const int mts[] = {12, 6, 4, 3, 2, 1};          // larger row tiles first: fewer halo rows
            for (int mt : mts) {
                const int nr = 8 * mt;
                if (nr - 7 > dm1 && mt > 1) continue;              // a tile taller than the block: use a smaller mt
                int nst = MD_MMA_WARPS / mt;
                auto need = [&](int nst_) {
                    size_t xb = 16;
                    for (int p = bra_begin[b]; p < bra_begin[b + 1]; ++p) {
                        const ProdD& q = op->h_prods[p];
                        xb = std::max(xb, (size_t)nst_ * std::min(q.dm2, nr + h_span[p]) * (q.dk2 | 1));
                    }
                    return md_smem_bytes((int)xb, mt, kt_max, nprod_b);
                };
                while (nst > 1 && need(nst) > MD_SMEM_MAX) --nst;
                if (need(nst) > MD_SMEM_MAX) continue;
                const int ntile = (dm1 + nr - 1) / nr;
                const double util = (double)dm1 / ((double)nr * ntile) * ((double)nst * mt / MD_MMA_WARPS);
                if (util > best_util + 1e-9) { best_util = util; best_mt = mt; best_nst = nst; }
            }
            if (best_mt > 0) {
                const int nr_t = 8 * best_mt;
                for (int c0 = 0; c0 < dk1; c0 += 8 * MD_NTMAX)
                    for (int r0 = 0; r0 < dm1; r0 += nr_t) {
                        ItemD2 it;
                        it.bra_off = poff[b];
                        it.dk1 = dk1;
                        it.dm1 = dm1;
                        it.r0 = r0;
                        it.nrows = std::min(nr_t, dm1 - r0);
                        it.c0 = c0;
                        it.nc = std::min(8 * MD_NTMAX, dk1 - c0);
                        it.nt = (it.nc + 7) / 8;
                        it.ldk = it.nt * 8 + 4;
                        if (it.ldk % 16 != 4) it.ldk += 8;            // == 4 (mod 16): conflict-free B fragments
                        it.p_begin = bra_begin[b];
                        it.p_end = bra_begin[b + 1];
                        it.nst = best_nst;
                        it.mt = best_mt;
                        it.desc_off = (int)gdesc.size();
                        int xbe = 16, ktd = 2;
                        for (int p = it.p_begin; p < it.p_end; ++p) {
                            const ProdD& q = op->h_prods[p];
                            int lo = q.dm2, hi = -1;
                            for (int r = it.r0; r < it.r0 + it.nrows; ++r)
                                for (int j = 0; j < q.nd; ++j) {
                                    const int col = ent_col[q.ent_off + (long long)r * q.nd + j];
                                    if (col >= 0) { lo = std::min(lo, col); hi = std::max(hi, col); }
                                }
                            ProdS ds;
                            ds.c_lo = hi < 0 ? 0 : lo;
                            ds.nr = hi < 0 ? 0 : hi - lo + 1;
                            ds.ket_off = q.ket_off + (long long)ds.c_lo * (q.dk2 | 1);
                            ds.ent_off = q.ent_off;
                            ds.dk2 = q.dk2;
                            ds.nnz = 0;
                            ds.xrs = q.dk2 | 1;
                            ds.tab = q.tab;
                            ds.mreal = ds.pad2 = 0;
                            gdesc.push_back(ds);
                            xbe = std::max(xbe, it.nst * ds.nr * (q.dk2 | 1));
                            ktd = std::max(ktd, ((q.dk2 + 3) & ~3) * it.ldk);
                        }
                        it.x_elems = xbe;
                        it.kt_doubles = (ktd + 1) & ~1;
                        // K^T images [dk2 padded to 4][ldk] per product, zero padded, shared by the row tiles of this
                        // (bra block, column tile)
                        auto key = std::make_pair(-1 - b, c0);
                        auto found = kt_index.find(key);
                        if (found == kt_index.end()) {
                            while (ktpool.size() % 2) ktpool.push_back(0.0);
                            const long long off = (long long)ktpool.size();
                            kt_index[key] = off;
                            for (int p = it.p_begin; p < it.p_end; ++p) {
                                const ProdD& q = op->h_prods[p];
                                const int k2p = (q.dk2 + 3) & ~3;
                                for (int k2 = 0; k2 < k2p; ++k2)
                                    for (int c = 0; c < it.ldk; ++c)
                                        ktpool.push_back((k2 < q.dk2 && c < it.nc)
                                                             ? kpool[(size_t)(q.koff + (long long)(c0 + c) * q.dk2 + k2)]
                                                             : 0.0);
                            }
                            it.kt_off = off;
                        } else {
                            it.kt_off = found->second;
                        }
                        // the CTA owns the SM anyway (416 threads x 128 registers): small items use the shared memory the large
                        // ones need for a deeper pipeline (their products compute faster than a bulk copy takes to land)
                        it.nstages = MD_STAGES;
                        while (it.nstages < MD_STAGES_MAX && it.nstages < it.p_end - it.p_begin &&
                               md_smem_bytes(it.x_elems, it.mt, it.kt_doubles, it.p_end - it.p_begin, it.nstages + 1) <= MD_SMEM_DEEP)
                            ++it.nstages;
                        if (const char* e = getenv("RMB_DMMA_STAGES")) it.nstages = std::min(it.nstages, std::max(MD_STAGES, atoi(e)));
                        op->matvecG_smem = std::max(op->matvecG_smem, md_smem_bytes(it.x_elems, it.mt, it.kt_doubles,
                                                                                    it.p_end - it.p_begin, it.nstages));
                        itemsG.push_back(it);
                    }
                continue;
            }
        }
        // ---- tiled kernel: row tiles (<= 128 rows, one thread per row and state pair) x column
        //      chunks (<= 16); falls back to the scalar kernel when the tile does not fit
        bool fast = !force_scalar && ndmax <= MV2_NDMAX &&   // diagonal slots handled in registers
                    bra_begin[b + 1] - bra_begin[b] <= 128;
        if (fast) {
            int best_nt = 0, best_nst = 0;
            double best_util = -1;
            const int nt0 = (dm1 + MV2_CONSUMERS - 1) / MV2_CONSUMERS;
            for (int nt = nt0; nt <= nt0 + 3 && nt <= dm1; ++nt) {
                const int nr = (dm1 + nt - 1) / nt;
                int nst = 2 * std::min(MV2_SMAX / 2, MV2_CONSUMERS / nr);   // two states per thread
                // shared memory of the item (exactly what the kernel carves): staging buffers for the ket
                // rows and MF diagonals, K^T of all products, descriptors, state offsets, barriers
                auto need = [&](int nst_) {
                    size_t kt = 0, xb = 256;
                    const int nc0 = std::min(dk1, MV2_NCMAX);
                    const int ncp = nc0 == 1 ? 1 : ((nc0 + 1) & ~1);
                    for (int p = bra_begin[b]; p < bra_begin[b + 1]; ++p) {
                        const ProdD& q = op->h_prods[p];
                        kt += (size_t)q.dk2 * ncp * (kc ? 2 : 1);
                        xb = std::max(xb, (size_t)nst_ * std::min(q.dm2, nr + h_span[p]) * (q.dk2 | 1));
                    }
                    kt = (kt + 1) & ~(size_t)1;
                    return item_smem(xb, nr, kt, bra_begin[b + 1] - bra_begin[b]);
                };
                while (nst > 2 && need(nst) > smem_budget) nst -= 2;
                if (need(nst) > smem_budget) continue;
                const double util = (double)nr * (nst / 2) / MV2_CONSUMERS * ((double)dm1 / (nr * nt));
                if (util > best_util + 1e-9) { best_util = util; best_nt = nt; best_nst = nst; }
            }
            if (best_nt == 0) fast = false;
            if (fast) {
                const int nr_t = (dm1 + best_nt - 1) / best_nt;
                for (int c0 = 0; c0 < dk1; c0 += MV2_NCMAX)
                    for (int r0 = 0; r0 < dm1; r0 += nr_t) {
                        Item2D it;
                        it.bra_off = poff[b];
                        it.dk1 = dk1;
                        it.dm1 = dm1;
                        it.r0 = r0;
                        it.nrows = std::min(nr_t, dm1 - r0);
                        it.c0 = c0;
                        it.nc = std::min(MV2_NCMAX, dk1 - c0);
                        it.p_begin = bra_begin[b];
                        it.p_end = bra_begin[b + 1];
                        it.nst = best_nst;
                        it.xr_off = (int)xranges.size();
                        it.desc_off = (int)gdesc.size();
                        const int ncp = it.nc == 1 ? 1 : ((it.nc + 1) & ~1);
                        const int kw = kc ? 2 : 1;
                        int ktd = 0;
                        int xbe = 256;
                        for (int p = it.p_begin; p < it.p_end; ++p) {
                            const ProdD& q = op->h_prods[p];
                            ktd += q.dk2 * ncp * kw;
                            int lo = q.dm2, hi = -1;
                            for (int r = it.r0; r < it.r0 + it.nrows; ++r)
                                for (int j = 0; j < q.nd; ++j) {
                                    const int col = ent_col[q.ent_off + (long long)r * q.nd + j];
                                    if (col >= 0) { lo = std::min(lo, col); hi = std::max(hi, col); }
                                }
                            XRange xr;
                            xr.c_lo = hi < 0 ? 0 : lo;
                            xr.nr = hi < 0 ? 0 : hi - lo + 1;
                            xranges.push_back(xr);
                            xbe = std::max(xbe, it.nst * xr.nr * (q.dk2 | 1));
                            ProdS ds;
                            ds.ket_off = q.ket_off + (long long)xr.c_lo * (q.dk2 | 1);
                            ds.ent_off = q.ent_off;
                            ds.dk2 = q.dk2;
                            ds.nnz = 0;
                            ds.c_lo = xr.c_lo;
                            ds.nr = xr.nr;
                            ds.xrs = q.dk2 | 1;
                            ds.tab = q.tab;
                            ds.mreal = ds.pad2 = 0;
                            gdesc.push_back(ds);
                        }
                        it.kt_total = (ktd + 1) & ~1;
                        // K^T image [product][k2][ncp] of this (bra block, column chunk), built once
                        auto key = std::make_pair(b, c0);
                        auto found = kt_index.find(key);
                        if (found == kt_index.end()) {
                            const long long off = (long long)ktpool.size();
                            kt_index[key] = off;
                            for (int p = it.p_begin; p < it.p_end; ++p) {
                                const ProdD& q = op->h_prods[p];
                                for (int k2 = 0; k2 < q.dk2; ++k2)
                                    for (int c = 0; c < ncp; ++c)
                                        for (int w = 0; w < kw; ++w)
                                            ktpool.push_back(c < it.nc ? kpool[(size_t)(q.koff + (long long)(c0 + c) * q.dk2 + k2) * kw + w] : 0.0);
                            }
                            while (ktpool.size() % 2) ktpool.push_back(0.0);   // 16-byte aligned images
                            it.kt_off = off;
                        } else {
                            it.kt_off = found->second;
                        }
                        it.xbuf_elems = xbe;
                        // every CTA reserves the maximum over items anyway: small items use it for a deeper pipeline
                        // (their products compute faster than a bulk copy takes to land)
                        it.nstages = MV2_STAGES;
                        while (it.nstages < MV2_STAGES_MAX && it.nstages < it.p_end - it.p_begin &&
                               item_smem_st((size_t)xbe, it.nrows, (size_t)it.kt_total, it.p_end - it.p_begin,
                                            it.nstages + 1) <= smem_budget)
                            ++it.nstages;
                        if (const char* e = getenv("RMB_MV2_STAGES")) it.nstages = std::min(it.nstages, std::max(MV2_STAGES, atoi(e)));
                        op->matvec2_smem = std::max(op->matvec2_smem,
                                                    item_smem_st((size_t)xbe, it.nrows, (size_t)it.kt_total,
                                                                 it.p_end - it.p_begin, it.nstages));
                        items2.push_back(it);
                    }
                continue;
            }
        }
        // ---- scalar kernel items
        const int ncols_t = std::min(dk1, max_out);
        int nrows_t = std::max(1, std::min(dm1, max_out / ncols_t));
        nrows_t = std::max(1, std::min(nrows_t, 1536 / dk2max));
        if ((long long)dk2max * S * 16 > 160 * 1024) {
            set_error("dim_k too large for the shared-memory tile of the matvec kernel");
            return RMB_ERR_INVALID;
        }
        for (int c0 = 0; c0 < dk1; c0 += ncols_t)
            for (int r0 = 0; r0 < dm1; r0 += nrows_t) {
                ItemD it;
                it.bra_off = poff[b];
                it.dk1 = dk1;
                it.r0 = r0;
                it.nrows = std::min(nrows_t, dm1 - r0);
                it.c0 = c0;
                it.ncols = std::min(ncols_t, dk1 - c0);
                it.p_begin = bra_begin[b];
                it.p_end = bra_begin[b + 1];
                it.dk2max = dk2max;
                op->h_items.push_back(it);
                zstride = std::max(zstride, it.nrows * dk2max);
            }
    }
    // heaviest tiled items first
    {
        std::vector<int> order(items2.size());
        std::iota(order.begin(), order.end(), 0);
        auto cost2 = [&](const Item2D& it) {
            double c = 0;
            for (int p = it.p_begin; p < it.p_end; ++p)
                c += (double)it.nrows * op->h_prods[p].dk2 * (it.nc + 2.0 * op->h_prods[p].nd);
            return c;
        };
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost2(items2[a]) > cost2(items2[b]); });
        std::vector<Item2D> sorted;
        for (int i : order) sorted.push_back(items2[i]);
        items2.swap(sorted);
    }
    {
        auto costG = [&](const ItemD2& it) {
            double c = 0;
            for (int p = it.p_begin; p < it.p_end; ++p) c += (double)it.mt * it.nst * op->h_prods[p].dk2 * it.nc;
            return c;
        };
        std::stable_sort(itemsG.begin(), itemsG.end(), [&](const ItemD2& a, const ItemD2& b) { return costG(a) > costG(b); });
    }
    op->nitemsG = (int)itemsG.size();
    for (auto& it : itemsG) op->h_itemG_states.push_back(it.nst);
    op->nitems2 = (int)items2.size();
    for (auto& it : items2) op->h_item2_states.push_back(it.nst);
    // <w,v> partial slots of a state: one per tiled item, then one per (DMMA item, m-tile)
    op->dot_slots = op->nitems2;
    for (auto& it : itemsG) {
        it.pslot = op->dot_slots;
        it.pad2 = 0;
        op->dot_slots += it.mt;
    }
    if (getenv("RMB_DEBUG")) {
        fprintf(stderr, "[rmb] tiled items %d (smem %zu B), DMMA items %d (smem %zu B), scalar items %zu\n",
                op->nitems2, op->matvec2_smem, op->nitemsG, op->matvecG_smem, op->h_items.size());
        for (auto& it : items2)
            fprintf(stderr, "[rmb]   item dm1 %d dk1 %d rows %d nc %d nst %d products %d stages %d xbuf %d B kt %d B\n", it.dm1,
                    it.dk1, it.nrows, it.nc, it.nst, it.p_end - it.p_begin, it.nstages, it.xbuf_elems * 16, it.kt_total * 8);
    }

    opbytes += 20.0 * (double)op->nent;   // MF values + column indices
    op->flops_per_state = flops;
    op->op_bytes = opbytes;
    // heaviest items first (tail balance)
    std::stable_sort(op->h_items.begin(), op->h_items.end(), [&](const ItemD& a, const ItemD& b) {
        auto cost = [&](const ItemD& it) {
            double c = 0;
            for (int p = it.p_begin; p < it.p_end; ++p)
                c += (double)it.nrows * op->h_prods[p].dk2 * (it.ncols + 2.0 * op->h_prods[p].nd);
            return c;
        };
        return cost(a) > cost(b);
    });
    op->nitems = (int)op->h_items.size();
    op->matvec_smem = (size_t)S * zstride * sizeof(cplx);
    int rc;
    if ((rc = upload(&op->d_prods, op->h_prods.data(), op->h_prods.size()))) return rc;
    if ((rc = upload(&op->d_items, op->h_items.data(), op->h_items.size()))) return rc;
    if ((rc = upload((Item2D**)&op->d_items2, items2.data(), items2.size()))) return rc;
    // fused single-launch step: every block has dim_k = 1 and three vectors fit in shared memory
    {
        bool all1 = true;
        for (int b = 0; b < d->nblocks; ++b) all1 = all1 && d->blk_dk[b] == 1;
        const size_t fused_smem = (size_t)3 * op->n * sizeof(cplx) + (size_t)op->nprod * sizeof(FusedProd);
        if (all1 && fused_smem <= 210 * 1024 && op->nd_max <= 32 && op->nent < (1LL << 31)) {
            std::vector<int> row_blk((size_t)op->n);
            std::vector<long long> boff(d->nblocks);
            for (int b = 0; b < d->nblocks; ++b) {
                boff[b] = d->blk_off[b];
                for (long long i = d->blk_off[b]; i < d->blk_off[b + 1]; ++i) row_blk[(size_t)i] = b;
            }
            if ((rc = upload(&op->d_row_blk, row_blk.data(), row_blk.size()))) return rc;
            if ((rc = upload(&op->d_blk_begin, bra_begin.data(), bra_begin.size()))) return rc;
            if ((rc = upload(&op->d_blk_off, boff.data(), boff.size()))) return rc;
            if ((rc = upload(&op->d_blk_dm, d->blk_dm, (size_t)d->nblocks))) return rc;
            if (fused_smem > g_fused_smem) {
                RMB_CUDA(cudaFuncSetAttribute(k_lanczos_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_smem));
                g_fused_smem = fused_smem;
            }
            const char* ff = getenv("RMB_FUSED");
            op->fused_ok = !(ff && strcmp(ff, "0") == 0);
        }
        // sliding-window matvec for large ensembles of linear rotors
        int W = 0, dmmax = 1;
        for (size_t p = 0; p < h_prod_ket.size(); ++p) W = std::max(W, std::abs(h_prod_ket[p] - h_prod_bra[p]));
        for (int b = 0; b < d->nblocks; ++b) dmmax = std::max(dmmax, d->blk_dm[b]);
        dmmax |= 1;                                               // odd row stride: conflict-free lanes
        const char* fl = getenv("RMB_LIN");
        int maxL = 0;
        for (int b = 0; b < d->nblocks; ++b) {
            int l = 0;
            for (int p = bra_begin[b]; p < bra_begin[b + 1]; ++p) l += std::min(op->h_prods[p].nd, MV2_NDMAX);
            maxL = std::max(maxL, l);
        }
        if (all1 && W <= 3 && op->nd_max <= MV2_NDMAX && op->nprod > 0 && maxL <= ML_LMAX &&
            !(fl && strcmp(fl, "0") == 0)) {
            // per-block entry capacity (all diagonals alive) and offsets of the folded K*MF values
            std::vector<long long> val_off(d->nblocks + 1, 0);
            long long ebuf_elems = 1;
            for (int b = 0; b < d->nblocks; ++b) {
                int l = 0;
                for (int p = bra_begin[b]; p < bra_begin[b + 1]; ++p) l += std::min(op->h_prods[p].nd, MV2_NDMAX);
                const long long cap = (long long)l * d->blk_dm[b];
                val_off[b + 1] = val_off[b] + cap;
                ebuf_elems = std::max(ebuf_elems, cap);
            }
            // T = 8 states per CTA (two groups of 4 per thread) with a ring of up to 2W+4 ket blocks (prefetch depth)
            // (RMB_LIN_T = 16 / 8 / 4 starts the search at that tile: testing and A/B runs)
            const char* ft = getenv("RMB_LIN_T");
            const int t_first = ft ? atoi(ft) : 8;
            for (int pass = t_first >= 16 ? 0 : (t_first >= 8 ? 3 : 6); pass < 9 && !op->lin_ok; ++pass) {
                const int T = pass < 3 ? 16 : (pass < 6 ? 8 : 4);
                const int NS = 2 * W + 4 - pass % 3;
                const size_t fixed = (size_t)NS * T * dmmax * 16 + (size_t)(2 * NS + ML_NBMAX + 1) * 8 +
                                     (size_t)d->nblocks * sizeof(LinBlk) + 128;
                const size_t need = fixed + (size_t)2 * (ebuf_elems * 16 + ML_FLAT * sizeof(LinEnt));
                if (need <= LIN_SMEM_MAX) {
                    op->lin_smem_fixed = fixed;
                    op->lin_ok = true;
                    op->lin_W = W;
                    op->lin_T = T;
                    op->lin_NS = NS;
                    op->lin_dm_max = dmmax;
                    op->lin_smem = need;
                    op->lin_ebuf = (int)ebuf_elems;
                }
            }
            if (op->lin_ok && op->nent >= (1ll << 31)) op->lin_ok = false;
            op->fused_lcap = op->lin_ok ? maxL : 0;
            if (op->lin_ok) {
                if (!op->d_blk_begin) {
                    std::vector<long long> boff(d->nblocks);
                    for (int b = 0; b < d->nblocks; ++b) boff[b] = d->blk_off[b];
                    if ((rc = upload(&op->d_blk_begin, bra_begin.data(), bra_begin.size()))) return rc;
                    if ((rc = upload(&op->d_blk_off, boff.data(), boff.size()))) return rc;
                    if ((rc = upload(&op->d_blk_dm, d->blk_dm, (size_t)d->nblocks))) return rc;
                }
                if ((rc = upload(&op->d_prod_ket, h_prod_ket.data(), h_prod_ket.size()))) return rc;
                std::vector<LinBlk> lb(d->nblocks);
                int chunk0 = 0, ubase = 0;
                op->lin_g1 = getenv("RMB_LIN_G1") && atoi(getenv("RMB_LIN_G1")) == 1;     // soak / sanitizer runs only
                const int G = op->lin_g1 ? 1 : op->lin_T / (op->lin_T <= ML_TS ? op->lin_T / 2 : ML_TS);
                for (int b = 0; b < d->nblocks; ++b) {
                    const int nch = (d->blk_dm[b] + 31) / 32;
                    lb[b].off = poff[b];
                    lb[b].val_off = val_off[b];
                    lb[b].dm = d->blk_dm[b];
                    lb[b].chunk0 = chunk0;
                    lb[b].ubase = ubase;
                    lb[b].L = 0;
                    chunk0 += nch;
                    ubase = (ubase + nch * G) % ML_CWARPS;
                }
                op->lin_npart = ML_CWARPS / G;         // <w,v> partials per state: one per warp of the state's group
                if ((rc = upload((LinBlk**)&op->d_lin_blk, lb.data(), lb.size()))) return rc;
                RMB_CUDA(cudaMalloc(&op->d_lin_flat, (size_t)d->nblocks * ML_FLAT * sizeof(LinEnt)));
                if ((rc = upload(&op->d_lin_val_off, val_off.data(), val_off.size()))) return rc;   // k_lin_entries
                RMB_CUDA(cudaMalloc((void**)&op->d_lin_val, (size_t)std::max<long long>(1, val_off[d->nblocks]) * sizeof(cplx)));
                static bool g_lin_attr = false;
                if (!g_lin_attr) {
                    RMB_CUDA(cudaFuncSetAttribute(k_matvec_linw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LIN_SMEM_MAX));
                    RMB_CUDA(cudaFuncSetAttribute(k_matvec_lin<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LIN_SMEM_MAX));
                    RMB_CUDA(cudaFuncSetAttribute(k_matvec_lin<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LIN_SMEM_MAX));
                    RMB_CUDA(cudaFuncSetAttribute(k_matvec_lin<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LIN_SMEM_MAX));
                    RMB_CUDA(cudaFuncSetAttribute(k_matvec_lin<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LIN_SMEM_MAX));
                    RMB_CUDA(cudaFuncSetAttribute(k_matvec_lin<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LIN_SMEM_MAX));
                    g_lin_attr = true;
                }
                op->h_bra_begin = bra_begin;
                op->h_blk_dm.assign(d->blk_dm, d->blk_dm + d->nblocks);
                // ---- register-window kernel: blocks ordered by J with symmetric contiguous m ranges (rows grow by an even
                //      number from block to block), block distance <= 2; per (product, diagonal) the m offset it couples
                {
                    // Both register-window kernels are opt-in (RMB_MW=1: k_matvec_mw, RMB_LINW=1: k_matvec_linw).  They are correct
                    // -- the whole GPU suite passes with either as the default -- but issue 3.4x the instructions of the ring
                    // kernel for the same work and are slower as they stand: 0.92 / 1.24-1.34 ms against 0.39 ms on the OCS
                    // batch (profiles/r02_lin_notes.md)
                    bool ok = W <= MW_DB;
                    std::vector<int> cshift(d->nblocks, 0);
                    for (int b = 0; b + 1 < d->nblocks; ++b) {
                        const int diff = d->blk_dm[b + 1] - d->blk_dm[b];
                        if (diff < 0 || (diff & 1)) ok = false;
                        cshift[b + 1] = cshift[b] + diff / 2;
                    }
                    op->h_prod_dm_off.assign(op->h_prods.size() + 1, 0);
                    for (size_t p = 0; p < op->h_prods.size(); ++p) {
                        const ProdD& q = op->h_prods[p];
                        op->h_prod_dm_off[p] = (int)op->h_prod_dm.size();
                        const int dm1p = op->h_prod_dm1[p];
                        for (int j = 0; j < q.nd; ++j) {
                            int doff = 0;
                            bool found = false;
                            for (int r = 0; r < dm1p && !found; ++r) {
                                const int col = ent_col[q.ent_off + (long long)r * q.nd + j];
                                if (col >= 0) { doff = col - r; found = true; }
                            }
                            const int dmq = doff - (cshift[h_prod_ket[p]] - cshift[h_prod_bra[p]]);
                            op->h_prod_dm.push_back((signed char)std::max(-100, std::min(100, found ? dmq : 0)));
                        }
                    }
                    op->h_prod_dm_off[op->h_prods.size()] = (int)op->h_prod_dm.size();
                    op->sym_ok = ok;
                    op->mw_static = ok && getenv("RMB_MW") && atoi(getenv("RMB_MW")) == 1;
                    op->lw_static = ok && d->blk_dm[d->nblocks - 1] <= 8 * LW_CWARPS && op->lin_NS >= W + 3 &&
                                    getenv("RMB_LINW") && atoi(getenv("RMB_LINW")) == 1;
                    op->lw_groups = (d->blk_dm[d->nblocks - 1] + 7) / 8;
                    op->mw_groups = (d->blk_dm[d->nblocks - 1] + 3) / 4;
                    op->h_cshift = cshift;
                    if ((rc = upload(&op->d_cshift, cshift.data(), cshift.size()))) return rc;
                    RMB_CUDA(cudaMalloc((void**)&op->d_cmap, (size_t)d->nblocks * 16));
                    RMB_CUDA(cudaMemset(op->d_cmap, 0xff, (size_t)d->nblocks * 16));
                    RMB_CUDA(cudaMalloc((void**)&op->d_mw_counter, sizeof(int)));
                    // first active block of every m-group
                    op->h_mw_bfirst.assign(op->mw_groups, d->nblocks);
                    for (int g = 0; g < op->mw_groups; ++g)
                        for (int b = 0; b < d->nblocks; ++b) {
                            bool any = false;
                            for (int qq = 0; qq < 4; ++qq) {
                                const int r = g * 4 + qq - (cshift[d->nblocks - 1] - cshift[b]);
                                any = any || (r >= 0 && r < d->blk_dm[b] && g * 4 + qq < d->blk_dm[d->nblocks - 1]);
                            }
                            if (any) { op->h_mw_bfirst[g] = b; break; }
                        }
                }
                lin_update_bound(op);
            }
        }
    }
    if ((rc = upload((ItemD2**)&op->d_itemsG, itemsG.data(), itemsG.size()))) return rc;
    {
        static size_t g_gemm_smem = 16 * 1024;
        if (op->matvecG_smem > g_gemm_smem) {
            RMB_CUDA(cudaFuncSetAttribute(k_matvec_dmma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)op->matvecG_smem));
            g_gemm_smem = op->matvecG_smem;
        }
    }
    if ((rc = upload((ProdS**)&op->d_gdesc, gdesc.data(), gdesc.size()))) return rc;
    op->ngdesc = (int)gdesc.size();
    ktpool.push_back(0.0);
    ktpool.push_back(0.0);
    if ((rc = upload(&op->d_ktpool, ktpool.data(), ktpool.size()))) return rc;
    // the opt-in limit is a per-function, process-wide attribute: only ever raise it
    static size_t g_tiled_smem = 48 * 1024;
    if (op->matvec2_smem > g_tiled_smem) {
        RMB_CUDA(cudaFuncSetAttribute(k_matvec_tiled<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)op->matvec2_smem));
        RMB_CUDA(cudaFuncSetAttribute(k_matvec_tiled<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)op->matvec2_smem));
        g_tiled_smem = op->matvec2_smem;
    }
    if ((rc = upload(&op->d_ent_col, ent_col.data(), ent_col.size()))) return rc;
    if ((rc = upload<cplx>(&op->d_ent_val, nullptr, (size_t)op->nent))) return rc;
    if ((rc = upload<double>(&op->d_ent_cent, nullptr, (size_t)op->nent * 4 + 4))) return rc;
    RMB_CUDA(cudaMemset(op->d_ent_cent, 0, ((size_t)op->nent * 4 + 4) * sizeof(double)));
    op->ntab = (int)tab_off.size() - 1;
    if ((rc = upload(&op->d_ent_tab, ent_tab.data(), ent_tab.size()))) return rc;
    if ((rc = upload(&op->d_tab_off, tab_off.data(), tab_off.size()))) return rc;
    if ((rc = upload(&op->d_tab_nd, tab_nd.data(), tab_nd.size()))) return rc;
    // [0, ntab]: masks of the non-zero diagonals; [ntab + 1, 2 ntab + 1]: table has an MF entry with a non-zero
    // imaginary part (fields in the XZ plane give real MF: the matvec then skips half of its z-stage)
    if ((rc = upload<unsigned>(&op->d_tab_mask, nullptr, 2 * ((size_t)op->ntab + 1)))) return rc;
    RMB_CUDA(cudaMemset(op->d_tab_mask, 0, 2 * ((size_t)op->ntab + 1) * sizeof(unsigned)));
    op->d_tab_cplx = op->d_tab_mask + op->ntab + 1;
    RMB_CUDA(cudaMemset(op->d_ent_val, 0, std::max<size_t>(1, (size_t)op->nent) * sizeof(cplx)));
    if ((rc = upload(&op->d_kpool, kpool.data(), kpool.size()))) return rc;
    if ((rc = upload<int>(&op->d_flags, nullptr, (size_t)d->nparts + 1))) return rc;
    RMB_CUDA(cudaMemset(op->d_flags, 0, ((size_t)d->nparts + 1) * sizeof(int)));
    static size_t g_scalar_smem = 48 * 1024;
    if (op->matvec_smem > g_scalar_smem) {
        RMB_CUDA(cudaFuncSetAttribute(k_matvec_scalar<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)op->matvec_smem));
        RMB_CUDA(cudaFuncSetAttribute(k_matvec_scalar<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)op->matvec_smem));
        g_scalar_smem = op->matvec_smem;
    }
    guard.p = nullptr;
    *out = op;
    return RMB_OK;
}

int64_t rmb_operator_dim(const rmb_operator* op) { return op ? op->n : 0; }

int64_t rmb_operator_nentries(const rmb_operator* op, int32_t part) {
    if (!op || part < 0 || part >= (int)op->parts.size()) return -1;
    return op->parts[part].ent_end - op->parts[part].ent_begin;
}

int32_t rmb_operator_set_field(rmb_operator* op, int32_t part, const double* fprod, double thresh,
                               int32_t all_dropped, void* stream) {
    if (!op || part < 0 || part >= (int)op->parts.size() || !fprod) {
        set_error("set_field: bad operator part");
        return RMB_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    PartH& ph = op->parts[part];
    FieldProds fp;
    const double* fdev = nullptr;
    if (ph.ncart <= 16) {
        for (int c = 0; c < 16; ++c) fp.f[c] = c < ph.ncart ? fprod[c] : 0.0;
    } else {
        RMB_CUDA(cudaMemcpyAsync(ph.d_fprod, fprod, sizeof(double) * ph.ncart, cudaMemcpyHostToDevice, st));
        fdev = ph.d_fprod;
    }
    const long long nent = ph.ent_end - ph.ent_begin;
    RMB_CUDA(cudaMemsetAsync(op->d_flags + 1 + part, 0, sizeof(int), st));
    if (ph.tab_end > ph.tab_begin)
    {
        RMB_CUDA(cudaMemsetAsync(op->d_tab_mask + ph.tab_begin, 0, sizeof(unsigned) * (ph.tab_end - ph.tab_begin), st));
        RMB_CUDA(cudaMemsetAsync(op->d_tab_cplx + ph.tab_begin, 0, sizeof(unsigned) * (ph.tab_end - ph.tab_begin), st));
    }
    if (nent > 0) {
        const int nt = 256;
        k_field_contract<<<(unsigned)((nent + nt - 1) / nt), nt, 0, st>>>(
            nent, ph.ncart, ph.d_coef, fdev, fp, thresh, all_dropped, op->d_ent_val + ph.ent_begin,
            op->d_flags + 1 + part, op->d_ent_tab, op->d_tab_off, op->d_tab_nd, op->d_tab_mask, op->d_tab_cplx,
            ph.ent_begin);
        k_compact_tables<<<(unsigned)((nent + nt - 1) / nt), nt, 0, st>>>(
            nent, ph.ent_begin, op->d_ent_val, op->d_ent_col, op->d_ent_tab, op->d_tab_off, op->d_tab_nd,
            op->d_tab_mask, op->d_ent_cent);
        RMB_CUDA(cudaGetLastError());
        op->n_launches += 2;
        op->lin_flat_dirty = true;
    }
    op->nnz_dirty = true;
    ph.has_field = true;
    {
        unsigned nz = 0;
        for (int c = 0; c < ph.ncart && c < 32; ++c)
            if (fprod[c] != 0.0) nz |= 1u << c;
        ph.nzmask = nz;
    }
    ph.all_dropped = all_dropped != 0;
    op->lin_flat_dirty = true;
    return RMB_OK;
}

int32_t rmb_operator_get_mf(rmb_operator* op, int32_t part, double* out_host, void* stream) {
    if (!op || part < 0 || part >= (int)op->parts.size() || !out_host) {
        set_error("get_mf: bad operator part");
        return RMB_ERR_INVALID;
    }
    PartH& ph = op->parts[part];
    if (!ph.has_field) {
        set_error("operator part has no field applied");
        return RMB_ERR_NOFIELD;
    }
    cudaStream_t st = (cudaStream_t)stream;
    RMB_CUDA(cudaMemcpyAsync(out_host, op->d_ent_val + ph.ent_begin,
                             sizeof(cplx) * (ph.ent_end - ph.ent_begin), cudaMemcpyDeviceToHost, st));
    RMB_CUDA(cudaStreamSynchronize(st));
    return RMB_OK;
}

int32_t rmb_operator_mf_nonempty(rmb_operator* op, void* stream) {
    if (!op) return RMB_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<int> flags(op->parts.size() + 1, 0);
    RMB_CUDA(cudaMemcpyAsync(flags.data(), op->d_flags, sizeof(int) * flags.size(), cudaMemcpyDeviceToHost, st));
    RMB_CUDA(cudaStreamSynchronize(st));
    for (size_t q = 0; q < op->parts.size(); ++q)
        if (op->parts[q].has_field && flags[1 + q]) return 1;
    return 0;
}

}  // extern "C"

namespace rmb {

static int check_field(rmb_operator* op) {
    for (auto& p : op->parts)
        if (!p.has_field) {
            set_error("you need to multiply tensor with field before applying it to a vector");
            return RMB_ERR_NOFIELD;
        }
    return RMB_OK;
}

struct MvEpilogue {
    const double* scale = nullptr;   // per-state factor applied to the product (rinv_k), stride in doubles
    int scale_stride = 0;
    cplx* pdot = nullptr;            // fused partial sums conj(y) * x per (state, tiled item)
    int npart = 0;
    bool use_lin = false;            // sliding-window kernel for linear rotors (npart = nblocks)
};

// Field-dependent tables of the matvec kernels, rebuilt on `st` after a field update: the entry lists of the sliding-window
// kernel, or the surviving-diagonal counts in the descriptors of the tiled / DMMA kernels.  launch_matvec calls it; the
// host-buffer pipeline calls it once on the caller's stream before it forks onto several compute streams.
static int matvec_prep(rmb_operator* op, cudaStream_t st, bool use_lin) {
    if (use_lin && op->lin_ok) {
        if (op->lin_flat_dirty) {
            lin_update_bound(op);
            k_lin_entries<<<(unsigned)op->nblocks, 128, 0, st>>>(
                op->nblocks, op->lin_NS, (unsigned)(op->lin_T * op->lin_dm_max * 16), op->d_blk_begin, op->d_blk_dm,
                op->d_prod_ket, op->d_prods, op->d_tab_mask, (const MfEntry*)op->d_ent_cent, op->d_kpool,
                op->k_complex ? 1 : 0, op->d_lin_val_off, (LinEnt*)op->d_lin_flat, op->d_lin_val,
                op->sym_ok ? op->d_cshift : nullptr, op->mw_static ? op->d_cmap : nullptr);
            op->lin_flat_dirty = false;
            op->n_launches++;
        }
    } else if (op->nnz_dirty && op->ngdesc > 0) {
        // surviving diagonals per (item, product) descriptor of the tiled and DMMA kernels, after a field update
        k_fill_nnz<<<(unsigned)((op->ngdesc + 255) / 256), 256, 0, st>>>(op->ngdesc, (ProdS*)op->d_gdesc, op->d_tab_mask, op->d_tab_cplx);
        op->nnz_dirty = false;
        op->n_launches++;
    }
    RMB_CUDA(cudaGetLastError());
    return RMB_OK;
}

static int launch_matvec(rmb_operator* op, const cplx* X, cplx* Y, long long nstates, long long ldx,
                         long long ldy, const int* active, cudaStream_t st, const MvEpilogue& ep = MvEpilogue()) {
    if ((op->nitems == 0 && op->nitems2 == 0 && op->nitemsG == 0) || nstates == 0) return RMB_OK;
    {
        const int rcp = matvec_prep(op, st, ep.use_lin);
        if (rcp) return rcp;
    }
    const int S = op->matvec_S;
    const int zstride = (int)(op->matvec_smem / (S * sizeof(cplx)));
    std::pair<cudaEvent_t, cudaEvent_t> ev;
    if (op->time_matvec) {
        if (!op->mv_event_pool.empty()) {
            ev = op->mv_event_pool.back();
            op->mv_event_pool.pop_back();
        } else {
            RMB_CUDA(cudaEventCreate(&ev.first));
            RMB_CUDA(cudaEventCreate(&ev.second));
        }
        RMB_CUDA(cudaEventRecord(ev.first, st));
    }
    if (ep.use_lin && op->lin_ok && op->lw_cur) {
        // ring + register window (k_matvec_linw): tiles of 4 states
        LinArgs la;
        la.nblocks = op->nblocks;
        la.W = op->lin_W;
        la.dms = op->lin_dm_max;
        la.NS = op->lw_NS;
        la.NB = op->lw_NB;
        la.ebuf_elems = op->lin_ebuf_cur;
        la.blk = (const LinBlk*)op->d_lin_blk;
        la.flat = (const LinEnt*)op->d_lin_flat;
        la.val = op->d_lin_val;
        const unsigned grid = (unsigned)((nstates + LW_T - 1) / LW_T);
        k_matvec_linw<<<grid, LW_THREADS, op->lw_smem, st>>>(la, op->h_blk_dm[op->nblocks - 1], op->lw_groups, X, Y, ldx, ldy,
                                                             (int)nstates, active, ep.scale, ep.scale_stride, ep.pdot,
                                                             ep.npart);
        op->n_launches++;
        RMB_CUDA(cudaGetLastError());
        if (op->time_matvec) {
            RMB_CUDA(cudaEventRecord(ev.second, st));
            op->mv_events.push_back(ev);
        }
        op->n_matvec_launches++;
        return RMB_OK;
    }
    if (ep.use_lin && op->lin_ok && op->mw_cur) {
        // register-window kernel: items (m-group, 64-state super tile), longest walks first
        auto hit = op->mw_items_cache.find(nstates);
        if (hit == op->mw_items_cache.end()) {
            std::vector<MwItem> items;
            std::vector<int> order(op->mw_groups);
            std::iota(order.begin(), order.end(), 0);
            std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return op->h_mw_bfirst[x] < op->h_mw_bfirst[y]; });
            for (int g : order)
                for (long long t0 = 0; t0 < nstates; t0 += 8 * MW_WARPS) items.push_back({g, (int)t0, op->h_mw_bfirst[g], 0});
            if (op->mw_items_cache.size() >= 16) {
                RMB_CUDA(cudaStreamSynchronize(st));
                for (auto& kv : op->mw_items_cache) cudaFree(kv.second.first);
                op->mw_items_cache.clear();
            }
            void* dptr = nullptr;
            RMB_CUDA(cudaMalloc(&dptr, sizeof(MwItem) * std::max<size_t>(1, items.size())));
            RMB_CUDA(cudaMemcpyAsync(dptr, items.data(), sizeof(MwItem) * items.size(), cudaMemcpyHostToDevice, st));
            RMB_CUDA(cudaStreamSynchronize(st));
            hit = op->mw_items_cache.emplace(nstates, std::make_pair(dptr, (int)items.size())).first;
        }
        MwArgs ma;
        ma.nblocks = op->nblocks;
        ma.nitems = hit->second.second;
        ma.dm_last = op->h_blk_dm[op->nblocks - 1];
        ma.blk_off = op->d_blk_off;
        ma.blk_dm = op->d_blk_dm;
        ma.cshift = op->d_cshift;
        ma.val_off = op->d_lin_val_off;
        ma.cmap = op->d_cmap;
        ma.val = op->d_lin_val;
        ma.items = (const MwItem*)hit->second.first;
        ma.counter = op->d_mw_counter;
        RMB_CUDA(cudaMemsetAsync(op->d_mw_counter, 0, sizeof(int), st));
        const unsigned grid = (unsigned)std::min<long long>(ma.nitems, 2LL * op->num_sms);
        k_matvec_mw<<<grid, MW_THREADS, mw_smem_bytes(op->nblocks), st>>>(ma, X, Y, ldx, ldy, (int)nstates, active, ep.scale,
                                                                          ep.scale_stride, ep.pdot, ep.npart);
        op->n_launches++;
        RMB_CUDA(cudaGetLastError());
        if (op->time_matvec) {
            RMB_CUDA(cudaEventRecord(ev.second, st));
            op->mv_events.push_back(ev);
        }
        op->n_matvec_launches++;
        return RMB_OK;
    }
    if (ep.use_lin && op->lin_ok) {
        LinArgs la;
        la.nblocks = op->nblocks;
        la.W = op->lin_W;
        la.dms = op->lin_dm_max;
        la.NS = op->lin_NS;
        la.NB = op->lin_NB;
        la.ebuf_elems = op->lin_ebuf_cur;
        la.blk = (const LinBlk*)op->d_lin_blk;
        la.flat = (const LinEnt*)op->d_lin_flat;
        la.val = op->d_lin_val;
        const int T = op->lin_T;
        const unsigned grid = (unsigned)((nstates + T - 1) / T);
        if (op->lin_g1 && T == 8)
            k_matvec_lin<8, true><<<grid, ML_THREADS, op->lin_smem, st>>>(la, X, Y, ldx, ldy, (int)nstates, active, ep.scale,
                                                                          ep.scale_stride, ep.pdot, ep.npart);
        else if (op->lin_g1)
            k_matvec_lin<4, true><<<grid, ML_THREADS, op->lin_smem, st>>>(la, X, Y, ldx, ldy, (int)nstates, active, ep.scale,
                                                                          ep.scale_stride, ep.pdot, ep.npart);
        else if (T == 16)
            k_matvec_lin<16><<<grid, ML_THREADS, op->lin_smem, st>>>(la, X, Y, ldx, ldy, (int)nstates, active, ep.scale,
                                                                     ep.scale_stride, ep.pdot, ep.npart);
        else if (T == 8)
            k_matvec_lin<8><<<grid, ML_THREADS, op->lin_smem, st>>>(la, X, Y, ldx, ldy, (int)nstates, active, ep.scale,
                                                                    ep.scale_stride, ep.pdot, ep.npart);
        else
            k_matvec_lin<4><<<grid, ML_THREADS, op->lin_smem, st>>>(la, X, Y, ldx, ldy, (int)nstates, active, ep.scale,
                                                                    ep.scale_stride, ep.pdot, ep.npart);
        op->n_launches++;
        RMB_CUDA(cudaGetLastError());
        if (op->time_matvec) {
            RMB_CUDA(cudaEventRecord(ev.second, st));
            op->mv_events.push_back(ev);
        }
        op->n_matvec_launches++;
        return RMB_OK;
    }
    if (op->nitems2 > 0) {
        // work units (item, first state) for this batch size; rebuilt only when the size changes
        if (op->units_nstates != nstates) {
            auto hit = op->units_cache.find(nstates);
            if (hit == op->units_cache.end()) {
                // a CTA walks `tiles` consecutive state tiles of one item (K^T image, descriptors and pipeline set up
                // once); fewer, longer CTAs as long as the grid still fills the 2 x 148 CTA slots several times over
                std::vector<Unit2D> units;
                long long ntile_total = 0;
                for (int i = 0; i < op->nitems2; ++i) ntile_total += (nstates + op->h_item2_states[i] - 1) / op->h_item2_states[i];
                int tiles = (int)std::max<long long>(1, std::min<long long>(MV2_TILES_MAX, ntile_total / (6 * 2 * 148)));
                if (const char* e = getenv("RMB_MV2_TILES")) tiles = std::max(1, std::min(MV2_TILES_MAX, atoi(e)));
                for (int i = 0; i < op->nitems2; ++i) {
                    const long long nst = op->h_item2_states[i];
                    const int tl = (int)std::min<long long>(tiles, std::max<long long>(1, 256 / nst));   // one flag thread per state
                    for (long long s0 = 0; s0 < nstates; s0 += nst * tl)
                        units.push_back({i, (int)s0, (int)std::min<long long>(tl, (nstates - s0 + nst - 1) / nst), 0});
                }
                // the lists of the batch sizes in use are kept (the chunks of the host-buffer pipeline and the
                // sub-batches of a large ensemble alternate between two or three sizes)
                if (op->units_cache.size() >= 16) {
                    RMB_CUDA(cudaStreamSynchronize(st));
                    for (auto& kv : op->units_cache) cudaFree(kv.second.first);
                    op->units_cache.clear();
                }
                void* d = nullptr;
                RMB_CUDA(cudaMalloc(&d, sizeof(Unit2D) * std::max<size_t>(1, units.size())));
                RMB_CUDA(cudaMemcpyAsync(d, units.data(), sizeof(Unit2D) * units.size(), cudaMemcpyHostToDevice, st));
                RMB_CUDA(cudaStreamSynchronize(st));   // `units` is pageable host memory
                hit = op->units_cache.emplace(nstates, std::make_pair(d, (int)units.size())).first;
            }
            op->d_units = hit->second.first;
            op->nunits = hit->second.second;
            op->units_nstates = nstates;
        }
        if (op->k_complex)
            k_matvec_tiled<true><<<op->nunits, MV2_THREADS, op->matvec2_smem, st>>>(
                (const Unit2D*)op->d_units, (const Item2D*)op->d_items2, (const ProdS*)op->d_gdesc,
                (const MfEntry*)op->d_ent_cent, op->d_tab_mask, op->d_ktpool, X, Y, ldx, ldy, (int)nstates, active,
                ep.scale, ep.scale_stride, ep.pdot, ep.npart);
        else
            k_matvec_tiled<false><<<op->nunits, MV2_THREADS, op->matvec2_smem, st>>>(
                (const Unit2D*)op->d_units, (const Item2D*)op->d_items2, (const ProdS*)op->d_gdesc,
                (const MfEntry*)op->d_ent_cent, op->d_tab_mask, op->d_ktpool, X, Y, ldx, ldy, (int)nstates, active,
                ep.scale, ep.scale_stride, ep.pdot, ep.npart);
        op->n_launches++;
    }
    if (op->nitemsG > 0) {
        if (op->unitsG_nstates != nstates) {
            auto hit = op->unitsG_cache.find(nstates);
            if (hit == op->unitsG_cache.end()) {
                // a CTA walks `tl` consecutive state tiles of one item (one CTA per SM: descriptors, barriers and the
                // producer pipeline are set up once); fewer, longer CTAs while the grid still fills the SMs ~6 times
                std::vector<Unit2D> units;
                long long ntile_total = 0;
                for (int i = 0; i < op->nitemsG; ++i) ntile_total += (nstates + op->h_itemG_states[i] - 1) / op->h_itemG_states[i];
                int tl = (int)std::max<long long>(1, std::min<long long>(MV2_TILES_MAX, ntile_total / (6 * op->num_sms)));
                if (const char* e = getenv("RMB_DMMA_TILES")) tl = std::max(1, std::min(MV2_TILES_MAX, atoi(e)));
                for (int i = 0; i < op->nitemsG; ++i) {
                    const long long nst = op->h_itemG_states[i];
                    for (long long s0 = 0; s0 < nstates; s0 += nst * tl)
                        units.push_back({i, (int)s0, (int)std::min<long long>(tl, (nstates - s0 + nst - 1) / nst), 0});
                }
                if (op->unitsG_cache.size() >= 16) {
                    RMB_CUDA(cudaStreamSynchronize(st));
                    for (auto& kv : op->unitsG_cache) cudaFree(kv.second.first);
                    op->unitsG_cache.clear();
                }
                void* d = nullptr;
                RMB_CUDA(cudaMalloc(&d, sizeof(Unit2D) * std::max<size_t>(1, units.size())));
                RMB_CUDA(cudaMemcpyAsync(d, units.data(), sizeof(Unit2D) * units.size(), cudaMemcpyHostToDevice, st));
                RMB_CUDA(cudaStreamSynchronize(st));
                hit = op->unitsG_cache.emplace(nstates, std::make_pair(d, (int)units.size())).first;
            }
            op->d_unitsG = hit->second.first;
            op->nunitsG = hit->second.second;
            op->unitsG_nstates = nstates;
        }
        k_matvec_dmma<<<op->nunitsG, MD_THREADS, op->matvecG_smem, st>>>(
            (const Unit2D*)op->d_unitsG, (const ItemD2*)op->d_itemsG, (const ProdS*)op->d_gdesc,
            (const MfEntry*)op->d_ent_cent, op->d_ktpool, X, Y, ldx, ldy, (int)nstates, active,
            ep.scale, ep.scale_stride, ep.pdot, ep.npart, op->nitems2);
        op->n_launches++;
    }
    // scalar kernel for the bra blocks the tiled kernel does not cover (no fused epilogue: callers
    // check op->nitems == 0 before asking for one); grid.y is limited to 65535
    const long long max_y = 65535;
    for (long long s0 = 0; op->nitems > 0 && Y != nullptr && s0 < nstates; s0 += max_y * S) {
        const long long ns = std::min(nstates - s0, max_y * S);
        dim3 grid((unsigned)op->nitems, (unsigned)((ns + S - 1) / S));
        if (op->k_complex)
            k_matvec_scalar<true><<<grid, MV_THREADS, op->matvec_smem, st>>>(
                op->d_items, op->d_prods, op->d_ent_col, op->d_ent_val, op->d_kpool, X + s0 * ldx,
                Y + s0 * ldy, ldx, ldy, (int)ns, S, active ? active + s0 : nullptr, zstride);
        else
            k_matvec_scalar<false><<<grid, MV_THREADS, op->matvec_smem, st>>>(
                op->d_items, op->d_prods, op->d_ent_col, op->d_ent_val, op->d_kpool, X + s0 * ldx,
                Y + s0 * ldy, ldx, ldy, (int)ns, S, active ? active + s0 : nullptr, zstride);
        op->n_launches++;
    }
    RMB_CUDA(cudaGetLastError());
    if (op->time_matvec) {
        RMB_CUDA(cudaEventRecord(ev.second, st));
        op->mv_events.push_back(ev);
    }
    op->n_matvec_launches++;
    return RMB_OK;
}

template <typename T>
static int ensure(T** p, size_t count) {
    if (*p) cudaFree(*p);
    *p = nullptr;
    RMB_CUDA(cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T)));
    return RMB_OK;
}

// speculative enqueue: the next call starts with as many iterations as this one needed
static inline void commit_spec(rmb_operator* op) {
    op->spec_guess = std::max(1, op->spec_seen);
    op->spec_seen = 0;
}

// partial sums per state written by the linear-rotor matvec in use: m-groups (register-window kernel) or 32-row chunks
static inline int lin_parts(const rmb_operator* op) {
    return op->lw_cur ? op->lw_groups : (op->mw_cur ? op->mw_groups : op->lin_npart);
}

// the tiled kernel can fuse the <w, V_k> partial sums only if it covers every bra block
static inline bool fused_dot(const rmb_operator* op) { return op->nitems == 0 && op->nitems2 + op->nitemsG > 0; }
static inline int dot_parts(const rmb_operator* op) { return fused_dot(op) ? op->dot_slots : op->W->nchunk; }

// (re)allocate the per-state small arrays and the product vector for `cap` states
static int ensure_workspace(rmb_operator* op, long long cap, int maxorder) {
    if (cap <= op->W->ws_states && maxorder <= op->W->ws_maxorder && op->W->d_w) return RMB_OK;
    cap = std::max(cap, op->W->ws_states);
    maxorder = std::max(maxorder, op->W->ws_maxorder);
    RMB_CUDA(cudaDeviceSynchronize());
    for (auto* s : op->W->slabs) cudaFree(s);
    op->W->slabs.clear();
    op->W->slab_ptrs_uploaded = 0;
    op->W->nchunk = nchunks(op->np);
    int rc;
    const size_t vec = (size_t)cap * (size_t)op->np;
    const size_t np = (size_t)std::max({op->W->nchunk, op->dot_slots, op->lin_npart});
    if ((rc = ensure(&op->W->d_w, vec))) return rc;
    RMB_CUDA(cudaMemset(op->W->d_w, 0, vec * sizeof(cplx)));   // pad elements stay zero forever
    if ((rc = ensure(&op->W->d_alpha, (size_t)cap * maxorder))) return rc;
    if ((rc = ensure(&op->W->d_beta, (size_t)cap * (maxorder + 1)))) return rc;
    if ((rc = ensure(&op->W->d_rinv, (size_t)cap * (maxorder + 1)))) return rc;
    if ((rc = ensure(&op->W->d_ccur, (size_t)cap * maxorder))) return rc;
    if ((rc = ensure(&op->W->d_ceff, (size_t)cap * maxorder))) return rc;
    if ((rc = ensure(&op->W->d_dc, (size_t)cap * maxorder))) return rc;
    if ((rc = ensure(&op->W->d_active, (size_t)cap))) return rc;
    if ((rc = ensure(&op->W->d_order, (size_t)cap))) return rc;
    if ((rc = ensure(&op->W->d_pdot, (size_t)cap * np))) return rc;
    if ((rc = ensure(&op->W->d_pnrm, (size_t)cap * op->W->nchunk))) return rc;
    if ((rc = ensure(&op->W->d_pconv, (size_t)cap * op->W->nchunk))) return rc;
    if ((rc = ensure(&op->W->d_pg0, (size_t)cap * op->W->nchunk))) return rc;
    if ((rc = ensure(&op->W->d_gdiag, (size_t)cap * (maxorder + 1)))) return rc;
    if ((rc = ensure(&op->W->d_ticket, (size_t)cap))) return rc;
    RMB_CUDA(cudaMemset(op->W->d_ticket, 0, std::max<size_t>(1, (size_t)cap) * sizeof(unsigned)));
    if ((rc = ensure(&op->W->d_ctrl, (size_t)4 * (maxorder + 2)))) return rc;
    // pinned AND mapped: the control words are published by a one-warp kernel (k_publish) instead of a D2H memcpy on the
    // compute stream, which would queue behind any large download in flight on the same copy engine (the host-buffer
    // pipeline's own downloads serialised every chunk's Lanczos loop that way)
    if (op->W->h_ctrl) cudaFreeHost(op->W->h_ctrl);
    RMB_CUDA(cudaHostAlloc((void**)&op->W->h_ctrl, sizeof(int) * 4 * (maxorder + 2), cudaHostAllocMapped));
    RMB_CUDA(cudaHostGetDevicePointer((void**)&op->W->hd_ctrl, op->W->h_ctrl, 0));
    while ((int)op->W->it_events.size() < maxorder + 2) {
        cudaEvent_t e;
        RMB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        op->W->it_events.push_back(e);
    }
    op->W->ws_states = cap;
    op->W->ws_maxorder = maxorder;
    return RMB_OK;
}

static int ensure_slab(rmb_operator* op, int k, cudaStream_t st) {
    while ((int)op->W->slabs.size() <= k) {
        cplx* p = nullptr;
        RMB_CUDA(cudaMalloc((void**)&p, (size_t)op->W->ws_states * (size_t)op->np * sizeof(cplx)));
        RMB_CUDA(cudaMemsetAsync(p, 0, (size_t)op->W->ws_states * (size_t)op->np * sizeof(cplx), st));
        op->W->slabs.push_back(p);
    }
    if (op->W->slab_ptrs_cap < (int)op->W->slabs.size() || !op->W->d_slab_ptrs) {
        RMB_CUDA(cudaStreamSynchronize(st));
        if (op->W->d_slab_ptrs) cudaFree(op->W->d_slab_ptrs);
        op->W->slab_ptrs_cap = std::max(16, (int)op->W->slabs.size() * 2);
        op->W->slab_ptrs_uploaded = 0;
        RMB_CUDA(cudaMalloc((void**)&op->W->d_slab_ptrs, op->W->slab_ptrs_cap * sizeof(cplx*)));
    }
    if (op->W->slab_ptrs_uploaded != (int)op->W->slabs.size()) {
        RMB_CUDA(cudaStreamSynchronize(st));   // kernels in flight read the table; slabs.data() is pageable
        RMB_CUDA(cudaMemcpyAsync(op->W->d_slab_ptrs, op->W->slabs.data(), op->W->slabs.size() * sizeof(cplx*),
                                 cudaMemcpyHostToDevice, st));
        RMB_CUDA(cudaStreamSynchronize(st));
        op->W->slab_ptrs_uploaded = (int)op->W->slabs.size();
    }
    return RMB_OK;
}

// One sub-batch of states through the literal Lanczos loop of tdse.py:417-486, in two halves so that several
// batches (the chunks of the host-buffer pipeline, each on its own stream and workspace) can be in flight at once:
//
//   lanczos_begin   phase/scatter, then iterations 0..guess enqueued WITHOUT asking the device anything (`guess` = the
//                   iteration at which the previous call of this operator ran out of active states: fields change
//                   slowly from step to step, so the loop length rarely changes);
//   lanczos_finish  reads the control words (one synchronisation), continues one iteration at a time if states are
//                   still active (host one iteration ahead of the device), then the final combination.
//
// Launches per iteration: matvec (+ fused scale and <w,V_k> partials) and k_recur_gram (recurrence, alpha/beta, small
// exponential, convergence metric, stop rule); beyond RMB_GRAM_KMAX iterations the explicit evaluation with k_small_a /
// k_recur_conv / k_small_b.  An iteration enqueued after every state has retired costs a few microseconds (every
// kernel returns on the `active` flags).
static int lz_enqueue_iter(rmb_operator* op, int k) {
    Workspace* W = op->W;
    LzRun& r = W->run;
    cudaStream_t st = r.st;
    const long long n = op->n, np = op->np, B = r.B;
    const int nch = W->nchunk;
    const int ts = W->ws_maxorder, bs = W->ws_maxorder + 1;
    const dim3 vgrid((unsigned)nch, (unsigned)B);
    static const bool gram = !(getenv("RMB_GRAM") && atoi(getenv("RMB_GRAM")) == 0);
    int rc;
    if ((rc = ensure_slab(op, k + 1, st))) return rc;
    cplx* Vk = W->slabs[k];
    MvEpilogue ep;
    if (r.fused) {
        ep.scale = W->d_rinv + k;
        ep.scale_stride = bs;
        ep.pdot = W->d_pdot;
        ep.npart = r.npart;
        ep.use_lin = r.lin;
    }
    if ((rc = launch_matvec(op, Vk, W->d_w, B, np, np, W->d_active, st, ep))) return rc;
    if (!r.fused) {
        // some bra blocks went through the scalar kernel, which has no epilogue: scale the product
        // (w = rinv_k * H slab_k) and form the partial dots in separate passes
        k_scale_rows<<<vgrid, VEC_THREADS, 0, st>>>(W->d_w, np, np, W->d_rinv + k, bs, W->d_active);
        k_dot<<<vgrid, VEC_THREADS, 0, st>>>(W->d_w, Vk, np, np, W->d_pdot, r.npart, W->d_active);
        op->n_launches += 2;
    }
    if (gram && k < RMB_GRAM_KMAX) {
        // one launch: recurrence + (last CTA per state) alpha/beta, small exponential, Gram-diagonal
        // convergence metric, stop rule, zero-beta fallback
        k_recur_gram<<<dim3((unsigned)r.nsl_p, (unsigned)B), VEC_THREADS, 0, st>>>(
            W->d_w, W->d_slab_ptrs, np, np, W->d_pdot, r.npart, W->d_alpha, W->d_beta, W->d_rinv, W->d_gdiag, ts, bs, k,
            r.fac, W->d_ccur, W->d_ceff, W->d_pnrm, W->d_pg0, r.nsl_p, r.cps_p, nch, W->d_ticket, r.tol, r.maxorder,
            W->d_active, W->d_order, W->d_ctrl, op->d_pmap, n);
        op->n_launches += 1;
    } else {
        // explicit evaluation of sum |u_k - u_{k-1}|^2 over the Krylov history (many vectors: the Gram matrix is no
        // longer close to diagonal once the three-term recurrence loses orthogonality)
        k_small_a<<<(unsigned)B, 32, 0, st>>>(W->d_pdot, r.npart, W->d_alpha, W->d_beta, W->d_rinv, ts, bs, k, r.fac,
                                              W->d_ccur, W->d_ceff, W->d_dc, W->d_active);
        k_recur_conv<<<vgrid, VEC_THREADS, 0, st>>>(W->d_w, W->d_slab_ptrs, np, np, W->d_alpha, W->d_beta, W->d_rinv,
                                                    W->d_dc, ts, bs, k, W->d_pnrm, W->d_pconv, nch, W->d_active);
        k_small_b<<<(unsigned)B, VEC_THREADS, 0, st>>>(W->d_pnrm, W->d_pconv, nch, W->d_beta, W->d_rinv, bs, k, r.tol,
                                                       r.maxorder, W->d_active, W->d_order, W->d_ctrl, W->d_slab_ptrs, np,
                                                       n, op->d_pmap);
        op->n_launches += 3;
    }
    op->n_iterations++;
    // control words of the iteration -> pinned, mapped host mirror (a kernel, not a memcpy: see ensure_workspace)
    k_publish<<<1, 32, 0, st>>>(W->d_ctrl + 4 * k, W->hd_ctrl + 4 * k, 4);
    RMB_CUDA(cudaEventRecord(W->it_events[k], st));
    r.k_last = k;
    return RMB_OK;
}

static int lanczos_begin(rmb_operator* op, cplx* psi, long long B, long long ld, cplx fac, double tol, int maxorder,
                         const cplx* ph, cudaStream_t st) {
    Workspace* W = op->W;
    LzRun& r = W->run;
    const long long n = op->n, np = op->np;
    const int nch = W->nchunk;
    const int bs = W->ws_maxorder + 1;
    r.psi = psi; r.B = B; r.ld = ld; r.fac = fac; r.tol = tol; r.maxorder = maxorder; r.ph = ph; r.st = st;
    r.lin = op->lin_ok && B >= 4 * op->lin_T;      // large batches of linear rotors: sliding window
    r.fused = r.lin || fused_dot(op);
    if (r.lin && op->lin_flat_dirty) lin_update_bound(op);       // decides which linear-rotor kernel runs (field dependent)
    r.npart = r.lin ? lin_parts(op) : dot_parts(op);
    // sliced grids of the vector kernels: a CTA walks `cps` consecutive chunks of one state (~12 CTAs per SM in total)
    auto slices = [&](int nchunk_, int* nsl_, int* cps_) {
        const long long want = std::max<long long>(1, (12LL * op->num_sms + B - 1) / B);
        const int nsl0 = (int)std::min<long long>(nchunk_, want);
        *cps_ = (nchunk_ + nsl0 - 1) / nsl0;
        *nsl_ = (nchunk_ + *cps_ - 1) / *cps_;
    };
    slices(nch, &r.nsl_p, &r.cps_p);
    slices(nchunks(n), &r.nsl_u, &r.cps_u);
    int rc;
    if ((rc = ensure_slab(op, 2, st))) return rc;
    RMB_CUDA(cudaMemsetAsync(W->d_ctrl, 0, sizeof(int) * 4 * (maxorder + 2), st));
    k_init_states<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(W->d_active, W->d_order, W->d_rinv, W->d_beta, bs, (int)B);
    k_phase_init_s<<<dim3((unsigned)r.nsl_u, (unsigned)B), VEC_THREADS, 0, st>>>(psi, ld, ph, W->slabs[0], np, n, op->d_pmap,
                                                                                  r.cps_u, nchunks(n));
    op->n_launches += 2;
    static const bool spec = !(getenv("RMB_SPEC") && atoi(getenv("RMB_SPEC")) == 0);
    const int guess = spec ? std::min(op->spec_guess, maxorder - 1) : 0;
    for (int k = 0; k <= std::max(0, guess); ++k)
        if ((rc = lz_enqueue_iter(op, k))) return rc;
    r.open = true;
    return RMB_OK;
}

static int lanczos_finish(rmb_operator* op, int* orders_host, bool* hit_maxorder) {
    Workspace* W = op->W;
    LzRun& r = W->run;
    cudaStream_t st = r.st;
    const long long n = op->n, np = op->np, B = r.B;
    const int ts = W->ws_maxorder;
    int rc;
    r.open = false;
    // control words of everything enqueued so far
    int checked = -1, stop = -1;
    std::vector<long long> act_hist;
    auto check_to = [&](int upto) -> int {
        RMB_CUDA(cudaEventSynchronize(W->it_events[upto]));
        while (checked < upto && stop < 0) {
            ++checked;
            act_hist.push_back(W->h_ctrl[4 * checked]);
            if (W->h_ctrl[4 * checked + 1]) *hit_maxorder = true;
            if (W->h_ctrl[4 * checked] == 0) stop = checked;
        }
        return RMB_OK;
    };
    if ((rc = check_to(r.k_last))) return rc;
    // states still active: continue, the host one iteration ahead of the device (its wake-up and launch latencies stay
    // hidden; the price is one enqueued iteration in which every state is inactive)
    while (stop < 0 && checked < r.maxorder) {
        if (r.k_last <= checked && r.k_last < r.maxorder)
            if ((rc = lz_enqueue_iter(op, r.k_last + 1))) return rc;
        if (r.k_last == checked + 1 && r.k_last < r.maxorder)
            if ((rc = lz_enqueue_iter(op, r.k_last + 1))) return rc;
        if ((rc = check_to(checked + 1))) return rc;
    }
    if (stop >= 0) op->spec_seen = std::max(op->spec_seen, stop);
    k_combine_s<<<dim3((unsigned)r.nsl_u, (unsigned)B), VEC_THREADS, 0, st>>>(W->d_slab_ptrs, np, n, W->d_ceff, ts, W->d_order,
                                                                               r.ph, r.psi, r.ld, op->d_pmap, r.cps_u,
                                                                               nchunks(n));
    op->n_launches++;
    RMB_CUDA(cudaGetLastError());
    // state-matvecs actually performed: all states in iteration 0, the survivors of k-1 in iteration k
    op->n_state_matvecs += B;
    for (long long a : act_hist) op->n_state_matvecs += a;
    if (op->pipe_orders) {
        // host-buffer pipeline: orders of all chunks are gathered on the device and downloaded once at the end
        k_publish<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(W->d_order, op->pipe_orders, (int)B);
        op->pipe_orders += B;
    } else if (orders_host) {
        // asynchronous: the caller synchronises the stream before reading (rmb_propagate_step documents
        // this; the Python layer reads `last_orders` lazily).  Pageable destinations make the copy
        // synchronous, pinned ones do not.
        RMB_CUDA(cudaMemcpyAsync(orders_host, W->d_order, sizeof(int) * B, cudaMemcpyDeviceToHost, st));
    }
    return RMB_OK;
}

static int lanczos_batch(rmb_operator* op, cplx* psi, long long B, long long ld, cplx fac, double tol,
                         int maxorder, const cplx* ph, int* orders_host, cudaStream_t st,
                         bool* hit_maxorder) {
    int rc = lanczos_begin(op, psi, B, ld, fac, tol, maxorder, ph, st);
    if (rc) return rc;
    return lanczos_finish(op, orders_host, hit_maxorder);
}

static int propagate_device(rmb_operator* op, cplx* psi, long long nstates, long long ld, cplx fac,
                            double tol, int maxorder, const cplx* ph, int skip, int* orders_host,
                            cudaStream_t st) {
    const long long n = op->n;
    if (nstates == 0) return RMB_OK;
    if (skip) {
        if (ph) {
            const long long max_y = 65535;
            for (long long s0 = 0; s0 < nstates; s0 += max_y) {
                const long long ns = std::min(nstates - s0, max_y);
                k_phase_mul2<<<dim3((unsigned)nchunks(n), (unsigned)ns), VEC_THREADS, 0, st>>>(psi + s0 * ld, ld, ph, n);
                op->n_launches++;
            }
            RMB_CUDA(cudaGetLastError());
        }
        if (op->pipe_orders) {
            RMB_CUDA(cudaMemsetAsync(op->pipe_orders, 0, sizeof(int) * nstates, st));
            op->pipe_orders += nstates;
        } else if (orders_host) {
            std::fill(orders_host, orders_host + nstates, 0);
        }
        return RMB_OK;
    }
    int rc = check_field(op);
    if (rc) return rc;
    if (maxorder < 1 || maxorder > MAX_ORDER_SMEM) {
        set_error("maxorder must be in [1, 128]");
        return RMB_ERR_INVALID;
    }
    // small linear-rotor problems: the whole step in one launch (rmb_fused.cuh)
    if (op->fused_ok && nstates <= 65535 &&
        (long long)nstates * n * (long long)sizeof(cplx) * (maxorder + 2) <= (1LL << 30)) {
        if ((rc = ensure_workspace(op, nstates, maxorder))) return rc;
        if ((rc = ensure_slab(op, maxorder, st))) return rc;
        if (!op->defer_error) RMB_CUDA(cudaMemsetAsync(op->W->d_ctrl, 0, sizeof(int) * 4, st));
        FusedArgs fa;
        fa.n = n;
        fa.row_blk = op->d_row_blk;
        fa.blk_begin = op->d_blk_begin;
        fa.blk_off = op->d_blk_off;
        fa.blk_dm = op->d_blk_dm;
        fa.prods = op->d_prods;
        fa.tab_mask = op->d_tab_mask;
        fa.cent = (const MfEntry*)op->d_ent_cent;
        fa.kpool = op->d_kpool;
        fa.k_complex = op->k_complex ? 1 : 0;
        fa.nprod = op->nprod;
        fa.slabs = op->W->d_slab_ptrs;
        fa.fac = fac;
        fa.tol = tol;
        fa.maxorder = maxorder;
        fa.ph = ph;
        fa.psi = psi;
        fa.ld = ld;
        fa.order = op->W->d_order;
        fa.ctrl = op->W->d_ctrl;
        fa.lin_blk = nullptr;
        fa.lin_flat = nullptr;
        fa.lin_val = nullptr;
        static const bool gram_f = !(getenv("RMB_GRAM") && atoi(getenv("RMB_GRAM")) == 0);
        static const bool fused_lin = !(getenv("RMB_FUSED_LIN") && atoi(getenv("RMB_FUSED_LIN")) == 0);
        fa.gram_kmax = gram_f ? RMB_GRAM_KMAX : 0;
        if (op->lin_ok && fused_lin) {
            // merged entry lists of the sliding-window kernel (rebuilt after a field update)
            if ((rc = matvec_prep(op, st, true))) return rc;
            fa.lin_blk = (const LinBlk*)op->d_lin_blk;
            fa.lin_flat = (const LinEnt*)op->d_lin_flat;
            fa.lin_val = op->d_lin_val;
        }
        // entry lists in shared memory behind the vectors when they fit (they share the room of the product descriptors)
        size_t fused_dyn = (size_t)3 * n * sizeof(cplx) + (size_t)op->nprod * sizeof(FusedProd);
        fa.nblocks = op->nblocks;
        fa.lin_lcap = 0;
        if (fa.lin_blk && op->fused_lcap > 0) {
            const size_t with_tab = (size_t)3 * n * sizeof(cplx) + fused_lin_table_bytes(op->nblocks, op->fused_lcap);
            if (with_tab <= 210 * 1024) {
                fa.lin_lcap = op->fused_lcap;
                fused_dyn = std::max(fused_dyn, with_tab);
                if (fused_dyn > g_fused_smem) {
                    RMB_CUDA(cudaFuncSetAttribute(k_lanczos_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_dyn));
                    g_fused_smem = fused_dyn;
                }
            }
        }
        // the history slabs are indexed [slab][state * n + i] with the leading dimension of this batch
        k_lanczos_fused<<<(unsigned)nstates, FUSED_THREADS, fused_dyn, st>>>(fa, nstates);
        RMB_CUDA(cudaGetLastError());
        op->n_launches++;
        op->n_iterations++;
        if (op->pipe_orders) {
            k_publish<<<(unsigned)((nstates + 255) / 256), 256, 0, st>>>(op->W->d_order, op->pipe_orders, (int)nstates);
            op->pipe_orders += nstates;
        } else if (orders_host) {
            RMB_CUDA(cudaMemcpyAsync(orders_host, op->W->d_order, sizeof(int) * nstates, cudaMemcpyDeviceToHost, st));
        }
        if (op->defer_error) return RMB_OK;      // rmb_propagate_many checks the flag once at the end
        k_publish<<<1, 32, 0, st>>>(op->W->d_ctrl, op->W->hd_ctrl, 1);
        RMB_CUDA(cudaStreamSynchronize(st));
        if (op->W->h_ctrl[0]) {
            char buf[128];
            snprintf(buf, sizeof(buf), "Lanczos reached maximum order of '%d' without convergence", maxorder);
            set_error(buf);
            return RMB_ERR_MAXORDER;
        }
        return RMB_OK;
    }
    // sub-batch size from the workspace budget: the product vector and ~15 Krylov vectors per state
    long long bc = op->W->ws_states;
    if (bc < std::min<long long>(nstates, 65535)) {
        // more states than the workspace holds: (re)size it from the budget (rare: first call / larger batch)
        long long budget = op->ws_budget;
        if (budget <= 0) {
            size_t fr = 0, tot = 0;
            RMB_CUDA(cudaMemGetInfo(&fr, &tot));
            long long held = (long long)(op->W->slabs.size() + 1) * op->W->ws_states * op->np * (long long)sizeof(cplx);
            budget = (long long)(0.4 * (double)(fr + (size_t)held));
        }
        const long long per_state = 16LL * op->np * (long long)sizeof(cplx);
        bc = std::max({1LL, op->W->ws_states, std::min({(long long)nstates, budget / per_state, 65535LL})});
    }
    bc = std::max(1LL, std::min(bc, (long long)nstates));
    if ((rc = ensure_workspace(op, bc, maxorder))) return rc;
    bc = std::min<long long>(op->W->ws_states, nstates);
    bool hit = false;
    for (long long s0 = 0; s0 < nstates; s0 += bc) {
        const long long b = std::min(bc, nstates - s0);
        rc = lanczos_batch(op, psi + s0 * ld, b, ld, fac, tol, maxorder, ph,
                           orders_host ? orders_host + s0 : nullptr, st, &hit);
        if (rc) return rc;
    }
    if (hit) {
        char buf[128];
        snprintf(buf, sizeof(buf), "Lanczos reached maximum order of '%d' without convergence", maxorder);
        set_error(buf);
        return RMB_ERR_MAXORDER;
    }
    return RMB_OK;
}

// scratch for the entry points that take user-layout vectors: one padded slab + the product vector
static int scratch_for(rmb_operator* op, long long nstates, cudaStream_t st) {
    int rc;
    if (op->W->ws_states == 0) {
        const long long cap = std::max<long long>(1, (1LL << 30) / (op->np * (long long)sizeof(cplx)));
        if ((rc = ensure_workspace(op, std::min<long long>({nstates, cap, 65535LL}), 1))) return rc;
    }
    return ensure_slab(op, 0, st);
}

}  // namespace rmb

extern "C" {

int32_t rmb_matvec(rmb_operator* op, const double* x_dev, double* y_dev, int64_t nstates, int64_t ld,
                   void* stream) {
    if (!op || !x_dev || !y_dev || x_dev == y_dev || ld < op->n) {
        set_error("matvec: bad arguments");
        return RMB_ERR_INVALID;
    }
    int rc = check_field(op);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (nstates == 0) return RMB_OK;
    if ((rc = scratch_for(op, nstates, st))) return rc;
    const long long bc = op->W->ws_states, n = op->n, np = op->np;
    for (long long s0 = 0; s0 < nstates; s0 += bc) {
        const long long b = std::min(bc, (long long)nstates - s0);
        const dim3 ugrid((unsigned)nchunks(n), (unsigned)b);
        // user layout -> padded scratch, product, back
        k_phase_init<<<ugrid, VEC_THREADS, 0, st>>>((const cplx*)x_dev + s0 * ld, ld, nullptr, op->W->slabs[0], np, n,
                                                    op->d_pmap);
        MvEpilogue epm;
        epm.use_lin = op->lin_ok && b >= 4 * op->lin_T;
        if ((rc = launch_matvec(op, op->W->slabs[0], op->W->d_w, b, np, np, nullptr, st, epm))) return rc;
        k_unpad<<<ugrid, VEC_THREADS, 0, st>>>(op->W->d_w, np, (cplx*)y_dev + s0 * ld, ld, n, op->d_pmap);
        op->n_launches += 2;
        op->n_state_matvecs += b;
    }
    RMB_CUDA(cudaGetLastError());
    return RMB_OK;
}

int32_t rmb_propagate_step(rmb_operator* op, double* psi_dev, int64_t nstates, int64_t ld, double fac_re,
                           double fac_im, double tol, int32_t maxorder, const double* h0phase_dev,
                           int32_t skip_krylov, int32_t* orders_host, void* stream) {
    if (!op || !psi_dev || ld < op->n || nstates < 0) {
        set_error("propagate_step: bad arguments");
        return RMB_ERR_INVALID;
    }
    const int rc = propagate_device(op, (cplx*)psi_dev, nstates, ld, make_double2(fac_re, fac_im), tol, maxorder,
                                    (const cplx*)h0phase_dev, skip_krylov, orders_host, (cudaStream_t)stream);
    commit_spec(op);
    return rc;
}

int32_t rmb_propagate_many(rmb_operator* op, double* psi_dev, int64_t nstates, int64_t ld, int32_t nsteps,
                           double fac_re, double fac_im, double tol, int32_t maxorder,
                           const double* h0phase_dev, int32_t ndyn, const int32_t* dyn_part,
                           const double* fprod, const double* thresh, const int32_t* all_dropped,
                           int32_t nobs, rmb_operator** obs, int32_t obs_every, double* expval_dev,
                           int32_t* orders_host, void* stream) {
    if (!op || !psi_dev || ld < op->n || nstates < 0 || nsteps < 0 || ndyn < 0 || nobs < 0 ||
        (ndyn > 0 && (!dyn_part || !fprod || !thresh || !all_dropped)) || (nobs > 0 && (!obs || !expval_dev))) {
        set_error("propagate_many: bad arguments");
        return RMB_ERR_INVALID;
    }
    for (int j = 0; j < ndyn; ++j)
        if (dyn_part[j] < 0 || dyn_part[j] >= (int)op->parts.size() || op->parts[dyn_part[j]].ncart > 16) {
            set_error("propagate_many: bad time-dependent part");
            return RMB_ERR_INVALID;
        }
    cudaStream_t st = (cudaStream_t)stream;
    if (obs_every < 1) obs_every = 1;
    const cplx fac = make_double2(fac_re, fac_im);
    int result = RMB_OK, rc;
    // the fused path needs its workspace before the flag can be deferred
    const bool fused = op->fused_ok && nstates <= 65535 && nstates > 0 &&
                       (long long)nstates * op->n * (long long)sizeof(cplx) * (maxorder + 2) <= (1LL << 30);
    if (fused) {
        if ((rc = ensure_workspace(op, nstates, maxorder))) return rc;
        RMB_CUDA(cudaMemsetAsync(op->W->d_ctrl, 0, sizeof(int) * 4, st));
        op->defer_error = true;
    }
    for (int i = 0; i < nsteps; ++i) {
        for (int j = 0; j < ndyn; ++j) {
            rc = rmb_operator_set_field(op, dyn_part[j], fprod + ((size_t)i * ndyn + j) * 16, thresh[j],
                                        all_dropped[(size_t)i * ndyn + j], stream);
            if (rc != RMB_OK) { op->defer_error = false; return rc; }
        }
        int skip = 1;
        for (auto& p : op->parts) skip = skip && p.has_field && p.all_dropped;
        if (op->parts.empty()) skip = 1;
        if (skip && !h0phase_dev) skip = 0;            // without H0 the reference always runs the Krylov part
        rc = propagate_device(op, (cplx*)psi_dev, nstates, ld, fac, tol, maxorder, (const cplx*)h0phase_dev, skip,
                              (i == nsteps - 1) ? orders_host : nullptr, st);
        commit_spec(op);
        if (rc == RMB_ERR_MAXORDER) result = rc;
        else if (rc != RMB_OK) { op->defer_error = false; return rc; }
        if (nobs > 0 && (i % obs_every) == obs_every - 1) {
            for (int o = 0; o < nobs; ++o) {
                rc = rmb_expectation(obs[o], psi_dev, nstates, ld,
                                     expval_dev + 2 * (((size_t)(i / obs_every) * nobs + o) * (size_t)nstates), stream);
                if (rc != RMB_OK) { op->defer_error = false; return rc; }
            }
        }
    }
    if (fused) {
        op->defer_error = false;
        k_publish<<<1, 32, 0, st>>>(op->W->d_ctrl, op->W->hd_ctrl, 1);
        RMB_CUDA(cudaStreamSynchronize(st));
        if (op->W->h_ctrl[0]) result = RMB_ERR_MAXORDER;
    }
    if (result == RMB_ERR_MAXORDER) {
        char buf[128];
        snprintf(buf, sizeof(buf), "Lanczos reached maximum order of '%d' without convergence", maxorder);
        set_error(buf);
    }
    return result;
}

int32_t rmb_propagate_step_host_obs(rmb_operator* op, const double* psi_in_host, double* psi_out_host,
                                    int64_t nstates, int64_t ld, double fac_re, double fac_im, double tol,
                                    int32_t maxorder, const double* h0phase_host, int32_t skip_krylov,
                                    int32_t* orders_host, int32_t nobs, rmb_operator** obs,
                                    double* expval_host, void* stream) {
    if (!op || !psi_in_host || !psi_out_host || ld < op->n || nstates < 0 || nobs < 0 ||
        (nobs > 0 && (!obs || !expval_host))) {
        set_error("propagate_step_host: bad arguments");
        return RMB_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const long long elems = (long long)nstates * ld;
    if (nstates == 0) return RMB_OK;
    int rc;
    if (elems > op->stage_elems) {
        RMB_CUDA(cudaStreamSynchronize(st));
        if ((rc = ensure(&op->d_stage, (size_t)elems))) return rc;
        op->stage_elems = elems;
    }
    if ((long long)nobs * nstates > op->expv_elems) {
        RMB_CUDA(cudaStreamSynchronize(st));
        if ((rc = ensure(&op->d_expv, (size_t)nobs * nstates))) return rc;
        op->expv_elems = (long long)nobs * nstates;
    }
    if (!op->s_in) {
        RMB_CUDA(cudaStreamCreateWithFlags(&op->s_in, cudaStreamNonBlocking));
        RMB_CUDA(cudaStreamCreateWithFlags(&op->s_out, cudaStreamNonBlocking));
    }
    const cplx* ph = nullptr;
    if (h0phase_host) {
        // the phase vector may already live on the device (the Python layer caches it there: uploading 16 N bytes of
        // pageable memory on every call is a synchronous copy of milliseconds at N ~ 10^6)
        cudaPointerAttributes pa;
        const bool on_device = cudaPointerGetAttributes(&pa, h0phase_host) == cudaSuccess &&
                               (pa.type == cudaMemoryTypeDevice || pa.type == cudaMemoryTypeManaged);
        cudaGetLastError();
        if (on_device) {
            ph = (const cplx*)h0phase_host;
        } else {
            if (op->n > op->phase_elems) {
                if ((rc = ensure(&op->d_phase, (size_t)op->n))) return rc;
                op->phase_elems = op->n;
            }
            RMB_CUDA(cudaMemcpyAsync(op->d_phase, h0phase_host, sizeof(cplx) * op->n, cudaMemcpyHostToDevice, st));
            ph = op->d_phase;
        }
    }
    // Chunked pipeline: upload of chunk c+1 and download of chunk c-1 overlap the propagation of chunk c
    // (PCIe is full duplex; uploads on s_in, downloads on s_out, kernels on the caller's stream).
    // chunk count: by bytes (chunks of >= 16 MB keep the copy engines efficient; up to 6 chunks hide all but the first
    // upload and the last download: measured on H2S / H2O / OCS, tools/r02_i.sh), never below 2 states per chunk
    const double mbytes = (double)elems * 16.0 / 1e6;
    int want = (int)std::max(1.0, std::min(6.0, mbytes / 16.0));
    if (const char* e = getenv("RMB_HOST_CHUNKS")) want = std::max(1, atoi(e));
    long long cs = (nstates + want - 1) / want;
    cs = std::max<long long>(2, (cs + 1) & ~1LL);
    const int nchunk = (int)((nstates + cs - 1) / cs);
    while ((int)op->pipe_events.size() < 2 * nchunk + 1) {
        cudaEvent_t e;
        RMB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        op->pipe_events.push_back(e);
    }
    // RMB_E2E_TRACE=1: device timeline of the pipeline on stderr (upload / compute / download end of every chunk)
    static const bool trace = getenv("RMB_E2E_TRACE") && atoi(getenv("RMB_E2E_TRACE")) != 0;
    std::vector<cudaEvent_t> tev;
    auto mark = [&](cudaStream_t s_) {
        if (!trace) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, s_);
        tev.push_back(e);
    };
    mark(st);
    // uploads must not overtake the previous call's downloads of the same staging buffer
    RMB_CUDA(cudaEventRecord(op->pipe_events[2 * nchunk], st));
    RMB_CUDA(cudaStreamWaitEvent(op->s_in, op->pipe_events[2 * nchunk], 0));
    for (int c = 0; c < nchunk; ++c) {
        const long long c0 = c * cs, b = std::min(cs, (long long)nstates - c0);
        RMB_CUDA(cudaMemcpyAsync(op->d_stage + c0 * ld, (const cplx*)psi_in_host + c0 * ld, sizeof(cplx) * b * ld,
                                 cudaMemcpyHostToDevice, op->s_in));
        RMB_CUDA(cudaEventRecord(op->pipe_events[2 * c], op->s_in));
        mark(op->s_in);
    }
    int result = RMB_OK;
    if (nstates > op->pipe_orders_cap) {
        RMB_CUDA(cudaStreamSynchronize(st));
        if ((rc = ensure(&op->d_pipe_orders, (size_t)nstates))) return rc;
        op->pipe_orders_cap = nstates;
    }
    op->pipe_orders = op->d_pipe_orders;
    // Co-running chunks: chunk c runs on compute stream c % NWS with workspace c % NWS (the caller's stream is stream
    // 0), its Lanczos loop enqueued speculatively (lanczos_begin) and finished (lanczos_finish) only when its workspace
    // is needed again, so that up to NWS chunks fill the GPU together: a chunk alone is too small to do so and its
    // per-iteration scalar tails leave gaps.  Falls back to one chunk after the other for the single-launch fused
    // step, for skipped steps and when a chunk does not fit one workspace.
    const cplx fac = make_double2(fac_re, fac_im);
    static const int nws_env = getenv("RMB_CORUN") ? std::max(1, std::min(RMB_NWS, atoi(getenv("RMB_CORUN")))) : RMB_NWS;
    const bool fused_step = op->fused_ok && cs <= 65535 && cs * op->n * (long long)sizeof(cplx) * (maxorder + 2) <= (1LL << 30);
    // (chunks that fill the GPU on their own -- 64 MB of state vector and more -- gain nothing from sharing it and would
    // only delay the first download: measured on the H2S and OCS batches, tools/r02_m.sh)
    const bool small_chunk = (double)cs * (double)op->np * 16.0 < 64e6 || getenv("RMB_CORUN") != nullptr;
    bool corun = nws_env > 1 && nchunk > 1 && small_chunk && !skip_krylov && !fused_step && check_field(op) == RMB_OK &&
                 maxorder >= 1 && maxorder <= MAX_ORDER_SMEM;
    const int NW = corun ? std::min(nws_env, nchunk) : 1;
    if (corun) {
        for (int i = 0; i < NW && corun; ++i) {
            op->W = &op->wsp[i];
            for (int o = 0; o < nobs; ++o) obs[o]->W = &obs[o]->wsp[i];
            if (i > 0 && !op->s_c[i]) RMB_CUDA(cudaStreamCreateWithFlags(&op->s_c[i], cudaStreamNonBlocking));
            if (ensure_workspace(op, cs, maxorder) != RMB_OK || op->W->ws_states < cs) corun = false;
        }
        op->W = &op->wsp[0];
        for (int o = 0; o < nobs; ++o) obs[o]->W = &obs[o]->wsp[0];
    }
    auto finish_chunk = [&](int c) -> int {
        // second half of chunk c: finish its Lanczos loop, observables, download
        const int wi = corun ? c % NW : 0;
        const long long c0 = c * cs, b = std::min(cs, (long long)nstates - c0);
        cudaStream_t sc = wi == 0 ? st : op->s_c[wi];
        int rc2;
        if (corun) {
            op->W = &op->wsp[wi];
            bool hit = false;
            rc2 = lanczos_finish(op, nullptr, &hit);
            op->W = &op->wsp[0];
            if (rc2) return rc2;
            if (hit) result = RMB_ERR_MAXORDER;
        }
        for (int o = 0; o < nobs; ++o) {
            obs[o]->W = &obs[o]->wsp[wi];
            rc2 = rmb_expectation(obs[o], (const double*)(op->d_stage + c0 * ld), b, ld,
                                  (double*)(op->d_expv + (long long)o * nstates + c0), sc);
            obs[o]->W = &obs[o]->wsp[0];
            if (rc2 != RMB_OK) return rc2;
        }
        RMB_CUDA(cudaEventRecord(op->pipe_events[2 * c + 1], sc));
        mark(sc);
        RMB_CUDA(cudaStreamWaitEvent(op->s_out, op->pipe_events[2 * c + 1], 0));
        RMB_CUDA(cudaMemcpyAsync((cplx*)psi_out_host + c0 * ld, op->d_stage + c0 * ld, sizeof(cplx) * b * ld,
                                 cudaMemcpyDeviceToHost, op->s_out));
        mark(op->s_out);
        return RMB_OK;
    };
    if (corun) {
        // field-dependent tables of the matvec kernels and everything else already enqueued on the caller's stream must
        // be visible to the other compute streams
        if ((rc = matvec_prep(op, st, op->lin_ok && cs >= 4 * op->lin_T))) return rc;
        for (int o = 0; o < nobs; ++o)
            if ((rc = matvec_prep(obs[o], st, obs[o]->lin_ok && cs >= 4 * obs[o]->lin_T))) return rc;
        RMB_CUDA(cudaEventRecord(op->pipe_events[2 * nchunk], st));
        for (int i = 1; i < NW; ++i) RMB_CUDA(cudaStreamWaitEvent(op->s_c[i], op->pipe_events[2 * nchunk], 0));
    }
    for (int c = 0; c < nchunk; ++c) {
        const long long c0 = c * cs, b = std::min(cs, (long long)nstates - c0);
        if (!corun) {
            RMB_CUDA(cudaStreamWaitEvent(st, op->pipe_events[2 * c], 0));
            rc = propagate_device(op, op->d_stage + c0 * ld, b, ld, fac, tol, maxorder, ph, skip_krylov, nullptr, st);
            if (rc == RMB_ERR_MAXORDER) result = rc;
            else if (rc != RMB_OK) { op->pipe_orders = nullptr; return rc; }
            if ((rc = finish_chunk(c))) { op->pipe_orders = nullptr; return rc; }
            continue;
        }
        if (c >= NW && (rc = finish_chunk(c - NW))) { op->pipe_orders = nullptr; op->W = &op->wsp[0]; return rc; }
        const int wi = c % NW;
        cudaStream_t sc = wi == 0 ? st : op->s_c[wi];
        RMB_CUDA(cudaStreamWaitEvent(sc, op->pipe_events[2 * c], 0));
        op->W = &op->wsp[wi];
        rc = lanczos_begin(op, op->d_stage + c0 * ld, b, ld, fac, tol, maxorder, ph, sc);
        op->W = &op->wsp[0];
        if (rc) { op->pipe_orders = nullptr; return rc; }
    }
    if (corun) {
        for (int c = std::max(0, nchunk - NW); c < nchunk; ++c)
            if ((rc = finish_chunk(c))) { op->pipe_orders = nullptr; op->W = &op->wsp[0]; return rc; }
        // the caller's stream continues only after the other compute streams are done
        for (int i = 1; i < NW; ++i) {
            RMB_CUDA(cudaEventRecord(op->pipe_events[2 * nchunk], op->s_c[i]));
            RMB_CUDA(cudaStreamWaitEvent(st, op->pipe_events[2 * nchunk], 0));
        }
    }
    commit_spec(op);
    op->pipe_orders = nullptr;
    // small results last, on the download stream (behind the last chunk: nothing on the compute stream ever waits for
    // the D2H copy engine)
    RMB_CUDA(cudaEventRecord(op->pipe_events[2 * nchunk], st));
    RMB_CUDA(cudaStreamWaitEvent(op->s_out, op->pipe_events[2 * nchunk], 0));
    if (nobs > 0)
        RMB_CUDA(cudaMemcpyAsync(expval_host, op->d_expv, sizeof(cplx) * (size_t)nobs * nstates,
                                 cudaMemcpyDeviceToHost, op->s_out));
    if (orders_host)
        RMB_CUDA(cudaMemcpyAsync(orders_host, op->d_pipe_orders, sizeof(int) * (size_t)nstates, cudaMemcpyDeviceToHost,
                                 op->s_out));
    RMB_CUDA(cudaStreamSynchronize(st));
    RMB_CUDA(cudaStreamSynchronize(op->s_out));
    if (trace && !tev.empty()) {
        fprintf(stderr, "[rmb e2e] %d chunks of %lld states:", nchunk, cs);
        for (size_t i = 1; i < tev.size(); ++i) {
            float ms = 0;
            cudaEventElapsedTime(&ms, tev[0], tev[i]);
            const int nc_ = nchunk;
            const char* what = (int)i <= nc_ ? "up" : (((int)i - nc_) % 2 ? "comp" : "down");
            fprintf(stderr, " %s %.2f", what, ms);
        }
        fprintf(stderr, " ms\n");
        for (auto e : tev) cudaEventDestroy(e);
    }
    if (result == RMB_ERR_MAXORDER) {
        char buf[128];
        snprintf(buf, sizeof(buf), "Lanczos reached maximum order of '%d' without convergence", maxorder);
        set_error(buf);
    }
    return result;
}

int32_t rmb_propagate_step_host(rmb_operator* op, const double* psi_in_host, double* psi_out_host,
                                int64_t nstates, int64_t ld, double fac_re, double fac_im, double tol,
                                int32_t maxorder, const double* h0phase_host, int32_t skip_krylov,
                                int32_t* orders_host, void* stream) {
    return rmb_propagate_step_host_obs(op, psi_in_host, psi_out_host, nstates, ld, fac_re, fac_im, tol, maxorder,
                                       h0phase_host, skip_krylov, orders_host, 0, nullptr, nullptr, stream);
}

int32_t rmb_expectation(rmb_operator* op, const double* psi_dev, int64_t nstates, int64_t ld,
                        double* expval_dev, void* stream) {
    if (!op || !psi_dev || !expval_dev || ld < op->n) {
        set_error("expectation: bad arguments");
        return RMB_ERR_INVALID;
    }
    int rc = check_field(op);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const long long n = op->n, np = op->np;
    if (nstates == 0) return RMB_OK;
    if ((rc = scratch_for(op, nstates, st))) return rc;
    const long long bc = op->W->ws_states;
    const int nch = op->W->nchunk;
    for (long long s0 = 0; s0 < nstates; s0 += bc) {
        const long long b = std::min(bc, (long long)nstates - s0);
        const dim3 ugrid((unsigned)nchunks(n), (unsigned)b);
        // user layout -> padded scratch; without padding (every dim_k odd, e.g. linear rotors) the kernels read psi in place
        const cplx* X = (const cplx*)psi_dev + s0 * ld;
        long long ldx = ld;
        if (np != n) {
            k_phase_init<<<ugrid, VEC_THREADS, 0, st>>>(X, ld, nullptr, op->W->slabs[0], np, n, op->d_pmap);
            op->n_launches++;
            X = op->W->slabs[0];
            ldx = np;
        }
        const bool lin = op->lin_ok && b >= 4 * op->lin_T;      // same routing as lanczos_batch
        if (lin || fused_dot(op)) {
            // <psi|O psi> = conj( sum conj(O psi) psi ): partial sums come out of the matvec epilogue and
            // the product vector itself is never written
            MvEpilogue ep;
            ep.pdot = op->W->d_pdot;
            ep.use_lin = lin;
            if (lin && op->lin_flat_dirty) lin_update_bound(op);
            ep.npart = ep.use_lin ? lin_parts(op) : dot_parts(op);
            if ((rc = launch_matvec(op, X, nullptr, b, ldx, np, nullptr, st, ep))) return rc;
            k_reduce_dot<<<(unsigned)b, 32, 0, st>>>(op->W->d_pdot, ep.npart, (cplx*)expval_dev + s0, -1.0);
            op->n_launches += 1;
        } else {
            if ((rc = launch_matvec(op, X, op->W->d_w, b, ldx, np, nullptr, st))) return rc;
            k_dot2<<<dim3((unsigned)nch, (unsigned)b), VEC_THREADS, 0, st>>>(X, ldx, op->W->d_w, np, np, op->W->d_pdot, nch);
            k_reduce_dot<<<(unsigned)b, 32, 0, st>>>(op->W->d_pdot, nch, (cplx*)expval_dev + s0, 1.0);
            op->n_launches += 2;
        }
        op->n_state_matvecs += b;
    }
    RMB_CUDA(cudaGetLastError());
    return RMB_OK;
}

int32_t rmb_populations(const double* psi_dev, int64_t nstates, int64_t n, int64_t ld, double* pop_dev,
                        void* stream) {
    if (!psi_dev || !pop_dev || ld < n) {
        set_error("populations: bad arguments");
        return RMB_ERR_INVALID;
    }
    if (n == 0) return RMB_OK;
    k_populations<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const cplx*)psi_dev, nstates, n, ld, pop_dev);
    RMB_CUDA(cudaGetLastError());
    return RMB_OK;
}

int32_t rmb_set_workspace_budget(rmb_operator* op, int64_t bytes) {
    if (!op) return RMB_ERR_INVALID;
    op->ws_budget = bytes;
    return RMB_OK;
}

int32_t rmb_get_counters(const rmb_operator* op, int64_t* out4) {
    if (!op || !out4) return RMB_ERR_INVALID;
    out4[0] = op->n_launches;
    out4[1] = op->n_matvec_launches;
    out4[2] = op->n_iterations;
    out4[3] = op->n_state_matvecs;
    return RMB_OK;
}

int32_t rmb_operator_work(rmb_operator* op, double* flops_per_state, double* op_bytes, void* stream) {
    if (!op) return RMB_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<unsigned> mask((size_t)op->ntab + 1, 0u);
    RMB_CUDA(cudaMemcpyAsync(mask.data(), op->d_tab_mask, sizeof(unsigned) * mask.size(), cudaMemcpyDeviceToHost, st));
    RMB_CUDA(cudaStreamSynchronize(st));
    double fl = 0, by = 0;
    for (size_t p = 0; p < op->h_prods.size(); ++p) {
        const ProdD& q = op->h_prods[p];
        const double dm1 = op->h_prod_dm1[p], dk1 = op->h_prod_dk1[p];
        const int nnz = __builtin_popcount(mask[q.tab]);
        if (nnz == 0) continue;                       // the reference drops the block (field.py:1137-1139)
        fl += (op->k_complex ? 8.0 : 4.0) * dm1 * dk1 * q.dk2 + 8.0 * nnz * dm1 * q.dk2;
        by += (op->k_complex ? 16.0 : 8.0) * dk1 * q.dk2 + 20.0 * nnz * dm1;
    }
    if (flops_per_state) *flops_per_state = fl;
    if (op_bytes) *op_bytes = by;
    return RMB_OK;
}

int32_t rmb_threej_band(int32_t j1, int32_t j2, int32_t omega, int32_t ncoef, const double* coef_host, double pref,
                        double* out_host, void* stream) {
    if (j1 < 0 || j2 < 0 || omega < 0 || ncoef <= 0 || !coef_host || !out_host) {
        set_error("threej_band: bad arguments");
        return RMB_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nco = (size_t)ncoef * (2 * omega + 1);
    const size_t nout = (size_t)ncoef * (2 * j1 + 1) * (2 * j2 + 1);
    cplx *d_coef = nullptr, *d_out = nullptr;
    RMB_CUDA(cudaMalloc((void**)&d_coef, nco * sizeof(cplx)));
    if (cudaMalloc((void**)&d_out, nout * sizeof(cplx)) != cudaSuccess) {
        cudaFree(d_coef);
        return cuda_fail(cudaGetLastError(), "cudaMalloc");
    }
    cudaMemcpyAsync(d_coef, coef_host, nco * sizeof(cplx), cudaMemcpyHostToDevice, st);
    const int nt = 256;
    const unsigned nb = (unsigned)std::min<size_t>((nout + nt - 1) / nt, 148 * 16);
    k_threej_band<<<nb, nt, 0, st>>>(j1, j2, omega, ncoef, d_coef, pref, d_out);
    cudaMemcpyAsync(out_host, d_out, nout * sizeof(cplx), cudaMemcpyDeviceToHost, st);
    const cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(d_coef);
    cudaFree(d_out);
    if (e != cudaSuccess) return cuda_fail(e, "threej_band");
    RMB_CUDA(cudaGetLastError());
    return RMB_OK;
}

int32_t rmb_small_expm(int32_t nmat, int32_t n, const double* alpha_host, const double* beta_host, double fac_re,
                       double fac_im, double* out_host, void* stream) {
    if (nmat <= 0 || n < 1 || n > MAX_ORDER_SMEM || !alpha_host || !beta_host || !out_host) {
        set_error("small_expm: bad arguments (1 <= n <= 128)");
        return RMB_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t cnt = (size_t)nmat * n;
    cplx *d_a = nullptr, *d_o = nullptr;
    double* d_b = nullptr;
    cudaError_t e = cudaMalloc((void**)&d_a, cnt * sizeof(cplx));
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_o, cnt * sizeof(cplx));
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_b, cnt * sizeof(double));
    if (e == cudaSuccess) {
        cudaMemcpyAsync(d_a, alpha_host, cnt * sizeof(cplx), cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(d_b, beta_host, cnt * sizeof(double), cudaMemcpyHostToDevice, st);
        k_small_expm<<<(unsigned)nmat, 32, 0, st>>>(n, d_a, d_b, make_double2(fac_re, fac_im), d_o);
        cudaMemcpyAsync(out_host, d_o, cnt * sizeof(cplx), cudaMemcpyDeviceToHost, st);
        e = cudaStreamSynchronize(st);
    }
    cudaFree(d_a);
    cudaFree(d_o);
    cudaFree(d_b);
    if (e != cudaSuccess) return cuda_fail(e, "small_expm");
    RMB_CUDA(cudaGetLastError());
    return RMB_OK;
}

int32_t rmb_operator_info(const rmb_operator* op, int64_t* out8) {
    if (!op || !out8) return RMB_ERR_INVALID;
    out8[0] = op->nitems2;
    out8[1] = op->nitemsG;
    out8[2] = op->nitems;
    out8[3] = op->lin_ok ? op->lin_T : 0;
    out8[4] = op->fused_ok ? 1 : 0;
    out8[5] = op->dk_max;
    out8[6] = op->np;
    out8[7] = op->nprod;
    return RMB_OK;
}

int32_t rmb_fp64_peak(double* dfma_tflops, double* dmma_tflops, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 0;
    RMB_CUDA(cudaGetDevice(&dev));
    RMB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int threads = 256, blocks = sms * 8, iters = 20000;
    double* out = nullptr;
    RMB_CUDA(cudaMalloc((void**)&out, sizeof(double) * (size_t)blocks * threads));
    cudaEvent_t e0, e1;
    RMB_CUDA(cudaEventCreate(&e0));
    RMB_CUDA(cudaEventCreate(&e1));
    double best[2] = {0, 0};
    for (int which = 0; which < 2; ++which)
        for (int rep = 0; rep < 4; ++rep) {          // rep 0 warms up
            RMB_CUDA(cudaEventRecord(e0, st));
            if (which == 0) k_peak_dfma<<<blocks, threads, 0, st>>>(out, rep ? iters : 100);
            else k_peak_dmma<<<blocks, threads, 0, st>>>(out, rep ? iters : 100);
            RMB_CUDA(cudaEventRecord(e1, st));
            RMB_CUDA(cudaEventSynchronize(e1));
            float ms = 0;
            RMB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            if (!rep) continue;
            const double fl = which == 0 ? 2.0 * 8 * iters * (double)blocks * threads
                                         : 2.0 * 8 * 8 * 4 * 4 * iters * (double)blocks * (threads / 32);
            best[which] = std::max(best[which], fl / ms * 1e-9);
        }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    RMB_CUDA(cudaGetLastError());
    if (dfma_tflops) *dfma_tflops = best[0];
    if (dmma_tflops) *dmma_tflops = best[1];
    return RMB_OK;
}

int32_t rmb_matvec_timing(rmb_operator* op, int32_t enable, double* ms_out, int64_t* launches_out) {
    if (!op) return RMB_ERR_INVALID;
    double ms = 0;
    long long cnt = 0;
    for (auto& e : op->mv_events) {
        RMB_CUDA(cudaEventSynchronize(e.second));
        float t = 0;
        RMB_CUDA(cudaEventElapsedTime(&t, e.first, e.second));
        ms += t;
        cnt++;
        op->mv_event_pool.push_back(e);
    }
    op->mv_events.clear();
    if (ms_out) *ms_out = ms;
    if (launches_out) *launches_out = cnt;
    op->time_matvec = enable != 0;
    return RMB_OK;
}

}  // extern "C"
