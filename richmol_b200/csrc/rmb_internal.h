// Internal declarations shared by the translation units of librichmol_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <map>
#include <vector>

#include "../../include/richmol_b200.h"

namespace rmb {

typedef double2 cplx;

// one (Jpair, sympair, irrep) block product, device layout
struct ProdD {
    long long ket_off;   // offset of the ket (J,sym) block in the flat state vector
    long long koff;      // offset of the dk1 x dk2 K block in the K pool (elements)
    long long ent_off;   // offset of the MF table (row 0) in the global entry pool
    int dk2;             // ket dim_k
    int dm2;             // ket dim_m
    int nd;              // ELL width (number of diagonals) of the MF table
    int tab;             // global MF table index (slot masks)
};

// one work item of the matvec: a tile of rows (m1) x columns (k1) of one bra block
struct ItemD {
    long long bra_off;   // offset of the bra block in the flat state vector
    int dk1;             // bra dim_k
    int r0, nrows;       // m1 tile
    int c0, ncols;       // k1 tile
    int p_begin, p_end;  // products with this bra block (sorted by bra)
    int dk2max;          // max ket dim_k over those products
};

// tiled matvec (rmb_matvec.cuh)
struct Item2D;
struct XRange;
struct ProdS;
struct Unit2D;

struct PartH {
    int ncart = 0;
    long long ent_begin = 0, ent_end = 0;   // range in the global entry pool
    int tab_begin = 0, tab_end = 0;         // range of global MF table indices
    cplx* d_coef = nullptr;                 // [ncart][nent]
    double* d_fprod = nullptr;              // [ncart]
    bool has_field = false;
    bool all_dropped = false;
    unsigned nzmask = 0xffffffffu;          // Cartesian components with a non-zero field product
};

void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);

#define RMB_CUDA(call)                                        \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return rmb::cuda_fail(e__, #call); \
    } while (0)

}  // namespace rmb

namespace rmb {
constexpr int RMB_NWS = 3;
// parameters of a Lanczos loop that has been started (lanczos_begin) and not yet finished (lanczos_finish)
struct LzRun {
    cplx* psi = nullptr;
    long long B = 0, ld = 0;
    cplx fac;
    double tol = 0;
    int maxorder = 0;
    const cplx* ph = nullptr;
    cudaStream_t st = nullptr;
    int nsl_p = 1, cps_p = 1, nsl_u = 1, cps_u = 1;   // sliced grids of the vector kernels
    bool lin = false, fused = false;
    int npart = 0;
    int k_last = -1;                 // last iteration enqueued
    bool open = false;
};
// ---- Krylov workspace (lazily sized): everything a Lanczos loop over one batch of states writes ----
struct Workspace {
    long long ws_states = 0;         // capacity in states of each slab
    std::vector<rmb::cplx*> slabs;   // V_0, V_1, ... each ws_states * n
    rmb::cplx* d_w = nullptr;        // H V_k
    rmb::cplx** d_slab_ptrs = nullptr;   // device copy of slab pointers
    int slab_ptrs_cap = 0;
    int slab_ptrs_uploaded = 0;
    // per-state small arrays (capacity ws_states, order capacity ws_maxorder)
    int ws_maxorder = 0;
    rmb::cplx* d_alpha = nullptr;    // [S][maxorder]
    double* d_beta = nullptr;        // [S][maxorder+1]
    rmb::cplx* d_ccur = nullptr;     // [S][maxorder]
    rmb::cplx* d_dc = nullptr;       // [S][maxorder] (c^k - c^{k-1}) * rinv
    rmb::cplx* d_ceff = nullptr;     // [S][maxorder] c^k * rinv
    double* d_rinv = nullptr;        // [S][maxorder+1] 1/beta_k (1 for k = 0 and after a fallback)
    std::vector<cudaEvent_t> it_events;
    int* d_active = nullptr;         // [S]
    int* d_order = nullptr;          // [S]
    rmb::cplx* d_pdot = nullptr;     // [S][nchunk]
    double* d_pnrm = nullptr;        // [S][nchunk]
    double* d_pconv = nullptr;       // [S][nchunk]
    double* d_pg0 = nullptr;         // [S][nchunk]  partial <V_0,V_0> (k_recur_gram, k = 0)
    double* d_gdiag = nullptr;       // [S][maxorder+1]  <V_i,V_i> (diagonal of the Gram matrix of the Krylov vectors)
    unsigned* d_ticket = nullptr;    // [S]  arrival counter of k_recur_gram's CTAs per state
    int* d_ctrl = nullptr;           // per iteration k: [4k] states still active, [4k+1] maxorder flag
    int* h_ctrl = nullptr;           // pinned + mapped mirror, written by k_publish
    int* hd_ctrl = nullptr;          // its device address
    int nchunk = 0;
    LzRun run;                       // the Lanczos loop in flight on this workspace
};
}  // namespace rmb

struct rmb_operator {
    int device = 0;
    int num_sms = 148;
    long long n = 0;                 // Hilbert-space dimension
    long long np = 0;                // padded length of the internal vectors (rows of dim_k | 1 elements)
    int* d_pmap = nullptr;           // [n] position of element i in the padded layout
    int nblocks = 0;
    int nprod = 0;
    int nitems = 0;
    bool k_complex = false;
    int dk_max = 1;                  // max dim_k over blocks
    int nd_max = 1;
    long long nent = 0;
    // device tables
    rmb::ProdD* d_prods = nullptr;
    rmb::ItemD* d_items = nullptr;
    int* d_ent_col = nullptr;
    rmb::cplx* d_ent_val = nullptr;
    double* d_ent_cent = nullptr;    // compacted MF entries {re, im, col, pad} (32 bytes), diagonal-major
    int* d_ent_tab = nullptr;        // entry -> global table index
    int* d_tab_off = nullptr;        // [ntab] first entry of each table (int: nent < 2^31)
    int* d_tab_nd = nullptr;         // [ntab] ELL width
    unsigned* d_tab_mask = nullptr;  // [ntab] bit j set <=> diagonal slot j has a non-zero MF entry
    unsigned* d_tab_cplx = nullptr;  // [ntab] non-zero <=> some MF entry of the table has a non-zero imaginary part (same allocation)
    int ntab = 0;
    double* d_kpool = nullptr;       // doubles, or interleaved complex if k_complex
    int* d_flags = nullptr;          // [0] mf non-empty flag (per set_field accumulates), [1..] scratch
    std::vector<rmb::PartH> parts;
    std::vector<rmb::ItemD> h_items;
    std::vector<rmb::ProdD> h_prods;
    std::vector<int> h_prod_dm1, h_prod_dk1;   // bra dims per product (work accounting)
    size_t matvec_smem = 0;          // dynamic shared memory of the scalar matvec launch
    int matvec_S = 1;                // states per CTA (scalar kernel)
    // tiled matvec
    int nitems2 = 0;
    void* d_items2 = nullptr;        // Item2D[]
    void* d_gdesc = nullptr;         // ProdS[]: static per-(item, product) descriptors
    double* d_ktpool = nullptr;      // K^T images in shared-memory layout
    void* d_units = nullptr;         // Unit2D[] for `units_nstates` states (owned by units_cache)
    std::map<long long, std::pair<void*, int>> units_cache, unitsG_cache;   // batch size -> device unit list
    long long units_nstates = -1;
    int nunits = 0;
    std::vector<int> h_item2_states; // states per CTA of each tiled item
    // DMMA matvec for wide K blocks (rmb_matvec_dmma.cuh)
    int nitemsG = 0;
    void* d_itemsG = nullptr;        // ItemD2[]
    void* d_unitsG = nullptr;
    int nunitsG = 0;
    long long unitsG_nstates = -1;
    std::vector<int> h_itemG_states;
    size_t matvecG_smem = 0;
    int kt_doubles = 0;
    int xbuf_elems = 0;
    int np_max = 0;
    int mf_elems = 0;                // MV2_NDMAX * max tile rows
    size_t matvec2_smem = 0;
    // algorithmic work per state-matvec (for DESIGN.md / bench roofline)
    double flops_per_state = 0;
    double op_bytes = 0;

    // ---- fused single-launch Lanczos step (linear rotors, small N): row -> block tables
    bool fused_ok = false;
    int dot_slots = 0;               // <w,v> partials per state written by the tiled / DMMA matvec epilogues
    int fused_lcap = 0;              // entries per bra block the fused kernel stages in shared memory (0: lists stay global)
    bool defer_error = false;        // multi-step mode: no per-step synchronisation for the maxorder flag
    int* d_row_blk = nullptr;
    int* d_blk_begin = nullptr;
    long long* d_blk_off = nullptr;
    int* d_blk_dm = nullptr;

    // ---- sliding-window matvec for linear rotors (rmb_matvec_lin.cuh)
    bool lin_ok = false;
    bool lin_g1 = false;         // RMB_LIN_G1=1: single-state-group instantiation (soak / sanitizer runs only)
    int lin_W = 0, lin_T = 0, lin_dm_max = 0, lin_NS = 0, lin_npart = 0;
    size_t lin_smem = 0;
    bool lin_flat_dirty = true;      // entry lists must be rebuilt (field changed)
    bool nnz_dirty = true;           // ProdS::nnz of the tiled-kernel descriptors must be refreshed (field changed)
    int ngdesc = 0;
    int* d_prod_ket = nullptr;
    void* d_lin_blk = nullptr;       // LinBlk[nblocks]
    void* d_lin_flat = nullptr;      // LinEnt[nblocks][ML_FLAT]
    long long* d_lin_val_off = nullptr;
    rmb::cplx* d_lin_val = nullptr;  // K * MF per (block, entry, row), rebuilt after every field update
    int lin_ebuf = 0;                // worst-case elements of one entry buffer (every diagonal alive)
    int lin_ebuf_cur = 0, lin_NB = 2;    // bound for the fields currently applied, entry buffers that fit
    size_t lin_smem_fixed = 0;       // ring + barriers + block table
    std::vector<unsigned> h_diag_cart;   // per (table, diagonal): Cartesian components with a non-zero coefficient
    std::vector<int> h_diag_off;     // [ntables + 1] offsets into h_diag_cart
    std::vector<int> h_tab_part;     // [ntables] part owning the table
    std::vector<int> h_bra_begin;    // [nblocks + 1] products sorted by bra block
    std::vector<int> h_blk_dm;       // [nblocks]
    // register-window kernel for linear rotors (rmb_matvec_mw.cuh)
    bool sym_ok = false;             // blocks ordered by J with symmetric contiguous m ranges, block distance <= 2
    bool lw_static = false, lw_cur = false;   // ring + register window kernel (k_matvec_linw) usable / usable with the current fields
    int lw_groups = 0, lw_NB = 2, lw_NS = 6;
    size_t lw_smem = 0;
    bool mw_static = false;          // standalone register-window kernel enabled (RMB_MW=1)
    bool mw_cur = false;             // ... and every diagonal that survives the current fields has |dm| <= 1
    int mw_groups = 0;               // groups of 4 m values
    std::vector<int> h_cshift, h_mw_bfirst, h_prod_dm_off;
    std::vector<signed char> h_prod_dm;   // per (product, diagonal slot): m offset of the diagonal
    int* d_cshift = nullptr;
    unsigned char* d_cmap = nullptr; // [nblocks][16] (block distance, dm) -> merged entry (k_lin_entries)
    int* d_mw_counter = nullptr;
    std::map<long long, std::pair<void*, int>> mw_items_cache;   // batch size -> device item list

    // ---- Krylov workspace (lazily sized) ----
    long long ws_budget = 0;         // bytes; 0 = auto
    // one workspace per compute stream of the host-buffer pipeline (chunks co-run on alternating streams); `W` is
    // the one the next enqueued work uses -- every other entry point runs on wsp[0]
    rmb::Workspace wsp[rmb::RMB_NWS];
    rmb::Workspace* W = &wsp[0];
    int spec_guess = 2;              // iteration at which the previous call ran out of active states (speculative enqueue)
    int spec_seen = 0;               // maximum over the batches of the call in flight
    cudaStream_t s_c[rmb::RMB_NWS] = {nullptr, nullptr, nullptr};   // compute streams 1.. of the pipeline ([0] = caller's)
    int* d_pipe_orders = nullptr;    // host-buffer pipeline: Lanczos orders of all chunks (one download at the end)
    int* pipe_orders = nullptr;      // write cursor into d_pipe_orders while a pipeline call is in flight, else nullptr
    long long pipe_orders_cap = 0;
    // host staging for the *_host entry point
    rmb::cplx* d_stage = nullptr;
    long long stage_elems = 0;
    rmb::cplx* d_expv = nullptr;     // [nobs][nstates] expectation values of the host entry point
    long long expv_elems = 0;
    cudaStream_t s_in = nullptr, s_out = nullptr;   // upload / download streams of the chunked pipeline
    std::vector<cudaEvent_t> pipe_events;
    rmb::cplx* d_phase = nullptr;
    long long phase_elems = 0;

    // counters
    long long n_launches = 0, n_matvec_launches = 0, n_iterations = 0, n_state_matvecs = 0;
    bool time_matvec = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> mv_events;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> mv_event_pool;
};
