/*
 * richmol_b200 -- C ABI of the B200-native TDSE propagation hot path.
 *
 * This is the drop-in boundary ("Boundary").  The reference (CFEL-CMI/richmol) is pure
 * Python; the seams this library replaces are the Python methods listed next to each entry point
 * (file:line relative to the reference tree) plus the one native FFI the reference already has on
 * this path, the f2py `expokit.zhexpv` Krylov exponential (expokit/expokit.pyf:200-216), whose role
 * `rmb_propagate_step` takes over.  INTEGRATION.md shows the ctypes stub a richmol maintainer
 * would add.
 *
 * Conventions
 *   - plain C types only; no torch / CUDA types in any signature.  `stream` arguments are a
 *     `cudaStream_t` passed as `void*` (NULL = legacy default stream).
 *   - all complex data is interleaved (re, im) IEEE double = numpy complex128.
 *   - state batches are state-major: `psi[s * ld + i]`, i < N, exactly the `(nstates, N)`
 *     C-contiguous layout `TDSE.update` receives (richmol/tdse.py:375,397).
 *   - `*_dev` pointers are device pointers owned by the caller (e.g. `torch.Tensor.data_ptr()`);
 *     the library never frees them.  `*_host` pointers are host memory.
 *   - every function returns RMB_OK (0) or a negative status; `rmb_last_error()` gives the text.
 *   - one host thread per GPU; a handle is bound to the device current at creation.
 */
#ifndef RICHMOL_B200_H
#define RICHMOL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RMB_ABI_VERSION 1

enum {
    RMB_OK = 0,
    RMB_ERR_INVALID = -1,    /* bad argument / inconsistent tables            -> ValueError        */
    RMB_ERR_CUDA = -2,       /* CUDA runtime failure (no device, OOM, launch) -> RuntimeError      */
    RMB_ERR_MAXORDER = -3,   /* Lanczos hit `maxorder` (richmol/tdse.py:480-484) -> ValueError     */
    RMB_ERR_NOFIELD = -4     /* operator part has no field applied (richmol/field.py:1163-1169)    */
};

/* One term of a (lazily summed) operator: a CarTens with K and M factors
 * (data model: richmol/field.py:58-170; sum: richmol/field.py:951-1070).                          */
typedef struct rmb_part_desc {
    int32_t ncart;             /* number of Cartesian components carrying M coefficients           */
    int32_t nprod;             /* number of (Jpair, sympair, irrep) block products                 */
    const int32_t* pr_bra;     /* [nprod] bra (J,sym) block index                                  */
    const int32_t* pr_ket;     /* [nprod] ket (J,sym) block index                                  */
    const int32_t* pr_table;   /* [nprod] M-table index (identical M factors are stored once)      */
    const int64_t* pr_koff;    /* [nprod] offset of the dense row-major dk1 x dk2 K block in kpool,
                                  counted in K elements                                            */
    int32_t k_is_complex;      /* 0: kpool holds doubles, 1: interleaved complex                   */
    const double* kpool;
    int64_t kpool_len;         /* number of K elements in kpool                                    */
    int32_t ntables;
    const int32_t* tb_dm1;     /* [ntables] rows (bra m quanta)                                    */
    const int32_t* tb_dm2;     /* [ntables] columns (ket m quanta)                                 */
    const int32_t* tb_nd;      /* [ntables] ELL width = number of distinct diagonals (col - row) of the
                                  union pattern; slot j of every row is the same diagonal            */
    const int64_t* tb_off;     /* [ntables+1] entry offset; table t has tb_dm1[t]*tb_nd[t] entries,
                                  row-major (m1, j)                                                */
    const int32_t* ent_col;    /* [nent] ket m index of the entry, -1 = padding                    */
    const double* ent_coef;    /* [ncart][nent] complex: M_{cart} value at the entry               */
} rmb_part_desc;

typedef struct rmb_operator_desc {
    int32_t nblocks;           /* (J,sym) blocks in `for J in Jlist2 for sym in symlist2[J]` order
                                  (richmol/tdse.py:343-348)                                         */
    const int64_t* blk_off;    /* [nblocks+1] offset of the block in the flat state vector          */
    const int32_t* blk_dm;     /* [nblocks] dim_m; block index = im*dim_k + ik (m-major)            */
    const int32_t* blk_dk;     /* [nblocks] dim_k                                                   */
    int32_t nparts;
    const rmb_part_desc* parts;
} rmb_operator_desc;

typedef struct rmb_operator rmb_operator;   /* opaque; owns device copies of the tables */

int32_t rmb_abi_version(void);
const char* rmb_last_error(void);
/* number of visible CUDA devices, or a negative status */
int32_t rmb_device_count(void);

/* Upload the block tables of an operator.  Replaces the per-call dict walking of
 * CarTens.vec (richmol/field.py:1212-1243).                                                         */
int32_t rmb_operator_create(const rmb_operator_desc* desc, rmb_operator** out);
void rmb_operator_destroy(rmb_operator* op);
int64_t rmb_operator_dim(const rmb_operator* op);            /* N                                   */
int64_t rmb_operator_nentries(const rmb_operator* op, int32_t part);

/* K1 -- CarTens.field (richmol/field.py:1073-1142):  MF = sum_cart fprod[cart] * M_cart on the
 * device, |MF| < thresh zeroed (thresh <= 0: no element threshold).  `fprod[ncart]` are the
 * products of field components per Cartesian label, already screened by the caller (a dropped
 * product is passed as 0).  `all_dropped` != 0 reproduces the early return of field.py:1107-1112
 * (empty mfmat).                                                                                    */
int32_t rmb_operator_set_field(rmb_operator* op, int32_t part, const double* fprod,
                               double thresh, int32_t all_dropped, void* stream);
/* Copy the contracted MF entries of a part back (for the `mfmat` attribute), [nent] complex.        */
int32_t rmb_operator_get_mf(rmb_operator* op, int32_t part, double* out_host, void* stream);
/* 1 if any MF entry of any part is non-zero (len(H.mfmat) > 0, richmol/tdse.py:377); syncs.         */
int32_t rmb_operator_mf_nonempty(rmb_operator* op, void* stream);

/* K2 -- CarTens.vec (richmol/field.py:1145-1245) for a batch of states:
 * y[s] = sum_products (MF (x) K) x[s].  x_dev != y_dev.                                             */
int32_t rmb_matvec(rmb_operator* op, const double* x_dev, double* y_dev, int64_t nstates,
                   int64_t ld, void* stream);

/* K3+K4 -- TDSE.update (richmol/tdse.py:265-414) with propag='internal'
 * (_expmv_lanczos, richmol/tdse.py:417-486) for every row of psi, in place:
 *     psi <- ph * exp(fac * H) * (ph * psi)          ph = h0phase_dev (may be NULL: no split)
 * with the reference's recurrences and stopping rule (sum |u_k - u_{k-1}|^2 <= tol, at most
 * `maxorder` basis vectors).  orders_host (may be NULL) receives per state the index of the last
 * Lanczos iteration (= matvecs - 1); the copy is enqueued on `stream` -- synchronise the stream before
 * reading it (pinned memory keeps the call asynchronous).  Returns RMB_ERR_MAXORDER if any state
 * reached maxorder (known without a final synchronisation: the flag is raised when a state retires).
 * If `skip_krylov` != 0 only the two phase multiplications are applied (richmol/tdse.py:377).        */
int32_t rmb_propagate_step(rmb_operator* op, double* psi_dev, int64_t nstates, int64_t ld,
                           double fac_re, double fac_im, double tol, int32_t maxorder,
                           const double* h0phase_dev, int32_t skip_krylov,
                           int32_t* orders_host, void* stream);
/* Same, with HOST buffers: copies psi in, propagates, copies the result out (this is the call a
 * numpy-array caller of TDSE.update makes; used for the end-to-end figure).  h0phase_host may be
 * NULL.                                                                                             */
int32_t rmb_propagate_step_host(rmb_operator* op, const double* psi_in_host, double* psi_out_host,
                                int64_t nstates, int64_t ld, double fac_re, double fac_im,
                                double tol, int32_t maxorder, const double* h0phase_host,
                                int32_t skip_krylov, int32_t* orders_host, void* stream);

/* Same as rmb_propagate_step_host, and additionally evaluates <psi_s|O_o|psi_s> of the PROPAGATED states for
 * `nobs` operators on the device before the download (expval_host: [nobs][nstates] complex) -- what the
 * reference's examples do on the host after every update (examples/ocs_alignment.py:96-100), without a
 * second upload.  The ensemble is processed in chunks so that uploads, kernels and downloads overlap.
 * `h0phase_host` may also be a DEVICE pointer (detected with cudaPointerGetAttributes): callers that keep the
 * phase vector resident avoid a pageable 16 N byte upload per call.                                        */
int32_t rmb_propagate_step_host_obs(rmb_operator* op, const double* psi_in_host, double* psi_out_host,
                                    int64_t nstates, int64_t ld, double fac_re, double fac_im,
                                    double tol, int32_t maxorder, const double* h0phase_host,
                                    int32_t skip_krylov, int32_t* orders_host, int32_t nobs,
                                    rmb_operator** obs, double* expval_host, void* stream);

/* Many steps in one call (extension; the reference drives this loop from Python, examples/ocs_alignment.py:
 * 89-100): for step i, the field products of the `ndyn` time-dependent parts are applied
 * (fprod[(i*ndyn + j)*16 + c], thresh[j], all_dropped[i*ndyn + j]; parts not listed keep their field),
 * the ensemble is propagated by one step exactly as rmb_propagate_step does (including the skip rule when
 * every part is screened out), and every `obs_every` steps the per-state expectation values of `nobs`
 * operators are written to expval_dev[((i / obs_every) * nobs + o) * nstates + s] (complex).  No host
 * synchronisation per step on the fused path; errors are reported at the end.                           */
int32_t rmb_propagate_many(rmb_operator* op, double* psi_dev, int64_t nstates, int64_t ld, int32_t nsteps,
                           double fac_re, double fac_im, double tol, int32_t maxorder,
                           const double* h0phase_dev, int32_t ndyn, const int32_t* dyn_part,
                           const double* fprod, const double* thresh, const int32_t* all_dropped,
                           int32_t nobs, rmb_operator** obs, int32_t obs_every, double* expval_dev,
                           int32_t* orders_host, void* stream);

/* K5 -- observables (user code in examples/ocs_alignment.py:99-100, tests/test_tdse.py:66).
 * expval_dev[s] = <psi_s| O |psi_s>  (complex, [nstates]), O given as an operator whose field has
 * been applied (rank-0 tensors: fprod = {1}).                                                       */
int32_t rmb_expectation(rmb_operator* op, const double* psi_dev, int64_t nstates, int64_t ld,
                        double* expval_dev, void* stream);
/* pop_dev[i] = sum_s |psi_s[i]|^2  ([N] doubles).                                                   */
int32_t rmb_populations(const double* psi_dev, int64_t nstates, int64_t n, int64_t ld,
                        double* pop_dev, void* stream);

/* Tuning / diagnostics (not part of the reference seam). */
/* workspace budget in bytes for the Krylov vectors of one sub-batch (default: 40% of free memory) */
int32_t rmb_set_workspace_budget(rmb_operator* op, int64_t bytes);
/* counters since handle creation: [0] kernel launches, [1] matvec launches, [2] Lanczos
 * iterations (batch-level), [3] state-matvecs                                                       */
int32_t rmb_get_counters(const rmb_operator* op, int64_t* out4);
/* algorithmic work of one state-matvec with the field currently applied: flops =
 * sum_products [4(8) dm1 dk1 dk2 + 8 nnz_diag dm1 dk2], counting only the M diagonals that survived the
 * field contraction; op_bytes = K blocks + compacted MF entries, read once per launch.  Syncs.        */
int32_t rmb_operator_work(rmb_operator* op, double* flops_per_state, double* op_bytes, void* stream);
/* device time (ms, CUDA events on `stream`) spent inside matvec launches since the last reset, and
 * the number of launches it covers; enabling timing serialises with event syncs at query time only */
int32_t rmb_matvec_timing(rmb_operator* op, int32_t enable, double* ms_out, int64_t* launches_out);
/* how the matvec of this operator is routed: out8 = { tiled items (k_matvec_tiled), DMMA items
 * (k_matvec_dmma), scalar items (k_matvec_scalar), sliding-window kernel usable (k_matvec_lin),
 * single-launch step usable (k_lanczos_fused), max dim_k, padded dimension, products }            */
int32_t rmb_operator_info(const rmb_operator* op, int64_t* out8);
/* SURVEY 8f-3 -- generator of the laboratory-frame tensor factors (richmol/rot/labtens.py:482-523, which loops over
 * py3nj calls in Python): out[c][a][b] = pref * (-1)^|qa| * sum_sigma coef[c][sigma+omega] * 3j(j2 omega j1; qb sigma -qa),
 * qa = a - j1, qb = b - j2; complex [ncoef][2 j1 + 1][2 j2 + 1], host buffers.  coef = Ux[cart,(omega,sigma)],
 * pref = sqrt((2 j1 + 1)(2 j2 + 1)) gives the M tensor of every Cartesian component; coef = (Us T)_{omega,sigma},
 * pref = 1 the primitive K tensor over |J,k>.                                                               */
int32_t rmb_threej_band(int32_t j1, int32_t j2, int32_t omega, int32_t ncoef, const double* coef_host, double pref,
                        double* out_host, void* stream);
/* The small exponential of the Lanczos loop on its own (richmol/tdse.py:474, `expm(fac * T_k)[:, 0]` via scipy there): nmat
 * Hermitian tridiagonal matrices of order n (n <= 128), diagonal alpha[mat][n] (complex), off-diagonal beta[mat][n] (real,
 * beta[mat][i] couples i-1 and i, beta[mat][0] unused), out[mat][n] complex; host buffers, one warp per matrix -- the same
 * device function the propagation kernels call (registers for n <= 32, shared memory above).  For tests.        */
int32_t rmb_small_expm(int32_t nmat, int32_t n, const double* alpha_host, const double* beta_host, double fac_re,
                       double fac_im, double* out_host, void* stream);
/* FP64 roofline denominators measured on the current device (MEASURED_PEAKS.json has no FP64 entry):
 * register-resident DFMA and DMMA (mma.sync.m8n8k4.f64) loops over all SMs, best of 3, TFLOP/s.    */
int32_t rmb_fp64_peak(double* dfma_tflops, double* dmma_tflops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RICHMOL_B200_H */
