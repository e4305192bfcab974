// vag_b200.cu -- sm_100a kernels and the C ABI (include/vag.h) of the B200-native
// VegasAfterglow model-evaluation path.  Kernel bodies live in vag_pipeline.cuh / vag_observer.cuh.
//
// Launch geometry (B200: 148 SMs, FP64 path, no tensor cores -- nothing here is a contraction):
//   k_grid<G>        : G = 8 / 16 / 32 lanes per model (32 / G models share a warp's scalar instruction stream)
//   k_dynamics<PAIR> : one thread per unique ODE row, 8 rows per warp for small batches; forward-only rows on a
//                      register-resident dopri5, pair rows with the stage vectors in shared memory
//   k_radiation      : one 64-thread CTA per unique row: finishes the shock tables from the raw node states,
//                      EATS node geometry, photon coefficients (SoA planes, stores coalesced along k)
//   k_eats<MODE,KIND>: one 128-thread CTA per (model, row-split, shock), 8 CTAs / SM; staged rows, the
//                      log2_softplus table and the accumulators in dynamic shared memory
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/vag.h"
#include "vag_libm.cuh"
#include "vag_pipeline.cuh"

using namespace vag;

static_assert(sizeof(vag_params) == 320, "vag_params layout must match vegasafterglow_b200/abi.py");

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
// K0: G lanes per model (vag_grid.cuh GroupPar), 32 / G models per warp, 4 warps per CTA.  Narrow groups
// share the scalar instruction stream between more models (throughput, large batches); wide groups
// finish one model sooner (latency, small batches).
#ifndef GRID_MIN_BLOCKS
#define GRID_MIN_BLOCKS 4
#endif
template <int G>
__global__ void __launch_bounds__(128, GRID_MIN_BLOCKS) k_grid(BatchWs w, const double* __restrict__ t_obs, int n_t_obs) {
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int mi = gt / G;
    if (mi >= w.n_models) return;  // whole groups exit together
    const int lane = threadIdx.x & 31;
    const int shift = lane & ~(G - 1);
    const GroupPar<G> par{lane & (G - 1), (G == 32 ? 0xffffffffu : ((1u << (G & 31)) - 1u)) << shift, shift};
    k0_grid_body(par, w, mi, t_obs[0], t_obs[n_t_obs - 1]);
}

// K0b: exclusive scan of the ragged row / cell counts over the models of the batch + totals.
// One CTA of 1024 threads; thread i owns a contiguous slice of models (k0b_scan_body is the
// sequential statement of the same result).
__global__ void __launch_bounds__(1024) k_scan(BatchWs w) {
    __shared__ int s_rows[1024];
    __shared__ long long s_cells[1024];
    __shared__ int s_max[3][32];
    __shared__ int s_or[32];
    const int tid = threadIdx.x, n = w.n_models;
    const int per = (n + 1023) / 1024;
    const int m0 = min(tid * per, n), m1 = min(m0 + per, n);
    int rows = 0, max_nt = 0, max_nth = 0, max_er = 0, st = 0;
    long long cells = 0;
    for (int mi = m0; mi < m1; ++mi) {
        const GridHeader& h = w.hdr[mi];
        const ModelCfg& cfg = w.cfg[mi];
        if (cfg.fwd.ssc || (cfg.has_rvs && cfg.rvs.ssc)) st |= (1 << 30);  // carries "any ssc" through the OR
        st |= cfg.has_rvs ? (1 << 29) : (1 << 28);                         // "any pair rows" / "any forward-only rows"
        if (cfg.spreading) st |= (1 << 27);                                // "any spreading model"
        rows += h.n_reps;
        cells += (long long)h.n_reps * h.n_t;
        max_nt = max(max_nt, h.n_t);
        max_nth = max(max_nth, h.n_theta);
        max_er = max(max_er, h.n_theta * h.n_phi_eff);
        st |= h.status;
    }
    s_rows[tid] = rows;
    s_cells[tid] = cells;
    __syncthreads();
    // Hillis-Steele inclusive scan over the 1024 slice totals
    for (int off = 1; off < 1024; off <<= 1) {
        int r = 0;
        long long c = 0;
        if (tid >= off) {
            r = s_rows[tid - off];
            c = s_cells[tid - off];
        }
        __syncthreads();
        s_rows[tid] += r;
        s_cells[tid] += c;
        __syncthreads();
    }
    int row0 = s_rows[tid] - rows;
    long long cell0 = s_cells[tid] - cells;
    for (int mi = m0; mi < m1; ++mi) {
        const GridHeader& h = w.hdr[mi];
        w.row_off[mi] = row0;
        w.cell_off[mi] = cell0;
        row0 += h.n_reps;
        cell0 += (long long)h.n_reps * h.n_t;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        max_nt = max(max_nt, __shfl_xor_sync(0xffffffffu, max_nt, o));
        max_nth = max(max_nth, __shfl_xor_sync(0xffffffffu, max_nth, o));
        max_er = max(max_er, __shfl_xor_sync(0xffffffffu, max_er, o));
        st |= __shfl_xor_sync(0xffffffffu, st, o);
    }
    if ((tid & 31) == 0) {
        s_max[0][tid >> 5] = max_nt;
        s_max[1][tid >> 5] = max_nth;
        s_max[2][tid >> 5] = max_er;
        s_or[tid >> 5] = st;
    }
    __syncthreads();
    if (tid == 0) {
        for (int i = 1; i < 32; ++i) {
            max_nt = max(max_nt, s_max[0][i]);
            max_nth = max(max_nth, s_max[1][i]);
            max_er = max(max_er, s_max[2][i]);
            st |= s_or[i];
        }
        w.row_off[n] = s_rows[1023];
        w.cell_off[n] = s_cells[1023];
        w.totals[TOT_ROWS] = s_rows[1023];
        w.totals[TOT_MAX_NT] = max_nt;
        w.totals[TOT_MAX_NTHETA] = max_nth;
        w.totals[TOT_MAX_EROWS] = max_er;
        w.totals[TOT_STATUS_OR] = st & ~(15 << 27);
        w.totals[TOT_ANY_SPREAD] = (st >> 27) & 1;
        w.totals[TOT_ANY_SSC] = (st >> 30) & 1;
        w.totals[TOT_ANY_PAIR] = (st >> 29) & 1;
        w.totals[TOT_ANY_FWD_ONLY] = (st >> 28) & 1;
    }
}

__global__ void k_rowmap(BatchWs w) {
    const int mi = blockIdx.x * blockDim.x + threadIdx.x;
    if (mi >= w.n_models) return;
    k0c_rowmap_body(w, mi);
}

// K1: one thread per unique row, `lanes` rows per 32-thread CTA.  The ODE is a dependent-latency chain
// and rows of one warp diverge (different accept/reject histories, branches of the RHS), so a small
// batch is spread thinly -- few rows per warp, about two warps per scheduler -- and only large batches
// fill whole warps.  The dopri5 stage vectors sit in shared memory, [slot][component][lane]
// (conflict-free: a lane only ever touches its own column).
// PAIR selects the rows a launch integrates (forward+reverse pairs with the shared-memory stepper, or
// forward-only rows with the register-resident one, which needs no shared memory).
// K1a: one warp per unique ODE row builds its time lattice (n_t independent pow(10, .) evaluations)
__global__ void __launch_bounds__(128) k_lattice(BatchWs w, int n_rows) {
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (row < n_rows) k1_lattice_body(w, row, threadIdx.x & 31, 32);
}

template <bool PAIR>
__global__ void __launch_bounds__(32) k_dynamics(BatchWs w, int n_rows, int lanes) {
    __shared__ double s_col[PAIR ? K1_COL_DOUBLES * 32 : 1];
    if ((int)threadIdx.x >= lanes) return;
    const int row = blockIdx.x * lanes + threadIdx.x;
    if (row >= n_rows) return;
    if ((w.cfg[w.row_model[row]].has_rvs != 0) != PAIR) return;
    k1_dynamics_body<PAIR>(w, row, s_col + threadIdx.x, 32);
}

// K1b + K2: one CTA per unique row.  Finishes the row's shock tables from the raw node states the ODE
// kernel left (thread <-> node), applies the early-time reverse-shock extrapolation, derives the EATS
// node geometry and then builds the photon coefficients of both shocks (thread <-> node, SoA plane
// stores coalesced along k).
__global__ void __launch_bounds__(64) k_radiation(BatchWs w) {
    __shared__ int s_cut;
    const int row = blockIdx.x;
    const RowCtx c = row_ctx(w, row);
    const int tid = threadIdx.x, nthr = blockDim.x;
    if (tid == 0) s_cut = c.n_t;
    for (int k = tid; k < c.n_t; k += nthr) k1b_finish_cell(w, row, c, k);
    __syncthreads();
    if (c.has_rvs && w.row_dyn[row].n_saved >= 0) {
        const int mine = extrap_scan(shock_row(w.rvs, c.off), c.n_t, tid, nthr);
        if (mine < c.n_t) atomicMin(&s_cut, mine);
        __syncthreads();
        const int idx_cut = s_cut;
        for (int k = tid; k < c.n_t; k += nthr) k1c_extrap_cell(w, row, c, idx_cut, k);
        __syncthreads();
    }
    const ModelCfg& cfg = w.cfg[c.mi];
    for (int k = tid; k < c.n_t; k += nthr) {
        k1d_geo_cell(w, c, k);
        if (cfg.spreading && w.sh_theta) k1e_spread_geo_cell(w, row, c, k);
        if (!cfg.fwd.ssc) k2_radiation_cell(w, row, k, 0);  // ssc shocks: k_ic_cooling
        if (cfg.has_rvs && !cfg.rvs.ssc) k2_radiation_cell(w, row, k, 1);
    }
}

// K2 for shocks with ssc=True: one thread per (row, shock), sequential in k (IC cooling of cell k
// starts from gamma_c of cell k-1)
__global__ void __launch_bounds__(32) k_ic_cooling(BatchWs w, int n_rows) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    const int which = blockIdx.y;
    if (row >= n_rows) return;
    const int mi = w.row_model[row];
    if (which && !w.cfg[mi].has_rvs) return;
    if (!(which ? w.cfg[mi].rvs : w.cfg[mi].fwd).ssc) return;
    k2_ic_cool_row(w, row, which);
}

__global__ void k_kn_lut(KnLut* lut) {
    const int i = threadIdx.x;
    if (i < KN_LUT_N) kn_lut_entry(i, lut->ratio[i], lut->lg2_ratio[i]);
}

// log2 of the smallest / largest observation frequency (code units, without the 1+z shift)
__global__ void k_nu_range(const double* __restrict__ lg2_nu, int n, double* out) {
    __shared__ double s_lo[256], s_hi[256];
    double lo = kInf, hi = -kInf;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        lo = fmin(lo, lg2_nu[i]);
        hi = fmax(hi, lg2_nu[i]);
    }
    s_lo[threadIdx.x] = lo;
    s_hi[threadIdx.x] = hi;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            s_lo[threadIdx.x] = fmin(s_lo[threadIdx.x], s_lo[threadIdx.x + o]);
            s_hi[threadIdx.x] = fmax(s_hi[threadIdx.x], s_hi[threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[0] = s_lo[0];
        out[1] = s_hi[0];
    }
}

// EATS row constants of every (model, (phi,theta) row): cos to the line of sight, time coefficient,
// log2 solid angle, representative shock row
__global__ void k_rowgeom(BatchWs w) {
    const int mi = blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (w.status[mi] & VAG_ST_CAPACITY) return;
    if (q >= w.hdr[mi].n_theta * w.hdr[mi].n_phi_eff) return;
    k_rowgeom_body(w, mi, q);
}

__global__ void k_rowcos(BatchWs w) {
    const int mi = blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const ModelCfg& cfg = w.cfg[mi];
    if (!(cfg.fwd.ssc || (cfg.has_rvs && cfg.rvs.ssc))) return;
    if (q >= w.hdr[mi].n_theta * w.hdr[mi].n_phi_eff) return;
    k_rowcos_body(w, mi, q);
}
__global__ void k_dop_extrema(BatchWs w) {
    const int mi = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const ModelCfg& cfg = w.cfg[mi];
    if (!(cfg.fwd.ssc || (cfg.has_rvs && cfg.rvs.ssc))) return;
    if (k >= w.hdr[mi].n_t) return;
    k_dop_extrema_body(w, mi, k);
}

// K2b: SSC spectrum tables, one warp per unique cell (both shocks); warps stride over the batch's cells
__global__ void __launch_bounds__(128) k_ic_spectrum(BatchWs w, int n_rows, int n_warps) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const WarpPar par{(int)(threadIdx.x & 31)};
    double* scratch = w.ic_scratch + (size_t)warp * IC_SCRATCH_DOUBLES;
    for (long long g = warp; g < w.n_cells; g += n_warps) {
        // row of global cell g: last row whose first cell is <= g
        int lo = 0, hi = n_rows;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (w.row_cell_off[mid] <= g)
                lo = mid;
            else
                hi = mid;
        }
        const int row = lo;
        const int k = (int)(g - w.row_cell_off[row]);
        const int mi = w.row_model[row];
        const ModelCfg& cfg = w.cfg[mi];
        for (int which = 0; which < 2; ++which) {
            if (which && !cfg.has_rvs) continue;
            if (!(which ? cfg.rvs : cfg.fwd).ssc) continue;
            const int st = k2b_ic_spectrum_cell(par, w, row, k, which, scratch);
            if (st && par.lane == 0) atomicOr(&w.status[mi], st);
        }
    }
}

// observation request pre-pass: log2 of the (unit-scaled) times and frequencies
__global__ void k_prep_obs(const double* __restrict__ t, int n_t, const double* __restrict__ nu, int n_nu,
                           double* lg2_t, double* t_lin, double* lg2_nu, double* nu_lin, double* nu23,
                           int nu_in_code_units) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_t) {
        const double tl = t[i] * unit::sec;
        t_lin[i] = tl;
        lg2_t[i] = log2(tl);
    }
    if (i < n_nu) {
        const double v = nu_in_code_units ? nu[i] : nu[i] * unit::Hz;
        const double l = log2(v);
        lg2_nu[i] = l;
        nu_lin[i] = v;
        nu23[i] = exp2((2. / 3) * l);
    }
}

// Distinct frequencies of a series request (one CTA).  out_n = their number, or EATS_NU_TILE + 1 when there are more
// (or a NaN); band_of[i] = index of point i's frequency in the list; the list is kept as log2 / linear / ^(2/3).
// The list grows in order of first appearance: each round appends the unmatched point of lowest index.
__global__ void __launch_bounds__(1024) k_series_bands(const double* __restrict__ lg2_nu, const double* __restrict__ nu_lin,
                                                        const double* __restrict__ nu23, int n, int* band_of, double* b_lg2,
                                                        double* b_lin, double* b_23, int* out_n) {
    __shared__ double s_u[EATS_NU_TILE + 1];
    __shared__ int s_cnt, s_cand;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const double v = i < n ? lg2_nu[i] : 0.0;
        int mine = -1;
        for (;;) {
            const int cnt = s_cnt;
            if (mine < 0)
                for (int q = 0; q < cnt; ++q)
                    if (s_u[q] == v) mine = q;
            __syncthreads();
            if (threadIdx.x == 0) s_cand = 0x7fffffff;
            __syncthreads();
            if (i < n && mine < 0) atomicMin(&s_cand, i);
            __syncthreads();
            const int cand = s_cand;
            if (cand == 0x7fffffff) break;
            if (cnt >= EATS_NU_TILE) {  // a ninth value (or a NaN, which matches nothing): not a banded request
                if (threadIdx.x == 0) *out_n = EATS_NU_TILE + 1;
                return;
            }
            if (threadIdx.x == 0) {
                s_u[cnt] = lg2_nu[cand];
                b_lg2[cnt] = lg2_nu[cand];
                b_lin[cnt] = nu_lin[cand];
                b_23[cnt] = nu23[cand];
                s_cnt = cnt + 1;
            }
            __syncthreads();
        }
        if (i < n) band_of[i] = mine;
        __syncthreads();
    }
    if (threadIdx.x == 0) *out_n = s_cnt;
}

// Observer::flux (src/core/observer.h:555-567): band[m][c][j] = sum_i F[m][c][i][j] * w[i]
__global__ void k_band_reduce(const double* __restrict__ F, const double* __restrict__ wgt, double* __restrict__ out,
                              size_t n_mc, int n_nu, int n_t) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_mc * n_t) return;
    const size_t mc = e / n_t;
    const int j = (int)(e - mc * n_t);
    const double* f = F + mc * (size_t)n_nu * n_t + j;
    double s = 0;
    for (int i = 0; i < n_nu; ++i) s += f[(size_t)i * n_t] * wgt[i];
    out[e] = s;
}

// K3.  grid = (model, split, shock).  out[model][comp][n_nu][n_t] (grid) or [model][comp][n] (series)
// MODE 0: synchrotron of shocks without ssc; MODE 1: synchrotron of shocks with ssc (IC-corrected
// spectrum); MODE 2: SSC component (per-cell tables).  blockIdx.z = shock (0 forward, 1 reverse).
#ifndef EATS_MIN_BLOCKS
#define EATS_MIN_BLOCKS 8
#endif
// KIND: the request kind, compiled separately (see eats_phase1)
enum { EATS_GRID = 0, EATS_POINT = 1, EATS_BANDED = 2 };
// MODE 1 (synchrotron with the inverse-Compton correction) carries the IC record of the cell next to the tile state: at
// 64 registers it spilled 120 M local loads per config-4 batch; 5 CTAs / SM (90 registers) hold it without spills.
template <int MODE, int KIND>
__global__ void __launch_bounds__(128, MODE == 1 ? 5 : EATS_MIN_BLOCKS) k_eats(BatchWs w, EatsRequest rq0, double* __restrict__ out, int n_split,
                                                 int row_chunk, int max_n_t, int nu_tile, size_t split_stride) {
    extern __shared__ __align__(16) double smem[];
    const int mi = blockIdx.x;
    const int split = blockIdx.y;
    const int which = (int)blockIdx.z + (MODE == 2 ? 2 : 0);  // 0 fwd sync, 1 rvs sync, 2 fwd ssc, 3 rvs ssc
    {
        const ModelCfg& cfg = w.cfg[mi];
        if ((which & 1) && !cfg.has_rvs) return;
        const bool ssc = ((which & 1) ? cfg.rvs : cfg.fwd).ssc != 0;
        if ((MODE == 0) == ssc) return;  // MODE 0 handles the shocks without ssc, MODE 1/2 those with
    }
    if (w.status[mi] & VAG_ST_CAPACITY) return;
    EatsModel M = make_eats_model(w, mi, which);
    M.breach = &w.status[mi];
    {
        // log2_softplus table -> head of the dynamic shared memory (16-byte aligned rows)
        for (int a = threadIdx.x; a < SPL_DOUBLES; a += blockDim.x) smem[a] = w.sp_lut[a];
        M.sp_lut = smem;
    }
    const int n_t = M.h->n_t;
    const int erows = M.h->n_theta * M.h->n_phi_eff;
    constexpr bool series = KIND != EATS_GRID;
    constexpr bool banded = KIND == EATS_BANDED;  // series points read a (node, band) tile staged as in grid mode
    double* acc = smem + SPL_DOUBLES;     // [acc_cols][acc_stride]
    const int acc_stride = rq0.acc_stride;
    const int acc_cols = series ? 1 : nu_tile;
    EatsShared sh = eats_carve(acc + acc_cols * acc_stride, max_n_t, series && !banded, row_chunk, nu_tile);
    EatsRequest rq = rq0;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int comp = which == 0 ? VAG_C_FWD_SYNC : which == 1 ? VAG_C_RVS_SYNC : which == 2 ? VAG_C_FWD_SSC : VAG_C_RVS_SSC;
    const size_t comp_sz = series ? (size_t)rq.n_t_obs : (size_t)rq.n_nu * rq.n_t_obs;
    // row-split batches (n_split > 1, small batches): each split owns a slab of `out` (= the partial buffer) and
    // k_sum_splits adds the slabs in split order -- deterministic, unlike an atomicAdd into one cell
    double* dst = out + (size_t)split * split_stride + ((size_t)mi * VAG_NCOMP + comp) * comp_sz;
    const int n_nu_tiles = series ? 1 : (rq.n_nu + nu_tile - 1) / nu_tile;
    const RowGeom* rowg = w.rowgeom + (size_t)mi * w.max_erows;
    // rows per pass: this model's lattice may be shorter than the batch maximum the buffers are sized for
    const int rpp = eats_rows_per_pass(n_t, row_chunk, nthr);
    for (int tile = 0; tile < n_nu_tiles; ++tile) {
        const int l0 = tile * nu_tile;
        const int nl = series ? (banded ? rq.n_bands : 1) : imin(nu_tile, rq.n_nu - l0);
        for (int i0 = 0; i0 < rq.n_t_obs; i0 += EATS_T_BLOCK) {
            rq.i0 = i0;
            rq.ni = imin(EATS_T_BLOCK, rq.n_t_obs - i0);
            // zeroed by the thread that accumulates into and finally reads the column (ii == tid mod nthr): no barrier is
            // needed even when this split has no pass to run (compute-sanitizer racecheck, profiles/r02_sanitize.txt)
            for (int ii = tid; ii < rq.ni; ii += nthr)
                for (int l = 0; l < acc_cols; ++l) acc[l * acc_stride + ii] = 0.0;
            for (int q0 = split * rpp; q0 < erows; q0 += n_split * rpp) {
                const int nrows = imin(rpp, erows - q0);
                sh.rowg = rowg + q0;
                __syncthreads();  // previous pass finished reading the staged rows (and the table is loaded)
                eats_phase1<MODE, KIND == EATS_POINT>(M, rq, sh, nrows, l0, nl, tid, nthr);
                __syncthreads();
                if (banded)
                    eats_phase2_banded(M, rq, sh, nrows, acc, tid, nthr);
                else if (series)
                    eats_phase2_series<MODE>(M, rq, sh, nrows, acc, tid, nthr);
                else
                    eats_phase2_grid(M, rq, sh, nrows, nl, acc, tid, nthr);
            }
            // accumulator columns are thread-owned (ii == tid mod nthr): no barrier needed here
            for (int ii = tid; ii < rq.ni; ii += nthr) {
                for (int l = 0; l < (series ? 1 : nl); ++l) {
                    const double v = flux_scale(M, acc[l * acc_stride + ii]);
                    double* p = series ? (dst + i0 + ii) : (dst + (size_t)(l0 + l) * rq.n_t_obs + i0 + ii);
                    *p = v;
                }
            }
            __syncthreads();
        }
    }
}

// out[e] = sum over the row splits, in split order
__global__ void k_sum_splits(double* __restrict__ out, const double* __restrict__ part, int n_split, size_t elems) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= elems) return;
    double s = 0;
    for (int q = 0; q < n_split; ++q) s += part[(size_t)q * elems + e];
    out[e] = s;
}

// total = sum of the present components (PyFlux::calc_total, pybind/pymodel.cpp:350-364)
__global__ void k_total(double* out, size_t n_models, size_t comp_sz) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_models * comp_sz) return;
    const size_t mi = i / comp_sz, e = i - mi * comp_sz;
    double* o = out + mi * VAG_NCOMP * comp_sz;
    o[e] = o[VAG_C_FWD_SYNC * comp_sz + e] + o[VAG_C_FWD_SSC * comp_sz + e] + o[VAG_C_RVS_SYNC * comp_sz + e] +
           o[VAG_C_RVS_SSC * comp_sz + e];
}

// K4: one warp per model; chi2 over the series total (fitter.py:497-501)
// A model whose ODE exhausted Boost's 500 consecutive step rejections is an exception in the reference
// (max_step_checker.hpp:99-106) that the samplers turn into logL = -inf (samplers.py:63-70): chi2 = +inf here.
// accumulate: add to chi2[] (band terms after the point term, fitter.py:525-531).
__global__ void k_chi2(const double* __restrict__ flux, size_t n_models, int n, const double* __restrict__ lnF,
                       const double* __restrict__ sigma_ln, const double* __restrict__ wgt, double* chi2,
                       const int* __restrict__ status, int accumulate) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= (int)n_models) return;
    const double* F = flux + (size_t)warp * VAG_NCOMP * n;  // total
    double s = 0;
    for (int i = lane; i < n; i += 32) s += chi2_term(lnF[i], F[i], sigma_ln[i], wgt[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        if (accumulate) s += chi2[warp];
        const bool fatal = status && (status[warp] & VAG_ST_ODE_FAIL500);
        chi2[warp] = (isfinite(s) && !fatal) ? s : kInf;
    }
}

// vag_details_ic: [VAG_IC_DETAIL_PLANES][n_cells] <- IcCell records
__global__ void k_ic_export(const IcCell* __restrict__ ic, size_t n_cells, double* __restrict__ out) {
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    const IcCell& e = ic[c];
    const double v[VAG_IC_DETAIL_PLANES] = {e.gamma_m, e.gamma_c, e.gamma_a, e.gamma_M, e.ys.gamma_m_hat, e.ys.gamma_c_hat, e.ys.Y_T};
#pragma unroll
    for (int a = 0; a < VAG_IC_DETAIL_PLANES; ++a) out[(size_t)a * n_cells + c] = v[a];
}

// Work counters of a batch (profiling only): [0] forward-only ODE rows, [1] pair ODE rows, [2] shock-table cells x shocks,
// [3] EATS cells = sum_models shocks n_phi_eff n_theta n_t, [4] EATS rows = sum shocks n_phi_eff n_theta,
// [5] theta-quadrature attempts, [6] phi-quadrature attempts x n_theta, [7] sum n_theta
__global__ void k_work(BatchWs w, double* out) {
    const int mi = blockIdx.x * blockDim.x + threadIdx.x;
    if (mi >= w.n_models) return;
    const GridHeader& h = w.hdr[mi];
    const ModelCfg& c = w.cfg[mi];
    const double S = c.has_rvs ? 2.0 : 1.0;
    atomicAdd(out + (c.has_rvs ? 1 : 0), (double)h.n_reps);
    atomicAdd(out + 2, S * h.n_reps * (double)h.n_t);
    atomicAdd(out + 3, S * h.n_phi_eff * (double)h.n_theta * h.n_t);
    atomicAdd(out + 4, S * h.n_phi_eff * (double)h.n_theta);
    atomicAdd(out + 5, (double)h.quad_attempts_theta);
    atomicAdd(out + 6, (double)h.quad_attempts_phi * h.n_theta);
    atomicAdd(out + 7, (double)h.n_theta);
}

__global__ void k_or_status(int32_t* acc, const int32_t* __restrict__ st, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) acc[i] |= st[i];
}

__global__ void k_nan_capacity(BatchWs w, double* out, size_t comp_sz_total) {
    // models whose grid exceeded the compiled capacity produce NaN, never a silent partial result
    const int mi = blockIdx.x;
    if (!(w.status[mi] & VAG_ST_CAPACITY)) return;
    for (size_t i = threadIdx.x; i < comp_sz_total; i += blockDim.x) out[(size_t)mi * comp_sz_total + i] = NAN;
}

// Self-test of vag_libm.cuh on the device: out[i] = gl::fn(x[i] [, y[i]]) (fn as in vag_selftest_libm, include/vag.h)
__global__ void k_libm_selftest(int fn, const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ out,
                                size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double a = x[i], b = y ? y[i] : 0.0;
        double r = 0;
        switch (fn) {
            case 0: r = gl::exp(a); break;
            case 1: r = gl::exp2(a); break;
            case 2: r = gl::log(a); break;
            case 3: r = gl::log2(a); break;
            case 4: r = gl::log10(a); break;
            case 5: r = gl::pow(a, b); break;
            case 6: r = gl::sin(a); break;
            case 7: r = gl::cos(a); break;
        }
        out[i] = r;
    }
}

// FP64 FMA throughput probe (roofline denominator for the ODE / radiation / EATS kernels, which are
// FP64-pipe or dependent-latency bound; MEASURED_PEAKS.json has no FP64 entry)
__global__ void __launch_bounds__(256) k_fp64_peak(double* sink, int iters) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.678) sink[0] = a0;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            return fail(VAG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

}  // namespace

struct vag_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    DevBuf model_buf, row_buf, cell_buf, obs_buf, io_params, io_t, io_nu, io_out, io_status, io_aux, ic_buf, lut_buf, sp_buf, geom_buf, io_w, io_obs, io_chi2, split_buf;
    int* h_totals = nullptr;        // pinned
    long long* h_cells = nullptr;   // pinned
    int* h_bands = nullptr;         // pinned: distinct frequencies of the series request in flight
    int series_mode = 0;            // vag_set_series_mode: 0 auto, 1 per-point spectra always, 2 banded whenever possible
    int cap_theta = 384, cap_phi = 128;            // capacities of the batch in flight (host calls derive them per batch)
    int dbg_max_ode_steps = 0, dbg_max_ode_fails = 0;
    int user_cap_theta = 384, user_cap_phi = 128;  // vag_set_capacity: what the *_dev entry points run with
    bool profiling = false;
    int out_mode = VAG_OUT_DENSE;
    int total_alias = -1;  // VAG_OUT_PRESENT_ALIAS_TOTAL: component plane the last host call's `total` equals, or -1
    cudaEvent_t ev[8] = {};
    float stage_ms[8] = {};
    double work[8] = {};   // k_work counters of the last profiled pass
    DevBuf work_buf;
    int launches = 0;
    int sm_count = 148;
    // warps of k_ic_spectrum per SM (one warp per cell in flight, 30 KB of scratch each): as many as are resident.  Measured
    // on config 4: 12 / 16 / 20 / 24 / 28 warps per SM -> 75.1 / 68.6 / 65.0 / 61.1 / 58.5 ms per batch (latency-bound, IPC 1.3)
    int ic_warps_per_sm = 16;
};

namespace {

template <class T>
T* carve(char*& p, size_t n) {
    T* r = reinterpret_cast<T*>(p);
    p += ((n * sizeof(T) + 255) / 256) * 256;
    return r;
}
template <class T>
size_t carve_sz(size_t n) {
    return ((n * sizeof(T) + 255) / 256) * 256;
}

// capacity planning from host-side parameters (upper bounds of auto_grid's node counts,
// src/core/grid-refinement.h:246-262,655,664-677)
void caps_for(const vag_params* p, size_t n, int& cap_theta, int& cap_phi) {
    cap_theta = 64;
    cap_phi = 8;
    for (size_t i = 0; i < n; ++i) {
        const bool r = p[i].has_rvs != 0;
        const double th_res = p[i].theta_resol > 0 ? p[i].theta_resol : (r ? 0.2 : 0.15);
        const double ph_res = p[i].phi_resol > 0 ? p[i].phi_resol : 0.06;
        const double lg = std::log10(std::max(1.0, std::max(p[i].Gamma0, p[i].Gamma0_w) * 1.5708));
        const int ct = 36 + (int)(90 * th_res) + (int)(std::max(0.0, lg - 1) * th_res * 55) + (int)(lg * th_res * 25) + 64;
        const int cp = std::max((int)(360 * ph_res), 1) * 5 + 8;
        cap_theta = std::max(cap_theta, ct);
        cap_phi = std::max(cap_phi, cp);
    }
}

constexpr size_t EATS_SMEM_BUDGET = 200 * 1024;  // dynamic shared memory k_eats may ask for
// models per pipeline pass: k_rowgeom / k_rowcos / k_dop_extrema index the model with gridDim.y (<= 65535), and the
// per-model workspaces of a pass stay below a few GB; larger batches are processed in consecutive passes
constexpr size_t MAX_MODELS_PER_PASS = 32768;

struct Request {
    bool series;
    const double* d_t;
    const double* d_nu;
    size_t n_t, n_nu;
    bool nu_code_units = false;
};

int setup_models(vag_context* ctx, BatchWs& w, const vag_params* d_params, size_t n) {
    w = BatchWs{};
    w.n_models = (int)n;
    w.dbg_max_ode_steps = ctx->dbg_max_ode_steps;
    w.dbg_max_ode_fails = ctx->dbg_max_ode_fails;
    w.cap_theta = ctx->cap_theta;
    w.cap_phi = ctx->cap_phi;
    w.work_per_model = grid_work_doubles(w.cap_theta, w.cap_phi);
    w.params = d_params;
    w.sp_lut = static_cast<const double*>(ctx->sp_buf.p);
    size_t bytes = carve_sz<ModelCfg>(n) + carve_sz<GridHeader>(n) + carve_sz<double>(n * w.cap_theta) * 2 +
                   carve_sz<double>(n * w.cap_phi) + carve_sz<double>(n * w.work_per_model) +
                   carve_sz<int>(n * w.cap_theta) * 2 + carve_sz<int>(n + 1) + carve_sz<long long>(n + 1) +
                   carve_sz<int>(TOT_N) + carve_sz<int>(n);
    CK(ctx->model_buf.ensure(bytes));
    char* p = static_cast<char*>(ctx->model_buf.p);
    w.cfg = carve<ModelCfg>(p, n);
    w.hdr = carve<GridHeader>(p, n);
    w.theta = carve<double>(p, n * w.cap_theta);
    w.t_dec = carve<double>(p, n * w.cap_theta);
    w.phi = carve<double>(p, n * w.cap_phi);
    w.work = carve<double>(p, n * w.work_per_model);
    w.reps = carve<int>(p, n * w.cap_theta);
    w.rep_of = carve<int>(p, n * w.cap_theta);
    w.row_off = carve<int>(p, n + 1);
    w.cell_off = carve<long long>(p, n + 1);
    w.totals = carve<int>(p, TOT_N);
    w.status = carve<int>(p, n);
    return VAG_OK;
}

int setup_rows(vag_context* ctx, BatchWs& w, int rows, long long cells, bool any_spread) {
    w.n_cells = cells;
    CK(ctx->row_buf.ensure(carve_sz<int>(rows) * 3 + carve_sz<RowDyn>(rows) + carve_sz<long long>(rows + 1)));
    char* p = static_cast<char*>(ctx->row_buf.p);
    w.row_model = carve<int>(p, rows);
    w.row_rep = carve<int>(p, rows);
    w.inj_idx = carve<int>(p, rows);
    w.row_dyn = carve<RowDyn>(p, rows);
    w.row_cell_off = carve<long long>(p, rows + 1);
    const size_t plane = carve_sz<double>((size_t)cells);
    CK(ctx->cell_buf.ensure(plane * (1 + 12 + 2 + (any_spread ? 4 : 0)) + carve_sz<double>((size_t)cells * PH_NCOEF) * 2));
    p = static_cast<char*>(ctx->cell_buf.p);
    w.t_rows = carve<double>(p, (size_t)cells);
    for (int a = 0; a < 6; ++a) w.fwd[a] = carve<double>(p, (size_t)cells);
    for (int a = 0; a < 6; ++a) w.rvs[a] = carve<double>(p, (size_t)cells);
    w.geo_u = carve<double>(p, (size_t)cells);
    w.geo_lg2r2 = carve<double>(p, (size_t)cells);
    w.coef_fwd = carve<double>(p, (size_t)cells * PH_NCOEF);
    w.coef_rvs = carve<double>(p, (size_t)cells * PH_NCOEF);
    w.sh_theta = w.geo_cth = w.geo_sth = w.geo_dcos = nullptr;
    if (any_spread) {
        w.sh_theta = carve<double>(p, (size_t)cells);
        w.geo_cth = carve<double>(p, (size_t)cells);
        w.geo_sth = carve<double>(p, (size_t)cells);
        w.geo_dcos = carve<double>(p, (size_t)cells);
    }
    return VAG_OK;
}

// inverse-Compton workspaces (only when some model of the batch has ssc=True)
int setup_ic(vag_context* ctx, BatchWs& w, size_t n, long long cells, int n_ic_warps, cudaStream_t s) {
    const size_t nc = (size_t)cells;
    const size_t bytes = 2 * (carve_sz<IcCell>(nc) + carve_sz<IcTable>(nc) + carve_sz<double>(nc * IC_CAP_OUT)) +
                         carve_sz<double>(n * w.max_erows) + 2 * carve_sz<double>(n * w.max_n_t) +
                         carve_sz<double>((size_t)n_ic_warps * IC_SCRATCH_DOUBLES);
    CK(ctx->ic_buf.ensure(bytes));
    char* p = static_cast<char*>(ctx->ic_buf.p);
    for (int a = 0; a < 2; ++a) {
        w.ic[a] = carve<IcCell>(p, nc);
        w.ictab_h[a] = carve<IcTable>(p, nc);
        w.ictab[a] = carve<double>(p, nc * IC_CAP_OUT);
    }
    w.rowcos = carve<double>(p, n * w.max_erows);
    w.dop_min = carve<double>(p, n * w.max_n_t);
    w.dop_max = carve<double>(p, n * w.max_n_t);
    w.ic_scratch = carve<double>(p, (size_t)n_ic_warps * IC_SCRATCH_DOUBLES);
    if (!ctx->lut_buf.p) {
        CK(ctx->lut_buf.ensure(sizeof(KnLut)));
        k_kn_lut<<<1, KN_LUT_N, 0, s>>>(static_cast<KnLut*>(ctx->lut_buf.p));
    }
    w.lut = static_cast<const KnLut*>(ctx->lut_buf.p);
    return VAG_OK;
}

void mark(vag_context* ctx, int i, cudaStream_t s) {
    if (ctx->profiling) cudaEventRecord(ctx->ev[i], s);
}

// grid + dynamics + radiation for a batch; fills w and the totals
int run_front(vag_context* ctx, BatchWs& w, const vag_params* d_params, size_t n, const double* d_t, size_t n_t,
              cudaStream_t s, int* totals_out, long long* cells_out) {
    int rc = setup_models(ctx, w, d_params, n);
    if (rc) return rc;
    mark(ctx, 0, s);
    if (n >= 8192)
        k_grid<8><<<(unsigned)((n * 8 + 127) / 128), 128, 0, s>>>(w, d_t, (int)n_t);
    else if (n >= 1024)
        k_grid<16><<<(unsigned)((n * 16 + 127) / 128), 128, 0, s>>>(w, d_t, (int)n_t);
    else
        k_grid<32><<<(unsigned)((n * 32 + 127) / 128), 128, 0, s>>>(w, d_t, (int)n_t);
    k_scan<<<1, 1024, 0, s>>>(w);
    ctx->launches += 2;
    CK(cudaMemcpyAsync(ctx->h_totals, w.totals, sizeof(int) * TOT_N, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(ctx->h_cells, w.cell_off + n, sizeof(long long), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const int rows = ctx->h_totals[TOT_ROWS];
    const long long cells = *ctx->h_cells;
    std::memcpy(totals_out, ctx->h_totals, sizeof(int) * TOT_N);
    *cells_out = cells;
    rc = setup_rows(ctx, w, std::max(rows, 1), std::max<long long>(cells, 1), ctx->h_totals[TOT_ANY_SPREAD] != 0);
    if (rc) return rc;
    w.max_n_t = std::max(ctx->h_totals[TOT_MAX_NT], 1);
    w.max_erows = std::max(ctx->h_totals[TOT_MAX_EROWS], 1);
    CK(ctx->geom_buf.ensure(sizeof(RowGeom) * n * (size_t)w.max_erows));
    w.rowgeom = static_cast<RowGeom*>(ctx->geom_buf.p);
    w.any_ssc = ctx->h_totals[TOT_ANY_SSC];
    if (w.any_ssc) {
        const int n_ic_warps = (int)std::min<long long>(std::max<long long>(cells, 1), (long long)ctx->sm_count * ctx->ic_warps_per_sm);
        rc = setup_ic(ctx, w, n, std::max<long long>(cells, 1), n_ic_warps, s);
        if (rc) return rc;
    }
    if (rows > 0) {
        k_rowmap<<<(unsigned)((n + 63) / 64), 64, 0, s>>>(w);
        k_rowgeom<<<dim3((unsigned)((w.max_erows + 127) / 128), (unsigned)n), 128, 0, s>>>(w);
        k_lattice<<<(unsigned)((rows + 3) / 4), 128, 0, s>>>(w, rows);
        ctx->launches += 2;
        mark(ctx, 1, s);
        {
            // rows per warp: up to one warp per scheduler before warps are filled up
            int lanes = 32;
            // Fewer rows per warp shorten a row (less divergence) but multiply the 255-register CTAs that other batches'
            // kernels have to share the SMs with.  Measured on the config-5 batch (4096 rows; 8 / 16 / 32 rows per warp):
            // one batch alone 1.80 / 1.88 / 2.03 ms, four batches in flight 1.13 / 1.33 / 1.36 M evaluations/s --
            // so warps are thinned only while they stay below two per SM.
            // (512 rows, the share of an 8-GPU ensemble: 1 / 2 / 4 / 8 rows per warp 1.56 / 1.59 / 1.64 / 1.71 ms)
            while (lanes > 1 && (rows + lanes / 2 - 1) / (lanes / 2) <= ctx->sm_count * 2) lanes >>= 1;
            const unsigned nb = (unsigned)((rows + lanes - 1) / lanes);
            if (ctx->h_totals[TOT_ANY_FWD_ONLY]) {
                k_dynamics<false><<<nb, 32, 0, s>>>(w, rows, lanes);
                ctx->launches++;
            }
            if (ctx->h_totals[TOT_ANY_PAIR]) {
                k_dynamics<true><<<nb, 32, 0, s>>>(w, rows, lanes);
                ctx->launches++;
            }
        }
        mark(ctx, 2, s);
        k_radiation<<<(unsigned)rows, 64, 0, s>>>(w);
        ctx->launches += 2;
        if (w.any_ssc) {
            k_ic_cooling<<<dim3((unsigned)((rows + 31) / 32), 2), 32, 0, s>>>(w, rows);
            ctx->launches++;
        }
    } else {
        mark(ctx, 1, s);
        mark(ctx, 2, s);
    }
    mark(ctx, 3, s);
    CK(cudaGetLastError());
    return VAG_OK;
}

template <int MODE>
void launch_eats(int kind, dim3 eg, size_t sb, cudaStream_t s, const BatchWs& w, const EatsRequest& rq, double* out,
                 int n_split, int row_chunk, int max_n_t, int nu_tile, size_t elems) {
    if (kind == EATS_GRID)
        k_eats<MODE, EATS_GRID><<<eg, 128, sb, s>>>(w, rq, out, n_split, row_chunk, max_n_t, nu_tile, elems);
    else if (kind == EATS_POINT)
        k_eats<MODE, EATS_POINT><<<eg, 128, sb, s>>>(w, rq, out, n_split, row_chunk, max_n_t, nu_tile, elems);
    else
        k_eats<MODE, EATS_BANDED><<<eg, 128, sb, s>>>(w, rq, out, n_split, row_chunk, max_n_t, nu_tile, elems);
}

int run_flux_pass(vag_context* ctx, const vag_params* d_params, size_t n, const Request& rq_in, double* d_out,
                  int32_t* d_status, const double* d_lnF, const double* d_sig, const double* d_w, double* d_chi2,
                  cudaStream_t s) {
    const size_t n_t = rq_in.n_t, n_nu = rq_in.n_nu;
    // observation arrays
    CK(ctx->obs_buf.ensure(sizeof(double) * (2 * n_t + 3 * n_nu + 8 + 3 * EATS_NU_TILE) + sizeof(int) * (n_t + 2)));
    double* lg2_t = static_cast<double*>(ctx->obs_buf.p);
    double* t_lin = lg2_t + n_t;
    double* lg2_nu = t_lin + n_t;
    double* nu_lin = lg2_nu + n_nu;
    double* nu23 = nu_lin + n_nu;
    double* nu_range = nu23 + n_nu;
    double* band_lg2 = nu_range + 8;
    double* band_lin = band_lg2 + EATS_NU_TILE;
    double* band_23 = band_lin + EATS_NU_TILE;
    int* band_of = reinterpret_cast<int*>(band_23 + EATS_NU_TILE);
    int* d_n_bands = band_of + n_t;
    {
        const size_t m = std::max(n_t, n_nu);
        k_prep_obs<<<(unsigned)((m + 127) / 128), 128, 0, s>>>(rq_in.d_t, (int)n_t, rq_in.d_nu, (int)n_nu, lg2_t, t_lin,
                                                              lg2_nu, nu_lin, nu23, rq_in.nu_code_units ? 1 : 0);
        ctx->launches++;
    }
    *ctx->h_bands = 0;
    if (rq_in.series && ctx->series_mode != 1) {
        // distinct frequencies of the request; the count reaches the host with run_front's synchronisation
        k_series_bands<<<1, 1024, 0, s>>>(lg2_nu, nu_lin, nu23, (int)n_nu, band_of, band_lg2, band_lin, band_23, d_n_bands);
        CK(cudaMemcpyAsync(ctx->h_bands, d_n_bands, sizeof(int), cudaMemcpyDeviceToHost, s));
        ctx->launches++;
    }
    BatchWs w;
    int totals[TOT_N];
    long long cells = 0;
    int rc = run_front(ctx, w, d_params, n, rq_in.d_t, n_t, s, totals, &cells);
    if (rc) return rc;

    if (w.any_ssc && totals[TOT_ROWS] > 0) {
        // SSC tables: observation band -> per-row line-of-sight cosines -> per-k Doppler extrema -> tables
        w.nu_range = nu_range;
        k_nu_range<<<1, 256, 0, s>>>(lg2_nu, (int)n_nu, nu_range);
        k_rowcos<<<dim3((unsigned)((w.max_erows + 127) / 128), (unsigned)n), 128, 0, s>>>(w);
        k_dop_extrema<<<dim3((unsigned)((w.max_n_t + 63) / 64), (unsigned)n), 64, 0, s>>>(w);
        const int n_ic_warps = (int)std::min<long long>(std::max<long long>(cells, 1), (long long)ctx->sm_count * ctx->ic_warps_per_sm);
        k_ic_spectrum<<<(unsigned)((n_ic_warps * 32 + 127) / 128), 128, 0, s>>>(w, totals[TOT_ROWS], n_ic_warps);
        ctx->launches += 4;
    }
    const size_t comp_sz = rq_in.series ? n_t : n_nu * n_t;
    CK(cudaMemsetAsync(d_out, 0, sizeof(double) * n * VAG_NCOMP * comp_sz, s));
    const int max_n_t = std::max(totals[TOT_MAX_NT], 2);
    const int max_erows = std::max(totals[TOT_MAX_EROWS], 1);
    if (totals[TOT_ROWS] > 0) {
        // shared-memory budget -> rows staged per pass
        const size_t budget = EATS_SMEM_BUDGET;
        int row_chunk = EATS_ROW_CHUNK;
        // Banded series: the boundary luminosities of every (node, band) are staged once per row instead of two spectra
        // per (point, row): n_bands * n_t tile evaluations against 2 * n_points per-point ones.  Measured on the
        // config-5 batch (5 bands, ~100-node lattices, scripts/series_ab.py): 3.18 ms banded, flat in the number of
        // points, against 1.97 ms per 100 points -- the break-even is n_bands * n_t ~ 3 n_points.
        const int n_bands_req = rq_in.series ? *ctx->h_bands : 0;
        const bool banded = n_bands_req >= 1 && n_bands_req <= EATS_NU_TILE &&
                            (ctx->series_mode == 2 || (size_t)n_bands_req * max_n_t <= 3 * n_t);
        const int nu_tile = rq_in.series ? (banded ? n_bands_req : 1) : (int)std::min<size_t>(EATS_NU_TILE, n_nu);
        const int acc_cols = rq_in.series ? 1 : nu_tile;
        auto smem_bytes = [&](int rc_) {
            return sizeof(double) * (acc_cols * eats_acc_stride((int)n_t) +
                                     eats_shared_doubles(max_n_t, rq_in.series && !banded, rc_, nu_tile) + SPL_DOUBLES);
        };
        // fewer staged rows per pass while that buys residency: 8 CTAs / SM need <= 27 KB each (227 KB per SM,
        // 1 KB per CTA reserved by the system)
        while (row_chunk > 4 && smem_bytes(row_chunk) > 27 * 1024) --row_chunk;
        while (row_chunk > 1 && smem_bytes(row_chunk) > budget) --row_chunk;
        if (smem_bytes(row_chunk) > budget)
            return fail(VAG_ERR_CAPACITY, "time lattice too long for the EATS shared-memory stage");
        // Row split: a (model, shock) whose rows do not fit one pass is cut into up to `chunks` CTAs, as many as it takes to
        // put ~2 waves of CTAs (8 resident per SM) on the device.  Tophat-on-axis batches of thousands of models need
        // none; a few hundred structured off-axis or SSC models (hundreds of rows each, BASELINE.json configs 2 and 4)
        // would otherwise run as one long CTA per model on a mostly idle device.  The splits write disjoint slabs that
        // k_sum_splits adds in a fixed order (deterministic).
        const int chunks = (max_erows + row_chunk - 1) / row_chunk;
        const int n_shock = totals[TOT_ANY_PAIR] ? 2 : 1;
        const size_t target_ctas = (size_t)ctx->sm_count * EATS_MIN_BLOCKS * 2;
        const size_t have = n * (size_t)n_shock;
        int n_split = (have < target_ctas) ? (int)std::min<size_t>(chunks, (target_ctas + have - 1) / have) : 1;
        const size_t slab_bytes = sizeof(double) * n * VAG_NCOMP * comp_sz;
        while (n_split > 1 && slab_bytes * n_split > ((size_t)1 << 30)) --n_split;  // <= 1 GiB of slabs
        n_split = std::max(n_split, 1);
        const size_t sb = smem_bytes(row_chunk);  // <= EATS_SMEM_BUDGET, the opt-in limit set once in vag_create
        EatsRequest rq{};
        rq.series = rq_in.series ? 1 : 0;
        rq.n_t_obs = (int)n_t;
        rq.n_nu = (int)n_nu;
        rq.lg2_t_obs = lg2_t;
        rq.lg2_nu_obs = banded ? band_lg2 : lg2_nu;
        rq.nu_obs_lin = banded ? band_lin : nu_lin;
        rq.nu23_obs = banded ? band_23 : nu23;
        rq.n_bands = banded ? n_bands_req : 0;
        rq.band_of = band_of;
        rq.t_obs_lin = t_lin;
        rq.acc_stride = eats_acc_stride((int)n_t);
        const dim3 eg((unsigned)n, (unsigned)n_split, (unsigned)n_shock);  // no reverse shock in the batch: no z = 1 CTAs
        const size_t elems = n * VAG_NCOMP * comp_sz;
        double* eats_out = d_out;
        if (n_split > 1) {
            CK(ctx->split_buf.ensure(sizeof(double) * elems * n_split));
            eats_out = static_cast<double*>(ctx->split_buf.p);
            CK(cudaMemsetAsync(eats_out, 0, sizeof(double) * elems * n_split, s));
        }
        const int kind = !rq_in.series ? EATS_GRID : banded ? EATS_BANDED : EATS_POINT;
        launch_eats<0>(kind, eg, sb, s, w, rq, eats_out, n_split, row_chunk, max_n_t, nu_tile, elems);
        ctx->launches++;
        if (w.any_ssc) {
            launch_eats<1>(kind, eg, sb, s, w, rq, eats_out, n_split, row_chunk, max_n_t, nu_tile, elems);
            launch_eats<2>(kind, eg, sb, s, w, rq, eats_out, n_split, row_chunk, max_n_t, nu_tile, elems);
            ctx->launches += 2;
        }
        if (n_split > 1) {
            k_sum_splits<<<(unsigned)((elems + 255) / 256), 256, 0, s>>>(d_out, eats_out, n_split, elems);
            ctx->launches++;
        }
    }
    mark(ctx, 4, s);
    {
        const size_t tot = n * comp_sz;
        k_total<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(d_out, n, comp_sz);
        ctx->launches++;
        // the SSC lattices can overflow later than the grid (k_ic_spectrum sets the bit after k_scan read the OR):
        // with ssc shocks in the batch the kernel runs unconditionally and tests each model's bit on the device
        if ((totals[TOT_STATUS_OR] & VAG_ST_CAPACITY) || w.any_ssc) {
            k_nan_capacity<<<(unsigned)n, 128, 0, s>>>(w, d_out, VAG_NCOMP * comp_sz);
            ctx->launches++;
        }
    }
    if (d_chi2) {
        k_chi2<<<(unsigned)((n * 32 + 127) / 128), 128, 0, s>>>(d_out, n, (int)n_t, d_lnF, d_sig, d_w, d_chi2, w.status, 0);
        ctx->launches++;
    }
    mark(ctx, 5, s);
    if (d_status) CK(cudaMemcpyAsync(d_status, w.status, sizeof(int) * n, cudaMemcpyDeviceToDevice, s));
    CK(cudaGetLastError());
    if (ctx->profiling) {
        CK(ctx->work_buf.ensure(sizeof(double) * 8));
        CK(cudaMemsetAsync(ctx->work_buf.p, 0, sizeof(double) * 8, s));
        k_work<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(w, static_cast<double*>(ctx->work_buf.p));
        CK(cudaMemcpyAsync(ctx->work, ctx->work_buf.p, sizeof(double) * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        for (int i = 0; i < 5; ++i) cudaEventElapsedTime(&ctx->stage_ms[i], ctx->ev[i], ctx->ev[i + 1]);
    }
    return VAG_OK;
}

int run_flux(vag_context* ctx, const vag_params* d_params, size_t n, const Request& rq, double* d_out, int32_t* d_status,
             const double* d_lnF, const double* d_sig, const double* d_w, double* d_chi2, cudaStream_t s) {
    ctx->launches = 0;
    const size_t comp_sz = rq.series ? rq.n_t : rq.n_nu * rq.n_t;
    for (size_t lo = 0; lo < n; lo += MAX_MODELS_PER_PASS) {
        const size_t m = std::min(MAX_MODELS_PER_PASS, n - lo);
        if (int rc = run_flux_pass(ctx, d_params + lo, m, rq, d_out + lo * VAG_NCOMP * comp_sz, d_status ? d_status + lo : nullptr,
                                   d_lnF, d_sig, d_w, d_chi2 ? d_chi2 + lo : nullptr, s))
            return rc;
    }
    return VAG_OK;
}

int check_ascending(const double* t, size_t n) {  // is_ascending, pybind/pybind.h:28
    for (size_t i = 1; i < n; ++i)
        if (!(t[i - 1] <= t[i])) return 0;
    return 1;
}

bool finite_pos(double x) { return std::isfinite(x) && x > 0; }

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

const char* vag_last_error(void) { return g_err.c_str(); }
const char* vag_version(void) { return "vegasafterglow_b200 0.1 (sm_100a)"; }

void vag_params_default(vag_params* p) {
    std::memset(p, 0, sizeof(*p));
    p->k_e = p->k_g = 2.0;
    p->duration = 1.0;
    p->theta_w = 0.3;
    p->E_iso_w = 1e50;
    p->Gamma0_w = 50.0;
    p->n0 = INFINITY;
    p->fwd = vag_radiation{0.1, 0.01, 2.3, 1.0, 0, 0};
    p->rvs = vag_radiation{0.1, 0.01, 2.3, 1.0, 0, 0};
    p->axisymmetric = 1;
    p->radiative_fireball = 1;
    p->wind_k_m = 2.0;
}

int vag_params_validate(const vag_params* p) {
    auto bad = [&](const std::string& m) { return fail(VAG_ERR_INVALID, m); };
    auto rng_oi = [](double v, double lo, double hi) { return std::isfinite(v) && v > lo && v <= hi; };
    if (p->jet_type < 0 || p->jet_type > 5) return bad("jet_type must be one of VAG_JET_* (0..5)");
    if (p->medium_type < 0 || p->medium_type > 1) return bad("medium_type must be 0 (ISM) or 1 (wind)");
    // PyTophatJet / PyGaussianJet / PyPowerLawJet: pybind/pymodel.cpp:47-95
    if (!rng_oi(p->theta_c, 0.0, con::pi / 2)) return bad("theta_c must be in (0, pi/2]");
    const bool has_core = p->jet_type != VAG_JET_POWERLAW_WING;
    const bool has_wing = p->jet_type >= VAG_JET_TWO_COMPONENT;
    if (has_core && !finite_pos(p->E_iso)) return bad("E_iso must be finite and > 0");
    if (has_core && !(std::isfinite(p->Gamma0) && p->Gamma0 > 1.0)) return bad("Gamma0 must be > 1");
    if (!finite_pos(p->duration)) return bad("duration must be finite and > 0");
    if ((p->jet_type == VAG_JET_POWERLAW || p->jet_type == VAG_JET_STEP_POWERLAW || p->jet_type == VAG_JET_POWERLAW_WING) &&
        (!finite_pos(p->k_e) || !finite_pos(p->k_g)))
        return bad("k_e and k_g must be finite and > 0");
    // PyTwoComponentJet / PyStepPowerLawJet / PyPowerLawWing: pybind/pymodel.cpp:97-146
    if (has_wing) {
        if (!finite_pos(p->E_iso_w)) return bad("E_iso_w must be finite and > 0");
        if (!(std::isfinite(p->Gamma0_w) && p->Gamma0_w > 1.0)) return bad("Gamma0_w must be > 1");
    }
    if (p->jet_type == VAG_JET_TWO_COMPONENT) {
        if (!rng_oi(p->theta_w, 0.0, con::pi / 2)) return bad("theta_w must be in (0, pi/2]");
        if (!(p->theta_w > p->theta_c))
            return bad("theta_w (wing angle) must be greater than theta_c (core angle), got theta_w=" +
                       std::to_string(p->theta_w) + ", theta_c=" + std::to_string(p->theta_c));
    }
    if (!(std::isfinite(p->sigma0) && p->sigma0 >= 0)) return bad("sigma0 must be finite and >= 0");
    if (p->has_magnetar) {  // PyMagnetar: pybind/pymodel.h:45-49
        if (!finite_pos(p->magnetar_L0)) return bad("L0 must be finite and > 0");
        if (!finite_pos(p->magnetar_t0)) return bad("t0 must be finite and > 0");
        if (!finite_pos(p->magnetar_q)) return bad("q must be finite and > 0");
        if (p->jet_type == VAG_JET_POWERLAW_WING) return bad("PowerLawWing takes no magnetar (pybind/pybind.cpp:220-223)");
    }
    // PyISM / PyWind: pybind/pymodel.cpp:148-186
    if (p->medium_type == VAG_MEDIUM_ISM) {
        if (!(std::isfinite(p->n_ism) && p->n_ism >= 0)) return bad("n_ism must be finite and >= 0");
    } else {
        if (!finite_pos(p->A_star)) return bad("A_star must be finite and > 0");
        if (!(std::isfinite(p->n_ism) && p->n_ism >= 0)) return bad("n_ism must be finite and >= 0");
        if (!(p->n0 > 0)) return bad("n0 must be > 0 (or +inf for no floor)");
        if (!(std::isfinite(p->wind_k_m))) return bad("k_m must be finite and > 0");
    }
    // PyObserver: pybind/pymodel.h:190-204
    if (!finite_pos(p->lumi_dist)) return bad("lumi_dist must be finite and > 0");
    if (!(std::isfinite(p->z) && p->z >= 0)) return bad("z must be finite and >= 0");
    if (!(std::isfinite(p->theta_obs) && p->theta_obs >= 0 && p->theta_obs <= con::pi))
        return bad("theta_obs must be in [0, pi]");
    if (!std::isfinite(p->phi_obs)) return bad("phi_obs must be finite");
    // PyRadiation: pybind/pymodel.h:303-313
    auto chk_rad = [&](const vag_radiation& r, const char* who) -> int {
        if (!rng_oi(r.eps_e, 0.0, 1.0)) return bad(std::string(who) + ".eps_e must be in (0, 1]");
        if (!rng_oi(r.eps_B, 0.0, 1.0)) return bad(std::string(who) + ".eps_B must be in (0, 1]");
        if (!rng_oi(r.xi_e, 0.0, 1.0)) return bad(std::string(who) + ".xi_e must be in (0, 1]");
        if (!(std::isfinite(r.p) && r.p > 1.0)) return bad(std::string(who) + ".p must be > 1");
        return VAG_OK;
    };
    if (int rc = chk_rad(p->fwd, "fwd_rad")) return rc;
    if (p->has_rvs)
        if (int rc = chk_rad(p->rvs, "rvs_rad")) return rc;
    // PyModel ctor: pybind/pymodel.h:642-647 (a non-positive value selects the default)
    if (p->rtol > 0 && !(p->rtol < 1)) return bad("rtol must be in (0, 1)");
    if (!std::isfinite(p->rtol) || !std::isfinite(p->phi_resol) || !std::isfinite(p->theta_resol) ||
        !std::isfinite(p->t_resol))
        return bad("resolutions and rtol must be finite");
    return VAG_OK;
}

int vag_create(int device, vag_context** out) {
    if (!out) return fail(VAG_ERR_INVALID, "out is NULL");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(VAG_ERR_CUDA, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count = 0"));
    if (device < 0 || device >= count) return fail(VAG_ERR_INVALID, "device index out of range");
    CK(cudaSetDevice(device));
    vag_context* c = new vag_context();
    c->device = device;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaMallocHost(&c->h_totals, sizeof(int) * TOT_N));
    CK(cudaMallocHost(&c->h_cells, sizeof(long long)));
    CK(cudaMallocHost(&c->h_bands, sizeof(int)));
    for (auto& ev : c->ev) CK(cudaEventCreate(&ev));
    {
        std::vector<double> lut(SPL_DOUBLES);
        build_softplus_lut(lut.data());
        CK(c->sp_buf.ensure(sizeof(double) * SPL_DOUBLES));
        CK(cudaMemcpy(c->sp_buf.p, lut.data(), sizeof(double) * SPL_DOUBLES, cudaMemcpyHostToDevice));
    }
    {
        // node sequences of find_theta_range (vag_grid.cuh ThetaWalk): plain IEEE running sums, identical on host and
        // device; uploaded once per device
        static std::mutex mu;
        static bool done[256] = {};
        std::lock_guard<std::mutex> lk(mu);
        if (device < 256 && !done[device]) {
            ThetaWalk tw;
            build_theta_walk(tw);
            CK(cudaMemcpyToSymbol(g_theta_walk, &tw, sizeof(tw)));
            done[device] = true;
        }
    }
    // k_grid spills its scratch to local memory: prefer L1.  k_dynamics keeps its dopri5 stage vectors
    // in shared memory (23 KB per 32-row CTA): give it the full carve-out so several CTAs share an SM.
    cudaFuncSetCacheConfig(k_grid<8>, cudaFuncCachePreferL1);
    cudaFuncSetCacheConfig(k_grid<16>, cudaFuncCachePreferL1);
    cudaFuncSetCacheConfig(k_grid<32>, cudaFuncCachePreferL1);
    cudaFuncSetAttribute(k_dynamics<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    // per-function process state: set once, to the largest request run_flux can make (not per call, where two contexts
    // with different request shapes would race on it)
#define VAG_EATS_ATTR(M_, K_) \
    CK(cudaFuncSetAttribute(k_eats<M_, K_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EATS_SMEM_BUDGET))
    VAG_EATS_ATTR(0, EATS_GRID);
    VAG_EATS_ATTR(0, EATS_POINT);
    VAG_EATS_ATTR(0, EATS_BANDED);
    VAG_EATS_ATTR(1, EATS_GRID);
    VAG_EATS_ATTR(1, EATS_POINT);
    VAG_EATS_ATTR(1, EATS_BANDED);
    VAG_EATS_ATTR(2, EATS_GRID);
    VAG_EATS_ATTR(2, EATS_POINT);
    VAG_EATS_ATTR(2, EATS_BANDED);
#undef VAG_EATS_ATTR
    {
        int blocks = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, k_ic_spectrum, 128, 0) == cudaSuccess && blocks > 0)
            c->ic_warps_per_sm = std::min(blocks, 8) * 4;
    }
    *out = c;
    return VAG_OK;
}

void vag_destroy(vag_context* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (DevBuf* b : {&c->model_buf, &c->row_buf, &c->cell_buf, &c->obs_buf, &c->io_params, &c->io_t, &c->io_nu,
                      &c->io_out, &c->io_status, &c->io_aux, &c->ic_buf, &c->lut_buf, &c->sp_buf, &c->geom_buf, &c->io_w, &c->io_obs,
                      &c->io_chi2, &c->split_buf, &c->work_buf})
        b->release();
    if (c->h_totals) cudaFreeHost(c->h_totals);
    if (c->h_cells) cudaFreeHost(c->h_cells);
    if (c->h_bands) cudaFreeHost(c->h_bands);
    for (auto& ev : c->ev)
        if (ev) cudaEventDestroy(ev);
    cudaStreamDestroy(c->stream);
    delete c;
}

int vag_set_profiling(vag_context* ctx, int enable) {
    ctx->profiling = enable != 0;
    return VAG_OK;
}
int vag_last_stage_ms(vag_context* ctx, float ms[8]) {
    for (int i = 0; i < 8; ++i) ms[i] = ctx->stage_ms[i];
    return VAG_OK;
}
int vag_last_launch_count(vag_context* ctx) { return ctx->launches; }
int vag_last_work(vag_context* ctx, double work[8]) {
    for (int i = 0; i < 8; ++i) work[i] = ctx->work[i];
    return VAG_OK;
}
int vag_synchronize(vag_context* ctx) {
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return VAG_OK;
}
int vag_set_output_mode(vag_context* ctx, int mode) {
    if (mode != VAG_OUT_DENSE && mode != VAG_OUT_PRESENT && mode != VAG_OUT_PRESENT_ALIAS_TOTAL)
        return fail(VAG_ERR_INVALID, "unknown output mode");
    ctx->out_mode = mode;
    return VAG_OK;
}
int vag_last_total_alias(vag_context* ctx) { return ctx ? ctx->total_alias : -1; }
int vag_set_series_mode(vag_context* ctx, int mode) {
    if (mode < 0 || mode > 2) return fail(VAG_ERR_INVALID, "unknown series mode");
    ctx->series_mode = mode;
    return VAG_OK;
}
int vag_set_capacity(vag_context* ctx, int cap_theta, int cap_phi) {
    if (cap_theta < 40 || cap_phi < 2) return fail(VAG_ERR_INVALID, "capacity too small");
    ctx->cap_theta = ctx->user_cap_theta = cap_theta;
    ctx->cap_phi = ctx->user_cap_phi = cap_phi;
    return VAG_OK;
}

int vag_debug_set_ode_limits(vag_context* ctx, int max_steps, int max_fails) {
    if (!ctx) return fail(VAG_ERR_INVALID, "ctx is NULL");
    ctx->dbg_max_ode_steps = max_steps > 0 ? max_steps : 0;
    ctx->dbg_max_ode_fails = max_fails > 0 ? max_fails : 0;
    return VAG_OK;
}

int vag_measure_fp64_peak(vag_context* ctx, double* tflops) {
    CK(cudaSetDevice(ctx->device));
    CK(ctx->io_aux.ensure(64));
    const int iters = 1 << 16, blocks = ctx->sm_count * 8, threads = 256;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(a, ctx->stream));
        k_fp64_peak<<<blocks, threads, 0, ctx->stream>>>(static_cast<double*>(ctx->io_aux.p), iters);
        CK(cudaEventRecord(b, ctx->stream));
        CK(cudaEventSynchronize(b));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, a, b));
        const double fl = 2.0 * 8 * (double)iters * blocks * threads;
        best = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *tflops = best;
    return VAG_OK;
}

int vag_selftest_libm(vag_context* ctx, int fn, const double* x, const double* y, double* out, size_t n) {
    if (fn < 0 || fn > 7) return fail(VAG_ERR_INVALID, "vag_selftest_libm: fn must be 0..7");
    if ((fn == 5) != (y != nullptr)) return fail(VAG_ERR_INVALID, "vag_selftest_libm: y is required for pow only");
    if (n == 0) return VAG_OK;
    CK(cudaSetDevice(ctx->device));
    CK(ctx->io_aux.ensure(3 * n * sizeof(double)));
    double* d_x = static_cast<double*>(ctx->io_aux.p);
    double* d_y = d_x + n;
    double* d_o = d_y + n;
    CK(cudaMemcpyAsync(d_x, x, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (y) CK(cudaMemcpyAsync(d_y, y, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    k_libm_selftest<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(fn, d_x, y ? d_y : nullptr, d_o, n);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, d_o, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return VAG_OK;
}

// ---- device-buffer entry points ----------------------------------------------------------------
int vag_flux_density_grid_dev(vag_context* ctx, const vag_params* d_params, size_t n_models, const double* d_t,
                              size_t n_t, const double* d_nu, size_t n_nu, double* d_out, int32_t* d_status,
                              void* stream) {
    if (!ctx) return fail(VAG_ERR_INVALID, "ctx is NULL");
    if (n_t == 0) return fail(VAG_ERR_INVALID, "time array must be non-empty");
    if (n_nu == 0) return fail(VAG_ERR_INVALID, "frequency array must be non-empty");
    ctx->cap_theta = ctx->user_cap_theta;  // host-buffer calls in between may have run with tighter per-batch capacities
    ctx->cap_phi = ctx->user_cap_phi;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    Request rq{false, d_t, d_nu, n_t, n_nu};
    return run_flux(ctx, d_params, n_models, rq, d_out, d_status, nullptr, nullptr, nullptr, nullptr, s);
}

int vag_flux_density_series_dev(vag_context* ctx, const vag_params* d_params, size_t n_models, const double* d_t,
                                const double* d_nu, size_t n, double* d_out, int32_t* d_status, void* stream) {
    if (!ctx) return fail(VAG_ERR_INVALID, "ctx is NULL");
    if (n == 0) return fail(VAG_ERR_INVALID, "time array must be non-empty");
    ctx->cap_theta = ctx->user_cap_theta;  // host-buffer calls in between may have run with tighter per-batch capacities
    ctx->cap_phi = ctx->user_cap_phi;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    Request rq{true, d_t, d_nu, n, n};
    return run_flux(ctx, d_params, n_models, rq, d_out, d_status, nullptr, nullptr, nullptr, nullptr, s);
}

int vag_chi2_series_dev(vag_context* ctx, const vag_params* d_params, size_t n_models, const double* d_t,
                        const double* d_nu, const double* d_lnF_obs, const double* d_sigma_ln, const double* d_w,
                        size_t n, double* d_chi2, int32_t* d_status, void* stream) {
    if (!ctx) return fail(VAG_ERR_INVALID, "ctx is NULL");
    if (n == 0) return fail(VAG_ERR_INVALID, "time array must be non-empty");
    ctx->cap_theta = ctx->user_cap_theta;  // host-buffer calls in between may have run with tighter per-batch capacities
    ctx->cap_phi = ctx->user_cap_phi;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    CK(ctx->io_out.ensure(sizeof(double) * n_models * VAG_NCOMP * n));
    Request rq{true, d_t, d_nu, n, n};
    return run_flux(ctx, d_params, n_models, rq, static_cast<double*>(ctx->io_out.p), d_status, d_lnF_obs, d_sigma_ln,
                    d_w, d_chi2, s);
}

// ---- host-buffer entry points --------------------------------------------------------------------
static int host_prepare(vag_context* ctx, const vag_params* params, size_t n_models, const double* t, size_t n_t,
                        const double* nu, size_t n_nu, bool series) {
    if (!ctx) return fail(VAG_ERR_INVALID, "ctx is NULL");
    if (!params && n_models) return fail(VAG_ERR_INVALID, "params is NULL");
    if (n_t == 0 || !t) return fail(VAG_ERR_INVALID, "time array must be non-empty");
    if (n_nu == 0 || !nu) return fail(VAG_ERR_INVALID, "frequency array must be non-empty");
    if (series && n_t != n_nu)
        return fail(VAG_ERR_INVALID,
                    "time and frequency arrays must have the same size\nIf you intend to get grid-like output, use "
                    "the generic `flux_density_grid` instead");
    if (!check_ascending(t, n_t)) return fail(VAG_ERR_INVALID, "time array must be in ascending order");
    // log10(t_end / t_start) sizes the time lattice: a zero, negative or non-finite epoch / frequency has no meaning
    for (size_t i = 0; i < n_t; ++i)
        if (!finite_pos(t[i])) return fail(VAG_ERR_INVALID, "observation times must be finite and > 0");
    for (size_t i = 0; i < n_nu; ++i)
        if (!finite_pos(nu[i])) return fail(VAG_ERR_INVALID, "observation frequencies must be finite and > 0");
    for (size_t i = 0; i < n_models; ++i)
        if (int rc = vag_params_validate(&params[i])) return rc;
    CK(cudaSetDevice(ctx->device));
    int ct, cp;
    caps_for(params, n_models, ct, cp);
    ctx->cap_theta = ct;
    ctx->cap_phi = cp;
    CK(ctx->io_params.ensure(sizeof(vag_params) * std::max<size_t>(n_models, 1)));
    CK(ctx->io_t.ensure(sizeof(double) * n_t));
    CK(ctx->io_nu.ensure(sizeof(double) * n_nu));
    CK(ctx->io_status.ensure(sizeof(int32_t) * std::max<size_t>(n_models, 1)));
    cudaStream_t s = ctx->stream;
    CK(cudaMemcpyAsync(ctx->io_params.p, params, sizeof(vag_params) * n_models, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->io_t.p, t, sizeof(double) * n_t, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->io_nu.p, nu, sizeof(double) * n_nu, cudaMemcpyHostToDevice, s));
    return VAG_OK;
}

// device -> host transfer of out[n_models][VAG_NCOMP][comp_sz] under the context's output mode
static int copy_out(vag_context* ctx, const vag_params* params, size_t n_models, size_t comp_sz, double* out,
                    cudaStream_t s) {
    const size_t row = sizeof(double) * comp_sz, pitch = row * VAG_NCOMP;
    if (ctx->out_mode == VAG_OUT_DENSE) {
        CK(cudaMemcpyAsync(out, ctx->io_out.p, pitch * n_models, cudaMemcpyDeviceToHost, s));
        return VAG_OK;
    }
    bool present[VAG_NCOMP] = {true, true, false, false, false};
    for (size_t i = 0; i < n_models; ++i) {
        if (params[i].fwd.ssc) present[VAG_C_FWD_SSC] = true;
        if (params[i].has_rvs) {
            present[VAG_C_RVS_SYNC] = true;
            if (params[i].rvs.ssc) present[VAG_C_RVS_SSC] = true;
        }
    }
    // exactly one emission component in the whole batch: `total` is that component, bit for bit -- do not ship it twice
    ctx->total_alias = -1;
    if (ctx->out_mode == VAG_OUT_PRESENT_ALIAS_TOTAL) {
        int n_present = 0, which = -1;
        for (int c = 1; c < VAG_NCOMP; ++c)
            if (present[c]) {
                ++n_present;
                which = c;
            }
        if (n_present == 1) {
            present[VAG_C_TOTAL] = false;
            ctx->total_alias = which;
        }
    }
    const char* src = static_cast<const char*>(ctx->io_out.p);
    char* dst = reinterpret_cast<char*>(out);
    for (int c = 0; c < VAG_NCOMP;) {
        if (!present[c]) {
            ++c;
            continue;
        }
        int c1 = c;
        while (c1 + 1 < VAG_NCOMP && present[c1 + 1]) ++c1;  // adjacent present planes travel together
        CK(cudaMemcpy2DAsync(dst + row * c, pitch, src + row * c, pitch, row * (c1 - c + 1), n_models,
                             cudaMemcpyDeviceToHost, s));
        c = c1 + 1;
    }
    return VAG_OK;
}

int vag_flux_density_grid(vag_context* ctx, const vag_params* params, size_t n_models, const double* t, size_t n_t,
                          const double* nu, size_t n_nu, double* out, int32_t* status) {
    if (int rc = host_prepare(ctx, params, n_models, t, n_t, nu, n_nu, false)) return rc;
    if (n_models == 0) return VAG_OK;
    const size_t out_bytes = sizeof(double) * n_models * VAG_NCOMP * n_nu * n_t;
    CK(ctx->io_out.ensure(out_bytes));
    cudaStream_t s = ctx->stream;
    Request rq{false, static_cast<double*>(ctx->io_t.p), static_cast<double*>(ctx->io_nu.p), n_t, n_nu};
    if (int rc = run_flux(ctx, static_cast<vag_params*>(ctx->io_params.p), n_models, rq,
                          static_cast<double*>(ctx->io_out.p), static_cast<int32_t*>(ctx->io_status.p), nullptr, nullptr,
                          nullptr, nullptr, s))
        return rc;
    if (int rc = copy_out(ctx, params, n_models, out_bytes / (sizeof(double) * n_models * VAG_NCOMP), out, s)) return rc;
    if (status) CK(cudaMemcpyAsync(status, ctx->io_status.p, sizeof(int32_t) * n_models, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return VAG_OK;
}

int vag_flux_density_series(vag_context* ctx, const vag_params* params, size_t n_models, const double* t,
                            const double* nu, size_t n, double* out, int32_t* status) {
    if (int rc = host_prepare(ctx, params, n_models, t, n, nu, n, true)) return rc;
    if (n_models == 0) return VAG_OK;
    const size_t out_bytes = sizeof(double) * n_models * VAG_NCOMP * n;
    CK(ctx->io_out.ensure(out_bytes));
    cudaStream_t s = ctx->stream;
    Request rq{true, static_cast<double*>(ctx->io_t.p), static_cast<double*>(ctx->io_nu.p), n, n};
    if (int rc = run_flux(ctx, static_cast<vag_params*>(ctx->io_params.p), n_models, rq,
                          static_cast<double*>(ctx->io_out.p), static_cast<int32_t*>(ctx->io_status.p), nullptr, nullptr,
                          nullptr, nullptr, s))
        return rc;
    if (int rc = copy_out(ctx, params, n_models, out_bytes / (sizeof(double) * n_models * VAG_NCOMP), out, s)) return rc;
    if (status) CK(cudaMemcpyAsync(status, ctx->io_status.p, sizeof(int32_t) * n_models, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return VAG_OK;
}

// Frequency grid (code units) and quadrature weights of PyModel::flux (pybind/pymodel.cpp:391-410): xt::logspace of the
// unit-scaled bounds, Boole's rule in ln(nu) with 3/8 / Simpson / trapezoid remainders (src/core/quadrature.h:138-191),
// times the Jacobian nu and the code-unit -> Hz factor, so that sum_i F_nu[cgs](nu_i) w_i is erg cm^-2 s^-1.
static int band_grid(double nu_min, double nu_max, size_t num_nu, std::vector<double>& nu, std::vector<double>& wgt) {
    if (!(nu_min > 0) || !std::isfinite(nu_min)) return fail(VAG_ERR_INVALID, "nu_min must be positive");
    if (!(nu_max > nu_min) || !std::isfinite(nu_max)) return fail(VAG_ERR_INVALID, "nu_max must be greater than nu_min");
    if (num_nu < 2) return fail(VAG_ERR_INVALID, "num_nu must be at least 2");
    nu.assign(num_nu, 0.0);
    wgt.assign(num_nu, 0.0);
    const double a = std::log10(nu_min * unit::Hz), b = std::log10(nu_max * unit::Hz);
    const double step = (b - a) / std::fmax(1.0, (double)(num_nu - 1));
    for (size_t i = 0; i < num_nu; ++i) nu[i] = std::pow(10.0, (i == num_nu - 1) ? b : a + step * (double)i);
    const double h = std::log(nu[1] / nu[0]);
    const double cb = 2.0 * h / 45.0;
    size_t j = 0;
    for (; j + 4 < num_nu; j += 4) {
        wgt[j] += cb * 7;
        wgt[j + 1] += cb * 32;
        wgt[j + 2] += cb * 12;
        wgt[j + 3] += cb * 32;
        wgt[j + 4] += cb * 7;
    }
    const size_t remaining = num_nu - 1 - j;
    if (remaining == 3) {
        const double c38 = 3.0 * h / 8.0;
        wgt[j] += c38;
        wgt[j + 1] += c38 * 3;
        wgt[j + 2] += c38 * 3;
        wgt[j + 3] += c38;
    } else if (remaining == 2) {
        const double c13 = h / 3.0;
        wgt[j] += c13;
        wgt[j + 1] += c13 * 4;
        wgt[j + 2] += c13;
    } else if (remaining == 1) {
        wgt[j] += 0.5 * h;
        wgt[j + 1] += 0.5 * h;
    }
    for (size_t i = 0; i < num_nu; ++i) wgt[i] = wgt[i] * nu[i] / unit::Hz;
    return VAG_OK;
}

// Band-integrated flux of a batch whose parameters are already on the device: grid evaluation on the band's
// frequency nodes, then the weighted sum over frequency.  *d_band -> [n_models][VAG_NCOMP][n_t] inside io_out.
static int run_band(vag_context* ctx, size_t n_models, const double* t, size_t n_t, double nu_min, double nu_max,
                    size_t num_nu, double** d_band, cudaStream_t s) {
    std::vector<double> nu, wgt;
    if (int rc = band_grid(nu_min, nu_max, num_nu, nu, wgt)) return rc;
    if (!check_ascending(t, n_t)) return fail(VAG_ERR_INVALID, "time array must be in ascending order");
    for (size_t i = 0; i < n_t; ++i)
        if (!finite_pos(t[i])) return fail(VAG_ERR_INVALID, "observation times must be finite and > 0");
    CK(ctx->io_t.ensure(sizeof(double) * n_t));
    CK(ctx->io_nu.ensure(sizeof(double) * num_nu));
    CK(ctx->io_w.ensure(sizeof(double) * num_nu));
    const size_t grid_elems = n_models * VAG_NCOMP * num_nu * n_t;
    CK(ctx->io_out.ensure(sizeof(double) * (grid_elems + n_models * VAG_NCOMP * n_t)));
    // (the previous pass over these staging buffers has completed on this stream; pageable copies are staged by the runtime)
    CK(cudaMemcpyAsync(ctx->io_t.p, t, sizeof(double) * n_t, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->io_nu.p, nu.data(), sizeof(double) * num_nu, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->io_w.p, wgt.data(), sizeof(double) * num_nu, cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));  // nu / wgt are locals
    double* d_grid = static_cast<double*>(ctx->io_out.p);
    *d_band = d_grid + grid_elems;
    Request rq{false, static_cast<double*>(ctx->io_t.p), static_cast<double*>(ctx->io_nu.p), n_t, num_nu, true};
    if (int rc = run_flux(ctx, static_cast<vag_params*>(ctx->io_params.p), n_models, rq, d_grid,
                          static_cast<int32_t*>(ctx->io_status.p), nullptr, nullptr, nullptr, nullptr, s))
        return rc;
    const size_t n_mc = n_models * VAG_NCOMP;
    k_band_reduce<<<(unsigned)((n_mc * n_t + 255) / 256), 256, 0, s>>>(d_grid, static_cast<double*>(ctx->io_w.p), *d_band,
                                                                     n_mc, (int)num_nu, (int)n_t);
    ctx->launches++;
    CK(cudaGetLastError());
    return VAG_OK;
}

// Replaces PyModel::flux (pybind/pymodel.cpp:391-410): band-integrated flux over [nu_min, nu_max]
int vag_flux_band(vag_context* ctx, const vag_params* params, size_t n_models, const double* t, size_t n_t,
                  double nu_min, double nu_max, size_t num_nu, double* out, int32_t* status) {
    std::vector<double> nu, wgt;
    if (int rc = band_grid(nu_min, nu_max, num_nu, nu, wgt)) return rc;
    if (int rc = host_prepare(ctx, params, n_models, t, n_t, nu.data(), num_nu, false)) return rc;
    if (n_models == 0) return VAG_OK;
    cudaStream_t s = ctx->stream;
    double* d_band = nullptr;
    if (int rc = run_band(ctx, n_models, t, n_t, nu_min, nu_max, num_nu, &d_band, s)) return rc;
    CK(cudaMemcpyAsync(out, d_band, sizeof(double) * n_models * VAG_NCOMP * n_t, cudaMemcpyDeviceToHost, s));
    if (status) CK(cudaMemcpyAsync(status, ctx->io_status.p, sizeof(int32_t) * n_models, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return VAG_OK;
}

// Replaces Fitter._evaluate (VegasAfterglow/fitting/fitter.py:503-533) for a batch of parameter sets: the chi-squared of the
// point data (flux_density series) plus one term per band-integrated data set (Model.flux), all in ln-flux space.
int vag_chi2(vag_context* ctx, const vag_params* params, size_t n_models, const double* t, const double* nu,
             const double* lnF_obs, const double* sigma_ln, const double* w, size_t n_points, const vag_band_obs* bands,
             size_t n_bands, double* chi2, int32_t* status) {
    if (!ctx || !chi2) return fail(VAG_ERR_INVALID, "NULL argument");
    if (n_points && (!t || !nu || !lnF_obs || !sigma_ln || !w)) return fail(VAG_ERR_INVALID, "point data arrays must not be NULL");
    if (n_bands && !bands) return fail(VAG_ERR_INVALID, "bands is NULL");
    if (n_points == 0 && n_bands == 0) return fail(VAG_ERR_INVALID, "no data: n_points and n_bands are both 0");
    for (size_t b = 0; b < n_bands; ++b) {
        const vag_band_obs& B = bands[b];
        if (B.n == 0 || !B.t || !B.lnF_obs || !B.sigma_ln || !B.w) return fail(VAG_ERR_INVALID, "band data arrays must be non-empty");
        std::vector<double> gnu, gw;
        if (int rc = band_grid(B.nu_min, B.nu_max, B.num_nu, gnu, gw)) return rc;
    }
    // validation + upload of the parameters (and of the point request when there is one)
    const double one = 1.0;
    if (int rc = n_points ? host_prepare(ctx, params, n_models, t, n_points, nu, n_points, true)
                          : host_prepare(ctx, params, n_models, bands[0].t, bands[0].n, &one, 1, false))
        return rc;
    if (n_models == 0) return VAG_OK;
    cudaStream_t s = ctx->stream;
    size_t max_obs = n_points;
    for (size_t b = 0; b < n_bands; ++b) max_obs = std::max(max_obs, bands[b].n);
    CK(ctx->io_obs.ensure(sizeof(double) * 3 * max_obs));
    CK(ctx->io_chi2.ensure(sizeof(double) * n_models + sizeof(int32_t) * n_models));
    double* d_lnF = static_cast<double*>(ctx->io_obs.p);
    double* d_sig = d_lnF + max_obs;
    double* d_w = d_sig + max_obs;
    double* d_chi2 = static_cast<double*>(ctx->io_chi2.p);
    int32_t* d_stat_or = reinterpret_cast<int32_t*>(d_chi2 + n_models);
    CK(cudaMemsetAsync(d_chi2, 0, sizeof(double) * n_models, s));
    CK(cudaMemsetAsync(d_stat_or, 0, sizeof(int32_t) * n_models, s));
    int launches = 0;
    auto upload = [&](const double* a, const double* b_, const double* c, size_t n) -> int {
        CK(cudaMemcpyAsync(d_lnF, a, sizeof(double) * n, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(d_sig, b_, sizeof(double) * n, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(d_w, c, sizeof(double) * n, cudaMemcpyHostToDevice, s));
        return VAG_OK;
    };
    const unsigned cg = (unsigned)((n_models * 32 + 127) / 128);
    if (n_points) {
        if (int rc = upload(lnF_obs, sigma_ln, w, n_points)) return rc;
        CK(ctx->io_out.ensure(sizeof(double) * n_models * VAG_NCOMP * n_points));
        Request rq{true, static_cast<double*>(ctx->io_t.p), static_cast<double*>(ctx->io_nu.p), n_points, n_points};
        if (int rc = run_flux(ctx, static_cast<vag_params*>(ctx->io_params.p), n_models, rq, static_cast<double*>(ctx->io_out.p),
                              static_cast<int32_t*>(ctx->io_status.p), nullptr, nullptr, nullptr, nullptr, s))
            return rc;
        k_chi2<<<cg, 128, 0, s>>>(static_cast<double*>(ctx->io_out.p), n_models, (int)n_points, d_lnF, d_sig, d_w, d_chi2,
                                  static_cast<int32_t*>(ctx->io_status.p), 1);
        k_or_status<<<(unsigned)((n_models + 255) / 256), 256, 0, s>>>(d_stat_or, static_cast<int32_t*>(ctx->io_status.p), n_models);
        launches += ctx->launches + 2;
    }
    for (size_t b = 0; b < n_bands; ++b) {
        const vag_band_obs& B = bands[b];
        double* d_band = nullptr;
        if (int rc = run_band(ctx, n_models, B.t, B.n, B.nu_min, B.nu_max, B.num_nu, &d_band, s)) return rc;
        if (int rc = upload(B.lnF_obs, B.sigma_ln, B.w, B.n)) return rc;
        k_chi2<<<cg, 128, 0, s>>>(d_band, n_models, (int)B.n, d_lnF, d_sig, d_w, d_chi2, static_cast<int32_t*>(ctx->io_status.p), 1);
        k_or_status<<<(unsigned)((n_models + 255) / 256), 256, 0, s>>>(d_stat_or, static_cast<int32_t*>(ctx->io_status.p), n_models);
        launches += ctx->launches + 2;
        CK(cudaStreamSynchronize(s));  // the observation staging buffers are reused by the next band
    }
    ctx->launches = launches;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(chi2, d_chi2, sizeof(double) * n_models, cudaMemcpyDeviceToHost, s));
    if (status) CK(cudaMemcpyAsync(status, d_stat_or, sizeof(int32_t) * n_models, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return VAG_OK;
}

// per-parameter-set validation of a batch: ok[i] = 1 where vag_params_validate accepts params[i] (a sampler masks the
// rejected walkers to logL = -inf the way the reference maps the constructor's exception, samplers.py:63-70)
int vag_params_validate_batch(const vag_params* params, size_t n, int32_t* ok) {
    if ((!params || !ok) && n) return fail(VAG_ERR_INVALID, "NULL argument");
    for (size_t i = 0; i < n; ++i) ok[i] = vag_params_validate(&params[i]) == VAG_OK ? 1 : 0;
    return VAG_OK;
}

int vag_chi2_series(vag_context* ctx, const vag_params* params, size_t n_models, const double* t, const double* nu,
                    const double* lnF_obs, const double* sigma_ln, const double* w, size_t n, double* chi2,
                    int32_t* status) {
    if (!lnF_obs || !sigma_ln || !w || !chi2) return fail(VAG_ERR_INVALID, "data arrays must not be NULL");
    if (int rc = host_prepare(ctx, params, n_models, t, n, nu, n, true)) return rc;
    if (n_models == 0) return VAG_OK;
    cudaStream_t s = ctx->stream;
    CK(ctx->io_out.ensure(sizeof(double) * n_models * VAG_NCOMP * n));
    CK(ctx->io_aux.ensure(sizeof(double) * (3 * n + n_models)));
    double* d_lnF = static_cast<double*>(ctx->io_aux.p);
    double* d_sig = d_lnF + n;
    double* d_w = d_sig + n;
    double* d_chi2 = d_w + n;
    CK(cudaMemcpyAsync(d_lnF, lnF_obs, sizeof(double) * n, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_sig, sigma_ln, sizeof(double) * n, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_w, w, sizeof(double) * n, cudaMemcpyHostToDevice, s));
    Request rq{true, static_cast<double*>(ctx->io_t.p), static_cast<double*>(ctx->io_nu.p), n, n};
    if (int rc = run_flux(ctx, static_cast<vag_params*>(ctx->io_params.p), n_models, rq,
                          static_cast<double*>(ctx->io_out.p), static_cast<int32_t*>(ctx->io_status.p), d_lnF, d_sig,
                          d_w, d_chi2, s))
        return rc;
    CK(cudaMemcpyAsync(chi2, d_chi2, sizeof(double) * n_models, cudaMemcpyDeviceToHost, s));
    if (status) CK(cudaMemcpyAsync(status, ctx->io_status.p, sizeof(int32_t) * n_models, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return VAG_OK;
}

// ---- introspection ---------------------------------------------------------------------------------
int vag_details(vag_context* ctx, const vag_params* p, double t_min, double t_max, vag_grid_info* info, double* theta,
                double* phi, int32_t* reps, double* t_rows, double* fwd_shock, double* rvs_shock, int32_t* inj_idx) {
    if (!ctx || !p) return fail(VAG_ERR_INVALID, "NULL argument");
    if (int rc = vag_params_validate(p)) return rc;
    CK(cudaSetDevice(ctx->device));
    int ct, cp;
    caps_for(p, 1, ct, cp);
    ctx->cap_theta = ct;
    ctx->cap_phi = cp;
    cudaStream_t s = ctx->stream;
    CK(ctx->io_params.ensure(sizeof(vag_params)));
    CK(ctx->io_t.ensure(sizeof(double) * 2));
    const double tt[2] = {t_min, t_max};
    CK(cudaMemcpyAsync(ctx->io_params.p, p, sizeof(vag_params), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->io_t.p, tt, sizeof(tt), cudaMemcpyHostToDevice, s));
    BatchWs w;
    int totals[TOT_N];
    long long cells = 0;
    ctx->launches = 0;
    if (int rc = run_front(ctx, w, static_cast<vag_params*>(ctx->io_params.p), 1, static_cast<double*>(ctx->io_t.p), 2,
                           s, totals, &cells))
        return rc;
    GridHeader h;
    int st = 0;
    CK(cudaMemcpyAsync(&h, w.hdr, sizeof(h), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&st, w.status, sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (info) {
        info->n_phi = h.n_phi;
        info->n_theta = h.n_theta;
        info->n_t = h.n_t;
        info->n_reps = h.n_reps;
        info->symmetry = h.symmetry;
        info->phi_mirrored = h.phi_mirrored;
        info->n_phi_eff = h.n_phi_eff;
        info->status = st;
    }
    const size_t nc = (size_t)h.n_reps * h.n_t;
    if (theta) CK(cudaMemcpyAsync(theta, w.theta, sizeof(double) * h.n_theta, cudaMemcpyDeviceToHost, s));
    if (phi) CK(cudaMemcpyAsync(phi, w.phi, sizeof(double) * h.n_phi, cudaMemcpyDeviceToHost, s));
    // theta index of every ODE row: the representatives of the symmetry groups, or r % n_theta for a rows3d model
    if (reps && !h.rows3d) CK(cudaMemcpyAsync(reps, w.reps, sizeof(int) * h.n_reps, cudaMemcpyDeviceToHost, s));
    if (reps && h.rows3d)
        for (int r = 0; r < h.n_reps; ++r) reps[r] = r % h.n_theta;
    if (t_rows) CK(cudaMemcpyAsync(t_rows, w.t_rows, sizeof(double) * nc, cudaMemcpyDeviceToHost, s));
    if (inj_idx) CK(cudaMemcpyAsync(inj_idx, w.inj_idx, sizeof(int) * h.n_reps, cudaMemcpyDeviceToHost, s));
    std::vector<double> th_host(h.n_theta);
    std::vector<int> reps_host(h.n_reps);
    CK(cudaMemcpyAsync(th_host.data(), w.theta, sizeof(double) * h.n_theta, cudaMemcpyDeviceToHost, s));
    if (!h.rows3d) CK(cudaMemcpyAsync(reps_host.data(), w.reps, sizeof(int) * h.n_reps, cudaMemcpyDeviceToHost, s));
    if (h.rows3d)
        for (int r = 0; r < h.n_reps; ++r) reps_host[r] = r % h.n_theta;
    auto dump = [&](double* const* pl, double* o) -> int {
        const int map[7] = {0, 1, -1, 2, 3, 4, 5};
        for (int a = 0; a < 7; ++a)
            if (map[a] >= 0) CK(cudaMemcpyAsync(o + (size_t)a * nc, pl[map[a]], sizeof(double) * nc, cudaMemcpyDeviceToHost, s));
        return VAG_OK;
    };
    if (fwd_shock)
        if (int rc = dump(w.fwd, fwd_shock)) return rc;
    if (rvs_shock && p->has_rvs)
        if (int rc = dump(w.rvs, rvs_shock)) return rc;
    CK(cudaStreamSynchronize(s));
    auto fill_theta = [&](double* o) {
        for (int r = 0; r < h.n_reps; ++r)
            for (int k = 0; k < h.n_t; ++k) o[(2 * (size_t)h.n_reps + r) * h.n_t + k] = th_host[reps_host[r]];
    };
    if (fwd_shock) fill_theta(fwd_shock);
    if (fwd_shock && w.sh_theta) {  // spreading model: Shock::theta varies along k
        CK(cudaMemcpyAsync(fwd_shock + 2 * nc, w.sh_theta, sizeof(double) * nc, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    }
    if (rvs_shock && p->has_rvs) fill_theta(rvs_shock);
    return VAG_OK;
}

int vag_details_photons(vag_context* ctx, const vag_params* p, double t_min, double t_max, double* fwd, double* rvs) {
    if (!ctx || !p || !fwd) return fail(VAG_ERR_INVALID, "NULL argument");
    if (int rc = vag_params_validate(p)) return rc;
    CK(cudaSetDevice(ctx->device));
    int ct, cp;
    caps_for(p, 1, ct, cp);
    ctx->cap_theta = ct;
    ctx->cap_phi = cp;
    cudaStream_t s = ctx->stream;
    CK(ctx->io_params.ensure(sizeof(vag_params)));
    CK(ctx->io_t.ensure(sizeof(double) * 2));
    const double tt[2] = {t_min, t_max};
    CK(cudaMemcpyAsync(ctx->io_params.p, p, sizeof(vag_params), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->io_t.p, tt, sizeof(tt), cudaMemcpyHostToDevice, s));
    BatchWs w;
    int totals[TOT_N];
    long long cells = 0;
    ctx->launches = 0;
    if (int rc = run_front(ctx, w, static_cast<vag_params*>(ctx->io_params.p), 1, static_cast<double*>(ctx->io_t.p), 2,
                           s, totals, &cells))
        return rc;
    const int planes[6] = {PH_LOG2_NU_M, PH_LOG2_NU_C, PH_LOG2_NU_A, PH_LOG2_NU_M_MAX, PH_LOG2_I_MAX, PH_INV_NU_M_MAX};
    const size_t nc = (size_t)cells;
    for (int a = 0; a < 6; ++a) {
        CK(cudaMemcpyAsync(fwd + a * nc, w.coef_fwd + (size_t)planes[a] * nc, sizeof(double) * nc, cudaMemcpyDeviceToHost, s));
        if (rvs && p->has_rvs)
            CK(cudaMemcpyAsync(rvs + a * nc, w.coef_rvs + (size_t)planes[a] * nc, sizeof(double) * nc, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));
    return VAG_OK;
}

// Electron / inverse-Compton bookkeeping of the shocks with ssc=True (after KN_cooling / Thomson_cooling), per unique cell
int vag_details_ic(vag_context* ctx, const vag_params* p, double t_min, double t_max, double* fwd, double* rvs) {
    if (!ctx || !p || !fwd) return fail(VAG_ERR_INVALID, "NULL argument");
    if (int rc = vag_params_validate(p)) return rc;
    CK(cudaSetDevice(ctx->device));
    int ct, cp;
    caps_for(p, 1, ct, cp);
    ctx->cap_theta = ct;
    ctx->cap_phi = cp;
    cudaStream_t s = ctx->stream;
    CK(ctx->io_params.ensure(sizeof(vag_params)));
    CK(ctx->io_t.ensure(sizeof(double) * 2));
    const double tt[2] = {t_min, t_max};
    CK(cudaMemcpyAsync(ctx->io_params.p, p, sizeof(vag_params), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->io_t.p, tt, sizeof(tt), cudaMemcpyHostToDevice, s));
    BatchWs w;
    int totals[TOT_N];
    long long cells = 0;
    ctx->launches = 0;
    if (int rc = run_front(ctx, w, static_cast<vag_params*>(ctx->io_params.p), 1, static_cast<double*>(ctx->io_t.p), 2,
                           s, totals, &cells))
        return rc;
    const size_t nc = (size_t)cells;
    CK(ctx->io_out.ensure(sizeof(double) * VAG_IC_DETAIL_PLANES * std::max<size_t>(nc, 1)));
    double* d_out = static_cast<double*>(ctx->io_out.p);
    for (int which = 0; which < 2; ++which) {
        double* dst = which ? rvs : fwd;
        if (which && !(rvs && p->has_rvs)) continue;
        const bool ssc = (which ? p->rvs.ssc : p->fwd.ssc) != 0;
        if (ssc && w.any_ssc && nc > 0) {
            k_ic_export<<<(unsigned)((nc + 127) / 128), 128, 0, s>>>(w.ic[which], nc, d_out);
            CK(cudaMemcpyAsync(dst, d_out, sizeof(double) * VAG_IC_DETAIL_PLANES * nc, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
        } else {
            std::memset(dst, 0, sizeof(double) * VAG_IC_DETAIL_PLANES * nc);
        }
    }
    CK(cudaGetLastError());
    return VAG_OK;
}

}  // extern "C"
