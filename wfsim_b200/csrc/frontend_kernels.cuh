// Sampling kernels of the front end (included by frontend.cu only).
//
// Every random quantity is philox(seed, stream, instruction identity, draw) -- see philox.cuh --
// so photons do not depend on batching or on which GPU simulates the instruction.
#pragma once
#include "frontend.cuh"
#include "sampling.cuh"

namespace wfs {

struct GenCtx {
    // instructions (batch-local SoA)
    int32_t *i_type;
    int64_t *i_time;
    float *i_x, *i_y, *i_z;
    int32_t *i_amp;
    uint64_t *i_gidx;
    double *i_lce, *i_scg, *i_cy;
    int32_t *i_pat;
    double *i_vd, *i_dl;          // per-instruction drift velocity / longitudinal diffusion, or nullptr
    double *i_xo, *i_yo;          // observed xy (field distortion), or nullptr
    double *i_hsr, *i_hsa;        // transverse-diffusion sigma (radial, azimuthal) [cm], or nullptr
    // 'simple' luminescence with per-position gas gaps (s2.py:317-378): gap of the instruction, largest gap
    // of its S2 call, field scale E0, and the mean emission time subtracted (filled by k_instr); or nullptr
    double *i_lgap, *i_lgapmax, *i_le0, *i_lavgt;
    double lumw_alpha, lumw_ue, lumw_pressure, lumw_ra, lumw_rw, lumw_dr;
    int32_t *i_recoil;
    int32_t *i_lrow;              // garfield luminescence: table row of the instruction
    // 'garfield_gas_gap' luminescence (s2.py:411-483): table, per-instruction rows / fraction, and the
    // per-instruction mean of the drawn excitation times (filled by k_gg_sum + k_gg_mean)
    const double *gg_cdf;
    int32_t gg_rows, gg_len;
    int32_t *i_gglo, *i_gghi;
    double *i_ggfrac, *i_ggmean;
    // externally supplied photons (wfs_instr_maps.opt_*): first list index / count per instruction
    // (count 0: sampled as usual), the lists, the cutoff; all nullptr when the call has none
    const int64_t *i_optfirst;
    const int32_t *i_optn;
    const int32_t *opt_ch;
    const int64_t *opt_t;
    int64_t opt_cutoff;
    double *i_dmean, *i_dspread;
    uint32_t *i_nemit, *i_emitoff;
    int64_t *i_nhits;
    int64_t *i_acc;
    // pattern CDF rows
    double *cdf;          // [rows][n_ch]
    int32_t *cdf_ok;      // [rows]
    const uint16_t *cdf_guide;   // [rows][kCdfGuide + 1], see k_pattern_cdf
    // emitters
    int64_t *e_t;
    int32_t *e_instr;
    uint32_t *e_nph, *e_phoff;
    // photons
    int64_t *ph_t;
    int32_t *ph_ch;
    double *ph_gain;
    int32_t *ph_instr;
    uint8_t *ph_flags, *ph_nap;
    uint32_t *ap_off;
    // tables
    const double *spe_ppf;
    const int32_t *spe_row;
    int32_t spe_len;
    const double *lum_cdf, *lum_t;
    const uint32_t *lum_guide;      // [kLumGuide + 1] guide of the search in lum_cdf (see k_pattern_cdf)
    int32_t lum_len;
    int32_t n_ap;
    int32_t ap_is_uniform[WFS_MAX_AP_ELEMENTS];
    const double *ap_delay_cdf[WFS_MAX_AP_ELEMENTS];
    int32_t ap_delay_len[WFS_MAX_AP_ELEMENTS];
    double ap_delay_bin[WFS_MAX_AP_ELEMENTS];
    const double *ap_amp_cdf[WFS_MAX_AP_ELEMENTS];
    int32_t ap_amp_len[WFS_MAX_AP_ELEMENTS], ap_amp_rows[WFS_MAX_AP_ELEMENTS];
    double ap_amp_bin[WFS_MAX_AP_ELEMENTS];
    const double *pi_time, *pi_prob;
    int32_t pi_len;
    const double *s1_op_top, *s1_op_bottom, *s2_op_top, *s2_op_bottom;
    int32_t s1_op_nz, s1_op_nu, s2_op_nu;
    double s1_op_z0, s1_op_z1, s1_op_u0, s1_op_u1, s2_op_u0, s2_op_u1;
    const int32_t *gf_t;
    const double *gf_x;
    int32_t gf_rows, gf_cols;
    uint64_t seed;
};

// Linear interpolation on a regular grid, extrapolating linearly outside it (what scipy's
// RegularGridInterpolator(method='linear', bounds_error=False, fill_value=None) returns).
__device__ __forceinline__ void grid_cell(double x, double lo, double hi, int n, int &i, double &w) {
    const double f = (x - lo) / (hi - lo) * (double)(n - 1);
    i = (int)floor(f);
    i = max(0, min(i, n - 2));
    w = f - (double)i;
}
__device__ __forceinline__ double grid_interp1(const double *tab, int n, double lo, double hi, double x) {
    if (n < 2) return tab[0];
    int i; double w;
    grid_cell(x, lo, hi, n, i, w);
    return tab[i] * (1.0 - w) + tab[i + 1] * w;
}
__device__ __forceinline__ double grid_interp2(const double *tab, int n0, int n1, double lo0, double hi0,
                                               double lo1, double hi1, double x0, double x1) {
    int i, j; double wi, wj;
    grid_cell(x0, lo0, hi0, n0, i, wi);
    grid_cell(x1, lo1, hi1, n1, j, wj);
    const double *r0 = tab + (int64_t)i * n1, *r1 = r0 + n1;
    return (r0[j] * (1.0 - wj) + r0[j + 1] * wj) * (1.0 - wi) + (r1[j] * (1.0 - wj) + r1[j + 1] * wj) * wi;
}

// first index in [0, n) with a[i] > key  (a ascending)
template <typename T, typename K>
__device__ __forceinline__ uint32_t upper_bound_dev(const T *a, uint32_t n, K key) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (a[mid] <= key) lo = mid + 1; else hi = mid;
    }
    return lo;
}
// first index in [0, n) with a[i] >= key
template <typename T, typename K>
__device__ __forceinline__ uint32_t lower_bound_dev(const T *a, uint32_t n, K key) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void k_widen_u8(const uint8_t *a, uint32_t *b, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = a[i];
}

// Pattern row -> normalised CDF (np.random.choice(p=...) builds exactly this: cumsum, /= last;
// s1.py:148-158, s2.py:646-677).  One thread per row.
// Beside every row a guide for the search of k_photons: guide[j] = first channel whose CDF value exceeds j / kCdfGuide,
// so that the channel of u lies in [guide[j], guide[j + 1]] for j = floor(u * kCdfGuide) -- a couple of entries
// instead of log2(n_ch) dependent loads (the result is the same upper bound).
constexpr int kCdfGuide = 256;
__global__ void k_pattern_cdf(int64_t rows, int n_ch, const float *pattern, const double *gains,
                              double *cdf, int32_t *ok, uint16_t *guide) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float *p = pattern + r * n_ch;
    double *c = cdf + r * n_ch;
    double s = 0.0;
    bool bad = false;
    for (int ch = 0; ch < n_ch; ch++) {
        double v = gains[ch] != 0.0 ? (double)p[ch] : 0.0;   // turned-off PMTs get no photons
        if (isnan(v)) bad = true;
        s += v;
        c[ch] = s;
    }
    if (bad || !(s > 0.0)) { ok[r] = 0; return; }
    for (int ch = 0; ch < n_ch; ch++) c[ch] /= s;
    ok[r] = 1;
    uint16_t *gd = guide + r * (kCdfGuide + 1);
    int i = 0;
    for (int j = 0; j <= kCdfGuide; j++) {
        const double thr = (double)j / (double)kCdfGuide;
        while (i < n_ch && c[i] <= thr) i++;
        gd[j] = (uint16_t)i;
    }
}

// Pattern rows of the instructions whose map lives on the device as a regular grid (rows >=
// first_dev_row): multilinear interpolation, linear extrapolation outside the grid -- the arithmetic of
// scipy's RegularGridInterpolator(fill_value=None) behind straxen.InterpolatingMap, in the operation
// order of wfsim_b200.resource.GridMap so that host- and device-evaluated rows are bit-identical.
// S1 (s1.py:148): (x, y, z); S2-like (s2.py:637-645): observed (x, y), top-only maps padded with ones.
// One warp per instruction, lanes over the PMTs (coalesced reads of the grid rows).
__global__ void __launch_bounds__(128)
k_pattern_eval(uint32_t n_instr, int32_t first_dev_row, const int32_t *__restrict__ i_type,
               const float *__restrict__ x, const float *__restrict__ y, const float *__restrict__ z,
               const double *__restrict__ xo, const double *__restrict__ yo,
               const int32_t *__restrict__ i_pat, PatGrid g1, PatGrid g2, int n_ch, float *__restrict__ pattern) {
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n_instr) return;
    const int32_t row = i_pat[i];
    if (row < first_dev_row) return;
    const bool is_s1 = i_type[i] == 1;
    const PatGrid &g = is_s1 ? g1 : g2;
    double p[3] = {0.0, 0.0, 0.0};
    if (is_s1) { p[0] = (double)x[i]; p[1] = (double)y[i]; p[2] = (double)z[i]; }
    else { p[0] = xo ? xo[i] : (double)x[i]; p[1] = yo ? yo[i] : (double)y[i]; }
    int64_t idx[3] = {0, 0, 0};
    double w[3] = {0.0, 0.0, 0.0};
    for (int d = 0; d < g.nd; d++) {
        const double f = (p[d] - g.lo[d]) / (g.hi[d] - g.lo[d]) * (double)(g.n[d] - 1);
        double fl = floor(f);
        int64_t i0 = isnan(fl) ? 0 : (fl < -1e18 ? (int64_t)-1000000000000000000ll : (fl > 1e18 ? (int64_t)1000000000000000000ll : (int64_t)fl));
        i0 = i0 < 0 ? 0 : (i0 > g.n[d] - 2 ? g.n[d] - 2 : i0);
        idx[d] = i0;
        w[d] = f - (double)i0;
    }
    const int nd = g.nd, ncorner = 1 << nd;
    float *out = pattern + (int64_t)row * n_ch;
    for (int ch = lane; ch < n_ch; ch += 32) {
        double acc = 1.0;                 // columns the map does not have (bottom array of a top-only S2 map)
        if (ch < g.npmt) {
            acc = 0.0;
            for (int corner = 0; corner < ncorner; corner++) {
                double wt = 1.0;
                int64_t flat = 0;
                for (int d = 0; d < nd; d++) {
                    const int bit = (corner >> (nd - 1 - d)) & 1;
                    flat = flat * g.n[d] + idx[d] + bit;
                    wt = wt * (bit ? w[d] : 1.0 - w[d]);
                }
                acc = acc + wt * g.v[flat * g.npmt + ch];
            }
        }
        out[ch] = (float)acc;
    }
}

// 'simple' S2 luminescence with a per-position gas gap (S2._luminescence_timings_simple, s2.py:317-341).
// An electron drifts from the liquid surface (r = dG from the wire) to the anode wire (r = rW) with velocity
// alpha * E(r), E(r) = E0 * rr(r), rr = clip(1/r, 1/rA, 1/rW), and emits photons at the rate
// dy(r) = E(r)/uE - 0.8 p per unit length.  The reference tabulates both on a radial grid of step dr that
// starts at the LARGEST gap of the call, subtracts the yield-weighted mean time over that whole grid
// (avgt, s2.py:329-331), and inverts the yield CDF from the instruction's own gap down (s2.py:333-339).
//   lumw_mean_time: avgt with the reference's own discrete sums (r_k = gapmax - k dr > rW).
__device__ inline double lumw_mean_time(const GenCtx &g, double gapmax, double e0) {
    const double dr = g.lumw_dr, inv_ra = 1.0 / g.lumw_ra, inv_rw = 1.0 / g.lumw_rw;
    const int64_t n = (int64_t)ceil((gapmax - g.lumw_rw) / dr);          // len(np.arange(gapmax, rW, -dr))
    double t = 0.0, num = 0.0, den = 0.0;
    for (int64_t k = 0; k < n; k++) {
        const double r = gapmax - (double)k * dr;
        const double rr = fmin(fmax(1.0 / r, inv_ra), inv_rw);
        t += dr / (g.lumw_alpha * e0 * rr);
        const double dy = e0 * rr / g.lumw_ue - 0.8 * g.lumw_pressure;
        num += t * dy;
        den += dy;
    }
    return den != 0.0 ? num / den : 0.0;
}
//   lumw_time: emission time of one photon -- the inverse of the yield CDF in closed form (the continuum
//   limit of np.interp(U, cumsum(dy)/sum, cumsum(dt)); the grid step is 0.1 um, i.e. < 1 ns): uniform
//   field above rA (linear), 1/r field below (a few Newton steps).
__device__ inline double lumw_time(const GenCtx &g, double gap, double e0, double avgt, double u) {
    const double ra = g.lumw_ra, rw = g.lumw_rw, c = e0 / g.lumw_ue, q = 0.8 * g.lumw_pressure;
    const double ae = g.lumw_alpha * e0;
    const double rb = fmin(ra, gap);                        // top of the 1/r region
    const double dy_a = c / ra - q;                         // yield density in the uniform-field region
    const double y_a = gap > ra ? dy_a * (gap - ra) : 0.0;
    const double y_b = c * log(rb / rw) - q * (rb - rw);
    const double target = u * (y_a + y_b);
    if (target < y_a) return (target / dy_a) * ra / ae - avgt;
    const double t_a = gap > ra ? (gap - ra) * ra / ae : 0.0;
    const double yb = target - y_a;
    double r = rb * exp(-yb / c);                           // exact for q = 0
#pragma unroll 1
    for (int it = 0; it < 8; it++) {
        const double f = c * log(rb / r) - q * (rb - r) - yb;
        const double step = f / (c / r - q);                // d/dr of the yield integral is -(c/r - q)
        r = fmin(fmax(r + step, rw), rb);
        if (fabs(step) < 1e-12) break;
    }
    return t_a + (rb * rb - r * r) / (2.0 * ae) - avgt;
}

// Transverse diffusion of the S2 hit pattern (S2.s2_pattern_map_diffuse, s2.py:560-613): the pattern of an
// S2 instruction is the average of the pattern map over its electrons, each displaced by
// N(0, sigma_r) along the radius and N(0, sigma_a) along the azimuth (rotated by theta = atan2(y, x),
// s2.py:588-594); electrons displaced beyond tpc_radius are left out of the average (s2.py:597-599), none
// left -> NaN row (np.average of nothing), whose photons get channel -1.  One CTA per instruction: the
// threads draw 128 electrons at a time into shared memory (grid cell + weights), then every thread
// accumulates its PMTs over them in electron order -- a fixed order, so the row is reproducible.
// Interpolation arithmetic as in k_pattern_eval.  Rows of instructions without electrons keep the
// pattern at the observed position (they emit nothing).
constexpr int kDiffuseMaxPerThread = 8;
__global__ void __launch_bounds__(128)
k_pattern_diffuse(GenCtx g, int32_t first_dev_row, PatGrid pg, int n_ch, double tpc_radius,
                  float *__restrict__ pattern) {
    __shared__ int32_t s_i0[128], s_i1[128];
    __shared__ double s_w0[128], s_w1[128];
    const uint32_t i = blockIdx.x;
    const int tid = threadIdx.x;
    const int32_t row = g.i_pat[i];
    const uint32_t ne = g.i_nemit[i];
    if (g.i_type[i] == 1 || row < first_dev_row || ne == 0) return;
    const double x0 = g.i_xo ? g.i_xo[i] : (double)g.i_x[i], y0 = g.i_yo ? g.i_yo[i] : (double)g.i_y[i];
    const double theta = atan2(y0, x0);
    const double ct = cos(theta), st = sin(theta);
    const double sr = g.i_hsr[i], sa = g.i_hsa[i];
    const uint64_t gidx = g.i_gidx[i];
    double acc[kDiffuseMaxPerThread];
#pragma unroll
    for (int k = 0; k < kDiffuseMaxPerThread; k++) acc[k] = 0.0;
    uint32_t n_in = 0;
    for (uint32_t base = 0; base < ne; base += 128) {
        const uint32_t e = base + tid;
        if (e < ne) {
            Rng rng(g.seed, RS_HDIFF, gidx, e * 2u);
            const double hr = normal_d(rng) * sr, ha = normal_d(rng) * sa;
            const double pos[2] = {x0 + (ct * hr - st * ha), y0 + (st * hr + ct * ha)};
            int32_t cell[2];
            double w[2];
            for (int d = 0; d < 2; d++) {
                const double f = (pos[d] - pg.lo[d]) / (pg.hi[d] - pg.lo[d]) * (double)(pg.n[d] - 1);
                const double fl = floor(f);
                int64_t c = isnan(fl) ? 0 : (fl < -1e9 ? (int64_t)-1000000000 : (fl > 1e9 ? (int64_t)1000000000 : (int64_t)fl));
                c = c < 0 ? 0 : (c > pg.n[d] - 2 ? pg.n[d] - 2 : c);
                cell[d] = (int32_t)c;
                w[d] = f - (double)c;
            }
            const bool inside = pos[0] * pos[0] + pos[1] * pos[1] <= tpc_radius * tpc_radius;
            s_i0[tid] = inside ? cell[0] : -1;
            s_i1[tid] = cell[1];
            s_w0[tid] = w[0];
            s_w1[tid] = w[1];
        }
        __syncthreads();
        const uint32_t m = min(128u, ne - base);
        for (uint32_t k = 0; k < m; k++) {
            const int32_t c0 = s_i0[k];
            if (c0 < 0) continue;
            const int32_t c1 = s_i1[k];
            const double w0 = s_w0[k], w1 = s_w1[k];
            const double *v00 = pg.v + ((int64_t)c0 * pg.n[1] + c1) * pg.npmt;
            const double *v10 = v00 + (int64_t)pg.n[1] * pg.npmt;
            n_in++;
#pragma unroll
            for (int q = 0; q < kDiffuseMaxPerThread; q++) {
                const int ch = tid + 128 * q;
                if (ch < pg.npmt) {
                    double a = 0.0;
                    a = a + ((1.0 - w0) * (1.0 - w1)) * v00[ch];
                    a = a + ((1.0 - w0) * w1) * v00[pg.npmt + ch];
                    a = a + (w0 * (1.0 - w1)) * v10[ch];
                    a = a + (w0 * w1) * v10[pg.npmt + ch];
                    acc[q] += a;
                }
            }
        }
        __syncthreads();
    }
    float *out = pattern + (int64_t)row * n_ch;
#pragma unroll
    for (int q = 0; q < kDiffuseMaxPerThread; q++) {
        const int ch = tid + 128 * q;
        if (ch < pg.npmt) out[ch] = (float)(acc[q] / (double)n_in);      // n_in == 0 -> NaN
    }
    for (int ch = pg.npmt + tid; ch < n_ch; ch += 128) out[ch] = 1.0f;      // top-only map: s2.py:642-644
}

// Per instruction: yields.  S1: s1.py:117-135.  S2-like: s2.py:157-179, 212-256.
__global__ void k_instr(GenCtx g, wfs_params p, uint32_t i0, uint32_t i1) {
    uint32_t i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= i1) return;
    const int type = g.i_type[i];
    const int64_t amp = g.i_amp[i];
    Rng rng(g.seed, RS_INSTR, g.i_gidx[i], 0);
    uint32_t nemit = 0;
    int64_t nhits = 0;
    if (type == 1 && g.i_optn && g.i_optn[i] > 0) {      // photons supplied by the caller (rawdata.py:478-495)
        nhits = g.i_optn[i];
        nemit = 1u;
    } else if (type == 1) {
        double ly = g.i_lce[i] / (1.0 + p.p_double_pe_emision) * p.s1_detection_efficiency;
        ly = fmin(fmax(ly, 0.0), 1.0);
        nhits = sample_binomial(amp, ly, rng.ud53());
        nemit = nhits > 0 ? 1u : 0u;
    } else if (type == 2 || type == 4 || type == 6) {
        const double z = (double)g.i_z[i];
        const double vd = g.i_vd ? g.i_vd[i] : p.drift_velocity_liquid;       // s2.py:139-155
        const double dl = g.i_dl ? g.i_dl[i] : p.diffusion_constant_longitudinal;
        double mean = -z / vd + p.drift_time_gate;
        if (mean < 0.0) mean = 0.0;
        double spread = sqrt(2.0 * dl * mean) / vd;
        if (p.s2_luminescence_model == 1 && g.gf_rows > 0) {
            // distance to the nearest anode wire -> nearest row of the garfield table (s2.py:396-404)
            double d;
            if (p.s2_garfield_confine_position > 0.0) {
                d = (2.0 * rng.ud53() - 1.0) * p.s2_garfield_confine_position;
            } else {
                const double x = g.i_xo ? g.i_xo[i] : (double)g.i_x[i], y = g.i_yo ? g.i_yo[i] : (double)g.i_y[i];
                const double rel = -x * sin(p.anode_xaxis_angle) + y * cos(p.anode_xaxis_angle);
                double m = fmod(rel + 0.5 * p.anode_pitch, p.anode_pitch);
                if (m < 0.0) m += p.anode_pitch;           // python modulo
                d = m - 0.5 * p.anode_pitch;
            }
            int best = 0;
            double bd = fabs(d - g.gf_x[0]);
            for (int r = 1; r < g.gf_rows; r++) {
                const double dd = fabs(d - g.gf_x[r]);
                if (dd < bd) { bd = dd; best = r; }
            }
            g.i_lrow[i] = best;
        }
        if (p.s2_luminescence_model == 0 && g.i_lgap) g.i_lavgt[i] = lumw_mean_time(g, g.i_lgapmax[i], g.i_le0[i]);
        double cy = p.electron_extraction_yield * exp(-mean / p.electron_lifetime_liquid) * g.i_cy[i];
        cy = fmin(fmax(cy, 0.0), 1.0);
        int64_t ne = sample_binomial(amp, cy, rng.ud53());
        g.i_dmean[i] = mean;
        g.i_dspread[i] = spread;
        nemit = (uint32_t)ne;
    }
    g.i_nhits[i] = nhits;
    g.i_nemit[i] = nemit;
}

// Per emitter: S1 -> the interaction itself; S2 -> one electron (s2.py:258-286, 300-310).
__global__ void k_emitters(GenCtx g, wfs_params p, uint32_t n_instr, uint32_t e0, uint32_t e1) {
    uint32_t e = e0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= e1) return;
    uint32_t i = upper_bound_dev(g.i_emitoff, n_instr + 1, e) - 1;
    const int type = g.i_type[i];
    int64_t t = g.i_time[i];
    int64_t nph;
    if (type == 1) {
        nph = g.i_nhits[i];
    } else {
        const uint32_t j = e - g.i_emitoff[i];
        Rng rng(g.seed, RS_ELECTRON, g.i_gidx[i], j * 2u);
        double tt = -log(1.0 - rng.ud53()) * p.electron_trapping_time;
        tt += g.i_dmean[i] + g.i_dspread[i] * normal_d(rng);
        t += (int64_t)tt;   // int(): truncation toward zero
        nph = sample_poisson(g.i_scg[i], rng.ud53());
        if (p.s2_gain_spread > 0.0) nph += (int64_t)(normal_d(rng) * p.s2_gain_spread);
        if (nph < 0) nph = 0;
    }
    g.e_t[e] = t;
    g.e_instr[e] = (int32_t)i;
    g.e_nph[e] = (uint32_t)nph;
}

// `guide` (optional, x in [0, 1)): guide[c] = upper bound of c / kLumGuide in xp, built on the host at wfs_create --
// the search then runs between guide[c] and guide[c + 1] for c = floor(x * kLumGuide) (same result)
constexpr int kLumGuide = 1024;
__device__ __forceinline__ double interp_table(const double *xp, const double *fp, int n, double x,
                                               const uint32_t *guide = nullptr) {
    // np.interp semantics for ascending xp
    if (x <= xp[0]) return fp[0];
    if (x >= xp[n - 1]) return fp[n - 1];
    uint32_t j;
    if (guide) {
        const int cell = (int)(x * (double)kLumGuide);
        const uint32_t lo = guide[cell], hi = guide[cell + 1];
        j = lo + upper_bound_dev(xp + lo, hi - lo, x) - 1;
    } else {
        j = upper_bound_dev(xp, (uint32_t)n, x) - 1;
    }
    double slope = (fp[j + 1] - fp[j]) / (xp[j + 1] - xp[j]);
    return slope * (x - xp[j]) + fp[j];
}

// Excitation time of one photon in the 'garfield_gas_gap' model (s2.py:441-449): the inverse CDF is the
// linear blend of the two rows around the local gas gap, sampled at U(0, len - 2) with linear
// interpolation between neighbouring entries.
__device__ __forceinline__ double gg_time(const GenCtx &g, int lo_row, int hi_row, double frac, uint32_t u32) {
    const double s = u01_32(u32) * (double)(g.gg_len - 2);
    const double fl = floor(s);
    const int k0 = (int)fl, k1 = (int)ceil(s);
    const double *lo = g.gg_cdf + (int64_t)lo_row * g.gg_len, *hi = g.gg_cdf + (int64_t)hi_row * g.gg_len;
    const double t1 = (hi[k0] - lo[k0]) * frac + lo[k0];
    const double t2 = (hi[k1] - lo[k1]) * frac + lo[k1];
    return (t2 - t1) * (s - fl) + t1;
}

// Sum of the excitation times of the photons of every S2-like instruction, in a FIXED order (thread-strided
// partial sums, shuffle tree, warps in order), so that the mean subtracted in k_photons (s2.py:450-451) --
// and with it every photon time -- is reproducible.  grid (instructions, slices); one partial per slice.
__global__ void __launch_bounds__(128)
k_gg_sum(GenCtx g, uint32_t i0, uint32_t i1, double *partial) {
    __shared__ double sm[4];
    const uint32_t i = i0 + blockIdx.x;
    if (i >= i1) return;
    const uint32_t ny = gridDim.y, yi = blockIdx.y;
    double acc = 0.0;
    if (g.i_type[i] != 1) {
        const uint32_t q0 = g.e_phoff[g.i_emitoff[i]], q1 = g.e_phoff[g.i_emitoff[i + 1]];
        const uint64_t n = q1 - q0;
        const uint32_t a = (uint32_t)(n * yi / ny), b = (uint32_t)(n * (yi + 1) / ny);
        const int lo = g.i_gglo[i], hi = g.i_gghi[i];
        const double frac = g.i_ggfrac[i];
        const uint64_t gidx = g.i_gidx[i];
        for (uint32_t ord = a + threadIdx.x; ord < b; ord += blockDim.x) {
            const Philox4 w0 = philox4x32(g.seed, RS_PHOTON, gidx, ord * 3u);
            acc += gg_time(g, lo, hi, frac, w0.v[1]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) partial[(int64_t)(i - i0) * ny + yi] = ((sm[0] + sm[1]) + sm[2]) + sm[3];
}

__global__ void k_gg_mean(GenCtx g, uint32_t i0, uint32_t i1, int ny, const double *partial) {
    const uint32_t i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= i1) return;
    double s = 0.0;
    for (int y = 0; y < ny; y++) s += partial[(int64_t)(i - i0) * ny + y];
    const uint32_t n = g.e_phoff[g.i_emitoff[i + 1]] - g.e_phoff[g.i_emitoff[i]];
    g.i_ggmean[i] = n ? s / (double)n : 0.0;
}

// The emitter of every photon, left in ph_instr for k_photons (which replaces it by the instruction): a warp takes 32
// emitters and writes the index of each over its photons, lanes side by side.  (k_photons used to search e_phoff per
// photon: ~18 dependent loads, 40 % of its stall samples and a fifth of its instructions.)
__global__ void __launch_bounds__(256)
k_photon_emitter(GenCtx g, uint32_t e0, uint32_t e1) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t e = e0 + blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t q0 = e < e1 ? g.e_phoff[e] : 0u, q1 = e < e1 ? g.e_phoff[e + 1] : 0u;
    for (int k = 0; k < 32; k++) {
        const uint32_t a = __shfl_sync(0xffffffffu, q0, k), b = __shfl_sync(0xffffffffu, q1, k);
        const int32_t ek = (int32_t)(e - lane + (uint32_t)k);
        for (uint32_t q = a + lane; q < b; q += 32u) g.ph_instr[q] = ek;
    }
}

// Per photon: channel (s1.py:138-159 / s2.py:616-682), arrival time (s1.py:162-238 /
// s2.py:504-557, pulse.py:321-341), then the PMT stage (pulse.py:53-56,76-79,95-103).
__global__ void __launch_bounds__(256)
k_photons(GenCtx g, wfs_params p, const double *gains, int n_ch, uint32_t n_emit, uint32_t p0,
          uint32_t p1) {
    uint32_t ph = p0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (ph >= p1) return;
    const uint32_t em = (uint32_t)g.ph_instr[ph];        // left there by k_photon_emitter
    // (tried before that: one search per warp and walking on from there, 3.5 vs 2.8 ms per 2e4 events; the search
    // narrowed to the emitters of the CTA's first and last photon, 305 vs 297 us per batch)
    const int32_t i = g.e_instr[em];
    const uint32_t ord = ph - g.e_phoff[g.i_emitoff[i]];
    const int type = g.i_type[i];
    const uint64_t gidx = g.i_gidx[i];
    const Philox4 w0 = philox4x32(g.seed, RS_PHOTON, gidx, ord * 3u);
    const Philox4 w1 = philox4x32(g.seed, RS_PHOTON, gidx, ord * 3u + 1u);
    const Philox4 w2 = philox4x32(g.seed, RS_PHOTON, gidx, ord * 3u + 2u);
    // channel
    int ch = -1;
    const int row = g.i_pat[i];
    if (g.cdf_ok[row]) {
        const double u = u01_32(w0.v[0]);
        const uint16_t *gd = g.cdf_guide + (int64_t)row * (kCdfGuide + 1);
        const int cell = (int)(u * (double)kCdfGuide);           // u < 1
        const uint32_t lo = gd[cell], hi = gd[cell + 1];
        ch = (int)(lo + upper_bound_dev(g.cdf + (int64_t)row * n_ch + lo, hi - lo, u));
        if (ch >= n_ch) ch = n_ch - 1;
    }
    float zs, zt;
    normal_pair(w1.v[0], w1.v[1], zs, zt);
    int64_t t = g.e_t[em];
    const bool external = type == 1 && g.i_optn && g.i_optn[i] > 0;
    if (external) {                                            // channel and time come from the lists
        const int64_t k = g.i_optfirst[i] + ord;
        const int64_t dt_k = g.opt_t[k];
        ch = g.opt_ch[k];
        if (dt_k < 0 || dt_k >= g.opt_cutoff || ch < 0 || ch >= n_ch) ch = -1;   // dropped (rawdata.py:484-487)
        t += dt_k;
    }
    const bool top = ch >= 0 && ch < p.n_top_pmts;
    if (external) {
    } else if (type == 1) {
        if (p.s1_model_optical && g.s1_op_top && ch >= 0)      // s1.py:186-189, 241-260
            t += (int64_t)grid_interp2(top ? g.s1_op_top : g.s1_op_bottom, g.s1_op_nz, g.s1_op_nu, g.s1_op_z0,
                                       g.s1_op_z1, g.s1_op_u0, g.s1_op_u1, (double)g.i_z[i], u01_32(w2.v[1]));
        if (p.s1_model_simple) {
            t += (int64_t)(exp1(w0.v[1]) * (float)p.s1_decay_time);
            t += (int64_t)(zs * (float)p.s1_decay_spread);
        }
        if (p.s1_model_custom) {                                // s1.py:201-217, 263-337
            const int rc = g.i_recoil[i];
            if (rc == 20) {                                     // LED: uniform in the pulse length
                t += (int64_t)(u01_32(w0.v[2]) * p.led_pulse_length);
            } else {                                            // NR (0) / alpha (6): singlet / triplet decay
                const double fs = rc == 0 ? p.s1_NR_singlet_fraction : p.s1_ER_alpha_singlet_fraction;
                const double delay = u01_32(w0.v[2]) < fs ? p.singlet_lifetime_liquid : p.triplet_lifetime_liquid;
                t += (int64_t)((double)exp1(w0.v[3]) * delay);
            }
        }
    } else {
        if (p.s2_luminescence_model == 0 && g.i_lgap)
            t += (int64_t)lumw_time(g, g.i_lgap[i], g.i_le0[i], g.i_lavgt[i], u01_32(w0.v[1]));
        else if (p.s2_luminescence_model == 0 && g.lum_len > 0)
            t += (int64_t)interp_table(g.lum_cdf, g.lum_t, g.lum_len, u01_32(w0.v[1]), g.lum_guide);
        else if (p.s2_luminescence_model == 1 && g.gf_rows > 0) {   // s2.py:405-409
            const int col = (int)(((uint64_t)w0.v[1] * (uint32_t)g.gf_cols) >> 32);
            t += (int64_t)g.gf_t[(int64_t)g.i_lrow[i] * g.gf_cols + col] - (int64_t)p.gf_avgt;
        } else if (p.s2_luminescence_model == 2 && g.gg_cdf) {      // s2.py:441-451, astype(int64) at :532
            t += (int64_t)(gg_time(g, g.i_gglo[i], g.i_gghi[i], g.i_ggfrac[i], w0.v[1]) - g.i_ggmean[i]);
        }
        const double delay = u01_32(w0.v[2]) < p.singlet_fraction_gas ? p.singlet_lifetime_gas
                                                                      : p.triplet_lifetime_gas;
        t += (int64_t)((double)exp1(w0.v[3]) * delay);
        if (p.s2_time_model == 2 && g.s2_op_top && ch >= 0)    // s2.py:486-501, 542-544
            t += (int64_t)grid_interp1(top ? g.s2_op_top : g.s2_op_bottom, g.s2_op_nu, g.s2_op_u0, g.s2_op_u1,
                                       u01_32(w2.v[1]));
        else if (p.s2_time_model == 0 && p.s2_time_spread > 0.0) t += (int64_t)(zs * (float)p.s2_time_spread);
    }
    // PMT transit time spread (FWHM -> sigma), truncated toward zero
    t += (int64_t)(p.pmt_transit_time_mean + (double)zt * (p.pmt_transit_time_spread / 2.35482));
    const bool dpe = u01_32(w1.v[2]) < p.p_double_pe_emision;
    double gain = 0.0;
    if (ch >= 0) {
        const double *ppf = g.spe_ppf + (int64_t)g.spe_row[ch] * g.spe_len;
        const double gch = gains[ch];
        int k1 = (int)(((uint64_t)w1.v[3] * 2000u) >> 32) + 1;
        gain = gch * ppf[k1];
        if (dpe) {
            int k2 = (int)(((uint64_t)w2.v[0] * 2000u) >> 32) + 1;
            gain += gch * ppf[k2];
        }
    }
    // PMT afterpulse selection (afterpulse.py:189-204): count children now, fill later
    uint32_t nap = 0;
    if (g.n_ap > 0 && ch >= 0) {
        Rng ar(g.seed, RS_AP, gidx, ord * 2u);
        for (int e = 0; e < g.n_ap; e++) {
            double rU0 = (1.0 - ar.ud32()) / p.pmt_ap_modifier;
            if (dpe) rU0 *= 0.5;
            const double *dc = g.ap_delay_cdf[e] + (int64_t)ch * g.ap_delay_len[e];
            if (rU0 <= dc[g.ap_delay_len[e] - 1]) nap++;
        }
    }
    g.ph_t[ph] = t;
    g.ph_ch[ph] = ch;
    g.ph_gain[ph] = gain;
    g.ph_instr[ph] = i;
    g.ph_flags[ph] = dpe ? 1 : 0;
    g.ph_nap[ph] = (uint8_t)nap;
}

// argmin(|cdf - x|) over an ascending row, first index on ties (np.argmin)
__device__ __forceinline__ int nearest_index(const double *cdf, int n, double x) {
    int j = (int)lower_bound_dev(cdf, (uint32_t)n, x);
    if (j == 0) return 0;
    if (j == n) j = n - 1;
    double d1 = fabs(cdf[j] - x), d0 = fabs(cdf[j - 1] - x);
    if (d0 <= d1) return (int)lower_bound_dev(cdf, (uint32_t)n, cdf[j - 1]);
    return (int)lower_bound_dev(cdf, (uint32_t)n, cdf[j]);
}

// Fill the PMT-afterpulse photons behind the parents (afterpulse.py:206-243).
__global__ void k_ap_fill(GenCtx g, wfs_params p, const double *gains, uint32_t n_emit, uint32_t n_ph,
                          uint32_t out0) {
    uint32_t ph = blockIdx.x * blockDim.x + threadIdx.x;
    if (ph >= n_ph || g.ph_nap[ph] == 0) return;
    const int32_t i = g.ph_instr[ph];
    const uint32_t em = upper_bound_dev(g.e_phoff, n_emit + 1, ph) - 1;
    (void)em;
    const uint32_t ord = ph - g.e_phoff[g.i_emitoff[i]];
    const uint64_t gidx = g.i_gidx[i];
    const int ch = g.ph_ch[ph];
    const bool dpe = g.ph_flags[ph] & 1;
    const int64_t tp = g.ph_t[ph];
    uint32_t o = out0 + g.ap_off[ph];
    Rng ar(g.seed, RS_AP, gidx, ord * 2u);
    for (int e = 0; e < g.n_ap; e++) {
        double rU0 = (1.0 - ar.ud32()) / p.pmt_ap_modifier;
        if (dpe) rU0 *= 0.5;
        const int len = g.ap_delay_len[e];
        const double *dc = g.ap_delay_cdf[e] + (int64_t)ch * len;
        if (!(rU0 <= dc[len - 1])) continue;
        Rng er(g.seed, RS_AP + 100u + (uint32_t)e, gidx, ord);
        const double rU1 = 1.0 - er.ud32();
        double delay, amp;
        if (g.ap_is_uniform[e]) {
            double lo = dc[0], hi = dc[1];
            delay = (lo + (hi - lo) * er.ud32()) * g.ap_delay_bin[e];
            amp = 1.0;
        } else {
            delay = nearest_index(dc, len, rU0) * g.ap_delay_bin[e] - p.pmt_ap_t_modifier;
            const int alen = g.ap_amp_len[e];
            const double *ac = g.ap_amp_cdf[e] + (g.ap_amp_rows[e] > 1 ? (int64_t)ch * alen : 0);
            amp = nearest_index(ac, alen, rU1) * g.ap_amp_bin[e];
        }
        g.ph_t[o] = (int64_t)((double)tp + delay);
        g.ph_ch[o] = ch;
        g.ph_gain[o] = gains[ch] * amp;
        g.ph_instr[o] = i;
        g.ph_flags[o] = 2;   // afterpulse photon
        g.ph_nap[o] = 0;
        o++;
    }
}

// Per-instruction truth accumulators: one CTA per instruction, fixed reduction order.
template <typename T, typename Op>
__device__ __forceinline__ T block_reduce(T v, T *sm, Op op) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    T r = sm[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) r = op(r, sm[w]);
    return r;
}

struct OpAdd { __device__ int64_t operator()(int64_t a, int64_t b) const { return a + b; } };
struct OpMin { __device__ int64_t operator()(int64_t a, int64_t b) const { return a < b ? a : b; } };
struct OpMax { __device__ int64_t operator()(int64_t a, int64_t b) const { return a > b ? a : b; } };

// warp-wide reductions of 64-bit integers on the 32-bit reduce unit
__device__ __forceinline__ int64_t warp_sum_i64(int64_t v) {      // modulo 2^64, any sign: chunks of 24 + 24 + 16 bits
    const uint64_t u = (uint64_t)v;
    const uint32_t s0 = __reduce_add_sync(0xffffffffu, (uint32_t)(u & 0xffffffu));
    const uint32_t s1 = __reduce_add_sync(0xffffffffu, (uint32_t)((u >> 24) & 0xffffffu));
    const uint32_t s2 = __reduce_add_sync(0xffffffffu, (uint32_t)(u >> 48));
    return (int64_t)((uint64_t)s0 + ((uint64_t)s1 << 24) + ((uint64_t)s2 << 48));
}
__device__ __forceinline__ int64_t warp_min_i64(int64_t v) {
    const int32_t hi = (int32_t)(v >> 32);
    const int32_t mh = __reduce_min_sync(0xffffffffu, hi);
    const uint32_t ml = __reduce_min_sync(0xffffffffu, hi == mh ? (uint32_t)v : 0xffffffffu);
    return (int64_t)(((uint64_t)(uint32_t)mh << 32) | ml);
}
__device__ __forceinline__ int64_t warp_max_i64(int64_t v) {
    const int32_t hi = (int32_t)(v >> 32);
    const int32_t mh = __reduce_max_sync(0xffffffffu, hi);
    const uint32_t ml = __reduce_max_sync(0xffffffffu, hi == mh ? (uint32_t)v : 0u);
    return (int64_t)(((uint64_t)(uint32_t)mh << 32) | ml);
}

__device__ __forceinline__ void time_moments(int64_t rel, int64_t &s, int64_t &hi2, int64_t &hilo, int64_t &lo2) {
    int64_t hi = rel >> 12, lo = rel & 4095;
    s += rel; hi2 += hi * hi; hilo += hi * lo; lo2 += lo * lo;
}

// neutral elements of the accumulators, for the multi-CTA (atomic) form of k_instr_truth
__global__ void k_acc_init(int64_t n, int64_t *acc) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int a = (int)(k % A_COUNT);
    acc[k] = (a == A_TMIN || a == A_ETMIN) ? LLONG_MAX : (a == A_TMAX || a == A_ETMAX || a == A_PTMAX) ? LLONG_MIN : 0;
}

// One CTA per instruction; an instruction with more than kTruthSlice photons (heavy S2s with millions of photons,
// which one CTA would walk alone) is cut into ny slices: slice 0 by its own CTA, the others through the work list
// of k_truth_items (a (n_instr, ny) grid launched ~1e7 empty CTAs for a batch with a few heavy S2s and thousands of
// single-electron secondaries: 10 ms).  ny == 1: plain stores; ny > 1: integer atomics into accumulators preset by
// k_acc_init -- sums, minima and maxima of integers, so the result does not depend on the order.
constexpr uint32_t kTruthSlice = 8192, kTruthSlicesMax = 256;      // photons per CTA of a heavy instruction
__device__ __forceinline__ uint32_t truth_slices(uint32_t n_photons) {
    return min(max((n_photons + kTruthSlice - 1) / kTruthSlice, 1u), kTruthSlicesMax);
}

// Work list of the heavy instructions: the slices 1 .. ny - 1 of every instruction with more than kTruthSlice
// photons (slice 0 is taken by the instruction's own CTA).  sum(ny_i - 1) <= n_photons / kTruthSlice.
__global__ void k_truth_items(GenCtx g, uint32_t n_instr, uint2 *items, uint32_t *n_items) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_instr) return;
    const uint32_t ny = truth_slices(g.e_phoff[g.i_emitoff[i + 1]] - g.e_phoff[g.i_emitoff[i]]);
    if (ny <= 1) return;
    const uint32_t base = atomicAdd(n_items, ny - 1);
    for (uint32_t y = 1; y < ny; y++) items[base + y - 1] = make_uint2(i, y);
}

// `split`: instructions may be heavy (accumulators preset by k_acc_init, combined by atomics).  `items` == nullptr:
// CTA i takes slice 0 of instruction i; else CTA k takes the slice items[k] names (k < *n_items).
__global__ void __launch_bounds__(128)
k_instr_truth(GenCtx g, DeviceConfig c, uint32_t n_instr, uint32_t n_ph, uint32_t ap0, int split,
              const uint2 *__restrict__ items, const uint32_t *__restrict__ n_items) {
    __shared__ int64_t sm[4][A_COUNT];
    // (one warp per instruction for light batches -- no CTA-wide combination of the 35 accumulators -- was tried: slower,
    // 3.4 vs 2.8 ms of front end per 2e4 C1 events: an S2 instruction's ~1200 photons are a long walk for 32 lanes)
    const int tid = (int)threadIdx.x, nthr = (int)blockDim.x;
    uint32_t i = blockIdx.x, yi = 0;
    if (items) {
        if (blockIdx.x >= *n_items) return;
        i = items[blockIdx.x].x;
        yi = items[blockIdx.x].y;
    }
    if (i >= n_instr) return;
    const int64_t T0 = g.i_time[i];
    const uint32_t e0_all = g.i_emitoff[i], e1_all = g.i_emitoff[i + 1];
    const uint32_t q0_all = g.e_phoff[e0_all], q1_all = g.e_phoff[e1_all];
    const uint32_t ny = split ? truth_slices(q1_all - q0_all) : 1u;
    auto slice = [&](uint32_t lo, uint32_t hi, uint32_t &a, uint32_t &b) {
        const uint64_t n = hi - lo;
        a = lo + (uint32_t)(n * yi / ny);
        b = lo + (uint32_t)(n * (yi + 1) / ny);
    };
    uint32_t e0, e1, q0, q1;
    slice(e0_all, e1_all, e0, e1);
    slice(q0_all, q1_all, q0, q1);
    if (ny > 1 && yi > 0 && e1 == e0 && q1 == q0) return;   // nothing in this slice (CTA 0 still reports)
    int64_t v[A_COUNT];
#pragma unroll
    for (int a = 0; a < A_COUNT; a++) v[a] = 0;
    v[A_TMIN] = LLONG_MAX; v[A_TMAX] = LLONG_MIN; v[A_ETMIN] = LLONG_MAX; v[A_ETMAX] = LLONG_MIN;
    v[A_PTMAX] = LLONG_MIN;
    const int dt = c.p.dt;
    int r0 = (int)(T0 % dt); if (r0 < 0) r0 += dt;          // T0 mod dt (floor)
    for (uint32_t q = q0 + tid; q < q1; q += nthr) {
        const int64_t t = g.ph_t[q];
        const int ch = g.ph_ch[q];
        v[A_NPHALL]++;
        v[A_TMIN] = t < v[A_TMIN] ? t : v[A_TMIN];
        v[A_TMAX] = t > v[A_TMAX] ? t : v[A_TMAX];
        time_moments(t - T0, v[A_SREL], v[A_SHI2], v[A_SHILO], v[A_SLO2]);
        if (ch < 0) continue;
        const double gch = c.gains[ch];
        if (gch == 0.0) continue;
        v[A_PTMAX] = t > v[A_PTMAX] ? t : v[A_PTMAX];
        const double gain = g.ph_gain[q];
        const int dpe = g.ph_flags[q] & 1;
        // t mod dt (floor) from the instruction's remainder and the photon's offset to it: 32-bit arithmetic unless the
        // photon is more than 2 s from its instruction
        int r;
        const int64_t rel = t - T0;
        if (rel >= 0 && rel < ((int64_t)1 << 31)) {
            r = (int)(((uint32_t)rel % (uint32_t)dt + (uint32_t)r0) % (uint32_t)dt);
        } else {
            int64_t q_ = t / dt; r = (int)(t - q_ * dt); if (r < 0) r += dt;
        }
        const double thr = (double)(c.p.baseline - 1 - c.zle_thr[ch]) - 0.5;
        const bool above = gain * c.current_max[r] * c.p.current_2_adc > thr;
        const int64_t area = llrint(gain / gch * kAreaScale);
        const bool bottom = ch >= c.p.n_top_pmts;
        v[A_NPH]++; v[A_NDPE] += dpe; v[A_AREA] += area;
        if (above) { v[A_NTRIG]++; v[A_AREA_TRIG] += area; }
        if (bottom) {
            v[A_NPH_B]++; v[A_NDPE_B] += dpe; v[A_AREA_B] += area;
            if (above) { v[A_NTRIG_B]++; v[A_AREA_TRIG_B] += area; }
        }
    }
    if (g.i_type[i] != 1) {
        for (uint32_t e = e0 + tid; e < e1; e += nthr) {
            const int64_t t = g.e_t[e];
            v[A_NE]++;
            v[A_ETMIN] = t < v[A_ETMIN] ? t : v[A_ETMIN];
            v[A_ETMAX] = t > v[A_ETMAX] ? t : v[A_ETMAX];
            time_moments(t - T0, v[A_ESREL], v[A_EHI2], v[A_EHILO], v[A_ELO2]);
        }
    }
    if (g.n_ap > 0 && q1 > q0) {   // PMT-afterpulse children (of my photons) extend the last pulse end
        const uint32_t a0 = ap0 + g.ap_off[q0], a1 = ap0 + (q1 < n_ph ? g.ap_off[q1] : g.ap_off[n_ph]);
        for (uint32_t a = a0 + tid; a < a1; a += nthr) {
            const int64_t t = g.ph_t[a];
            v[A_NAP]++;
            v[A_PTMAX] = t > v[A_PTMAX] ? t : v[A_PTMAX];
        }
    }
    // warp partials of all accumulators, ONE barrier, then accumulator a is combined by thread a
    // (integer sums / min / max: the order does not matter, the result is deterministic)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // (warp reductions through the 32-bit reduce instruction: counters directly, 64-bit sums in three chunks, minima /
    // maxima by high word then low word -- a fifth of the instructions of 35 64-bit shuffle trees, which were three
    // quarters of this kernel for an instruction of a few hundred photons)
#pragma unroll
    for (int a = 0; a < A_COUNT; a++) {
        const bool is_min = a == A_TMIN || a == A_ETMIN, is_max = a == A_TMAX || a == A_ETMAX || a == A_PTMAX;
        const bool is_count = a == A_NPH || a == A_NDPE || a == A_NTRIG || a == A_NPH_B || a == A_NDPE_B ||
                              a == A_NTRIG_B || a == A_NPHALL || a == A_NE || a == A_NAP;
        const int64_t r = is_min ? warp_min_i64(v[a]) : is_max ? warp_max_i64(v[a])
                          : is_count ? (int64_t)__reduce_add_sync(0xffffffffu, (uint32_t)v[a]) : warp_sum_i64(v[a]);
        if (lane == 0) sm[warp][a] = r;
    }
    __syncthreads();
    if (threadIdx.x < A_COUNT) {
        const int a = threadIdx.x;
        const bool is_min = a == A_TMIN || a == A_ETMIN, is_max = a == A_TMAX || a == A_ETMAX || a == A_PTMAX;
        int64_t r = sm[0][a];
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) {
            const int64_t u = sm[w][a];
            r = is_min ? (u < r ? u : r) : is_max ? (u > r ? u : r) : r + u;
        }
        int64_t *out = &g.i_acc[(int64_t)i * A_COUNT + a];
        if (ny == 1) *out = r;
        else if (is_min) atomicMin((long long *)out, (long long)r);
        else if (is_max) atomicMax((long long *)out, (long long)r);
        else if (r) atomicAdd((unsigned long long *)out, (unsigned long long)r);
    }
}

// Photo-ionisation electrons (afterpulse.py:29-80): n_e ~ Poisson(n * N_photons * modifier) delays
// drawn from the delay histogram and binned on the coarse grid == independent Poisson counts per
// coarse bin with mean lambda * P(bin).  One warp per instruction; pass 0 counts, pass 1 fills.
__global__ void __launch_bounds__(128)
k_photoionization(GenCtx g, wfs_params p, uint32_t n_prim, int fill, uint32_t *count,
                  const uint32_t *offset, uint32_t out0, int32_t *sec_parent) {
    const int lane = threadIdx.x & 31;
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n_prim) return;
    uint32_t total = 0;
    if (g.i_type[i] == 2) {
        const uint32_t q0 = g.e_phoff[g.i_emitoff[i]], q1 = g.e_phoff[g.i_emitoff[i + 1]];
        const uint32_t nph = q1 - q0;
        if (nph > 0) {
            const double lam = p.ele_ap_n * (double)nph * p.photoionization_modifier;
            const uint64_t gidx = g.i_gidx[i];
            uint32_t base = fill ? out0 + offset[i] : 0;
            for (int b0 = 0; b0 < g.pi_len; b0 += 32) {
                const int b = b0 + lane;
                int64_t n = 0;
                Philox4 w;
                if (b < g.pi_len) {
                    w = philox4x32(g.seed, RS_PI, gidx, (uint32_t)b * 2u);
                    n = sample_poisson(lam * g.pi_prob[b], u01_53(w.v[0], w.v[1]));
                }
                const unsigned m = __ballot_sync(0xffffffffu, n > 0);
                if (fill && n > 0) {
                    const uint32_t o = base + __popc(m & ((1u << lane) - 1u));
                    const Philox4 w2 = philox4x32(g.seed, RS_PI, gidx, (uint32_t)b * 2u + 1u);
                    // random parent photon as time zero (afterpulse.py:49-56)
                    const uint32_t pick = q0 + (uint32_t)(((uint64_t)w.v[2] * nph) >> 32);
                    const double t0 = (double)g.ph_t[pick];
                    const double R = p.tpc_radius;
                    const double r = sqrt(u01_53(w2.v[0], w2.v[1]) * R * R);
                    const double ang = -3.141592653589793 + 6.283185307179586 * u01_32(w2.v[2]);
                    g.i_type[o] = 4;
                    g.i_time[o] = (int64_t)(t0 - p.drift_time_gate);
                    g.i_x[o] = (float)(r * cos(ang));
                    g.i_y[o] = (float)(r * sin(ang));
                    g.i_z[o] = (float)(-g.pi_time[b] * p.drift_velocity_liquid);
                    g.i_amp[o] = (int32_t)n;
                    g.i_gidx[o] = (1ull << 40) + gidx * 4096ull + (uint64_t)b;
                    g.i_lce[o] = 1.0;
                    g.i_scg[o] = g.i_scg[i];     // see DESIGN.md: map value of the parent position
                    g.i_cy[o] = g.i_cy[i];
                    g.i_pat[o] = g.i_pat[i];
                    g.i_recoil[o] = g.i_recoil[i];
                    if (g.i_gglo) { g.i_gglo[o] = g.i_gglo[i]; g.i_gghi[o] = g.i_gghi[i]; g.i_ggfrac[o] = g.i_ggfrac[i]; }
                    if (g.i_lgap) { g.i_lgap[o] = g.i_lgap[i]; g.i_lgapmax[o] = g.i_lgapmax[i]; g.i_le0[o] = g.i_le0[i]; }
                    if (g.i_vd) g.i_vd[o] = g.i_vd[i];
                    if (g.i_dl) g.i_dl[o] = g.i_dl[i];
                    if (g.i_xo) { g.i_xo[o] = r * cos(ang); g.i_yo[o] = r * sin(ang); }
                    sec_parent[o] = (int32_t)i;
                }
                total += __popc(m);
                base += __popc(m);
            }
        }
    }
    if (!fill && lane == 0) count[i] = total;
}

// Photo-electric (gate) electrons (afterpulse.py:105-135): n_e ~ Poisson(p * N_photons * modifier)
// single-electron instructions of type 6, delayed by N(t_center + gate, t_spread) clipped at 0
// behind a random photon of the S2.  One warp per primary instruction; pass 0 counts, pass 1 fills.
__global__ void __launch_bounds__(128)
k_photoelectric(GenCtx g, wfs_params p, uint32_t n_prim, int fill, uint32_t *count,
                const uint32_t *offset, uint32_t out0, int32_t *sec_parent) {
    const int lane = threadIdx.x & 31;
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n_prim) return;
    uint32_t n = 0;
    uint32_t q0 = 0, nph = 0;
    if (g.i_type[i] == 2) {
        q0 = g.e_phoff[g.i_emitoff[i]];
        nph = g.e_phoff[g.i_emitoff[i + 1]] - q0;
        if (nph > 0) {
            const Philox4 w = philox4x32(g.seed, RS_PE, g.i_gidx[i], 0u);
            const int64_t k = sample_poisson(p.photoelectric_p * (double)nph * p.photoelectric_modifier,
                                             u01_53(w.v[0], w.v[1]));
            n = (uint32_t)min((int64_t)(1 << 24) - 1, k);
        }
    }
    if (!fill) {
        if (lane == 0) count[i] = n;
        return;
    }
    const uint64_t gidx = g.i_gidx[i];
    for (uint32_t j = lane; j < n; j += 32) {
        const uint32_t o = out0 + offset[i] + j;
        const Philox4 a = philox4x32(g.seed, RS_PE, gidx, 1u + 2u * j), b = philox4x32(g.seed, RS_PE, gidx, 2u + 2u * j);
        const double u1 = 1.0 - u01_53(a.v[0], a.v[1]), u2 = u01_32(a.v[2]);
        const double nrm = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
        double delay = p.photoelectric_t_center + p.drift_time_gate + p.photoelectric_t_spread * nrm;
        if (delay < 0.0) delay = 0.0;
        const uint32_t pick = q0 + (uint32_t)(((uint64_t)a.v[3] * nph) >> 32);
        const double R = p.tpc_radius;
        const double r = sqrt(u01_53(b.v[0], b.v[1]) * R * R);
        const double ang = -3.141592653589793 + 6.283185307179586 * u01_32(b.v[2]);
        g.i_type[o] = 6;
        g.i_time[o] = (int64_t)((double)g.ph_t[pick] + p.drift_time_gate);
        g.i_x[o] = (float)(r * cos(ang));
        g.i_y[o] = (float)(r * sin(ang));
        g.i_z[o] = (float)(-delay * p.drift_velocity_liquid);
        g.i_amp[o] = 1;
        g.i_gidx[o] = (1ull << 62) | (gidx << 24) | (uint64_t)j;
        g.i_lce[o] = 1.0;
        g.i_scg[o] = g.i_scg[i];
        g.i_cy[o] = g.i_cy[i];
        g.i_pat[o] = g.i_pat[i];
        g.i_recoil[o] = g.i_recoil[i];
        if (g.i_gglo) { g.i_gglo[o] = g.i_gglo[i]; g.i_gghi[o] = g.i_gghi[i]; g.i_ggfrac[o] = g.i_ggfrac[i]; }
        if (g.i_lgap) { g.i_lgap[o] = g.i_lgap[i]; g.i_lgapmax[o] = g.i_lgapmax[i]; g.i_le0[o] = g.i_le0[i]; }
        if (g.i_vd) g.i_vd[o] = g.i_vd[i];
        if (g.i_dl) g.i_dl[o] = g.i_dl[i];
        if (g.i_xo) { g.i_xo[o] = r * cos(ang); g.i_yo[o] = r * sin(ang); }
        sec_parent[o] = (int32_t)i;
    }
}

}  // namespace wfs
