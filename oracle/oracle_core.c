/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the deterministic back end of the
 * WFSim hot path.  Used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs as the checker / reported baseline.  The product (wfsim_b200/) never
 * links, loads or calls this file.
 *
 * Each function cites the reference (WFSim v1.2.2, paths relative to /root/reference) it
 * restates.  Parity of this restatement is pinned by tests/test_oracle_golden.py against the
 * golden vectors produced by the reference itself (tests/golden/make_golden.py).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: the reference's numba code multiplies
 * and adds separately, so no FMA contraction is allowed).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* wfsim/core/pulse.py:276-318  Pulse.add_current
 * Photons must already be ordered by ascending time (the reference argsorts inside).
 * Equal-ns photons are merged (gains summed) before the template multiply. */
void orc_add_current(const int64_t *t, const double *g, int64_t n, int64_t pulse_left,
                     int64_t dt, const double *templates, int tlen, double *cur)
{
    if (n <= 0) return;
    double gain_total = 0.0;
    int64_t tmp = t[0];
    for (int64_t i = 0; i < n; i++) {
        if (t[i] > tmp) {
            int64_t q = tmp / dt, r = tmp % dt;
            if (r < 0) { r += dt; q -= 1; }          /* python floor semantics */
            double *c = cur + (q - pulse_left);
            const double *tm = templates + r * tlen;
            for (int k = 0; k < tlen; k++) c[k] += tm[k] * gain_total;
            gain_total = g[i];
            tmp = t[i];
        } else {
            gain_total += g[i];
        }
    }
    int64_t q = tmp / dt, r = tmp % dt;
    if (r < 0) { r += dt; q -= 1; }
    double *c = cur + (q - pulse_left);
    const double *tm = templates + r * tlen;
    for (int k = 0; k < tlen; k++) c[k] += tm[k] * gain_total;
}

/* wfsim/utils.py:13-58  find_intervals_below_threshold (strict <, hold-off merge) */
int64_t orc_find_intervals(const int64_t *w, int64_t n, int64_t threshold, int64_t holdoff,
                           int64_t *out_lr, int64_t cap)
{
    int in_itv = 0;
    int64_t cur = 0, start = -1, end = -1;
    for (int64_t i = 0; i < n; i++) {
        int64_t x = w[i];
        if (x < threshold) {
            if (!in_itv) { in_itv = 1; start = i; }
            end = i;
        }
        if ((i == n - 1 && in_itv) || (x >= threshold && i >= end + holdoff && in_itv)) {
            in_itv = 0;
            if (cur < cap) { out_lr[2 * cur] = start; out_lr[2 * cur + 1] = end; }
            cur++;
        }
    }
    return cur;
}

typedef struct {
    double current_2_adc;
    int64_t trigger_window;
    int64_t baseline;
    int64_t he_first;        /* channel_map['he'][0]; <0 when detector is not XENONnT */
    int64_t he_mult;         /* int(high_energy_deamplification_factor) */
    int64_t n_top;           /* n_top_pmts */
    int64_t n_rows;          /* 801 */
    int64_t enable_noise;
    int64_t noise_len;
    int64_t noise_nch;
    int64_t ix_rand;         /* the one random offset add_noise draws per call */
} orc_digi_cfg;

/* wfsim/core/rawdata.py:204-272 digitize_pulse_cache (+ njit helpers :392-458) followed by
 * wfsim/core/rawdata.py:274-311 ZLE, for ONE pulse cache (= one digitisation group).
 * Instead of the reference's dense (801, span) int64 array the rows are built per channel over
 * the channel window [min left - tw, max right + tw]; the values inside the window are the
 * same by construction (nothing outside a channel's window is ever read by ZLE).
 *
 * thr[ch] is the ZLE threshold per row (baseline - zle_threshold - 1 or the special one).
 * Outputs intervals as (channel, abs_left, abs_right, offset into samples[]).
 * Returns 0, or -1 if a capacity is too small (n_itv/n_samples then hold the needed sizes).
 */
int orc_digitize_zle(int64_t n_pulses, const int32_t *p_ch, const int64_t *p_left,
                     const int64_t *p_right, const int64_t *p_off, const double *currents,
                     const orc_digi_cfg *cfg, const int64_t *thr, const double *noise,
                     int64_t cap_itv, int32_t *itv_ch, int64_t *itv_left, int64_t *itv_right,
                     int64_t *itv_off, int64_t cap_samples, int16_t *samples,
                     int64_t *n_itv_out, int64_t *n_samples_out, int64_t *grp_left_right)
{
    const int64_t tw = cfg->trigger_window;
    const int64_t R = cfg->n_rows;
    int64_t *wl = (int64_t *)malloc(sizeof(int64_t) * R);
    int64_t *wr = (int64_t *)malloc(sizeof(int64_t) * R);
    char *mask = (char *)calloc(R, 1);
    int64_t gl = INT64_MAX, gr = INT64_MIN;
    for (int64_t r = 0; r < R; r++) { wl[r] = INT64_MAX; wr[r] = INT64_MIN; }
    for (int64_t i = 0; i < n_pulses; i++) {
        int64_t ch = p_ch[i];
        mask[ch] = 1;
        if (p_left[i] < wl[ch]) wl[ch] = p_left[i];
        if (p_right[i] > wr[ch]) wr[ch] = p_right[i];
        if (p_left[i] < gl) gl = p_left[i];
        if (p_right[i] > gr) gr = p_right[i];
    }
    /* rawdata.py:215-222: group window, left made even */
    int64_t left = gl - tw, right = gr + tw;
    if (left % 2 != 0) left -= 1;
    grp_left_right[0] = left;
    grp_left_right[1] = right;
    /* HE rows mirror the top-channel windows (rawdata.py:243-249) */
    if (cfg->he_first >= 0) {
        for (int64_t ch = 0; ch < cfg->n_top; ch++)
            if (mask[ch]) {
                mask[cfg->he_first + ch] = 1;
                wl[cfg->he_first + ch] = wl[ch];
                wr[cfg->he_first + ch] = wr[ch];
            }
    }
    int64_t n_itv = 0, n_samples = 0, rc = 0;
    int64_t maxlen = gr - gl + 2 * tw + 1;
    int64_t *row = (int64_t *)malloc(sizeof(int64_t) * (maxlen > 0 ? maxlen : 1));
    int64_t itv_cap_local = maxlen / 2 + 2;
    int64_t *lr = (int64_t *)malloc(sizeof(int64_t) * 2 * itv_cap_local);
    for (int64_t ch = 0; ch < R; ch++) {
        if (!mask[ch]) continue;
        int64_t a = wl[ch] - tw, b = wr[ch] + tw, len = b - a + 1;
        memset(row, 0, sizeof(int64_t) * len);
        int64_t src = ch, mult = 1;
        if (cfg->he_first >= 0 && ch >= cfg->he_first && ch < cfg->he_first + cfg->n_top) {
            src = ch - cfg->he_first;
            mult = cfg->he_mult;
        }
        /* rawdata.py:236-239: one rounding per pulse, integer sum over pulses */
        for (int64_t i = 0; i < n_pulses; i++) {
            if (p_ch[i] != src) continue;
            const double *c = currents + p_off[i];
            int64_t n = p_right[i] - p_left[i] + 1;
            int64_t *dst = row + (p_left[i] - a);
            for (int64_t k = 0; k < n; k++) {
                int64_t adc = -(int64_t)rint(c[k] * cfg->current_2_adc);
                dst[k] += adc * mult;
            }
        }
        /* rawdata.py:398-437 add_noise: int64 += float64 truncates toward zero */
        if (cfg->enable_noise && ch < cfg->noise_nch) {
            for (int64_t k = 0; k < len; k++) {
                int64_t ix = cfg->ix_rand + k;
                if (ix >= cfg->noise_len) ix -= cfg->noise_len * (ix / cfg->noise_len);
                row[k] = (int64_t)((double)row[k] + noise[ix * cfg->noise_nch + ch]);
            }
        }
        /* rawdata.py:439-458 baseline, clamp at zero only */
        for (int64_t k = 0; k < len; k++) {
            row[k] += cfg->baseline;
            if (row[k] < 0) row[k] = 0;
        }
        /* rawdata.py:296-308 */
        int64_t found = orc_find_intervals(row, len, thr[ch], 2 * tw + 1, lr, itv_cap_local);
        for (int64_t j = 0; j < found; j++) {
            int64_t l = lr[2 * j] - tw, r = lr[2 * j + 1] + tw;
            if (l < 0) l = 0;
            if (l > len - 1) l = len - 1;
            if (r < 0) r = 0;
            if (r > len - 1) r = len - 1;
            l = (int64_t)(ceil(l / 2.0) * 2);
            r = (int64_t)(floor(r / 2.0) * 2);
            int64_t m = r - l + 1;
            if (m < 0) m = 0;
            if (n_itv < cap_itv && n_samples + m <= cap_samples) {
                itv_ch[n_itv] = (int32_t)ch;
                itv_left[n_itv] = a + l;
                itv_right[n_itv] = a + r;
                itv_off[n_itv] = n_samples;
                for (int64_t k = 0; k < m; k++) samples[n_samples + k] = (int16_t)row[l + k];
            } else {
                rc = -1;
            }
            n_itv++;
            n_samples += m;
        }
    }
    *n_itv_out = n_itv;
    *n_samples_out = n_samples;
    free(wl); free(wr); free(mask); free(row); free(lr);
    return (int)rc;
}

/* wfsim/strax_interface.py:391-436: one ZLE interval -> ceil(len/110) raw_records (244 B each).
 * rec points at packed strax.raw_record_dtype rows. */
int64_t orc_pack_records(int64_t n_itv, const int32_t *itv_ch, const int64_t *itv_left,
                         const int64_t *itv_right, const int64_t *itv_off, const int16_t *samples,
                         int64_t dt, int64_t spr, uint8_t *rec, int64_t cap_rec)
{
    const int64_t stride = 24 + 2 * spr;
    int64_t n = 0;
    for (int64_t j = 0; j < n_itv; j++) {
        int64_t pl = itv_right[j] - itv_left[j] + 1;
        int64_t need = (pl + spr - 1) / spr;
        if (pl <= 0) need = 0;
        for (int64_t i = 0; i < need; i++) {
            if (n < cap_rec) {
                uint8_t *p = rec + n * stride;
                int64_t time = dt * (itv_left[j] + spr * i);
                int32_t length = (int32_t)((pl < spr * (i + 1) ? pl : spr * (i + 1)) - spr * i);
                int16_t dt16 = (int16_t)dt, ch16 = (int16_t)itv_ch[j], ri = (int16_t)i, bl = 0;
                int32_t pl32 = (int32_t)pl;
                memcpy(p, &time, 8);
                memcpy(p + 8, &length, 4);
                memcpy(p + 12, &dt16, 2);
                memcpy(p + 14, &ch16, 2);
                memcpy(p + 16, &pl32, 4);
                memcpy(p + 20, &ri, 2);
                memcpy(p + 22, &bl, 2);
                memset(p + 24, 0, 2 * spr);
                memcpy(p + 24, samples + itv_off[j] + spr * i, 2 * (size_t)length);
            }
            n++;
        }
    }
    return n;
}

/* wfsim/core/pulse.py:82-144 (preset-gain form): photons of ONE Pulse call sorted by
 * (channel, time).  Pass 1 (currents == NULL) returns the number of pulses and the total current
 * length in out_n[0], out_n[1]; pass 2 fills p_ch/p_left/p_right/p_off and the currents. */
void orc_pulse_call(int64_t n, const int64_t *t, const int32_t *ch, const double *gain,
                    const double *gains, int64_t dt, int64_t left_margin, int64_t right_margin,
                    const double *templates, int tlen, int32_t *p_ch, int64_t *p_left,
                    int64_t *p_right, int64_t *p_off, double *currents, int64_t *out_n)
{
    int64_t np = 0, off = 0;
    int64_t a = 0;
    while (a < n) {
        int64_t b = a + 1;
        while (b < n && ch[b] == ch[a]) b++;
        int32_t c = ch[a];
        if (c >= 0 && gains[c] != 0.0) {      /* turned-off PMTs: pulse.py:89-90 */
            int64_t q0 = t[a] / dt, q1 = t[b - 1] / dt;
            if (t[a] % dt != 0 && t[a] < 0) q0--;
            if (t[b - 1] % dt != 0 && t[b - 1] < 0) q1--;
            int64_t left = q0 - left_margin, right = q1 + right_margin;
            if (currents) {
                p_ch[np] = c; p_left[np] = left; p_right[np] = right; p_off[np] = off;
                orc_add_current(t + a, gain + a, b - a, left, dt, templates, tlen, currents + off);
            }
            np++;
            off += right - left + 1;
        }
        a = b;
    }
    out_n[0] = np;
    out_n[1] = off;
}
