/* TEST INFRASTRUCTURE ONLY -- see oracle.h. Float64 host implementation of V-trace
 * (Espeholt et al. 2018, arXiv:1802.01561). The reference has NO V-trace, policy head or
 * entropy loss, so parity against the reference is UNPINNED for this
 * file; BASELINE.json's north_star names exactly this float64 host recurrence as the checker.
 *
 * Paper eq. (1) with the lambda extension of Remark 2, trajectory b, time s in [0,T):
 *   rho_s = min(rho_bar, exp(log_rho_s))        c_s = lambda * min(c_bar, exp(log_rho_s))
 *   delta_s = rho_s (r_s + gamma_s V_{s+1} - V_s),   V_T = bootstrap
 *   vs_s - V_s = delta_s + gamma_s c_s (vs_{s+1} - V_{s+1}),   vs_T - V_T = 0
 * Section 4.2 policy-gradient advantage:
 *   pg_adv_s = min(pg_rho_bar, exp(log_rho_s)) (r_s + gamma_s vs_{s+1} - V_s),  vs_T = bootstrap
 * Losses (section 4.2, sums over time and batch; vs and pg_adv treated as constants):
 *   pg = - sum log pi(a_s|x_s) pg_adv_s     baseline = 1/2 sum (vs_s - V_s)^2
 *   entropy = sum_s sum_a pi log pi         total = pg + c_v baseline + c_e entropy */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>

void orc_vtrace(int m, int t, const double* log_rho, const double* discount, const double* reward,
                const double* value, const double* bootstrap, double rho_bar, double c_bar,
                double pg_rho_bar, double lambda, double* vs, double* pg_adv) {
    for (int b = 0; b < m; b++) {
        const size_t o = (size_t)b * t;
        double acc = 0.0;
        for (int s = t - 1; s >= 0; s--) {
            double is = exp(log_rho[o + s]);
            double rho = is < rho_bar ? is : rho_bar;
            double c = lambda * (is < c_bar ? is : c_bar);
            double v_next = s == t - 1 ? bootstrap[b] : value[o + s + 1];
            double delta = rho * (reward[o + s] + discount[o + s] * v_next - value[o + s]);
            acc = delta + discount[o + s] * c * acc;
            vs[o + s] = value[o + s] + acc;
        }
        if (pg_adv)
            for (int s = 0; s < t; s++) {
                double is = exp(log_rho[o + s]);
                double rho_pg = is < pg_rho_bar ? is : pg_rho_bar;
                double vs_next = s == t - 1 ? bootstrap[b] : vs[o + s + 1];
                pg_adv[o + s] = rho_pg * (reward[o + s] + discount[o + s] * vs_next - value[o + s]);
            }
    }
}

void orc_vtrace_closed_form(int m, int t, const double* log_rho, const double* discount,
                            const double* reward, const double* value, const double* bootstrap,
                            double rho_bar, double c_bar, double lambda, double* vs) {
    for (int b = 0; b < m; b++) {
        const size_t o = (size_t)b * t;
        for (int s = 0; s < t; s++) {
            double sum = 0.0, coef = 1.0; /* coef = prod_{i=s}^{k-1} gamma_i c_i */
            for (int k = s; k < t; k++) {
                double is = exp(log_rho[o + k]);
                double rho = is < rho_bar ? is : rho_bar;
                double c = lambda * (is < c_bar ? is : c_bar);
                double v_next = k == t - 1 ? bootstrap[b] : value[o + k + 1];
                sum += coef * rho * (reward[o + k] + discount[o + k] * v_next - value[o + k]);
                coef *= discount[o + k] * c;
            }
            vs[o + s] = value[o + s] + sum;
        }
    }
}

static double log_softmax_row(const double* z, int a, double* logp) {
    double mx = z[0];
    for (int i = 1; i < a; i++) mx = z[i] > mx ? z[i] : mx;
    double se = 0.0;
    for (int i = 0; i < a; i++) se += exp(z[i] - mx);
    double lse = mx + log(se);
    for (int i = 0; i < a; i++) logp[i] = z[i] - lse;
    return lse;
}

void orc_vtrace_losses(int m, int t, int a, const double* logits, const double* value,
                       const float* mu_logits, const int32_t* action, const float* reward,
                       const float* discount, const float* bootstrap, const orc_vtrace_cfg* cfg,
                       double* out_losses, double* dlogits, double* dvalue, double* vs_out,
                       double* pg_adv_out) {
    const size_t n = (size_t)m * t;
    double* log_rho = (double*)malloc(sizeof(double) * n);
    double* disc = (double*)malloc(sizeof(double) * n);
    double* rew = (double*)malloc(sizeof(double) * n);
    double* boot = (double*)malloc(sizeof(double) * m);
    double* vs = (double*)malloc(sizeof(double) * n);
    double* adv = (double*)malloc(sizeof(double) * n);
    double* logp = (double*)malloc(sizeof(double) * n * a);
    double* tmp = (double*)malloc(sizeof(double) * a);
    double* mu = (double*)malloc(sizeof(double) * a);
    for (size_t i = 0; i < n; i++) {
        log_softmax_row(logits + i * a, a, logp + i * a);
        for (int j = 0; j < a; j++) mu[j] = (double)mu_logits[i * a + j];
        log_softmax_row(mu, a, tmp);
        log_rho[i] = logp[i * a + action[i]] - tmp[action[i]];
        disc[i] = (double)discount[i];
        rew[i] = (double)reward[i];
    }
    for (int b = 0; b < m; b++) boot[b] = (double)bootstrap[b];
    orc_vtrace(m, t, log_rho, disc, rew, value, boot, cfg->rho_bar, cfg->c_bar, cfg->pg_rho_bar,
               cfg->lambda, vs, adv);

    double pg = 0.0, bl = 0.0, ent = 0.0;
    for (size_t i = 0; i < n; i++) {
        const double* lp = logp + i * a;
        double plogp = 0.0;
        for (int j = 0; j < a; j++) plogp += exp(lp[j]) * lp[j];
        pg += -lp[action[i]] * adv[i];
        bl += 0.5 * (vs[i] - value[i]) * (vs[i] - value[i]);
        ent += plogp;
        if (dvalue) dvalue[i] = -cfg->baseline_cost * (vs[i] - value[i]);
        if (dlogits)
            for (int j = 0; j < a; j++) {
                double pi = exp(lp[j]);
                double d_pg = adv[i] * (pi - (j == action[i] ? 1.0 : 0.0));
                double d_ent = pi * (lp[j] - plogp);
                dlogits[i * a + j] = d_pg + cfg->entropy_cost * d_ent;
            }
    }
    if (out_losses) {
        out_losses[1] = pg;
        out_losses[2] = bl;
        out_losses[3] = ent;
        out_losses[0] = pg + cfg->baseline_cost * bl + cfg->entropy_cost * ent;
    }
    if (vs_out) for (size_t i = 0; i < n; i++) vs_out[i] = vs[i];
    if (pg_adv_out) for (size_t i = 0; i < n; i++) pg_adv_out[i] = adv[i];
    free(log_rho); free(disc); free(rew); free(boot); free(vs); free(adv); free(logp); free(tmp); free(mu);
}
