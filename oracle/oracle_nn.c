/* TEST INFRASTRUCTURE ONLY -- see oracle.h. Float64 restatement of the numerics the
 * reference obtains from libtorch in cmd/libtorch_bench/main.cpp:
 *   FarmerLstmModel::forward  main.cpp:25-37   (LSTM 162->128 batch_first, last step,
 *                                               cat with x, 5x Linear+ReLU, Linear->1)
 *   criterion                 main.cpp:105-114 (mse_loss / l1_loss / smooth_l1_loss, mean)
 *   train_step                main.cpp:117-135 (zero_grad, fwd, loss, backward, opt.step)
 *   make_optimizer            main.cpp:94-103  (Adam / SGD / AdamW with libtorch defaults)
 * libtorch is an un-vendored dependency (reference pins 2.7.1 CPU, Dockerfile:6,35; 2.11.0
 * is installed here). Restated semantics: LSTM gate row blocks i,f,g,o of the [4H,.] weights,
 * c' = f*c + i*g, h' = o*tanh(c'), h0=c0=0, two bias vectors; Linear y = x W^T + b;
 * Adam: m=b1 m+(1-b1)g; v=b2 v+(1-b2)g^2; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t)+eps).
 * Pinned against oracle/_ref/libfi_ref_nn.so by tests/test_oracle_pinned.py. */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define H 128
#define G4 512
#define ZD ORC_Z_DIM
#define XD ORC_X_DIM
#define FEAT (H + XD) /* 612 */
#define HID 512

/* ---------------------------------------------------------------- optimiser ---------- */
void orc_opt_update(int opt_kind, double lr, int64_t step, size_t n, double* p, const double* g,
                    double* m, double* v) {
    const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
    if (opt_kind == ORC_OPT_SGD) { /* SGDOptions(lr): momentum 0, wd 0 */
        for (size_t i = 0; i < n; i++) p[i] -= lr * g[i];
        return;
    }
    const double wd = (opt_kind == ORC_OPT_ADAMW) ? 1e-2 : 0.0; /* AdamWOptions default */
    const double bc1 = 1.0 - pow(b1, (double)step), bc2 = 1.0 - pow(b2, (double)step);
    const double step_size = lr / bc1, bc2_sqrt = sqrt(bc2);
    for (size_t i = 0; i < n; i++) {
        if (wd != 0.0) p[i] *= (1.0 - lr * wd); /* decoupled decay (AdamW) */
        m[i] = b1 * m[i] + (1.0 - b1) * g[i];
        v[i] = b2 * v[i] + (1.0 - b2) * g[i] * g[i];
        double denom = sqrt(v[i]) / bc2_sqrt + eps;
        p[i] -= step_size * (m[i] / denom);
    }
}

void orc_opt_update_f32(int opt_kind, double lr, int64_t step, size_t n, float* p, const float* g,
                        float* m, float* v) {
    const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
    if (opt_kind == ORC_OPT_SGD) {
        for (size_t i = 0; i < n; i++) p[i] = p[i] - (float)lr * g[i];
        return;
    }
    const double wd = (opt_kind == ORC_OPT_ADAMW) ? 1e-2 : 0.0;
    const double bc1 = 1.0 - pow(b1, (double)step), bc2 = 1.0 - pow(b2, (double)step);
    const float step_size = (float)(lr / bc1), bc2_sqrt = (float)sqrt(bc2);
    const float fb1 = (float)b1, fb2 = (float)b2, omb1 = (float)(1.0 - b1), omb2 = (float)(1.0 - b2);
    for (size_t i = 0; i < n; i++) {
        if (wd != 0.0) p[i] = p[i] * (float)(1.0 - lr * wd);
        m[i] = fb1 * m[i] + omb1 * g[i];
        v[i] = fb2 * v[i] + omb2 * g[i] * g[i];
        float denom = sqrtf(v[i]) / bc2_sqrt + (float)eps;
        p[i] = p[i] - step_size * (m[i] / denom);
    }
}

/* ---------------------------------------------------------------- dense helpers ------ */
/* Y[m,n] = X[m,k](ldx) W[n,k]^T + b, optional ReLU. */
static void dense_fwd(const double* X, int ldx, const double* W, const double* b, int m, int n,
                      int k, double* Y, int relu) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < m; i++) {
        const double* xi = X + (size_t)i * ldx;
        for (int j = 0; j < n; j++) {
            const double* wj = W + (size_t)j * k;
            double acc = b[j];
            for (int q = 0; q < k; q++) acc += xi[q] * wj[q];
            Y[(size_t)i * n + j] = (relu && acc < 0.0) ? 0.0 : acc;
        }
    }
}
/* dW[n,k] += dY^T X ; db[n] += colsum(dY) */
static void dense_wgrad(const double* dY, const double* X, int ldx, int m, int n, int k, double* dW,
                        double* db) {
#pragma omp parallel for schedule(static)
    for (int j = 0; j < n; j++) {
        double* dwj = dW + (size_t)j * k;
        double sb = 0.0;
        for (int i = 0; i < m; i++) {
            double d = dY[(size_t)i * n + j];
            if (d == 0.0) continue;
            sb += d;
            const double* xi = X + (size_t)i * ldx;
            for (int q = 0; q < k; q++) dwj[q] += d * xi[q];
        }
        db[j] += sb;
    }
}
/* dX[m,k] = dY[m,n] W[n,k]; if act != NULL, multiply by relu'(act) (act is the ReLU output
 * that produced X, so derivative is act>0). */
static void dense_dgrad(const double* dY, const double* W, int m, int n, int k, double* dX,
                        const double* act) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < m; i++) {
        double* dxi = dX + (size_t)i * k;
        for (int q = 0; q < k; q++) dxi[q] = 0.0;
        for (int j = 0; j < n; j++) {
            double d = dY[(size_t)i * n + j];
            if (d == 0.0) continue;
            const double* wj = W + (size_t)j * k;
            for (int q = 0; q < k; q++) dxi[q] += d * wj[q];
        }
        if (act)
            for (int q = 0; q < k; q++)
                if (!(act[(size_t)i * k + q] > 0.0)) dxi[q] = 0.0;
    }
}

static double sigmoid(double x) { return 1.0 / (1.0 + exp(-x)); }

/* ---------------------------------------------------------------- FarmerLstmModel ---- */
/* model.parameters() order (main.cpp:16-22): lstm.weight_ih_l0 [512,162], weight_hh_l0
 * [512,128], bias_ih_l0 [512], bias_hh_l0 [512], dense1.weight [512,612], dense1.bias [512],
 * dense2..5 weight [512,512] + bias [512], dense6.weight [1,512], dense6.bias [1]. */
static const int64_t FARMER_NUMEL[ORC_FARMER_TENSORS] = {
    G4 * ZD, G4 * H, G4, G4, HID * FEAT, HID, HID * HID, HID, HID * HID, HID,
    HID * HID, HID, HID * HID, HID, HID, 1};

void orc_farmer_tensor_table(int64_t offsets[ORC_FARMER_TENSORS], int64_t numels[ORC_FARMER_TENSORS]) {
    int64_t o = 0;
    for (int i = 0; i < ORC_FARMER_TENSORS; i++) {
        offsets[i] = o;
        numels[i] = FARMER_NUMEL[i];
        o += FARMER_NUMEL[i];
    }
}

struct orc_farmer {
    double *p, *g, *m, *v;
    int64_t off[ORC_FARMER_TENSORS], numel[ORC_FARMER_TENSORS];
    int opt_kind, loss_kind;
    double lr;
    int64_t step;
};

orc_farmer* orc_farmer_create(const float* params, int opt_kind, double lr, int loss_kind) {
    orc_farmer* f = (orc_farmer*)calloc(1, sizeof(*f));
    orc_farmer_tensor_table(f->off, f->numel);
    size_t n = ORC_FARMER_PARAMS;
    f->p = (double*)malloc(n * sizeof(double));
    f->g = (double*)calloc(n, sizeof(double));
    f->m = (double*)calloc(n, sizeof(double));
    f->v = (double*)calloc(n, sizeof(double));
    for (size_t i = 0; i < n; i++) f->p[i] = params[i];
    f->opt_kind = opt_kind;
    f->loss_kind = loss_kind;
    f->lr = lr;
    f->step = 0;
    return f;
}
void orc_farmer_destroy(orc_farmer* f) {
    if (!f) return;
    free(f->p); free(f->g); free(f->m); free(f->v); free(f);
}
void orc_farmer_get_params(const orc_farmer* f, double* out) { memcpy(out, f->p, sizeof(double) * ORC_FARMER_PARAMS); }
void orc_farmer_get_grads(const orc_farmer* f, double* out) { memcpy(out, f->g, sizeof(double) * ORC_FARMER_PARAMS); }
void orc_farmer_set_grads(orc_farmer* f, const double* in) { memcpy(f->g, in, sizeof(double) * ORC_FARMER_PARAMS); }
void orc_farmer_opt_step(orc_farmer* f) {
    f->step++;
    orc_opt_update(f->opt_kind, f->lr, f->step, ORC_FARMER_PARAMS, f->p, f->g, f->m, f->v);
}

typedef struct {
    int b, t;
    double *gates; /* [b,t,4H] post-activation i,f,g,o */
    double *c;     /* [b,t,H] */
    double *h;     /* [b,t,H] */
    double *feat;  /* [b,612] */
    double *act[5];/* [b,512] ReLU outputs of dense1..5 */
    double *y;     /* [b] */
} farmer_ws;

static void ws_alloc(farmer_ws* w, int b, int t) {
    w->b = b; w->t = t;
    w->gates = (double*)malloc(sizeof(double) * (size_t)b * t * G4);
    w->c = (double*)malloc(sizeof(double) * (size_t)b * t * H);
    w->h = (double*)malloc(sizeof(double) * (size_t)b * t * H);
    w->feat = (double*)malloc(sizeof(double) * (size_t)b * FEAT);
    for (int l = 0; l < 5; l++) w->act[l] = (double*)malloc(sizeof(double) * (size_t)b * HID);
    w->y = (double*)malloc(sizeof(double) * (size_t)b);
}
static void ws_free(farmer_ws* w) {
    free(w->gates); free(w->c); free(w->h); free(w->feat);
    for (int l = 0; l < 5; l++) free(w->act[l]);
    free(w->y);
}

static void farmer_fwd(const orc_farmer* f, const float* z, const float* x, farmer_ws* w) {
    const int b = w->b, t = w->t;
    const double* Wih = f->p + f->off[0];
    const double* Whh = f->p + f->off[1];
    const double* bih = f->p + f->off[2];
    const double* bhh = f->p + f->off[3];
#pragma omp parallel for schedule(static)
    for (int n = 0; n < b; n++) {
        double hprev[H], cprev[H], pre[G4];
        for (int j = 0; j < H; j++) hprev[j] = cprev[j] = 0.0;
        for (int s = 0; s < t; s++) {
            const float* zt = z + ((size_t)n * t + s) * ZD;
            for (int r = 0; r < G4; r++) {
                double acc = bih[r] + bhh[r];
                const double* wi = Wih + (size_t)r * ZD;
                for (int q = 0; q < ZD; q++) acc += wi[q] * (double)zt[q];
                const double* wh = Whh + (size_t)r * H;
                for (int q = 0; q < H; q++) acc += wh[q] * hprev[q];
                pre[r] = acc;
            }
            double* gt = w->gates + ((size_t)n * t + s) * G4;
            double* ct = w->c + ((size_t)n * t + s) * H;
            double* ht = w->h + ((size_t)n * t + s) * H;
            for (int j = 0; j < H; j++) {
                double ig = sigmoid(pre[j]), fg = sigmoid(pre[H + j]);
                double gg = tanh(pre[2 * H + j]), og = sigmoid(pre[3 * H + j]);
                gt[j] = ig; gt[H + j] = fg; gt[2 * H + j] = gg; gt[3 * H + j] = og;
                ct[j] = fg * cprev[j] + ig * gg;
                ht[j] = og * tanh(ct[j]);
            }
            for (int j = 0; j < H; j++) { hprev[j] = ht[j]; cprev[j] = ct[j]; }
        }
        double* ft = w->feat + (size_t)n * FEAT; /* cat({last_timestep, x}), main.cpp:30 */
        for (int j = 0; j < H; j++) ft[j] = hprev[j];
        for (int j = 0; j < XD; j++) ft[H + j] = (double)x[(size_t)n * XD + j];
    }
    const double* in = w->feat;
    int k = FEAT;
    for (int l = 0; l < 5; l++) {
        dense_fwd(in, k, f->p + f->off[4 + 2 * l], f->p + f->off[5 + 2 * l], b, HID, k, w->act[l], 1);
        in = w->act[l];
        k = HID;
    }
    dense_fwd(in, HID, f->p + f->off[14], f->p + f->off[15], b, 1, HID, w->y, 0);
}

void orc_farmer_forward(orc_farmer* f, const float* z, const float* x, int b, int t, double* y) {
    farmer_ws w;
    ws_alloc(&w, b, t);
    farmer_fwd(f, z, x, &w);
    memcpy(y, w.y, sizeof(double) * b);
    ws_free(&w);
}

/* per-sample loss and d(loss_i)/dy (before the 1/denominator of the mean) */
static double loss_elem(int kind, double y, double tgt, double* dy) {
    double d = y - tgt;
    switch (kind) {
        case ORC_LOSS_MAE:
            *dy = (d > 0) - (d < 0);
            return fabs(d);
        case ORC_LOSS_HUBER: /* smooth_l1_loss, beta = 1 */
            if (fabs(d) < 1.0) { *dy = d; return 0.5 * d * d; }
            *dy = (d > 0) - (d < 0);
            return fabs(d) - 0.5;
        default:
            *dy = 2.0 * d;
            return d * d;
    }
}

double orc_farmer_loss_grad(orc_farmer* f, const float* z, const float* x, const float* target,
                            int b, int t, int loss_denom) {
    farmer_ws w;
    ws_alloc(&w, b, t);
    farmer_fwd(f, z, x, &w);
    memset(f->g, 0, sizeof(double) * ORC_FARMER_PARAMS);

    double loss = 0.0;
    double* dy = (double*)malloc(sizeof(double) * b);
    for (int n = 0; n < b; n++) {
        double d;
        loss += loss_elem(f->loss_kind, w.y[n], (double)target[n], &d);
        dy[n] = d / (double)loss_denom;
    }
    loss /= (double)loss_denom;

    /* dense6 .. dense1 */
    double* da = (double*)malloc(sizeof(double) * (size_t)b * FEAT);
    double* db_ = (double*)malloc(sizeof(double) * (size_t)b * FEAT);
    dense_wgrad(dy, w.act[4], HID, b, 1, HID, f->g + f->off[14], f->g + f->off[15]);
    dense_dgrad(dy, f->p + f->off[14], b, 1, HID, da, w.act[4]);
    for (int l = 4; l >= 0; l--) {
        const double* in = l == 0 ? w.feat : w.act[l - 1];
        int k = l == 0 ? FEAT : HID;
        dense_wgrad(da, in, k, b, HID, k, f->g + f->off[4 + 2 * l], f->g + f->off[5 + 2 * l]);
        dense_dgrad(da, f->p + f->off[4 + 2 * l], b, HID, k, db_, l == 0 ? NULL : w.act[l - 1]);
        double* tmp = da; da = db_; db_ = tmp;
    }
    /* da now holds dfeat [b,612]; first 128 columns are dh_{T-1} */

    /* BPTT: dG[b,t,4H] pre-activation gate gradients */
    double* dG = (double*)malloc(sizeof(double) * (size_t)b * t * G4);
    const double* Whh = f->p + f->off[1];
#pragma omp parallel for schedule(static)
    for (int n = 0; n < b; n++) {
        double dh[H], dc[H], dhn[H];
        for (int j = 0; j < H; j++) { dh[j] = da[(size_t)n * FEAT + j]; dc[j] = 0.0; }
        for (int s = t - 1; s >= 0; s--) {
            const double* gt = w.gates + ((size_t)n * t + s) * G4;
            const double* ct = w.c + ((size_t)n * t + s) * H;
            const double* cp = s > 0 ? w.c + ((size_t)n * t + s - 1) * H : NULL;
            double* dg = dG + ((size_t)n * t + s) * G4;
            for (int j = 0; j < H; j++) {
                double ig = gt[j], fg = gt[H + j], gg = gt[2 * H + j], og = gt[3 * H + j];
                double tc = tanh(ct[j]);
                double dct = dc[j] + dh[j] * og * (1.0 - tc * tc);
                double d_o = dh[j] * tc;
                dg[j] = dct * gg * ig * (1.0 - ig);
                dg[H + j] = dct * (cp ? cp[j] : 0.0) * fg * (1.0 - fg);
                dg[2 * H + j] = dct * ig * (1.0 - gg * gg);
                dg[3 * H + j] = d_o * og * (1.0 - og);
                dc[j] = dct * fg;
            }
            for (int j = 0; j < H; j++) dhn[j] = 0.0;
            for (int r = 0; r < G4; r++) {
                const double* wh = Whh + (size_t)r * H;
                double d = dg[r];
                for (int j = 0; j < H; j++) dhn[j] += d * wh[j];
            }
            for (int j = 0; j < H; j++) dh[j] = dhn[j];
        }
    }
    double* gWih = f->g + f->off[0];
    double* gWhh = f->g + f->off[1];
    double* gbih = f->g + f->off[2];
    double* gbhh = f->g + f->off[3];
#pragma omp parallel for schedule(static)
    for (int r = 0; r < G4; r++) {
        double sb = 0.0;
        double* gi = gWih + (size_t)r * ZD;
        double* gh = gWhh + (size_t)r * H;
        for (int n = 0; n < b; n++)
            for (int s = 0; s < t; s++) {
                double d = dG[((size_t)n * t + s) * G4 + r];
                sb += d;
                const float* zt = z + ((size_t)n * t + s) * ZD;
                for (int q = 0; q < ZD; q++) gi[q] += d * (double)zt[q];
                if (s > 0) {
                    const double* hp = w.h + ((size_t)n * t + s - 1) * H;
                    for (int q = 0; q < H; q++) gh[q] += d * hp[q];
                }
            }
        gbih[r] += sb;
        gbhh[r] += sb;
    }
    free(dG); free(da); free(db_); free(dy);
    ws_free(&w);
    return loss;
}

double orc_farmer_train_step(orc_farmer* f, const float* z, const float* x, const float* target,
                             int b, int t) {
    double loss = orc_farmer_loss_grad(f, z, x, target, b, t, b);
    orc_farmer_opt_step(f);
    return loss;
}

/* ---------------------------------------------------------------- MLP actor-critic --- */
/* This build's V-trace model (not in the reference): the reference trunk shapes
 * (main.cpp:17-21) applied per transition to the 162-feature observation, and one fused
 * head [A+1, 512]: rows 0..A-1 policy logits, row A the value. Parameter order:
 * dense1.w [512,162], dense1.b, dense2..5 .w [512,512] .b, head.w [17,512], head.b [17]. */
#define NA ORC_NUM_ACTIONS
#define NHEAD (NA + 1)
static const int64_t AC_NUMEL[ORC_AC_TENSORS] = {
    HID * ZD, HID, HID * HID, HID, HID * HID, HID, HID * HID, HID, HID * HID, HID, NHEAD * HID, NHEAD};

void orc_ac_tensor_table(int64_t offsets[ORC_AC_TENSORS], int64_t numels[ORC_AC_TENSORS]) {
    int64_t o = 0;
    for (int i = 0; i < ORC_AC_TENSORS; i++) {
        offsets[i] = o;
        numels[i] = AC_NUMEL[i];
        o += AC_NUMEL[i];
    }
}

struct orc_ac {
    double *p, *g, *m, *v;
    int64_t off[ORC_AC_TENSORS], numel[ORC_AC_TENSORS];
    int opt_kind;
    double lr;
    int64_t step;
    orc_vtrace_cfg cfg;
};

orc_ac* orc_ac_create(const float* params, int opt_kind, double lr, const orc_vtrace_cfg* cfg) {
    orc_ac* f = (orc_ac*)calloc(1, sizeof(*f));
    orc_ac_tensor_table(f->off, f->numel);
    size_t n = ORC_AC_PARAMS;
    f->p = (double*)malloc(n * sizeof(double));
    f->g = (double*)calloc(n, sizeof(double));
    f->m = (double*)calloc(n, sizeof(double));
    f->v = (double*)calloc(n, sizeof(double));
    for (size_t i = 0; i < n; i++) f->p[i] = params[i];
    f->opt_kind = opt_kind;
    f->lr = lr;
    f->cfg = *cfg;
    return f;
}
void orc_ac_destroy(orc_ac* f) {
    if (!f) return;
    free(f->p); free(f->g); free(f->m); free(f->v); free(f);
}
void orc_ac_get_params(const orc_ac* f, double* out) { memcpy(out, f->p, sizeof(double) * ORC_AC_PARAMS); }
void orc_ac_get_grads(const orc_ac* f, double* out) { memcpy(out, f->g, sizeof(double) * ORC_AC_PARAMS); }
void orc_ac_set_grads(orc_ac* f, const double* in) { memcpy(f->g, in, sizeof(double) * ORC_AC_PARAMS); }
void orc_ac_opt_step(orc_ac* f) {
    f->step++;
    orc_opt_update(f->opt_kind, f->lr, f->step, ORC_AC_PARAMS, f->p, f->g, f->m, f->v);
}

/* Forward. gate[l] (may be NULL) receives the ReLU decisions as 1.0 / 0.0. With `mask` ([5][rows*HID]
 * bytes, may be NULL) the decisions are taken from the caller instead of z > 0 -- used by the GPU parity
 * tests to pin the units whose pre-activation lies within fp32 rounding of the kink, where an fp32 and
 * a float64 forward legitimately disagree; *n_override counts the units where the mask differs from
 * z > 0 and *max_override is the largest |z| / rms(z of that layer) among them. */
static void ac_fwd(const orc_ac* f, const float* obs, int rows, double* in0, double* act[5],
                   double* head, double* gate[5], const uint8_t* mask, int64_t* n_override,
                   double* max_override) {
    for (size_t i = 0; i < (size_t)rows * ZD; i++) in0[i] = (double)obs[i];
    const double* in = in0;
    int k = ZD;
    if (n_override) *n_override = 0;
    if (max_override) *max_override = 0.0;
    for (int l = 0; l < 5; l++) {
        const size_t n = (size_t)rows * HID;
        dense_fwd(in, k, f->p + f->off[2 * l], f->p + f->off[2 * l + 1], rows, HID, k, act[l], 0);
        double ss = 0.0;
        for (size_t i = 0; i < n; i++) ss += act[l][i] * act[l][i];
        const double rms = sqrt(ss / (double)n) + 1e-300;
        for (size_t i = 0; i < n; i++) {
            const double z = act[l][i];
            int on = z > 0.0;
            if (mask) {
                const int want = mask[(size_t)l * n + i] != 0;
                if (want != on) {
                    if (n_override) (*n_override)++;
                    if (max_override && fabs(z) / rms > *max_override) *max_override = fabs(z) / rms;
                    on = want;
                }
            }
            act[l][i] = on ? z : 0.0;
            if (gate) gate[l][i] = on ? 1.0 : 0.0;
        }
        in = act[l];
        k = HID;
    }
    dense_fwd(in, HID, f->p + f->off[10], f->p + f->off[11], rows, NHEAD, HID, head, 0);
}

void orc_ac_forward(orc_ac* f, const float* obs, int rows, double* logits, double* value) {
    double* in0 = (double*)malloc(sizeof(double) * (size_t)rows * ZD);
    double* act[5];
    for (int l = 0; l < 5; l++) act[l] = (double*)malloc(sizeof(double) * (size_t)rows * HID);
    double* head = (double*)malloc(sizeof(double) * (size_t)rows * NHEAD);
    ac_fwd(f, obs, rows, in0, act, head, NULL, NULL, NULL, NULL);
    for (int i = 0; i < rows; i++) {
        for (int a = 0; a < NA; a++) logits[(size_t)i * NA + a] = head[(size_t)i * NHEAD + a];
        value[i] = head[(size_t)i * NHEAD + NA];
    }
    free(in0); free(head);
    for (int l = 0; l < 5; l++) free(act[l]);
}

void orc_ac_loss_grad_masked(orc_ac* f, const float* obs, const float* mu_logits, const int32_t* action,
                             const float* reward, const float* discount, const float* bootstrap, int m,
                             int t, double* out_losses, const uint8_t* relu_mask, int64_t* n_override,
                             double* max_override) {
    const int rows = m * t;
    double* in0 = (double*)malloc(sizeof(double) * (size_t)rows * ZD);
    double *act[5], *gate[5];
    for (int l = 0; l < 5; l++) {
        act[l] = (double*)malloc(sizeof(double) * (size_t)rows * HID);
        gate[l] = (double*)malloc(sizeof(double) * (size_t)rows * HID);
    }
    double* head = (double*)malloc(sizeof(double) * (size_t)rows * NHEAD);
    ac_fwd(f, obs, rows, in0, act, head, gate, relu_mask, n_override, max_override);

    double* logits = (double*)malloc(sizeof(double) * (size_t)rows * NA);
    double* value = (double*)malloc(sizeof(double) * (size_t)rows);
    for (int i = 0; i < rows; i++) {
        for (int a = 0; a < NA; a++) logits[(size_t)i * NA + a] = head[(size_t)i * NHEAD + a];
        value[i] = head[(size_t)i * NHEAD + NA];
    }
    double* dlogits = (double*)malloc(sizeof(double) * (size_t)rows * NA);
    double* dvalue = (double*)malloc(sizeof(double) * (size_t)rows);
    orc_vtrace_losses(m, t, NA, logits, value, mu_logits, action, reward, discount, bootstrap,
                      &f->cfg, out_losses, dlogits, dvalue, NULL, NULL);
    double* dhead = head; /* reuse */
    for (int i = 0; i < rows; i++) {
        for (int a = 0; a < NA; a++) dhead[(size_t)i * NHEAD + a] = dlogits[(size_t)i * NA + a];
        dhead[(size_t)i * NHEAD + NA] = dvalue[i];
    }
    memset(f->g, 0, sizeof(double) * ORC_AC_PARAMS);
    double* da = (double*)malloc(sizeof(double) * (size_t)rows * HID);
    double* db_ = (double*)malloc(sizeof(double) * (size_t)rows * HID);
    dense_wgrad(dhead, act[4], HID, rows, NHEAD, HID, f->g + f->off[10], f->g + f->off[11]);
    dense_dgrad(dhead, f->p + f->off[10], rows, NHEAD, HID, da, gate[4]);
    for (int l = 4; l >= 0; l--) {
        const double* in = l == 0 ? in0 : act[l - 1];
        int k = l == 0 ? ZD : HID;
        dense_wgrad(da, in, k, rows, HID, k, f->g + f->off[2 * l], f->g + f->off[2 * l + 1]);
        if (l > 0) {
            dense_dgrad(da, f->p + f->off[2 * l], rows, HID, HID, db_, gate[l - 1]);
            double* tmp = da; da = db_; db_ = tmp;
        }
    }
    free(in0); free(head); free(logits); free(value); free(dlogits); free(dvalue); free(da); free(db_);
    for (int l = 0; l < 5; l++) { free(act[l]); free(gate[l]); }
}

void orc_ac_loss_grad(orc_ac* f, const float* obs, const float* mu_logits, const int32_t* action,
                      const float* reward, const float* discount, const float* bootstrap, int m,
                      int t, double* out_losses) {
    orc_ac_loss_grad_masked(f, obs, mu_logits, action, reward, discount, bootstrap, m, t, out_losses,
                            NULL, NULL, NULL);
}

/* Threads the OpenMP loops above use (bench.py reports it as cpu_baseline.cores). */
#ifdef _OPENMP
#include <omp.h>
int orc_num_threads(void) { return omp_get_max_threads(); }
#else
int orc_num_threads(void) { return 1; }
#endif
