/* TEST INFRASTRUCTURE ONLY. CPU restatement (plain C, float64 arithmetic) of the reference's
 * learner hot path, used solely as the checker by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg. The product (freeimpala_b200/) never links or calls this.
 *
 * Pinning: the reference holds NO tests, golden vectors or fixtures
 * (find -iname '*test*' is empty). The oracle is therefore pinned against outputs of the
 * reference itself, compiled here from its own sources into oracle/_ref (oracle/Makefile):
 *   - ring semantics      vs SharedBuffer            (tests/test_oracle_pinned.py)
 *   - FarmerLstm step     vs libtorch train_step     (same file; fixtures in tests/golden/)
 * V-trace and the policy/value/entropy losses do not exist in the reference at all
 * (SURVEY.md section 0): for that part parity against the reference is UNPINNED; the oracle
 * restates Espeholt et al. 2018 (arXiv:1802.01561) eq. (1) and section 4.1-4.2 in float64
 * and is cross-checked against an independent closed-form (non-recursive) evaluation.
 */
#ifndef FI_ORACLE_H
#define FI_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------- trajectory ring: SharedBuffer, data_structures.h:191-307 ---------- */
#define ORC_ELEMENT_SIZE 1024 /* data_structures.h:35 */
typedef struct orc_ring orc_ring;
orc_ring* orc_ring_create(size_t entry_size, size_t capacity);      /* ctor, :205-210 */
void orc_ring_destroy(orc_ring* r);
/* write, :219-241. 1 = written; 0 = data larger than the slot (nothing changes);
 * -1 = the reference would block here (ring full). Bytes [n, slot) keep old content. */
int orc_ring_write(orc_ring* r, const void* data, size_t n);
/* readBatch, :267-300. Returns M and fills out[M*slot] FIFO with wraparound; 0 = empty
 * batch (draining and count<M, :278-280); -1 = the reference would block. */
long orc_ring_read_batch(orc_ring* r, size_t m, void* out);
void orc_ring_set_draining(orc_ring* r);                            /* :212-216 */
size_t orc_ring_filled_count(const orc_ring* r);                    /* :303-306 */
size_t orc_ring_slot_bytes(const orc_ring* r);

/* ---------------- trajectory record layout (defined by this build) ---- */
#define ORC_REC_WORDS 256
#define ORC_Z_DIM 162
#define ORC_X_DIM 484
#define ORC_NUM_ACTIONS 16
#define ORC_W_MU 162      /* behaviour logits [16] */
#define ORC_W_ACTION 178  /* int32 */
#define ORC_W_REWARD 179
#define ORC_W_DISCOUNT 180
#define ORC_W_AUX 181     /* record 0: regression target; record S-1: bootstrap value */
#define ORC_W_X 192       /* 64 words of x per record, records 0..7 */
#define ORC_X_PER_REC 64
/* batch: m slots of s records. Outputs are [m,s,162], [m,484], [m]. */
void orc_decode_farmer(const void* batch, size_t m, size_t s, float* z, float* x, float* target);
/* Outputs trajectory-major: obs [m,s,162], mu_logits [m,s,16], action [m,s], reward [m,s],
 * discount [m,s], bootstrap [m]. */
void orc_decode_vtrace(const void* batch, size_t m, size_t s, float* obs, float* mu_logits,
                       int32_t* action, float* reward, float* discount, float* bootstrap);

/* ---------------- optimiser (torch::optim, main.cpp:94-103) -------------------------- */
enum { ORC_OPT_ADAM = 0, ORC_OPT_SGD = 1, ORC_OPT_ADAMW = 2 };
enum { ORC_LOSS_MSE = 0, ORC_LOSS_MAE = 1, ORC_LOSS_HUBER = 2 };
/* One libtorch-default Adam/AdamW/SGD update on n float64 values. step counts from 1. */
void orc_opt_update(int opt_kind, double lr, int64_t step, size_t n, double* p, const double* g,
                    double* m, double* v);
/* Same arithmetic carried in float32 storage (p,g,m,v float) with float64 bias corrections,
 * i.e. what torch::optim::Adam does elementwise; used to bound the fused Adam kernel. */
void orc_opt_update_f32(int opt_kind, double lr, int64_t step, size_t n, float* p, const float* g,
                        float* m, float* v);

/* ---------------- FarmerLstmModel step: libtorch_bench main.cpp:14-42,105-135 -------- */
#define ORC_FARMER_PARAMS 1514497
#define ORC_FARMER_TENSORS 16
typedef struct orc_farmer orc_farmer;
orc_farmer* orc_farmer_create(const float* params, int opt_kind, double lr, int loss_kind);
void orc_farmer_destroy(orc_farmer* f);
void orc_farmer_tensor_table(int64_t offsets[ORC_FARMER_TENSORS], int64_t numels[ORC_FARMER_TENSORS]);
void orc_farmer_forward(orc_farmer* f, const float* z, const float* x, int b, int t, double* y);
/* forward + criterion + backward; gradients are zeroed first (zero_grad, main.cpp:124).
 * loss_denom = number of samples the mean is taken over (b for one process; the global
 * batch under data parallelism). Returns sum-of-per-sample-loss / loss_denom. */
double orc_farmer_loss_grad(orc_farmer* f, const float* z, const float* x, const float* target,
                            int b, int t, int loss_denom);
void orc_farmer_opt_step(orc_farmer* f);
double orc_farmer_train_step(orc_farmer* f, const float* z, const float* x, const float* target,
                             int b, int t);
void orc_farmer_get_params(const orc_farmer* f, double* out);
void orc_farmer_get_grads(const orc_farmer* f, double* out);
void orc_farmer_set_grads(orc_farmer* f, const double* in);

/* ---------------- V-trace (Espeholt et al. 2018); NOT in the reference ---------------- */
/* All arrays trajectory-major [m,t]. vs and pg_adv are outputs. */
void orc_vtrace(int m, int t, const double* log_rho, const double* discount, const double* reward,
                const double* value, const double* bootstrap, double rho_bar, double c_bar,
                double pg_rho_bar, double lambda, double* vs, double* pg_adv);
/* Non-recursive evaluation of paper eq. (1): vs_s = V_s + sum_{k>=s} gamma-products * (prod c) * delta_k.
 * O(T^2); independent of the scan formulation. */
void orc_vtrace_closed_form(int m, int t, const double* log_rho, const double* discount,
                            const double* reward, const double* value, const double* bootstrap,
                            double rho_bar, double c_bar, double lambda, double* vs);

typedef struct {
    double rho_bar, c_bar, pg_rho_bar, lambda, baseline_cost, entropy_cost;
} orc_vtrace_cfg;
/* Losses and their gradients w.r.t. learner logits [m,t,A] and values [m,t].
 * out_losses[4] = {total, pg, baseline, entropy}. */
void orc_vtrace_losses(int m, int t, int a, const double* logits, const double* value,
                       const float* mu_logits, const int32_t* action, const float* reward,
                       const float* discount, const float* bootstrap, const orc_vtrace_cfg* cfg,
                       double* out_losses, double* dlogits, double* dvalue, double* vs_out,
                       double* pg_adv_out);

/* ---------------- MLP actor-critic V-trace learner (this build's model) --------------- */
#define ORC_AC_PARAMS 1142801
#define ORC_AC_TENSORS 12
typedef struct orc_ac orc_ac;
orc_ac* orc_ac_create(const float* params, int opt_kind, double lr, const orc_vtrace_cfg* cfg);
void orc_ac_destroy(orc_ac* f);
void orc_ac_tensor_table(int64_t offsets[ORC_AC_TENSORS], int64_t numels[ORC_AC_TENSORS]);
/* obs [m,t,162] etc. as produced by orc_decode_vtrace. out_losses[4] as above. */
void orc_ac_loss_grad(orc_ac* f, const float* obs, const float* mu_logits, const int32_t* action,
                      const float* reward, const float* discount, const float* bootstrap, int m,
                      int t, double* out_losses);
/* Same, with the ReLU decisions of the 5 hidden layers supplied by the caller: relu_mask is
 * [5][m*t*512] bytes (non-zero = unit active). The GPU parity tests pass the CUDA run's own decisions
 * so that units whose pre-activation lies within fp32 rounding of zero (where an fp32 and a float64
 * forward legitimately disagree, and one disagreement changes a weight-gradient row by O(1/rows)) are
 * pinned. *n_override = units where the mask differs from z > 0; *max_override = largest |z|/rms(z)
 * among those (the test asserts both are tiny). */
void orc_ac_loss_grad_masked(orc_ac* f, const float* obs, const float* mu_logits, const int32_t* action,
                             const float* reward, const float* discount, const float* bootstrap, int m,
                             int t, double* out_losses, const uint8_t* relu_mask, int64_t* n_override,
                             double* max_override);
void orc_ac_forward(orc_ac* f, const float* obs, int rows, double* logits, double* value);
void orc_ac_opt_step(orc_ac* f);
void orc_ac_get_params(const orc_ac* f, double* out);
void orc_ac_get_grads(const orc_ac* f, double* out);
void orc_ac_set_grads(orc_ac* f, const double* in);

int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
