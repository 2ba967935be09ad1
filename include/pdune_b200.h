/*
 * pdune_b200.h -- C ABI of the B200-native batched putting-dune simulator.
 *
 * The reference (google/putting-dune) has no FFI: its boundary for this path
 * is three synchronous Python protocols.  Every entry
 * point below names the reference interface it replaces (paths relative to
 * putting_dune/ in the reference).  Signatures use only plain pointers and
 * sizes; "device pointer" means a CUDA global-memory address (the Python host
 * passes torch.Tensor.data_ptr()); `stream` is a cudaStream_t passed as
 * void* (NULL = legacy default stream).
 *
 * Conventions
 *   - All calls return PD_OK (0) or a negative pd_status; pd_last_error()
 *     returns a thread-local message for the last failure.
 *   - No call allocates device memory; the caller owns every buffer.  Calls
 *     ending in _host take HOST pointers for the per-step inputs/outputs and
 *     copy them inside the call through caller-provided device staging.
 *   - Calls are asynchronous on `stream` unless documented otherwise; they are
 *     re-entrant across states, not thread-safe on one state.
 *   - Data-dependent failures that the reference raises per object
 *     (AssertionError on negative rates, graphene.py:258) are reported per
 *     env in pd_state.status (PD_ENV_* bits), never by aborting the batch.
 *
 * Random draws: Philox4x32-10, key = (seed lo, seed hi), counter =
 *   (global env id, seq, slot, stream).  See DESIGN.md "Random streams"; the
 *   same function in NumPy (oracle/pdune_oracle.py) lets identical draws be
 *   injected into the unmodified reference.
 */
#ifndef PDUNE_B200_H_
#define PDUNE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDUNE_B200_ABI_VERSION 2

typedef enum pd_status {
  PD_OK = 0,
  PD_ERR_INVALID_ARGUMENT = -1,
  PD_ERR_CUDA = -2,
  PD_ERR_UNSUPPORTED = -3,
  PD_ERR_NO_DEVICE = -4
} pd_status;

/* pd_state.status bits (per env). */
#define PD_ENV_OK 0u
#define PD_ENV_BAD_RATE 1u      /* rate < 0 or NaN (graphene.py:258 assert)   */
#define PD_ENV_LOG_OVERFLOW 2u  /* more transitions than the event log holds */
#define PD_ENV_NOT_RESET 4u     /* stepped before reset (simulator.py:224)    */

/* Rate functions (the RateFunction / CanonicalRatePredictionFn seam,
 * graphene.py:52-78). */
typedef enum pd_rate_fn {
  PD_RATE_SIMPLE = 0,   /* graphene.py:133-166 simple_canonical_rate_function */
  PD_RATE_PRIOR = 1,    /* graphene.py:169-229 HumanPriorRatePredictor.predict */
  PD_RATE_LEARNED = 2,  /* rate_learning/learn_rates.py:925-972 predict        */
  PD_RATE_CONSTANT = 3, /* fixed rates: the seam the reference's own tests     */
                        /* mock (simulator_test.py:139-168, graphene_test.py)  */
  PD_RATE_GMM = 4       /* graphene.py:279-390 GaussianMixtureRateFunction     */
} pd_rate_fn;

/* graphene.py:279-390: per neighbour, a mixture of Gaussians placed along the
 * Si->neighbour vector (mean = si + delta * loc_distance[m]) with variances
 * (along, across) that vector, scaled so that the largest mixture mode equals
 * max_rate.  Rates stay float64 in the reference (no float32 cast before the
 * total, graphene.py:375-388). */
#define PD_GMM_MAX_MIXTURES 16
typedef struct pd_gmm {
  int32_t n_mixtures;
  int32_t reserved_;
  double max_rate;
  double mixture_weights[PD_GMM_MAX_MIXTURES];
  double loc_distances[PD_GMM_MAX_MIXTURES];
  double variances[PD_GMM_MAX_MIXTURES][2];
} pd_gmm;

/* Philox stream ids (counter word 3). */
#define PD_STREAM_KMC 0u
#define PD_STREAM_RESET 1u
#define PD_STREAM_RENDER_POISSON 2u
#define PD_STREAM_RENDER_SP 3u
#define PD_STREAM_JITTER 4u
#define PD_STREAM_GOAL 5u
#define PD_STREAM_AGENT 6u
#define PD_STREAM_RENDER_UNIFORM 7u
#define PD_STREAM_RENDER_EXP 8u
#define PD_STREAM_RENDER_GAUSS 9u
#define PD_STREAM_SYNTH 10u
/* Renderer noise fields (counter = (env, frame_count, index, stream)): pixel
 * p = row * S + col takes word p % 4 of the call with index p / 4;
 * u24(w) = (w >> 8) * 2^-24.
 *   RENDER_POISSON  Poisson(image * mult): smallest k with CDF(k) > u24(w),
 *                   CDF by the float64 recurrence p_k = p_{k-1} * lam / k
 *   RENDER_SP       flip if u24(w) <= amount; salt if (w & 255) < 128
 *   RENDER_UNIFORM  + scale * u24(w)
 *   RENDER_EXP      + -log1p(-u24(w)) * scale
 *   RENDER_GAUSS    Box-Muller: words (0,1) -> pixels 4g, 4g+1; words (2,3) ->
 *                   4g+2, 4g+3; r = sqrt(-2 ln(((wa >> 8) + 1) 2^-24)),
 *                   t = 2 pi u24(wb); first pixel r cos t, second r sin t
 *   JITTER          index = image row, word 0: Poisson(jitter_rate) as above */

/* Shared lattice (graphene.py:464-559): device pointers, env-independent. */
typedef struct pd_lattice {
  int32_t n_cols;        /* grid_columns (reference default 50)              */
  int32_t n_sites;       /* 1881 for 50 columns                              */
  const double* base_xy; /* [n_sites][2] (G*1.42 - mean), graphene.py:537-543 */
  const int32_t* nbr;    /* [n_sites][4] 3-NN site ids, geometry.py:93; column
                          * 3 is library data: bits 0-23 list the sites near
                          * the lattice centre (terminated by 0xFFFFFF), bits
                          * 24-25 hold the site's neighbour-geometry class    */
} pd_lattice;

/* Per-env simulator state, struct of arrays, all device pointers of length
 * n_envs (x the inner extent shown).  This is everything that
 * PuttingDuneSimulator + PristineSingleDopedGraphene hold between calls
 * (SURVEY.md appendix A.1). */
typedef struct pd_state {
  int64_t n_envs;
  uint64_t seed;          /* Philox key                                        */
  uint32_t env_offset;    /* global id of local env 0 (multi-GPU sharding)     */
  uint32_t reserved_;
  int32_t* si_idx;        /* lattice site of the Si dopant                     */
  double* lattice;        /* [n][4] off_x, off_y, cos, sin (graphene.py:544-557)*/
  double* fov;            /* [n][4] ll_x, ll_y, ur_x, ur_y (simulator.py:79-82) */
  double* fov_scale;      /* [n]    FOV width (simulator.py:77)                */
  double* image_params;   /* [n][9] imaging.py:42-54, dataclass order          */
  uint32_t* episode;      /* resets seen (Philox seq of the RESET stream)      */
  uint32_t* ctrl_count;   /* apply_control calls seen (seq of the KMC stream)  */
  uint32_t* frame_count;  /* frames rendered (seq of the RENDER streams)       */
  int64_t* sim_time_us;   /* cumulative simulated time                         */
  int64_t* n_events;      /* cumulative rate evaluations (KMC iterations)      */
  int64_t* n_transitions; /* cumulative Si hops                                */
  uint8_t* status;        /* PD_ENV_* bits                                     */
} pd_state;

/* Learned rate model: eval-mode forward of get_mlp_fn
 * (rate_learning/learn_rates.py:80-99). float32 device pointers, row-major
 * [in][out] like the Haiku `w` leaves. */
typedef struct pd_mlp {
  int32_t context_dim;  /* D; only 2 is a valid drop-in          */
  int32_t hidden1;
  int32_t hidden2;
  int32_t batchnorm;    /* 0/1                                                */
  const float* bn_scale;   /* [D] batch_norm/scale                             */
  const float* bn_offset;  /* [D] batch_norm/offset                            */
  const float* bn_mean;    /* [D] batch_norm/~/mean_ema/average                */
  const float* bn_var;     /* [D] batch_norm/~/var_ema/average                 */
  const float* w0;         /* [D][H1]   mlp/~/linear_0/w                       */
  const float* b0;         /* [H1]                                             */
  const float* w1;         /* [H1][H2]  mlp/~/linear_1/w                       */
  const float* b1;         /* [H2]                                             */
  const float* w2;         /* [H2][4]   mlp/~/linear_2/w                       */
  const float* b2;         /* [4]                                              */
  /* Hidden contraction on the 5th-gen tensor cores (tcgen05.mma kind::f16,
   * BF16 operands, FP32 accumulate in TMEM) instead of FP32 FMA.  Rates then
   * agree with the FP32 path to ~1e-2 relative (bf16 operand rounding); the
   * default (0) is the FP32 parity path.  w1_umma: device copy of W1
   * transposed to [H2][H1], bf16, in the canonical no-swizzle K-major UMMA
   * layout (see pd_mlp_umma_layout_bytes / DESIGN.md).  Needs H1 % 16 == 0,
   * H2 % 16 == 0, 32 <= H2 <= 256, and H1*(128 + H2)*2 B of shared memory.
   * tensor_core = 2: both operands as fp16 hi + fp16 lo, three MMAs per K
   * step (hi hi + hi lo + lo hi) -- rates within 3e-7 of the FP32 path's, i.e.
   * a parity path; w1_umma then holds the hi tile followed by the lo tile
   * (fp16, same layout), and the shared memory doubles: hidden sizes up to
   * 128. */
  int32_t tensor_core;
  int32_t reserved_;
  const void* w1_umma;
} pd_mlp;

/* HumanPriorRatePredictor(mean, cov, max_rate) (graphene.py:181-189): the
 * peak position relative to a neighbour at (1, 0) in bond lengths, the
 * covariance of the Gaussian fall-off (symmetric positive definite; it stays
 * in the material frame, only the mean is rotated towards each neighbour,
 * graphene.py:222-227) and the rate at the peak. */
typedef struct pd_prior {
  double mean[2];
  double cov[2][2];
  double max_rate;
} pd_prior;

/* Rate-function selection passed to every stepping call. */
typedef struct pd_rate_config {
  int32_t rate_fn;         /* pd_rate_fn                                       */
  int32_t reserved_;
  const pd_mlp* mlp;       /* HOST pointer to the struct; PD_RATE_LEARNED only */
  float constant_rates[3]; /* PD_RATE_CONSTANT only                            */
  float reserved2_;
  const pd_gmm* gmm;       /* HOST pointer; PD_RATE_GMM only                   */
  const pd_prior* prior;   /* HOST pointer; PD_RATE_PRIOR only; NULL = the
                            * defaults of constants.py:26-28 ((0.85, 0),
                            * 0.1 I, ln 2 / 3)                                 */
} pd_rate_config;

/* Optional per-call outputs of the stepping calls (any pointer may be NULL).
 * The event log carries the observe_transition payload
 * (microscope_utils.py:516-521): for env e, transition k < log_count[e] of
 * this call happened log_elapsed_us[e][k] after its control was applied and
 * moved the Si to site log_site[e][k] during control log_ctrl[e][k]. */
typedef struct pd_step_out {
  int64_t* elapsed_us;     /* [n] MicroscopeObservation.elapsed_time           */
  int32_t* transitions;    /* [n] hops during this call                        */
  int32_t* events;         /* [n] rate evaluations during this call            */
  uint8_t* recentred;      /* [n] 1 if the FOV was re-centred (simulator.py:156)*/
  double* si_xy;           /* [n][2] Si position after the call, material frame */
  int32_t log_capacity;    /* K                                                */
  int32_t reserved_;
  int32_t* log_count;      /* [n]                                              */
  int64_t* log_elapsed_us; /* [n][K]                                           */
  int32_t* log_site;       /* [n][K]                                           */
  int32_t* log_ctrl;       /* [n][K] index of the control within the call      */
} pd_step_out;

/* ---- library ---------------------------------------------------------- */
int pd_abi_version(void);
const char* pd_last_error(void);
/* Number of SMs of the current device (grid sizing; 148 on B200). */
int pd_device_sm_count(int* out_sm_count);

/* ---- lattice: graphene.py:464-501 _generate_hexagonal_grid, :537-543,
 *      geometry.py:93-111 nearest_neighbors3 (canonical order) ------------ */
/* Host-only arithmetic: number of sites/rows for a column count. */
int pd_lattice_size(int32_t n_cols, int32_t* out_n_sites, int32_t* out_n_rows);
/* Builds base_xy [n_sites][2] and nbr [n_sites][4] on the device. */
int pd_build_lattice(int32_t n_cols, double* base_xy, int32_t* nbr,
                     void* stream);

/* ---- reset: simulator.py:65-105 PuttingDuneSimulator.reset,
 *      graphene.py:584-598 PristineSingleDopedGraphene.reset,
 *      imaging.py:42-54 sample_image_parameters --------------------------- */
/* mask: device uint8 [n] (NULL = all envs). */
int pd_reset(const pd_lattice* lat, const pd_state* st, const uint8_t* mask,
             void* stream);

/* ---- RateFunction seam: graphene.py:238-276
 *      PristineSingleSiGrRatePredictor.__call__ --------------------------- */
/* beam_xy: device double [n][2] material frame.  rates_out: float [n][3];
 * nbr_out: int32 [n][3] successor Si sites (either may be NULL). */
int pd_rates(const pd_lattice* lat, const pd_state* st,
             const pd_rate_config* rc, const double* beam_xy, float* rates_out,
             int32_t* nbr_out, void* stream);

/* ---- Material seam: graphene.py:646-694
 *      PristineSingleDopedGraphene.apply_control -------------------------- */
/* beam_xy: device double [n][2] MATERIAL frame; dwell_us: device int64 [n]
 * or NULL to use dwell_us_scalar for every env. */
int pd_apply_control(const pd_lattice* lat, const pd_state* st,
                     const pd_rate_config* rc, const double* beam_xy,
                     const int64_t* dwell_us, int64_t dwell_us_scalar,
                     const pd_step_out* out, void* stream);

/* ---- Simulator seam: simulator.py:107-182
 *      PuttingDuneSimulator.step_and_image (return_image=False part) ------ */
/* controls_xy: device double [n][n_controls][2] MICROSCOPE frame;
 * dwell_us: device int64 [n][n_controls] or NULL (scalar). */
int pd_step_and_image(const pd_lattice* lat, const pd_state* st,
                      const pd_rate_config* rc, const double* controls_xy,
                      const int64_t* dwell_us, int64_t dwell_us_scalar,
                      int32_t n_controls, int64_t image_duration_us,
                      const pd_step_out* out, void* stream);

/* Same call with HOST buffers for the per-step inputs and outputs: copies
 * controls (and dwell if non-NULL) host->device into the staging buffers,
 * steps, copies elapsed_us / si_xy / fov back, and synchronises `stream`.
 * h_* are host pointers (pinned for full speed), d_* device staging of the
 * same extents.  This is the call a reference user's step() maps to. */
int pd_step_and_image_host(const pd_lattice* lat, const pd_state* st,
                           const pd_rate_config* rc, const double* h_controls_xy,
                           const int64_t* h_dwell_us, int64_t dwell_us_scalar,
                           int32_t n_controls, int64_t image_duration_us,
                           double* d_controls_xy, int64_t* d_dwell_us,
                           const pd_step_out* d_out, int64_t* h_elapsed_us,
                           double* h_si_xy, double* h_fov, void* stream);

/* ---- beam-action stream: n_steps consecutive step_and_image calls with one
 *      control each, fused in one launch (state stays on chip between steps;
 *      semantics identical to calling pd_step_and_image n_steps times).
 *      controls_xy: device double [n_steps][n][2] MICROSCOPE frame.
 *      Per-step outputs (any may be NULL): si_idx_out int32 [n_steps][n],
 *      elapsed_us_out int64 [n_steps][n]. -------------------------------- */
int pd_rollout(const pd_lattice* lat, const pd_state* st,
               const pd_rate_config* rc, const double* controls_xy,
               int64_t dwell_us_scalar, int32_t n_steps,
               int64_t image_duration_us, int32_t* si_idx_out,
               int64_t* elapsed_us_out, void* stream);

/* Action adapters (action_adapters.py): how a value of the action stream
 * becomes a beam position in the microscope frame. */
typedef enum pd_action_mode {
  PD_ACTION_DIRECT = 0,   /* the value IS the position (DirectActionAdapter,
                             action_adapters.py:53-84, without the clip)     */
  PD_ACTION_RELATIVE_TO_SILICON = 1 /* RelativeToSiliconActionAdapter
                             (:131-216): clip(si + clip(a,-1,1) *
                             max_distance / fov_extent, 0, 1)                */
} pd_action_mode;

/* pd_rollout with an action adapter applied on the device before each step
 * (the observation the adapter reads is the env's own current one).
 * actions_xy: device double [n_steps][n][2]. */
int pd_rollout_actions(const pd_lattice* lat, const pd_state* st,
                       const pd_rate_config* rc, const double* actions_xy,
                       int32_t action_mode, double max_distance_angstroms,
                       int64_t dwell_us_scalar, int32_t n_steps,
                       int64_t image_duration_us, int32_t* si_idx_out,
                       int64_t* elapsed_us_out, void* stream);

/* Same with HOST buffers: copies the action stream host->device into
 * d_controls_xy, runs the fused steps, copies the per-step Si sites and
 * elapsed times back and synchronises `stream`.  h_si_idx / h_elapsed_us may
 * be NULL (then the matching d_* staging may be NULL too).  Small batches with
 * page-locked buffers take the streamed launch described at
 * pd_rollout_actions_host_f32 (float64 actions, int64 elapsed). */
int pd_rollout_host(const pd_lattice* lat, const pd_state* st,
                    const pd_rate_config* rc, const double* h_controls_xy,
                    int64_t dwell_us_scalar, int32_t n_steps,
                    int64_t image_duration_us, double* d_controls_xy,
                    int32_t* d_si_idx, int64_t* d_elapsed_us,
                    int32_t* h_si_idx, int64_t* h_elapsed_us, void* stream);
int pd_rollout_actions_host(const pd_lattice* lat, const pd_state* st,
                            const pd_rate_config* rc,
                            const double* h_actions_xy, int32_t action_mode,
                            double max_distance_angstroms,
                            int64_t dwell_us_scalar, int32_t n_steps,
                            int64_t image_duration_us, double* d_actions_xy,
                            int32_t* d_si_idx, int64_t* d_elapsed_us,
                            int32_t* h_si_idx, int64_t* h_elapsed_us,
                            void* stream);
/* pd_rollout_actions_host with compact host formats: actions as float32 (the
 * dtype the action adapters' action_spec declares, action_adapters.py:80-84,
 * 202-216; widened exactly on the device) and per-step elapsed time as int32
 * microseconds (dwell + 2 * image_duration must fit).  Results equal
 * pd_rollout_actions_host on the widened actions.  Staging (device):
 * d_actions_f32 float [n_steps][n][2], d_controls_xy double [n_steps][n][2],
 * d_si_idx int32 / d_elapsed_us int64 / d_elapsed_us32 int32 [n_steps][n];
 * their contents after the call are unspecified.  All five may be NULL: the
 * library then keeps its own (grow-only, per calling thread and device) and
 * prepares them for the next call behind the current one.
 *
 * Small batches on the prior / simple rates (what k_rollout_pre covers in one
 * wave of CTAs; n a multiple of 16, >= 2^18 env-steps) with page-locked host
 * buffers (cudaHostAlloc / torch pin_memory; the result buffers must be
 * device-visible) run as ONE launch that overlaps both copies: the actions
 * come in through a single copy-engine copy that the stepping CTAs follow
 * element by element, the CTAs of a few SMs write finished result rows
 * straight to h_si_idx / h_elapsed_us32.  An action whose two float32 words
 * are both 0xFFFFFFFF is only consumed once the copy has ended.  Everything
 * else (and PD_HOST_STREAMED=0) takes the chunked copy-engine pipeline of
 * pd_rollout_actions_host.  Both forms return the same bytes. */
int pd_rollout_actions_host_f32(
    const pd_lattice* lat, const pd_state* st, const pd_rate_config* rc,
    const float* h_actions_xy, int32_t action_mode,
    double max_distance_angstroms, int64_t dwell_us_scalar, int32_t n_steps,
    int64_t image_duration_us, float* d_actions_f32, double* d_controls_xy,
    int32_t* d_si_idx, int64_t* d_elapsed_us, int32_t* d_elapsed_us32,
    int32_t* h_si_idx, int32_t* h_elapsed_us32, void* stream);

/* Packed host format of the same call (prior / simple rates, one positive
 * dwell time below 3000 s, at most 32768 lattice sites): float32 actions
 * [n_steps][n][2] in, one uint16 per env-step out, [n_steps][n]:
 *   bits 0-14  Si lattice site after the step,
 *   bit 15     the step re-centred the FOV (simulator.py:156-169),
 * from which MicroscopeObservation.elapsed_time follows as dwell + image
 * duration * (1 + bit 15) (simulator.py:131-169).  8 + 2 bytes per env-step
 * cross PCIe instead of 8 + 8.  The library keeps the device stagings; the
 * action stream goes through the copy engines in a few chunks of whole steps
 * that the stepping kernels and the result copies follow (no kernel reads a
 * buffer a copy is still writing).  Page-locked host buffers make the copies
 * asynchronous; pageable ones work.  Synchronises `stream`. */
int pd_rollout_actions_host_packed(
    const pd_lattice* lat, const pd_state* st, const pd_rate_config* rc,
    const float* h_actions_xy, int32_t action_mode,
    double max_distance_angstroms, int64_t dwell_us_scalar, int32_t n_steps,
    int64_t image_duration_us, uint16_t* h_packed, void* stream);

/* ---- the guarded float32 iteration (rollouts on the prior / simple rates).
 * pd_rollout_actions with one positive dwell time below 3000 s decides every
 * iteration of graphene.py:658-694 (hop or not, which neighbour) in float32
 * when the float32 result is further from the deciding threshold than a bound
 * of its own error, and replays the control with the float64 code otherwise,
 * so results equal the float64 kernels' bit for bit (csrc/pd_fast.cuh).
 * pd_set_fast_path(0) sends every iteration through the float64 code (A/B
 * timing, parity tests); returns the previous setting.  Default 1; the
 * environment variable PD_FAST=0 sets the default to 0. */
int pd_set_fast_path(int enabled);

/* Process-wide kernel-selection options (A/B timing and the parity tests that
 * compare the kernels with each other); each is initialised once from the
 * environment variable of the same meaning.  value: 0 / 1.
 *   "fast_path"     PD_FAST          guarded float32 iteration (above)
 *   "prepass"       PD_PREPASS       float32 pre-pass of the float64 kernels
 *   "rollout_spec"  PD_ROLLOUT_SPEC  look-ahead over idle lanes in small
 *                                    float64 rollouts
 *   "plan"          PD_PLAN          (default 0) small-batch rollouts under the
 *                                    relative adapter through k_rollout_plan
 *                                    (per-CTA plan of every (env, step) in
 *                                    shared memory) instead of k_rollout_fast;
 *                                    same results, slower on the benchmarked
 *                                    workload
 *   "walk_plan"     PD_WALK_PLAN     (default 1) large-batch rollouts under the
 *                                    relative adapter (>= 3 waves of CTAs, >=
 *                                    8 steps) through k_walk_plan, followed by
 *                                    k_walk_fast over the envs it hands over,
 *                                    instead of k_walk_fast alone; 2 = for
 *                                    every large batch (parity tests)
 *   "mlp_slim"      PD_MLP_SLIM      (default 1) learned-rate step on the
 *                                    tensor cores with 256 threads and two
 *                                    CTAs per SM where two sets of operand
 *                                    tiles fit in shared memory (H <= 64 in
 *                                    the fp16 hi + lo form), so that one
 *                                    CTA's item build / event phases run under
 *                                    the other's wave; 0 = 512 threads, one
 *                                    CTA per SM for every shape; same results
 *   "race_sampling" PD_SAMPLING_RACE (default 0) events by the race of
 *                                    competing exponentials -- each neighbour
 *                                    draws Exp(rate_i), the smallest wins --
 *                                    instead of the reference's direct method
 *                                    (graphene.py:658-694).  Equal in
 *                                    distribution, not draw for draw: never
 *                                    the parity path.  Scalar rate functions
 *                                    only (not PD_RATE_LEARNED). */
int pd_set_option(const char* name, int value);

/* Measures the float32 quantities of that iteration against the float64 ones
 * over n_samples random iterations (random site, lattice angle, beam offset
 * within max_distance of the Si, Philox draw, clock).  The *_over_bound
 * fields are the largest observed error divided by the bound the fast path
 * assumes for it (must stay below 1, in practice below ~0.3);
 * wrong_decision / wrong_slot count decided iterations that differ from the
 * exact code (must be 0). */
typedef struct pd_fast_audit {
  int64_t samples, no_hop, hop, unsure;
  int64_t wrong_decision, wrong_slot, waiting_time_outside_bounds;
  double total_rate_error_over_bound;
  double waiting_time_error_over_bound;
  double choice_error_over_bound;
  /* the float32 unit-exponential draw against float64 over all 2^24 values
   * of its 24-bit uniform (exhaustive), over the bound assumed for it */
  double draw_error_over_bound;
} pd_fast_audit;
int pd_fast_path_audit(const pd_lattice* lat, int32_t rate_fn, uint64_t seed,
                       int64_t n_samples, int64_t dwell_us,
                       double max_distance_angstroms, pd_fast_audit* out,
                       void* stream);

/* The float64 rate expressions of the kernels against the reference's
 * operation sequence (graphene.py:151-166, :210-229), each neighbour of
 * n_samples random (site, lattice angle, beam offset within max_distance)
 * triples = 3 n_samples evaluations per rate function:
 *   simple_max_ulps / prior_max_ulps  largest difference of the two float64
 *       forms, in float64 ulps;
 *   simple_cast_differs_unguarded     float32 casts of the short form that
 *       differ from the operation sequence's, before the guard;
 *   simple_guard_taken                evaluations the guard (within
 *       guard_ulps of a float32 rounding boundary) sends to the operation
 *       sequence;
 *   simple_cast_differs               differing casts the guard missed (must
 *       be 0: the simple rate is the reference's float32 value, bit for bit);
 *   prior_cast_differs                differing casts of the human prior
 *       (unguarded: the reference evaluates it in JAX float32, no float64
 *       form is "the" reference there). */
typedef struct pd_rate_ops_stats {
  int64_t evaluations;
  int64_t simple_cast_differs_unguarded, simple_cast_differs;
  int64_t simple_guard_taken, prior_cast_differs;
  uint32_t simple_max_ulps, prior_max_ulps, guard_ulps, reserved_;
} pd_rate_ops_stats;
int pd_rate_ops_audit(const pd_lattice* lat, uint64_t seed, int64_t n_samples,
                      double max_distance_angstroms, pd_rate_ops_stats* out,
                      void* stream);

/* ---- imaging.py:42-72: re-draws the nine image parameters of the envs'
 *      current episode from the uniforms the last pd_reset used (RESET draws
 *      4..12): sample_image_parameters (DEFAULT, what pd_reset itself applies)
 *      or sample_noisy_image_parameters (NOISY).  mask: device uint8 [n] or
 *      NULL.  Call after pd_reset. ---------------------------------------- */
typedef enum pd_image_params_mode {
  PD_IMAGE_PARAMS_DEFAULT = 0,
  PD_IMAGE_PARAMS_NOISY = 1
} pd_image_params_mode;
int pd_sample_image_params(const pd_state* st, const uint8_t* mask,
                           int32_t mode, void* stream);

/* ---- queries: graphene.py:600-644 get_atoms_in_bounds,
 *      graphene.py:696-700 get_silicon_position --------------------------- */
/* fov_override: device double [n][4] or NULL (use st->fov).  Outputs are
 * padded to max_atoms per env, in lattice order: out_xy double
 * [n][max_atoms][2] normalised to the box, out_z uint8 [n][max_atoms]
 * (6 / 14), out_site int32 [n][max_atoms] (may be NULL), out_count int32 [n]
 * (the true count even if it exceeds max_atoms). */
int pd_get_atoms_in_bounds(const pd_lattice* lat, const pd_state* st,
                           const double* fov_override, int32_t max_atoms,
                           double* out_xy, uint8_t* out_z, int32_t* out_site,
                           int32_t* out_count, void* stream);
int pd_get_silicon_position(const pd_lattice* lat, const pd_state* st,
                            double* out_xy, void* stream);
/* All atom positions of a subset of envs (the `.grid` attribute,
 * graphene.py:581): env_ids device int32 [m]; out_xy double [m][n_sites][2]. */
int pd_get_grid(const pd_lattice* lat, const pd_state* st,
                const int32_t* env_ids, int32_t m, double* out_xy,
                void* stream);

/* ---- trajectory export: microscope_utils.py:72-131 AtomicGrid.to_proto,
 *      :180-230 BeamControl.to_proto, :496-501 MicroscopeFieldOfView.to_proto,
 *      :589-604 MicroscopeObservation.to_proto, :737-757 Trajectory.to_proto,
 *      io.py:65-82 write_records (putting_dune.proto:7-49) ----------------- */
/* Serialises every env's current observation -- observed grid (atoms inside
 * st->fov in lattice order, microscope frame), FOV, the controls just applied
 * and the elapsed time -- as the protobuf wire bytes of MicroscopeObservation
 * (fields 1-4; float64 -> float32 round-to-nearest as the protobuf runtime
 * does), on the device.
 *   controls_xy device double [n][n_controls][2] (microscope frame);
 *   dwell_us device int64 [n][n_controls] or NULL (dwell_us_scalar);
 *   elapsed_us device int64 [n] or NULL (st->sim_time_us);
 *   max_atoms: staging bound per record; records with more atoms (or beyond
 *   `capacity` bytes of out_bytes) are skipped and flagged in out_overflow
 *   (uint8 [n], may be NULL);
 *   out_offsets device int64 [n + 1]: record e starts at out_offsets[e]
 *   (16-byte aligned), out_offsets[n] = bytes needed; out_len int32 [n]: its
 *   length (= pd_observation_bytes(out_atoms[e], n_controls)); out_atoms
 *   int32 [n]. */
int pd_encode_observations(const pd_lattice* lat, const pd_state* st,
                           const double* controls_xy, const int64_t* dwell_us,
                           int64_t dwell_us_scalar, int32_t n_controls,
                           const int64_t* elapsed_us, float voltage_kv,
                           float current_na, int32_t max_atoms,
                           uint8_t* out_bytes, int64_t capacity,
                           int64_t* out_offsets, int32_t* out_len,
                           int32_t* out_atoms, uint8_t* out_overflow,
                           void* stream);
/* Wire size of an observation with `atoms` atoms and n_controls controls. */
int64_t pd_observation_bytes(int32_t atoms, int32_t n_controls);
/* HOST: frames one Trajectory record per env (its observations of steps
 * 0..n_steps-1, each as field 1) into a TFRecord stream: uint64 length,
 * masked CRC-32C of the length, payload, masked CRC-32C of the payload.
 * step_bytes[t] / step_offsets[t] / step_len[t] are HOST copies of one
 * pd_encode_observations call's outputs.  out == NULL only sizes the stream
 * (*out_size). */
int pd_tfrecord_trajectories(int32_t n_steps, int64_t n_envs,
                             const uint8_t* const* step_bytes,
                             const int64_t* const* step_offsets,
                             const int32_t* const* step_len, uint8_t* out,
                             int64_t out_capacity, int64_t* out_size);
/* HOST: CRC-32C (Castagnoli) of a buffer. */
uint32_t pd_crc32c(const void* data, int64_t size);

/* ---- synthetic rate-learning data: rate_learning/data_utils.py:158-303
 *      generate_synthetic_data, PRIOR mode (sample_from_prior :237-283) and
 *      NETWORK mode (sample_network_rates :201-234) ------------------------- */
/* n samples of split `split` (0 = train, 1 = test; data_utils.py:297-300):
 * next_state int32 [n] (0 = no transition, k + 1 = state k), dt float [n]
 * (the observation window), rates float [n][num_states], context float
 * [n][context_dim], position float [n][2]; all device pointers.  Draws are
 * keyed by Philox (PD_STREAM_SYNTH), not jax.random. */
int pd_generate_synthetic_data(uint64_t seed, int32_t split, int64_t n,
                               int32_t num_states, int32_t context_dim,
                               float time_lo, float time_hi,
                               int32_t* next_state, float* dt, float* rates,
                               float* context, float* position, void* stream);

/* NETWORK mode: x ~ N(0, I) [context_dim + position_dim], rates =
 * softplus(MLP(x))[:num_states] with the MLP of learn_rates.py:80-99
 * (batchnorm=False; swish between layers; sizes x -> hidden0 -> hidden1 ->
 * num_states + 1; the reference uses (1, 64)).  Weights: device float,
 * row-major [in][out] like Haiku's hk.Linear (the reference draws them with
 * jax.random; here they are the caller's).  context float [n][context_dim] =
 * x[:context_dim], position float [n][position_dim] = x[context_dim:]. */
int pd_generate_synthetic_data_network(
    uint64_t seed, int32_t split, int64_t n, int32_t num_states,
    int32_t context_dim, int32_t position_dim, float time_lo, float time_hi,
    const float* w0, const float* b0, const float* w1, const float* b1,
    const float* w2, const float* b2, int32_t hidden0, int32_t hidden1,
    int32_t* next_state, float* dt, float* rates, float* context,
    float* position, void* stream);

/* ---- whole goal-reaching episodes on the device (BASELINE configs[4]):
 *      eval_lib.py:77-184 evaluate for the greedy_on_neighbor experiment
 *      (experiments/registry.py:287-298) = PuttingDuneEnvironment.reset/step
 *      (putting_dune_environment.py:87-158) + SingleSiliconGoalReaching
 *      (goals.py:70-185) + SingleSiliconMaterialFrameFeatureConstructor
 *      (feature_constructors.py:157-228) + GreedyAgent.step
 *      (agents/agent_lib.py:163-183) +
 *      RelativeToSiliconMaterialFrameActionAdapter
 *      (action_adapters.py:219-274) + StepLimitWrapper (run_helpers.py:120). */
typedef struct pd_episode_config {
  int64_t dwell_us;           /* 5 s  (registry.py:291-294)                    */
  int64_t image_duration_us;  /* 2 s  (simulator.py:37)                        */
  int64_t timeout_us;         /* 10 min simulated (eval_lib.py:82)             */
  int32_t step_limit;         /* 600  (run_helpers.py:34)                      */
  int32_t reserved_;
  double argmax_x, argmax_y;  /* greedy beam offset for a neighbour on +x,     */
                              /* angstroms: (1.42, 0) (registry.py:289)        */
} pd_episode_config;

/* One record per env: eval_lib.EvalResult (eval_lib.py:47-59), 16 bytes, the
 * unit that is all-gathered across GPUs. */
typedef struct pd_episode_stats {
  int32_t num_actions;   /* num_actions_taken                                  */
  float env_seconds;     /* environment_seconds_to_goal (NaN if not reached)   */
  float total_reward;    /* gamma^elapsed of the terminal step, else 0         */
  uint8_t reached_goal;
  uint8_t pad_[3];
} pd_episode_stats;

/* Resets every env (pd_reset), draws its goal (draw 13 of the RESET stream)
 * and runs the greedy controller until the goal is reached, the step limit or
 * the simulated-time limit.  goal_xy: device double [n][2] workspace/output
 * (goal position, material frame); goal_site: device int32 [n] or NULL;
 * stats: device pd_episode_stats [n]. */
int pd_run_episodes(const pd_lattice* lat, const pd_state* st,
                    const pd_rate_config* rc, const pd_episode_config* cfg,
                    double* goal_xy, int32_t* goal_site,
                    pd_episode_stats* stats, void* stream);

/* ---- batched RL environment layer:
 *      PuttingDuneEnvironment.reset/step (putting_dune_environment.py:87-158)
 *      under StepLimitWrapper (run_helpers.py:120-153), with the four action
 *      adapters (action_adapters.py:53-274), the two 10-float feature
 *      constructors (feature_constructors.py:79-228) and
 *      SingleSiliconGoalReaching (goals.py:70-185), for every env at once. -- */
typedef enum pd_adapter {
  PD_ADAPTER_DIRECT = 0,            /* action_adapters.py:53-84               */
  PD_ADAPTER_DELTA = 1,             /* :87-128 (stateful beam position)       */
  PD_ADAPTER_RELATIVE = 2,          /* :131-216                               */
  PD_ADAPTER_RELATIVE_MATERIAL = 3  /* :219-274                               */
} pd_adapter;
typedef enum pd_features {
  PD_FEATURES_MICROSCOPE = 0,  /* feature_constructors.py:79-154              */
  PD_FEATURES_MATERIAL = 1     /* :157-228                                    */
} pd_features;
/* dm_env.StepType */
#define PD_STEP_FIRST 0
#define PD_STEP_MID 1
#define PD_STEP_LAST 2

typedef struct pd_env_config {
  int32_t adapter;      /* pd_adapter                                        */
  int32_t features;     /* pd_features                                       */
  int32_t action_dim;   /* 2, or 3 when the relative adapters take a dwell   */
  int32_t step_limit;   /* StepLimitWrapper (600 in run_helpers.py:34)       */
  double min_dwell_s, max_dwell_s;   /* adapter dwell range (1.5, 1.5)       */
  double max_distance_angstroms;     /* RelativeToSilicon adapter            */
  int64_t image_duration_us;
} pd_env_config;

/* Per-env environment state and scratch, device arrays of length n. */
typedef struct pd_env_buffers {
  double* goal_xy;          /* [n][2] goal position, material frame          */
  double* beam_pos;         /* [n][2] DeltaPositionActionAdapter state       */
  int32_t* elapsed_steps;   /* StepLimitWrapper counter (-1 = truncated)     */
  uint8_t* needs_reset;     /* 1 = next step resets (initially 1)            */
  double* controls_xy;      /* [n][2] scratch: adapter output                */
  int64_t* dwell_us;        /* [n]    scratch                                */
  int64_t* elapsed_us;      /* [n]    scratch: step elapsed time             */
  uint8_t* resetting;       /* [n]    scratch: envs that reset in this call  */
} pd_env_buffers;

/* One env.step(action) per env: envs whose previous step was LAST (or that
 * were never reset) reset instead and return FIRST.  actions: device double
 * [n][action_dim]; outputs: observation float [n][10], reward float [n],
 * discount float [n], step_type int32 [n]. */
int pd_env_step(const pd_lattice* lat, const pd_state* st,
                const pd_rate_config* rc, const pd_env_config* cfg,
                const pd_env_buffers* buf, const double* actions,
                float* observation, float* reward, float* discount,
                int32_t* step_type, void* stream);

/* ---- learned model, batched form: rate_learning/learn_rates.py:704-732
 *      LearnedTransitionRatePredictor.apply_model -- mean over an ensemble of
 *      softmax(out[:3]) * out[3].  models: HOST array of n_models pd_mlp;
 *      x: device float [n][2]; out: device float [n][3]. ------------------- */
int pd_mlp_apply_model(const pd_mlp* models, int32_t n_models, const float* x,
                       int64_t n, float* out, void* stream);

/* ---- renderer: imaging.py:239-265 generate_stem_image (and the stages it
 *      chains, :117-236) for a set of envs, using each env's current FOV,
 *      image_params and Si position; simulator.py:206-221 _generate_image. -- */
typedef enum pd_render_stage {
  PD_RENDER_CLEAN = 0,       /* imaging.py:117-173 generate_clean_image       */
  PD_RENDER_BLUR = 1,        /* :212-214 apply_blur                           */
  PD_RENDER_POISSON = 2,     /* :199-203 apply_poisson_noise                  */
  PD_RENDER_JITTER = 3,      /* :188-196 apply_jitter                         */
  PD_RENDER_UNIFORM = 4,     /* :206-209 s&p, :217-218 gamma, :231-236 uniform */
  PD_RENDER_EXPONENTIAL = 5, /* :221-228 apply_exponential_noise              */
  PD_RENDER_GAUSSIAN = 6,    /* :176-185 apply_gaussian_noise                 */
  PD_RENDER_FINAL = 7        /* :264 exposure.equalize_adapthist (CLAHE)      */
} pd_render_stage;

/* Workspace the renderer needs: *out_bytes of device memory (any alignment
 * of 256 B), independent of the number of frames. */
int pd_render_workspace_bytes(int32_t image_size, int64_t* out_bytes);

/* Frames the device renders concurrently: one thread-block cluster of eight
 * CTAs (eight SMs) per frame; benchmarks size their batches as a multiple. */
int pd_render_clusters(int32_t image_size, int32_t* out_clusters);

/* Renders frames for envs env_ids[0..m) (device int32; NULL = envs 0..m-1).
 * frames_out: device float [m][image_size][image_size], values in [0, 1]
 * (the reference returns float64; pixels agree to float32 tolerance, see
 * DESIGN.md).  image_size: power of two in [64, 512].  stop_stage < FINAL
 * returns the image after that stage (parity tests).  advance_frame_count != 0
 * increments pd_state.frame_count of the rendered envs (the Philox sequence of
 * the noise fields).  Frames with at most 1024 atoms in view, a clean-image
 * kernel radius <= 128 px and blur_amount < 1.125 (everything imaging.py:42-72
 * samples at FOV >= 7.5 A) stay in shared memory of an 8-CTA cluster; others
 * take a slower one-CTA path through the workspace.  Same results.
 * buffer_size (imaging.py:123, in [0, 0.25]): 0 renders the atoms inside the
 * FOV, as the simulator does; > 0 is generate_clean_image's buffered form with
 * the whole lattice as its grid -- atoms up to buffer_size FOV widths outside
 * the frame contribute their Gaussian tails. */
int pd_render(const pd_lattice* lat, const pd_state* st, const int32_t* env_ids,
              int32_t m, int32_t image_size, int32_t stop_stage,
              int32_t advance_frame_count, double buffer_size,
              float* frames_out, void* workspace, int64_t workspace_bytes,
              void* stream);

/* ---- label masks: imaging.py:75-114 generate_grid_mask for envs
 *      env_ids[0..m) (NULL = envs 0..m-1) with each env's current FOV.
 *      mask_out: device uint8 [m][image_size][image_size], 0 / 6 / 14.
 *      radius_carbon / radius_silicon = (Z / 6)^intensity_exponent * 0.1,
 *      evaluated by the caller in float64 (the host mirror does) so that the
 *      comparison `squared distance < radius` sees the reference's value.
 *      image_size: multiple of 8. ----------------------------------------- */
int pd_render_mask(const pd_lattice* lat, const pd_state* st,
                   const int32_t* env_ids, int32_t m, int32_t image_size,
                   double radius_carbon, double radius_silicon,
                   uint8_t* mask_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PDUNE_B200_H_ */
