/* gode.h — C ABI of the B200-native latent-motion ODE hot path (libgode.so).
 *
 * Drop-in boundary for chechaohp/gan-ode: these entry points are what a binding of the
 * reference's solver calls would bind.  The reference is pure Python and reaches its solver
 * through two third-party call signatures:
 *
 *   torchdiffeq.odeint_adjoint(func, y0, t, method='rk4')      models/mocogan_ode.py:48-50,105-107,142-144
 *   torchdiffeq.odeint_adjoint(func, h, [0,1])  (dopri5)       models/mocogan_ode_rnn.py:47-48
 *   nn.GRUCell(e_t, h')  jump after each solve                 models/mocogan_ode_rnn.py:49
 *   torchsde.sdeint_adjoint(sde, x, ts, method='euler', dt)    models/mocogan_sde.py:57-59
 *
 * with func = ODEFunc (models/mocogan_ode.py:6-17): f(t,x) = W2 tanh(W1 x + b1) + b2.
 * The Python host side (gan_ode_b200/odeint.py …) keeps those signatures and calls the
 * functions below through ctypes.  INTEGRATION.md shows the stub.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / C++ types.  All tensors fp32, row-major,
 *    contiguous.  Weights use nn.Linear layout: W1 (H,D), b1 (H), W2 (D,H), b2 (D).
 *  - pointers are DEVICE pointers unless the name ends in _host.
 *  - every call is asynchronous on `stream` (a cudaStream_t), holds no global mutable state,
 *    never synchronises the device, and allocates nothing: outputs and workspaces come from
 *    the caller (PyTorch's caching allocator on the Python side).
 *  - WORKSPACES ARE PERSISTENT (v0.2): the first GODE_SYNC_REGION_BYTES of every `workspace`
 *    argument are the persistent region: grid-synchronisation words (counters, tagged all-reduce
 *    slots) and the tagged rows through which the backward kernels and the continuous adjoint sum
 *    parameter gradients over the grid without a barrier.  Only kernels of this library write there,
 *    always as {value | tag} words whose tags keep counting across launches.  The owner zero-fills
 *    it ONCE after allocation (gode_workspace_init) and may then pass the same workspace to any
 *    number of calls that are ordered on one stream, so no memset is enqueued per call.  Never share one workspace between launches
 *    that can run concurrently (different streams, concurrently replayed graphs).
 *  - return value: 0 OK; <0 argument / capability error (gode_strerror); launch failures are
 *    returned as -(1000 + cudaError_t).  Solver conditions that are only known on the device
 *    (dt underflow, non-finite state, step budget exhausted, checkpoint overflow) are written
 *    to the int status word of the GodeStepLog the caller passed.
 */
#ifndef GODE_H_
#define GODE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* gode_stream_t; /* cudaStream_t */

/* arithmetic of the MLP contractions */
enum {
  GODE_PREC_FP32 = 0, /* CUDA-core FFMA, <=1e-5 rel vs torchdiffeq                      */
  GODE_PREC_TF32 = 1, /* tcgen05 kind::tf32, fp32 accumulate in TMEM, <=2e-3 rel        */
  GODE_PREC_BF16 = 2  /* tcgen05 kind::f16 (bf16 operands), fp32 accumulate, <=2e-3 rel */
};

/* trajectory layout */
enum {
  GODE_LAYOUT_TBD = 0, /* (T,B,D): what odeint returns                                   */
  GODE_LAYOUT_BTD = 1  /* (B,T,D): == traj.transpose(0,1).reshape(-1,D) of models/mocogan_ode.py:146 */
};

/* error-norm scope of the adaptive controller */
enum {
  GODE_NORM_BATCH = 0, /* one RMS norm over all B*D elements, one (t,dt) for the batch — torchdiffeq */
  GODE_NORM_TRAJ = 1   /* per-trajectory RMS over D, per-trajectory (t,dt) — opt-in                 */
};

/* error codes (<0) */
enum {
  GODE_OK = 0,
  GODE_ERR_SHAPE = -1,       /* (D,H) has no compiled kernel for this precision            */
  GODE_ERR_ARG = -2,         /* null pointer, B<=0, T<2 ...                                */
  GODE_ERR_T_TOO_LONG = -3,  /* a time / step table exceeds its launch-parameter limit (below) */
  GODE_ERR_WORKSPACE = -4,   /* workspace smaller than gode_*_workspace_bytes()            */
  GODE_ERR_COOP = -5,        /* batch-global dopri5 needs all CTAs co-resident; B too big  */
  GODE_ERR_PRECISION = -6    /* precision mode not available for this entry point          */
};

/* device-side solver status bits (GodeStepLog.status) */
enum {
  GODE_ST_DT_UNDERFLOW = 1, /* torchdiffeq: assert t0 + dt > t0, 'underflow in dt'         */
  GODE_ST_NONFINITE = 2,    /* torchdiffeq: assert isfinite(y0), 'non-finite values in state' */
  GODE_ST_MAX_STEPS = 4,    /* torchdiffeq: 'max_num_steps exceeded'                       */
  GODE_ST_CKPT_OVERFLOW = 8, /* more accepted steps than checkpoint / log capacity         */
  GODE_ST_PEER_TIMEOUT = 16 /* world-scope norm: a peer rank did not publish within 10 s   */
};

/* adaptive tableaus (GodeAdaptiveOpts.tableau; 0 keeps every existing caller on dopri5) */
enum {
  GODE_TAB_DOPRI5 = 0,        /* dopri5.py: Dormand–Prince 5(4), FSAL                                   */
  GODE_TAB_BOSH3 = 1,         /* bosh3.py: Bogacki–Shampine 3(2), FSAL                                  */
  GODE_TAB_ADAPTIVE_HEUN = 2  /* adaptive_heun.py: Heun–Euler 2(1); f1 = k[-1] handed on as upstream does.  Not FSAL: the
                                 checkpoint buffer must hold 2 * ckpt_capacity rows (y0 of every step, then its f0) */
};

#define GODE_MAX_HOST_STEPS 255
/* other tables that travel in the launch parameters (exceeding one returns GODE_ERR_T_TOO_LONG) */
#define GODE_ADAPTIVE_MAX_T 256    /* output times of gode_dopri5_* / gode_adaptive_*                                 */
#define GODE_SDE_MAX_STEPS 320     /* Euler–Maruyama steps of one gode_sde_* solve (the reference: 41)                 */
#define GODE_SDE_MAX_FRAMES 64     /* output frames of one gode_sde_* solve (the reference: 16)                        */
#define GODE_SDE_MAX_CELLS 768     /* Brownian cells = distinct forward + reverse step end points (gode_sde_em_fwd_cells) */
#define GODE_SDE_MAX_REV_STEPS 384 /* reverse steps of gode_sde_adjoint_bwd (the reference: 45)                        */
/* persistent region at the front of every workspace: 256 KB of grid-sync words (counters, tagged all-reduce slots) followed by
 * 6 MB of tagged rows: the backward kernels' final reduction (296 CTAs x 544 values x 8 bytes) and the per-attempt reduction
 * of the continuous dopri5 adjoint (up to 296 CTAs x 2184 values x 8 bytes + one row of totals) */
#define GODE_SYNC_REGION_BYTES (6400 * 1024)

/* per-thread launch flags (gode_set_thread_launch_flags) */
enum {
  GODE_LAUNCH_PDL_BWD = 1 /* backward kernels are launched with programmatic stream serialisation: their
                             weight-staging prologue may overlap the tail of the kernel enqueued just before
                             them on the stream.  The caller guarantees that that kernel does not write the
                             weights (it normally is the matching forward).                               */
};

/* Step log of one adaptive solve, device resident, written by the forward kernel and read by the
 * backward kernel (no host round trip).  One entry per ATTEMPTED step in attempt order. */
typedef struct GodeStepLog {
  int32_t status;      /* OR of GODE_ST_*                                                  */
  int32_t n_attempts;  /* attempted steps                                                  */
  int32_t n_accepted;  /* accepted steps                                                   */
  int32_t nfe;         /* vector-field evaluations per trajectory                          */
  double dt0;          /* result of the initial-step heuristic                             */
  double t_final;      /* t1 of the last accepted step                                     */
} GodeStepLog;

/* adaptive controller options — torchdiffeq RKAdaptiveStepsizeODESolver defaults in comments */
typedef struct GodeAdaptiveOpts {
  double rtol;         /* 1e-7 */
  double atol;         /* 1e-9 */
  double first_step;   /* <=0: use the initial-step heuristic */
  double safety;       /* 0.9  */
  double ifactor;      /* 10   */
  double dfactor;      /* 0.2  */
  double min_step;     /* 0    */
  double max_step;     /* inf  */
  int32_t max_num_steps; /* per output interval; 2^31-1 */
  int32_t norm_scope;    /* GODE_NORM_BATCH */
  int32_t log_capacity;  /* entries in the attempt arrays (t0/dt/er/accepted)              */
  int32_t ckpt_capacity; /* accepted steps the checkpoint buffer can hold (0: none kept)   */
  float fsign;           /* +1, or -1 when the caller negated a decreasing time grid (torchdiffeq _ReverseFunc) */
  int32_t tableau;       /* GODE_TAB_*: gode_dopri5_fwd / gode_dopri5_backprop_bwd (batch-global control); others: dopri5 only */
} GodeAdaptiveOpts;

/* ---- introspection ------------------------------------------------------------------------- */
const char* gode_strerror(int code);
const char* gode_version(void);
/* zero-fill the sync region of a freshly allocated workspace (once; see "WORKSPACES ARE PERSISTENT") */
int gode_workspace_init(void* workspace, size_t ws_bytes, gode_stream_t stream);
/* capture state of `stream`: returns 1 and the capture's unique id while the stream is being captured into a CUDA
 * graph, 0 otherwise (<0: error).  Hosts key persistent workspaces by it: one workspace per captured graph. */
int gode_stream_capture_id(gode_stream_t stream, unsigned long long* id_out);
/* Status mailbox: ONE int in pinned, device-mapped HOST memory (set once per process, before the first solve; NULL turns
 * it off).  Every adaptive solve whose device status word is non-zero (GODE_ST_*) also stores it there, so a host that
 * does not synchronise per call still learns of a failed solve at its next natural check point by reading plain host
 * memory — nothing is written on the success path. */
int gode_set_status_mailbox(int32_t* host_mapped);
/* launch flags of the CALLING THREAD (thread-local; default 0), OR of GODE_LAUNCH_*; returns the previous value */
int gode_set_thread_launch_flags(int flags);
/* 1 if a kernel exists for (D,H) at this precision, else 0 */
int gode_supported(int D, int H, int precision);
/* number of floats in the flat parameter vector [W1|b1|W2|b2] = H*D + H + D*H + D */
int gode_param_count(int D, int H);

/* ---- a3: fixed-grid rk4 (3/8 rule) forward -------------------------------------------------- */
/* dt: T-1 step sizes dt_j = t[j+1]-t[j] rounded as torchdiffeq does (in t's dtype, then fp32).
 * dt_on_device=0: dt is a HOST array, passed by value in the launch (T-1 <= GODE_MAX_HOST_STEPS).
 * traj: (T,B,D) or (B,T,D); traj[0] = y0 bit-exact. */
int gode_rk4_fwd(const float* y0, const float* W1, const float* b1, const float* W2, const float* b2,
                 const float* dt, int dt_on_device, int B, int D, int H, int T, int precision,
                 int out_layout, float* traj, gode_stream_t stream);

/* ---- a4: rk4 continuous adjoint (torchdiffeq OdeintAdjointMethod.backward with method='rk4') -- */
/* grad_params: flat [W1|b1|W2|b2], OVERWRITTEN with this call's gradient (may be an NCCL buffer).
 * workspace: gode_bwd_workspace_bytes(B,D,H) bytes, contents undefined on entry. */
size_t gode_bwd_workspace_bytes(int B, int D, int H);
/* same, for any supported shape: wide fields (D,H multiples of 32, e.g. 64/256) additionally need scratch rows
 * proportional to B*(T-1) for the two-pass parameter-gradient contraction */
size_t gode_rk4_bwd_workspace_bytes(int B, int D, int H, int T);
int gode_rk4_adjoint_bwd(const float* traj, const float* grad_traj, const float* W1, const float* b1,
                         const float* W2, const float* b2, const float* dt, int dt_on_device, int B, int D,
                         int H, int T, int precision, int layout, float* grad_y0, float* grad_params,
                         void* workspace, size_t ws_bytes, gode_stream_t stream);

/* ---- A.5: rk4 backprop-through-solver (== autograd through torchdiffeq.odeint, method='rk4') -- */
int gode_rk4_backprop_bwd(const float* traj, const float* grad_traj, const float* W1, const float* b1,
                          const float* W2, const float* b2, const float* dt, int dt_on_device, int B, int D,
                          int H, int T, int precision, int layout, float* grad_y0, float* grad_params,
                          void* workspace, size_t ws_bytes, gode_stream_t stream);

/* ---- a5: adaptive dopri5 forward ------------------------------------------------------------- */
/* t_host: T output times (fp64, increasing; the caller negates decreasing grids as torchdiffeq does).
 * log: device GodeStepLog; att_t0/att_dt (double), att_er (float), att_acc (uint8): device arrays of
 * opts->log_capacity entries (may be NULL when log_capacity==0).
 * ckpt: (ckpt_capacity,B,D) state at the START of each accepted step, for backprop (NULL if capacity 0).
 * acc_t0/acc_dt: (ckpt_capacity) doubles, (t0,dt) of each accepted step.
 * workspace: gode_dopri5_workspace_bytes(B,D,H) bytes.
 * Shapes: the reference's D = H = 16 (lane-split kernels, every trajectory in registers, batch co-resident) and the wide
 * fields D=64/H=256, D=32/H=32, D=32/H=64 (csrc/wide_dopri5.cu: one warp per trajectory, the state between attempts in global
 * memory behind the sync region of the workspace, any batch size; dopri5 tableau only). */
size_t gode_dopri5_workspace_bytes(int B, int D, int H);
/* workspace of gode_dopri5_backprop_bwd: equal to the above for D = H = 16; for the wide fields it also holds the per-(step,
 * trajectory, stage) gradient rows of up to ckpt_capacity recorded steps, B * ckpt_capacity * 7 * (2D + 2H) floats. */
size_t gode_dopri5_backprop_workspace_bytes(int B, int D, int H, int ckpt_capacity);
int gode_dopri5_fwd(const float* y0, const float* W1, const float* b1, const float* W2, const float* b2,
                    const double* t_host, int B, int D, int H, int T, const GodeAdaptiveOpts* opts,
                    int out_layout, float* traj, GodeStepLog* log, double* att_t0, double* att_dt,
                    float* att_er, uint8_t* att_acc, float* ckpt, double* acc_t0, double* acc_dt,
                    void* workspace, size_t ws_bytes, gode_stream_t stream);

/* ---- e (optional): dopri5 with a WORLD-scope error norm ---------------------------------------------------------- */
/* Data-parallel ranks each hold a shard of the batch; with this entry point the step controller sees the RMS norm over
 * the trajectories of ALL ranks, so N ranks take the (t, dt) sequence one process would take on the whole batch
 * (torchdiffeq's batch-global norm).  The exchange is fused into the solver kernel: after the grid-wide
 * reduction of an attempt, CTA 0 stores this rank's partial as one tagged 64-bit word {fp32 | tag} into slot [rank] of
 * EVERY peer's exchange buffer over NVLink peer memory, and warp 0 of every CTA polls its own GPU's buffer until the
 * `world` words of this epoch are there, then adds them in rank order (identical on every rank).  No NCCL call, no host
 * round trip; parity double-buffering as in the intra-grid reduction.
 * slots_dev: device array[world] of device pointers, entry r = rank r's exchange buffer (GODE_WORLD_SLOT_WORDS(world)
 * uint64 words, zero-initialised once, mapped into this process: symmetric memory); launch_ctr: device word in local
 * memory, zero-initialised once, holding the cumulative count of exchanges (advanced by the kernel: tags stay unique
 * under CUDA-graph replay).  Every rank must
 * launch the same sequence of world-scope solves.  total_B: trajectories over all ranks. */
#define GODE_WORLD_SLOT_WORDS(world) (2 * 4 * (world))
typedef struct GodeWorld {
  int32_t rank, world;
  int64_t total_B;
  void* const* slots_dev;
  uint32_t* launch_ctr;
} GodeWorld;
int gode_dopri5_fwd_world(const float* y0, const float* W1, const float* b1, const float* W2, const float* b2,
                          const double* t_host, int B, int D, int H, int T, const GodeAdaptiveOpts* opts,
                          int out_layout, float* traj, GodeStepLog* log, double* att_t0, double* att_dt,
                          float* att_er, uint8_t* att_acc, float* ckpt, double* acc_t0, double* acc_dt,
                          void* workspace, size_t ws_bytes, const GodeWorld* world, gode_stream_t stream);

/* ---- A.5: dopri5 backprop-through-solver ------------------------------------------------------ */
/* Replays the accepted steps recorded by gode_dopri5_fwd in reverse (dt sequence treated as data),
 * including the dense-output interpolation of every requested time. */
int gode_dopri5_backprop_bwd(const float* grad_traj, const float* W1, const float* b1, const float* W2,
                             const float* b2, const double* t_host, int B, int D, int H, int T, int layout,
                             const GodeStepLog* log, const float* ckpt, const double* acc_t0,
                             const double* acc_dt, int ckpt_capacity, float fsign, float* grad_y0, float* grad_params,
                             void* workspace, size_t ws_bytes, gode_stream_t stream);

/* f4: the same replay for the other adaptive tableaus (`tableau` = the GODE_TAB_* the forward ran with through
 * GodeAdaptiveOpts.tableau: bosh3, adaptive_heun); tableau = GODE_TAB_DOPRI5 is gode_dopri5_backprop_bwd. */
int gode_adaptive_backprop_bwd(int tableau, const float* grad_traj, const float* W1, const float* b1, const float* W2,
                               const float* b2, const double* t_host, int B, int D, int H, int T, int layout,
                               const GodeStepLog* log, const float* ckpt, const double* acc_t0, const double* acc_dt,
                               int ckpt_capacity, float fsign, float* grad_y0, float* grad_params, void* workspace,
                               size_t ws_bytes, gode_stream_t stream);

/* ---- e: gode_dopri5_backprop_bwd with the data-parallel all-reduce of grad_params fused into its reduction tail ------ */
/* Same computation; in addition grad_params leaves the kernel summed over all ranks (rank-order sum, bit-identical on every
 * rank): the warp that finishes a float4 column of the parameter gradient stores it as tagged 64-bit words into every
 * peer's exchange buffer over NVLink peer memory and gathers the peers' words for that column — no separate all-reduce
 * launch.  exchange: rank / world; slots_dev[r] = rank r's buffer of GODE_GRAD_SLOT_WORDS(world, gode_param_count(D,H))
 * uint64 words, zero-initialised once and peer-mapped; launch_ctr: zero-initialised device word in local memory (cumulative
 * launch count, CUDA-graph replayable); total_B unused.  Every rank must launch the same sequence.  grad_y0 stays local. */
#define GODE_GRAD_SLOT_WORDS(world, P) (2 * (size_t)(world) * (size_t)(P))
int gode_dopri5_backprop_bwd_world(const float* grad_traj, const float* W1, const float* b1, const float* W2,
                                   const float* b2, const double* t_host, int B, int D, int H, int T, int layout,
                                   const GodeStepLog* log, const float* ckpt, const double* acc_t0,
                                   const double* acc_dt, int ckpt_capacity, float fsign, float* grad_y0,
                                   float* grad_params, void* workspace, size_t ws_bytes, const GodeWorld* exchange,
                                   gode_stream_t stream);

/* The same fused exchange for the FP32 rk4 backward kernels of the reference shape (D = H = 16): gode_rk4_adjoint_bwd
 * (adjoint != 0) or gode_rk4_backprop_bwd (adjoint == 0) with grad_params summed over all ranks inside the kernel. */
int gode_rk4_bwd_world(int adjoint, const float* traj, const float* grad_traj, const float* W1, const float* b1,
                       const float* W2, const float* b2, const float* dt, int dt_on_device, int B, int D, int H, int T,
                       int layout, float* grad_y0, float* grad_params, void* workspace, size_t ws_bytes,
                       const GodeWorld* exchange, gode_stream_t stream);

/* ---- a4 with the adaptive solver: torchdiffeq's continuous adjoint, method = adjoint_method = 'dopri5' ------------- */
/* Replaces OdeintAdjointMethod.backward for the call `odeint(self.ode_fn, h, tensor([0,1]))` of the ODE-RNN sampler
 * (models/mocogan_ode_rnn.py:47-48; `odeint` there is odeint_adjoint, :4).  Per output interval, from the last to the first:
 * a fresh dopri5 solve of the augmented state (y, a, theta_bar) against the forward direction — initial-step heuristic and
 * error ratios under torchdiffeq's default adjoint norm max(rms(y), rms(a), max over the four parameter tensors of
 * rms(theta_bar_k)); dense-output value at the interval's end; then y <- stored forward value, a += grad_traj[i-1].
 * traj / grad_traj: (T,B,D) or (B,T,D) per `layout`; t_host: the forward grid as passed to gode_dopri5_fwd (increasing,
 * opts->fsign = -1 if the caller's t was decreasing); opts: ADJOINT tolerances and controller options (rtol, atol,
 * first_step, safety, ifactor, dfactor, min/max_step, max_num_steps, log_capacity for the optional att_* arrays, which
 * hold the attempts of all intervals back to back; NULL = no log).  log (optional): status / attempts / accepted / nfe
 * summed over the intervals, dt0 of the last one.  param_mask: bit k set = parameter tensor k (W1, b1, W2, b2) is an adjoint
 * parameter (adjoint.py keeps those with requires_grad) and enters the norm; 15 = all.  One cooperative launch: the batch must be co-resident (<= 18944
 * trajectories; GODE_ERR_COOP beyond).  workspace: gode_dopri5_adjoint_workspace_bytes(B,D,H) bytes.  Deterministic. */
size_t gode_dopri5_adjoint_workspace_bytes(int B, int D, int H);
int gode_dopri5_adjoint_bwd(const float* traj, const float* grad_traj, const float* W1, const float* b1, const float* W2,
                            const float* b2, const double* t_host, int B, int D, int H, int T, int layout,
                            const GodeAdaptiveOpts* opts, int param_mask, float* grad_y0, float* grad_params,
                            GodeStepLog* log, double* att_dt, float* att_er, uint8_t* att_acc, void* workspace,
                            size_t ws_bytes, gode_stream_t stream);

/* ---- a5 (opt-in): dopri5 with PER-TRAJECTORY step control (GODE_NORM_TRAJ) ----------------------------------- */
/* Every trajectory has its own (t, dt), RMS error norm over its own D components and accept/reject sequence — what
 * torchdiffeq computes when called with B = 1 per trajectory.  No grid-wide reduction, ordinary launch, any B.
 * n_acc / n_att: (B) int32 accepted / attempted steps per trajectory.  att_dt (f64), att_er (f32), att_acc (u8):
 * (log_capacity, B) per-attempt logs or NULL.  ckpt: (ckpt_capacity, B, D); acc_t0 / acc_dt: (ckpt_capacity, B) f64.
 * log: status = OR over trajectories, n_attempts / n_accepted / nfe = max over trajectories. */
int gode_dopri5_traj_fwd(const float* y0, const float* W1, const float* b1, const float* W2, const float* b2,
                         const double* t_host, int B, int D, int H, int T, const GodeAdaptiveOpts* opts,
                         int out_layout, float* traj, GodeStepLog* log, int32_t* n_acc, int32_t* n_att,
                         double* att_dt, float* att_er, uint8_t* att_acc, float* ckpt, double* acc_t0,
                         double* acc_dt, gode_stream_t stream);
int gode_dopri5_traj_backprop_bwd(const float* grad_traj, const float* W1, const float* b1, const float* W2,
                                  const float* b2, const double* t_host, int B, int D, int H, int T, int layout,
                                  const GodeStepLog* log, const int32_t* n_acc, const float* ckpt,
                                  const double* acc_t0, const double* acc_dt, int ckpt_capacity, float fsign,
                                  float* grad_y0, float* grad_params, void* workspace, size_t ws_bytes,
                                  gode_stream_t stream);

/* ---- a7: neural SDE, fixed-step Euler–Maruyama (torchsde sdeint, method='euler', diagonal Ito noise) -------------- */
/* drift / diffusion: HOST arrays of 4 DEVICE pointers {W1, b1, W2, b2} (SDEFunc.drift_fn / diffusion_fn,
 * models/mocogan_sde.py:10-19).  The step grid is built on the host exactly as torchsde's fixed-step driver does
 * (fp32 time accumulation) and passed by value: h_host[n_steps] step sizes; frame j (1 <= j < T) is emitted after
 * step out_step_host[j] as w0_host[j]*y_k + w1_host[j]*y_{k+1} (linear interpolation); frame 0 = y0.
 * dW: (n_steps,B,D) Brownian increments, or NULL to generate them in the kernel: Philox4x32-10, key = seed,
 * counter = (traj_offset + b, step, d/4, 0) -> Box–Muller -> N(0,1) * sqrt(h)   (oracle/philox.py::normals).
 * states: (n_steps,B,D) state at the start of every step, kept for the backward (NULL: not kept). */
size_t gode_sde_workspace_bytes(int B, int D, int H);
int gode_sde_em_fwd(const float* y0, const float* const* drift, const float* const* diffusion, const float* h_host,
                    int n_steps, const int* out_step_host, const float* w0_host, const float* w1_host, int B, int D,
                    int H, int T, const float* dW, uint64_t seed, int64_t traj_offset, int out_layout, float* frames,
                    float* states, gode_stream_t stream);
/* exact reverse-mode through the Euler–Maruyama steps given the same increments (table or regenerated Philox).
 * grad_params: flat [drift: W1|b1|W2|b2 | diffusion: W1|b1|W2|b2], overwritten. */
int gode_sde_em_bwd(const float* states, const float* grad_frames, const float* const* drift,
                    const float* const* diffusion, const float* h_host, int n_steps, const int* out_step_host,
                    const float* w0_host, const float* w1_host, int B, int D, int H, int T, const float* dW,
                    uint64_t seed, int64_t traj_offset, int layout, float* grad_y0, float* grad_params,
                    void* workspace, size_t ws_bytes, gode_stream_t stream);

/* ---- f4: options['step_size'] under odeint_adjoint (torchdiffeq FixedGridODESolver + adjoint.py) -------------------------
 * The forward integrates on the grid t0, t0+h, ... and interpolates the requested times linearly (host: fine-grid
 * gode_fixed_fwd + interpolation); the adjoint re-solves every output interval [t_i, t_{i-1}] on ITS OWN grid t_i, t_i -+ h, ...
 * (last step clamped), carrying y along inside the interval and resetting it to traj[i-1] at the end.  Interval i
 * (i = T-1 .. 1) takes the sub-steps sub_dt[sub_beg[i] .. sub_end[i]) — device arrays, signed like the dt of
 * gode_rk4_adjoint_bwd (positive for an increasing t).  method: GODE_METHOD_RK4 / EULER / MIDPOINT.  traj: the forward's
 * (interpolated) outputs. */
int gode_fixed_adjoint_bwd_substep(int method, const float* traj, const float* grad_traj, const float* W1, const float* b1,
                                   const float* W2, const float* b2, const float* sub_dt, const int32_t* sub_beg,
                                   const int32_t* sub_end, int B, int D, int H, int T, int layout, float* grad_y0,
                                   float* grad_params, void* workspace, size_t ws_bytes, gode_stream_t stream);

/* ---- f2: the latent-motion sampler fused around the solve (models/mocogan_ode.py:133-148, models/mocogan.py:259-269) ----
 * OPT-IN (it changes which random numbers are consumed).  One launch does what sample_z_m does in five:
 *   x = randn(B, D)              Philox4x32-10, counter (global trajectory, 0, d/4, 2), oracle/philox.py::normals
 *   y0 = linear(x)               LeakyReLU(Wb LeakyReLU(Wa x + ba) + bb), Wa (pre_hidden, D), Wb (D, pre_hidden); pre_Wa NULL:
 *                                linear is nn.Identity (models/mocogan_ode.py:36-37)
 *   odeint(..., method='rk4')    3/8 rule on the grid dt, as gode_rk4_fwd
 *   .transpose(0,1).reshape(-1,D) and the torch.cat into z: out_layout GODE_LAYOUT_BTD with row stride ld_out floats (0 = D)
 *                                writes row b*T + j of a (B*T, ld_out) buffer at `out` (= &z[0][dim_z_content]).
 * traj_ids (device, B int64, or NULL): global trajectory index of every row — sample_images keeps num_samples rows of
 * num_samples*T*2 trajectories (models/mocogan.py:287-291); solving only those gives bit-identical codes.
 * noise_out (B,D) or NULL: the drawn x, for the pre-MLP's backward. */
int gode_rk4_sampler_fwd(const float* pre_Wa, const float* pre_ba, const float* pre_Wb, const float* pre_bb, float pre_slope,
                         int pre_hidden, const float* W1, const float* b1, const float* W2, const float* b2, const float* dt,
                         int dt_on_device, int B, int D, int H, int T, uint64_t seed, int64_t traj_offset,
                         const int64_t* traj_ids, int out_layout, float* out, int ld_out, float* noise_out,
                         gode_stream_t stream);
/* gode_rk4_adjoint_bwd for (B,T,D) codes that live inside wider buffers: traj rows are ld_traj floats apart, upstream-gradient
 * rows ld_grad floats apart (the gradient of z arrives as one (B*T, dim_z) tensor; its motion columns are read in place). */
int gode_rk4_adjoint_bwd_strided(const float* traj, int ld_traj, const float* grad_traj, int ld_grad, const float* W1,
                                 const float* b1, const float* W2, const float* b2, const float* dt, int dt_on_device, int B,
                                 int D, int H, int T, float* grad_y0, float* grad_params, void* workspace, size_t ws_bytes,
                                 gode_stream_t stream);

/* ---- f3: sdeint_adjoint with torchsde's stochastic adjoint (models/mocogan_sde.py:57-59, adjoint_method='euler') -------
 * torchsde/_core/adjoint.py re-solves every output interval backwards in time with its own dt grid, so the Brownian
 * path must answer increments over intervals that are not forward steps.  The path is sampled on CELLS = the union of the
 * forward and reverse step times (host-built, fp32 time accumulation as torchsde): cell r carries an N(0, len_r) increment,
 * Philox counter (traj_offset + b, r, d/4, 1) * cell_sqrt_host[r], or row r of dW_cells (R,B,D); a step's increment is
 * the left-to-right sum of its cells.  fwd_lo_host[k] .. fwd_lo_host[k+1] are the cells of forward step k (n_steps+1 ints).
 * The forward keeps nothing but its output frames. */
int gode_sde_em_fwd_cells(const float* y0, const float* const* drift, const float* const* diffusion, const float* h_host,
                          int n_steps, const int* out_step_host, const float* w0_host, const float* w1_host,
                          const int* fwd_lo_host, const float* cell_sqrt_host, int R, int B, int D, int H, int T,
                          const float* dW_cells, uint64_t seed, int64_t traj_offset, int out_layout, float* frames,
                          gode_stream_t stream);
/* The adjoint solve: for output interval i = T-1 .. 1 the reverse Euler steps [ibeg_host[i], iend_host[i]) (sizes
 * h_rev_host[n], increment = cells [rev_lo_host[n], rev_hi_host[n]) in forward time) of the augmented state
 * (y, a, theta_bar) with the Ito-corrected adjoint drift for diagonal noise (adjoint_sde.py::f_corrected_diagonal), then
 * y <- frames[i-1], a += grad_frames[i-1].  grad_params: flat [drift | diffusion], overwritten.  workspace as gode_sde_em_bwd. */
int gode_sde_adjoint_bwd(const float* frames, const float* grad_frames, const float* const* drift,
                         const float* const* diffusion, int n_rev, const float* h_rev_host, const int* rev_lo_host,
                         const int* rev_hi_host, const int* ibeg_host, const int* iend_host, const float* cell_sqrt_host,
                         int R, int B, int D, int H, int T, const float* dW_cells, uint64_t seed, int64_t traj_offset,
                         int layout, float* grad_y0, float* grad_params, void* workspace, size_t ws_bytes,
                         gode_stream_t stream);

/* ---- a6: ODE-RNN sampler (models/mocogan_ode_rnn.py:40-54) ---------------------------------------------------- */
/* The GRU jump h_out = GRUCell(x, h) of models/mocogan_ode_rnn.py:49 (nn.GRUCell semantics, models/mocogan.py:198:
 * gate order r, z, n; n = tanh(W_in x + b_in + r (W_hn h + b_hn)); h_out = (1 - z) n + z h).  D = 16.
 * Parameters as PyTorch stores them: w_ih (3D,D), w_hh (3D,D), b_ih (3D), b_hh (3D); flat gradient in that order. */
int gode_gru_param_count(int D);
int gode_gru_jump_fwd(const float* x, const float* h, const float* w_ih, const float* w_hh, const float* b_ih,
                      const float* b_hh, int B, int D, float* h_out, gode_stream_t stream);
/* grad_x may be NULL.  grad_params is OVERWRITTEN.  workspace: gode_odernn_workspace_bytes(B, D, D). */
int gode_gru_jump_bwd(const float* x, const float* h, const float* w_ih, const float* w_hh, const float* b_ih,
                      const float* b_hh, const float* grad_out, int B, int D, float* grad_x, float* grad_h,
                      float* grad_params, void* workspace, size_t ws_bytes, gode_stream_t stream);

/* The whole sampler in one call: for f in 0..F-1: h' = dopri5 solve of the ODEFunc over [0,1] from h_{f-1} (h_{-1} = h0)
 * with `opts` (the reference passes none: rtol 1e-7, atol 1e-9), h_f = GRUCell(eps[f], h').  codes: (F,B,D), codes[f] = h_f
 * (the reference's torch.cat(...).view(-1, D), models/mocogan_ode_rnn.py:51-52, is its (B,F,D) transpose).
 * Saved for the backward (all caller-allocated, device): seg (F,2,B,D) = each solve's [input copy, h']; logs = F slots of
 * gode_odernn_log_stride(opts->log_capacity) bytes ([GodeStepLog | attempt arrays], as gode_dopri5_fwd lays them out);
 * ckpt (F, ckpt_capacity, B, D) and acc (F, 2, ckpt_capacity) doubles, or NULL with opts->ckpt_capacity = 0 (no backward).
 * opts->norm_scope = GODE_NORM_TRAJ (opt-in): every trajectory controls its own step (gode_dopri5_traj_fwd); then n_acc is
 * an (F + 1, B) int32 buffer (accepted steps per frame and trajectory + one scratch row) and acc is (F, 2, ckpt_capacity, B);
 * with GODE_NORM_BATCH n_acc may be NULL.  The same n_acc / acc go to gode_odernn_bwd (n_acc = NULL selects batch mode). */
size_t gode_odernn_log_stride(int log_capacity);
size_t gode_odernn_workspace_bytes(int B, int D, int H);
int gode_odernn_fwd(const float* h0, const float* eps, const float* W1, const float* b1, const float* W2, const float* b2,
                    const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, int B, int D, int H,
                    int F, const GodeAdaptiveOpts* opts, float* codes, float* seg, void* logs, float* ckpt, double* acc,
                    int32_t* n_acc, void* workspace, size_t ws_bytes, gode_stream_t stream);
/* Reverse-mode through the F (solve, jump) pairs: GRU VJP, then the discrete adjoint of that frame's recorded solve.
 * grad_eps may be NULL.  grad_ode ([W1|b1|W2|b2]) and grad_gru are OVERWRITTEN (per-frame slots summed in frame order).
 * scratch: (3*B*D + F*gode_param_count(D,H)) floats.  adjoint_opts non-NULL: each frame's solve is differentiated by
 * torchdiffeq's continuous adjoint instead (gode_dopri5_adjoint_bwd with these tolerances / controller options, what the
 * reference loop computes); logs / ckpt / acc are then unused and may be NULL, and B is limited as there.
 * adjoint_param_mask: as gode_dopri5_adjoint_bwd's param_mask (bit k = tensor k of W1, b1, W2, b2 is an adjoint parameter,
 * i.e. part of the augmented state and of its error norm: torchdiffeq takes the parameters with requires_grad); ignored
 * when adjoint_opts is NULL. */
int gode_odernn_bwd(const float* grad_codes, const float* eps, const float* W1, const float* b1, const float* W2,
                    const float* b2, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, int B,
                    int D, int H, int F, int log_capacity, int ckpt_capacity, const float* seg, const void* logs,
                    const float* ckpt, const double* acc, const int32_t* n_acc, const GodeAdaptiveOpts* adjoint_opts,
                    int adjoint_param_mask, float* grad_h0, float* grad_eps, float* grad_ode, float* grad_gru, float* scratch, void* workspace,
                    size_t ws_bytes, gode_stream_t stream);

/* ---- e: data-parallel exchange ------------------------------------------------------------------------------------ */
/* One-shot all-reduce (sum, in place) of `n` floats over peer memory (NVLink / NVSwitch), the fused alternative to
 * ncclAllReduce for the flat parameter-gradient buffer a backward kernel has just written on the same stream.
 * bufs_dev / pads_dev: DEVICE arrays of `world` pointers to every rank's symmetric buffer (2 * world * cap floats) and
 * signal pad (>= 2 * world uint32, zero-initialised) as mapped into THIS process (e.g. torch symmetric memory's
 * buffer_ptrs_dev / signal_pad_ptrs_dev).  epoch_ctr: one zero-initialised device uint32 owned by this rank.
 * Every rank must issue the same sequence of calls.  Slots are added in rank order: results are bit-identical on all
 * ranks and from run to run.  data must be 16-byte aligned, n <= cap, cap % 4 == 0. */
int gode_allreduce_p2p(float* data, int n, void* const* bufs_dev, void* const* pads_dev, int rank, int world, int cap,
                       uint32_t* epoch_ctr, gode_stream_t stream);

/* ---- f4: the other fixed-grid methods of torchdiffeq for the same field (FP32, D = H = 16) ------------------------------- */
/* method: GODE_METHOD_RK4 (torchdiffeq 'rk4' = 3/8 rule; identical to the gode_rk4_* entry points), GODE_METHOD_EULER
 * (fixed_grid.py::Euler), GODE_METHOD_MIDPOINT (fixed_grid.py::Midpoint).  Same arguments, layouts and workspace as
 * gode_rk4_fwd / gode_rk4_adjoint_bwd (continuous adjoint re-solved per interval with the SAME method, adjoint.py) /
 * gode_rk4_backprop_bwd (reverse mode through the forward's arithmetic). */
#define GODE_METHOD_RK4 0
#define GODE_METHOD_EULER 1
#define GODE_METHOD_MIDPOINT 2
int gode_fixed_fwd(int method, const float* y0, const float* W1, const float* b1, const float* W2, const float* b2,
                   const float* dt, int dt_on_device, int B, int D, int H, int T, int out_layout, float* traj,
                   gode_stream_t stream);
int gode_fixed_adjoint_bwd(int method, const float* traj, const float* grad_traj, const float* W1, const float* b1,
                           const float* W2, const float* b2, const float* dt, int dt_on_device, int B, int D, int H, int T,
                           int layout, float* grad_y0, float* grad_params, void* workspace, size_t ws_bytes,
                           gode_stream_t stream);
int gode_fixed_backprop_bwd(int method, const float* traj, const float* grad_traj, const float* W1, const float* b1,
                            const float* W2, const float* b2, const float* dt, int dt_on_device, int B, int D, int H, int T,
                            int layout, float* grad_y0, float* grad_params, void* workspace, size_t ws_bytes,
                            gode_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GODE_H_ */
