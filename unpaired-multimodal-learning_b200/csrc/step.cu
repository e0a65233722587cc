// One UML iteration behind ONE C-ABI call: the host enqueues every kernel of the step
// (reference loop body, vision_language/finetune.py:163-195) without returning to Python in between.
// At the reference's batch sizes the step is launch-latency bound, at throughput batch sizes the host
// must stay ahead of a ~150 us GPU step - either way per-kernel Python dispatch is what limits it.
#include <cstdlib>

#include "common.cuh"
#include "optim.cuh"
#include "gather.cuh"

namespace {
inline void rec(void* ev, void* stream) {
  if (ev) cudaEventRecord(static_cast<cudaEvent_t>(ev), uml::as_stream(stream));
}

bool overlap_fixup();

// Side stream + events for the gather prefetch of uml_linear_run (one set per device, created on first use).
struct Pipe {
  cudaStream_t aux = nullptr;
  cudaStream_t aux2 = nullptr;          // the dW GEMM when it overlaps the fix-up launch
  cudaEvent_t dw_done = nullptr, fwd_done = nullptr;
  unsigned* split_done = nullptr;       // [8] fix-up CTAs finished per dW split, then one int: watchdog flag
  cudaEvent_t ready[3] = {nullptr, nullptr, nullptr}, freed[3] = {nullptr, nullptr, nullptr}, start = nullptr, mid = nullptr;
  // the pipeline of uml_linear_run survives the call boundary when the caller vouches for its index batches (idx_ready):
  bool warm = false;                    // freed[] / fwd_mid[] describe the last steps of the previous call
  int phase = 0;                        // operand buffer of the next step
  unsigned gstep = 0;                   // steps enqueued so far (fwd_mid[gstep & 1] is recorded after step gstep's forward)
  const void* bufs[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t fwd_mid[2] = {nullptr, nullptr};
  bool ok = false;
};

// any other user of the operand buffers (single-step calls, Python-side kernels) makes the next run start cold
void pipe_invalidate();

static Pipe g_pipes[64];
void pipe_invalidate() {
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64) g_pipes[dev].warm = false;
}

Pipe* get_pipe() {
  Pipe* pipes = g_pipes;
  static bool tried[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  Pipe& p = pipes[dev];
  if (!tried[dev]) {
    tried[dev] = true;
    const char* e = getenv("UML_PREFETCH");
    if (e && e[0] == '0') return nullptr;
    int lo = 0, hi = 0;
    if (cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess) return nullptr;
    // lowest priority: the prefetch must never delay a kernel of the step it hides under
    const char* pe = getenv("UML_AUX_PRIORITY");  // experiments: "hi" = the GEMMs' priority class and above
    const int aux_prio = (pe && pe[0] == 'h') ? hi : lo;
    if (cudaStreamCreateWithPriority(&p.aux, cudaStreamNonBlocking, aux_prio) != cudaSuccess) return nullptr;
    bool good = true;
    for (int i = 0; i < 3; ++i) {
      good = good && cudaEventCreateWithFlags(&p.ready[i], cudaEventDisableTiming) == cudaSuccess;
      good = good && cudaEventCreateWithFlags(&p.freed[i], cudaEventDisableTiming) == cudaSuccess;
    }
    good = good && cudaEventCreateWithFlags(&p.start, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 2; ++i) good = good && cudaEventCreateWithFlags(&p.fwd_mid[i], cudaEventDisableTiming) == cudaSuccess;
    good = good && cudaEventCreateWithFlags(&p.mid, cudaEventDisableTiming) == cudaSuccess;
    good = good && cudaEventCreateWithFlags(&p.dw_done, cudaEventDisableTiming) == cudaSuccess;
    good = good && cudaEventCreateWithFlags(&p.fwd_done, cudaEventDisableTiming) == cudaSuccess;
    good = good && cudaStreamCreateWithPriority(&p.aux2, cudaStreamNonBlocking, hi) == cudaSuccess;
    if (overlap_fixup())  // (a device allocation synchronises: only when the experiment is switched on)
      good = good && cudaMalloc(&p.split_done, 64) == cudaSuccess && cudaMemset(p.split_done, 0, 64) == cudaSuccess;
    p.ok = good;
  }
  return p.ok ? &p : nullptr;
}

// Where in step i the gather of step i+1 is enqueued.  Measured on B200 (cfg3, 2 x 18944 rows): at the START of the
// step (it then shares the machine with the forward and, mostly, the bandwidth-bound fix-up) 0.1925 ms/step;
// AFTER forward + fix-up (sharing with the dW GEMM, which streams both operands from HBM) 0.237 ms/step - slower
// than no prefetch at all (0.207); right AFTER THE FORWARD KERNEL (sharing with fix-up, dW, update) 0.189 ms/step
// with the forward kernel running undisturbed (67.5 us instead of 84 us).  UML_PREFETCH_AT = 0 | 1 | 2 selects
// start / after fix-up / after the forward kernel (default 2).  Placement 3 needs no side stream: the copy of step
// i+1's rows rides in step i's fix-up launch as extra, interleaved CTAs - measured 0.1936 ms/step against 0.1862 for
// placement 2 (two bandwidth-bound jobs in one launch just add up: 49.7 us for the launch instead of 24.5 us).
// Placement 4: after the dW GEMM, i.e. under the step's data-parallel tail (that kernel waits on NVLink most of the time).
static bool g_dp_step = false;  // set by linear_run for the step being enqueued
// UML_UPDATE_IN_DW=1: the split-K sum and AdamW run in the tail of the dW GEMM instead of a launch of their own.
// Measured on B200 (cfg3, 73 728 rows): 0.2376 ms/step against 0.2319 with the separate launch - every CTA waits for the
// slowest split of its tile with the whole SM in hand, and its share of the update (21 rows x 256 columns, nine
// arrays) is latency-bound, while the separate kernel spreads the same bytes over 750 CTAs.  Off by default.
bool update_in_dw_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("UML_UPDATE_IN_DW");
    on = (e && e[0] == '1') ? 1 : 0;
  }
  return on == 1;
}

int prefetch_placement() {
  static int cached = -2;
  if (cached == -2) {
    const char* e = getenv("UML_PREFETCH_AT");
    cached = (e && e[0] >= '0' && e[0] <= '4') ? e[0] - '0' : -1;
  }
  if (cached >= 0) return cached;
  (void)g_dp_step;  // measured on 2 B200s: placement 4 0.1627 ms/step, placement 2 0.1618 - no gain, 2 stays the default
  return 2;
}

// UML_FUSE_FIX=1: the deferred softmax normalisation is applied in the dW prologue (tc_gemm kFix: each G stage is
// rescaled in shared memory between the TMA arrival and the MMA) instead of by a fix-up pass over G after the
// forward kernel.  Bit-identical results, but measured SLOWER on B200 (cfg3): 0.231 vs 0.189 ms/step - the dW
// GEMM with two MN-major operands already moves ~100 KB through shared memory per 950-cycle stage, the transform's
// extra 32 KB per stage makes it shared-memory bound.  Default: the fix-up kernel.
bool fuse_fix() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("UML_FUSE_FIX");
    cached = (e && e[0] == '1') ? 1 : 0;
  }
  return cached == 1;
}

// UML_OVERLAP_FIXUP=1: the dW GEMM is launched on a second stream as soon as the forward kernel ends and each of its K
// splits starts when the fix-up CTAs of ITS rows are done (per-split counters), instead of the whole GEMM waiting for
// the whole fix-up pass.  Bit-identical, but measured SLOWER on B200 (cfg3): 0.205 vs 0.185 ms/step - the resident,
// waiting GEMM CTAs plus the prefetching gather leave the fix-up CTAs almost no registers to run in.  Default: off.
bool overlap_fixup() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("UML_OVERLAP_FIXUP");
    cached = (e && e[0] == '1') ? 1 : 0;
  }
  return cached == 1;
}

// true when every non-empty run of the step can be gathered from bf16 shadow banks by one copy launch
bool shadow_gatherable(const uml_linear_step_args* a) {
  for (int i = 0; i < a->nseg; ++i) {
    const uml_segment& s = a->seg[i];
    if (s.n > 0 && !(s.rows16 && s.idx && !s.label_idx)) return false;
  }
  return true;
}

// both runs' rows + labels -> (X16, labels32) by one launch of the TMA copy kernel (or its small-footprint twin)
int shadow_gather(const uml_linear_step_args* a, uint16_t* X16, int32_t* labels32, bool light, void* stream) {
  const uml_segment* g[2] = {nullptr, nullptr};
  int ng = 0;
  for (int i = 0; i < a->nseg; ++i)
    if (a->seg[i].n > 0) g[ng++] = &a->seg[i];
  if (ng == 0) return 0;
  // The register-copy kernel is also the faster one stand-alone (ncu, cfg3: 17.8 us = 6.5 TB/s, the measured copy
  // peak, against 30 us for the TMA ring kernel), so it serves both the prefetch and the in-line gather.
  (void)light;
  auto fn = uml_gather2_rows_bf16_light;
  return fn(g[0]->rows16, g[0]->labels, g[0]->idx, g[0]->n, ng > 1 ? g[1]->rows16 : nullptr, ng > 1 ? g[1]->labels : nullptr,
            ng > 1 ? g[1]->idx : nullptr, ng > 1 ? g[1]->n : 0, a->dim, X16, a->dim, labels32, stream);
}
}  // namespace

// `mid` (optional) runs on the host right after the forward + fix-up launches were enqueued: the place where
// uml_linear_run enqueues the next step's gather, so that it overlaps dW / update but not the bandwidth-bound fix-up
struct StepHooks {
  bool pregathered = false;
  cudaEvent_t operand_free = nullptr;
  int (*mid)(void*) = nullptr;
  void* mid_arg = nullptr;
  cudaEvent_t after_fwd_kernel = nullptr;  // placement 2: recorded between the forward kernel and its fix-up
  const uml::GatherJob* merged = nullptr;  // placement 3: the fix-up launch also copies the NEXT step's rows
};
static int linear_step_impl(const uml_linear_step_args* a, void* stream, const StepHooks& hooks);
int uml_head_fwd_ce_bf16_ev(const uint16_t* X, int64_t n_rows, int32_t dim, const uint16_t* W, int32_t n_classes,
                            const int32_t* labels, const uml_tc_segments* segs, uint16_t* G, int64_t ldg, float* row_loss,
                            int32_t* row_pred, int32_t* row_correct, float* row_dscale, float* tile_ws, uml_seg_stats* stats,
                            void* ev_after_fwd, void* stream, int defer_fixup, void* ev_after_fwd2 = nullptr,
                            const uml::GatherJob* job = nullptr, const uml::FixupSignal* sig = nullptr);  // tc_fwd.cu
int uml_head_bwd_dw_gated_bf16(const uint16_t* G, int64_t ldg, const uint16_t* X, int64_t n_rows, int32_t dim, int32_t n_classes,
                               float* partials, int32_t n_splits, const unsigned* done, int* failed, void* stream);  // tc_gemm.cu

// tc_fwd2.cu / tc_gemm.cu: exchange forward kernel (G final after one pass) and the dW GEMM that reduces its statistics
bool uml_fwd_x_eligible(int64_t n_rows, int32_t n_classes);
void uml_fwd_x_partials(float* tile_ws, int64_t n_rows, int32_t n_classes, const float** part, int64_t* n_entries);
int uml_head_bwd_dw_stats_bf16(const uint16_t* G, int64_t ldg, const uint16_t* X, int64_t n_rows, int32_t dim, int32_t n_classes,
                               float* partials, int32_t n_splits, const float* part, int64_t part_entries, int32_t nseg,
                               uml_seg_stats* stats, void* stream);

int uml_head_bwd_dw_update_bf16(const uint16_t* G, int64_t ldg, const uint16_t* X, int64_t n_rows, int32_t dim, int32_t n_classes,
                                float* partials, int32_t n_splits, const float* part, int64_t part_entries, int32_t nseg,
                                uml_seg_stats* stats, float* W, float* m, float* v, uint16_t* W16, const uml::AdamArgs* adam,
                                unsigned* failed, void* stream);

float* uml_dp_p2p_input(int64_t n);   // dp.cu: this rank's exchange buffers of the peer-memory all-reduce (or NULL)
float* uml_dp_p2p_output(int64_t n);

// where the local gradient sum of a data-parallel step goes: straight into the peer-visible exchange buffer when
// the NVLink all-reduce is set up, else into dW_out (NCCL reduces that in place)
static float* dp_local_sum_target(const uml_linear_step_args* a, int64_t np) {
  float* x = uml_dp_p2p_input(np);
  return x ? x : a->dW_out;
}

// data parallel tail of a step: sum dW over the ranks, then the optimizer update on every rank
static int dp_reduce_and_update(const uml_linear_step_args* a, int64_t np, void* stream) {
  int rc;
  const float* grad = a->dW_out;
  if (uml_dp_p2p_input(np)) {
    if (a->upd.kind != 3) {  // exchange + Adam fused (the local sum already is in the exchange buffer)
      rec(a->ev[6], stream);
      rc = uml_dp_fused_adam_update(nullptr, 0, 0, np, a->W, a->upd.m, a->upd.v, a->upd.lr, a->upd.beta1, a->upd.beta2,
                                    a->upd.eps, a->upd.weight_decay, a->upd.step, a->upd.kind == 1,
                                    a->precision == 1 ? a->W16 : nullptr, stream);
      rec(a->ev[7], stream);
      return rc;
    }
    rc = uml_dp_allreduce_p2p(np, stream);
    grad = uml_dp_p2p_output(np);
  } else {
    rc = uml_dp_allreduce_f32(a->dW_out, np, stream);
  }
  if (rc) return rc;
  uint16_t* shadow = a->precision == 1 ? a->W16 : nullptr;
  rec(a->ev[6], stream);
  if (a->upd.kind == 3)
    rc = uml_sgd_step(a->W, grad, nullptr, 0.f, a->upd.m, np, a->upd.lr, a->upd.momentum, a->upd.weight_decay,
                      a->upd.step, shadow, stream);
  else
    rc = uml_adamw_step(a->W, grad, nullptr, 0.f, a->upd.m, a->upd.v, np, a->upd.lr, a->upd.beta1, a->upd.beta2,
                        a->upd.eps, a->upd.weight_decay, a->upd.step, a->upd.kind == 1, shadow, stream);
  rec(a->ev[7], stream);
  return rc;
}


static bool continuous_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("UML_PREFETCH_CONTINUOUS");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

static int linear_run_continuous(const uml_linear_step_args* base, const uml_run_step* steps, int32_t n_steps, void* stream,
                                 Pipe* pipe) {
  using namespace uml;
  constexpr int kNb = 3, kDepth = 2;
  uml_linear_step_args a = *base;
  auto patch = [&](uml_linear_step_args& t, const uml_run_step& s) {
    for (int k = 0; k < t.nseg; ++k) {
      t.seg[k].idx = s.idx[k];
      t.seg[k].n = s.n[k];
      t.seg[k].loss_weight = s.loss_weight[k];
      t.scale_step[k] = s.scale_step[k];
    }
    t.upd.lr = s.lr;
    t.upd.step = s.opt_step;
    t.stats = s.stats;
    for (int k = 0; k < 8; ++k) t.ev[k] = s.ev[k];
  };
  uint16_t* xbuf[kNb] = {base->X16, base->X16_alt, base->X16_alt2};
  int32_t* lbuf[kNb] = {base->labels32, base->labels32_alt, base->labels32_alt2};
  cudaStream_t main_st = as_stream(stream);
  const cudaEvent_t idx_ready = static_cast<cudaEvent_t>(base->idx_ready);
  if (!(pipe->warm && pipe->bufs[0] == xbuf[0] && pipe->bufs[1] == xbuf[1] && pipe->bufs[2] == xbuf[2])) {
    // cold: everything enqueued so far may still read the buffers
    pipe->warm = false;
    pipe->phase = 0;
    for (int k = 0; k < kNb; ++k) pipe->bufs[k] = xbuf[k];
    UML_CUDA(cudaEventRecord(pipe->start, main_st));
  }
  const bool warm = pipe->warm;
  const int phase = pipe->phase;
  const unsigned g0 = pipe->gstep;

  // the gather of step t of this call -> side stream.  `after_fwd`: the forward kernel it follows (so that it shares the
  // machine with a dW GEMM and an update, never - but for its tail - with a forward kernel)
  auto launch_gather = [&](int t, cudaEvent_t after_fwd) -> int {
    const int tb = (phase + t) % kNb;
    uml_linear_step_args g = a;
    patch(g, steps[t]);
    UML_CUDA(cudaStreamWaitEvent(pipe->aux, idx_ready, 0));
    // the dW kernel that last read this buffer: step t - 3 of this call, or of the previous one (warm), else `start`
    UML_CUDA(cudaStreamWaitEvent(pipe->aux, (t >= kNb || warm) ? pipe->freed[tb] : pipe->start, 0));
    if (after_fwd) UML_CUDA(cudaStreamWaitEvent(pipe->aux, after_fwd, 0));
    rec(g.ev[0], pipe->aux);
    const int rc = shadow_gather(&g, xbuf[tb], lbuf[tb], true, pipe->aux);
    if (rc) return rc;
    rec(g.ev[1], pipe->aux);
    UML_CUDA(cudaEventRecord(pipe->ready[tb], pipe->aux));
    return 0;
  };
  for (int t = 0; t < kDepth && t < n_steps; ++t) {
    // inside one long call step t's gather would have followed the forward kernel of step t - 2: steps -2 and -1 are the
    // previous call's last two
    const int rc = launch_gather(t, warm ? pipe->fwd_mid[(g0 + t) & 1u] : nullptr);  // (g0 + t - 2) & 1
    if (rc) return rc;
  }

  g_dp_step = base->dp_allreduce != 0;
  for (int i = 0; i < n_steps; ++i) {
    patch(a, steps[i]);
    if (i > 0) a.w16_valid = 1;  // the optimizer kernel of the previous step refreshed the shadow
    const int b = (phase + i) % kNb;
    a.X16 = xbuf[b];
    a.labels32 = lbuf[b];
    UML_CUDA(cudaStreamWaitEvent(main_st, pipe->ready[b], 0));
    struct Next {
      decltype(launch_gather)* launch;
      int t;
      cudaEvent_t mid;
      cudaStream_t main_st;
    } next{&launch_gather, i + kDepth < n_steps ? i + kDepth : -1, pipe->fwd_mid[(g0 + i) & 1u], main_st};
    StepHooks hooks;
    hooks.pregathered = true;
    hooks.operand_free = pipe->freed[b];
    hooks.mid_arg = &next;
    hooks.after_fwd_kernel = next.mid;  // recorded right after the forward kernel, also for the next call's first gathers
    hooks.mid = [](void* p) -> int {
      Next* n = static_cast<Next*>(p);
      return n->t >= 0 ? (*n->launch)(n->t, n->mid) : 0;
    };
    const int rc = linear_step_impl(&a, stream, hooks);
    if (rc) {
      pipe->warm = false;
      return rc;
    }
  }
  pipe->phase = (phase + n_steps) % kNb;
  pipe->gstep = g0 + static_cast<unsigned>(n_steps);
  pipe->warm = true;
  return 0;
}

extern "C" {

int uml_linear_run(const uml_linear_step_args* base, const uml_run_step* steps, int32_t n_steps, void* stream) {
  using namespace uml;
  UML_REQUIRE(base && steps && n_steps >= 0, "linear_run: bad arguments");
  uml_linear_step_args a = *base;
  auto patch = [&](uml_linear_step_args& t, const uml_run_step& s) {
    for (int k = 0; k < t.nseg; ++k) {
      t.seg[k].idx = s.idx[k];
      t.seg[k].n = s.n[k];
      t.seg[k].loss_weight = s.loss_weight[k];
      t.scale_step[k] = s.scale_step[k];
    }
    t.upd.lr = s.lr;
    t.upd.step = s.opt_step;
    t.stats = s.stats;
    for (int k = 0; k < 8; ++k) t.ev[k] = s.ev[k];
  };

  // Gather prefetch: step i+1's rows are copied into the operand buffer step i does not use, on a low-priority
  // side stream, while step i's forward / fix-up / dW / update run.  (The copy depends only on the indices and
  // the bank, never on the weights.)  Buffer b becomes free again once the dW kernel that read it has finished.
  Pipe* pipe = nullptr;
  if (a.precision == 1 && a.X16_alt && a.labels32_alt && n_steps > 1) {
    bool all = true;
    for (int i = 0; i < n_steps && all; ++i) {
      uml_linear_step_args t = a;
      patch(t, steps[i]);
      all = shadow_gatherable(&t);
    }
    if (all) pipe = get_pipe();
  }
  // ---- continuous mode: three operand buffers and an event that vouches for the index batches (idx_ready) ----------
  // The gather pipeline then runs through the call boundaries: the first two steps' gathers go to the side stream at once,
  // ordered after the index upload, after the dW kernel that last read their buffer and after the forward kernel of the
  // step they would have followed inside one long call - when the host is a chunk ahead of the GPU (it normally is) they
  // run beside the previous call's last steps, and a call costs no gather on the main stream at all.
  if (a.precision == 1 && base->X16_alt && base->labels32_alt && base->X16_alt2 && base->labels32_alt2 && base->idx_ready &&
      n_steps >= 1 && prefetch_placement() == 2 && !fuse_fix() && continuous_enabled()) {
    bool all = true;
    for (int i = 0; i < n_steps && all; ++i) {
      uml_linear_step_args t = a;
      patch(t, steps[i]);
      all = shadow_gatherable(&t);
    }
    Pipe* cp = all ? get_pipe() : nullptr;
    if (cp) return linear_run_continuous(base, steps, n_steps, stream, cp);
  }
  pipe_invalidate();
  // With a third operand buffer the gather runs TWO steps ahead: step i + 2's rows are copied while step i's dW and
  // update run and may spill into step i + 1 - no forward kernel ever waits for an event that is signalled at the last
  // moment (one step ahead, the gather of step i + 1 ended about when the update did, and the cross-stream hand-over
  // put ~9 us between the update and the next forward kernel).
  int nb = 2;
  if (pipe && base->X16_alt2 && base->labels32_alt2 && n_steps > 2 && prefetch_placement() != 3) {
    const char* e = getenv("UML_PREFETCH_DEPTH");
    if (!(e && e[0] == '1')) nb = 3;
  }
  const int depth = nb - 1;
  uint16_t* xbuf[3] = {base->X16, base->X16_alt, base->X16_alt2};
  int32_t* lbuf[3] = {base->labels32, base->labels32_alt, base->labels32_alt2};
  cudaStream_t main_st = as_stream(stream);
  if (pipe) UML_CUDA(cudaEventRecord(pipe->start, main_st));  // everything enqueued so far may still read the other buffers

  g_dp_step = base->dp_allreduce != 0;
  for (int i = 0; i < n_steps; ++i) {
    patch(a, steps[i]);
    if (i > 0 && a.precision == 1) a.w16_valid = 1;  // the optimizer kernel of the previous step refreshed the shadow
    if (!pipe) {
      const int rc = uml_linear_step(&a, stream);
      if (rc) return rc;
      continue;
    }
    const int b = i % nb;
    a.X16 = xbuf[b];
    a.labels32 = lbuf[b];
    if (i == 0) {
      // the first step's rows in line; with two steps of look-ahead the first step's hook (after its forward kernel)
      // starts BOTH the second and the third step's gathers: beside the first forward kernel a gather would only have
      // the SMs that kernel leaves free (4 of 148) and hold up the second step
      rec(a.ev[0], stream);
      const int rc0 = shadow_gather(&a, xbuf[0], lbuf[0], false, stream);
      if (rc0) return rc0;
      rec(a.ev[1], stream);
    } else if (prefetch_placement() != 3 || fuse_fix()) {
      UML_CUDA(cudaStreamWaitEvent(main_st, pipe->ready[b], 0));
    }
    struct Job {
      uml_linear_step_args nx;
      uint16_t* x;
      int32_t* l;
      cudaEvent_t wait_a, ready;
    };
    struct Next {
      Pipe* pipe;
      Job job[2];
      int n_jobs;
      cudaEvent_t wait_b;
      cudaStream_t main_st;
      bool have, record_mid;
    } next;
    next.pipe = pipe;
    next.n_jobs = 0;
    next.record_mid = prefetch_placement() != 2;
    next.wait_b = pipe->mid;
    next.main_st = main_st;
    // step i's hook starts the gather of step i + depth (the first step's: of every step up to `depth`)
    for (int t = (i == 0 ? 1 : i + depth); t <= i + depth && t < n_steps; ++t) {
      const int tb = t % nb;  // last read by the dW kernel of step t - nb (never, when that is before this call)
      Job& j = next.job[next.n_jobs++];
      j.nx = a;
      patch(j.nx, steps[t]);
      j.x = xbuf[tb];
      j.l = lbuf[tb];
      j.wait_a = t < nb ? pipe->start : pipe->freed[tb];
      j.ready = pipe->ready[tb];
    }
    next.have = next.n_jobs > 0;
    StepHooks hooks;
    hooks.pregathered = true;
    hooks.operand_free = pipe->freed[b];
    GatherJob gjob;
    memset(&gjob, 0, sizeof(gjob));
    const bool merged = prefetch_placement() == 3 && !fuse_fix();
    if (merged) {
      if (next.have) {
        const uml_segment* g[2] = {nullptr, nullptr};
        int ng = 0;
        for (int k = 0; k < next.job[0].nx.nseg; ++k)
          if (next.job[0].nx.seg[k].n > 0) g[ng++] = &next.job[0].nx.seg[k];
        if (ng > 0) {
          gjob.s0 = CopySeg{reinterpret_cast<const unsigned char*>(g[0]->rows16), g[0]->idx, g[0]->labels, g[0]->n};
          if (ng > 1) gjob.s1 = CopySeg{reinterpret_cast<const unsigned char*>(g[1]->rows16), g[1]->idx, g[1]->labels, g[1]->n};
          gjob.vec_per_row = a.dim / 8;
          gjob.out = reinterpret_cast<uint4*>(next.job[0].x);
          gjob.out_pitch_vec = a.dim / 8;
          gjob.out_labels = next.job[0].l;
          gjob.blocks = 4 * sm_count();
          hooks.merged = &gjob;
        }
      }
      if (next.have) {  // (the merged copy has no launch of its own to bracket)
        rec(next.job[0].nx.ev[0], stream);
        rec(next.job[0].nx.ev[1], stream);
      }
      next.have = false;  // nothing for the side stream to do
    }
    hooks.mid_arg = &next;
    hooks.after_fwd_kernel = (prefetch_placement() == 2 && next.have) ? pipe->mid : nullptr;
    hooks.mid = [](void* p) -> int {
      Next* n = static_cast<Next*>(p);
      if (!n->have) return 0;
      if (n->record_mid) UML_CUDA(cudaEventRecord(n->pipe->mid, n->main_st));
      for (int k = 0; k < n->n_jobs; ++k) {
        Job& j = n->job[k];
        UML_CUDA(cudaStreamWaitEvent(n->pipe->aux, j.wait_a, 0));
        UML_CUDA(cudaStreamWaitEvent(n->pipe->aux, n->wait_b, 0));
        rec(j.nx.ev[0], n->pipe->aux);  // ev[0]..ev[1] bracket the gather where it really runs: on the side stream
        const int rc = shadow_gather(&j.nx, j.x, j.l, true, n->pipe->aux);
        if (rc) return rc;
        rec(j.nx.ev[1], n->pipe->aux);
        UML_CUDA(cudaEventRecord(j.ready, n->pipe->aux));
      }
      return 0;
    };
    const int rc = linear_step_impl(&a, stream, hooks);
    if (rc) return rc;
  }
  return 0;
}

int uml_linear_step(const uml_linear_step_args* a, void* stream) {
  pipe_invalidate();  // (it gathers into the first operand buffer on the main stream)
  return linear_step_impl(a, stream, StepHooks());
}

// the operand buffers were used by something the library did not see (Python-side kernels on them): start cold next time
int uml_linear_run_reset(void) {
  pipe_invalidate();
  return 0;
}

}  // extern "C"

static int linear_step_impl(const uml_linear_step_args* a, void* stream, const StepHooks& hooks) {
  using namespace uml;
  const bool pregathered = hooks.pregathered;
  const cudaEvent_t operand_free = hooks.operand_free;
  Pipe* opipe = nullptr;  // set when this step's dW runs on the second stream, overlapping the fix-up launch
  bool stats_in_dw = false;  // the forward's per-run statistics are reduced inside the dW kernel
  UML_REQUIRE(a != nullptr, "linear_step: null args");
  UML_REQUIRE(a->nseg >= 1 && a->nseg <= UML_MAX_SEGMENTS, "linear_step: 1..2 segments");
  UML_REQUIRE(a->W && a->G && a->row_loss && a->row_correct && a->stats, "linear_step: null buffers");
  const int64_t n0 = a->seg[0].n, n1 = a->nseg > 1 ? a->seg[1].n : 0, total = n0 + n1;
  const bool dp = a->dp_allreduce != 0;
  UML_REQUIRE(!dp || a->dW_out, "linear_step: data-parallel mode needs dW_out");
  const bool fused = a->dW_out == nullptr;
  int rc;

  if (a->precision == 0) {
    // ------------------------------------------------------------------ fp32 exact path (one fused launch, else 3 + 1)
    UML_REQUIRE(a->row_dscale, "linear_step: fp32 path needs row_dscale");
    rec(a->ev[2], stream);
    if (fused && !dp && total > 0) {  // the reference's batch sizes: forward, gradient and update in one cooperative launch
      int32_t launched = 0;
      rc = uml_head_step_fused_f32(a->seg, a->nseg, a->dim, a->W, a->n_classes, static_cast<float*>(a->G), a->ldg, a->row_loss,
                                   a->row_correct, a->row_dscale, a->stats, &a->upd, &launched, stream);
      if (rc) return rc;
      if (launched) {
        rec(a->ev[3], stream);
        rec(a->ev[4], stream);
        rec(a->ev[5], stream);
        return 0;
      }
    }
    rc = uml_head_fwd_ce_f32(a->seg, a->nseg, a->dim, a->W, a->n_classes, static_cast<float*>(a->G), a->ldg,
                             a->row_loss, a->row_correct, a->row_dscale, a->stats, a->g_capacity_rows, stream);
    if (rc) return rc;
    rec(a->ev[3], stream);
  } else {
    // ------------------------------------------------------------------ bf16 tensor-core path
    UML_REQUIRE(a->X16 && a->W16 && a->labels32 && a->partials && a->tile_ws && a->max_splits >= 1,
                "linear_step: bf16 buffers");
    if (!a->w16_valid) {
      rc = uml_cast_f32_to_bf16(a->W, a->W16, static_cast<int64_t>(a->n_classes) * a->dim, stream);
      if (rc) return rc;
    }
    uml_tc_segments ts;
    memset(&ts, 0, sizeof(ts));
    ts.nseg = 0;
    int64_t off = 0;
    if (!pregathered) rec(a->ev[0], stream);  // (a pre-gathered step's ev[0]..ev[1] were recorded around its gather)
    const bool shadow = pregathered || shadow_gatherable(a);
    if (shadow && !pregathered) {
      rc = shadow_gather(a, a->X16, a->labels32, false, stream);
      if (rc) return rc;
    }
    for (int i = 0; i < a->nseg; ++i) {
      const uml_segment& s = a->seg[i];
      if (s.n == 0) continue;
      if (!shadow) {
        const float* rows = static_cast<const float*>(s.rows);
        UML_REQUIRE(s.ld == a->dim, "linear_step: bf16 path needs dense bank rows (ld == dim)");
        if (s.idx && !s.label_idx) {
          rc = uml_gather_rows_labels_bf16(rows, s.labels, a->dim, s.idx, s.n, a->X16 + off * a->dim, a->dim,
                                           a->labels32 + off, stream);
          if (rc) return rc;
        } else {
          if (s.idx)
            rc = uml_gather_rows_bf16(rows, INT64_MAX / 2, a->dim, s.idx, s.n, a->X16 + off * a->dim, a->dim, stream);
          else
            rc = uml_cast_f32_to_bf16(rows, a->X16 + off * a->dim, s.n * a->dim, stream);
          if (rc) return rc;
          rc = uml_gather_labels_i32(s.labels, s.label_idx ? s.label_idx : s.idx, s.n, a->labels32 + off, stream);
          if (rc) return rc;
        }
      }
      ts.seg_rows[ts.nseg] = s.n;
      ts.scale[ts.nseg] = s.scale;
      ts.loss_weight[ts.nseg] = s.loss_weight;
      ts.scale_dev[ts.nseg] = s.scale_dev;
      ts.nseg++;
      off += s.n;
    }
    if (!pregathered) rec(a->ev[1], stream);
    if (hooks.mid && prefetch_placement() == 0) {
      rc = hooks.mid(hooks.mid_arg);
      if (rc) return rc;
    }
    if (total > 0) {
      // fix-up / dW overlap: counters zeroed before the forward kernel, dW launched on aux2 right after it
      int splits_o = uml_tc_dw_splits(total, a->dim, a->n_classes);
      if (splits_o > a->max_splits) splits_o = a->max_splits;
      opipe = (overlap_fixup() && !fuse_fix() && splits_o >= 1 && splits_o <= 8) ? get_pipe() : nullptr;
      FixupSignal sig;
      memset(&sig, 0, sizeof(sig));
      cudaEvent_t after_fwd = hooks.after_fwd_kernel;
      if (opipe) {
        UML_CUDA(cudaMemsetAsync(opipe->split_done, 0, 64, as_stream(stream)));
        sig.done = opipe->split_done;
        sig.n_splits = splits_o;
        sig.num_kb = (total + 63) / 64;
        if (!after_fwd) after_fwd = opipe->fwd_done;
      }
      rec(a->ev[2], stream);
      // ev[2]..ev[3] bracket the tensor-core kernel alone.  Exchange kernel (default): G is final when it ends and the
      // per-run statistics are reduced by an idle warp of the dW GEMM - unless a learnable temperature needs them
      // before that (then a small reduction launch follows the forward).  Chunk-sequential kernel (UML_FWD_IMPL=old):
      // the fix-up launch, which also reduces the statistics, follows it.
      stats_in_dw = !fuse_fix() && !opipe && uml_fwd_x_eligible(total, a->n_classes) && !a->scale_param[0] &&
                    !a->scale_param[1];
      rc = uml_head_fwd_ce_bf16_ev(a->X16, total, a->dim, a->W16, a->n_classes, a->labels32, &ts,
                                   static_cast<uint16_t*>(a->G), a->ldg, nullptr, nullptr, nullptr, nullptr, a->tile_ws,
                                   stats_in_dw ? nullptr : a->stats, a->ev[3], stream, fuse_fix() ? 1 : 0, after_fwd,
                                   fuse_fix() ? nullptr : hooks.merged, opipe ? &sig : nullptr);
      if (rc) return rc;
      if (opipe) {
        // the GEMM goes to the second stream behind "forward kernel done"; the fix-up launch just enqueued on the main
        // stream runs concurrently and releases the GEMM's splits one by one
        cudaStream_t s2 = opipe->aux2;
        UML_CUDA(cudaStreamWaitEvent(s2, after_fwd, 0));
        rec(a->ev[4], s2);
        rc = uml_head_bwd_dw_gated_bf16(static_cast<const uint16_t*>(a->G), a->ldg, a->X16, total, a->dim, a->n_classes,
                                        a->partials, splits_o, opipe->split_done, reinterpret_cast<int*>(opipe->split_done + 8), s2);
        if (rc) return rc;
        rec(a->ev[5], s2);
        UML_CUDA(cudaEventRecord(opipe->dw_done, s2));
      }
    } else if (hooks.merged) {
      // no rows on this rank this step, hence no fix-up launch to ride in: copy the next step's rows directly
      const GatherJob& j = *hooks.merged;
      rc = uml_gather2_rows_bf16_light(reinterpret_cast<const uint16_t*>(j.s0.bank), j.s0.labels, j.s0.idx, j.s0.n,
                                       reinterpret_cast<const uint16_t*>(j.s1.bank), j.s1.labels, j.s1.idx, j.s1.n, a->dim,
                                       reinterpret_cast<uint16_t*>(j.out), a->dim, j.out_labels, stream);
      if (rc) return rc;
    }
    if (hooks.mid && prefetch_placement() != 0 && prefetch_placement() != 4) {
      rc = hooks.mid(hooks.mid_arg);
      if (rc) return rc;
    }
  }

  // learnable temperatures: scalar Adam(W) steps fed straight from the stats record on the device
  for (int i = 0; i < a->nseg; ++i) {
    if (!a->scale_param[i]) continue;
    float* g = &a->stats[i].dscale;
    if (dp) {  // the temperature gradient is a sum over the global batch as well
      rc = uml_dp_allreduce_f32(g, 1, stream);
      if (rc) return rc;
    }
    if (a->upd.kind == 3)
      rc = uml_sgd_step(a->scale_param[i], g, nullptr, 0.f, a->scale_m[i], 1, a->upd.lr, a->upd.momentum,
                        a->upd.weight_decay, a->scale_step[i], nullptr, stream);
    else
      rc = uml_adamw_step(a->scale_param[i], g, nullptr, 0.f, a->scale_m[i], a->scale_v[i], 1, a->upd.lr, a->upd.beta1,
                          a->upd.beta2, a->upd.eps, a->upd.weight_decay, a->scale_step[i], a->upd.kind == 1, nullptr,
                          stream);
    if (rc) return rc;
  }

  const int64_t np_all = static_cast<int64_t>(a->n_classes) * a->dim;
  if (total == 0 && a->precision == 1 && hooks.mid && prefetch_placement() == 4) {  // (no dW launch to follow on this rank)
    rc = hooks.mid(hooks.mid_arg);
    if (rc) return rc;
  }
  if (total == 0 && !dp) return 0;
  if (total == 0) {  // a rank without rows still takes part in the all-reduce, contributing zeros
    UML_CUDA(cudaMemsetAsync(dp_local_sum_target(a, np_all), 0, np_all * sizeof(float), as_stream(stream)));
    return dp_reduce_and_update(a, np_all, stream);
  }
  if (a->precision == 0) {
    uml_update none;
    memset(&none, 0, sizeof(none));
    rec(a->ev[4], stream);
    rc = uml_head_bwd_dw_f32(a->seg, a->nseg, a->dim, static_cast<const float*>(a->G), a->ldg, a->n_classes, a->W,
                             dp ? dp_local_sum_target(a, np_all) : a->dW_out, fused ? &a->upd : &none, stream);
    rec(a->ev[5], stream);
    if (rc || !dp) return rc;
    return dp_reduce_and_update(a, np_all, stream);
  }
  int splits = uml_tc_dw_splits(total, a->dim, a->n_classes);
  if (splits > a->max_splits) splits = a->max_splits;
  const bool update_in_dw = fused && !dp && a->upd.kind != 3 && a->W16 && splits <= 8 && a->dim % 4 == 0 && update_in_dw_enabled();
  if (opipe) {
    UML_CUDA(cudaStreamWaitEvent(as_stream(stream), opipe->dw_done, 0));  // the GEMM was launched with the forward
    rc = 0;
  } else {
  rec(a->ev[4], stream);
  if (fuse_fix()) {
    // the forward skipped its fix-up pass: the dW prologue normalises every G stage in shared memory
    uml_tc_segments ts2;
    memset(&ts2, 0, sizeof(ts2));
    for (int i = 0; i < a->nseg; ++i) {
      const uml_segment& sg = a->seg[i];
      if (sg.n == 0) continue;
      ts2.seg_rows[ts2.nseg] = sg.n;
      ts2.scale[ts2.nseg] = sg.scale;
      ts2.loss_weight[ts2.nseg] = sg.loss_weight;
      ts2.scale_dev[ts2.nseg] = sg.scale_dev;
      ts2.nseg++;
    }
    rc = uml_head_bwd_dw_fix_bf16(static_cast<const uint16_t*>(a->G), a->ldg, a->X16, total, a->dim, a->n_classes,
                                  a->partials, splits, &ts2, a->labels32, a->tile_ws, a->stats, stream);
  } else if (stats_in_dw && update_in_dw) {
    // single GPU, AdamW: statistics, split-K sum and the update all run inside the dW launch (no update launch)
    const float* part = nullptr;
    int64_t entries = 0;
    uml_fwd_x_partials(a->tile_ws, total, a->n_classes, &part, &entries);
    int nseg_live = 0;
    for (int i = 0; i < a->nseg; ++i) nseg_live += a->seg[i].n > 0 ? 1 : 0;
    const uml::AdamArgs adam = uml::make_adam(a->upd.lr, a->upd.beta1, a->upd.beta2, a->upd.eps, a->upd.weight_decay, a->upd.step,
                                              a->upd.kind == 1);
    rc = uml_head_bwd_dw_update_bf16(static_cast<const uint16_t*>(a->G), a->ldg, a->X16, total, a->dim, a->n_classes,
                                     a->partials, splits, part, entries, nseg_live, a->stats, a->W, a->upd.m, a->upd.v, a->W16,
                                     &adam, reinterpret_cast<unsigned*>(a->tile_ws) + 2, stream);
  } else if (stats_in_dw) {
    const float* part = nullptr;
    int64_t entries = 0;
    uml_fwd_x_partials(a->tile_ws, total, a->n_classes, &part, &entries);
    int nseg_live = 0;
    for (int i = 0; i < a->nseg; ++i) nseg_live += a->seg[i].n > 0 ? 1 : 0;
    rc = uml_head_bwd_dw_stats_bf16(static_cast<const uint16_t*>(a->G), a->ldg, a->X16, total, a->dim, a->n_classes,
                                    a->partials, splits, part, entries, nseg_live, a->stats, stream);
  } else {
    rc = uml_head_bwd_dw_bf16(static_cast<const uint16_t*>(a->G), a->ldg, a->X16, total, a->dim, a->n_classes, a->partials,
                              splits, stream);
  }
  if (rc) return rc;
  rec(a->ev[5], stream);
  }
  if (operand_free) UML_CUDA(cudaEventRecord(operand_free, as_stream(stream)));  // X16 / labels32 may be overwritten
  if (a->precision == 1 && hooks.mid && prefetch_placement() == 4) {  // the next step's gather starts when dW has finished
    rc = hooks.mid(hooks.mid_arg);
    if (rc) return rc;
  }
  const int64_t np = static_cast<int64_t>(a->n_classes) * a->dim;
  if (stats_in_dw && update_in_dw) {  // (the events of the update bracket nothing: it ran inside the dW kernel)
    rec(a->ev[6], stream);
    rec(a->ev[7], stream);
    return 0;
  }
  if (!fused) {
    if (dp && a->upd.kind != 3 && uml_dp_p2p_input(np)) {
      // data parallel over NVLink peer memory: split-K sum + exchange + Adam in ONE kernel per rank
      rec(a->ev[6], stream);
      rc = uml_dp_fused_adam_update(a->partials, splits, np, np, a->W, a->upd.m, a->upd.v, a->upd.lr, a->upd.beta1,
                                    a->upd.beta2, a->upd.eps, a->upd.weight_decay, a->upd.step, a->upd.kind == 1, a->W16,
                                    stream);
      rec(a->ev[7], stream);
      return rc;
    }
    rc = uml_sum_partials(a->partials, splits, np, np, dp ? dp_local_sum_target(a, np) : a->dW_out, stream);
    if (rc || !dp) return rc;
    return dp_reduce_and_update(a, np, stream);
  }
  if (a->upd.kind == 3) {
    UML_REQUIRE(a->dW_scratch, "linear_step: SGD on the bf16 path needs dW_scratch");
    rc = uml_sum_partials(a->partials, splits, np, np, a->dW_scratch, stream);
    if (rc) return rc;
    return uml_sgd_step(a->W, a->dW_scratch, nullptr, 0.f, a->upd.m, np, a->upd.lr, a->upd.momentum, a->upd.weight_decay,
                        a->upd.step, a->W16, stream);
  }
  rec(a->ev[6], stream);
  rc = uml_adamw_step_partials(a->W, a->partials, splits, np, a->upd.m, a->upd.v, np, a->upd.lr, a->upd.beta1,
                               a->upd.beta2, a->upd.eps, a->upd.weight_decay, a->upd.step, a->upd.kind == 1, a->W16,
                               nullptr, stream);
  rec(a->ev[7], stream);
  return rc;
}
