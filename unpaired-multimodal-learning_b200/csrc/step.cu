// One UML iteration behind ONE C-ABI call: the host enqueues every kernel of the step
// (reference loop body, vision_language/finetune.py:163-195) without returning to Python in between.
// At the reference's batch sizes the step is launch-latency bound, at throughput batch sizes the host
// must stay ahead of a ~150 us GPU step - either way per-kernel Python dispatch is what limits it.
#include "common.cuh"

namespace {
inline void rec(void* ev, void* stream) {
  if (ev) cudaEventRecord(static_cast<cudaEvent_t>(ev), uml::as_stream(stream));
}
}  // namespace

// data parallel tail of a step: sum dW over the ranks, then the optimizer update on every rank
static int dp_reduce_and_update(const uml_linear_step_args* a, int64_t np, void* stream) {
  int rc = uml_dp_allreduce_f32(a->dW_out, np, stream);
  if (rc) return rc;
  uint16_t* shadow = a->precision == 1 ? a->W16 : nullptr;
  rec(a->ev[6], stream);
  if (a->upd.kind == 3)
    rc = uml_sgd_step(a->W, a->dW_out, nullptr, 0.f, a->upd.m, np, a->upd.lr, a->upd.momentum, a->upd.weight_decay,
                      a->upd.step, shadow, stream);
  else
    rc = uml_adamw_step(a->W, a->dW_out, nullptr, 0.f, a->upd.m, a->upd.v, np, a->upd.lr, a->upd.beta1, a->upd.beta2,
                        a->upd.eps, a->upd.weight_decay, a->upd.step, a->upd.kind == 1, shadow, stream);
  rec(a->ev[7], stream);
  return rc;
}

extern "C" {

int uml_linear_run(const uml_linear_step_args* base, const uml_run_step* steps, int32_t n_steps, void* stream) {
  using namespace uml;
  UML_REQUIRE(base && steps && n_steps >= 0, "linear_run: bad arguments");
  uml_linear_step_args a = *base;
  for (int i = 0; i < n_steps; ++i) {
    const uml_run_step& s = steps[i];
    for (int k = 0; k < a.nseg; ++k) {
      a.seg[k].idx = s.idx[k];
      a.seg[k].n = s.n[k];
      a.seg[k].loss_weight = s.loss_weight[k];
      a.scale_step[k] = s.scale_step[k];
    }
    a.upd.lr = s.lr;
    a.upd.step = s.opt_step;
    a.stats = s.stats;
    a.ev[2] = s.ev_fwd[0];
    a.ev[3] = s.ev_fwd[1];
    if (i > 0 && a.precision == 1) a.w16_valid = 1;  // the optimizer kernel of the previous step refreshed the shadow
    const int rc = uml_linear_step(&a, stream);
    if (rc) return rc;
  }
  return 0;
}

int uml_linear_step(const uml_linear_step_args* a, void* stream) {
  using namespace uml;
  UML_REQUIRE(a != nullptr, "linear_step: null args");
  UML_REQUIRE(a->nseg >= 1 && a->nseg <= UML_MAX_SEGMENTS, "linear_step: 1..2 segments");
  UML_REQUIRE(a->W && a->G && a->row_loss && a->row_correct && a->stats, "linear_step: null buffers");
  const int64_t n0 = a->seg[0].n, n1 = a->nseg > 1 ? a->seg[1].n : 0, total = n0 + n1;
  const bool dp = a->dp_allreduce != 0;
  UML_REQUIRE(!dp || a->dW_out, "linear_step: data-parallel mode needs dW_out");
  const bool fused = a->dW_out == nullptr;
  int rc;

  if (a->precision == 0) {
    // ------------------------------------------------------------------ fp32 exact path (3 launches)
    UML_REQUIRE(a->row_dscale, "linear_step: fp32 path needs row_dscale");
    rec(a->ev[2], stream);
    rc = uml_head_fwd_ce_f32(a->seg, a->nseg, a->dim, a->W, a->n_classes, static_cast<float*>(a->G), a->ldg,
                             a->row_loss, a->row_correct, a->row_dscale, a->stats, stream);
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
    rec(a->ev[0], stream);
    bool shadow_gather = true;  // every non-empty run has a bf16 shadow bank and gather indices
    for (int i = 0; i < a->nseg; ++i) {
      const uml_segment& s = a->seg[i];
      if (s.n > 0 && !(s.rows16 && s.idx && !s.label_idx)) shadow_gather = false;
    }
    if (shadow_gather) {
      // one launch of the TMA copy kernel for both runs (rows + labels)
      const uml_segment* g[2] = {nullptr, nullptr};
      int ng = 0;
      for (int i = 0; i < a->nseg; ++i)
        if (a->seg[i].n > 0) g[ng++] = &a->seg[i];
      if (ng > 0) {
        rc = uml_gather2_rows_bf16(g[0]->rows16, g[0]->labels, g[0]->idx, g[0]->n, ng > 1 ? g[1]->rows16 : nullptr,
                                   ng > 1 ? g[1]->labels : nullptr, ng > 1 ? g[1]->idx : nullptr, ng > 1 ? g[1]->n : 0,
                                   a->dim, a->X16, a->dim, a->labels32, stream);
        if (rc) return rc;
      }
    }
    for (int i = 0; i < a->nseg; ++i) {
      const uml_segment& s = a->seg[i];
      if (s.n == 0) continue;
      if (!shadow_gather) {
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
    rec(a->ev[1], stream);
    if (total > 0) {
      rec(a->ev[2], stream);
      rc = uml_head_fwd_ce_bf16(a->X16, total, a->dim, a->W16, a->n_classes, a->labels32, &ts,
                                static_cast<uint16_t*>(a->G), a->ldg, nullptr, nullptr, nullptr, nullptr, a->tile_ws,
                                a->stats, stream);  // the fix-up launch also reduces the per-run statistics
      if (rc) return rc;
      rec(a->ev[3], stream);
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
  if (total == 0 && !dp) return 0;
  if (total == 0) {  // a rank without rows still takes part in the all-reduce, contributing zeros
    UML_CUDA(cudaMemsetAsync(a->dW_out, 0, np_all * sizeof(float), as_stream(stream)));
    return dp_reduce_and_update(a, np_all, stream);
  }
  if (a->precision == 0) {
    uml_update none;
    memset(&none, 0, sizeof(none));
    rec(a->ev[4], stream);
    rc = uml_head_bwd_dw_f32(a->seg, a->nseg, a->dim, static_cast<const float*>(a->G), a->ldg, a->n_classes, a->W,
                             a->dW_out, fused ? &a->upd : &none, stream);
    rec(a->ev[5], stream);
    if (rc || !dp) return rc;
    return dp_reduce_and_update(a, np_all, stream);
  }
  int splits = uml_tc_dw_splits(total, a->dim, a->n_classes);
  if (splits > a->max_splits) splits = a->max_splits;
  rec(a->ev[4], stream);
  rc = uml_head_bwd_dw_bf16(static_cast<const uint16_t*>(a->G), a->ldg, a->X16, total, a->dim, a->n_classes, a->partials,
                            splits, stream);
  if (rc) return rc;
  rec(a->ev[5], stream);
  const int64_t np = static_cast<int64_t>(a->n_classes) * a->dim;
  if (!fused) {
    rc = uml_sum_partials(a->partials, splits, np, np, a->dW_out, stream);
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

}  // extern "C"
