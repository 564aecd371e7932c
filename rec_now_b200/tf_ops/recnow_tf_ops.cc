// TensorFlow 2 custom-op shim over the C ABI of librecnow_b200.so (include/recnow_b200.h).
//
// STATUS: source only.  TensorFlow is not installed in the build image (and cannot be: no network), so this file
// has never been compiled or loaded; it is the binding a maintainer adds on a box that has TensorFlow
// (build: python -m rec_now_b200.tf_ops.build).  All logic lives behind the C ABI, which IS tested (ctypes).
//
// Ops (registered for DEVICE_GPU; there is no CPU kernel on purpose):
//   RecNowCanonKeys      float/double group ids -> canonical int64 keys + row_ok        (PW:33-35 equality)
//   RecNowPairwiseLoss   pairwise_loss with bpr_loss_func, forward + d loss/d logits     (PW:228-279, PW:96-127)
//   RecNowPairIndices    kept pairs in the reference's row-major order                  (PW:206-217)
//   RecNowListwiseLoss   to_listwise_sample + softmax-CE, forward + d loss/d logits      (LW:89-173)
//   RecNowListwiseDense  the (V,B) dense_mask / dense_labels / dense_logits outputs      (LW:142-145)
// Gradients are registered from Python (rec_now_b200/tf_ops/__init__.py) as upstream * saved dlogits.
//
// TF conventions honoured: Compute() keeps no state (re-entrant across inter-op threads); every buffer is
// TF-owned (allocate_output / allocate_temp, 256-byte aligned); work is enqueued on TF's stream and never
// synchronised, except the unavoidable count -> allocate step of the two data-dependent-shape ops; errors go
// through OP_REQUIRES, never exceptions or aborts.
#define EIGEN_USE_GPU
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"

#include <cuda_runtime.h>

#include "recnow_b200.h"

namespace tf = tensorflow;
using tf::OpKernel; using tf::OpKernelConstruction; using tf::OpKernelContext; using tf::Tensor; using tf::TensorShape;
using tf::errors::InvalidArgument; using tf::errors::Internal;

namespace {

inline void* Stream(OpKernelContext* ctx) { return static_cast<void*>(ctx->eigen_device<Eigen::GpuDevice>().stream()); }

template <typename T> const T* OptPtr(const Tensor& t) { return t.NumElements() ? t.flat<T>().data() : nullptr; }

tf::Status RnStatus(int code, const char* where) {
  if (code == RN_OK) return tf::OkStatus();
  if (code == RN_ERR_LAUNCH || code == RN_ERR_INTERNAL) return Internal(where, ": ", rn_strerror(code));
  return InvalidArgument(where, ": ", rn_strerror(code));
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------
REGISTER_OP("RecNowCanonKeys")
    .Input("ids: T").Input("row_ok_in: uint8")
    .Output("keys: int64").Output("row_ok: uint8")
    .Attr("T: {float, double}")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      c->set_output(0, c->Vector(c->NumElements(c->input(0))));
      c->set_output(1, c->Vector(c->NumElements(c->input(0))));
      return tf::OkStatus();
    });

template <typename T>
class RecNowCanonKeysOp : public OpKernel {
 public:
  explicit RecNowCanonKeysOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor& ids = ctx->input(0);
    const Tensor& ok_in = ctx->input(1);
    const int64_t B = ids.NumElements();
    OP_REQUIRES(ctx, B > 0, InvalidArgument("RecNowCanonKeys: empty batch"));
    OP_REQUIRES(ctx, ok_in.NumElements() == 0 || ok_in.NumElements() == B, InvalidArgument("row_ok_in size"));
    Tensor *keys = nullptr, *ok = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({B}), &keys));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, TensorShape({B}), &ok));
    uint8_t* okp = ok->flat<uint8_t>().data();
    int and_into = 0;
    if (ok_in.NumElements()) {
      OP_REQUIRES(ctx, cudaMemcpyAsync(okp, ok_in.flat<uint8_t>().data(), B, cudaMemcpyDeviceToDevice,
                                       static_cast<cudaStream_t>(Stream(ctx))) == cudaSuccess, Internal("copy row_ok"));
      and_into = 1;
    }
    int rc;
    if (std::is_same<T, float>::value)
      rc = rn_canon_keys_f32(reinterpret_cast<const float*>(ids.flat<T>().data()), B, keys->flat<tf::int64>().data(), okp, and_into, Stream(ctx));
    else
      rc = rn_canon_keys_f64(reinterpret_cast<const double*>(ids.flat<T>().data()), B, keys->flat<tf::int64>().data(), okp, and_into, Stream(ctx));
    OP_REQUIRES_OK(ctx, RnStatus(rc, "rn_canon_keys"));
  }
};
REGISTER_KERNEL_BUILDER(Name("RecNowCanonKeys").Device(tf::DEVICE_GPU).TypeConstraint<float>("T"), RecNowCanonKeysOp<float>);
REGISTER_KERNEL_BUILDER(Name("RecNowCanonKeys").Device(tf::DEVICE_GPU).TypeConstraint<double>("T"), RecNowCanonKeysOp<double>);

// ---------------------------------------------------------------------------------------------------------
REGISTER_OP("RecNowPairwiseLoss")
    .Input("logits: float").Input("labels: float").Input("group_keys: int64")      // keys: [K, B]
    .Input("row_ok: uint8").Input("rw_pos: float").Input("rw_neg: float")          // empty tensor = absent
    .Input("weight_lut: float")                                                    // [8, 8] level table (label_func = RN_LABEL_LUT) or empty
    .Attr("label_func: int = 0").Attr("factor: float = 1.0").Attr("power: float = 0.0")
    .Attr("only_wrong: bool = false").Attr("reduce_mean: bool = true")
    .Output("loss: float").Output("n_pair: float").Output("n_pair_i64: int64").Output("dlogits: float")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      c->set_output(0, c->Scalar()); c->set_output(1, c->Scalar()); c->set_output(2, c->Scalar());
      c->set_output(3, c->Vector(c->NumElements(c->input(0))));
      return tf::OkStatus();
    });

struct PairAttrs {
  int label_func; float factor, power; bool only_wrong, reduce_mean;
  explicit PairAttrs(OpKernelConstruction* c) {
    OP_REQUIRES_OK(c, c->GetAttr("label_func", &label_func));
    OP_REQUIRES_OK(c, c->GetAttr("factor", &factor));
    OP_REQUIRES_OK(c, c->GetAttr("power", &power));
    OP_REQUIRES_OK(c, c->GetAttr("only_wrong", &only_wrong));
    reduce_mean = true;
    if (c->HasAttr("reduce_mean")) OP_REQUIRES_OK(c, c->GetAttr("reduce_mean", &reduce_mean));
  }
};

static tf::Status FillPairArgs(OpKernelContext* ctx, const PairAttrs& at, rn_pairwise_args* a) {
  const Tensor &s = ctx->input(0), &y = ctx->input(1), &k = ctx->input(2), &ok = ctx->input(3), &wp = ctx->input(4),
               &wn = ctx->input(5);
  const int64_t B = s.NumElements();
  if (B <= 0 || y.NumElements() != B || k.NumElements() % B != 0 || k.NumElements() == 0)
    return InvalidArgument("logits / labels / group_keys sizes disagree");
  for (const Tensor* t : {&ok, &wp, &wn})
    if (t->NumElements() != 0 && t->NumElements() != B) return InvalidArgument("optional per-row input has the wrong size");
  *a = rn_pairwise_args{};
  a->B = B; a->K = static_cast<int32_t>(k.NumElements() / B); a->label_func = at.label_func;
  a->keys = reinterpret_cast<const int64_t*>(k.flat<tf::int64>().data());
  a->logits = s.flat<float>().data(); a->labels = y.flat<float>().data();
  a->row_ok = OptPtr<uint8_t>(ok); a->rw_pos = OptPtr<float>(wp); a->rw_neg = OptPtr<float>(wn);
  a->factor = at.factor; a->power = at.power; a->only_wrong = at.only_wrong; a->reduce_mean = at.reduce_mean;
  a->part_rank = 0; a->part_count = 1;
  return tf::OkStatus();
}

class RecNowPairwiseLossOp : public OpKernel {
 public:
  explicit RecNowPairwiseLossOp(OpKernelConstruction* c) : OpKernel(c), at_(c) {}
  void Compute(OpKernelContext* ctx) override {
    rn_pairwise_args a;
    OP_REQUIRES_OK(ctx, FillPairArgs(ctx, at_, &a));
    {
      // SURVEY 8b weight_lut: any label-only label_pair_to_weight_func as a table over the label levels (recnow_b200.h)
      const Tensor& lut = ctx->input(6);
      OP_REQUIRES(ctx, lut.NumElements() == 0 || lut.NumElements() == 64, InvalidArgument("weight_lut must be empty or hold 8 x 8 floats"));
      a.weight_lut = OptPtr<float>(lut);
    }
    Tensor *loss = nullptr, *n = nullptr, *ni = nullptr, *d = nullptr, scratch;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({}), &loss));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, TensorShape({}), &n));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, TensorShape({}), &ni));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(3, TensorShape({a.B}), &d));
    const size_t bytes = rn_pairwise_scratch_bytes(a.B, a.K);
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_UINT8, TensorShape({static_cast<int64_t>(bytes)}), &scratch));
    a.loss = loss->scalar<float>().data(); a.n_pair_f32 = n->scalar<float>().data();
    a.n_pair = reinterpret_cast<int64_t*>(ni->scalar<tf::int64>().data()); a.dlogits = d->flat<float>().data();
    OP_REQUIRES_OK(ctx, RnStatus(rn_pairwise_fwd_bwd(&a, scratch.flat<uint8_t>().data(), bytes, Stream(ctx)), "rn_pairwise_fwd_bwd"));
  }
 private:
  PairAttrs at_;
};
REGISTER_KERNEL_BUILDER(Name("RecNowPairwiseLoss").Device(tf::DEVICE_GPU), RecNowPairwiseLossOp);

// ---------------------------------------------------------------------------------------------------------
REGISTER_OP("RecNowPairIndices")
    .Input("logits: float").Input("labels: float").Input("group_keys: int64")
    .Input("row_ok: uint8").Input("rw_pos: float").Input("rw_neg: float")
    .Attr("label_func: int = 0").Attr("factor: float = 1.0").Attr("power: float = 0.0")
    .Attr("only_wrong: bool = false").Attr("label_cond: bool = true")
    .Output("pos_idx: int32").Output("neg_idx: int32").Output("weights: float")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      for (int i = 0; i < 3; ++i) c->set_output(i, c->Vector(c->UnknownDim()));
      return tf::OkStatus();
    });

class RecNowPairIndicesOp : public OpKernel {
 public:
  explicit RecNowPairIndicesOp(OpKernelConstruction* c) : OpKernel(c), at_(c) {
    OP_REQUIRES_OK(c, c->GetAttr("label_cond", &label_cond_));
  }
  void Compute(OpKernelContext* ctx) override {
    rn_pairwise_args a;
    OP_REQUIRES_OK(ctx, FillPairArgs(ctx, at_, &a));
    Tensor scratch;
    const size_t bytes = rn_pair_indices_scratch_bytes(a.B, a.K);
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_UINT8, TensorShape({static_cast<int64_t>(bytes)}), &scratch));
    int64_t P = 0;   // data-dependent output size: the one place this shim synchronises the stream
    OP_REQUIRES_OK(ctx, RnStatus(rn_pair_indices_count(&a, label_cond_, scratch.flat<uint8_t>().data(), bytes, &P, Stream(ctx)),
                                 "rn_pair_indices_count"));
    Tensor *pi = nullptr, *ni = nullptr, *w = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({P}), &pi));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, TensorShape({P}), &ni));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, TensorShape({P}), &w));
    if (P == 0) return;
    OP_REQUIRES_OK(ctx, RnStatus(rn_pair_indices_fill(&a, label_cond_, scratch.flat<uint8_t>().data(), bytes,
                                                      pi->flat<tf::int32>().data(), ni->flat<tf::int32>().data(),
                                                      w->flat<float>().data(), P, Stream(ctx)), "rn_pair_indices_fill"));
  }
 private:
  PairAttrs at_;
  bool label_cond_;
};
REGISTER_KERNEL_BUILDER(Name("RecNowPairIndices").Device(tf::DEVICE_GPU), RecNowPairIndicesOp);

// ---------------------------------------------------------------------------------------------------------
REGISTER_OP("RecNowListwiseLoss")
    .Input("group_keys: int64").Input("row_ok: uint8").Input("labels: float").Input("logits: float").Input("list_w: float")
    .Attr("pos_neg_th: float = 0.5").Attr("do_reduce: bool = true")
    .Output("loss: float").Output("list_loss: float").Output("n_valid: int32").Output("n_group: int32").Output("dlogits: float")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      auto b = c->Vector(c->NumElements(c->input(3)));
      c->set_output(0, c->Scalar()); c->set_output(1, b); c->set_output(2, c->Scalar()); c->set_output(3, c->Scalar());
      c->set_output(4, b);
      return tf::OkStatus();
    });

static tf::Status FillListArgs(OpKernelContext* ctx, float th, bool do_reduce, rn_listwise_args* a) {
  const Tensor &k = ctx->input(0), &ok = ctx->input(1), &y = ctx->input(2), &s = ctx->input(3), &lw = ctx->input(4);
  const int64_t B = s.NumElements();
  if (B <= 0 || y.NumElements() != B || k.NumElements() != B) return InvalidArgument("group_keys / labels / logits sizes disagree");
  if (ok.NumElements() != 0 && ok.NumElements() != B) return InvalidArgument("row_ok size");
  *a = rn_listwise_args{};
  a->B = B; a->keys = reinterpret_cast<const int64_t*>(k.flat<tf::int64>().data()); a->row_ok = OptPtr<uint8_t>(ok);
  a->labels = y.flat<float>().data(); a->logits = s.flat<float>().data(); a->list_w = OptPtr<float>(lw);
  a->pos_neg_th = th; a->do_reduce = do_reduce;
  return tf::OkStatus();
}

class RecNowListwiseLossOp : public OpKernel {
 public:
  explicit RecNowListwiseLossOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("pos_neg_th", &th_));
    OP_REQUIRES_OK(c, c->GetAttr("do_reduce", &do_reduce_));
  }
  void Compute(OpKernelContext* ctx) override {
    rn_listwise_args a;
    OP_REQUIRES_OK(ctx, FillListArgs(ctx, th_, do_reduce_, &a));
    Tensor *loss = nullptr, *ll = nullptr, *nv = nullptr, *ng = nullptr, *d = nullptr, scratch;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({}), &loss));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, TensorShape({a.B}), &ll));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, TensorShape({}), &nv));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(3, TensorShape({}), &ng));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(4, TensorShape({a.B}), &d));
    const size_t bytes = rn_listwise_scratch_bytes(a.B);
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_UINT8, TensorShape({static_cast<int64_t>(bytes)}), &scratch));
    a.loss = loss->scalar<float>().data(); a.list_loss = ll->flat<float>().data();
    a.n_valid = nv->scalar<tf::int32>().data(); a.n_group = ng->scalar<tf::int32>().data(); a.dlogits = d->flat<float>().data();
    OP_REQUIRES(ctx, cudaMemsetAsync(a.loss, 0, sizeof(float), static_cast<cudaStream_t>(Stream(ctx))) == cudaSuccess, Internal("memset"));
    OP_REQUIRES_OK(ctx, RnStatus(rn_listwise_fwd_bwd(&a, scratch.flat<uint8_t>().data(), bytes, Stream(ctx)), "rn_listwise_fwd_bwd"));
  }
 private:
  float th_; bool do_reduce_;
};
REGISTER_KERNEL_BUILDER(Name("RecNowListwiseLoss").Device(tf::DEVICE_GPU), RecNowListwiseLossOp);

// ---------------------------------------------------------------------------------------------------------
REGISTER_OP("RecNowListwiseDense")
    .Input("group_keys: int64").Input("row_ok: uint8").Input("labels: float").Input("logits: float")
    .Attr("pos_neg_th: float = 0.5").Attr("do_mask_logits: bool = true").Attr("value_of_masked_logit: float = -1e9")
    .Output("dense_mask: bool").Output("dense_labels: float").Output("dense_logits: float")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      auto s = c->Matrix(c->UnknownDim(), c->NumElements(c->input(3)));
      for (int i = 0; i < 3; ++i) c->set_output(i, s);
      return tf::OkStatus();
    });

class RecNowListwiseDenseOp : public OpKernel {
 public:
  explicit RecNowListwiseDenseOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("pos_neg_th", &th_));
    OP_REQUIRES_OK(c, c->GetAttr("do_mask_logits", &mask_));
    OP_REQUIRES_OK(c, c->GetAttr("value_of_masked_logit", &pad_));
  }
  void Compute(OpKernelContext* ctx) override {
    const Tensor &k = ctx->input(0), &ok = ctx->input(1), &y = ctx->input(2), &s = ctx->input(3);
    const int64_t B = s.NumElements();
    OP_REQUIRES(ctx, B > 0 && y.NumElements() == B && k.NumElements() == B, InvalidArgument("sizes disagree"));
    rn_listwise_args a{};
    a.B = B; a.keys = reinterpret_cast<const int64_t*>(k.flat<tf::int64>().data()); a.row_ok = OptPtr<uint8_t>(ok);
    a.labels = y.flat<float>().data(); a.logits = s.flat<float>().data(); a.pos_neg_th = th_; a.do_reduce = 1;
    Tensor scratch, tmp_f, tmp_i;
    const size_t bytes = rn_listwise_scratch_bytes(B);
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_UINT8, TensorShape({static_cast<int64_t>(bytes)}), &scratch));
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_FLOAT, TensorShape({B + 4}), &tmp_f));
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_INT32, TensorShape({8}), &tmp_i));
    a.loss = tmp_f.flat<float>().data(); a.dlogits = tmp_f.flat<float>().data() + 4;
    a.n_valid = tmp_i.flat<tf::int32>().data(); a.n_group = tmp_i.flat<tf::int32>().data() + 4;
    cudaStream_t st = static_cast<cudaStream_t>(Stream(ctx));
    OP_REQUIRES_OK(ctx, RnStatus(rn_listwise_fwd_bwd(&a, scratch.flat<uint8_t>().data(), bytes, st), "rn_listwise_fwd_bwd"));
    int32_t V = 0;   // data-dependent output shape: one stream synchronisation
    OP_REQUIRES(ctx, cudaMemcpyAsync(&V, a.n_valid, sizeof(V), cudaMemcpyDeviceToHost, st) == cudaSuccess &&
                         cudaStreamSynchronize(st) == cudaSuccess, Internal("read n_valid"));
    Tensor *dm = nullptr, *dl = nullptr, *dz = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({V, B}), &dm));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, TensorShape({V, B}), &dl));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, TensorShape({V, B}), &dz));
    if (V == 0) return;
    OP_REQUIRES_OK(ctx, RnStatus(rn_listwise_dense(&a, scratch.flat<uint8_t>().data(), bytes, V,
                                                   reinterpret_cast<uint8_t*>(dm->flat<bool>().data()), dl->flat<float>().data(),
                                                   dz->flat<float>().data(), mask_, pad_, st), "rn_listwise_dense"));
  }
 private:
  float th_, pad_; bool mask_;
};
REGISTER_KERNEL_BUILDER(Name("RecNowListwiseDense").Device(tf::DEVICE_GPU), RecNowListwiseDenseOp);
