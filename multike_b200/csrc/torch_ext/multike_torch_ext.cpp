// PyTorch C++ extension over the C-ABI (BASELINE north_star: "the training step binds through a PyTorch C++/CUDA
// extension (thin C-ABI) to hand-written sm_100a kernels").  Nothing is computed here: every op checks its tensors,
// fills the C structs of include/multike_b200.h and calls libmultike_b200.so on at::cuda::getCurrentCUDAStream().
// The ops are registered as torch.ops.multike_b200.* (multike_b200/torch_ops.py loads the library):
//   rel_step(ent_var, ent_grad, ent_touched?, rel_var, rel_grad, pos, neg_ent?, neg_side?, K, w?, pos_scale,
//            loss_accum, dim, ent_normalised, rel_normalised, variant)
//        = mke_rel_step_structured: MultiKE_model.py:123-131 + losses.py:4-12 (+ :30-50 with K = 0 / w / pos_scale)
//   rows_apply_adagrad(var, grad, touched?, acc, lr, dim, normalised)
//        = mke_rows_apply_adagrad: MultiKE_model.py:15-31 for the rows the step touched
//   sim_rank(emb1, idx1?, emb2, idx2?, gold?, dim, normalize) -> (rank, top1)
//        = mke_sim_rank: base/alignment.py:8-79, 141-163
#include <c10/cuda/CUDAStream.h>
#include <torch/extension.h>

#include "multike_b200.h"

namespace {

void check_rows(const at::Tensor& t, const char* what, at::ScalarType dtype = at::kFloat) {
  TORCH_CHECK(t.is_cuda() && t.scalar_type() == dtype && t.is_contiguous(), what, ": contiguous CUDA tensor of the right dtype expected");
}
void check_rc(int rc) { TORCH_CHECK(rc == 0, "libmultike_b200: ", mke_last_error()); }

mke_table_t table_of(const at::Tensor& var, const c10::optional<at::Tensor>& grad, const c10::optional<at::Tensor>& touched,
                     int64_t dim, bool normalised) {
  check_rows(var, "table");
  TORCH_CHECK(var.dim() == 2 && dim > 0 && dim <= var.size(1), "table must be [rows, stride >= dim]");
  mke_table_t t{};
  t.var = var.data_ptr<float>();
  t.rows = (int32_t)var.size(0);
  t.stride = (int32_t)var.size(1);
  t.dim = (int32_t)dim;
  t.normalised = normalised ? 1 : 0;
  t.grad_replicas = 1;
  if (grad.has_value()) {
    check_rows(*grad, "gradient accumulator");
    TORCH_CHECK(grad->dim() == 2 || grad->dim() == 3, "gradient accumulator: [rows, stride] or [replicas, rows, stride]");
    if (grad->dim() == 3) t.grad_replicas = (int32_t)grad->size(0);
    TORCH_CHECK(grad->size(-1) == var.size(1) && grad->size(-2) == var.size(0), "gradient accumulator shape");
    t.grad = grad->data_ptr<float>();
  }
  if (touched.has_value()) {
    check_rows(*touched, "touched flags", at::kByte);
    TORCH_CHECK(touched->numel() == var.size(0), "one flag per row");
    t.touched = touched->data_ptr<uint8_t>();
  }
  return t;
}

void rel_step(const at::Tensor& ent_var, const at::Tensor& ent_grad, const c10::optional<at::Tensor>& ent_touched,
              const at::Tensor& rel_var, const at::Tensor& rel_grad, const at::Tensor& pos,
              const c10::optional<at::Tensor>& neg_ent, const c10::optional<at::Tensor>& neg_side, int64_t K,
              const c10::optional<at::Tensor>& w, double pos_scale, at::Tensor loss_accum, int64_t dim, bool ent_normalised,
              bool rel_normalised, int64_t variant) {
  const mke_table_t ent = table_of(ent_var, ent_grad, ent_touched, dim, ent_normalised);
  const mke_table_t rel = table_of(rel_var, rel_grad, c10::nullopt, dim, rel_normalised);
  check_rows(pos, "positives", at::kInt);
  TORCH_CHECK(pos.dim() == 2 && pos.size(1) == 3, "positives: int32 [n, 3]");
  const int32_t n = (int32_t)pos.size(0);
  const int32_t* ne = nullptr;
  const uint32_t* ns = nullptr;
  if (K > 0) {
    TORCH_CHECK(neg_ent.has_value() && neg_side.has_value(), "K > 0 needs neg_ent [n, K] and neg_side [n]");
    check_rows(*neg_ent, "neg_ent", at::kInt);
    check_rows(*neg_side, "neg_side", at::kInt);
    TORCH_CHECK(neg_ent->numel() == (int64_t)n * K && neg_side->numel() == n, "negatives shape");
    ne = neg_ent->data_ptr<int32_t>();
    ns = reinterpret_cast<const uint32_t*>(neg_side->data_ptr<int32_t>());
  }
  const float* wp = nullptr;
  if (w.has_value()) {
    check_rows(*w, "weights");
    TORCH_CHECK(w->numel() == n, "one weight per positive");
    wp = w->data_ptr<float>();
  }
  check_rows(loss_accum, "loss accumulator", at::kDouble);
  check_rc(mke_rel_step_structured(&ent, &rel, pos.data_ptr<int32_t>(), n, (int32_t)K, ne, ns, wp, (float)pos_scale,
                                   loss_accum.data_ptr<double>(), (int32_t)variant,
                                   (mke_stream_t)at::cuda::getCurrentCUDAStream().stream()));
}

void rows_apply_adagrad(const at::Tensor& var, const at::Tensor& grad, const c10::optional<at::Tensor>& touched,
                        const at::Tensor& acc, double lr, int64_t dim, bool normalised) {
  const mke_table_t t = table_of(var, grad, touched, dim, normalised);
  check_rows(acc, "Adagrad accumulator");
  TORCH_CHECK(acc.sizes() == var.sizes(), "accumulator shape");
  check_rc(mke_rows_apply_adagrad(&t, acc.data_ptr<float>(), (float)lr, (mke_stream_t)at::cuda::getCurrentCUDAStream().stream()));
}

std::tuple<at::Tensor, at::Tensor> sim_rank(const at::Tensor& emb1, const c10::optional<at::Tensor>& idx1, const at::Tensor& emb2,
                                            const c10::optional<at::Tensor>& idx2, const c10::optional<at::Tensor>& gold,
                                            int64_t dim, bool normalize) {
  check_rows(emb1, "emb1");
  check_rows(emb2, "emb2");
  TORCH_CHECK(emb1.dim() == 2 && emb2.dim() == 2 && emb1.size(1) == emb2.size(1), "both row sets need the same row stride");
  auto idx_ptr = [](const c10::optional<at::Tensor>& t, const char* what) -> const int32_t* {
    if (!t.has_value()) return nullptr;
    check_rows(*t, what, at::kInt);
    return t->data_ptr<int32_t>();
  };
  const int32_t n1 = (int32_t)(idx1.has_value() ? idx1->numel() : emb1.size(0));
  const int32_t n2 = (int32_t)(idx2.has_value() ? idx2->numel() : emb2.size(0));
  auto opts = emb1.options().dtype(at::kInt);
  at::Tensor rank = at::empty({n1}, opts), top1 = at::empty({n1}, opts);
  if (n1 == 0) return {rank, top1};
  at::Tensor ws = at::empty({mke_sim_rank_workspace_floats(n1, n2, (int32_t)dim)}, emb1.options());
  check_rc(mke_sim_rank(emb1.data_ptr<float>(), idx_ptr(idx1, "idx1"), n1, emb2.data_ptr<float>(), idx_ptr(idx2, "idx2"), n2,
                        (int32_t)emb1.size(1), (int32_t)dim, normalize ? 1 : 0, idx_ptr(gold, "gold"), ws.data_ptr<float>(),
                        rank.data_ptr<int32_t>(), top1.data_ptr<int32_t>(),
                        (mke_stream_t)at::cuda::getCurrentCUDAStream().stream()));
  return {rank, top1};
}

}  // namespace

TORCH_LIBRARY(multike_b200, m) {
  m.def("rel_step(Tensor ent_var, Tensor(a!) ent_grad, Tensor(b!)? ent_touched, Tensor rel_var, Tensor(c!) rel_grad, Tensor pos, "
        "Tensor? neg_ent, Tensor? neg_side, int K, Tensor? w, float pos_scale, Tensor(d!) loss_accum, int dim, "
        "bool ent_normalised, bool rel_normalised, int variant) -> ()", &rel_step);
  m.def("rows_apply_adagrad(Tensor(a!) var, Tensor(b!) grad, Tensor(c!)? touched, Tensor(d!) acc, float lr, int dim, "
        "bool normalised) -> ()", &rows_apply_adagrad);
  m.def("sim_rank(Tensor emb1, Tensor? idx1, Tensor emb2, Tensor? idx2, Tensor? gold, int dim, bool normalize) -> (Tensor, Tensor)",
        &sim_rank);
}
