// jt_torch.cpp -- PyTorch extension over the C ABI of libjt_b200.so (include/jt_b200.h).
//
// BASELINE.json's north_star asks for "a thin C-ABI layer, called from Python via a PyTorch
// extension".  This file is that extension: it registers the stage calls as operators of the
// `jt_b200` namespace (torch.ops.jt_b200.*).  Each operator takes tensors instead of raw
// pointers, checks device, dtype, contiguity and sizes against the plan, takes the stream from
// torch's current CUDA stream of the workspace's device and forwards to the C entry point of the
// same name.  No compute lives here and there is no CPU implementation: a CPU tensor is an error.
//
// Reference interfaces behind the operators (paths relative to the reference checkout):
//   propagate   JunctionTree.propagate        junctiontree/junctiontree.py:297-331
//   init        CliqueGraph.evaluate          junctiontree/junctiontree.py:203-226
//   collect / distribute   compute_beliefs    junctiontree/computation.py:37-246
//   marginal    CliqueGraph.marginalize       junctiontree/junctiontree.py:229-274
//   ratio       SumProduct.absorb(old=...)    junctiontree/sum_product.py:24-35
//
// Built by csrc/Makefile into junctiontree/_jt_torch.so (g++ against the torch headers, linked
// to libjt_b200.so through $ORIGIN); loaded by junctiontree/torch_ops.py with
// torch.ops.load_library.
#include <ATen/ATen.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include <optional>

#include "../../include/jt_b200.h"

namespace {

jt_plan* as_plan(int64_t handle) {
    TORCH_CHECK(handle != 0, "jt_b200: null plan handle");
    return reinterpret_cast<jt_plan*>(static_cast<intptr_t>(handle));
}

void check(int rc, const char* call) {
    TORCH_CHECK(rc == JT_OK, "libjt_b200 ", call, ": ", jt_last_error_string(), " (code ", rc, ")");
}

int dtype_code(at::ScalarType st) {
    if (st == at::kDouble) return JT_F64;
    if (st == at::kFloat) return JT_F32;
    TORCH_CHECK(false, "jt_b200: only float32 and float64 potentials are supported, got ", st);
    return -1;
}

int64_t query(jt_plan* plan, int what) {
    int64_t out = 0;
    check(jt_plan_query(plan, what, &out), "jt_plan_query");
    return out;
}

void need_cuda(const at::Tensor& t, const char* name) {
    TORCH_CHECK(t.is_cuda(), "jt_b200: ", name, " must be a CUDA tensor (there is no CPU path), got ", t.device());
    TORCH_CHECK(t.is_contiguous(), "jt_b200: ", name, " must be contiguous");
}

void same_device(const at::Tensor& t, const at::Tensor& ws, const char* name) {
    TORCH_CHECK(t.device() == ws.device(), "jt_b200: ", name, " lives on ", t.device(), ", the workspace on ",
                ws.device());
}

// the workspace is a flat byte buffer of at least jt_workspace_bytes(plan, B, dtype)
void check_workspace(jt_plan* plan, const at::Tensor& ws, int64_t B, int dtype) {
    need_cuda(ws, "workspace");
    TORCH_CHECK(B > 0, "jt_b200: batch size must be positive, got ", B);
    size_t need = 0;
    check(jt_workspace_bytes(plan, B, dtype, &need), "jt_workspace_bytes");
    TORCH_CHECK(static_cast<size_t>(ws.nbytes()) >= need, "jt_b200: workspace holds ", ws.nbytes(),
                " bytes, a batch of ", B, " needs ", need);
}

void check_factors(jt_plan* plan, const at::Tensor& factors, bool batched, const at::Tensor& ws, int64_t B) {
    need_cuda(factors, "factor_tables");
    same_device(factors, ws, "factor_tables");
    const int64_t fin = query(plan, JT_H_FIN_ENTRIES);
    const int64_t want = batched ? fin * B : fin;
    TORCH_CHECK(factors.numel() == want, "jt_b200: factor_tables has ", factors.numel(), " values, the plan takes ",
                want, batched ? " ([fin_entries][B])" : " (fin_entries, shared by the batch)");
}

const int32_t* check_evidence(jt_plan* plan, const std::optional<at::Tensor>& evidence, const at::Tensor& ws,
                              int64_t B) {
    const int64_t n_ev = query(plan, JT_H_NEVID);
    if (!evidence.has_value() || !evidence->defined()) {
        TORCH_CHECK(n_ev == 0, "jt_b200: the plan has ", n_ev, " evidence variables but no evidence was given");
        return nullptr;
    }
    const at::Tensor& ev = *evidence;
    need_cuda(ev, "evidence");
    same_device(ev, ws, "evidence");
    TORCH_CHECK(ev.scalar_type() == at::kInt, "jt_b200: evidence must be int32, got ", ev.scalar_type());
    TORCH_CHECK(ev.numel() == B * n_ev, "jt_b200: evidence must hold [", B, "][", n_ev, "] states, got ",
                ev.numel(), " values");
    return n_ev ? ev.data_ptr<int32_t>() : nullptr;
}

void check_out(jt_plan* plan, const at::Tensor& out, const at::Tensor& ws, int64_t B, int dtype) {
    need_cuda(out, "factor_out");
    same_device(out, ws, "factor_out");
    TORCH_CHECK(dtype_code(out.scalar_type()) == dtype, "jt_b200: factor_out has another dtype than the factor tables");
    const int64_t want = query(plan, JT_H_FOUT_ENTRIES) * B;
    TORCH_CHECK(out.numel() == want, "jt_b200: factor_out has ", out.numel(), " values, [fout_entries][B] = ", want);
}

void* stream_of(const at::Tensor& t) {
    return c10::cuda::getCurrentCUDAStream(t.get_device()).stream();
}

// ---- library / plan ----

int64_t abi_version() { return jt_abi_version(); }
int64_t launch_count() { return jt_launch_count(); }

int64_t plan_create(const at::Tensor& blob) {
    TORCH_CHECK(blob.device().is_cpu() && blob.scalar_type() == at::kByte && blob.is_contiguous(),
                "jt_b200: the plan blob must be a contiguous CPU uint8 tensor");
    jt_plan* plan = nullptr;
    check(jt_plan_create(blob.data_ptr(), static_cast<size_t>(blob.numel()), &plan), "jt_plan_create");
    return static_cast<int64_t>(reinterpret_cast<intptr_t>(plan));
}

void plan_destroy(int64_t plan) {
    if (plan) jt_plan_destroy(as_plan(plan));
}

int64_t plan_query(int64_t plan, int64_t what) { return query(as_plan(plan), static_cast<int>(what)); }

void plan_upload(int64_t plan, int64_t device) {
    c10::cuda::CUDAGuard guard(static_cast<c10::DeviceIndex>(device));
    check(jt_plan_upload(as_plan(plan)), "jt_plan_upload");
}

int64_t workspace_bytes(int64_t plan, int64_t B, at::ScalarType dtype) {
    size_t out = 0;
    check(jt_workspace_bytes(as_plan(plan), B, dtype_code(dtype), &out), "jt_workspace_bytes");
    return static_cast<int64_t>(out);
}

// ---- stages (enqueue on torch's current stream of the workspace's device) ----

void init(int64_t plan, const at::Tensor& factors, bool factors_batched, const std::optional<at::Tensor>& evidence,
          at::Tensor workspace, int64_t B, int64_t flags) {
    jt_plan* p = as_plan(plan);
    const int dtype = dtype_code(factors.scalar_type());
    check_workspace(p, workspace, B, dtype);
    check_factors(p, factors, factors_batched, workspace, B);
    const int32_t* ev = check_evidence(p, evidence, workspace, B);
    c10::cuda::CUDAGuard guard(workspace.device());
    check(jt_init(p, factors.data_ptr(), factors_batched ? 1 : 0, ev, B, dtype, workspace.data_ptr(),
                  static_cast<int>(flags), stream_of(workspace)),
          "jt_init");
}

void collect(int64_t plan, at::Tensor workspace, int64_t B, at::ScalarType dtype, int64_t flags) {
    jt_plan* p = as_plan(plan);
    check_workspace(p, workspace, B, dtype_code(dtype));
    c10::cuda::CUDAGuard guard(workspace.device());
    check(jt_collect(p, B, dtype_code(dtype), workspace.data_ptr(), static_cast<int>(flags), stream_of(workspace)),
          "jt_collect");
}

void distribute(int64_t plan, at::Tensor workspace, int64_t B, at::ScalarType dtype, int64_t flags) {
    jt_plan* p = as_plan(plan);
    check_workspace(p, workspace, B, dtype_code(dtype));
    c10::cuda::CUDAGuard guard(workspace.device());
    check(jt_distribute(p, B, dtype_code(dtype), workspace.data_ptr(), static_cast<int>(flags),
                        stream_of(workspace)),
          "jt_distribute");
}

void marginal(int64_t plan, at::Tensor workspace, at::Tensor factor_out, int64_t B, int64_t flags) {
    jt_plan* p = as_plan(plan);
    const int dtype = dtype_code(factor_out.scalar_type());
    check_workspace(p, workspace, B, dtype);
    check_out(p, factor_out, workspace, B, dtype);
    c10::cuda::CUDAGuard guard(workspace.device());
    check(jt_marginal(p, B, dtype, workspace.data_ptr(), factor_out.data_ptr(), static_cast<int>(flags),
                      stream_of(workspace)),
          "jt_marginal");
}

void propagate(int64_t plan, const at::Tensor& factors, bool factors_batched,
               const std::optional<at::Tensor>& evidence, at::Tensor workspace,
               const std::optional<at::Tensor>& factor_out, int64_t B, int64_t flags) {
    jt_plan* p = as_plan(plan);
    const int dtype = dtype_code(factors.scalar_type());
    check_workspace(p, workspace, B, dtype);
    check_factors(p, factors, factors_batched, workspace, B);
    const int32_t* ev = check_evidence(p, evidence, workspace, B);
    void* out = nullptr;
    if (factor_out.has_value() && factor_out->defined()) {
        check_out(p, *factor_out, workspace, B, dtype);
        out = factor_out->data_ptr();
    } else {
        TORCH_CHECK(flags & JT_SKIP_MARGINAL, "jt_b200: propagate without factor_out needs JT_SKIP_MARGINAL");
    }
    c10::cuda::CUDAGuard guard(workspace.device());
    check(jt_propagate(p, factors.data_ptr(), factors_batched ? 1 : 0, ev, B, dtype, workspace.data_ptr(), out,
                       static_cast<int>(flags), stream_of(workspace)),
          "jt_propagate");
}

void normalize(int64_t plan, at::Tensor factor_out, const std::optional<at::Tensor>& logz, int64_t B, int64_t flags) {
    jt_plan* p = as_plan(plan);
    const int dtype = dtype_code(factor_out.scalar_type());
    check_out(p, factor_out, factor_out, B, dtype);
    void* lz = nullptr;
    if (logz.has_value() && logz->defined()) {
        need_cuda(*logz, "logz");
        same_device(*logz, factor_out, "logz");
        TORCH_CHECK(logz->scalar_type() == factor_out.scalar_type() && logz->numel() == B,
                    "jt_b200: logz must hold B values of the outputs' dtype");
        lz = logz->data_ptr();
    }
    c10::cuda::CUDAGuard guard(factor_out.device());
    check(jt_normalize(p, B, dtype, factor_out.data_ptr(), lz, static_cast<int>(flags), stream_of(factor_out)),
          "jt_normalize");
}

int64_t evidence_errors(int64_t plan, at::Tensor workspace, int64_t B, at::ScalarType dtype) {
    jt_plan* p = as_plan(plan);
    check_workspace(p, workspace, B, dtype_code(dtype));
    c10::cuda::CUDAGuard guard(workspace.device());
    int64_t out = 0;
    check(jt_evidence_errors(p, B, dtype_code(dtype), workspace.data_ptr(), stream_of(workspace), &out),
          "jt_evidence_errors");
    return out;
}

// Hugin separator ratio new / old with x / 0 = 0 (SumProduct.absorb(old=...))
at::Tensor ratio(const at::Tensor& new_values, const at::Tensor& old_values) {
    need_cuda(new_values, "new");
    need_cuda(old_values, "old");
    same_device(old_values, new_values, "old");
    TORCH_CHECK(new_values.scalar_type() == old_values.scalar_type() && new_values.sizes() == old_values.sizes(),
                "jt_b200: ratio takes two tensors of one dtype and shape");
    const int dtype = dtype_code(new_values.scalar_type());
    at::Tensor out = at::empty_like(new_values);
    c10::cuda::CUDAGuard guard(new_values.device());
    check(jt_ratio(new_values.data_ptr(), old_values.data_ptr(), out.data_ptr(), new_values.numel(), dtype,
                   stream_of(new_values)),
          "jt_ratio");
    return out;
}

}  // namespace

TORCH_LIBRARY(jt_b200, m) {
    m.def("abi_version() -> int", &abi_version);
    m.def("launch_count() -> int", &launch_count);
    m.def("plan_create(Tensor blob) -> int", &plan_create);
    m.def("plan_destroy(int plan) -> ()", &plan_destroy);
    m.def("plan_query(int plan, int what) -> int", &plan_query);
    m.def("plan_upload(int plan, int device) -> ()", &plan_upload);
    m.def("workspace_bytes(int plan, int B, ScalarType dtype) -> int", &workspace_bytes);
    m.def("init(int plan, Tensor factor_tables, bool factors_batched, Tensor? evidence, Tensor(a!) workspace, "
          "int B, int flags) -> ()",
          &init);
    m.def("collect(int plan, Tensor(a!) workspace, int B, ScalarType dtype, int flags) -> ()", &collect);
    m.def("distribute(int plan, Tensor(a!) workspace, int B, ScalarType dtype, int flags) -> ()", &distribute);
    m.def("marginal(int plan, Tensor(a!) workspace, Tensor(b!) factor_out, int B, int flags) -> ()", &marginal);
    m.def("propagate(int plan, Tensor factor_tables, bool factors_batched, Tensor? evidence, Tensor(a!) workspace, "
          "Tensor(b!)? factor_out, int B, int flags) -> ()",
          &propagate);
    m.def("normalize(int plan, Tensor(a!) factor_out, Tensor(b!)? logz, int B, int flags) -> ()", &normalize);
    m.def("evidence_errors(int plan, Tensor workspace, int B, ScalarType dtype) -> int", &evidence_errors);
    m.def("ratio(Tensor new_values, Tensor old_values) -> Tensor", &ratio);
}
