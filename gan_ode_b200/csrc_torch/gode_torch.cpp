// gode_torch.cpp — the thin PyTorch C++ host over the C ABI (include/gode.h) for the calls the reference actually makes:
//   odeint_adjoint(func, x, t, method='rk4')            models/mocogan_ode.py:142-144   -> Rk4Function
//   odeint / odeint_adjoint(func, h, t)  (dopri5)       models/mocogan_ode_rnn.py:47-48 -> Dopri5Function / Dopri5AdjointFunction
// The Python host (gan_ode_b200/odeint.py) keeps torchdiffeq's signatures, argument checks and error texts and turns a
// call into a cached PLAN (time grid, tolerances, layout, precision ...); this file is what runs per call after that: the
// autograd node, the output / gradient allocations, the persistent workspace and the two C-ABI launches — no Python in
// the backward at all (autograd calls it from its device thread).  It links nothing from libgode.so: the entry points'
// addresses are handed over once by the Python side (bind), which already holds the library through ctypes.
// Everything that is not on this fast path (other methods, data-parallel exchanges, NVTX, eager status checks) stays on
// the Python autograd.Functions, which call the same C ABI.
#include <torch/extension.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include <cstring>
#include <map>
#include <mutex>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "gode.h"

namespace {

using torch::autograd::AutogradContext;
using torch::autograd::variable_list;

// ---- entry points of libgode.so, bound by address --------------------------------------------------------------------
using rk4_fwd_t = decltype(&gode_rk4_fwd);
using rk4_bwd_t = decltype(&gode_rk4_adjoint_bwd);
using dp5_fwd_t = decltype(&gode_dopri5_fwd);
using dp5_bwd_t = decltype(&gode_dopri5_backprop_bwd);
using dp5_adj_t = decltype(&gode_dopri5_adjoint_bwd);
struct Api {
  rk4_fwd_t rk4_fwd = nullptr;
  rk4_bwd_t rk4_adjoint_bwd = nullptr, rk4_backprop_bwd = nullptr;
  dp5_fwd_t dopri5_fwd = nullptr;
  dp5_bwd_t dopri5_backprop_bwd = nullptr;
  dp5_adj_t dopri5_adjoint_bwd = nullptr;
  decltype(&gode_rk4_bwd_workspace_bytes) rk4_bwd_workspace_bytes = nullptr;
  decltype(&gode_dopri5_workspace_bytes) dopri5_workspace_bytes = nullptr;
  decltype(&gode_dopri5_adjoint_workspace_bytes) dopri5_adjoint_workspace_bytes = nullptr;
  decltype(&gode_param_count) param_count = nullptr;
  decltype(&gode_stream_capture_id) stream_capture_id = nullptr;
  decltype(&gode_set_thread_launch_flags) set_thread_launch_flags = nullptr;
  decltype(&gode_strerror) strerror_ = nullptr;
} api;

void bind(const std::unordered_map<std::string, uint64_t>& a) {
  auto get = [&](const char* n) -> void* {
    auto it = a.find(n);
    TORCH_CHECK(it != a.end() && it->second != 0, "gode_torch.bind: missing entry point ", n);
    return reinterpret_cast<void*>(it->second);
  };
  api.rk4_fwd = reinterpret_cast<rk4_fwd_t>(get("gode_rk4_fwd"));
  api.rk4_adjoint_bwd = reinterpret_cast<rk4_bwd_t>(get("gode_rk4_adjoint_bwd"));
  api.rk4_backprop_bwd = reinterpret_cast<rk4_bwd_t>(get("gode_rk4_backprop_bwd"));
  api.dopri5_fwd = reinterpret_cast<dp5_fwd_t>(get("gode_dopri5_fwd"));
  api.dopri5_backprop_bwd = reinterpret_cast<dp5_bwd_t>(get("gode_dopri5_backprop_bwd"));
  api.dopri5_adjoint_bwd = reinterpret_cast<dp5_adj_t>(get("gode_dopri5_adjoint_bwd"));
  api.rk4_bwd_workspace_bytes = reinterpret_cast<decltype(api.rk4_bwd_workspace_bytes)>(get("gode_rk4_bwd_workspace_bytes"));
  api.dopri5_workspace_bytes = reinterpret_cast<decltype(api.dopri5_workspace_bytes)>(get("gode_dopri5_workspace_bytes"));
  api.dopri5_adjoint_workspace_bytes =
      reinterpret_cast<decltype(api.dopri5_adjoint_workspace_bytes)>(get("gode_dopri5_adjoint_workspace_bytes"));
  api.param_count = reinterpret_cast<decltype(api.param_count)>(get("gode_param_count"));
  api.stream_capture_id = reinterpret_cast<decltype(api.stream_capture_id)>(get("gode_stream_capture_id"));
  api.set_thread_launch_flags = reinterpret_cast<decltype(api.set_thread_launch_flags)>(get("gode_set_thread_launch_flags"));
  api.strerror_ = reinterpret_cast<decltype(api.strerror_)>(get("gode_strerror"));
}

// ---- errors: the same Python exception types the Python host raises ------------------------------------------------------
// GodeError for C-ABI failures; for a failed adaptive solve (status mailbox != 0) the Python check_status() is called under
// the GIL so that torchdiffeq's AssertionError texts come out identically.  A backward runs on autograd's device thread:
// the Python error travels to the caller as a pybind11 error_already_set, which torch's boundary restores.
PyObject* g_error_type = nullptr;       // gan_ode_b200.GodeError
PyObject* g_check_status = nullptr;     // gan_ode_b200.odeint.check_status
const volatile int32_t* g_mailbox = nullptr;

void set_hooks(py::object error_type, py::object check_status, uint64_t mailbox_addr) {
  g_error_type = error_type.release().ptr();
  g_check_status = check_status.release().ptr();
  g_mailbox = reinterpret_cast<const volatile int32_t*>(mailbox_addr);
}

[[noreturn]] void fail(const std::string& msg) {
  if (g_error_type) {
    py::gil_scoped_acquire gil;
    PyErr_SetString(g_error_type, msg.c_str());
    throw py::error_already_set();
  }
  TORCH_CHECK(false, msg);
}

void check(int rc, const char* what) {
  if (rc != 0) fail(c10::str(what, " failed: ", api.strerror_ ? api.strerror_(rc) : "?", " (code ", rc, ")"));
}

inline void poll_mailbox() {   // a plain host-memory read on the success path
  if (g_mailbox && *g_mailbox != 0 && g_check_status) {
    py::gil_scoped_acquire gil;
    PyObject* r = PyObject_CallNoArgs(g_check_status);
    if (!r) throw py::error_already_set();
    Py_DECREF(r);
  }
}

// ---- plans ---------------------------------------------------------------------------------------------------------------
struct Plan {
  int kind = 0;  // 0 rk4 family, 1 dopri5 + gradient of the recorded steps, 2 dopri5 + continuous adjoint
  int T = 0, layout = 0, precision = 0, bwd_precision = 0;
  bool adjoint = true;
  std::vector<float> dt_host;   // rk4: step table passed by value (empty: dt_dev)
  at::Tensor dt_dev;
  std::vector<double> t64;      // dopri5: the (possibly negated) time grid
  GodeAdaptiveOpts opts{}, adj_opts{};
  int param_mask = 15;
  bool keep_ckpt = false;
};
std::mutex g_mu;
std::unordered_map<int64_t, Plan> g_plans;
int64_t g_next_plan = 1;

int64_t make_plan(int kind, int T, int layout, int precision, int bwd_precision, bool adjoint, py::bytes dt_host,
                  c10::optional<at::Tensor> dt_dev, py::bytes t64, py::bytes opts, py::bytes adj_opts, int param_mask,
                  bool keep_ckpt) {
  Plan p;
  p.kind = kind; p.T = T; p.layout = layout; p.precision = precision; p.bwd_precision = bwd_precision; p.adjoint = adjoint;
  std::string s = dt_host;
  p.dt_host.resize(s.size() / sizeof(float));
  std::memcpy(p.dt_host.data(), s.data(), p.dt_host.size() * sizeof(float));
  if (dt_dev.has_value()) p.dt_dev = *dt_dev;
  s = t64;
  p.t64.resize(s.size() / sizeof(double));
  std::memcpy(p.t64.data(), s.data(), p.t64.size() * sizeof(double));
  s = opts;
  if (!s.empty()) { TORCH_CHECK(s.size() == sizeof(GodeAdaptiveOpts), "opts size"); std::memcpy(&p.opts, s.data(), sizeof(GodeAdaptiveOpts)); }
  s = adj_opts;
  if (!s.empty()) { TORCH_CHECK(s.size() == sizeof(GodeAdaptiveOpts), "adj_opts size"); std::memcpy(&p.adj_opts, s.data(), sizeof(GodeAdaptiveOpts)); }
  p.param_mask = param_mask; p.keep_ckpt = keep_ckpt;
  std::lock_guard<std::mutex> lk(g_mu);
  const int64_t id = g_next_plan++;
  g_plans.emplace(id, std::move(p));
  return id;
}

void drop_plan(int64_t id) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_plans.erase(id);
}

const Plan& plan_of(int64_t id) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_plans.find(id);
  TORCH_CHECK(it != g_plans.end(), "gode_torch: unknown plan ", id);
  return it->second;   // plans are immutable once made; entries of an unordered_map are address-stable
}

// ---- persistent workspaces (include/gode.h "WORKSPACES ARE PERSISTENT"; same policy as odeint.py::_workspace) ---------------
constexpr int64_t kWsDefault = 8 << 20;
constexpr size_t kWsSpares = 4;
std::map<std::tuple<int, int, unsigned long long>, at::Tensor> g_ws;
std::map<int, std::vector<at::Tensor>> g_spare;

at::Tensor new_ws(const at::Device& dev, int64_t nbytes) {
  return at::zeros({std::max<int64_t>(nbytes, kWsDefault)}, at::TensorOptions().dtype(at::kByte).device(dev));
}

at::Tensor workspace(const at::Device& dev, size_t nbytes, cudaStream_t st) {
  unsigned long long cap_id = 0;
  const bool capturing = api.stream_capture_id(st, &cap_id) == 1;
  const auto key = capturing ? std::make_tuple((int)dev.index(), 1, cap_id)
                             : std::make_tuple((int)dev.index(), 0, (unsigned long long)reinterpret_cast<uintptr_t>(st));
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_ws.find(key);
  if (it != g_ws.end() && (size_t)it->second.numel() >= nbytes) return it->second;
  auto& spares = g_spare[(int)dev.index()];
  at::Tensor ws;
  if (capturing) {
    if (!spares.empty() && (int64_t)nbytes <= kWsDefault) { ws = spares.back(); spares.pop_back(); }
    else ws = new_ws(dev, (int64_t)nbytes);   // zero-filled inside the capture: one memset node per replay
  } else {
    ws = new_ws(dev, (int64_t)nbytes);
    while (spares.size() < kWsSpares) spares.push_back(new_ws(dev, kWsDefault));
    if (g_ws.size() > 256)
      for (auto i = g_ws.begin(); i != g_ws.end();) i = std::get<1>(i->first) == 0 ? g_ws.erase(i) : std::next(i);
  }
  g_ws[key] = ws;
  return ws;
}

// ---- helpers ---------------------------------------------------------------------------------------------------------------
// Device guard + current stream of the tensor's device.  (CPU tensors only occur in the host-wiring tests, which bind
// recording stubs instead of libgode.so's entry points: no guard, null stream.)
struct OnDevice {
  c10::cuda::OptionalCUDAGuard guard;
  cudaStream_t stream = nullptr;
  explicit OnDevice(const at::Tensor& t) {
    if (t.is_cuda()) {
      guard.set_device(t.device());
      stream = c10::cuda::getCurrentCUDAStream(t.device().index()).stream();
    }
  }
};

inline at::Tensor f32c(const at::Tensor& t) {   // fp32, contiguous, 16-byte aligned; the common case returns t itself
  if (t.scalar_type() == at::kFloat && t.is_contiguous() && !(reinterpret_cast<uintptr_t>(t.data_ptr()) & 15)) return t;
  at::Tensor u = t.detach().to(at::kFloat).contiguous();
  if (reinterpret_cast<uintptr_t>(u.data_ptr()) & 15) u = u.clone();
  return u;
}
inline const float* fp(const at::Tensor& t) { return t.defined() ? t.data_ptr<float>() : nullptr; }

inline at::Tensor grad_in_layout(const at::Tensor& g, int layout) {
  if (layout == GODE_LAYOUT_TBD && g.scalar_type() == at::kFloat && g.is_contiguous() &&
      !(reinterpret_cast<uintptr_t>(g.data_ptr()) & 15))
    return g;
  at::Tensor u = g.detach().to(at::kFloat);
  if (layout == GODE_LAYOUT_BTD) u = u.transpose(0, 1);
  u = u.contiguous();
  if (reinterpret_cast<uintptr_t>(u.data_ptr()) & 15) u = u.clone();
  return u;
}

struct LaunchFlags {   // thread-local launch flags of libgode.so around one backward launch
  bool on;
  explicit LaunchFlags(bool pdl) : on(pdl) { if (on) api.set_thread_launch_flags(GODE_LAUNCH_PDL_BWD); }
  ~LaunchFlags() { if (on) api.set_thread_launch_flags(0); }
};

variable_list split_params(const at::Tensor& gp, int64_t D, int64_t H, AutogradContext* ctx, int first) {
  const int64_t n1 = H * D;
  variable_list out(4);
  if (ctx->needs_input_grad(first + 0)) out[0] = gp.narrow(0, 0, n1).view({H, D});
  if (ctx->needs_input_grad(first + 1)) out[1] = gp.narrow(0, n1, H);
  if (ctx->needs_input_grad(first + 2)) out[2] = gp.narrow(0, n1 + H, D * H).view({D, H});
  if (ctx->needs_input_grad(first + 3)) out[3] = gp.narrow(0, n1 + H + D * H, D);
  return out;
}

// ---- fixed-grid rk4 (3/8 rule) -------------------------------------------------------------------------------------------------
struct Rk4Function : public torch::autograd::Function<Rk4Function> {
  static at::Tensor forward(AutogradContext* ctx, const at::Tensor& y0, const at::Tensor& W1, const at::Tensor& b1,
                            const at::Tensor& W2, const at::Tensor& b2, int64_t plan_id, bool pdl) {
    const Plan& p = plan_of(plan_id);
    OnDevice on(y0);
    cudaStream_t st = on.stream;
    const at::Tensor y = f32c(y0), w1 = f32c(W1), c1 = f32c(b1), w2 = f32c(W2), c2 = f32c(b2);
    const int64_t B = y.size(0), D = y.size(1), H = w1.size(0), T = p.T;
    at::Tensor buf = p.layout == GODE_LAYOUT_TBD ? at::empty({T, B, D}, y.options()) : at::empty({B, T, D}, y.options());
    const bool dev_dt = p.dt_host.empty();
    check(api.rk4_fwd(fp(y), fp(w1), fp(c1), fp(w2), fp(c2), dev_dt ? fp(p.dt_dev) : p.dt_host.data(), dev_dt ? 1 : 0,
                      (int)B, (int)D, (int)H, (int)T, p.precision, p.layout, buf.data_ptr<float>(), st),
          "gode_rk4_fwd");
    ctx->save_for_backward({buf, w1, c1, w2, c2});
    ctx->saved_data["plan"] = plan_id;
    ctx->saved_data["pdl"] = pdl;
    return p.layout == GODE_LAYOUT_TBD ? buf : buf.transpose(0, 1);
  }

  static variable_list backward(AutogradContext* ctx, variable_list grads) {
    const auto saved = ctx->get_saved_variables();
    const at::Tensor &buf = saved[0], &w1 = saved[1], &c1 = saved[2], &w2 = saved[3], &c2 = saved[4];
    const Plan& p = plan_of(ctx->saved_data["plan"].toInt());
    const bool pdl = ctx->saved_data["pdl"].toBool();
    OnDevice on(buf);
    cudaStream_t st = on.stream;
    const int64_t T = p.T, H = w1.size(0);
    const int64_t B = p.layout == GODE_LAYOUT_TBD ? buf.size(1) : buf.size(0), D = buf.size(2);
    const at::Tensor g = grad_in_layout(grads[0], p.layout);
    const int64_t n_param = api.param_count((int)D, (int)H), n_pad = (n_param + 3) & ~int64_t(3);
    at::Tensor gbuf = at::empty({n_pad + B * D}, buf.options());   // one allocation: [flat parameter gradient | grad_y0]
    at::Tensor gp = gbuf.narrow(0, 0, n_param), gy = gbuf.narrow(0, n_pad, B * D).view({B, D});
    const size_t ws_bytes = api.rk4_bwd_workspace_bytes((int)B, (int)D, (int)H, (int)T);
    at::Tensor ws = workspace(buf.device(), ws_bytes, st);
    const bool dev_dt = p.dt_host.empty();
    {
      LaunchFlags lf(pdl);
      auto fn = p.adjoint ? api.rk4_adjoint_bwd : api.rk4_backprop_bwd;
      check(fn(fp(buf), fp(g), fp(w1), fp(c1), fp(w2), fp(c2), dev_dt ? fp(p.dt_dev) : p.dt_host.data(), dev_dt ? 1 : 0,
               (int)B, (int)D, (int)H, (int)T, p.bwd_precision, p.layout, gy.data_ptr<float>(), gp.data_ptr<float>(),
               ws.data_ptr(), ws_bytes, st),
            "gode_rk4_bwd");
    }
    variable_list out(7);
    if (ctx->needs_input_grad(0)) out[0] = gy;
    auto ps = split_params(gp, D, H, ctx, 1);
    for (int k = 0; k < 4; ++k) out[1 + k] = ps[k];
    return out;
  }
};

// ---- adaptive dopri5 --------------------------------------------------------------------------------------------------------------
inline int64_t log_bytes(int64_t cap) { return 64 + 8 * cap + 8 * cap + 4 * cap + cap; }

// shared forward launch: returns (buf, raw log, ckpt, acc)
struct Dp5Fwd {
  at::Tensor buf, raw, ckpt, acc, w1, c1, w2, c2;
  int kc = 0;
};
Dp5Fwd dopri5_forward(const Plan& p, const at::Tensor& y0, const at::Tensor& W1, const at::Tensor& b1, const at::Tensor& W2,
                      const at::Tensor& b2, bool keep, cudaStream_t st) {
  Dp5Fwd r;
  const at::Tensor y = f32c(y0);
  r.w1 = f32c(W1); r.c1 = f32c(b1); r.w2 = f32c(W2); r.c2 = f32c(b2);
  const int64_t B = y.size(0), D = y.size(1), H = r.w1.size(0), T = p.T;
  r.buf = p.layout == GODE_LAYOUT_TBD ? at::empty({T, B, D}, y.options()) : at::empty({B, T, D}, y.options());
  const int64_t cap = p.opts.log_capacity;
  r.raw = at::empty({log_bytes(cap)}, y.options().dtype(at::kByte));
  GodeAdaptiveOpts o = p.opts;
  r.kc = keep ? p.opts.ckpt_capacity : 0;
  o.ckpt_capacity = r.kc;
  if (keep) {
    r.ckpt = at::empty({std::max(r.kc, 1), B, D}, y.options());
    r.acc = at::empty({2 * (int64_t)std::max(r.kc, 1)}, y.options().dtype(at::kDouble));
  }
  const size_t ws_bytes = api.dopri5_workspace_bytes((int)B, (int)D, (int)H);
  at::Tensor ws = workspace(y.device(), ws_bytes, st);
  uint8_t* base = r.raw.data_ptr<uint8_t>();
  double* accp = keep ? r.acc.data_ptr<double>() : nullptr;
  check(api.dopri5_fwd(fp(y), fp(r.w1), fp(r.c1), fp(r.w2), fp(r.c2), p.t64.data(), (int)B, (int)D, (int)H, (int)T, &o,
                       p.layout, r.buf.data_ptr<float>(), reinterpret_cast<GodeStepLog*>(base),
                       reinterpret_cast<double*>(base + 64), reinterpret_cast<double*>(base + 64 + 8 * cap),
                       reinterpret_cast<float*>(base + 64 + 16 * cap), base + 64 + 20 * cap,
                       keep ? r.ckpt.data_ptr<float>() : nullptr, accp, keep ? accp + std::max(r.kc, 1) : nullptr,
                       ws.data_ptr(), ws_bytes, st),
        "gode_dopri5_fwd");
  return r;
}

struct Dopri5Function : public torch::autograd::Function<Dopri5Function> {
  static variable_list forward(AutogradContext* ctx, const at::Tensor& y0, const at::Tensor& W1, const at::Tensor& b1,
                               const at::Tensor& W2, const at::Tensor& b2, int64_t plan_id, bool pdl) {
    const Plan& p = plan_of(plan_id);
    OnDevice on(y0);
    Dp5Fwd r = dopri5_forward(p, y0, W1, b1, W2, b2, p.keep_ckpt, on.stream);
    if (p.keep_ckpt) ctx->save_for_backward({r.raw, r.ckpt, r.acc, r.w1, r.c1, r.w2, r.c2});
    ctx->saved_data["plan"] = plan_id;
    ctx->saved_data["pdl"] = pdl;
    ctx->saved_data["kc"] = (int64_t)r.kc;
    ctx->mark_non_differentiable({r.raw});
    return {p.layout == GODE_LAYOUT_TBD ? r.buf : r.buf.transpose(0, 1), r.raw};
  }

  static variable_list backward(AutogradContext* ctx, variable_list grads) {
    const auto saved = ctx->get_saved_variables();
    if (saved.size() != 7) fail("dopri5 forward ran without checkpoints (inputs did not require grad)");
    poll_mailbox();   // a forward that has already failed on the device is reported here, before the replay is launched
    const at::Tensor &raw = saved[0], &ckpt = saved[1], &acc = saved[2], &w1 = saved[3], &c1 = saved[4], &w2 = saved[5],
                     &c2 = saved[6];
    const Plan& p = plan_of(ctx->saved_data["plan"].toInt());
    const bool pdl = ctx->saved_data["pdl"].toBool();
    const int kc = (int)ctx->saved_data["kc"].toInt();
    OnDevice on(ckpt);
    cudaStream_t st = on.stream;
    const int64_t T = p.T, B = ckpt.size(1), D = ckpt.size(2), H = w1.size(0);
    const at::Tensor g = grad_in_layout(grads[0], p.layout);
    at::Tensor gy = at::empty({B, D}, ckpt.options());
    at::Tensor gp = at::empty({api.param_count((int)D, (int)H)}, ckpt.options());
    const size_t ws_bytes = api.dopri5_workspace_bytes((int)B, (int)D, (int)H);
    at::Tensor ws = workspace(ckpt.device(), ws_bytes, st);
    {
      LaunchFlags lf(pdl);
      check(api.dopri5_backprop_bwd(fp(g), fp(w1), fp(c1), fp(w2), fp(c2), p.t64.data(), (int)B, (int)D, (int)H, (int)T,
                                    p.layout, reinterpret_cast<const GodeStepLog*>(raw.data_ptr<uint8_t>()), fp(ckpt),
                                    acc.data_ptr<double>(), acc.data_ptr<double>() + kc, kc, p.opts.fsign,
                                    gy.data_ptr<float>(), gp.data_ptr<float>(), ws.data_ptr(), ws_bytes, st),
            "gode_dopri5_backprop_bwd");
    }
    variable_list out(7);
    if (ctx->needs_input_grad(0)) out[0] = gy;
    auto ps = split_params(gp, D, H, ctx, 1);
    for (int k = 0; k < 4; ++k) out[1 + k] = ps[k];
    return out;
  }
};

struct Dopri5AdjointFunction : public torch::autograd::Function<Dopri5AdjointFunction> {
  static variable_list forward(AutogradContext* ctx, const at::Tensor& y0, const at::Tensor& W1, const at::Tensor& b1,
                               const at::Tensor& W2, const at::Tensor& b2, int64_t plan_id) {
    const Plan& p = plan_of(plan_id);
    OnDevice on(y0);
    Dp5Fwd r = dopri5_forward(p, y0, W1, b1, W2, b2, false, on.stream);
    ctx->save_for_backward({r.buf, r.w1, r.c1, r.w2, r.c2});
    ctx->saved_data["plan"] = plan_id;
    ctx->mark_non_differentiable({r.raw});
    return {p.layout == GODE_LAYOUT_TBD ? r.buf : r.buf.transpose(0, 1), r.raw};
  }

  static variable_list backward(AutogradContext* ctx, variable_list grads) {
    const auto saved = ctx->get_saved_variables();
    const at::Tensor &buf = saved[0], &w1 = saved[1], &c1 = saved[2], &w2 = saved[3], &c2 = saved[4];
    const Plan& p = plan_of(ctx->saved_data["plan"].toInt());
    OnDevice on(buf);
    cudaStream_t st = on.stream;
    const int64_t T = p.T, H = w1.size(0);
    const int64_t B = p.layout == GODE_LAYOUT_TBD ? buf.size(1) : buf.size(0), D = buf.size(2);
    const at::Tensor g = grad_in_layout(grads[0], p.layout);
    at::Tensor gy = at::empty({B, D}, buf.options());
    at::Tensor gp = at::empty({api.param_count((int)D, (int)H)}, buf.options());
    const int64_t cap = p.adj_opts.log_capacity;
    at::Tensor raw = at::zeros({log_bytes(cap)}, buf.options().dtype(at::kByte));
    uint8_t* base = raw.data_ptr<uint8_t>();
    const size_t ws_bytes = api.dopri5_adjoint_workspace_bytes((int)B, (int)D, (int)H);
    at::Tensor ws = workspace(buf.device(), ws_bytes, st);
    const int rc = api.dopri5_adjoint_bwd(fp(buf), fp(g), fp(w1), fp(c1), fp(w2), fp(c2), p.t64.data(), (int)B, (int)D, (int)H,
                                          (int)T, p.layout, &p.adj_opts, p.param_mask, gy.data_ptr<float>(),
                                          gp.data_ptr<float>(), reinterpret_cast<GodeStepLog*>(base),
                                          reinterpret_cast<double*>(base + 64 + 8 * cap),
                                          reinterpret_cast<float*>(base + 64 + 16 * cap), base + 64 + 20 * cap, ws.data_ptr(),
                                          ws_bytes, st);
    if (rc == GODE_ERR_COOP)
      fail("the continuous dopri5 adjoint keeps the whole batch co-resident (at most 18944 trajectories per GPU); shard the "
           "batch, or pass options={'adjoint': 'discrete'} for the gradient of the recorded steps (gode_dopri5_backprop_bwd)");
    check(rc, "gode_dopri5_adjoint_bwd");
    {
      std::lock_guard<std::mutex> lk(g_mu);
      last_adjoint_log() = raw;
    }
    variable_list out(6);
    if (ctx->needs_input_grad(0)) out[0] = gy;
    auto ps = split_params(gp, D, H, ctx, 1);
    for (int k = 0; k < 4; ++k) out[1 + k] = ps[k];
    return out;
  }

  static at::Tensor& last_adjoint_log() {
    static at::Tensor t;
    return t;
  }
};

at::Tensor rk4(const at::Tensor& y0, const at::Tensor& W1, const at::Tensor& b1, const at::Tensor& W2, const at::Tensor& b2,
               int64_t plan, bool pdl) {
  return Rk4Function::apply(y0, W1, b1, W2, b2, plan, pdl);
}
std::vector<at::Tensor> dopri5(const at::Tensor& y0, const at::Tensor& W1, const at::Tensor& b1, const at::Tensor& W2,
                               const at::Tensor& b2, int64_t plan, bool pdl) {
  return Dopri5Function::apply(y0, W1, b1, W2, b2, plan, pdl);
}
std::vector<at::Tensor> dopri5_adjoint(const at::Tensor& y0, const at::Tensor& W1, const at::Tensor& b1, const at::Tensor& W2,
                                       const at::Tensor& b2, int64_t plan) {
  return Dopri5AdjointFunction::apply(y0, W1, b1, W2, b2, plan);
}
at::Tensor last_adjoint_log() {
  std::lock_guard<std::mutex> lk(g_mu);
  return Dopri5AdjointFunction::last_adjoint_log();
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "thin PyTorch C++ host of gan_ode_b200 over the C ABI of libgode.so (include/gode.h)";
  m.def("bind", &bind, "hand over the addresses of the libgode.so entry points (name -> address)");
  m.def("set_hooks", &set_hooks, "GodeError type, check_status callable, address of the status mailbox (0: none)");
  m.def("make_plan", &make_plan);
  m.def("drop_plan", &drop_plan);
  m.def("rk4", &rk4);
  m.def("dopri5", &dopri5);
  m.def("dopri5_adjoint", &dopri5_adjoint);
  m.def("last_adjoint_log", &last_adjoint_log);
}
