// tests/tf_mock -- a MOCK of the slice of TensorFlow's C++ op-kernel API that tf_shim/gaccum_tf_op.cc uses.
// TEST INFRASTRUCTURE ONLY.  TensorFlow's headers are not in this image, so the adapter could never be compiled; with
// these headers first on the include path it compiles unmodified, registers its ops and kernels in a tiny registry, and
// tests/tf_mock/tf_op_driver.cc runs its Compute() against real device memory through libgaccum.so.
// What this pins: the adapter's own logic (attribute parsing, input indexing against ITS OWN REGISTER_OP order, resource vs
// ref variables, host-memory scalars, error paths).  What it cannot pin: conformance of these mock signatures to real
// TensorFlow -- they are restated from the TF 2.x headers from memory (tensorflow/core/framework/op_kernel.h, op.h,
// resource_var.h, resource_mgr.h), which is said here once and in DESIGN.md.
#pragma once
#include <atomic>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

namespace Eigen { struct GpuDevice {}; }

namespace tensorflow {

// ---- Status / errors -------------------------------------------------------------------------------------------
class Status {
 public:
  Status() = default;
  Status(int code, std::string msg) : code_(code), msg_(std::move(msg)) {}
  bool ok() const { return code_ == 0; }
  const std::string& message() const { return msg_; }
  std::string ToString() const { return ok() ? "OK" : msg_; }
 private:
  int code_ = 0;
  std::string msg_;
};
namespace errors {
template <typename... A> std::string StrCat(const A&... a) { std::ostringstream o; (o << ... << a); return o.str(); }
template <typename... A> Status InvalidArgument(const A&... a) { return Status(3, StrCat(a...)); }
template <typename... A> Status Internal(const A&... a) { return Status(13, StrCat(a...)); }
}  // namespace errors

// ---- Tensor: a typed view of caller-owned memory (device or host) ---------------------------------------------------
enum DataType { DT_FLOAT = 1, DT_INT64 = 9, DT_RESOURCE = 20 };
class Tensor {
 public:
  Tensor() = default;
  Tensor(DataType dt, void* data, int64_t numel) : dt_(dt), data_(data), n_(numel) {}
  int64_t NumElements() const { return n_; }
  DataType dtype() const { return dt_; }
  template <typename T> struct Flat {
    T* p; int64_t n;
    T* data() const { return p; }
    T& operator()(int64_t i) const { return p[i]; }
    int64_t size() const { return n; }
  };
  template <typename T> struct Scalar { T* p; T& operator()() const { return *p; } };
  template <typename T> Flat<T> flat() const { return Flat<T>{static_cast<T*>(data_), n_}; }
  template <typename T> Scalar<T> scalar() const { return Scalar<T>{static_cast<T*>(data_)}; }
  void* raw() const { return data_; }
 private:
  DataType dt_ = DT_FLOAT;
  void* data_ = nullptr;
  int64_t n_ = 0;
};

// ---- attribute bag / construction --------------------------------------------------------------------------------
struct AttrValue { int64_t i = 0; std::string s; std::vector<bool> bl; };
class OpKernelConstruction {
 public:
  std::map<std::string, AttrValue> attrs;
  Status status;
  Status GetAttr(const std::string& n, int* v) const { auto it = attrs.find(n); if (it == attrs.end()) return errors::InvalidArgument("no attr ", n); *v = (int)it->second.i; return Status(); }
  Status GetAttr(const std::string& n, std::string* v) const { auto it = attrs.find(n); if (it == attrs.end()) return errors::InvalidArgument("no attr ", n); *v = it->second.s; return Status(); }
  Status GetAttr(const std::string& n, std::vector<bool>* v) const { auto it = attrs.find(n); if (it == attrs.end()) return errors::InvalidArgument("no attr ", n); *v = it->second.bl; return Status(); }
  void CtxFailure(const Status& s) { if (status.ok()) status = s; }
  void CtxFailure(const char*, int, const Status& s) { CtxFailure(s); }
};

// ---- device / stream ---------------------------------------------------------------------------------------------------
struct PlatformSpecificHandle { void* stream = nullptr; void* bound_stream = nullptr; };
class Stream { public: explicit Stream(void* s) { h_.stream = s; } PlatformSpecificHandle platform_specific_handle() const { return h_; } private: PlatformSpecificHandle h_; };
class DeviceContext { public: explicit DeviceContext(Stream* s) : s_(s) {} Stream* stream() const { return s_; } private: Stream* s_; };
struct AcceleratorDeviceInfo { int gpu_id = 0; };
class DeviceBase { public: AcceleratorDeviceInfo info; const AcceleratorDeviceInfo* tensorflow_accelerator_device_info() const { return &info; } };

// ---- resource variables ------------------------------------------------------------------------------------------------
class mutex : public std::mutex {};
class mutex_lock { public: explicit mutex_lock(mutex& m) : l_(m) {} private: std::lock_guard<std::mutex> l_; };
class Var {
 public:
  mutex* mu() { return &mu_; }
  Tensor* tensor() { return &t_; }
  std::atomic<bool> copy_on_read_mode{false};
  int prepared = 0;                 // how often PrepareToUpdateVariable saw this variable (the driver checks it)
 private:
  mutex mu_;
  Tensor t_;
};
namespace core {
// real TF: an owning smart pointer that Unref()s; the mock's variables are owned by the driver
template <typename T> class RefCountPtr {
 public:
  RefCountPtr() = default;
  explicit RefCountPtr(T* p) : p_(p) {}
  RefCountPtr(RefCountPtr&& o) noexcept : p_(o.p_) { o.p_ = nullptr; }
  RefCountPtr& operator=(RefCountPtr&& o) noexcept { p_ = o.p_; o.p_ = nullptr; return *this; }
  T* operator->() const { return p_; }
  T* get() const { return p_; }
  void reset(T* p) { p_ = p; }
 private:
  T* p_ = nullptr;
};
}  // namespace core
struct ResourceHandle { Var* var = nullptr; };

// ---- the kernel context ------------------------------------------------------------------------------------------------
class OpKernelContext {
 public:
  std::vector<Tensor> inputs;             // in the op's declared input order
  std::vector<Var*> resources;            // parallel to inputs: non-null where the input is a DT_RESOURCE handle
  std::vector<bool> is_ref;               // parallel: declared as Ref(...)
  DeviceContext* dev_ctx = nullptr;
  DeviceBase* dev = nullptr;
  Status status_;
  const Tensor& input(int i) const { return inputs.at(i); }
  Tensor mutable_input(int i, bool /*lock_held*/) {
    if (!is_ref.at(i)) status_ = errors::InvalidArgument("mutable_input(", i, ") on an input that is not a Ref");
    return inputs.at(i);
  }
  DeviceContext* op_device_context() const { return dev_ctx; }
  DeviceBase* device() const { return dev; }
  void CtxFailure(const Status& s) { if (status_.ok()) status_ = s; }
  void CtxFailure(const char*, int, const Status& s) { CtxFailure(s); }
  const Status& status() const { return status_; }
};
inline ResourceHandle HandleFromInput(OpKernelContext* c, int i) { return ResourceHandle{c->resources.at(i)}; }
template <typename T> Status LookupResource(OpKernelContext*, const ResourceHandle& h, core::RefCountPtr<T>* out) {
  if (h.var == nullptr) return errors::InvalidArgument("input is not a resource handle");
  out->reset(h.var);
  return Status();
}
template <typename Device, typename T> Status PrepareToUpdateVariable(OpKernelContext*, Tensor*, bool) { return Status(); }

// ---- OpKernel + registries ---------------------------------------------------------------------------------------------
class OpKernel {
 public:
  explicit OpKernel(OpKernelConstruction*) {}
  virtual ~OpKernel() = default;
  virtual void Compute(OpKernelContext* ctx) = 0;
};

struct OpDef {
  std::string name, doc;
  std::vector<std::string> inputs, attrs;
  bool stateful = false;
};
namespace shape_inference { class InferenceContext; inline Status NoOutputs(InferenceContext*) { return Status(); } }
class OpDefBuilderWrapper {
 public:
  explicit OpDefBuilderWrapper(const char* name) { def_.name = name; }
  OpDefBuilderWrapper& Input(const std::string& s) { def_.inputs.push_back(s); return *this; }
  OpDefBuilderWrapper& Attr(const std::string& s) { def_.attrs.push_back(s); return *this; }
  OpDefBuilderWrapper& SetIsStateful() { def_.stateful = true; return *this; }
  template <typename F> OpDefBuilderWrapper& SetShapeFn(F) { return *this; }
  OpDefBuilderWrapper& Doc(const std::string& s) { def_.doc = s; return *this; }
  const OpDef& def() const { return def_; }
 private:
  OpDef def_;
};
struct KernelDef { std::string op, device; std::vector<std::string> host_memory; };
class KernelDefBuilder {
 public:
  explicit KernelDefBuilder(const char* op) { def_.op = op; }
  KernelDefBuilder& Device(const char* d) { def_.device = d; return *this; }
  KernelDefBuilder& HostMemory(const char* n) { def_.host_memory.push_back(n); return *this; }
  const KernelDef& def() const { return def_; }
 private:
  KernelDef def_;
};
// real TF: the macro pastes `::tensorflow::register_kernel::` in front of its first argument, so `Name(...)` needs no qualifier
namespace register_kernel { inline KernelDefBuilder Name(const char* op) { return KernelDefBuilder(op); } }
static const char* const DEVICE_GPU = "GPU";
static const char* const DEVICE_CPU = "CPU";

struct MockRegistry {
  std::map<std::string, OpDef> ops;
  std::map<std::string, std::pair<KernelDef, std::function<OpKernel*(OpKernelConstruction*)>>> kernels;
  static MockRegistry& Get() { static MockRegistry r; return r; }
};
struct OpRegistrar { OpRegistrar(const OpDefBuilderWrapper& b) { MockRegistry::Get().ops[b.def().name] = b.def(); } };
struct KernelRegistrar {
  KernelRegistrar(const KernelDefBuilder& b, std::function<OpKernel*(OpKernelConstruction*)> f) { MockRegistry::Get().kernels[b.def().op] = {b.def(), std::move(f)}; }
};

}  // namespace tensorflow

#define TF_MOCK_CAT2(a, b) a##b
#define TF_MOCK_CAT(a, b) TF_MOCK_CAT2(a, b)
#define REGISTER_OP(name) static ::tensorflow::OpRegistrar TF_MOCK_CAT(tf_mock_op_, __COUNTER__) = ::tensorflow::OpDefBuilderWrapper(name)
#define REGISTER_KERNEL_BUILDER(builder, ...) \
  static ::tensorflow::KernelRegistrar TF_MOCK_CAT(tf_mock_kernel_, __COUNTER__)(::tensorflow::register_kernel::builder, [](::tensorflow::OpKernelConstruction* c) -> ::tensorflow::OpKernel* { return new __VA_ARGS__(c); })
#define OP_REQUIRES(CTX, EXP, STATUS) do { if (!(EXP)) { (CTX)->CtxFailure(__FILE__, __LINE__, (STATUS)); return; } } while (0)
#define OP_REQUIRES_OK(CTX, ...) do { ::tensorflow::Status _s(__VA_ARGS__); if (!_s.ok()) { (CTX)->CtxFailure(__FILE__, __LINE__, _s); return; } } while (0)
#define TF_RETURN_IF_ERROR(...) do { ::tensorflow::Status _s(__VA_ARGS__); if (!_s.ok()) return _s; } while (0)
