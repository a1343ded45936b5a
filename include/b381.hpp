// b381.hpp -- header-only C++17 host layer over the C ABI (b381.h), mirroring the reference's Rust
// `core/` API by name and behaviour so that a port of its call sites is mechanical:
//
//   Rust (reference)                                  C++ (here)
//   core/stream.rs:96-198   ManagedStream              b381::ManagedStream
//   core/msm.rs:174-262     PrecomputedBases           b381::PrecomputedBases
//   core/msm.rs:271-1419    GpuMsmContext              b381::GpuMsmContext
//   core/msm.rs:1439-1503   MsmHandle                  b381::MsmHandle
//   core/ntt.rs:303-1390    GpuNttContext              b381::GpuNttContext
//   core/vecops.rs:140-365  vector_add/sub/mul, scalar_mul   b381::vecops::*
//   core/types.rs:126-450   TypeConverter (layouts)    b381::Scalar / G1Affine / ... PODs below
//
// The image has no Rust toolchain, which is why the compiled host layer is C++ (and the tested one is
// the ctypes twin in midnight_bls12_381_cuda_b200/).  Errors are exceptions carrying the eIcicleError
// code; nothing here falls back to the CPU.
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "b381.h"

namespace b381 {

// ---- layouts (core/types.rs:89-108): little-endian u64 limbs, Montgomery form
using Scalar = b381_fr;               // midnight_curves::Fq  == [u64; 4]
using G1Affine = b381_g1_affine;      // x || y, infinity = (0,0)
using G2Affine = b381_g2_affine;
using G1Projective = b381_g1_projective;   // results: (x, y, 1) STANDARD form, identity (0,1,0)
using G2Projective = b381_g2_projective;
static_assert(sizeof(Scalar) == 32 && sizeof(G1Affine) == 96 && sizeof(G2Affine) == 192, "layout");
static_assert(sizeof(G1Projective) == 144 && sizeof(G2Projective) == 288, "layout");

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& where) : std::runtime_error(where + ": b381 error " + std::to_string(c)), code(c) {}
};
using MsmError = Error;
using NttError = Error;
using VecOpsError = Error;
inline void check(int code, const char* where) {
  if (code != B381_SUCCESS) throw Error(code, where);
}

inline void set_device(int id) { check(b381_set_device(id), "set_device"); }
inline bool is_gpu_available() {
  int n = 0;
  return b381_device_count(&n) == B381_SUCCESS && n > 0;
}

class ManagedStream {
 public:
  static ManagedStream create() {
    void* s = nullptr;
    check(b381_stream_create(&s), "stream create");
    return ManagedStream(s, true);
  }
  static ManagedStream default_stream() { return ManagedStream(nullptr, false); }
  ManagedStream(ManagedStream&& o) noexcept : h_(o.h_), owned_(o.owned_), destroyed_(o.destroyed_) { o.owned_ = false; }
  ManagedStream(const ManagedStream&) = delete;
  ~ManagedStream() { try { destroy(); } catch (...) {} }
  void synchronize() { check(b381_stream_synchronize(h_), "stream synchronize"); }
  void destroy() {
    if (owned_ && !destroyed_) check(b381_stream_destroy(h_), "stream destroy");
    destroyed_ = true;
  }
  bool is_destroyed() const { return destroyed_; }
  void* handle() const { return h_; }

 private:
  ManagedStream(void* h, bool owned) : h_(h), owned_(owned) {}
  void* h_;
  bool owned_;
  bool destroyed_ = false;
};

// owned device buffer of T (icicle_runtime::memory::DeviceVec as used by core/)
template <class T>
class DeviceVec {
 public:
  explicit DeviceVec(size_t n) : n_(n) { check(b381_malloc(&p_, n ? n * sizeof(T) : 1), "device_malloc"); }
  DeviceVec(DeviceVec&& o) noexcept : p_(o.p_), n_(o.n_) { o.p_ = nullptr; }
  DeviceVec(const DeviceVec&) = delete;
  ~DeviceVec() { if (p_) b381_free(p_); }
  void copy_from_host(const T* src, size_t n) { check(b381_copy_to_device(p_, src, n * sizeof(T)), "copy_from_host"); }
  void copy_to_host(T* dst, size_t n) const { check(b381_copy_to_host(dst, p_, n * sizeof(T)), "copy_to_host"); }
  T* data() const { return static_cast<T*>(p_); }
  size_t len() const { return n_; }

 private:
  void* p_ = nullptr;
  size_t n_;
};

class PrecomputedBases {
 public:
  // Precomputed buffers hold point i's multiples interleaved at [i*factor + k]; `window` is the cfg.c the table was
  // built with (0 = the backend's fixed default for precomputed bases) and is reused by every MSM over it.
  PrecomputedBases(DeviceVec<G1Affine>&& buf, size_t size, int factor = 1, int window = 0)
      : buf_(std::move(buf)), size_(size), factor_(factor), window_(window) {}
  int window() const { return window_; }
  bool is_precomputed() const { return factor_ > 1; }
  int factor() const { return factor_; }
  size_t original_size() const { return size_; }
  size_t buffer_size() const { return buf_.len(); }
  size_t len() const { return size_; }
  const G1Affine* device_ptr() const { return buf_.data(); }

 private:
  DeviceVec<G1Affine> buf_;
  size_t size_;
  int factor_;
  int window_ = 0;
};

class MsmHandle {
 public:
  MsmHandle(ManagedStream&& s, DeviceVec<Scalar>&& staged) : stream_(std::move(s)), staged_(std::move(staged)) {}
  G1Projective wait() {
    stream_.synchronize();
    stream_.destroy();
    return result_;
  }
  G1Projective* slot() { return &result_; }
  void* stream() const { return stream_.handle(); }
  const Scalar* scalars() const { return staged_.data(); }

 private:
  ManagedStream stream_;
  DeviceVec<Scalar> staged_;
  G1Projective result_{};
};

class GpuMsmContext {
 public:
  explicit GpuMsmContext(int device_id = 0, int window = 0) : window_(window) { set_device(device_id); }

  PrecomputedBases upload_g1_bases(const std::vector<G1Affine>& pts) const {
    DeviceVec<G1Affine> d(pts.size());
    d.copy_from_host(pts.data(), pts.size());
    return PrecomputedBases(std::move(d), pts.size());
  }
  DeviceVec<G2Affine> upload_g2_bases(const std::vector<G2Affine>& pts) const {
    DeviceVec<G2Affine> d(pts.size());
    d.copy_from_host(pts.data(), pts.size());
    return d;
  }
  PrecomputedBases precompute_bases(const PrecomputedBases& bases, int factor) const {
    DeviceVec<G1Affine> out(bases.original_size() * (size_t)factor);
    b381_msm_config cfg = config(true);
    cfg.precompute_factor = factor;
    cfg.are_results_on_device = true;
    check(b381_g1_msm_precompute_bases(bases.device_ptr(), (int)bases.original_size(), &cfg, out.data()), "precompute_bases");
    return PrecomputedBases(std::move(out), bases.original_size(), factor, window_);
  }
  // core/msm.rs:519-592
  G1Projective msm(const std::vector<Scalar>& scalars, const std::vector<G1Affine>& points) const {
    if (scalars.size() != points.size()) throw Error(B381_INVALID_ARGUMENT, "Scalar count != base count");
    G1Projective r{};
    b381_msm_config cfg = config(false);
    check(b381_g1_msm(scalars.data(), points.data(), (int)scalars.size(), &cfg, &r), "msm");
    return r;
  }
  // core/msm.rs:594-682 -- the KZG-commit hot path
  G1Projective msm_with_device_bases(const std::vector<Scalar>& scalars, const PrecomputedBases& bases) const {
    if (scalars.size() > bases.original_size()) throw Error(B381_INVALID_ARGUMENT, "more scalars than bases");
    G1Projective r{};
    b381_msm_config cfg = config(true);
    apply(cfg, bases);
    check(b381_g1_msm(scalars.data(), bases.device_ptr(), (int)scalars.size(), &cfg, &r), "msm_with_device_bases");
    return r;
  }
  // core/msm.rs:715-798
  MsmHandle msm_with_device_bases_async(const std::vector<Scalar>& scalars, const PrecomputedBases& bases) const {
    if (scalars.size() > bases.original_size()) throw Error(B381_INVALID_ARGUMENT, "more scalars than bases");
    ManagedStream st = ManagedStream::create();
    DeviceVec<Scalar> staged(scalars.size());
    check(b381_copy_to_device_async(staged.data(), scalars.data(), scalars.size() * sizeof(Scalar), st.handle()), "h2d");
    MsmHandle h(std::move(st), std::move(staged));
    b381_msm_config cfg = config(true);
    cfg.are_scalars_on_device = true;
    cfg.is_async = true;
    cfg.stream = h.stream();
    apply(cfg, bases);
    check(b381_g1_msm(h.scalars(), bases.device_ptr(), (int)scalars.size(), &cfg, h.slot()), "msm async");
    return h;
  }
  // core/msm.rs:1179-1295: one backend call, batch_size results
  std::vector<G1Projective> msm_batch_with_device_bases(const std::vector<std::vector<Scalar>>& batch, const PrecomputedBases& bases) const {
    if (batch.empty()) return {};
    const size_t n = batch[0].size();
    if (n > bases.original_size()) throw Error(B381_INVALID_ARGUMENT, "more scalars than bases");
    std::vector<Scalar> flat;
    flat.reserve(n * batch.size());
    for (auto& v : batch) {
      if (v.size() != n) throw Error(B381_INVALID_ARGUMENT, "ragged batch");
      flat.insert(flat.end(), v.begin(), v.end());
    }
    std::vector<G1Projective> out(batch.size());
    b381_msm_config cfg = config(true);
    cfg.batch_size = (int)batch.size();
    cfg.are_points_shared_in_batch = true;
    apply(cfg, bases);
    check(b381_g1_msm(flat.data(), bases.device_ptr(), (int)n, &cfg, out.data()), "msm batch");
    return out;
  }
  G2Projective g2_msm(const std::vector<Scalar>& scalars, const std::vector<G2Affine>& points) const {
    if (scalars.size() != points.size()) throw Error(B381_INVALID_ARGUMENT, "Scalar count != base count");
    G2Projective r{};
    b381_msm_config cfg = config(false);
    check(b381_g2_msm(scalars.data(), points.data(), (int)scalars.size(), &cfg, &r), "g2_msm");
    return r;
  }
  G2Projective g2_msm_with_device_bases(const std::vector<Scalar>& scalars, const DeviceVec<G2Affine>& bases) const {
    G2Projective r{};
    b381_msm_config cfg = config(true);
    check(b381_g2_msm(scalars.data(), bases.data(), (int)scalars.size(), &cfg, &r), "g2_msm_with_device_bases");
    return r;
  }

 private:
  b381_msm_config config(bool points_on_device) const {
    b381_msm_config c = b381_default_msm_config();
    c.c = window_;
    c.are_scalars_montgomery_form = true;     // midnight-curves scalars are Montgomery (core/msm.rs:639-651)
    c.are_points_montgomery_form = true;
    c.are_points_on_device = points_on_device;
    return c;
  }
  // precomputed tables fix the factor AND the window they were built with
  static void apply(b381_msm_config& c, const PrecomputedBases& bases) {
    c.precompute_factor = bases.factor();
    if (bases.is_precomputed()) c.c = bases.window();
  }
  int window_;
};

enum class Ordering : int { kNN = B381_kNN, kNR = B381_kNR, kRN = B381_kRN, kRR = B381_kRR, kNM = B381_kNM, kMN = B381_kMN };

class GpuNttContext {
 public:
  // `root_of_unity`: primitive 2^max_log_size-th root, standard or Montgomery form (core/ntt.rs:380-442 passes
  // icicle's get_root_of_unity; here the caller supplies it because this header carries no field code).
  GpuNttContext(uint32_t max_log_size, const Scalar& root_of_unity, int device_id = 0, Ordering ordering = Ordering::kNN)
      : max_log_(max_log_size), ordering_(ordering) {
    set_device(device_id);
    b381_ntt_init_domain_config cfg{};
    check(b381_ntt_init_domain(&root_of_unity, &cfg), "ntt_init_domain");
  }
  uint32_t max_log_size() const { return max_log_; }

  std::vector<Scalar> forward_ntt(const std::vector<Scalar>& v) const { return run(v, B381_NTT_FORWARD, 1, nullptr); }
  std::vector<Scalar> inverse_ntt(const std::vector<Scalar>& v) const { return run(v, B381_NTT_INVERSE, 1, nullptr); }
  void forward_ntt_inplace(std::vector<Scalar>& v) const { run_inplace(v, B381_NTT_FORWARD, 1, nullptr); }
  void inverse_ntt_inplace(std::vector<Scalar>& v) const { run_inplace(v, B381_NTT_INVERSE, 1, nullptr); }
  std::vector<Scalar> forward_ntt_batch(const std::vector<Scalar>& v, size_t poly_size) const { return run(v, B381_NTT_FORWARD, batch_of(v, poly_size), nullptr); }
  std::vector<Scalar> inverse_ntt_batch(const std::vector<Scalar>& v, size_t poly_size) const { return run(v, B381_NTT_INVERSE, batch_of(v, poly_size), nullptr); }
  std::vector<Scalar> forward_coset_ntt(const std::vector<Scalar>& v, const Scalar& g) const { return run(v, B381_NTT_FORWARD, 1, &g); }
  std::vector<Scalar> inverse_coset_ntt(const std::vector<Scalar>& v, const Scalar& g) const { return run(v, B381_NTT_INVERSE, 1, &g); }
  // device-resident, in place (core/ntt.rs:610-919)
  void ntt_on_device(Scalar* device_data, size_t size, int dir, int batch = 1, const Scalar* coset_gen = nullptr,
                     void* stream = nullptr, bool is_async = false) const {
    b381_ntt_config c = config(batch, coset_gen);
    c.are_inputs_on_device = c.are_outputs_on_device = true;
    c.stream = stream;
    c.is_async = is_async;
    check(b381_ntt(device_data, (int)size, dir, &c, device_data), "ntt_on_device");
  }

 private:
  static int batch_of(const std::vector<Scalar>& v, size_t poly) {
    if (!poly || v.size() % poly) throw Error(B381_INVALID_ARGUMENT, "batch length is not a multiple of poly_size");
    return (int)(v.size() / poly);
  }
  b381_ntt_config config(int batch, const Scalar* g) const {
    b381_ntt_config c = b381_default_ntt_config();
    c.batch_size = batch;
    c.ordering = (int)ordering_;
    if (g) c.coset_gen = *g;
    return c;
  }
  std::vector<Scalar> run(const std::vector<Scalar>& v, int dir, int batch, const Scalar* g) const {
    std::vector<Scalar> out(v.size());
    b381_ntt_config c = config(batch, g);
    check(b381_ntt(v.data(), (int)(v.size() / (size_t)batch), dir, &c, out.data()), "ntt");
    return out;
  }
  void run_inplace(std::vector<Scalar>& v, int dir, int batch, const Scalar* g) const {
    b381_ntt_config c = config(batch, g);
    check(b381_ntt(v.data(), (int)(v.size() / (size_t)batch), dir, &c, v.data()), "ntt");
  }
  uint32_t max_log_;
  Ordering ordering_;
};

namespace vecops {
namespace detail {
using fn_t = int (*)(const b381_fr*, const b381_fr*, uint64_t, const b381_vecops_config*, b381_fr*);
inline std::vector<Scalar> binary(fn_t f, const Scalar* a, const std::vector<Scalar>& b, const char* where) {
  std::vector<Scalar> out(b.size());
  b381_vecops_config c = b381_default_vecops_config();
  check(f(a, b.data(), b.size(), &c, out.data()), where);
  return out;
}
inline void same_len(const std::vector<Scalar>& a, const std::vector<Scalar>& b) {
  if (a.size() != b.size()) throw Error(B381_INVALID_ARGUMENT, "length mismatch");
}
}  // namespace detail
inline std::vector<Scalar> vector_add(const std::vector<Scalar>& a, const std::vector<Scalar>& b) { detail::same_len(a, b); return detail::binary(b381_vector_add, a.data(), b, "vector_add"); }
inline std::vector<Scalar> vector_sub(const std::vector<Scalar>& a, const std::vector<Scalar>& b) { detail::same_len(a, b); return detail::binary(b381_vector_sub, a.data(), b, "vector_sub"); }
inline std::vector<Scalar> vector_mul(const std::vector<Scalar>& a, const std::vector<Scalar>& b) { detail::same_len(a, b); return detail::binary(b381_vector_mul, a.data(), b, "vector_mul"); }
inline std::vector<Scalar> scalar_mul(const Scalar& s, const std::vector<Scalar>& a) { return detail::binary(b381_scalar_mul_vec, &s, a, "scalar_mul"); }
inline std::vector<Scalar> scalar_add(const Scalar& s, const std::vector<Scalar>& a) { return detail::binary(b381_scalar_add_vec, &s, a, "scalar_add"); }
// core/vecops.rs:392-535
inline std::vector<Scalar> bit_reverse(const std::vector<Scalar>& a) {
  if (a.size() & (a.size() - 1)) throw Error(B381_INVALID_ARGUMENT, "bit_reverse requires power of 2 length");
  std::vector<Scalar> out(a.size());
  b381_vecops_config c = b381_default_vecops_config();
  check(b381_bit_reverse(a.data(), a.size(), &c, out.data()), "bit_reverse");
  return out;
}
inline void bit_reverse_inplace(std::vector<Scalar>& a) { a = bit_reverse(a); }
inline Scalar vector_sum(const std::vector<Scalar>& a) {
  Scalar out{};
  b381_vecops_config c = b381_default_vecops_config();
  check(b381_vector_sum(a.data(), a.size(), &c, &out), "vector_sum");
  return out;
}
inline std::vector<Scalar> batch_inverse(const std::vector<Scalar>& a) {
  std::vector<Scalar> out(a.size());
  b381_vecops_config c = b381_default_vecops_config();
  check(b381_vector_inv(a.data(), a.size(), &c, out.data()), "batch_inverse");
  return out;
}
}  // namespace vecops

}  // namespace b381
