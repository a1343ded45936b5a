// The slice of ICICLE v4.0.0's C++ backend ABI that this backend binds to, declared from the
// interface the reference documents (bls12-381/include/icicle_backend_api.cuh:98-225,
// include/icicle_types.cuh:47-201, vendored include/icicle/{device,errors,device_api}.h).
// Upstream headers are not on disk in this image, so everything ABI-relevant is isolated HERE:
// if a real ICICLE install disagrees, this is the one file to edit.
//
// What must match upstream exactly, because Itanium name mangling and vtable layout depend on it:
//   * the global-namespace type templates Field<>, Affine<>, Projective<>, ComplexExtensionField<>
//     and the tag structs bls12_381::{fp_config,fq_config,G1,G2}  (only their NAMES matter);
//   * icicle::{Device, eIcicleError, NTTDir, Ordering, MSMConfig, NTTConfig<S>, NTTInitDomainConfig,
//     VecOpsConfig} and the register_* prototypes;
//   * the virtual-method order of icicle::DeviceAPI.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <string>

#include "../../../include/b381.h"

// Only the NAMES of these templates/tags enter the mangled register_* symbols; the bodies just give
// the types their wire size so that std::function signatures over them are complete types.
namespace bls12_381 {
struct fp_config { static constexpr int limbs64 = 4; };   // scalar field tag (Fr)
struct fq_config { static constexpr int limbs64 = 6; };   // base field tag (Fq)
struct G1 {};
struct G2 {};
}  // namespace bls12_381
template <typename Config> class Field { public: uint64_t limbs[Config::limbs64]; };
template <typename BaseConfig, typename BaseField> class ComplexExtensionField { public: BaseField c0, c1; };
template <typename BaseField> class Affine { public: BaseField x, y; };
template <typename BaseField, typename ScalarField, typename Gen> class Projective { public: BaseField x, y, z; };

namespace icicle {

enum class eIcicleError {   // include/icicle/errors.h:37-53 (dense 0..14)
  SUCCESS = 0, INVALID_DEVICE, OUT_OF_MEMORY, INVALID_POINTER, ALLOCATION_FAILED, DEALLOCATION_FAILED,
  COPY_FAILED, SYNCHRONIZATION_FAILED, STREAM_CREATION_FAILED, STREAM_DESTRUCTION_FAILED, API_NOT_IMPLEMENTED,
  INVALID_ARGUMENT, BACKEND_LOAD_FAILED, LICENSE_CHECK_ERROR, UNKNOWN_ERROR
};
inline eIcicleError to_icicle(int b381_code) { return static_cast<eIcicleError>(b381_code); }

struct Device {             // include/icicle/device.h:55-57
  char type[32];
  int id;
};
struct DeviceProperties { bool using_host_memory; int num_memory_regions; bool supports_pinned_memory; };

enum class NTTDir { kForward = 0, kInverse = 1 };
enum class Ordering { kNN = 0, kNR = 1, kRN = 2, kRR = 3, kNM = 4, kMN = 5 };

// Config structs: same members, order and types as include/icicle_types.cuh; the b381_* C structs in
// include/b381.h are their C spelling and the glue reinterpret_casts between the two
// (static_asserts in *_api.cu keep them in lock step).
struct MSMConfig {
  void* stream; int precompute_factor; int c; int bitsize; int batch_size;
  bool are_points_shared_in_batch, are_scalars_on_device, are_scalars_montgomery_form, are_points_on_device,
      are_points_montgomery_form, are_results_on_device, is_async;
  void* ext;
};
template <typename S> struct NTTConfig {
  void* stream; S coset_gen; int batch_size; bool columns_batch; Ordering ordering;
  bool are_inputs_on_device, are_outputs_on_device, is_async;
  void* ext;
};
struct NTTInitDomainConfig { void* stream; bool is_async; void* ext; };
#ifdef B381_ICICLE_UPSTREAM_VECOPS
// upstream ICICLE v4.0.0 (icicle/include/icicle/vec_ops.h): two more members than the reference's copy of the struct.
// The reference's Rust layer sets batch_size (core/vecops.rs:345-346), so a real ICICLE install needs THIS layout;
// build.py emits that variant of the field library under lib/upstream_v4/ (same SONAME, same mangled symbols --
// the layout of a struct is not part of its mangled name).
struct VecOpsConfig { void* stream; bool is_a_on_device, is_b_on_device, is_result_on_device, is_async; int batch_size; bool columns_batch; void* ext; };
#else
struct VecOpsConfig { void* stream; bool is_a_on_device, is_b_on_device, is_result_on_device, is_async; void* ext; };   // icicle_types.cuh:194-201
#endif

using scalar_t = ::Field<bls12_381::fp_config>;
using fq_field_t = ::Field<bls12_381::fq_config>;
using g1_affine_t = ::Affine<fq_field_t>;
using g1_projective_t = ::Projective<fq_field_t, scalar_t, bls12_381::G1>;
using fq2_field_t = ::ComplexExtensionField<bls12_381::fq_config, fq_field_t>;
using g2_affine_t = ::Affine<fq2_field_t>;
using g2_projective_t = ::Projective<fq2_field_t, scalar_t, bls12_381::G2>;

// ---- callback types + registration entry points (weak: null when ICICLE's frontend libs are absent)
using MsmImpl = std::function<eIcicleError(const Device&, const scalar_t*, const g1_affine_t*, int, const MSMConfig&, g1_projective_t*)>;
using MsmPreComputeImpl = std::function<eIcicleError(const Device&, const g1_affine_t*, int, const MSMConfig&, g1_affine_t*)>;
using MsmG2Impl = std::function<eIcicleError(const Device&, const scalar_t*, const g2_affine_t*, int, const MSMConfig&, g2_projective_t*)>;
using MsmG2PreComputeImpl = std::function<eIcicleError(const Device&, const g2_affine_t*, int, const MSMConfig&, g2_affine_t*)>;
using NttImpl = std::function<eIcicleError(const Device&, const scalar_t*, int, NTTDir, const NTTConfig<scalar_t>&, scalar_t*)>;
using NttInitDomainImpl = std::function<eIcicleError(const Device&, const scalar_t&, const NTTInitDomainConfig&)>;
using NttReleaseDomainImpl = std::function<eIcicleError(const Device&, const scalar_t&)>;
using NttGetRouFromDomainImpl = std::function<eIcicleError(const Device&, uint64_t, scalar_t*)>;
using scalarVectorOpImpl = std::function<eIcicleError(const Device&, const scalar_t*, const scalar_t*, uint64_t, const VecOpsConfig&, scalar_t*)>;

__attribute__((weak)) void register_msm(const std::string& deviceType, MsmImpl impl);
__attribute__((weak)) void register_msm_precompute_bases(const std::string& deviceType, MsmPreComputeImpl impl);
__attribute__((weak)) void register_ntt(const std::string& deviceType, NttImpl impl);
__attribute__((weak)) void register_ntt_init_domain(const std::string& deviceType, NttInitDomainImpl impl);
__attribute__((weak)) void register_ntt_release_domain(const std::string& deviceType, NttReleaseDomainImpl impl);
__attribute__((weak)) void register_ntt_get_rou_from_domain(const std::string& deviceType, NttGetRouFromDomainImpl impl);
__attribute__((weak)) void register_vector_add(const std::string& deviceType, scalarVectorOpImpl impl);
__attribute__((weak)) void register_vector_sub(const std::string& deviceType, scalarVectorOpImpl impl);
__attribute__((weak)) void register_vector_mul(const std::string& deviceType, scalarVectorOpImpl impl);
__attribute__((weak)) void register_scalar_mul_vec(const std::string& deviceType, scalarVectorOpImpl impl);
__attribute__((weak)) void register_scalar_add_vec(const std::string& deviceType, scalarVectorOpImpl impl);
// G2: ICICLE core only exports these when built with G2, so -- like the reference
// (src/backend/g2_registry.cu:72-101) -- the backend carries its own registry, defined in g2_registry.cu.
void register_g2_msm(const std::string& deviceType, MsmG2Impl impl);
void register_g2_msm_precompute_bases(const std::string& deviceType, MsmG2PreComputeImpl impl);
MsmG2Impl get_g2_msm_backend(const std::string& deviceType);
MsmG2PreComputeImpl get_g2_msm_precompute_bases_backend(const std::string& deviceType);

// ---- device API (include/icicle/device_api.h); virtual order is ABI
enum eCopyDirection { HostToDevice, DeviceToHost, DeviceToDevice, HostToHost };
typedef void* icicleStreamHandle;
class DeviceAPI {
 public:
  virtual ~DeviceAPI() {}
  virtual eIcicleError set_device(const Device& device) = 0;
  virtual eIcicleError get_device_count(int& device_count) const = 0;
  virtual eIcicleError allocate_memory(void** ptr, size_t size) const = 0;
  virtual eIcicleError allocate_memory_async(void** ptr, size_t size, icicleStreamHandle stream) const = 0;
  virtual eIcicleError free_memory(void* ptr) const = 0;
  virtual eIcicleError free_memory_async(void* ptr, icicleStreamHandle stream) const = 0;
  virtual eIcicleError get_available_memory(size_t& total, size_t& free) const = 0;
  virtual eIcicleError memset(void* ptr, int value, size_t size) const = 0;
  virtual eIcicleError memset_async(void* ptr, int value, size_t size, icicleStreamHandle stream) const = 0;
  virtual eIcicleError copy(void* dst, const void* src, size_t size, eCopyDirection direction) const = 0;
  virtual eIcicleError copy_async(void* dst, const void* src, size_t size, eCopyDirection direction, icicleStreamHandle stream) const = 0;
  virtual eIcicleError synchronize(icicleStreamHandle stream = nullptr) const = 0;
  virtual eIcicleError create_stream(icicleStreamHandle* stream) const = 0;
  virtual eIcicleError destroy_stream(icicleStreamHandle stream) const = 0;
  virtual eIcicleError get_device_properties(DeviceProperties& properties) const = 0;
};
__attribute__((weak)) void register_deviceAPI(const std::string& deviceType, std::shared_ptr<DeviceAPI> api);

}  // namespace icicle

// run `fn` once when the shared object is loaded
#define B381_AT_LOAD(tag) \
  static void b381_at_load_##tag(); \
  namespace { struct B381AtLoad_##tag { B381AtLoad_##tag() { b381_at_load_##tag(); } } b381_at_load_instance_##tag; } \
  static void b381_at_load_##tag()
