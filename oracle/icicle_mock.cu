// TEST INFRASTRUCTURE: a stand-in for ICICLE's frontend libraries (libicicle_field/curve/device), built from the
// REFERENCE's own declarations -- bls12-381/include/icicle_backend_api.cuh:98-225 (register_* prototypes, callback
// std::function types) and include/icicle/device_api.h:54-227 (DeviceAPI vtable, register_deviceAPI) -- so that the
// Itanium-mangled names and the vtable layout the product's backend libraries must bind to are pinned by the
// reference, not by this repo's csrc/icicle/icicle_abi.h.  It DEFINES the register_* functions the backend libraries
// import weakly, stashes what they register, and lets a test invoke the stored callbacks through plain C entry points:
//   load this with RTLD_GLOBAL, THEN dlopen libicicle_backend_cuda_{device,field,curve_bls12_381}.so: their static
//   initialisers now find non-null register_* symbols and register "CUDA" (SURVEY.md 3.5 / 8b).
// Built by `make -C oracle mock` into oracle/_ref/libicicle_mock.so (needs /root/reference; the built .so travels).
#include <map>
#include <memory>
#include <string>

#include "icicle_backend_api.cuh"

namespace {
struct Table {
  std::map<std::string, icicle::MsmImpl> msm;
  std::map<std::string, icicle::MsmPreComputeImpl> msm_pre;
  std::map<std::string, icicle::NttImpl> ntt;
  std::map<std::string, icicle::NttInitDomainImpl> ntt_init;
  std::map<std::string, icicle::NttReleaseDomainImpl> ntt_release;
  std::map<std::string, icicle::NttGetRouFromDomainImpl> ntt_rou;
  std::map<std::string, icicle::scalarVectorOpImpl> vec[5];   // add sub mul scalar_mul scalar_add
  icicle::MsmG2Impl g2;
  icicle::MsmG2PreComputeImpl g2_pre;
};
Table& T() { static Table t; return t; }
icicle::Device cuda_device() {
  icicle::Device d;
  d.type = "CUDA";   // this header's Device carries a pointer (icicle_types.cuh:69-72); the callbacks ignore the argument
  d.id = 0;
  return d;
}
}  // namespace

namespace icicle {
void register_msm(const std::string& t, MsmImpl f) { T().msm[t] = std::move(f); }
void register_msm_precompute_bases(const std::string& t, MsmPreComputeImpl f) { T().msm_pre[t] = std::move(f); }
void register_ntt(const std::string& t, NttImpl f) { T().ntt[t] = std::move(f); }
void register_ntt_init_domain(const std::string& t, NttInitDomainImpl f) { T().ntt_init[t] = std::move(f); }
void register_ntt_release_domain(const std::string& t, NttReleaseDomainImpl f) { T().ntt_release[t] = std::move(f); }
void register_ntt_get_rou_from_domain(const std::string& t, NttGetRouFromDomainImpl f) { T().ntt_rou[t] = std::move(f); }
void register_vector_add(const std::string& t, scalarVectorOpImpl f) { T().vec[0][t] = std::move(f); }
void register_vector_sub(const std::string& t, scalarVectorOpImpl f) { T().vec[1][t] = std::move(f); }
void register_vector_mul(const std::string& t, scalarVectorOpImpl f) { T().vec[2][t] = std::move(f); }
void register_scalar_mul_vec(const std::string& t, scalarVectorOpImpl f) { T().vec[3][t] = std::move(f); }
void register_scalar_add_vec(const std::string& t, scalarVectorOpImpl f) { T().vec[4][t] = std::move(f); }
}  // namespace icicle

using scalar = icicle::icicle_scalar_t;
#define HAVE(m) (T().m.count("CUDA") ? 1 : 0)
#define ERR(e) static_cast<int>(e)

extern "C" int mock_device_registered();   // icicle_mock_device.cu
extern "C" {
// bit i set = callback i registered for "CUDA": msm, msm_pre, ntt, ntt_init, ntt_release, ntt_rou, 5 vecops, deviceAPI
int mock_registered_mask() {
  int m = HAVE(msm) | HAVE(msm_pre) << 1 | HAVE(ntt) << 2 | HAVE(ntt_init) << 3 | HAVE(ntt_release) << 4 | HAVE(ntt_rou) << 5;
  for (int i = 0; i < 5; i++) m |= (T().vec[i].count("CUDA") ? 1 : 0) << (6 + i);
  return m | mock_device_registered() << 11;
}
int mock_msm(const void* s, const void* b, int n, const void* cfg, void* r) {
  return ERR(T().msm.at("CUDA")(cuda_device(), (const scalar*)s, (const icicle::icicle_affine_t*)b, n, *(const icicle::MSMConfig*)cfg,
                                 (icicle::icicle_projective_t*)r));
}
int mock_msm_precompute(const void* in, int n, const void* cfg, void* out) {
  return ERR(T().msm_pre.at("CUDA")(cuda_device(), (const icicle::icicle_affine_t*)in, n, *(const icicle::MSMConfig*)cfg,
                                     (icicle::icicle_affine_t*)out));
}
// G2: the backend keeps its own registry (src/backend/g2_registry.cu:84-101); the test hands over the addresses of the
// two getters it found in the curve library (their mangled names come from the same reference prototypes)
typedef icicle::MsmG2Impl (*g2_getter_t)(const std::string&);
typedef icicle::MsmG2PreComputeImpl (*g2_pre_getter_t)(const std::string&);
int mock_fetch_g2(void* get_msm, void* get_pre) {
  T().g2 = ((g2_getter_t)get_msm)("CUDA");
  T().g2_pre = ((g2_pre_getter_t)get_pre)("CUDA");
  const bool unknown_empty = !((g2_getter_t)get_msm)("NO_SUCH_DEVICE");
  return (T().g2 ? 1 : 0) | (T().g2_pre ? 2 : 0) | (unknown_empty ? 4 : 0);
}
int mock_g2_msm(const void* s, const void* b, int n, const void* cfg, void* r) {
  return ERR(T().g2(cuda_device(), (const scalar*)s, (const icicle::icicle_g2_affine_t*)b, n, *(const icicle::MSMConfig*)cfg,
                    (icicle::icicle_g2_projective_t*)r));
}
int mock_g2_msm_precompute(const void* in, int n, const void* cfg, void* out) {
  return ERR(T().g2_pre(cuda_device(), (const icicle::icicle_g2_affine_t*)in, n, *(const icicle::MSMConfig*)cfg,
                        (icicle::icicle_g2_affine_t*)out));
}
int mock_ntt(const void* in, int size, int dir, const void* cfg, void* out) {
  return ERR(T().ntt.at("CUDA")(cuda_device(), (const scalar*)in, size, (icicle::NTTDir)dir,
                                 *(const icicle::NTTConfig<scalar>*)cfg, (scalar*)out));
}
int mock_ntt_init_domain(const void* root, const void* cfg) {
  return ERR(T().ntt_init.at("CUDA")(cuda_device(), *(const scalar*)root, *(const icicle::NTTInitDomainConfig*)cfg));
}
int mock_ntt_release_domain() {
  static const uint64_t phantom[4] = {0, 0, 0, 0};
  return ERR(T().ntt_release.at("CUDA")(cuda_device(), *(const scalar*)phantom));
}
int mock_ntt_get_rou(uint64_t logn, void* rou) { return ERR(T().ntt_rou.at("CUDA")(cuda_device(), logn, (scalar*)rou)); }
int mock_vecop(int which, const void* a, const void* b, uint64_t n, const void* cfg, void* out) {
  return ERR(T().vec[which].at("CUDA")(cuda_device(), (const scalar*)a, (const scalar*)b, n, *(const icicle::VecOpsConfig*)cfg,
                                        (scalar*)out));
}
}
