// ICICLE registration of the Fr NTT (+ domain init/release/get-root) and the vector ops for "CUDA".
// Takes the place of bls12-381/src/backend/icicle_field_api.cu:97-352; forwards to the C ABI.
#include "icicle_abi.h"

using namespace icicle;

// scalar_t is an incomplete type on this side of the ABI; its bytes are b381_fr.
struct ntt_config_bytes {   // NTTConfig<scalar_t> with S spelled as b381_fr
  void* stream; b381_fr coset_gen; int batch_size; bool columns_batch; Ordering ordering;
  bool are_inputs_on_device, are_outputs_on_device, is_async; void* ext;
};
static_assert(sizeof(ntt_config_bytes) == sizeof(b381_ntt_config) && offsetof(ntt_config_bytes, ext) == offsetof(b381_ntt_config, ext) &&
                  offsetof(ntt_config_bytes, ordering) == offsetof(b381_ntt_config, ordering),
              "NTTConfig layout drifted from include/b381.h");
static_assert(sizeof(NTTInitDomainConfig) == sizeof(b381_ntt_init_domain_config), "config layout drifted from include/b381.h");

static const b381_fr* fr(const scalar_t* p) { return reinterpret_cast<const b381_fr*>(p); }
static b381_fr* fr(scalar_t* p) { return reinterpret_cast<b381_fr*>(p); }

static eIcicleError ntt_cb(const Device&, const scalar_t* in, int size, NTTDir dir, const NTTConfig<scalar_t>& cfg, scalar_t* out) {
  return to_icicle(b381_ntt(fr(in), size, dir == NTTDir::kInverse ? B381_NTT_INVERSE : B381_NTT_FORWARD,
                            &reinterpret_cast<const b381_ntt_config&>(cfg), fr(out)));
}
static eIcicleError ntt_init_cb(const Device&, const scalar_t& root, const NTTInitDomainConfig& cfg) {
  return to_icicle(b381_ntt_init_domain(fr(&root), reinterpret_cast<const b381_ntt_init_domain_config*>(&cfg)));
}
static eIcicleError ntt_release_cb(const Device&, const scalar_t&) { return to_icicle(b381_ntt_release_domain()); }
static eIcicleError ntt_rou_cb(const Device&, uint64_t logn, scalar_t* rou) { return to_icicle(b381_ntt_get_rou_from_domain(logn, fr(rou))); }

#ifdef B381_ICICLE_UPSTREAM_VECOPS
// upstream layout: copy the common members, honour batch_size / columns_batch.  Element-wise ops on `batch_size`
// vectors of n elements are one op on n * batch_size elements whatever the layout (all operands share it).
static b381_vecops_config vcfg(const VecOpsConfig& c) {
  b381_vecops_config k;
  k.stream = c.stream; k.is_a_on_device = c.is_a_on_device; k.is_b_on_device = c.is_b_on_device;
  k.is_result_on_device = c.is_result_on_device; k.is_async = c.is_async; k.ext = nullptr;
  return k;
}
static uint64_t batch_of(const VecOpsConfig& c) { return c.batch_size > 1 ? (uint64_t)c.batch_size : 1u; }
#define VEC_CB(name, fn)                                                                                              \
  static eIcicleError name(const Device&, const scalar_t* a, const scalar_t* b, uint64_t n, const VecOpsConfig& cfg, \
                           scalar_t* out) {                                                                           \
    const b381_vecops_config k = vcfg(cfg);                                                                           \
    return to_icicle(fn(fr(a), fr(b), n * batch_of(cfg), &k, fr(out)));                                               \
  }
#define SCALAR_CB(name, fn)                                                                                           \
  static eIcicleError name(const Device&, const scalar_t* a, const scalar_t* b, uint64_t n, const VecOpsConfig& cfg, \
                           scalar_t* out) {                                                                           \
    const b381_vecops_config k = vcfg(cfg);                                                                           \
    return to_icicle(fn##_batch(fr(a), fr(b), n, (int)batch_of(cfg), cfg.columns_batch, &k, fr(out)));                \
  }
#else
static_assert(sizeof(VecOpsConfig) == sizeof(b381_vecops_config), "config layout drifted from include/b381.h");
static const b381_vecops_config* vcfg(const VecOpsConfig& c) { return reinterpret_cast<const b381_vecops_config*>(&c); }
#define VEC_CB(name, fn)                                                                                              \
  static eIcicleError name(const Device&, const scalar_t* a, const scalar_t* b, uint64_t n, const VecOpsConfig& cfg, \
                           scalar_t* out) {                                                                           \
    return to_icicle(fn(fr(a), fr(b), n, vcfg(cfg), fr(out)));                                                        \
  }
#define SCALAR_CB(name, fn) VEC_CB(name, fn)
#endif
VEC_CB(vec_add_cb, b381_vector_add)
VEC_CB(vec_sub_cb, b381_vector_sub)
VEC_CB(vec_mul_cb, b381_vector_mul)
SCALAR_CB(scalar_mul_cb, b381_scalar_mul_vec)
SCALAR_CB(scalar_add_cb, b381_scalar_add_vec)

B381_AT_LOAD(field) {
  if (register_ntt) register_ntt("CUDA", ntt_cb);
  if (register_ntt_init_domain) register_ntt_init_domain("CUDA", ntt_init_cb);
  if (register_ntt_release_domain) register_ntt_release_domain("CUDA", ntt_release_cb);
  if (register_ntt_get_rou_from_domain) register_ntt_get_rou_from_domain("CUDA", ntt_rou_cb);
  if (register_vector_add) register_vector_add("CUDA", vec_add_cb);
  if (register_vector_sub) register_vector_sub("CUDA", vec_sub_cb);
  if (register_vector_mul) register_vector_mul("CUDA", vec_mul_cb);
  if (register_scalar_mul_vec) register_scalar_mul_vec("CUDA", scalar_mul_cb);
  if (register_scalar_add_vec) register_scalar_add_vec("CUDA", scalar_add_cb);
}
