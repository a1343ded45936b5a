// ICICLE registration of the G1/G2 MSM for device type "CUDA".
// Takes the place of bls12-381/src/backend/icicle_curve_api.cu:243-665: the callbacks forward to the
// C ABI (b381_g1_msm & co. in msm_impl.cuh), which implements the config-flag semantics (host/device
// residency, Montgomery flags, batch, precompute_factor, async).
#include "icicle_abi.h"

using namespace icicle;

static_assert(sizeof(MSMConfig) == sizeof(b381_msm_config) && offsetof(MSMConfig, ext) == offsetof(b381_msm_config, ext) &&
                  offsetof(MSMConfig, is_async) == offsetof(b381_msm_config, is_async),
              "MSMConfig layout drifted from include/b381.h");

static const b381_msm_config* cfg_of(const MSMConfig& c) { return reinterpret_cast<const b381_msm_config*>(&c); }

static eIcicleError msm_g1(const Device&, const scalar_t* scalars, const g1_affine_t* bases, int n, const MSMConfig& cfg,
                           g1_projective_t* results) {
  return to_icicle(b381_g1_msm(reinterpret_cast<const b381_fr*>(scalars), reinterpret_cast<const b381_g1_affine*>(bases), n,
                               cfg_of(cfg), reinterpret_cast<b381_g1_projective*>(results)));
}
static eIcicleError msm_g1_precompute(const Device&, const g1_affine_t* in, int n, const MSMConfig& cfg, g1_affine_t* out) {
  return to_icicle(b381_g1_msm_precompute_bases(reinterpret_cast<const b381_g1_affine*>(in), n, cfg_of(cfg),
                                                reinterpret_cast<b381_g1_affine*>(out)));
}
static eIcicleError msm_g2(const Device&, const scalar_t* scalars, const g2_affine_t* bases, int n, const MSMConfig& cfg,
                           g2_projective_t* results) {
  return to_icicle(b381_g2_msm(reinterpret_cast<const b381_fr*>(scalars), reinterpret_cast<const b381_g2_affine*>(bases), n,
                               cfg_of(cfg), reinterpret_cast<b381_g2_projective*>(results)));
}
static eIcicleError msm_g2_precompute(const Device&, const g2_affine_t* in, int n, const MSMConfig& cfg, g2_affine_t* out) {
  return to_icicle(b381_g2_msm_precompute_bases(reinterpret_cast<const b381_g2_affine*>(in), n, cfg_of(cfg),
                                                reinterpret_cast<b381_g2_affine*>(out)));
}

B381_AT_LOAD(curve) {
  if (register_msm_precompute_bases) register_msm_precompute_bases("CUDA", msm_g1_precompute);
  if (register_msm) register_msm("CUDA", msm_g1);
  register_g2_msm_precompute_bases("CUDA", msm_g2_precompute);
  register_g2_msm("CUDA", msm_g2);
}
