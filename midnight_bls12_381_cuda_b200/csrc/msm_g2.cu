// C ABI of the G2 MSM (include/b381.h).  Templates in msm_impl.cuh.
#include "msm_impl.cuh"

using namespace b381;

static_assert(sizeof(b381_g2_affine) == sizeof(g2_affine) && sizeof(b381_g2_projective) == sizeof(g2_jac), "layout");

static b381_msm_config flat_cfg(const b381_msm_config* cfg) {   // see msm_g1.cu
  b381_msm_config c = *cfg;
  c.are_points_montgomery_form = true;
  c.are_scalars_montgomery_form = false;
  c.batch_size = 1;
  c.precompute_factor = 1;
  return c;
}

extern "C" {

int b381_g2_msm(const b381_fr* s, const b381_g2_affine* p, int n, const b381_msm_config* cfg, b381_g2_projective* r) {
  return msm_entry<fq2_t>((const fr_t*)s, (const g2_affine*)p, n, cfg, r, ResultKind::IcicleStd);
}
int bls12_381_g2_msm_cuda(const b381_fr* s, const b381_g2_affine* p, int n, const b381_msm_config* cfg,
                          b381_g2_projective* r) {
  if (!cfg) return B381_INVALID_POINTER;
  b381_msm_config c = flat_cfg(cfg);
  return msm_entry<fq2_t>((const fr_t*)s, (const g2_affine*)p, n, &c, r, ResultKind::JacobianMont);
}
int b381_g2_msm_partial(const b381_fr* s, const b381_g2_affine* p, int n, const b381_msm_config* cfg, void* out) {
  return msm_entry<fq2_t>((const fr_t*)s, (const g2_affine*)p, n, cfg, out, ResultKind::PartialXyzz);
}
int b381_g2_msm_combine(const void* parts, int count, void* stream, bool on_device, b381_g2_projective* r) {
  return combine_entry<fq2_t>(parts, count, stream, on_device, r);
}
int b381_g2_msm_precompute_bases(const b381_g2_affine* in, int n, const b381_msm_config* cfg, b381_g2_affine* out) {
  return precompute_entry<fq2_t>((const g2_affine*)in, n, cfg, (g2_affine*)out);
}

}  // extern "C"
