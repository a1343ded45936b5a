// TEST INFRASTRUCTURE: C entry point over the reference's OWN G1 MSM (msm::msm_cuda<Fr, G1Affine, G1Projective>,
// bls12-381/src/curve/msm_kernels.cu:603-903), compiled from where the sources lie into oracle/_ref/libref_msm.so
// (oracle/Makefile, target `ref_msm`).  Same name and signature as the reference's flat test API
// (src/backend/icicle_curve_api.cu:679-692) -- that file is not linked because its REGISTER_* initialisers need
// libicicle.  Used only by tests/ and by bench.py's `reference_gpu` leg (the R-GPU comparator of SURVEY.md 2.2).
#include "msm.cuh"

extern "C" int bls12_381_g1_msm_cuda(const bls12_381::Fr* scalars, const bls12_381::G1Affine* bases, int msm_size,
                                     const icicle::MSMConfig* config, bls12_381::G1Projective* result) {
  cudaError_t err = msm::msm_cuda<bls12_381::Fr, bls12_381::G1Affine, bls12_381::G1Projective>(scalars, bases, msm_size,
                                                                                                 *config, result);
  return err == cudaSuccess ? 0 : 14;   // eIcicleError::SUCCESS / UNKNOWN_ERROR, as the reference maps it
}
