// pc_quad.cu -- instantiation table and launcher of the per-element quadrature kernel.
#include <algorithm>
#include <cstring>
#include <vector>

#include "pc_plan.h"
#include "pc_quad.cuh"

namespace pc {

namespace {

struct LaunchCfg { int qc, epb, threads; size_t smem; };

template <int DIM, int P, int DOF, int TM>
int launch_one(petiga_cuda_plan* Pl, KParams prm) {
  using Cfg = QCfg<DIM, P, DOF, TM>;
  constexpr int G = Cfg::G;
  static_assert(G <= 1024, "element group too large");
  const int NC = prm.c1 - prm.c0, NA = prm.mc1 - prm.mc0, NV = prm.vc1 - prm.vc0;
  int nqp = 1;
  for (int d = 0; d < 3; d++) nqp *= prm.ax[d].nqp;
  // pick the chunk size and the elements per block so that two CTAs fit an SM when possible
  const size_t budget = 100 * 1024, hard = 220 * 1024;
  int qc = std::min(nqp, 32), epb = std::max(1, 256 / G);
  auto bytes = [&](int q, int e) { return (size_t)QSmem(Cfg::M, Cfg::N, DIM, DOF, NC, NA, NV, q, Cfg::NEN1, G).total * 8 * e; };
  while (qc > 1 && bytes(qc, epb) > budget) qc = (qc + 1) / 2;
  while (epb > 1 && bytes(qc, epb) > budget) epb--;
  if (bytes(qc, epb) > hard) { set_error("quadrature kernel: element does not fit shared memory"); return PETIGA_CUDA_ERR_SUP; }
  prm.qc = qc;
  prm.epb = epb;
  const size_t smem = bytes(qc, epb);
  const int threads = ((G * epb + 31) / 32) * 32;
  auto kern = quad_kernel<DIM, P, DOF, TM>;
  PC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int blocks = (prm.nelem + epb - 1) / epb;
  // FP64 operations executed: the contraction K_e += Psi^T T (2 per FMA), the T-builder and the element vector
  Pl->last_flops = (double)prm.nelem * nqp * (2.0 * NA * Cfg::M * Cfg::N + 2.0 * NA * NA * Cfg::N + 2.0 * NV * Cfg::M * DOF);
  if (blocks > 0) {
    kern<<<blocks, threads, smem, Pl->stream>>>(prm);
    PC_CUDA(cudaGetLastError());
    Pl->launches++;
  }
  return 0;
}

}  // namespace

#define PC_CASE(DIM_, P_, DOF_, TM_) \
  if (dim == DIM_ && p == P_ && dof == DOF_) return launch_one<DIM_, P_, DOF_, TM_>(Pl, base);

int launch_quadrature(petiga_cuda_plan* Pl, const KParams& base) {
  const int dim = base.dim, dof = base.dof, p = base.ax[0].p;
  for (int d = 1; d < dim; d++)
    if (base.ax[d].p != p) { set_error("quadrature kernel: mixed degrees per axis are not instantiated"); return PETIGA_CUDA_ERR_SUP; }
  PC_CASE(1, 1, 1, 2) PC_CASE(1, 2, 1, 3) PC_CASE(1, 3, 1, 4) PC_CASE(1, 4, 1, 5)
  PC_CASE(1, 1, 2, 2) PC_CASE(1, 2, 2, 3) PC_CASE(1, 3, 2, 4) PC_CASE(1, 4, 2, 5)
  PC_CASE(1, 1, 3, 2) PC_CASE(1, 2, 3, 3) PC_CASE(1, 3, 3, 4) PC_CASE(1, 4, 3, 5)
  PC_CASE(2, 1, 1, 2) PC_CASE(2, 2, 1, 3) PC_CASE(2, 3, 1, 4) PC_CASE(2, 4, 1, 5)
  PC_CASE(2, 1, 2, 2) PC_CASE(2, 2, 2, 3) PC_CASE(2, 3, 2, 4) PC_CASE(2, 4, 2, 5)
  PC_CASE(2, 1, 3, 2) PC_CASE(2, 2, 3, 3) PC_CASE(2, 3, 3, 4) PC_CASE(2, 4, 3, 5)
  PC_CASE(3, 1, 1, 2) PC_CASE(3, 2, 1, 3) PC_CASE(3, 3, 1, 4) PC_CASE(3, 4, 1, 5)
  PC_CASE(3, 1, 2, 2) PC_CASE(3, 2, 2, 3) PC_CASE(3, 3, 2, 4)
  PC_CASE(3, 1, 3, 2) PC_CASE(3, 2, 3, 3)
  set_error("quadrature kernel: (dim, degree, dof) combination not instantiated");
  return PETIGA_CUDA_ERR_SUP;
}


}  // namespace pc
