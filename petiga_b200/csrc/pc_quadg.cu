// pc_quadg.cu -- launcher of the generic runtime-degree quadrature kernel (pc_quadg.cuh): interior pass and face mode.
#include <algorithm>
#include <cstring>

#include "pc_plan.h"
#include "pc_quadg.cuh"

namespace pc {

namespace {

template <int DIM>
int launch_gen(petiga_cuda_plan* Pl, const KParams& base, const FormInfo& fi) {
  GenParams gp;
  memset(&gp, 0, sizeof(gp));
  gp.k = base;
  const int dof = base.dof;
  int nen = 1, nqp = 1;
  for (int d = 0; d < 3; d++) {
    if (base.ax[d].nen > kGenMaxN1) { set_error("generic kernel: degree > 8"); return PETIGA_CUDA_ERR_SUP; }
    nen *= base.ax[d].nen;
    nqp *= (d == base.face_axis) ? 1 : base.ax[d].nqp;
  }
  if (dof > kMaxDof) { set_error("generic kernel: more than 8 dofs per node"); return PETIGA_CUDA_ERR_SUP; }
  const bool mapped = base.X != nullptr, rational = base.Wt != nullptr;
  const bool order2 = fi.order >= 2;
  gp.ncomp = DIM + 1 + (order2 ? 1 : 0);
  gp.full2 = (order2 && (mapped || rational)) ? 1 : 0;
  gp.nct = gp.full2 ? 1 + DIM + gen_nsym(DIM) : gp.ncomp;
  const int NA = base.mc1 - base.mc0, NV = base.vc1 - base.vc0;
  const int want_mat = (NA > 0 && slot_has_mat(base.slot)) ? 1 : 0;
  const int R = nen * dof;
  const size_t budget = 224 * 1024;
  auto bytes = [&](int qc, int pw) { return (size_t)GenSmem(nen, dof, DIM, gp.nct, NA, NV, qc, pw, want_mat).total * 8; };
  int pw = R;
  // keep at least half of the budget for the quadrature-point chunk when the element matrix is large
  while (pw > 1 && (bytes(1, pw) > budget || (size_t)R * pw * 8 > budget * 5 / 8)) pw = (pw + 1) / 2;
  if (bytes(1, pw) > budget) { set_error("generic kernel: element does not fit shared memory"); return PETIGA_CUDA_ERR_SUP; }
  int qc = std::min(nqp, 32);
  while (qc > 1 && bytes(qc, pw) > budget) qc--;
  gp.pw = pw;
  gp.k.qc = qc;
  int blocks;
  if (base.face_axis < 0) blocks = base.nelem;
  else {
    blocks = 1;
    for (int d = 0; d < 3; d++) if (d != base.face_axis) blocks *= base.ax[d].ew;
  }
  const size_t smem = bytes(qc, pw);
  auto kern = quad_gen_kernel<DIM>;
  PC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (blocks > 0) {
    kern<<<blocks, kGenThreads, smem, Pl->stream>>>(gp);
    PC_CUDA(cudaGetLastError());
    Pl->launches++;
  }
  // FP64 operations executed: the reference's loop nest (contraction + T-builder + vector), times the panel passes of the tabulation
  const double npan = want_mat ? (double)((R + pw - 1) / pw) : 1.0;
  Pl->last_flops += (double)blocks * nqp * (2.0 * NA * (double)R * R + 2.0 * NA * NA * nen * dof * dof * npan + 2.0 * NV * R);
  return 0;
}

}  // namespace

int launch_quadrature_gen(petiga_cuda_plan* Pl, const KParams& base) {
  FormInfo fi = form_info(base.form, base.slot, base.dim, base.dof);
  if (!fi.valid) return PETIGA_CUDA_ERR_SUP;
  switch (base.dim) {
    case 1: return launch_gen<1>(Pl, base, fi);
    case 2: return launch_gen<2>(Pl, base, fi);
    case 3: return launch_gen<3>(Pl, base, fi);
  }
  return PETIGA_CUDA_ERR_ARG;
}

}  // namespace pc
