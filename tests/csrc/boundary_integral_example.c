/* main() of demo/BoundaryIntegral.c:138-200 against the host mirror: the IGAForm object API, a Dirichlet value on one end of
   the axis and the boundary-integral pass on the other.  Exit code: 0 assembled on a GPU, 3 stopped at IGACreateMat without
   one (there is no CPU fallback), anything else = an API call failed. */
#include "petiga_host.h"

int main(void) {
  IGA iga; IGAForm form; IGAAxis ax; Mat A; Vec x, b; IGAComm comm = {0, 1, NULL, 0};
  PetscInt axis = 0, side = 1, dim = 2, i;
  if (IGACreate(comm, &iga)) return 10;
  if (IGASetDof(iga, 1)) return 11;
  if (IGASetDim(iga, dim)) return 12;
  for (i = 0; i < dim; i++) {
    if (IGAGetAxis(iga, i, &ax)) return 13;
    if (IGAAxisSetDegree(ax, 2)) return 14;
    if (IGAAxisInitUniform(ax, 8, 0.0, 1.0, PETSC_DECIDE)) return 15;
  }
  if (IGAGetForm(iga, &form)) return 16;
  {
    PetscInt d = !side, n = !d;
    if (IGAFormSetSystem(form, IGADeviceForm_BoundaryIntegral_System, NULL)) return 17;
    if (IGAFormSetBoundaryValue(form, axis, d, 0, 1.0)) return 18;
    if (IGAFormSetBoundaryForm(form, axis, n, PETSC_TRUE)) return 19;
    if (IGAFormSetBoundaryLoad(form, 1, 0, 0, 5.0)) return 20;
    if (IGAFormClearBoundary(form, 1, 0)) return 21;              /* and take it back */
    if (IGAFormSetBoundaryForm(form, 3, 0, PETSC_TRUE) != PETSC_ERR_ARG_OUTOFRANGE) return 22;   /* IGAFormCheckArg */
  }
  if (IGASetUp(iga)) return 23;
  if (IGACreateMat(iga, &A)) { IGADestroy(&iga); return 3; }
  if (IGACreateVec(iga, &x) || IGACreateVec(iga, &b)) return 24;
  if (IGAComputeSystem(iga, A, b)) return 25;
  {
    PetscReal norm[1];
    if (IGAComputeErrorNorm(iga, 0, b, NULL, norm, NULL)) return 26;   /* any functional call: ||b_h||_L2 */
  }
  MatDestroy(&A); VecDestroy(&x); VecDestroy(&b); IGADestroy(&iga);
  return 0;
}
