#include "petiga_cuda.h"
#include "petiga_host.h"
int main(void) {
  IGA iga; Mat A; Vec b; IGAAxis ax; IGAComm comm = {0, 1, NULL, 0};
  IGACreate(comm, &iga); IGASetDim(iga, 3); IGASetDof(iga, 1);
  for (int d = 0; d < 3; d++) { IGAGetAxis(iga, d, &ax); IGAAxisSetDegree(ax, 3); IGAAxisInitUniform(ax, 8, 0.0, 1.0, 2);
    for (int s = 0; s < 2; s++) IGASetBoundaryValue(iga, d, s, 0, 1.0); }
  IGASetUp(iga);
  IGASetFormSystem(iga, IGADeviceForm_Poisson_System, NULL);
  int rc = IGACreateMat(iga, &A);          /* needs a device: PETSC_ERR_LIB-style failure without one, never a CPU fallback */
  (void)b;
  IGADestroy(&iga);
  return rc == 0 ? 0 : 3;
}
