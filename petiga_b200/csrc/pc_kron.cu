// pc_kron.cu -- separable ("Kronecker") assembly path. (stub; filled in below)
#include "pc_plan.h"
namespace pc {
bool kron_applicable(const petiga_cuda_plan*, int, int) { return false; }
int launch_kronecker(petiga_cuda_plan*, int, int, double*, double*) { set_error("separable path not built"); return PETIGA_CUDA_ERR_SUP; }
}  // namespace pc
