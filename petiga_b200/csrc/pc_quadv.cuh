// pc_quadv.cuh -- quad_vec3_kernel<P>: vector-only assembly (IGAComputeVector, and the load vector of IGAComputeSystem when the
// separable path has already written the matrix: BASELINE cfg 3, demo/L2Projection.c) by sum factorisation, one CTA per element.
//
// Reference semantics: the vector half of the element loop -- IGAElementBuildTabulation (src/petigaelem.c:794-1033), the user's
// point callback F[a] += N_a f(x) JW (demo/L2Projection.c:63-78) through IGAPointAddVec (src/petigapoint.c:467-492),
// IGAElementFixSystem's vector part (:1365-1387) and IGAElementAssembleVec (:1543-1559).  3-D, one dof per node, one degree P on
// all axes with the default P + 1 point rule, first-order forms without state, identity or mapped (non-rational) geometry.
// Before this kernel the hybrid path ran the matrix kernel's full set-up for the vector alone: 11.3 ms at cfg 3 against 0.27 ms
// for the matrix.  Work per element here: 3 x 2 N^4 flops for the geometry + 3 x 2 N^4 per tensor component, N = P + 1.
#pragma once
#include "pc_quad3.cuh"

namespace pc {

// MAPPED is a template parameter so that the identity-geometry instantiation carries no geometry scratch: 13 KB instead of 35 KB of
// shared memory at p = 4, i.e. 16 instead of 6 resident CTAs per SM for a kernel that is latency bound (table loads, 5 barriers)
template <int P, bool MAPPED>
__global__ void __launch_bounds__(((P + 1) * (P + 1) * (P + 1) + 31) / 32 * 32) quad_vec3_kernel(const __grid_constant__ SF3Params sp) {
  constexpr int N = P + 1, NN = N * N, NNN = N * N * N, T = (NNN + 31) / 32 * 32;
  const KParams& prm = sp.k;
  const SFLists& ls = sp.l;
  __shared__ double gB[3 * 2 * NN], wJ[3 * N], pt[3 * N], Fp[4 * NNN];
  __shared__ double Xs[MAPPED ? 3 * NNN : 1], Ev[MAPPED ? 3 * 4 * NNN : 1];
  __shared__ double T1[MAPPED ? 3 * 2 * N * NN : 4 * NNN], T2[MAPPED ? 3 * 3 * NNN : 4 * NNN];   // (also R1 / R2 of the transposed stages)
  const int gt = threadIdx.x;
  const int NV = prm.vc1 - prm.vc0, NT = ls.NT;
  constexpr bool mapped = MAPPED;
  const int ID[3] = {(int)blockIdx.x + prm.ax[0].es, (int)blockIdx.y + prm.ax[1].es, (int)blockIdx.z + prm.ax[2].es};   // 3-D grid: no divisions
  for (int t = gt; t < 3 * 2 * NN; t += T) {
    const int d = t / (2 * NN), r = t % (2 * NN), o = r / NN, q = (r / N) % N, a = r % N;
    gB[t] = prm.ax[d].value[((size_t)(ID[d] * N + q) * N + a) * 5 + o];                 // Bt[d][o][q][a]
  }
  if (gt < 3 * N) {
    const int d = gt / N, q = gt % N;
    wJ[gt] = prm.ax[d].weight[ID[d] * N + q] * prm.ax[d].detJac[ID[d]];
    pt[gt] = prm.ax[d].point[ID[d] * N + q];
  }
  const int a = gt, ai[3] = {a % N, (a / N) % N, a / NN};
  int gidx = 0;
  if (a < NNN) {
    int mul = 1;
#pragma unroll
    for (int d = 0; d < 3; d++) { gidx += (prm.ax[d].offset[ID[d]] + ai[d] - prm.ax[d].gs) * mul; mul *= prm.ax[d].gw; }
    if (mapped) {
#pragma unroll
      for (int i = 0; i < 3; i++) Xs[i * NNN + a] = prm.X[(size_t)gidx * 3 + i];
    }
  }
  __syncthreads();
  if (mapped) {   // X and dX/du at the points (petigamapgeo.f90.in:28-43), one axis at a time
    for (int t = gt; t < 3 * 2 * N * NN; t += T) {            // T1[i][o0][q0][a12]
      const int i = t / (2 * N * NN), r = t % (2 * N * NN), o0 = r / (N * NN), q0 = (r / NN) % N, a12 = r % NN;
      const double* b = gB + o0 * NN + q0 * N;
      const double* x = Xs + i * NNN + a12 * N;
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < N; k++) acc = fma(b[k], x[k], acc);
      T1[t] = acc;
    }
    __syncthreads();
    for (int t = gt; t < 3 * 3 * NNN; t += T) {               // T2[i][oc][q0][q1][a2], oc: 0 = (1,0), 1 = (0,1), 2 = (0,0)
      const int i = t / (3 * NNN), r = t % (3 * NNN), oc = r / NNN, q0 = (r / NN) % N, q1 = (r / N) % N, a2 = r % N;
      const int o0 = (oc == 0), o1 = (oc == 1);
      const double* b = gB + 2 * NN + o1 * NN + q1 * N;
      const double* s = T1 + i * (2 * N * NN) + o0 * (N * NN) + q0 * NN + a2 * N;
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < N; k++) acc = fma(b[k], s[k], acc);
      T2[t] = acc;
    }
    __syncthreads();
    for (int t = gt; t < 3 * 4 * NNN; t += T) {               // Ev[i][d][q], d = 3: the point itself
      const int i = t / (4 * NNN), r = t % (4 * NNN), d = r / NNN, q = r % NNN, q0 = q % N, q1 = (q / N) % N, q2 = q / NN;
      const int oc = (d == 0) ? 0 : (d == 1 ? 1 : 2), o2 = (d == 2);
      const double* b = gB + 4 * NN + o2 * NN + q2 * N;
      const double* s = T2 + i * (3 * NNN) + oc * NNN + q0 * NN + q1 * N;
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < N; k++) acc = fma(b[k], s[k], acc);
      Ev[t] = acc;
    }
    __syncthreads();
  }
  if (gt < NNN) {  // one thread per quadrature point: inverse map (petigamapinv.f90.in:28-31), weights, the form's vector coefficient
    const int q = gt, q0 = q % N, q1 = (q / N) % N, q2 = q / NN;
    double E[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, x[3] = {pt[q0], pt[N + q1], pt[2 * N + q2]};
    double jw = wJ[q0] * wJ[N + q1] * wJ[2 * N + q2];
    if (mapped) {
      double X1[3][3];
#pragma unroll
      for (int i = 0; i < 3; i++) {
#pragma unroll
        for (int d = 0; d < 3; d++) X1[i][d] = Ev[i * (4 * NNN) + d * NNN + q];
        x[i] = Ev[i * (4 * NNN) + 3 * NNN + q];
      }
      const double a00 = X1[0][0], a01 = X1[0][1], a02 = X1[0][2], a10 = X1[1][0], a11 = X1[1][1], a12 = X1[1][2], a20 = X1[2][0], a21 = X1[2][1], a22 = X1[2][2];
      const double det = a00 * (a11 * a22 - a12 * a21) - a01 * (a10 * a22 - a12 * a20) + a02 * (a10 * a21 - a11 * a20);
      E[0][0] = (a11 * a22 - a12 * a21) / det; E[0][1] = -(a01 * a22 - a02 * a21) / det; E[0][2] = (a01 * a12 - a02 * a11) / det;
      E[1][0] = -(a10 * a22 - a12 * a20) / det; E[1][1] = (a00 * a22 - a02 * a20) / det; E[1][2] = -(a00 * a12 - a02 * a10) / det;
      E[2][0] = (a10 * a21 - a11 * a20) / det; E[2][1] = -(a00 * a21 - a01 * a20) / det; E[2][2] = (a00 * a11 - a01 * a10) / det;
      jw *= det;                                                 // detJac *= detX (petigaelem.c:1024-1029)
    }
    double fv[4] = {sp.fconst[0], sp.fconst[1], sp.fconst[2], sp.fconst[3]};
    if (prm.per_qp) {
      QPoint qp;
      qp.atboundary = 0;
      qp.x[0] = x[0]; qp.x[1] = x[1]; qp.x[2] = x[2];
      fv[0] = fv[1] = fv[2] = fv[3] = 0.0;
      form_coefficients<3, 1>(prm.form, prm.slot, prm.prm, prm.shift, prm.t, qp, 0, NV, nullptr, fv);
    }
    // f'[s][q] = JW sum_al A[vc0 + al][s] f[al],  A[0][tN] = 1, A[1 + i][tG_d] = E[d][i]
    if (!MAPPED) {   // E = I: component ca feeds exactly one tensor slot
#pragma unroll
      for (int s = 0; s < 4; s++) {
        double acc = 0.0;
#pragma unroll
        for (int al = 0; al < 4; al++) {
          const int ca = prm.vc0 + al, ts = (ca == 0) ? ls.tN : (ca <= 3 ? ls.tG[ca - 1] : -1);
          if (al < NV && ts == s) acc += fv[al];
        }
        if (s < NT) Fp[s * NNN + q] = acc * jw;
      }
    } else {
#pragma unroll
      for (int s = 0; s < 4; s++) {
        double acc = 0.0;
#pragma unroll
        for (int al = 0; al < 4; al++) {
          const int ca = prm.vc0 + al;
          double as = 0.0;
          if (ca == 0) as = (s == ls.tN) ? 1.0 : 0.0;
          else if (ca <= 3) {
#pragma unroll
            for (int d = 0; d < 3; d++) if (s == ls.tG[d]) as = E[d][ca - 1];
          }
          if (al < NV) acc = fma(as, fv[al], acc);
        }
        if (s < NT) Fp[s * NNN + q] = acc * jw;
      }
    }
  }
  __syncthreads();
  // ---- element vector by the transposed sum factorisation ----
  double* R1 = T1;                                             // [s][q2][q1][a0]
  double* R2 = T2;                                             // [s][q2][a1][a0]
  for (int t = gt; t < NT * NNN; t += T) {
    const int s = t / NNN, r = t % NNN, q12 = r / N, a0 = r % N, o = ls.torder[s][0];
    if (!((sp.vslots >> s) & 1)) continue;                       // tensor slots the load never feeds
    const double* f = Fp + s * NNN + q12 * N;
    const double* b = gB + o * NN + a0;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < N; k++) acc = fma(b[k * N], f[k], acc);
    R1[t] = acc;
  }
  __syncthreads();
  for (int t = gt; t < NT * NNN; t += T) {
    const int s = t / NNN, r = t % NNN, q2 = r / NN, a1 = (r / N) % N, a0 = r % N, o = ls.torder[s][1];
    if (!((sp.vslots >> s) & 1)) continue;
    const double* b = gB + 2 * NN + o * NN + a1;
    const double* x = R1 + s * NNN + q2 * NN + a0;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < N; k++) acc = fma(b[k * N], x[k * N], acc);
    R2[t] = acc;
  }
  __syncthreads();
  if (a >= NNN) return;
  double F = 0.0;
  {
    const int a2 = a / NN, a01 = a % NN;
    for (int s = 0; s < NT; s++) {
      if (!((sp.vslots >> s) & 1)) continue;
      const int o = ls.torder[s][2];
      const double* b = gB + 4 * NN + o * NN + a2;
      const double* x = R2 + s * NNN + a01;
#pragma unroll
      for (int k = 0; k < N; k++) F = fma(b[k * N], x[k * NN], F);
    }
  }
  if (prm.slot == PETIGA_SLOT_SYSTEM && sf3_elem_on_bc(prm, ID, false)) {          // FixSystem vector part (petigaelem.c:1365-1387)
    int onfix; double vfix, vflux;
    sf3_node_bc(prm, ID, ai, gidx, onfix, vfix, vflux, P);
    F += vflux;
    if (onfix) F = vfix;
  }
  if (F != 0.0) atomicAdd(&prm.rhs[prm.localrow[gidx]], F);
}

}  // namespace pc
