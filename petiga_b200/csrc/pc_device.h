// pc_device.h -- device-resident plan tables and kernel parameter blocks (internal).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "pc_forms.cuh"
#include "pc_layout.h"

namespace pc {

struct DevAxis {
  const double* value;    // [nel][nqp][nen][5]   IGABasis.value (include/petiga.h:122-141)
  const double* weight;   // [nel][nqp]
  const double* point;    // [nel][nqp]
  const double* detJac;   // [nel]
  const int* offset;      // [nel]
  const int* W;           // [gw]  1-D row width by ghost coordinate
  const int* lo;          // [gw]  row coordinate - first column
  const uint32_t* seg;    // [gw][kMaxW] packed (B | S<<8 | L<<16)
  const int* simple;      // [gw]  1: columns of this row are already in storage order
  int nel, nqp, nen, p, gs, gw, es, ew, periodic, nnp;
};

struct FixSide {          // IGAFormBC value/load of one face, restricted to fields < dof
  int vcount, lcount;
  int vfield[kMaxDof];
  double vvalue[kMaxDof];
  int lfield[kMaxDof];
  double lvalue[kMaxDof];
};

struct KParams {
  DevAxis ax[3];
  int dim, dof;
  int nelem;                       // local elements
  int nown;                        // owned rows
  int64_t nnz_own;                 // blocks in owned rows
  const int* localrow;             // [ghost box] -> local row
  const int64_t* rowbase;          // [nloc+1]
  double* values;                  // owned rows (caller's array)
  double* ghost_values;            // ghost rows (plan's buffer), indexed from block nnz_own
  double* rhs;                     // [nloc*dof] unified local vector (owned first)
  const double* U;                 // [nloc*dof] unified local state (owned first) or NULL
  const double* V;
  const double* X;                 // geometry [ghost box][dim] or NULL
  const double* Wt;                // rational weights [ghost box] or NULL
  const double* fixtable;          // [ghost box][dof] or NULL
  FixSide bc[3][2];
  int any_bc;
  int form, slot, block;           // block: 1 = BAIJ value layout, 0 = AIJ
  int mc0, mc1, vc0, vc1, per_qp, needs_x, needs_state;
  int c0, c1;                      // tabulated component range (union)
  double prm[8];
  double shift, t;
  int qc;                          // quadrature points per chunk
  int epb;                         // elements per block
};

// BoundaryArea of an element face on a mapped geometry: sum over the face's quadrature points of the surface Jacobian
// sqrt|det(F F^T)|, F = d x / d(face parameters), times the weights (src/petigaelem.c:1132-1162 ->
// IGA_BoundaryArea_{2,3}D, src/petiga{2,3}d.F90; Rationalize + Jacobian there).  The rationalised sums are folded:
// F[r][s] = (Q[r][s] - S1[r] P[s] / W0) / W0 with W0 = sum W N0, S1 = sum W N1, P = sum W N0 X, Q = sum W N1 X.
template <int DIM>
__device__ inline double face_area_factor(const DevAxis* ax, const int* ID, int dir, int side, const double* __restrict__ X,
                                          const double* __restrict__ Wt) {
  if (DIM == 1) return 1.0;
  int fa[2] = {0, 0}, n = 0;
  for (int i = 0; i < DIM; i++) if (i != dir) fa[n++] = i;
  const int sd = DIM - 1;
  const DevAxis& A0 = ax[fa[0]];
  const DevAxis& A1 = ax[(sd > 1) ? fa[1] : fa[0]];
  const int ne0 = A0.nen, ne1 = (sd > 1) ? A1.nen : 1, nq0 = A0.nqp, nq1 = (sd > 1) ? A1.nqp : 1;
  const int kfix = side ? ax[dir].nen - 1 : 0;
  int g[3] = {0, 0, 0};
  g[dir] = ax[dir].offset[ID[dir]] + kfix - ax[dir].gs;
  const int b0 = A0.offset[ID[fa[0]]] - A0.gs, b1 = (sd > 1) ? A1.offset[ID[fa[1]]] - A1.gs : 0;
  double dS = 0.0;
  for (int jq = 0; jq < nq1; jq++)
    for (int iq = 0; iq < nq0; iq++) {
      double W0 = 0.0, S1[2] = {0, 0}, P[3] = {0, 0, 0}, Q[2][3] = {{0, 0, 0}, {0, 0, 0}};
      for (int ja = 0; ja < ne1; ja++) {
        const double j0 = (sd > 1) ? A1.value[((size_t)(ID[fa[1]] * nq1 + jq) * ne1 + ja) * 5] : 1.0;
        const double j1 = (sd > 1) ? A1.value[((size_t)(ID[fa[1]] * nq1 + jq) * ne1 + ja) * 5 + 1] : 0.0;
        if (sd > 1) g[fa[1]] = b1 + ja;
        for (int ia = 0; ia < ne0; ia++) {
          const double i0 = A0.value[((size_t)(ID[fa[0]] * nq0 + iq) * ne0 + ia) * 5];
          const double i1 = A0.value[((size_t)(ID[fa[0]] * nq0 + iq) * ne0 + ia) * 5 + 1];
          g[fa[0]] = b0 + ia;
          const int gidx = g[0] + ax[0].gw * (g[1] + ax[1].gw * g[2]);
          const double w = Wt ? Wt[gidx] : 1.0;
          const double N0 = w * i0 * j0, N1a = w * i1 * j0, N1b = w * i0 * j1;
          W0 += N0; S1[0] += N1a; S1[1] += N1b;
#pragma unroll
          for (int s = 0; s < DIM; s++) {
            const double x = X[(size_t)gidx * DIM + s];
            P[s] = fma(N0, x, P[s]); Q[0][s] = fma(N1a, x, Q[0][s]); Q[1][s] = fma(N1b, x, Q[1][s]);
          }
        }
      }
      double F[2][3];
#pragma unroll
      for (int r = 0; r < 2; r++)
#pragma unroll
        for (int s = 0; s < DIM; s++) F[r][s] = Wt ? (Q[r][s] - S1[r] * P[s] / W0) / W0 : Q[r][s];
      double m00 = 0, m01 = 0, m11 = 0;
#pragma unroll
      for (int s = 0; s < DIM; s++) { m00 = fma(F[0][s], F[0][s], m00); m01 = fma(F[0][s], F[1][s], m01); m11 = fma(F[1][s], F[1][s], m11); }
      const double det = (sd > 1) ? m00 * m11 - m01 * m01 : m00;
      double wq = A0.weight[ID[fa[0]] * nq0 + iq];
      if (sd > 1) wq *= A1.weight[ID[fa[1]] * nq1 + jq];
      dS += sqrt(fabs(det)) * wq;
    }
  return dS;
}

}  // namespace pc
