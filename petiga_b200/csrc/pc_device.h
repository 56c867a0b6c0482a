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
  const double* Wv;                // third state vector: U0 of the IE drivers / A of the I2 drivers, or NULL
  double shift2, t0;               // second shift (I2: shiftV) / second time (IE: t0)
  int face_axis, face_side;        // generic kernel in face mode: the visited face (IGAElementNextForm, petigaelem.c:427-447), else -1
  const double* bnd_value[3][2];   // IGABasis.bnd_value of every axis end ([p+1][5]) for the face mode
  double bnd_point[3][2];
  int maxdeg;
  const double* X;                 // geometry [ghost box][dim] or NULL
  const double* Wt;                // rational weights [ghost box] or NULL
  const double* fixtable;          // [ghost box][dof] or NULL
  const double* face_dS[3][2];     // mapped geometry + loads: BoundaryArea factor per face element [e_fa0 + ew_fa0*e_fa1] (local box) or NULL
  FixSide bc[3][2];
  int any_bc;
  int form, slot, block;           // block: 1 = BAIJ value layout, 0 = AIJ
  int mc0, mc1, vc0, vc1, per_qp, needs_x, needs_state;
  int c0, c1;                      // tabulated component range (union)
  double prm[kMaxPrm];
  double shift, t;
  int qc;                          // quadrature points per chunk
  int epb;                         // elements per block
  int noscatter;                   // profiling only (set_option "scatter" = 99): integrate but skip the global reductions
};

}  // namespace pc
