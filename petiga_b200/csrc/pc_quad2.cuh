// pc_quad2.cuh -- sum-factorised per-element quadrature kernel (general path, second generation).
//
// Same contract as quad_kernel (pc_quad.cuh): it replaces the reference's element loop body
// (src/petigaelem.c:375-410,693-1033,1166-1559; src/petigapoint.c:414-492) for a batch of elements per CTA.
// What changes is the algorithm of the O(nen^2 nqp) part.  The reference tabulates every shape function at every
// point (src/petiga3d.F90:32-233) and the callback loops over all (a,b) pairs per point.  Here the polynomial
// tensor-product structure psi_s(a,q) = prod_d B_d^{o_sd}(a_d,q_d) is kept to the end:
//
//   K_e[(a,i),(b,j)] = W_a W_b sum_{s,t} sum_q psi_s(a,q) D_q[ij][s][t] psi_t(b,q)
//
// where D_q = JW_q * A_q^T C_q A_q folds the form's coefficient tensor C_q (pc_forms.cuh), the inverse geometry
// map (K6/K7: d/dx_i = sum_d E[d][i] d/du_d) and the NURBS quotient rule (K4) into one small matrix over the
// parametric tensor components {N, d/du_d, d2/du_d2}.  The sum over q is contracted one axis at a time
//   stage A (q0): U1[g1][a0 b0][q1 q2];  stage B (q1) and C (q2) fused per thread (a0,b0,a1,b1) with a register
//   tile over (a2,b2)
// which needs ~7x fewer FP64 operations than the pair loop at p=3 in 3-D (SURVEY 8d anticipates this).
// Geometry, NURBS weights and state fields are evaluated at the points by the same axis-by-axis contraction.
#pragma once
#include "pc_device.h"

namespace pc {

constexpr int kMaxT = 7;          // tensor components: N, 3 first derivatives, 3 pure second derivatives
constexpr int kMaxPairs = 49;
constexpr int kMaxEval = 40;

struct SFLists {
  int NT;
  int tN, tG[3], tL[3];           // tensor index of N / d_d / d_dd, or -1
  int torder[kMaxT][3];
  int npairs; unsigned char pair_s[kMaxPairs], pair_t[kMaxPairs], pair_g1[kMaxPairs], pair_oo0[kMaxPairs];
  int ng1; unsigned char g1_oo1[kMaxPairs], g1_g2[kMaxPairs], g1_first[kMaxPairs + 1];   // pairs of g1 are [g1_first[g1], g1_first[g1+1])
  int ng2; unsigned char g2_oo2[9]; unsigned char g2_first[10];   // g1 groups of g2 are [g2_first[g2], g2_first[g2+1])
  int nev; unsigned char ev_field[kMaxEval], ev_t[kMaxEval];   // evaluation combos (field, tensor comp)
  int ev_index[16][kMaxT];        // (field, tensor comp) -> combo index or -1
  int nfields;                    // WX_0..WX_{DIM-1}, W, WU_c, WV_c
  int f_x0, f_w, f_u0, f_v0;      // first field index of each kind or -1
  int ijmask;                     // bit (i*DOF+j): block may be nonzero
  int dense_pairs;                // 1: all NT^2 pairs active (mapped / rational)
};

struct SFParams {
  KParams k;
  SFLists l;
  const double* pp[3];            // per axis [nel][9][nq][n*n] products B^{os}(a,q) B^{ot}(b,q)
  double cconst[kMaxPairs * 9];   // identity geometry + constant form: D'[ij][pair] / JW
  int const_dp;                   // 1: D'[pair][q] = JW_q * cconst[ij][pair]
};

}  // namespace pc
#include <vector>
namespace pc {
// host side (pc_quad2.cu), shared with the third-generation kernel (pc_quad3.cu)
template <int DIM, int DOF>
void host_matrix_pattern(int form, int slot, const double* prm, const FormInfo& fi, std::vector<char>& pat, int& ijmask, std::vector<double>& Cout);
int build_sf_lists(const KParams& kp, const FormInfo& fi, bool mapped, bool rational, bool state, bool transient,
                   const std::vector<char>& cpat, int ijmask, SFLists& l);

// plan-level table: PP[e][os*3+ot][q][a*n+b]
static __global__ void sf_pp_kernel(DevAxis ax, double* __restrict__ out) {
  const int n = ax.nen, nq = ax.nqp, per = 9 * nq * n * n;
  const size_t total = (size_t)ax.nel * per;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int e = (int)(t / per), r = (int)(t - (size_t)e * per), oo = r / (nq * n * n), r2 = r - oo * nq * n * n, q = r2 / (n * n), ab = r2 - q * n * n, a = ab / n, b = ab - a * n;
    const double* v = ax.value + ((size_t)(e * nq + q) * n) * 5;
    out[t] = v[a * 5 + oo / 3] * v[b * 5 + oo % 3];
  }
}

__host__ __device__ constexpr int sf_n(int dim, int p, int d) { return d < dim ? p + 1 : 1; }

// shared-memory carve-up of one element slot (offsets in doubles)
struct SFSmem {
  int b1d, pp0, pp1, p2, dp, fp, u1, ev, s1, s2, aq, cq, fq, fld, fe, r1, r2, geo, fixval, flux, ufix, rbase, ints, total;
  __host__ __device__ SFSmem(int n0, int n1, int n2, int nq0, int nq1, int nq2, int dim, int dof, const SFLists& l, int NA, int NV, int per_qp, int NC) {
    const int nqp = nq0 * nq1 * nq2, nen = n0 * n1 * n2;
    int o = 0;
    b1d = o; o += (3 * (nq0 * n0 + nq1 * n1 + nq2 * n2)); o += (o & 1);                 // B_d[o][q][a], o = 0..2
    pp0 = o; o += (9 * nq0 * n0 * n0); o += (o & 1);                                     // PP0[os*3+ot][q0][a0][b0]
    pp1 = o; o += (9 * nq1 * n1 * n1); o += (o & 1);
    p2 = o; o += (9 * nq2 * n2 * n2); o += (o & 1);
    dp = o; o += ((l.npairs > 0 ? l.npairs : 1) * nqp); o += (o & 1);                    // D'[pair][q] of the current (i,j) block
    fp = o; o += (dof * l.NT * nqp); o += (o & 1);                                       // f'[i][s][q]
    u1 = o; { int a = l.ng1 * nq2 * nq1 * n0 * n0, b = dof * l.NT * nq2 * nq1 * n0; o += ((a > b ? a : b) + 1); o += (o & 1); }   // U1[g1][q2][q1][a0 b0] / R1
    ev = o; o += ((l.nev > 0 ? l.nev : 1) * nqp); o += (o & 1);                          // evaluated polynomial fields
    s1 = o; { int a = l.nev * nq0 * n1 * n2, b = dof * l.NT * n0 * nq1 * nq2; o += ((a > b ? a : b) + 1); o += (o & 1); }
    s2 = o; { int a = l.nev * nq0 * nq1 * n2, b = dof * l.NT * n0 * n1 * nq2, c = (per_qp ? nqp : 1) * dof * dof * NA * NA; a = a > b ? a : b; o += ((a > c ? a : c) + 1); o += (o & 1); }
    aq = o; o += (nqp * NC * l.NT); o += (o & 1);                                        // A_q[al][s]
    cq = o; o += ((per_qp ? nqp : 1) * dof * dof * (NA > 0 ? NA * NA : 1)); o += (o & 1);
    fq = o; o += ((per_qp ? nqp : 1) * dof * (NV > 0 ? NV : 1)); o += (o & 1);
    fld = o; o += ((l.nfields > 0 ? l.nfields : 1) * nen); o += (o & 1);                 // nodal fields (already multiplied by W_a)
    fe = o; o += (nen * dof); o += (o & 1);
    r1 = o; o += (3 * nqp + nqp); o += (o & 1);                                          // per point: x[3], JW
    r2 = o; o += (nen); o += (o & 1);                                                    // W_a
    geo = o; o += (nqp * 13); o += (o & 1);                                             // per point: E[3][3], 1/w, grad w
    fixval = o; o += (nen * dof); o += (o & 1);
    flux = o; o += (nen * dof); o += (o & 1);
    ufix = o; o += (nen * dof); o += (o & 1);
    rbase = o; o += nen;                                  // int64 value-array offset of every local node's row (prefetched in the header)
    ints = o; o += ((nen + nen * dof + 3 * 25 + 3 * 5 + 8) / 2 + 1); o += (o & 1);
    total = o + (o & 1);
  }
};

template <int DIM, int P, int DOF>
struct SFCfg {
  static constexpr int n0 = sf_n(DIM, P, 0), n1 = sf_n(DIM, P, 1), n2 = sf_n(DIM, P, 2);
  static constexpr int NEN = n0 * n1 * n2;
  static constexpr int G = n0 * n0 * n1 * n1;     // threads per element: (b0, b1, a0, a1), b0 fastest
  static constexpr int THREADS = (G > 256) ? ((G + 31) / 32 * 32) : 256;
};

// NQ > 0: every used axis has exactly NQ quadrature points (the default rule NQ = p+1), loops unroll fully; NQ = 0: runtime
// MINB: resident CTAs per SM the register allocation is sized for.  MINB = 4 (<= 64 registers; the launcher picks it when
// four of the element's shared-memory slots fit an SM) hides more of the ~10 barriers per element: 69.5 -> 64.5 ms at cfg 2;
// with a larger slot (mapped geometry: 3 CTAs fit) the 77-register MINB = 1 build is the faster one.
template <int DIM, int P, int DOF, int NQ, int MINB>
__global__ void __launch_bounds__(SFCfg<DIM, P, DOF>::THREADS, MINB) quad_sf_kernel(const __grid_constant__ SFParams sp) {
  using Cfg = SFCfg<DIM, P, DOF>;
  constexpr int n0 = Cfg::n0, n1 = Cfg::n1, n2 = Cfg::n2, NEN = Cfg::NEN, G = Cfg::G, NEN1 = P + 1;
  const KParams& prm = sp.k;
  const SFLists& ls = sp.l;
  extern __shared__ double smem_all[];
  const int nq0 = (NQ > 0) ? NQ : prm.ax[0].nqp, nq1 = (NQ > 0) ? (DIM > 1 ? NQ : 1) : prm.ax[1].nqp,
            nq2 = (NQ > 0) ? (DIM > 2 ? NQ : 1) : prm.ax[2].nqp, nqp = nq0 * nq1 * nq2;
  const int NA = prm.mc1 - prm.mc0, NV = prm.vc1 - prm.vc0, NT = ls.NT;
  const SFSmem lay(n0, n1, n2, nq0, nq1, nq2, DIM, DOF, ls, NA, NV, prm.per_qp, prm.c1 - prm.c0);
  const int grp = threadIdx.x / G, lt = threadIdx.x - grp * G;
  const bool ingrp = grp < prm.epb;
  const int elem = blockIdx.x * prm.epb + grp;
  const bool valid = ingrp && elem < prm.nelem;
  double* sm = smem_all + (size_t)(ingrp ? grp : 0) * lay.total;
  double *B1d = sm + lay.b1d, *PP0 = sm + lay.pp0, *PP1 = sm + lay.pp1, *P2 = sm + lay.p2, *Dp = sm + lay.dp, *Fp = sm + lay.fp;
  double *U1 = sm + lay.u1, *Ev = sm + lay.ev, *S1 = sm + lay.s1, *S2 = sm + lay.s2, *Aq = sm + lay.aq, *Cq = sm + lay.cq, *Fq = sm + lay.fq;
  double *Fld = sm + lay.fld, *Fe = sm + lay.fe, *Xq = sm + lay.r1, *JW = sm + lay.r1 + 3 * nqp, *We = sm + lay.r2;
  double *FixVal = sm + lay.fixval, *Flux = sm + lay.flux, *UFix = sm + lay.ufix, *Geo = sm + lay.geo;
  int* lrow = reinterpret_cast<int*>(sm + lay.ints);
  int64_t* rbase = reinterpret_cast<int64_t*>(sm + lay.rbase);
  int* fixflag = lrow + NEN;
  uint32_t* segs = reinterpret_cast<uint32_t*>(fixflag + NEN * DOF);
  int* Wd = reinterpret_cast<int*>(segs + 3 * NEN1 * NEN1);
  const int nB[3] = {n0, n1, n2}, nQ[3] = {nq0, nq1, nq2};
  const int bOff[3] = {0, 3 * nq0 * n0, 3 * (nq0 * n0 + nq1 * n1)};
  // B_d[o][q][a]
#define BD(d, o, q, a) B1d[bOff[d] + ((o) * nQ[d] + (q)) * nB[d] + (a)]

  const bool mapped = prm.X != nullptr, rational = prm.Wt != nullptr;
  const bool want_mat = NA > 0, want_vec = (prm.slot != PETIGA_SLOT_MATRIX && prm.slot != PETIGA_SLOT_JACOBIAN && prm.slot != PETIGA_SLOT_IJACOBIAN);
  const bool state = prm.needs_state && prm.U != nullptr;
  const bool transient = (prm.slot == PETIGA_SLOT_IFUNCTION || prm.slot == PETIGA_SLOT_IJACOBIAN);

  int ID[3] = {0, 0, 0};
  if (valid) {
    int idx = elem;
#pragma unroll
    for (int d = 0; d < 3; d++) { int c = idx % prm.ax[d].ew; idx /= prm.ax[d].ew; ID[d] = c + prm.ax[d].es; }
  }

  // ---------------- header: closure, gathers, fix lists, position tables, 1-D tables ----------------
  if (valid) {
    for (int a = lt; a < NEN; a += G) {
      const int ia = a % n0, ja = (a / n0) % n1, ka = a / (n0 * n1);
      const int g0 = prm.ax[0].offset[ID[0]] + ia - prm.ax[0].gs;
      const int g1 = (DIM > 1) ? prm.ax[1].offset[ID[1]] + ja - prm.ax[1].gs : 0;
      const int g2 = (DIM > 2) ? prm.ax[2].offset[ID[2]] + ka - prm.ax[2].gs : 0;
      const int gidx = g0 + prm.ax[0].gw * (g1 + prm.ax[1].gw * g2);
      const int lr = prm.localrow[gidx];
      lrow[a] = lr;
      rbase[a] = want_mat ? prm.rowbase[lr] : 0;   // its L2 latency is paid here, under the header's other loads, not in the scatter
      const double wa = rational ? prm.Wt[gidx] : 1.0;
      We[a] = wa;
      if (mapped) {
#pragma unroll
        for (int i = 0; i < DIM; i++) Fld[(ls.f_x0 + i) * NEN + a] = wa * prm.X[(size_t)gidx * DIM + i];
      }
      if (ls.f_w >= 0) Fld[ls.f_w * NEN + a] = wa;
      int onfix[DOF];
      double vfix[DOF], vflux[DOF];
#pragma unroll
      for (int c = 0; c < DOF; c++) { onfix[c] = 0; vfix[c] = 0.0; vflux[c] = 0.0; }
      if (prm.any_bc) {   // BuildFix/AddFixa/AddFlux (petigaelem.c:1166-1283)
        const int ai[3] = {ia, ja, ka};
#pragma unroll
        for (int d = 0; d < DIM; d++) {
          if (prm.ax[d].periodic) continue;
          for (int s = 0; s < 2; s++) {
            const FixSide& fs = prm.bc[d][s];
            if (!(fs.vcount || fs.lcount)) continue;
            if (ID[d] != (s ? prm.ax[d].nel - 1 : 0)) continue;
            if (ai[d] != (s ? NEN1 - 1 : 0)) continue;
            for (int k = 0; k < fs.vcount; k++) {
              const int c = fs.vfield[k];
#pragma unroll
              for (int cc = 0; cc < DOF; cc++)
                if (cc == c) { onfix[cc] = 1; vfix[cc] = prm.fixtable ? prm.fixtable[(size_t)gidx * DOF + cc] : fs.vvalue[k]; }
            }
            if (fs.lcount) {
              double A = 1.0;
              if (DIM > 1) {
                for (int e = 0; e < DIM; e++) if (e != d) A *= prm.ax[e].detJac[ID[e]] / (double)NEN1;
                if (prm.face_dS[d][s]) {   // mapped geometry: surface Jacobian integrated over the face (face_area_kernel, pc_api.cu)
                  const int f0 = (d == 0) ? 1 : 0, f1 = (d == 2) ? 1 : 2;
                  const int fidx = (ID[f0] - prm.ax[f0].es) + ((DIM > 2) ? prm.ax[f0].ew * (ID[f1] - prm.ax[f1].es) : 0);
                  A *= prm.face_dS[d][s][fidx];
                } else A *= (DIM == 2) ? 2.0 : 4.0;
              }
              for (int k = 0; k < fs.lcount; k++)
#pragma unroll
                for (int cc = 0; cc < DOF; cc++) if (cc == fs.lfield[k]) vflux[cc] += fs.lvalue[k] * A;
            }
          }
        }
      }
#pragma unroll
      for (int c = 0; c < DOF; c++) {
        const int idx = a * DOF + c;
        double u = 0.0, v = 0.0;
        if (state) { u = prm.U[(size_t)lr * DOF + c]; if (transient && prm.V) v = prm.V[(size_t)lr * DOF + c]; }
        fixflag[idx] = onfix[c]; FixVal[idx] = vfix[c]; Flux[idx] = vflux[c]; UFix[idx] = u;
        if (onfix[c]) { u = vfix[c]; v = 0.0; }
        if (ls.f_u0 >= 0) Fld[(ls.f_u0 + c) * NEN + a] = wa * u;
        if (ls.f_v0 >= 0) Fld[(ls.f_v0 + c) * NEN + a] = wa * v;
        Fe[idx] = 0.0;
      }
    }
    for (int t = lt; t < 3 * NEN1 * NEN1; t += G) {
      const int d = t / (NEN1 * NEN1), r = t - d * NEN1 * NEN1, ia = r / NEN1, ib = r - ia * NEN1;
      uint32_t s = 0x00000100u;
      if (d < DIM) {
        const int g = prm.ax[d].offset[ID[d]] + ia - prm.ax[d].gs;
        s = prm.ax[d].seg[g * kMaxW + ib - ia + prm.ax[d].lo[g]];
        if (ib == 0) Wd[d * NEN1 + ia] = prm.ax[d].W[g];
      } else if (ib == 0) Wd[d * NEN1 + ia] = 1;
      segs[t] = s;
    }
    // 1-D tables B_d[o][q][a] of this element (IGABasis.value slice, include/petiga.h:122-141)
    for (int d = 0; d < 3; d++)
      for (int t = lt; t < 3 * nQ[d] * nB[d]; t += G) {
        const int o = t / (nQ[d] * nB[d]), r = t - o * nQ[d] * nB[d], q = r / nB[d], a = r - q * nB[d];
        BD(d, o, q, a) = prm.ax[d].value[((size_t)(ID[d] * nQ[d] + q) * nB[d] + a) * 5 + o];
      }
  }
  __syncthreads();
  if (valid) {  // pair products PP_d[os*3+ot][q][a][b] = B_d^{os}(a,q) B_d^{ot}(b,q): per-axis tables built once per plan
    {
      const double* g0p = sp.pp[0] + (size_t)ID[0] * 9 * nq0 * n0 * n0;
      for (int t = lt; t < 9 * nq0 * n0 * n0; t += G) PP0[t] = g0p[t];
      const double* g1p = sp.pp[1] + (size_t)ID[1] * 9 * nq1 * n1 * n1;
      for (int t = lt; t < 9 * nq1 * n1 * n1; t += G) PP1[t] = g1p[t];
      const double* g2p = sp.pp[2] + (size_t)ID[2] * 9 * nq2 * n2 * n2;
      for (int t = lt; t < 9 * nq2 * n2 * n2; t += G) P2[t] = g2p[t];
    }
    // ---- field evaluation at the points, axis by axis: Ev[c][q] = sum_a psi_t(a,q) F_f[a]  (K5, K11/K12) ----
    for (int t = lt; t < ls.nev * nq0 * n1 * n2; t += G) {
      const int c = t / (nq0 * n1 * n2), r = t - c * nq0 * n1 * n2, q0 = r / (n1 * n2), a12 = r - q0 * n1 * n2;
      const double* F = Fld + ls.ev_field[c] * NEN + a12 * n0;
      const int o = ls.torder[ls.ev_t[c]][0];
      double s = 0.0;
#pragma unroll
      for (int a0 = 0; a0 < n0; a0++) s += BD(0, o, q0, a0) * F[a0];
      S1[t] = s;                                     // [c][q0][a2][a1]  (a12 = a1 + n1*a2)
    }
  }
  __syncthreads();
  if (valid)
    for (int t = lt; t < ls.nev * nq0 * nq1 * n2; t += G) {
      const int c = t / (nq0 * nq1 * n2), r = t - c * nq0 * nq1 * n2, q0 = r / (nq1 * n2), r2 = r - q0 * nq1 * n2, q1 = r2 / n2, a2 = r2 - q1 * n2;
      const double* s1 = S1 + (c * nq0 + q0) * n1 * n2 + a2 * n1;
      const int o = ls.torder[ls.ev_t[c]][1];
      double s = 0.0;
#pragma unroll
      for (int a1 = 0; a1 < n1; a1++) s += BD(1, o, q1, a1) * s1[a1];
      S2[t] = s;                                     // [c][q0][q1][a2]
    }
  __syncthreads();
  if (valid)
    for (int t = lt; t < ls.nev * nqp; t += G) {
      const int c = t / nqp, q = t - c * nqp, q0 = q % nq0, q1 = (q / nq0) % nq1, q2 = q / (nq0 * nq1);
      const double* s2 = S2 + ((c * nq0 + q0) * nq1 + q1) * n2;
      const int o = ls.torder[ls.ev_t[c]][2];
      double s = 0.0;
#pragma unroll
      for (int a2 = 0; a2 < n2; a2++) s += BD(2, o, q2, a2) * s2[a2];
      Ev[t] = s;                                     // [c][q], q = q0 + nq0*(q1 + nq1*q2) as the reference orders points
    }
  __syncthreads();

  // ---------------- per point: geometry, weights, state, coefficient tensors in parametric components ----------------
  // A[al][s]: physical component al of the (rational, mapped) shape function of node a = W_a * sum_s A[al][s] psi_s(a,q)
  const int NC = prm.c1 - prm.c0;
  // (1) one thread per point: weights, NURBS denominator, geometry map and its inverse (K4-K6)
  if (valid)
    for (int q = lt; q < nqp; q += G) {
      const int qi[3] = {q % nq0, (q / nq0) % nq1, q / (nq0 * nq1)};
      double w = 1.0, J = 1.0, x[3] = {0, 0, 0};
#pragma unroll
      for (int d = 0; d < DIM; d++) {
        w *= prm.ax[d].weight[ID[d] * nQ[d] + qi[d]];
        J *= prm.ax[d].detJac[ID[d]];
        x[d] = prm.ax[d].point[ID[d] * nQ[d] + qi[d]];
      }
      auto ev = [&](int field, int tc) -> double {
        if (field < 0 || tc < 0) return 0.0;
        const int c = ls.ev_index[field][tc];
        return c >= 0 ? Ev[c * nqp + q] : 0.0;
      };
      double w0 = 1.0, wg[3] = {0, 0, 0};
      if (rational) {
        w0 = ev(ls.f_w, ls.tN);
#pragma unroll
        for (int d = 0; d < DIM; d++) wg[d] = ev(ls.f_w, ls.tG[d]);
      }
      const double iw = 1.0 / w0;
      double E[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};   // E[d][i] = du_d/dx_i
      if (mapped) {
        double X0[3] = {0, 0, 0}, X1[3][3];                  // X1[i][d] = dX_i/du_d
#pragma unroll
        for (int i = 0; i < DIM; i++) {
          X0[i] = ev(ls.f_x0 + i, ls.tN) * iw;
#pragma unroll
          for (int d = 0; d < DIM; d++) X1[i][d] = (ev(ls.f_x0 + i, ls.tG[d]) - X0[i] * wg[d]) * iw;   // quotient rule (K4) on W*X
          x[i] = X0[i];
        }
        double det;
        if (DIM == 1) { det = X1[0][0]; E[0][0] = 1.0 / det; }
        else if (DIM == 2) {
          det = X1[0][0] * X1[1][1] - X1[0][1] * X1[1][0];
          E[0][0] = X1[1][1] / det; E[0][1] = -X1[0][1] / det; E[1][0] = -X1[1][0] / det; E[1][1] = X1[0][0] / det;
        } else {
          const double a00 = X1[0][0], a01 = X1[0][1], a02 = X1[0][2], a10 = X1[1][0], a11 = X1[1][1], a12 = X1[1][2], a20 = X1[2][0], a21 = X1[2][1], a22 = X1[2][2];
          det = a00 * (a11 * a22 - a12 * a21) - a01 * (a10 * a22 - a12 * a20) + a02 * (a10 * a21 - a11 * a20);
          E[0][0] = (a11 * a22 - a12 * a21) / det; E[0][1] = -(a01 * a22 - a02 * a21) / det; E[0][2] = (a01 * a12 - a02 * a11) / det;
          E[1][0] = -(a10 * a22 - a12 * a20) / det; E[1][1] = (a00 * a22 - a02 * a20) / det; E[1][2] = -(a00 * a12 - a02 * a10) / det;
          E[2][0] = (a10 * a21 - a11 * a20) / det; E[2][1] = -(a00 * a21 - a01 * a20) / det; E[2][2] = (a00 * a11 - a01 * a10) / det;
        }
        J *= det;                                            // detJac *= detX (petigaelem.c:1024-1029)
      }
      JW[q] = J * w;
#pragma unroll
      for (int d = 0; d < 3; d++) Xq[d * nqp + q] = x[d];
#pragma unroll
      for (int d = 0; d < 3; d++)
#pragma unroll
        for (int i = 0; i < 3; i++) Geo[(d * 3 + i) * nqp + q] = E[d][i];
      Geo[9 * nqp + q] = iw;
#pragma unroll
      for (int d = 0; d < 3; d++) Geo[(10 + d) * nqp + q] = wg[d];
    }
  __syncthreads();
  // (2) all threads: component transformation matrix A[al][s][q]
  if (valid)
    for (int t = lt; t < NC * NT * nqp; t += G) {
      const int q = t % nqp, r = t / nqp, s = r % NT, al = r / NT, c = al + prm.c0;
      const double iw = Geo[9 * nqp + q];
      double v = 0.0;
      if (c == 0) v = (s == ls.tN) ? iw : 0.0;
      else if (c <= DIM) {
        const int i = c - 1;
#pragma unroll
        for (int d = 0; d < DIM; d++) if (s == ls.tG[d]) v = Geo[(d * 3 + i) * nqp + q] * iw;
        if (rational && s == ls.tN) {
          double sN = 0.0;
#pragma unroll
          for (int d = 0; d < DIM; d++) sN -= Geo[(d * 3 + i) * nqp + q] * Geo[(10 + d) * nqp + q];
          v = sN * iw * iw;
        }
      } else {
#pragma unroll
        for (int d = 0; d < DIM; d++) if (s == ls.tL[d]) v = 1.0;   // Laplacian on the identity map only (host checks)
      }
      Aq[t] = v;
    }
  __syncthreads();
  // (3) one thread per point (only the points that need it): state (K12) and the form's coefficient tensors
  if (valid)
    for (int q = lt; q < nqp; q += G) {
      if (!(prm.per_qp || q == 0)) continue;
      QPoint qp;
      qp.atboundary = 0;
#pragma unroll
      for (int d = 0; d < 3; d++) qp.x[d] = Xq[d * nqp + q];
      if (state) {
        const double* A = Aq + q;
#pragma unroll
        for (int i = 0; i < DOF; i++) {
          double ph[kMaxComp];
          for (int al = 0; al < NC; al++) {
            double sacc = 0.0;
            for (int t = 0; t < NT; t++) { const int c = ls.ev_index[ls.f_u0 + i][t]; if (c >= 0) sacc += A[(al * NT + t) * nqp] * Ev[c * nqp + q]; }
            ph[al] = sacc;
          }
          qp.u[i] = (prm.c0 == 0) ? ph[0] : 0.0;
#pragma unroll
          for (int d = 0; d < DIM; d++) { const int al = 1 + d - prm.c0; qp.gu[i][d] = (al >= 0 && al < NC) ? ph[al] : 0.0; }
          { const int al = DIM + 1 - prm.c0; qp.d2u[i] = (al >= 0 && al < NC) ? ph[al] : 0.0; }
          qp.v[i] = 0.0;
          if (ls.f_v0 >= 0 && prm.c0 == 0) { const int c = ls.ev_index[ls.f_v0 + i][ls.tN]; if (c >= 0) qp.v[i] = A[ls.tN * nqp] * Ev[c * nqp + q]; }
        }
      }
      double* fv = Fq + (size_t)(prm.per_qp ? q : 0) * DOF * (NV > 0 ? NV : 1);
      // the form writes a dense [DOF][DOF][NA][NA] block; stage it in this point's slice of S2 (free at this time), then
      // store it with the point index fastest so that the D' pass reads it without bank conflicts
      double* C = S2 + (size_t)(prm.per_qp ? q : 0) * DOF * DOF * NA * NA;
      for (int k = 0; k < DOF * DOF * NA * NA; k++) C[k] = 0.0;
      for (int k = 0; k < DOF * NV; k++) fv[k] = 0.0;
      form_coefficients<DIM, DOF>(prm.form, prm.slot, prm.prm, prm.shift, prm.t, qp, NA, NV, NA ? C : nullptr, NV ? fv : nullptr);
      const int cstride = prm.per_qp ? nqp : 1;
      for (int k = 0; k < DOF * DOF * NA * NA; k++) Cq[(size_t)k * cstride + (prm.per_qp ? q : 0)] = C[k];
    }
  __syncthreads();
  // (4) all threads: vector coefficients in tensor components f'[i][s][q] = JW_q * sum_al A[vc0-c0+al][s][q] f_q[i][al]
  if (want_vec && NV > 0) {   // uniform over the CTA
    if (valid)
      for (int t = lt; t < nqp * DOF * NT; t += G) {
        const int q = t % nqp, r = t / nqp, s = r % NT, i = r / NT;
        const double* fsrc = Fq + (size_t)(prm.per_qp ? q : 0) * DOF * NV + i * NV;
        double acc = 0.0;
        for (int al = 0; al < NV; al++) acc += Aq[((prm.vc0 - prm.c0 + al) * NT + s) * nqp + q] * fsrc[al];
        Fp[(i * NT + s) * nqp + q] = acc * JW[q];
      }
    __syncthreads();
  }

  // ---------------- element vector by axis-by-axis contraction (transpose of the evaluation) ----------------
  if (want_vec && NV > 0) {
    double* R1 = U1;    // [i*NT+s][q2][q1][a0]
    double* R2 = S2;    // [i*NT+s][q2][a1][a0]
    if (valid)
      for (int t = lt; t < DOF * NT * nq2 * nq1 * n0; t += G) {
        const int c = t / (nq2 * nq1 * n0), r = t - c * nq2 * nq1 * n0, q12 = r / n0, a0 = r - q12 * n0;
        const int o = ls.torder[c % NT][0];
        const double* f = Fp + c * nqp + q12 * nq0;
        double s = 0.0;
        for (int q0 = 0; q0 < nq0; q0++) s += BD(0, o, q0, a0) * f[q0];
        R1[t] = s;
      }
    __syncthreads();
    if (valid)
      for (int t = lt; t < DOF * NT * nq2 * n1 * n0; t += G) {
        const int c = t / (nq2 * n1 * n0), r = t - c * nq2 * n1 * n0, q2 = r / (n1 * n0), a01 = r - q2 * n1 * n0, a1 = a01 / n0, a0 = a01 - a1 * n0;
        const int o = ls.torder[c % NT][1];
        double s = 0.0;
        for (int q1 = 0; q1 < nq1; q1++) s += BD(1, o, q1, a1) * R1[((c * nq2 + q2) * nq1 + q1) * n0 + a0];
        R2[t] = s;
      }
    __syncthreads();
    if (valid)
      for (int t = lt; t < NEN * DOF; t += G) {
        const int a = t / DOF, i = t - a * DOF, a01 = a % (n0 * n1), a2 = a / (n0 * n1);
        double s = 0.0;
        for (int sc = 0; sc < NT; sc++) {
          const int o = ls.torder[sc][2], c = i * NT + sc;
          for (int q2 = 0; q2 < nq2; q2++) s += BD(2, o, q2, a2) * R2[(c * nq2 + q2) * n1 * n0 + a01];
        }
        Fe[t] = s * We[a];
      }
    __syncthreads();
  }

  // ---------------- element matrix: one pass per (i,j) block ----------------
  const int a0 = (lt / (n0 * n1)) % n0, a1 = lt / (n0 * n1 * n0), b0 = lt % n0, b1 = (lt / n0) % n1;
  const int ab0 = a0 * n0 + b0, ab1 = a1 * n1 + b1;
  const bool fix_mat = (prm.slot == PETIGA_SLOT_SYSTEM || prm.slot == PETIGA_SLOT_JACOBIAN || prm.slot == PETIGA_SLOT_IJACOBIAN);
  bool elem_fix = false;   // does this element hold Dirichlet dofs? (boundary element on a face with values, petigaelem.c:1263-1283)
  if (prm.any_bc)
#pragma unroll
    for (int d = 0; d < DIM; d++)
      if (!prm.ax[d].periodic)
        elem_fix = elem_fix || (ID[d] == 0 && prm.bc[d][0].vcount > 0) || (ID[d] == prm.ax[d].nel - 1 && prm.bc[d][1].vcount > 0);
  if (want_mat) {
    for (int ij = 0; ij < DOF * DOF; ij++) {
      if (!((ls.ijmask >> ij) & 1)) continue;     // uniform over the grid
      const int bi = ij / DOF, bj = ij - bi * DOF;
      // D'[pair][q] = JW_q * sum_{al,be} A[mc0-c0+al][s] C_q[i][j][al][be] A[mc0-c0+be][t]
      if (valid && sp.const_dp) {
        for (int t = lt; t < ls.npairs * nqp; t += G) { const int pr = t / nqp, q = t - pr * nqp; Dp[t] = sp.cconst[ij * kMaxPairs + pr] * JW[q]; }
      } else if (valid)
        for (int t = lt; t < ls.npairs * nqp; t += G) {
          const int pr = t / nqp, q = t - pr * nqp, s = ls.pair_s[pr], tt = ls.pair_t[pr];
          const int cstride = prm.per_qp ? nqp : 1;
          const double* C = Cq + (size_t)ij * NA * NA * cstride + (prm.per_qp ? q : 0);
          const double* Am = Aq + (size_t)(prm.mc0 - prm.c0) * NT * nqp + q;
          double acc = 0.0;
          for (int al = 0; al < NA; al++) {
            const double as = Am[(al * NT + s) * nqp];
            if (as == 0.0) continue;
            double inner = 0.0;
            for (int be = 0; be < NA; be++) inner += C[(al * NA + be) * cstride] * Am[(be * NT + tt) * nqp];
            acc += as * inner;
          }
          Dp[t] = acc * JW[q];
        }
      __syncthreads();
      // stage A: U1[g1][q2][q1][a0 b0] = sum_{pairs in g1} sum_q0 PP0[os0,ot0][q0][a0 b0] * D'[pair][q0,q1,q2]
      if (valid)
        for (int t = lt; t < ls.ng1 * nq2 * nq1 * n0 * n0; t += G) {
          const int g1 = t / (nq2 * nq1 * n0 * n0), r = t - g1 * nq2 * nq1 * n0 * n0, q12 = r / (n0 * n0), ab = r - q12 * n0 * n0;
          double acc = 0.0;
          for (int pr = ls.g1_first[g1]; pr < ls.g1_first[g1 + 1]; pr++) {
            const int oo = ls.pair_oo0[pr];
            const double* pp = PP0 + (size_t)oo * nq0 * n0 * n0 + ab;
            const double* dq = Dp + (size_t)pr * nqp + q12 * nq0;
            for (int q0 = 0; q0 < nq0; q0++) acc += pp[q0 * n0 * n0] * dq[q0];
          }
          U1[t] = acc;
        }
      __syncthreads();
      // stages B + C per thread (a0,b0,a1,b1): register tile over (a2,b2)
      double acc[n2][n2];
#pragma unroll
      for (int x = 0; x < n2; x++)
#pragma unroll
        for (int y = 0; y < n2; y++) acc[x][y] = 0.0;
      if (valid) {
        for (int q2 = 0; q2 < nq2; q2++)
          for (int g2 = 0; g2 < ls.ng2; g2++) {
            double u2 = 0.0;
            for (int g1 = ls.g2_first[g2]; g1 < ls.g2_first[g2 + 1]; g1++) {
              const double* pp = PP1 + (size_t)ls.g1_oo1[g1] * nq1 * n1 * n1 + ab1;
              const double* u1 = U1 + ((size_t)(g1 * nq2 + q2) * nq1) * n0 * n0 + ab0;
              for (int q1 = 0; q1 < nq1; q1++) u2 += pp[q1 * n1 * n1] * u1[q1 * n0 * n0];
            }
            // acc[x][y] += B2^{os}(x,q2) * (B2^{ot}(y,q2) * u2): 2*n2 broadcast loads instead of n2*n2
            const int oo2 = ls.g2_oo2[g2];
            const double* bs = &BD(2, oo2 / 3, q2, 0);
            const double* bt = &BD(2, oo2 % 3, q2, 0);
            double ty[n2];
#pragma unroll
            for (int y = 0; y < n2; y++) ty[y] = bt[y] * u2;
#pragma unroll
            for (int x = 0; x < n2; x++) {
              const double bx = bs[x];
#pragma unroll
              for (int y = 0; y < n2; y++) acc[x][y] = fma(bx, ty[y], acc[x][y]);
            }
          }
      }
      // NURBS node weights, fix-up (petigaelem.c:1360-1389,1483-1501) and scatter of this block.  The closed-form position
      // pos = Bk*W1*W0 + Sk*(Bj*W0 + Sj*Bi) + (Lk*Sj + Lj)*Si + Li is affine in the (a2,b2)-dependent bytes (Bk,Sk,Lk) with
      // thread-constant coefficients, and only boundary elements can hold fixed dofs, so an entry costs one table load,
      // three multiply-adds and the reduction (the address arithmetic was ~1/3 of the kernel's instructions before).
      if (valid) {
        const uint32_t s0 = segs[a0 * NEN1 + b0], s1 = segs[NEN1 * NEN1 + a1 * NEN1 + b1];
        const int Bi = s0 & 255, Si = (s0 >> 8) & 255, Li = (s0 >> 16) & 255;
        const int Bj = s1 & 255, Sj = (s1 >> 8) & 255, Lj = (s1 >> 16) & 255;
        const int W0 = Wd[a0], W1 = Wd[NEN1 + a1];
        const int cB = W1 * W0, cS = Bj * W0 + Sj * Bi, cL = Sj * Si, c0 = Lj * Si + Li;
        const bool fixing = fix_mat && elem_fix;
#pragma unroll
        for (int x = 0; x < n2; x++) {
          const int a = a0 + n0 * (a1 + n1 * x);
          const int ra = a * DOF + bi;
          const int lr = lrow[a];
          int64_t base = rbase[a];
          double* dst = prm.values;
          if (lr >= prm.nown) { dst = prm.ghost_values; base -= prm.nnz_own; }
          if (DOF == 1) dst += base;
          else if (prm.block) dst += (size_t)base * DOF * DOF + bj * DOF + bi;
          else dst += (size_t)base * DOF * DOF + (size_t)bi * (cB * Wd[2 * NEN1 + x]) * DOF + bj;
          const bool fr = fixing && fixflag[ra];
          const double wa = rational ? We[a] : 1.0;
#pragma unroll
          for (int y = 0; y < n2; y++) {
            double v = acc[x][y];
            if (rational) v *= wa * We[b0 + n0 * (b1 + n1 * y)];
            if (fixing) {
              const int cb = (b0 + n0 * (b1 + n1 * y)) * DOF + bj;
              const bool fc = fixflag[cb];
              if (fr || fc) {
                if (prm.slot == PETIGA_SLOT_SYSTEM && fc && !fr) atomicAdd(&Fe[ra], -v * FixVal[cb]);
                v = (ra == cb) ? 1.0 : 0.0;
              }
            }
            if (v == 0.0) continue;
            const uint32_t s2 = segs[2 * NEN1 * NEN1 + x * NEN1 + y];
            const int pos = (int)(s2 & 255) * cB + (int)((s2 >> 8) & 255) * cS + (int)((s2 >> 16) & 255) * cL + c0;
            double* p = (DOF == 1) ? dst + pos : (prm.block ? dst + (size_t)pos * DOF * DOF : dst + (size_t)pos * DOF);
            if (!prm.noscatter) atomicAdd(p, v);
            else if (v == 1.2345e300) *p = v;   // keeps the value and the address live
          }
        }
      }
      __syncthreads();
    }
    // structurally present but skipped blocks of fixed rows still need their unit diagonal: handled because the
    // diagonal block (i,i) is always in ijmask for every built-in form
  }
  __syncthreads();
  if (valid && want_vec) {  // FixSystem / FixFunction vector part, then VecSetValuesLocal(ADD_VALUES)
    for (int t = lt; t < NEN * DOF; t += G) {
      const int a = t / DOF, i = t - a * DOF;
      double F = Fe[t];
      if (prm.slot == PETIGA_SLOT_SYSTEM) { F += Flux[t]; if (fixflag[t]) F = FixVal[t]; }
      else if (prm.slot == PETIGA_SLOT_FUNCTION || prm.slot == PETIGA_SLOT_IFUNCTION) { F -= Flux[t]; if (fixflag[t]) F = UFix[t] - FixVal[t]; }
      if (F != 0.0) atomicAdd(&prm.rhs[(size_t)lrow[a] * DOF + i], F);
    }
  }
#undef BD
}

}  // namespace pc
