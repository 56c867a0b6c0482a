// pc_kron.cu -- the separable ("Kronecker") assembly path.
//
// When the geometry map is the identity and the form's coefficient tensor is constant (Poisson, Laplace,
// mass/L2, elasticity: every linear BASELINE config), the sum over elements of the element matrices of the
// reference's quadrature loop (src/petigaksp.c:149-202) factorises over the axes:
//
//   A[(A,i),(B,j)] = sum_{al,be} C[i][j][al][be] * prod_d M_d^{(r_d,s_d)}[A_d][B_d],
//   M_d^{(r,s)}[a][b] = sum_{e in supp(a) & supp(b)} sum_q w_q J_e N_a^{(r)}(u_q) N_b^{(s)}(u_q)      (1-D, banded)
//
// with r_d = 1 iff component al is the derivative along axis d.  So every CSR value can be *written once*
// from three tiny 1-D tables: no per-element work, no zeroing pass, no atomics, no ghost-row exchange (each rank
// writes exactly its owned rows).  The kernel is bound by the 8 bytes it stores per nonzero (HBM roofline).
//
// The reference's per-element Dirichlet/Neumann fix-up (IGAElementFixSystem, src/petigaelem.c:1360-1389) is
// reproduced in aggregate: a boundary node is fixed in every element that contains it, so after ADD_VALUES
//   fixed row    -> zero, diagonal = number of elements containing the node, rhs = that count * value
//   fixed column -> zero, its unfixed entry times the value is subtracted from the row's rhs
//   loads        -> rhs += value * sum over the face elements of BoundaryArea (petigaelem.c:1118-1131)
#include <algorithm>
#include <cstring>

#include "pc_plan.h"

namespace pc {

namespace {

constexpr int kMaxTerms = 40;
constexpr int kMaxWW = kMaxW * kMaxW;

struct KronTerm { unsigned char ij, rs0, rs1, rs2; double c; };
struct KronVTerm { unsigned char i, r0, r1, r2; double c; };

struct KronParams {
  const double* M[3];     // [4][nnp][kMaxW]  1-D matrices, rs = 2*r+s major
  const double* mv[3];    // [2][nnp]         1-D load vectors  sum_e sum_q wJ N^{(r)}
  const int* nsup[3];     // [nnp]            elements containing basis i
  const double* lsum[3];  // [nnp]            sum over those elements of detJac/nen
  const int* first[3];    // [nnp]
  const int* Wg[3];       // [gw] widths by ghost coordinate
  const int* lo[3];       // [gw]
  const uint32_t* seg[3]; // [gw][kMaxW]
  int ls[3], lw[3], gs[3], nnp[3], periodic[3];
  const int64_t* rowbase;
  const int* localrow;    // ghost box -> local row (for the fix table)
  const double* fixtable; // [ghost box][dof] or NULL
  int gw[3];
  int dim, dof, block, slot, simple;
  int nterms, nvterms, rsmask0;
  KronTerm terms[kMaxTerms];
  KronVTerm vterms[8];
  FixSide bc[3][2];
  int any_bc;
  double* values;
  double* rhs;
};

__device__ __forceinline__ int wrapi(int i, int n) { return i < 0 ? n + i : (i >= n ? i % n : i); }

// 1-D tables: one thread per (axis-local) output; deterministic summation order (elements ascending)
__global__ void kron_1d_kernel(DevAxis ax, const int* __restrict__ first, double* __restrict__ M, double* __restrict__ mv,
                               int* __restrict__ nsup, double* __restrict__ lsum) {
  const int nnp = ax.nnp, nen = ax.nen, nq = ax.nqp;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = nnp * kMaxW * 4;
  if (t < total) {
    const int rs = t / (nnp * kMaxW), rem = t - rs * nnp * kMaxW, i = rem / kMaxW, c = rem - i * kMaxW;
    const int r = rs >> 1, s = rs & 1;
    const int col = wrapi(first[i] + c, nnp);
    double acc = 0.0;
    for (int e = 0; e < ax.nel; e++) {
      int a = -1, b = -1;
      for (int l = 0; l < nen; l++) {
        int node = wrapi(ax.offset[e] + l, nnp);
        if (node == i) a = l;
        if (node == col) b = l;
      }
      // the unwrapped column must be the one this element sees: first+c - i == b - a  (guards short periodic axes)
      if (a < 0 || b < 0) continue;
      const double J = ax.detJac[e];
      double se = 0.0;
      for (int q = 0; q < nq; q++) {
        const double* va = ax.value + ((size_t)(e * nq + q) * nen + a) * 5;
        const double* vb = ax.value + ((size_t)(e * nq + q) * nen + b) * 5;
        se += (ax.weight[e * nq + q] * J) * va[r] * vb[s];
      }
      acc += se;
    }
    M[t] = acc;
  }
  if (t < nnp * 2) {
    const int r = t / nnp, i = t - r * nnp;
    double acc = 0.0, ls = 0.0;
    int cnt = 0;
    for (int e = 0; e < ax.nel; e++) {
      int a = -1;
      for (int l = 0; l < nen; l++) if (wrapi(ax.offset[e] + l, nnp) == i) a = l;
      if (a < 0) continue;
      const double J = ax.detJac[e];
      double se = 0.0;
      for (int q = 0; q < nq; q++) se += (ax.weight[e * nq + q] * J) * ax.value[((size_t)(e * nq + q) * nen + a) * 5 + r];
      acc += se;
      cnt++;
      ls += J / (double)nen;
    }
    mv[t] = acc;
    if (r == 0) { nsup[i] = cnt; lsum[i] = ls; }
  }
}

// Which faces fix dof component c of a node with per-axis boundary codes (0 interior, 1 side 0, 2 side 1)?
// Returns true and the value of the last face in the reference's order (d ascending, side 0 then 1).
template <int DOF>
__device__ __forceinline__ void node_fix(const KronParams& kp, const int code[3], bool fixed[DOF], double val[DOF]) {
#pragma unroll
  for (int c = 0; c < DOF; c++) { fixed[c] = false; val[c] = 0.0; }
  for (int d = 0; d < kp.dim; d++) {
    if (!code[d]) continue;
    const FixSide& fs = kp.bc[d][code[d] - 1];
    for (int k = 0; k < fs.vcount; k++) { fixed[fs.vfield[k]] = true; val[fs.vfield[k]] = fs.vvalue[k]; }
  }
}

template <int DOF, bool SIMPLE>
__global__ void __launch_bounds__(256) kron_rows_kernel(const __grid_constant__ KronParams kp) {
  __shared__ double G[4][DOF * DOF][kMaxWW];     // G^{rs0}_{ij}[cjk]
  __shared__ int colcode[kMaxWW];                // boundary codes of the (j,k) part of the column node: cj | ck<<2
  const int Aj = kp.ls[1] + (int)(blockIdx.x % kp.lw[1]), Ak = kp.ls[2] + (int)(blockIdx.x / kp.lw[1]);
  const int gj = Aj - kp.gs[1], gk = Ak - kp.gs[2];
  const int Wj = kp.Wg[1][gj], Wk = kp.Wg[2][gk], Wjk = Wj * Wk;
  const int fj = kp.first[1][Aj], fk = kp.first[2][Ak];
  for (int t = threadIdx.x; t < 4 * DOF * DOF * kMaxWW; t += blockDim.x) (&G[0][0][0])[t] = 0.0;
  __syncthreads();
  for (int t = threadIdx.x; t < Wjk; t += blockDim.x) {
    const int cj = t % Wj, ck = t / Wj;
    for (int n = 0; n < kp.nterms; n++) {
      const KronTerm tm = kp.terms[n];
      const double mj = kp.M[1][((size_t)tm.rs1 * kp.nnp[1] + Aj) * kMaxW + cj];
      const double mk = kp.M[2][((size_t)tm.rs2 * kp.nnp[2] + Ak) * kMaxW + ck];
      G[tm.rs0][tm.ij][t] += tm.c * mj * mk;
    }
    int code = 0;
    if (!kp.periodic[1]) { int col = fj + cj; code |= (col == 0) ? 1 : ((col == kp.nnp[1] - 1) ? 2 : 0); }
    if (!kp.periodic[2]) { int col = fk + ck; code |= ((col == 0) ? 1 : ((col == kp.nnp[2] - 1) ? 2 : 0)) << 2; }
    colcode[t] = code;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const bool want_mat = kp.values != nullptr, want_vec = kp.rhs != nullptr;
  const bool fixing = kp.any_bc && kp.slot == PETIGA_SLOT_SYSTEM;
  const int rcodej = kp.periodic[1] ? 0 : ((Aj == 0) ? 1 : ((Aj == kp.nnp[1] - 1) ? 2 : 0));
  const int rcodek = kp.periodic[2] ? 0 : ((Ak == 0) ? 1 : ((Ak == kp.nnp[2] - 1) ? 2 : 0));
  for (int il = warp; il < kp.lw[0]; il += nwarps) {
    const int Ai = kp.ls[0] + il, gi = Ai - kp.gs[0];
    const int Wi = kp.Wg[0][gi], fi = kp.first[0][Ai], W = Wi * Wjk;
    const int lr = il + kp.lw[0] * ((Aj - kp.ls[1]) + kp.lw[1] * (Ak - kp.ls[2]));
    const int64_t base = kp.rowbase[lr];
    int rcode[3] = {kp.periodic[0] ? 0 : ((Ai == 0) ? 1 : ((Ai == kp.nnp[0] - 1) ? 2 : 0)), rcodej, rcodek};
    bool rfix[DOF];
    double rval[DOF];
#pragma unroll
    for (int c = 0; c < DOF; c++) { rfix[c] = false; rval[c] = 0.0; }
    if (fixing) {
      node_fix<DOF>(kp, rcode, rfix, rval);
      if (kp.fixtable) {
        const int gidx = gi + kp.gw[0] * (gj + kp.gw[1] * gk);
#pragma unroll
        for (int c = 0; c < DOF; c++) if (rfix[c]) rval[c] = kp.fixtable[(size_t)gidx * DOF + c];
      }
    }
    const double nelem = (double)(kp.nsup[0][Ai] * kp.nsup[1][Aj] * kp.nsup[2][Ak]);
    const int dci = Ai - fi, dcj = Aj - fj, dck = Ak - fk;     // column offsets of the diagonal entry
    double racc[DOF];
#pragma unroll
    for (int c = 0; c < DOF; c++) racc[c] = 0.0;
    const unsigned inv = (65536u + Wi - 1) / Wi;
    for (int e = lane; e < W; e += 32) {
      const int cjk = (int)((e * inv) >> 16), ci = e - cjk * Wi;
      double a[4];
#pragma unroll
      for (int rs = 0; rs < 4; rs++) a[rs] = (kp.rsmask0 >> rs) & 1 ? kp.M[0][((size_t)rs * kp.nnp[0] + Ai) * kMaxW + ci] : 0.0;
      int pos = e;
      if (!SIMPLE) {
        const int cj = cjk % Wj, ck = cjk / Wj;
        const uint32_t s0 = kp.seg[0][gi * kMaxW + ci], s1 = kp.seg[1][gj * kMaxW + cj], s2 = kp.seg[2][gk * kMaxW + ck];
        const int Bi = s0 & 255, Si = (s0 >> 8) & 255, Li = (s0 >> 16) & 255;
        const int Bj = s1 & 255, Sj = (s1 >> 8) & 255, Lj = (s1 >> 16) & 255;
        const int Bk = s2 & 255, Sk = (s2 >> 8) & 255, Lk = (s2 >> 16) & 255;
        pos = Bk * Wj * Wi + Sk * (Bj * Wi + Sj * Bi) + (Lk * Sj + Lj) * Si + Li;
      }
      bool cfix[DOF];
      double cval[DOF];
#pragma unroll
      for (int c = 0; c < DOF; c++) { cfix[c] = false; cval[c] = 0.0; }
      bool isdiag = false;
      if (fixing) {
        const int cc = colcode[cjk];
        int ccode[3] = {0, cc & 3, cc >> 2};
        if (!kp.periodic[0]) { int col = fi + ci; ccode[0] = (col == 0) ? 1 : ((col == kp.nnp[0] - 1) ? 2 : 0); }
        if (ccode[0] | ccode[1] | ccode[2]) {
          node_fix<DOF>(kp, ccode, cfix, cval);
          if (kp.fixtable) {
            const int cj = cjk % Wj, ck = cjk / Wj;
            // the column node inside this rank's ghost box (single-rank use; see kron_applicable)
            const int hi = wrapi(fi + ci, kp.nnp[0]) - kp.gs[0], hj = wrapi(fj + cj, kp.nnp[1]) - kp.gs[1], hk = wrapi(fk + ck, kp.nnp[2]) - kp.gs[2];
            const int gidx = hi + kp.gw[0] * (hj + kp.gw[1] * hk);
#pragma unroll
            for (int c = 0; c < DOF; c++) if (cfix[c]) cval[c] = kp.fixtable[(size_t)gidx * DOF + c];
          }
        }
        const int cj = cjk % Wj, ck = cjk / Wj;
        isdiag = (ci == dci) && (cj == dcj) && (ck == dck);
      }
#pragma unroll
      for (int i = 0; i < DOF; i++)
#pragma unroll
        for (int j = 0; j < DOF; j++) {
          const int ij = i * DOF + j;
          double v = a[0] * G[0][ij][cjk];
          v = fma(a[1], G[1][ij][cjk], v);
          v = fma(a[2], G[2][ij][cjk], v);
          v = fma(a[3], G[3][ij][cjk], v);
          if (fixing) {
            if (rfix[i]) v = (isdiag && i == j) ? nelem : 0.0;
            else if (cfix[j]) { racc[i] -= v * cval[j]; v = 0.0; }
          }
          if (want_mat) {
            size_t off;
            if (DOF == 1) off = (size_t)(base + pos);
            else if (kp.block) off = (size_t)(base + pos) * DOF * DOF + j * DOF + i;
            else off = (size_t)base * DOF * DOF + (size_t)i * W * DOF + (size_t)pos * DOF + j;
            kp.values[off] = v;
          }
        }
    }
    if (want_vec) {
#pragma unroll
      for (int c = 0; c < DOF; c++) {
        double r = racc[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
        racc[c] = r;
      }
      if (lane == 0) {
#pragma unroll
        for (int c = 0; c < DOF; c++) {
          double F = 0.0;
          for (int n = 0; n < kp.nvterms; n++) {
            const KronVTerm vt = kp.vterms[n];
            if (vt.i != c) continue;
            F += vt.c * kp.mv[0][vt.r0 * kp.nnp[0] + Ai] * kp.mv[1][vt.r1 * kp.nnp[1] + Aj] * kp.mv[2][vt.r2 * kp.nnp[2] + Ak];
          }
          if (fixing) {
            // AddFlux: loads on the faces this node lies on, summed over the face elements that contain it
            for (int d = 0; d < kp.dim; d++) {
              if (!rcode[d]) continue;
              const FixSide& fs = kp.bc[d][rcode[d] - 1];
              if (!fs.lcount) continue;
              double A = 1.0;
              if (kp.dim > 1) {
                const int An[3] = {Ai, Aj, Ak};
                for (int e2 = 0; e2 < kp.dim; e2++) if (e2 != d) A *= kp.lsum[e2][An[e2]];
                A *= (kp.dim == 2) ? 2 : 4;
              }
              for (int k = 0; k < fs.lcount; k++) if (fs.lfield[k] == c) F += fs.lvalue[k] * A;
            }
            F += racc[c];
            if (rfix[c]) F = nelem * rval[c];
          }
          kp.rhs[(size_t)lr * DOF + c] = F;
        }
      }
    }
  }
}

template <int DIM, int DOF>
void host_terms(int form, int slot, const double* prm, const FormInfo& fi, KronParams& kp) {
  const int NA = fi.mc1 - fi.mc0, NV = fi.vc1 - fi.vc0;
  std::vector<double> C((size_t)DOF * DOF * std::max(NA, 1) * std::max(NA, 1), 0.0), fv((size_t)DOF * std::max(NV, 1), 0.0);
  QPoint q;
  memset(&q, 0, sizeof(q));
  form_coefficients<DIM, DOF>(form, slot, prm, 0.0, 0.0, q, NA, NV, NA ? C.data() : nullptr, NV ? fv.data() : nullptr);
  kp.nterms = kp.nvterms = kp.rsmask0 = 0;
  for (int i = 0; i < DOF; i++)
    for (int j = 0; j < DOF; j++)
      for (int al = 0; al < NA; al++)
        for (int be = 0; be < NA; be++) {
          double c = C[(((size_t)i * DOF + j) * NA + al) * NA + be];
          if (c == 0.0) continue;
          KronTerm t;
          const int ca = al + fi.mc0, cb = be + fi.mc0;
          unsigned char rs[3];
          for (int d = 0; d < 3; d++) rs[d] = (unsigned char)(((ca == 1 + d) ? 2 : 0) + ((cb == 1 + d) ? 1 : 0));
          t.ij = (unsigned char)(i * DOF + j); t.rs0 = rs[0]; t.rs1 = rs[1]; t.rs2 = rs[2]; t.c = c;
          if (kp.nterms < kMaxTerms) kp.terms[kp.nterms] = t;
          kp.nterms++;
          kp.rsmask0 |= 1 << rs[0];
        }
  for (int i = 0; i < DOF; i++)
    for (int al = 0; al < NV; al++) {
      double c = fv[(size_t)i * NV + al];
      if (c == 0.0) continue;
      const int ca = al + fi.vc0;
      KronVTerm t;
      t.i = (unsigned char)i; t.r0 = (ca == 1); t.r1 = (ca == 2); t.r2 = (ca == 3); t.c = c;
      if (kp.nvterms < 8) kp.vterms[kp.nvterms] = t;
      kp.nvterms++;
    }
}

}  // namespace

bool kron_applicable(const petiga_cuda_plan* P, int slot, int form) {
  const Layout& L = P->L;
  if (slot != PETIGA_SLOT_VECTOR && slot != PETIGA_SLOT_MATRIX && slot != PETIGA_SLOT_SYSTEM) return false;
  if (P->d_X) return false;                                  // mapped geometry: coefficients vary per point
  FormInfo fi = form_info(form, slot, L.dim, L.dof);
  if (!fi.valid || fi.per_qp || !fi.constant_f) return false;
  if (L.dof > 3) return false;
  if (P->d_fixtable && L.nranks > 1) return false;          // table values of off-box columns are not local
  for (int d = 0; d < L.dim; d++)
    if (L.ax[d].periodic && L.ax[d].nnp < 2 * L.ax[d].p + 1) return false;
  return true;
}

int launch_kronecker(petiga_cuda_plan* P, int slot, int block, double* values, double* rhs) {
  const Layout& L = P->L;
  const int form = P->slots[slot].form;
  FormInfo fi = form_info(form, slot, L.dim, L.dof);
  if (!P->d_kronrow[0]) {   // 1-D tables, once per plan
    for (int d = 0; d < 3; d++) {
      const int nnp = L.ax[d].nnp;
      void* buf = nullptr;
      const size_t nM = (size_t)4 * nnp * kMaxW, nmv = (size_t)2 * nnp;
      PC_CUDA(cudaMalloc(&buf, (nM + nmv + nnp) * sizeof(double) + (size_t)nnp * sizeof(int)));
      P->allocs.push_back(buf);
      P->d_kronrow[d] = (double*)buf;
      double* M = (double*)buf; double* mv = M + nM; double* lsum = mv + nmv; int* nsup = (int*)(lsum + nnp);
      const int total = (int)nM;
      kron_1d_kernel<<<(total + 127) / 128, 128, 0, P->stream>>>(P->dax[d], P->d_first[d], M, mv, nsup, lsum);
      PC_CUDA(cudaGetLastError());
      P->launches++;
    }
  }
  KronParams kp;
  memset(&kp, 0, sizeof(kp));
  bool simple = (L.nranks == 1);
  for (int d = 0; d < 3; d++) {
    const int nnp = L.ax[d].nnp;
    const size_t nM = (size_t)4 * nnp * kMaxW, nmv = (size_t)2 * nnp;
    kp.M[d] = P->d_kronrow[d]; kp.mv[d] = kp.M[d] + nM; kp.lsum[d] = kp.mv[d] + nmv; kp.nsup[d] = (const int*)(kp.lsum[d] + nnp);
    kp.first[d] = P->d_first[d]; kp.Wg[d] = P->dax[d].W; kp.lo[d] = P->dax[d].lo; kp.seg[d] = P->dax[d].seg;
    kp.ls[d] = L.ax[d].ls; kp.lw[d] = L.ax[d].lw; kp.gs[d] = L.ax[d].gs; kp.gw[d] = L.ax[d].gw; kp.nnp[d] = nnp; kp.periodic[d] = L.ax[d].periodic;
    if (L.ax[d].periodic) simple = false;
  }
  kp.rowbase = P->d_rowbase; kp.localrow = P->d_localrow; kp.fixtable = P->d_fixtable;
  kp.dim = L.dim; kp.dof = L.dof; kp.block = block; kp.slot = slot; kp.simple = simple;
  kp.values = values; kp.rhs = rhs;
  const double* prm = P->slots[slot].prm;
#define HT(DIM_, DOF_) if (L.dim == DIM_ && L.dof == DOF_) host_terms<DIM_, DOF_>(form, slot, prm, fi, kp);
  HT(1, 1) HT(1, 2) HT(1, 3) HT(2, 1) HT(2, 2) HT(2, 3) HT(3, 1) HT(3, 2) HT(3, 3)
#undef HT
  if (kp.nterms > kMaxTerms || kp.nvterms > 8) { set_error("separable path: too many coefficient terms"); return PETIGA_CUDA_ERR_SUP; }
  if (slot == PETIGA_SLOT_SYSTEM && P->has_bc)
    for (int d = 0; d < L.dim; d++)
      for (int s = 0; s < 2; s++) {
        FixSide& fs = kp.bc[d][s];
        for (int k = 0; k < P->bc.vcount[d][s]; k++) {
          int c = P->bc.vfield[d][s][k];
          if (c >= L.dof) continue;
          fs.vfield[fs.vcount] = c; fs.vvalue[fs.vcount] = P->bc.vvalue[d][s][k]; fs.vcount++;
        }
        for (int k = 0; k < P->bc.lcount[d][s]; k++) {
          int c = P->bc.lfield[d][s][k];
          if (c >= L.dof) continue;
          fs.lfield[fs.lcount] = c; fs.lvalue[fs.lcount] = P->bc.lvalue[d][s][k]; fs.lcount++;
        }
        if (!L.ax[d].periodic && (fs.vcount || fs.lcount)) kp.any_bc = 1;
      }
  const int blocks = L.ax[1].lw * L.ax[2].lw;
  const int threads = std::min(256, std::max(32, ((L.ax[0].lw + 0) * 32)));
#define KL(DOF_)                                                                              \
  if (L.dof == DOF_) {                                                                        \
    if (simple) kron_rows_kernel<DOF_, true><<<blocks, threads, 0, P->stream>>>(kp);          \
    else kron_rows_kernel<DOF_, false><<<blocks, threads, 0, P->stream>>>(kp);                \
  }
  KL(1) KL(2) KL(3)
#undef KL
  PC_CUDA(cudaGetLastError());
  P->launches++;
  return 0;
}

}  // namespace pc
