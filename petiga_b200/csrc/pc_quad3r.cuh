// pc_quad3r.cuh -- quad_sf3r_kernel: the third-generation matrix kernel with the pencil window carried in the DMMA accumulators.
//
// Same contract and the same stages as quad_sf3_kernel (pc_quad3.cuh): the reference's element loop (src/petigaelem.c:375-410,
// 693-1033,1360-1389,1525-1559; src/petigapoint.c:451-465) for 3-D, dof 1, degree 3, 4-point rule, constant-coefficient
// first-order forms on identity or mapped geometry.  It applies when consecutive elements of a pencil advance by ONE axis-0 row
// (maximally smooth axis, the headline configuration); other knot vectors run quad_sf3_kernel.
//
// What the ncu capture of quad_sf3_kernel showed (profiles/r2_ncu_sf3_v4_mesh64.json): 4 050 cycles per element against a DMMA
// floor of 1 056; 8 192 + 3 584 shared-memory accesses per element for the window S (+= of the 4 096 element entries, read + clear
// of the 1 792 flushed ones), two CTA barriers per element, the flush on every warp's critical path.  Here:
//  * a row (i0 + a0) receives its entries from the four consecutive elements with a0 = 3, 2, 1, 0.  A warp pair owns the rows with
//    (row mod 4) = rs; within the pair a warp owns the column offsets c0 = b0 - a0 + 3 of one parity.  Stage C's accumulator
//    fragments are therefore simply NOT cleared between those four elements: mma.sync accumulates the window in registers
//    (4 keys x 8 doubles per lane), and every warp has exactly two (a0, b0) combinations per element -- the same DMMA count as before.
//  * after the element with a0 = 0 the pair writes the finished row (1 792 entries) once into a staging buffer in flush order and
//    hands it to two FLUSH warps through an mbarrier pair; they issue the coalesced red.global.add.f64 while the assembly warps
//    go on: 1 792 STS + 1 792 LDS per element instead of 11 776 accesses, and no flush latency on the assembly warps.
//  * stages A + B of element e + 1 run in the same phase as stage C of element e (U2 double-buffered): one CTA barrier per element.
#pragma once
#include "pc_quad3.cuh"

namespace pc {

constexpr int k3rFlushWarps = 2;
constexpr int k3rThreads = k3AsmThreads + 32 + 32 * k3rFlushWarps;   // 8 assembly warps, 1 producer warp, flush warps
constexpr int k3rU2 = 4 * 4 * k3U2Q;                                  // one U2 buffer
constexpr int k3rMaxOps = 32;                                         // stage A + B operations (one per pair and q2) of one warp
constexpr int k3rStage = 16 * k3SA;                                   // one staging buffer: [a2][a1][e = c0 + 7 (b1 + 4 b2)] (+1 pad)

struct SF3RSmem {   // offsets in doubles
  int U2, PP1, PP2, ring, stage, rowoff, seg0, W0, wj0, PT, fix, prog, cc, bars, total, slot;
  __host__ __device__ SF3RSmem(int npairs, int mapped) {
    int o = 0;
    U2 = o; o += 2 * k3rU2;
    PP1 = o; o += 576; PP2 = o; o += 576;
    slot = 576 + (mapped ? npairs * 64 : 0);
    ring = o; o += k3Ring * slot;
    stage = o; o += 2 * k3rStage;
    rowoff = o; o += k3MaxRows * 16;
    seg0 = o; o += k3MaxRows * 8 / 2;
    W0 = o; o += k3MaxRows / 2;
    wj0 = o; o += k3MaxSeg * 4 + 4;
    PT = o; o += 128;
    fix = o; o += 2 * (64 + 32 + 32);                  // two buffers of fixval[64], fixflag int[64], rowlr int[64]
    prog = o; o += 8 * k3rMaxOps / 2;                  // uint32 [assembly warp][op]: the warp's stage A + B program
    cc = o; o += k3MaxPairs;                           // cconst
    bars = o; o += 2 * k3Ring + 4;
    total = o;
  }
};

// fire-and-forget reduction: atomicAdd() in the flush warps' loop compiled to ATOMG (the warp waits for the returned value, 245 cycles
// per instruction with two warps); the PTX red has no destination, so ptxas must emit REDG
__device__ __forceinline__ void red_add_f64(double* p, double v) {
  asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "d"(v) : "memory");
}

// Stages A + B of one element for one warp.  The (g2, q2) combinations of a warp never change, so the nested loops over the
// form's lists (g1 groups of g2, pairs of g1) are flattened once per launch into a program in shared memory: one word per pair,
//   oo0 | pr << 4 | oo1 << 8 | q2 << 12 | g2 << 14 | first-of-g1 << 16 | last-of-g1 << 17 | first-of-combination << 18 | last << 19.
// The operands of operation i + 1 are loaded while the DMMAs of operation i issue (the loop-and-list version spent ~1 000 cycles per
// element in dependent constant loads -> address -> LDS -> DMMA chains, profiles/r2_ncu_sf3r_v4_mesh64.json).
struct SF3ROperands { double pa0, pa1, bf, pb0, pb1; };
__device__ __forceinline__ SF3ROperands sf3r_load_op(uint32_t op, bool mapped, const double* P0, const double* D, const double* PP1, const double* ccS,
                                                     int lane, int dfrag, double w0q, const double (&wq12)[2]) {
  SF3ROperands o;
  const int oo0 = op & 15, pr = (op >> 4) & 15, oo1 = (op >> 8) & 15, b = (op >> 13) & 1;
  const double* pa = P0 + oo0 * 64 + lane;
  o.pa0 = pa[0]; o.pa1 = pa[32];
  o.bf = mapped ? D[pr * 64 + dfrag + 32 * b] : ccS[pr] * (w0q * (b ? wq12[1] : wq12[0]));
  const double* pb = PP1 + oo1 * 64 + lane;
  o.pb0 = pb[0]; o.pb1 = pb[32];
  return o;
}
__device__ __forceinline__ void sf3r_stage_ab(const uint32_t* prog, int nops, bool mapped, const double* P0, const double* D, const double* PP1,
                                              const double* ccS, double* U2, int lane, int r, int c, int dfrag, double w0q, const double (&wq12)[2]) {
  if (nops == 0) return;
  uint32_t op = prog[0];
  SF3ROperands nx = sf3r_load_op(op, mapped, P0, D, PP1, ccS, lane, dfrag, w0q, wq12);
  double cA[2][2] = {{0, 0}, {0, 0}}, cB[2][2][2] = {{{0, 0}, {0, 0}}, {{0, 0}, {0, 0}}};
  for (int i = 0; i < nops; i++) {
    const uint32_t cur = op;
    const SF3ROperands o = nx;
    if (i + 1 < nops) { op = prog[i + 1]; nx = sf3r_load_op(op, mapped, P0, D, PP1, ccS, lane, dfrag, w0q, wq12); }
    if (cur & (1u << 18)) {
#pragma unroll
      for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int nt = 0; nt < 2; nt++) cB[mt][nt][0] = cB[mt][nt][1] = 0.0;
    }
    if (cur & (1u << 16)) { cA[0][0] = cA[0][1] = cA[1][0] = cA[1][1] = 0.0; }
    dmma(cA[0][0], cA[0][1], o.pa0, o.bf);        // stage A: U1[ab0 = r + 8 mt][q1 = c][q2 = 2 b + reg] += PP0 x D'
    dmma(cA[1][0], cA[1][1], o.pa1, o.bf);
    if (cur & (1u << 17)) {                       // stage B: stage A's accumulators are the A operand as they stand
      const bool x = (cur >> 12) & 1;
      const double ax0 = x ? cA[0][1] : cA[0][0], ax1 = x ? cA[1][1] : cA[1][0];
      dmma(cB[0][0][0], cB[0][0][1], ax0, o.pb0);
      dmma(cB[0][1][0], cB[0][1][1], ax0, o.pb1);
      dmma(cB[1][0][0], cB[1][0][1], ax1, o.pb0);
      dmma(cB[1][1][0], cB[1][1][1], ax1, o.pb1);
    }
    if (cur & (1u << 19)) {
      double* u = U2 + ((cur >> 12) & 15) * k3U2Q + r * 24 + 2 * c;       // bits 12..15 = g2 * 4 + q2
#pragma unroll
      for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int nt = 0; nt < 2; nt++)
          *reinterpret_cast<double2*>(u + mt * 8 * 24 + nt * 8) = make_double2(cB[mt][nt][0], cB[mt][nt][1]);
    }
  }
}

// ---- compile-time form structures.  The two list structures that matter -- first-order forms with a diagonal coefficient tensor
//      on identity geometry (Poisson, Laplace: pairs G0G0, G1G1, G2G2) and on mapped geometry (all nine gradient pairs) -- are written
//      down as constants, in the order build_sf_lists (pc_quad2.cu) produces them; the launcher compares the run-time lists with them
//      and picks the specialised kernel only on an exact match (anything else runs the generic program from shared memory).  With the
//      structure known, a warp's stage A + B program is straight-line code: no flags to test, immediate offsets, loads hoisted by the
//      compiler (the interpreted program cost ~97 warp instructions per operation, a quarter of the kernel's instructions on identity
//      and ~40 % on mapped geometry, profiles/r2_ncu_sf3r_v5_mesh64.json). ----
struct SF3RStruct { int ng2, ng1, npairs; int g2_first[5], g2_oo2[4], g1_first[17], g1_oo1[16], pair_oo0[16]; };
constexpr SF3RStruct k3rStructDiag = {2, 3, 3, {0, 2, 3, 0, 0}, {0, 4, 0, 0}, {0, 1, 2, 3}, {0, 4, 0}, {4, 0, 0}};
constexpr SF3RStruct k3rStructFull = {4, 9, 9, {0, 4, 6, 8, 9}, {0, 1, 3, 4}, {0, 1, 2, 3, 4, 5, 6, 7, 8, 9}, {0, 1, 3, 4, 0, 3, 0, 1, 0}, {4, 3, 1, 0, 3, 0, 1, 0, 0}};
struct SF3RProg { uint32_t op[8][8]; int n[8]; };
constexpr SF3RProg sf3r_make_prog(const SF3RStruct& L) {      // the same longest-first deal as the run-time builder in the kernel
  SF3RProg P = {};
  bool done[16] = {};
  const int ncmb = L.ng2 * 4;
  for (int k = 0; k < ncmb; k++) {
    int best = -1, bestn = -1;
    for (int cmb = 0; cmb < ncmb; cmb++) {
      if (done[cmb]) continue;
      const int g2 = cmb >> 2;
      const int nop = L.g1_first[L.g2_first[g2 + 1]] - L.g1_first[L.g2_first[g2]];
      if (nop > bestn) { bestn = nop; best = cmb; }
    }
    done[best] = true;
    int wmin = 0;
    for (int w = 1; w < 8; w++) if (P.n[w] < P.n[wmin]) wmin = w;
    const int g2 = best >> 2, q2 = best & 3;
    bool firstc = true;
    for (int g1 = L.g2_first[g2]; g1 < L.g2_first[g2 + 1]; g1++)
      for (int pr = L.g1_first[g1]; pr < L.g1_first[g1 + 1]; pr++) {
        const bool f1 = pr == L.g1_first[g1], l1 = pr + 1 == L.g1_first[g1 + 1], lc = l1 && g1 + 1 == L.g2_first[g2 + 1];
        P.op[wmin][P.n[wmin]++] = (uint32_t)L.pair_oo0[pr] | ((uint32_t)pr << 4) | ((uint32_t)L.g1_oo1[g1] << 8) | ((uint32_t)q2 << 12) | ((uint32_t)g2 << 14) |
                                  ((uint32_t)f1 << 16) | ((uint32_t)l1 << 17) | ((uint32_t)firstc << 18) | ((uint32_t)lc << 19);
        firstc = false;
      }
  }
  return P;
}
template <int FS> struct SF3RProgOf;
template <> struct SF3RProgOf<1> { static constexpr SF3RProg P = sf3r_make_prog(k3rStructDiag); static constexpr bool mapped = false; };
template <> struct SF3RProgOf<2> { static constexpr SF3RProg P = sf3r_make_prog(k3rStructFull); static constexpr bool mapped = true; };

template <uint32_t OP, bool MAPPED>
__device__ __forceinline__ void sf3r_op(double (&cA)[2][2], double (&cB)[2][2][2], const double* P0, const double* D, const double* PP1, const double* ccS,
                                        double* U2, int lane, int r, int c, int dfrag, double w0q, const double (&wq12)[2]) {
  constexpr int oo0 = OP & 15, pr = (OP >> 4) & 15, oo1 = (OP >> 8) & 15, q2 = (OP >> 12) & 3, b = q2 >> 1, x = q2 & 1;
  constexpr bool f1 = (OP >> 16) & 1, l1 = (OP >> 17) & 1, fc = (OP >> 18) & 1, lc = (OP >> 19) & 1;
  const double* pa = P0 + oo0 * 64 + lane;
  const double bf = MAPPED ? D[pr * 64 + dfrag + 32 * b] : ccS[pr] * (w0q * wq12[b]);
  if (fc) {
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
      for (int nt = 0; nt < 2; nt++) cB[mt][nt][0] = cB[mt][nt][1] = 0.0;
  }
  if (f1) { cA[0][0] = cA[0][1] = cA[1][0] = cA[1][1] = 0.0; }
  dmma(cA[0][0], cA[0][1], pa[0], bf);
  dmma(cA[1][0], cA[1][1], pa[32], bf);
  if (l1) {
    const double* pb = PP1 + oo1 * 64 + lane;
    const double b0 = pb[0], b1 = pb[32];
    dmma(cB[0][0][0], cB[0][0][1], cA[0][x], b0);
    dmma(cB[0][1][0], cB[0][1][1], cA[0][x], b1);
    dmma(cB[1][0][0], cB[1][0][1], cA[1][x], b0);
    dmma(cB[1][1][0], cB[1][1][1], cA[1][x], b1);
  }
  if (lc) {
    double* u = U2 + ((OP >> 12) & 15) * k3U2Q + r * 24 + 2 * c;
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
      for (int nt = 0; nt < 2; nt++)
        *reinterpret_cast<double2*>(u + mt * 8 * 24 + nt * 8) = make_double2(cB[mt][nt][0], cB[mt][nt][1]);
  }
}
template <int FS, int W, int I>
__device__ __forceinline__ void sf3r_ops_from(double (&cA)[2][2], double (&cB)[2][2][2], const double* P0, const double* D, const double* PP1, const double* ccS,
                                              double* U2, int lane, int r, int c, int dfrag, double w0q, const double (&wq12)[2]) {
  if constexpr (I < SF3RProgOf<FS>::P.n[W]) {
    sf3r_op<SF3RProgOf<FS>::P.op[W][I], SF3RProgOf<FS>::mapped>(cA, cB, P0, D, PP1, ccS, U2, lane, r, c, dfrag, w0q, wq12);
    sf3r_ops_from<FS, W, I + 1>(cA, cB, P0, D, PP1, ccS, U2, lane, r, c, dfrag, w0q, wq12);
  }
}
template <int FS>
__device__ __forceinline__ void sf3r_stage_ab_static(int warp, const double* P0, const double* D, const double* PP1, const double* ccS, double* U2,
                                                     int lane, int r, int c, int dfrag, double w0q, const double (&wq12)[2]) {
  double cA[2][2] = {{0, 0}, {0, 0}}, cB[2][2][2] = {{{0, 0}, {0, 0}}, {{0, 0}, {0, 0}}};
  switch (warp) {   // warp-uniform: each warp runs its own straight-line program
    case 0: sf3r_ops_from<FS, 0, 0>(cA, cB, P0, D, PP1, ccS, U2, lane, r, c, dfrag, w0q, wq12); break;
    case 1: sf3r_ops_from<FS, 1, 0>(cA, cB, P0, D, PP1, ccS, U2, lane, r, c, dfrag, w0q, wq12); break;
    case 2: sf3r_ops_from<FS, 2, 0>(cA, cB, P0, D, PP1, ccS, U2, lane, r, c, dfrag, w0q, wq12); break;
    case 3: sf3r_ops_from<FS, 3, 0>(cA, cB, P0, D, PP1, ccS, U2, lane, r, c, dfrag, w0q, wq12); break;
    case 4: sf3r_ops_from<FS, 4, 0>(cA, cB, P0, D, PP1, ccS, U2, lane, r, c, dfrag, w0q, wq12); break;
    case 5: sf3r_ops_from<FS, 5, 0>(cA, cB, P0, D, PP1, ccS, U2, lane, r, c, dfrag, w0q, wq12); break;
    case 6: sf3r_ops_from<FS, 6, 0>(cA, cB, P0, D, PP1, ccS, U2, lane, r, c, dfrag, w0q, wq12); break;
    default: sf3r_ops_from<FS, 7, 0>(cA, cB, P0, D, PP1, ccS, U2, lane, r, c, dfrag, w0q, wq12); break;
  }
}

// stage C of one element for the warp (rs, par) at local row a0: its two (a0, b0) combinations, b0 = B0 and B0 + 2 with
// B0 = (par + a0 + 1) & 1, accumulate onto the keys K0 and K0 + 1 of the running row, K0 = (B0 - a0 + 3) >> 1.
// acc[k][nt][mt][reg]: entry (ab2 = r + 8 mt, ab1 = 2 c + reg + 8 nt) at column offset c0 = 2 k + par.  Only the key index has to be a
// compile-time constant (registers); everything else is an address.
// FS > 0: the number of g2 groups and their order pairs are compile-time constants (all loads of the loop can be issued up front);
// the values are those of k3rStructDiag / k3rStructFull (checked at compile time below)
__host__ __device__ constexpr int sf3r_ng2(int fs) { return fs == 1 ? 2 : 4; }
__host__ __device__ constexpr int sf3r_g2_oo2(int fs, int g2) { return fs == 1 ? (g2 == 0 ? 0 : 4) : (g2 == 0 ? 0 : g2 == 1 ? 1 : g2 == 2 ? 3 : 4); }
static_assert(sf3r_ng2(1) == k3rStructDiag.ng2 && sf3r_ng2(2) == k3rStructFull.ng2, "structure constants");
static_assert(sf3r_g2_oo2(1, 0) == k3rStructDiag.g2_oo2[0] && sf3r_g2_oo2(1, 1) == k3rStructDiag.g2_oo2[1], "structure constants");
static_assert(sf3r_g2_oo2(2, 0) == k3rStructFull.g2_oo2[0] && sf3r_g2_oo2(2, 1) == k3rStructFull.g2_oo2[1] &&
              sf3r_g2_oo2(2, 2) == k3rStructFull.g2_oo2[2] && sf3r_g2_oo2(2, 3) == k3rStructFull.g2_oo2[3], "structure constants");
template <int K0, int FS>
__device__ __forceinline__ void sf3r_stage_c(const SFLists& ls, const double* PP2, const double* ub0, double (&acc)[4][2][2][2], int lane) {
  constexpr int NG2 = FS > 0 ? sf3r_ng2(FS) : 0;
  const int ng2 = FS > 0 ? NG2 : ls.ng2;
#pragma unroll
  for (int g2 = 0; g2 < (FS > 0 ? NG2 : 4); g2++) {
    if (g2 >= ng2) break;
    const int oo2 = FS > 0 ? sf3r_g2_oo2(FS, g2) : ls.g2_oo2[g2];
    const double* pa = PP2 + oo2 * 64 + lane;
    const double a0f = pa[0], a1f = pa[32];
    const double* ub = ub0 + g2 * 4 * k3U2Q;
#pragma unroll
    for (int hb = 0; hb < 2; hb++)
#pragma unroll
      for (int nt = 0; nt < 2; nt++) {
        const double bf = ub[hb * 48 + nt * 8];
        dmma(acc[K0 + hb][nt][0][0], acc[K0 + hb][nt][0][1], a0f, bf);
        dmma(acc[K0 + hb][nt][1][0], acc[K0 + hb][nt][1][1], a1f, bf);
      }
  }
}

// boundary element: the element's own entries are needed for IGAElementFixSystem (petigaelem.c:1360-1389) before they join the row
__device__ __noinline__ void sf3r_stage_c_fix(const KParams& prm, const SFLists& ls, const double* PP2, const double* ub0, double (&T)[2][2][2][2],
                                              int lane, int a0, int b00, const double* fixval, const int* fixflag, const int* rowlr) {
  const int r = lane >> 2, c = lane & 3;
#pragma unroll
  for (int hb = 0; hb < 2; hb++)
#pragma unroll
    for (int nt = 0; nt < 2; nt++)
#pragma unroll
      for (int mt = 0; mt < 2; mt++) T[hb][nt][mt][0] = T[hb][nt][mt][1] = 0.0;
  for (int g2 = 0; g2 < ls.ng2; g2++) {
    const double* pa = PP2 + ls.g2_oo2[g2] * 64 + lane;
    const double a0f = pa[0], a1f = pa[32];
    const double* ub = ub0 + g2 * 4 * k3U2Q;
#pragma unroll
    for (int hb = 0; hb < 2; hb++)
#pragma unroll
      for (int nt = 0; nt < 2; nt++) {
        const double bf = ub[hb * 48 + nt * 8];
        dmma(T[hb][nt][0][0], T[hb][nt][0][1], a0f, bf);
        dmma(T[hb][nt][1][0], T[hb][nt][1][1], a1f, bf);
      }
  }
#pragma unroll
  for (int hb = 0; hb < 2; hb++)
#pragma unroll
    for (int nt = 0; nt < 2; nt++)
#pragma unroll
      for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int x = 0; x < 2; x++) {
          const double v = T[hb][nt][mt][x];
          const int a2 = 2 * mt + (r >> 2), b2 = r & 3, a1 = 2 * nt + (c >> 1), b1 = 2 * (c & 1) + x, b0 = b00 + 2 * hb;
          const int ra = a0 + 4 * a1 + 16 * a2, cb = b0 + 4 * b1 + 16 * b2;
          const bool fr = fixflag[ra], fcx = fixflag[cb];
          if (fr || fcx) {
            if (fcx && !fr) red_add_f64(&prm.rhs[rowlr[ra]], -v * fixval[cb]);
            T[hb][nt][mt][x] = (ra == cb) ? 1.0 : 0.0;
          }
        }
}

template <int K0>
__device__ __forceinline__ void sf3r_add(double (&acc)[4][2][2][2], const double (&T)[2][2][2][2]) {
#pragma unroll
  for (int hb = 0; hb < 2; hb++)
#pragma unroll
    for (int nt = 0; nt < 2; nt++)
#pragma unroll
      for (int mt = 0; mt < 2; mt++) { acc[K0 + hb][nt][mt][0] += T[hb][nt][mt][0]; acc[K0 + hb][nt][mt][1] += T[hb][nt][mt][1]; }
}

// the pair's finished (or, at the end of a segment, partial) row goes to the staging buffer in flush order
template <int PAR>
__device__ __forceinline__ void sf3r_stage_row(double* stg, double (&acc)[4][2][2][2], int r, int c) {
  double* base = stg + ((r >> 2) * 4 + (c >> 1)) * k3SA + (2 * (c & 1) + 4 * (r & 3)) * 7 + PAR;
#pragma unroll
  for (int k = 0; k < 4 - PAR; k++)
#pragma unroll
    for (int nt = 0; nt < 2; nt++)
#pragma unroll
      for (int mt = 0; mt < 2; mt++) {
        double* p = base + (8 * mt + 2 * nt) * k3SA + 2 * k;
        p[0] = acc[k][nt][mt][0];
        p[7] = acc[k][nt][mt][1];
      }
#pragma unroll
  for (int k = 0; k < 4; k++)
#pragma unroll
    for (int nt = 0; nt < 2; nt++)
#pragma unroll
      for (int mt = 0; mt < 2; mt++) acc[k][nt][mt][0] = acc[k][nt][mt][1] = 0.0;
}

template <int FS>      // 0: the form's lists interpreted at run time; 1 / 2: k3rStructDiag / k3rStructFull compiled in
__global__ void __launch_bounds__(k3rThreads, 1) quad_sf3r_kernel(const __grid_constant__ SF3Params sp) {
  const KParams& prm = sp.k;
  const SFLists& ls = sp.l;
  extern __shared__ __align__(128) double sm3[];
  const bool mapped = prm.X != nullptr;
  const SF3RSmem lay(ls.npairs, mapped ? 1 : 0);
  double *U2 = sm3 + lay.U2, *PP1 = sm3 + lay.PP1, *PP2 = sm3 + lay.PP2, *Ring = sm3 + lay.ring, *Stage = sm3 + lay.stage;
  int64_t* rowoffT = reinterpret_cast<int64_t*>(sm3 + lay.rowoff);
  uint32_t* seg0T = reinterpret_cast<uint32_t*>(sm3 + lay.seg0);
  int* W0T = reinterpret_cast<int*>(sm3 + lay.W0);
  double* wj0T = sm3 + lay.wj0;
  uint32_t* PT = reinterpret_cast<uint32_t*>(sm3 + lay.PT);
  double* fixbuf = sm3 + lay.fix;
  uint32_t* progS = reinterpret_cast<uint32_t*>(sm3 + lay.prog);
  double* ccS = sm3 + lay.cc;
  uint64_t* full = reinterpret_cast<uint64_t*>(sm3 + lay.bars);
  uint64_t* empty = full + k3Ring;
  uint64_t* staged = empty + k3Ring;            // [2]: a staging buffer holds a row (arrivals: the two warps of the pair)
  uint64_t* drained = staged + 2;               // [2]: the flush warps have read it (arrivals: one per flush warp).  Only the pair that
                                                //      stages event n waits here, for event n - 2: it is never more than one phase away from
                                                //      the barrier, which is what a parity wait needs; everybody else meets the flush warps at
                                                //      the end of a work item on a named barrier (bar.sync 2)
  const int tid = threadIdx.x;
  const int ew0 = prm.ax[0].ew, ew1 = prm.ax[1].ew;
  const int nwork = sp.npencils * sp.nseg;
  const uint32_t slot_bytes = (uint32_t)lay.slot * 8;

  if (tid == 0) {
    for (int k = 0; k < k3Ring; k++) { mbar_init(&full[k], 1); mbar_init(&empty[k], k3AsmThreads / 32); }
    for (int k = 0; k < 2; k++) { mbar_init(&staged[k], 2); mbar_init(&drained[k], k3rFlushWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < k3MaxPairs) ccS[tid] = sp.cconst[tid];
  if (tid == 0) {   // the warps' stage A + B programs (see sf3r_stage_ab): combinations dealt out longest first to the least loaded warp
    int load[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int ncmb = ls.ng2 * 4;
    bool done[16];
    for (int k = 0; k < 16; k++) done[k] = false;
    for (int k = 0; k < ncmb; k++) {
      int best = -1, bestn = -1;
      for (int cmb = 0; cmb < ncmb; cmb++) {
        if (done[cmb]) continue;
        const int g2 = cmb >> 2;
        const int nop = ls.g1_first[ls.g2_first[g2 + 1]] - ls.g1_first[ls.g2_first[g2]];
        if (nop > bestn) { bestn = nop; best = cmb; }
      }
      done[best] = true;
      int wmin = 0;
      for (int w = 1; w < 8; w++) if (load[w] < load[wmin]) wmin = w;
      uint32_t* pg = progS + wmin * k3rMaxOps;
      int n = load[wmin];
      const int g2 = best >> 2, q2 = best & 3;
      bool firstc = true;
      for (int g1 = ls.g2_first[g2]; g1 < ls.g2_first[g2 + 1]; g1++)
        for (int pr = ls.g1_first[g1]; pr < ls.g1_first[g1 + 1]; pr++) {
          const bool f1 = pr == ls.g1_first[g1], l1 = pr + 1 == ls.g1_first[g1 + 1], lc = l1 && g1 + 1 == ls.g2_first[g2 + 1];
          if (n < k3rMaxOps - 1)
            pg[n++] = (uint32_t)ls.pair_oo0[pr] | ((uint32_t)pr << 4) | ((uint32_t)ls.g1_oo1[g1] << 8) | ((uint32_t)q2 << 12) | ((uint32_t)g2 << 14) |
                      ((uint32_t)f1 << 16) | ((uint32_t)l1 << 17) | ((uint32_t)firstc << 18) | ((uint32_t)lc << 19);
          firstc = false;
        }
      load[wmin] = n;
    }
    for (int w = 0; w < 8; w++) progS[w * k3rMaxOps + k3rMaxOps - 1] = (uint32_t)load[w];
  }
  __syncthreads();

  if (tid >= k3AsmThreads + 32) {
    // ===================== FLUSH WARPS: staged rows -> coalesced red.global.add.f64 (IGAElementAssembleMat, petigaelem.c:1525-1541) =====================
    const int lane = tid & 31, fw = (tid - k3AsmThreads - 32) >> 5;
    int fc0[4], fbb[4];
#pragma unroll
    for (int k = 0; k < 4; k++) { const int e = lane + 32 * k; fc0[k] = e % 7; fbb[k] = (e / 7) & 15; }
    uint32_t n = 0;                                                       // flush events so far
    constexpr int NJ = 16 / k3rFlushWarps;
    int A13[NJ][4];                                                       // storage-order rows: pos = A13 * W0 + c0 - c0first
    for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
      const int seg = w % sp.nseg;
      const int le0 = seg * sp.seglen, le1 = min(ew0, le0 + sp.seglen), nel = le1 - le0;
      for (int row = 0; row < nel + 3; row++, n++) {
        const int b = n & 1;
        mbar_wait(&staged[b], (n >> 1) & 1, 1);
        const double* stg = Stage + b * k3rStage;
        if (row == 0) {                                                   // the work item's tables are complete before its first row is staged
#pragma unroll
          for (int j = 0; j < NJ; j++)
#pragma unroll
            for (int k = 0; k < 4; k++) { const uint32_t ptv = PT[(fw * NJ + j) * 16 + fbb[k]]; A13[j][k] = (int)(ptv & 0xFFFF) + (int)(ptv >> 24); }
        }
        const int W0 = W0T[row];
        const uint32_t sfirst = seg0T[row * 8 + 3];
        const int c0first = 3 - (int)((sfirst >> 16) & 255);
        const bool row_simple = (seg0T[row * 8 + 7] == 1u);
        if (row_simple) {   // all loads first, then the reductions back to back: the two warps have to keep up with eight assembly warps
          double vv[NJ][4];
          int64_t ro[NJ];
#pragma unroll
          for (int j = 0; j < NJ; j++) {
            ro[j] = rowoffT[row * 16 + fw * NJ + j];
#pragma unroll
            for (int k = 0; k < 4; k++) vv[j][k] = stg[(fw * NJ + j) * k3SA + lane + 32 * k];
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&drained[b]);                   // the row is in registers: the buffer may be refilled
          if (!prm.noscatter) {
#pragma unroll
            for (int j = 0; j < NJ; j++) {
              double* dst = (((ro[j] >> 62) & 1) ? prm.ghost_values + (ro[j] & (((int64_t)1 << 62) - 1)) : prm.values + ro[j]) - c0first;
#pragma unroll
              for (int k = 0; k < 4; k++)
                if (vv[j][k] != 0.0 && (k < 3 || lane < 16)) red_add_f64(dst + (A13[j][k] * W0 + fc0[k]), vv[j][k]);
            }
          }
        } else {            // rows whose columns are split between owners along axis 0: position from the packed (B, S, L) bytes
          for (int j = 0; j < NJ; j++) {
            const int a12 = fw * NJ + j;
            const int64_t ro = rowoffT[row * 16 + a12];
            double* dst = ((ro >> 62) & 1) ? prm.ghost_values + (ro & (((int64_t)1 << 62) - 1)) : prm.values + ro;
#pragma unroll
            for (int k = 0; k < 4; k++) {
              const double v = (k == 3 && lane >= 16) ? 0.0 : stg[a12 * k3SA + lane + 32 * k];
              if (v == 0.0 || prm.noscatter) continue;
              const uint32_t ptv = PT[a12 * 16 + fbb[k]];
              const uint32_t s0 = seg0T[row * 8 + fc0[k]];
              const int p23 = (int)(ptv >> 16);
              const int pos = (int)(ptv & 0xFFFF) * W0 + (p23 & 255) * (int)(s0 & 255) + (p23 >> 8) * (int)((s0 >> 8) & 255) + (int)((s0 >> 16) & 255);
              red_add_f64(dst + pos, v);
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&drained[b]);
        }
      }
      asm volatile("bar.sync 2, %0;" ::"n"(k3AsmThreads + 32 * k3rFlushWarps) : "memory");   // the work item's rows are out: its tables may be rewritten
    }
    return;
  }
  if (tid >= k3AsmThreads) {
    // ===================== PRODUCER WARP: cp.async.bulk of (PP0 slice, D') per element into the ring =====================
    if (tid == k3AsmThreads) {
      uint32_t it = 0;
      for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
        const int pen = w / sp.nseg, seg = w - pen * sp.nseg;
        const int l1 = pen % ew1, l2 = pen / ew1;
        const int le0 = seg * sp.seglen, le1 = min(ew0, le0 + sp.seglen);
        for (int le = le0; le < le1; le++, it++) {
          const int slot = it % k3Ring;
          mbar_wait_backoff(&empty[slot], ((it / k3Ring) & 1) ^ 1);
          double* dst = Ring + (size_t)slot * lay.slot;
          mbar_arrive_expect_tx(&full[slot], slot_bytes);
          bulk_g2s(dst, sp.pp[0] + (size_t)(prm.ax[0].es + le) * 576, 576 * 8, &full[slot]);
          if (mapped) {
            const size_t el = (size_t)le + (size_t)ew0 * (l1 + (size_t)ew1 * l2);
            bulk_g2s(dst + 576, sp.dprime + el * ls.npairs * 64, (uint32_t)ls.npairs * 512, &full[slot]);
          }
        }
      }
    }
    return;
  }

  // =========================================== ASSEMBLY WARPS ===========================================
  const int lane = tid & 31, warp = tid >> 5, r = lane >> 2, c = lane & 3;
  const int rs = warp >> 1, par = warp & 1;
  const int q1l = r >> 1, parl = r & 1;
  const int dfrag = c + 4 * (q1l ^ (2 * parl)) + 16 * parl;
  const uint32_t* prog = progS + warp * k3rMaxOps;
  const int nops = (int)prog[k3rMaxOps - 1];
  double acc[4][2][2][2];
#pragma unroll
  for (int k = 0; k < 4; k++)
#pragma unroll
    for (int nt = 0; nt < 2; nt++)
#pragma unroll
      for (int mt = 0; mt < 2; mt++) acc[k][nt][mt][0] = acc[k][nt][mt][1] = 0.0;
  uint32_t it = 0;       // ring position of the next element whose stages A + B have not run yet
  uint32_t nflush = 0;   // flush events before the current work item
  for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
    const int pen = w / sp.nseg, seg = w - pen * sp.nseg;
    const int e1 = prm.ax[1].es + pen % ew1, e2 = prm.ax[2].es + pen / ew1;
    const int le0 = seg * sp.seglen, le1 = min(ew0, le0 + sp.seglen), nel = le1 - le0;
    const int G1 = prm.ax[1].offset[e1] - prm.ax[1].gs, G2 = prm.ax[2].offset[e2] - prm.ax[2].gs;
    const int off_first = prm.ax[0].offset[prm.ax[0].es + le0];
    const int Gf = off_first - prm.ax[0].gs;
    const int nrows = nel + 3;
    bar_asm();
    {
      const double* g1p = sp.pp[1] + (size_t)e1 * 576;
      const double* g2p = sp.pp[2] + (size_t)e2 * 576;
      for (int t = tid; t < 576; t += k3AsmThreads) { PP1[t] = g1p[t]; PP2[t] = g2p[t]; }
      {
        const int a12 = tid >> 4, bb = tid & 15, a1 = a12 & 3, a2 = a12 >> 2, b1 = bb & 3, b2 = bb >> 2;
        const int g1 = G1 + a1, g2 = G2 + a2;
        const uint32_t s1 = prm.ax[1].seg[g1 * kMaxW + b1 - a1 + prm.ax[1].lo[g1]], s2 = prm.ax[2].seg[g2 * kMaxW + b2 - a2 + prm.ax[2].lo[g2]];
        const int Bj = s1 & 255, Sj = (s1 >> 8) & 255, Lj = (s1 >> 16) & 255, Bk = s2 & 255, Sk = (s2 >> 8) & 255, Lk = (s2 >> 16) & 255;
        const int W1 = prm.ax[1].W[g1];
        PT[tid] = (uint32_t)(Bk * W1 + Sk * Bj) | ((uint32_t)(Sk * Sj) << 16) | ((uint32_t)(Lk * Sj + Lj) << 24);
      }
      for (int t = tid; t < nrows * 16; t += k3AsmThreads) {
        const int i0 = t >> 4, a12 = t & 15, a1 = a12 & 3, a2 = a12 >> 2;
        const int gidx = (Gf + i0) + prm.ax[0].gw * ((G1 + a1) + prm.ax[1].gw * (G2 + a2));
        const int lr = prm.localrow[gidx];
        int64_t base = prm.rowbase[lr];
        if (lr >= prm.nown) base = (base - prm.nnz_own) | ((int64_t)1 << 62);
        rowoffT[t] = base;
      }
      for (int t = tid; t < nrows * 8; t += k3AsmThreads) {
        const int i0 = t >> 3, c0 = t & 7, g = Gf + i0;
        const int W = prm.ax[0].W[g], cc = c0 - 3 + prm.ax[0].lo[g];
        if (c0 < 7) seg0T[t] = (cc >= 0 && cc < W) ? prm.ax[0].seg[g * kMaxW + cc] : 0xFFFFFFFFu;
        else seg0T[t] = (uint32_t)prm.ax[0].simple[g];
        if (c0 == 0) W0T[i0] = W;
      }
      for (int t = tid; t < nel * 4; t += k3AsmThreads) {
        const int e0 = prm.ax[0].es + le0 + (t >> 2), q = t & 3;
        wj0T[t] = prm.ax[0].weight[e0 * 4 + q] * prm.ax[0].detJac[e0];
      }
    }
    bool pen_fix = false, lo0_fix = false, hi0_fix = false;
    if (sp.fixsys) {
      if (!prm.ax[1].periodic) pen_fix = pen_fix || (e1 == 0 && prm.bc[1][0].vcount) || (e1 == prm.ax[1].nel - 1 && prm.bc[1][1].vcount);
      if (!prm.ax[2].periodic) pen_fix = pen_fix || (e2 == 0 && prm.bc[2][0].vcount) || (e2 == prm.ax[2].nel - 1 && prm.bc[2][1].vcount);
      if (!prm.ax[0].periodic) { lo0_fix = prm.bc[0][0].vcount > 0; hi0_fix = prm.bc[0][1].vcount > 0; }
    }
    auto elem_fixed = [&](int le) { const int id0 = prm.ax[0].es + le0 + le; return pen_fix || (lo0_fix && id0 == 0) || (hi0_fix && id0 == prm.ax[0].nel - 1); };
    auto build_fix = [&](int le) {     // Dirichlet data of element le (threads 0..63), buffer le & 1
      const int IDs[3] = {prm.ax[0].es + le0 + le, e1, e2};
      const int ai[3] = {tid & 3, (tid >> 2) & 3, tid >> 4};
      const int gidx = (Gf + le + ai[0]) + prm.ax[0].gw * ((G1 + ai[1]) + prm.ax[1].gw * (G2 + ai[2]));
      int onfix; double vfix, vflux;
      sf3_node_bc(prm, IDs, ai, gidx, onfix, vfix, vflux);
      double* fv = fixbuf + (le & 1) * 128;
      int* ff = reinterpret_cast<int*>(fv + 64);
      fv[tid] = vfix; ff[tid] = onfix; ff[64 + tid] = prm.localrow[gidx];
    };
    double wq12[2] = {0, 0};
    if (!mapped) {
      const double w1 = prm.ax[1].weight[e1 * 4 + q1l] * prm.ax[1].detJac[e1];
#pragma unroll
      for (int b = 0; b < 2; b++) wq12[b] = w1 * prm.ax[2].weight[e2 * 4 + 2 * b + parl] * prm.ax[2].detJac[e2];
    }
    if (tid < 64 && elem_fixed(0)) build_fix(0);
    bar_asm();
    // ---- stages A + B of the first element: all eight warps ----
    {
      const int slot = it % k3Ring;
      const double* P0 = Ring + (size_t)slot * lay.slot;
      mbar_wait(&full[slot], (it / k3Ring) & 1, 4);
      const double w0q = mapped ? 0.0 : wj0T[c];
      if constexpr (FS > 0) sf3r_stage_ab_static<FS>(warp, P0, P0 + 576, PP1, ccS, U2, lane, r, c, dfrag, w0q, wq12);
      else sf3r_stage_ab(prog, nops, mapped, P0, P0 + 576, PP1, ccS, U2, lane, r, c, dfrag, w0q, wq12);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);
      it++;
    }
    bar_asm();
    for (int le = 0; le < nel; le++) {
      const int a0 = (rs - le) & 3;                                        // this pair's local row of element le: row le + a0
      const bool efix = elem_fixed(le);
      const bool more = le + 1 < nel;
      if (more && tid < 64 && elem_fixed(le + 1)) build_fix(le + 1);
      const double* U2c = U2 + (le & 1) * k3rU2;
      const double* fv = fixbuf + (le & 1) * 128;
      const int* ff = reinterpret_cast<const int*>(fv + 64);
      // ---- stage C: accumulate onto the running row ----
      {
        const int b00 = (par + a0 + 1) & 1, K0 = (b00 - a0 + 3) >> 1;
        const double* ub0 = U2c + c * k3U2Q + (a0 * 4 + b00) * 24 + r;
        if (!efix) {
          if (K0 == 0) sf3r_stage_c<0, FS>(ls, PP2, ub0, acc, lane);
          else if (K0 == 1) sf3r_stage_c<1, FS>(ls, PP2, ub0, acc, lane);
          else sf3r_stage_c<2, FS>(ls, PP2, ub0, acc, lane);
        } else {
          double T[2][2][2][2];
          sf3r_stage_c_fix(prm, ls, PP2, ub0, T, lane, a0, b00, fv, ff, ff + 64);
          if (K0 == 0) sf3r_add<0>(acc, T);
          else if (K0 == 1) sf3r_add<1>(acc, T);
          else sf3r_add<2>(acc, T);
        }
      }
      if (a0 == 0) {   // row le is complete: stage it for the flush warps (event nflush + le)
        const uint32_t n = nflush + le;
        const int b = n & 1;
        if (n >= 2) mbar_wait(&drained[b], ((n >> 1) & 1) ^ 1, 5);        // the buffer's previous row (event n - 2) has been read
        double* stg = Stage + b * k3rStage;
        if (par == 0) sf3r_stage_row<0>(stg, acc, r, c); else sf3r_stage_row<1>(stg, acc, r, c);
        __syncwarp();
        if (lane == 0) mbar_arrive(&staged[b]);
      }
      // ---- stages A + B of the next element ----
      if (more) {
        const int slot = it % k3Ring;
        if (nops > 0) {
          const double* P0 = Ring + (size_t)slot * lay.slot;
          mbar_wait(&full[slot], (it / k3Ring) & 1, 6);
          const double w0q = mapped ? 0.0 : wj0T[(le + 1) * 4 + c];
          if constexpr (FS > 0) sf3r_stage_ab_static<FS>(warp, P0, P0 + 576, PP1, ccS, U2 + ((le + 1) & 1) * k3rU2, lane, r, c, dfrag, w0q, wq12);
          else sf3r_stage_ab(prog, nops, mapped, P0, P0 + 576, PP1, ccS, U2 + ((le + 1) & 1) * k3rU2, lane, r, c, dfrag, w0q, wq12);
          __syncwarp();
        }
        if (lane == 0) mbar_arrive(&empty[slot]);
        it++;
      }
      bar_asm();
    }
    // ---- the three rows past the last element are partial: flush what this segment contributed ----
    for (int row = nel; row < nel + 3; row++) {
      if ((row & 3) == rs) {
        const uint32_t n = nflush + row;
        const int b = n & 1;
        if (n >= 2) mbar_wait(&drained[b], ((n >> 1) & 1) ^ 1, 7);
        double* stg = Stage + b * k3rStage;
        if (par == 0) sf3r_stage_row<0>(stg, acc, r, c); else sf3r_stage_row<1>(stg, acc, r, c);
        __syncwarp();
        if (lane == 0) mbar_arrive(&staged[b]);
      }
      bar_asm();   // the three tail rows are staged in event order: the parity wait above is only valid one phase away from the barrier
    }
    nflush += nel + 3;
    // the flush warps read this work item's row tables until its last rows are out
    asm volatile("bar.sync 2, %0;" ::"n"(k3AsmThreads + 32 * k3rFlushWarps) : "memory");
  }
}

}  // namespace pc
