// pc_quad3.cuh -- third-generation quadrature kernel: persistent, warp-specialised, FP64 tensor cores, shared-memory scatter.
//
// Same contract as quad_sf_kernel (pc_quad2.cuh) -- the body of the reference's element loop (src/petigaelem.c:375-410,
// 693-1033,1166-1559; src/petigapoint.c:414-492) for 3-D, dof 1, degree 3 with the default 4-point rule, constant-coefficient
// forms on identity or mapped (non-rational) geometry: the headline configuration on the general quadrature path.  What changed,
// each step on a measurement (tools/microbench.cu, profiles/r2_microbench.json):
//
//  * red.global.add.f64 runs at <= 258 G/s on a B200 (1.09 cycles per lane and SM): the 8.6e9 element-matrix entries of
//    BASELINE cfg 2 cost >= 33 ms of LSU time however fast the arithmetic is.  A CTA therefore walks a whole PENCIL of elements
//    along axis 0 and sums the contributions of consecutive elements in a shared-memory window S[row slot][...] with plain
//    conflict-free LDS/DADD/STS (2.26 T/s); a row slice is flushed with coalesced reductions when no later element of the
//    pencil touches it: 7/16 of the reductions, in runs of 28 contiguous doubles.
//  * mma.sync.m8n8k4.f64 (SASS DMMA) sustains 37.1 TFLOP/s against 34.1 for DFMA, at 1/8 of the issue slots: the three
//    sum-factorisation stages are GEMMs  U1 = PP0 x D',  U2 = U1 x PP1,  K = PP2 x U2  with 16x16 tiles at p = 3; stage A's
//    accumulator fragment is stage B's A operand as it stands (columns <-> contraction index), so only U2 crosses shared memory.
//  * the CTA is split into an ASSEMBLY group (8 warps: stages A-C, window, flush) and a GEOMETRY group (4 warps: 1-D tables by
//    cp.async.bulk (SASS UBLKCP) + mbarrier transaction counts, closure, geometry map and its inverse at the points, D', the
//    element vector) that runs one element ahead through a two-slot ring guarded by full/empty mbarriers (SASS SYNCS); the groups
//    use named barriers (bar.sync 1 / 2), never __syncthreads().
#pragma once
#include "pc_quad2.cuh"

namespace pc {

constexpr int k3AsmThreads = 256, k3GeoThreads = 128, k3Threads = k3AsmThreads + k3GeoThreads;
constexpr int k3DS = 96;          // doubles per pair in D': q = q0 + 4 q1 + 24 q2 (q2 stride padded for conflict-free B fragments)
constexpr int k3U2Q = 388;        // q2 stride of U2 (16 rows of 24 + 4: conflict-free B fragments in stage C)
constexpr int k3MaxPairs = 16;
constexpr int k3SRow = 80;        // S inner block [a1 (stride 20)][b1 (4)][b2]
constexpr int k3SSize = 4 * 4 * 7 * k3SRow;
constexpr int k3Meta = 256;       // doubles per ring slot of element metadata

struct SF3Params {
  KParams k;
  SFLists l;
  const double* pp[3];            // [e][os*3+ot][a*4+b][q]  (q fastest: A/B fragments are 32 consecutive doubles)
  double cconst[k3MaxPairs];      // identity geometry: D'[pair] / JW
  double Cc[16];                  // the form's constant coefficient tensor [NA][NA]
  double fconst[4];               // constant vector coefficient (when !per_qp)
  int const_dp;
  int npencils;
  int fixsys;                     // slot == SYSTEM with boundary conditions
};

// plan-level table for this kernel: PP3[e][oo][ab][q]
static __global__ void sf3_pp_kernel(DevAxis ax, double* __restrict__ out) {
  const int per = 9 * 16 * 4;
  const size_t total = (size_t)ax.nel * per;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int e = (int)(t / per), r = (int)(t - (size_t)e * per), oo = r / 64, ab = (r / 4) % 16, q = r % 4, a = ab / 4, b = ab % 4;
    const double* v = ax.value + ((size_t)(e * 4 + q) * 4) * 5;
    out[t] = v[a * 5 + oo / 3] * v[b * 5 + oo % 3];
  }
}

struct SF3Smem {   // offsets in doubles
  int S, U2, PP1, PP2, PP0, D, meta, PT, gB, gX, gT1, gT2, gEv, gGeo, gFp, gFe, bars, total;
  __host__ __device__ SF3Smem() {
    int o = 0;
    S = o; o += k3SSize;
    U2 = o; o += 4 * 4 * k3U2Q;
    PP1 = o; o += 576; PP2 = o; o += 576; PP0 = o; o += 2 * 576;
    D = o; o += 2 * k3MaxPairs * k3DS;
    meta = o; o += 2 * k3Meta;
    PT = o; o += 128;
    gB = o; o += 3 * 32 + 3 * 8;       // Bt[d][o][q][a], wJ[d][q], pt[d][q]
    gX = o; o += 192;
    gT1 = o; o += 384;
    gT2 = o; o += 576;
    gEv = o; o += 12 * 64;
    gGeo = o; o += 64 * 16;
    gFp = o; o += 4 * 64;
    gFe = o; o += 64 * 4;
    bars = o; o += 8;
    total = o;
  }
};

__device__ __forceinline__ uint32_t s3_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s3_u32(b)), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s3_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(s3_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(s3_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(s3_u32(dst)), "l"(src), "r"(bytes), "r"(s3_u32(b)) : "memory");
}
__device__ __forceinline__ void bar_asm() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void bar_geo() { asm volatile("bar.sync 2, 128;" ::: "memory"); }
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// element metadata of one ring slot (doubles unless noted)
//   [0,64)    rowoff[a0*16 + a12]  int64: value offset of the row, bit 62 set for ghost rows
//   [64,96)   rowlr[a] int32 x 64: local row of node a = a0 + 4 a1 + 16 a2
//   [96,112)  seg0[a0*8 + c0] uint32 (0xFFFFFFFF: column outside the row) ; [112,114) W0[a0] int32 x 4
//   [114,146) fixflag[a] int32 x 64 ; [146,210) fixval[a] ; [210] elemfix (int) ; [211] nflush (int)
constexpr int kM_rowoff = 0, kM_rowlr = 64, kM_seg0 = 96, kM_W0 = 112, kM_fixflag = 114, kM_fixval = 146, kM_flags = 210;

__global__ void __launch_bounds__(k3Threads, 1) quad_sf3_kernel(const __grid_constant__ SF3Params sp) {
  const KParams& prm = sp.k;
  const SFLists& ls = sp.l;
  extern __shared__ __align__(128) double sm3[];
  const SF3Smem lay;
  double *S = sm3 + lay.S, *U2 = sm3 + lay.U2, *PP1 = sm3 + lay.PP1, *PP2 = sm3 + lay.PP2, *PP0 = sm3 + lay.PP0, *Dd = sm3 + lay.D;
  double* Meta = sm3 + lay.meta;
  uint32_t* PT = reinterpret_cast<uint32_t*>(sm3 + lay.PT);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm3 + lay.bars);   // full[0], full[1], empty[0], empty[1]
  const int tid = threadIdx.x;
  const int NA = prm.mc1 - prm.mc0, NV = prm.vc1 - prm.vc0, NT = ls.NT;
  const bool mapped = prm.X != nullptr;
  const bool want_mat = NA > 0, want_vec = slot_has_vec(prm.slot);
  const int ew0 = prm.ax[0].ew, ew1 = prm.ax[1].ew;

  if (tid == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1);                    // full: one arrive by the geometry group + PP0 bytes
    mbar_init(&bars[2], k3AsmThreads / 32); mbar_init(&bars[3], k3AsmThreads / 32);   // empty: one arrive per assembly warp
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int t = tid; t < k3SSize; t += k3Threads) S[t] = 0.0;
  __syncthreads();

  if (tid >= k3AsmThreads) {
    // =========================================== GEOMETRY GROUP ===========================================
    const int gt = tid - k3AsmThreads;
    double* gB = sm3 + lay.gB;                    // Bt[d][o][q][a] at gB + d*32 + o*16 + q*4 + a
    double* wJ = gB + 96;                         // [d][q]
    double* pt = wJ + 12;                         // [d][q]
    double *Xs = sm3 + lay.gX, *T1 = sm3 + lay.gT1, *T2 = sm3 + lay.gT2, *Ev = sm3 + lay.gEv, *Gq = sm3 + lay.gGeo, *Fp = sm3 + lay.gFp, *Fe = sm3 + lay.gFe;
    uint32_t it = 0;
    for (int pen = blockIdx.x; pen < sp.npencils; pen += gridDim.x) {
      const int e1 = prm.ax[1].es + pen % ew1, e2 = prm.ax[2].es + pen / ew1;
      const int ID12[3] = {0, e1, e2};
      bar_geo();                                  // previous pencil's users of gB are done
      for (int t = gt; t < 2 * 32; t += k3GeoThreads) {
        const int d = 1 + t / 32, r = t % 32, o = r / 16, q = (r / 4) % 4, a = r % 4;
        gB[d * 32 + r] = prm.ax[d].value[((size_t)(ID12[d] * 4 + q) * 4 + a) * 5 + o];
      }
      if (gt < 8) {
        const int d = 1 + gt / 4, q = gt % 4;
        wJ[d * 4 + q] = prm.ax[d].weight[ID12[d] * 4 + q] * prm.ax[d].detJac[ID12[d]];
        pt[d * 4 + q] = prm.ax[d].point[ID12[d] * 4 + q];
      }
      const int G1 = prm.ax[1].offset[e1] - prm.ax[1].gs, G2 = prm.ax[2].offset[e2] - prm.ax[2].gs;
      for (int le = 0; le < ew0; le++, it++) {
        const int e0 = prm.ax[0].es + le, buf = it & 1;
        const int G0 = prm.ax[0].offset[e0] - prm.ax[0].gs;
        double* D = Dd + buf * k3MaxPairs * k3DS;
        double* M = Meta + buf * k3Meta;
        mbar_wait(&bars[2 + buf], ((it >> 1) & 1) ^ 1);             // ring slot free (assembly finished element it-2)
        if (gt == 0 && want_mat) {                                     // 1-D pair-product slice of this element, asynchronously
          mbar_expect_tx(&bars[buf], 576 * 8);
          bulk_g2s(PP0 + buf * 576, sp.pp[0] + (size_t)e0 * 576, 576 * 8, &bars[buf]);
        }
        // ---- axis-0 tables, closure, metadata ----
        if (gt < 32) { const int o = gt / 16, q = (gt / 4) % 4, a = gt % 4; gB[gt] = prm.ax[0].value[((size_t)(e0 * 4 + q) * 4 + a) * 5 + o]; }
        if (gt >= 32 && gt < 36) { const int q = gt - 32; wJ[q] = prm.ax[0].weight[e0 * 4 + q] * prm.ax[0].detJac[e0]; pt[q] = prm.ax[0].point[e0 * 4 + q]; }
        bool elem_fix = false, elem_bc = false;
        if (prm.any_bc) {
          const int IDs[3] = {e0, e1, e2};
#pragma unroll
          for (int d = 0; d < 3; d++)
            if (!prm.ax[d].periodic) {
              const bool lo = IDs[d] == 0, hi = IDs[d] == prm.ax[d].nel - 1;
              elem_fix = elem_fix || (lo && prm.bc[d][0].vcount > 0) || (hi && prm.bc[d][1].vcount > 0);
              elem_bc = elem_bc || (lo && (prm.bc[d][0].vcount || prm.bc[d][0].lcount)) || (hi && (prm.bc[d][1].vcount || prm.bc[d][1].lcount));
            }
        }
        if (gt < 64) {
          const int a = gt, a0 = a & 3, a1 = (a >> 2) & 3, a2 = a >> 4;
          const int gidx = (G0 + a0) + prm.ax[0].gw * ((G1 + a1) + prm.ax[1].gw * (G2 + a2));
          const int lr = prm.localrow[gidx];
          int64_t base = want_mat ? prm.rowbase[lr] : 0;
          if (lr >= prm.nown) base = (base - prm.nnz_own) | ((int64_t)1 << 62);
          reinterpret_cast<int64_t*>(M + kM_rowoff)[a0 * 16 + a1 + 4 * a2] = base;
          reinterpret_cast<int*>(M + kM_rowlr)[a] = lr;
          if (mapped) {
#pragma unroll
            for (int i = 0; i < 3; i++) Xs[i * 64 + a] = prm.X[(size_t)gidx * 3 + i];
          }
          int onfix = 0; double vfix = 0.0, vflux = 0.0;
          if (elem_bc) {   // BuildFix / AddFixa / AddFlux (petigaelem.c:1166-1283); dof = 1
            const int ai[3] = {a0, a1, a2}, IDs[3] = {e0, e1, e2};
#pragma unroll
            for (int d = 0; d < 3; d++) {
              if (prm.ax[d].periodic) continue;
              for (int s = 0; s < 2; s++) {
                const FixSide& fs = prm.bc[d][s];
                if (!(fs.vcount || fs.lcount)) continue;
                if (IDs[d] != (s ? prm.ax[d].nel - 1 : 0)) continue;
                if (ai[d] != (s ? 3 : 0)) continue;
                for (int k = 0; k < fs.vcount; k++) if (fs.vfield[k] == 0) { onfix = 1; vfix = prm.fixtable ? prm.fixtable[gidx] : fs.vvalue[k]; }
                if (fs.lcount) {
                  double A = 1.0;
                  for (int e = 0; e < 3; e++) if (e != d) A *= prm.ax[e].detJac[IDs[e]] / 4.0;
                  if (prm.face_dS[d][s]) {
                    const int f0 = (d == 0) ? 1 : 0, f1 = (d == 2) ? 1 : 2;
                    A *= prm.face_dS[d][s][(IDs[f0] - prm.ax[f0].es) + prm.ax[f0].ew * (IDs[f1] - prm.ax[f1].es)];
                  } else A *= 4.0;
                  for (int k = 0; k < fs.lcount; k++) if (fs.lfield[k] == 0) vflux += fs.lvalue[k] * A;
                }
              }
            }
          }
          reinterpret_cast<int*>(M + kM_fixflag)[a] = onfix;
          M[kM_fixval + a] = vfix;
          Fe[64 + a] = vflux;                                         // flux, private to this group
        } else if (gt < 64 + 28) {                                    // axis-0 position bytes of the four row slots
          const int t = gt - 64, a0 = t / 7, c0 = t - a0 * 7;
          const int g = G0 + a0, W = prm.ax[0].W[g], cc = c0 - 3 + prm.ax[0].lo[g];
          reinterpret_cast<uint32_t*>(M + kM_seg0)[a0 * 8 + c0] = (cc >= 0 && cc < W) ? prm.ax[0].seg[g * kMaxW + cc] : 0xFFFFFFFFu;
          if (c0 == 0) reinterpret_cast<int*>(M + kM_W0)[a0] = W;
        } else if (gt == 96) {
          int* fl = reinterpret_cast<int*>(M + kM_flags);
          fl[0] = (elem_fix && sp.fixsys) ? 1 : 0;
          fl[1] = (le == ew0 - 1) ? 4 : min(4, prm.ax[0].offset[e0 + 1] - prm.ax[0].offset[e0]);   // row slots complete after this element
          fl[2] = G0;
        }
        bar_geo();
        // ---- geometry at the points: X1 = dX/du by sum factorisation (K5), inverse map (K6) ----
        if (mapped) {
          for (int t = gt; t < 384; t += k3GeoThreads) {              // T1[i][o0][q0][a12]
            const int i = t / 128, r = t % 128, o0 = r / 64, q0 = (r / 16) % 4, a12 = r % 16;
            const double* b = gB + o0 * 16 + q0 * 4;
            const double* x = Xs + i * 64 + a12 * 4;
            T1[t] = b[0] * x[0] + b[1] * x[1] + b[2] * x[2] + b[3] * x[3];
          }
          bar_geo();
          for (int t = gt; t < 576; t += k3GeoThreads) {              // T2[i][oc][q0][q1][a2], oc: 0 = (1,0), 1 = (0,1), 2 = (0,0)
            const int i = t / 192, r = t % 192, oc = r / 64, q0 = (r / 16) % 4, q1 = (r / 4) % 4, a2 = r % 4;
            const int o0 = (oc == 0), o1 = (oc == 1);
            const double* b = gB + 32 + o1 * 16 + q1 * 4;
            const double* s = T1 + i * 128 + o0 * 64 + q0 * 16 + a2 * 4;
            T2[t] = b[0] * s[0] + b[1] * s[1] + b[2] * s[2] + b[3] * s[3];
          }
          bar_geo();
          for (int t = gt; t < 768; t += k3GeoThreads) {              // Ev[i][d][q], d = 3: the point itself
            const int i = t / 256, r = t % 256, d = r / 64, q = r % 64, q0 = q & 3, q1 = (q >> 2) & 3, q2 = q >> 4;
            const int oc = (d == 0) ? 0 : (d == 1 ? 1 : 2), o2 = (d == 2);
            const double* b = gB + 64 + o2 * 16 + q2 * 4;
            const double* s = T2 + i * 192 + oc * 64 + q0 * 16 + q1 * 4;
            Ev[t] = b[0] * s[0] + b[1] * s[1] + b[2] * s[2] + b[3] * s[3];
          }
          bar_geo();
        }
        if (gt < 64) {
          const int q = gt, q0 = q & 3, q1 = (q >> 2) & 3, q2 = q >> 4;
          double E[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, x[3] = {pt[q0], pt[4 + q1], pt[8 + q2]};
          double jw = wJ[q0] * wJ[4 + q1] * wJ[8 + q2];                // W = iW jW kW, J = iJ jJ kJ (petiga3d.F90:22-28)
          if (mapped) {
            double X1[3][3];
#pragma unroll
            for (int i = 0; i < 3; i++) {
#pragma unroll
              for (int d = 0; d < 3; d++) X1[i][d] = Ev[i * 256 + d * 64 + q];
              x[i] = Ev[i * 256 + 192 + q];
            }
            const double a00 = X1[0][0], a01 = X1[0][1], a02 = X1[0][2], a10 = X1[1][0], a11 = X1[1][1], a12 = X1[1][2], a20 = X1[2][0], a21 = X1[2][1], a22 = X1[2][2];
            const double det = a00 * (a11 * a22 - a12 * a21) - a01 * (a10 * a22 - a12 * a20) + a02 * (a10 * a21 - a11 * a20);
            E[0][0] = (a11 * a22 - a12 * a21) / det; E[0][1] = -(a01 * a22 - a02 * a21) / det; E[0][2] = (a01 * a12 - a02 * a11) / det;
            E[1][0] = -(a10 * a22 - a12 * a20) / det; E[1][1] = (a00 * a22 - a02 * a20) / det; E[1][2] = -(a00 * a12 - a02 * a10) / det;
            E[2][0] = (a10 * a21 - a11 * a20) / det; E[2][1] = -(a00 * a21 - a01 * a20) / det; E[2][2] = (a00 * a11 - a01 * a10) / det;
            jw *= det;                                                 // detJac *= detX (petigaelem.c:1024-1029)
          }
          double* g = Gq + q * 16;
#pragma unroll
          for (int d = 0; d < 3; d++)
#pragma unroll
            for (int i = 0; i < 3; i++) g[d * 3 + i] = E[d][i];
          g[9] = jw; g[10] = x[0]; g[11] = x[1]; g[12] = x[2];
          if (want_vec && NV > 0) {                                    // the form's vector coefficient at the point
            double fv[4] = {sp.fconst[0], sp.fconst[1], sp.fconst[2], sp.fconst[3]};
            if (prm.per_qp) {
              QPoint qp;
              qp.atboundary = 0;
              qp.x[0] = x[0]; qp.x[1] = x[1]; qp.x[2] = x[2];
              fv[0] = fv[1] = fv[2] = fv[3] = 0.0;
              form_coefficients<3, 1>(prm.form, prm.slot, prm.prm, prm.shift, prm.t, qp, 0, NV, nullptr, fv);
            }
            // f'[s][q] = JW sum_al A[vc0+al][s] f[al]:  A[0][tN] = 1, A[1+i][tG_d] = E[d][i]
            for (int s = 0; s < NT; s++) {
              double acc = 0.0;
              for (int al = 0; al < NV; al++) {
                const int cph = prm.vc0 + al;
                double av = 0.0;
                if (cph == 0) av = (s == ls.tN) ? 1.0 : 0.0;
                else {
#pragma unroll
                  for (int d = 0; d < 3; d++) if (s == ls.tG[d]) av = E[d][cph - 1];
                }
                acc += av * fv[al];
              }
              Fp[s * 64 + q] = acc * jw;
            }
          }
        }
        bar_geo();
        // ---- D'[pair][q] = JW_q sum_{al,be} A[al][s] C[al][be] A[be][t] ----
        if (want_mat)
          for (int t = gt; t < ls.npairs * 64; t += k3GeoThreads) {
            const int pr = t >> 6, q = t & 63;
            const double* g = Gq + q * 16;
            double v;
            if (sp.const_dp) v = sp.cconst[pr] * g[9];
            else {
              const int s = ls.pair_s[pr], tt = ls.pair_t[pr];
              double acc = 0.0;
              for (int al = 0; al < NA; al++) {
                const int ca = prm.mc0 + al;
                double as = 0.0;
                if (ca == 0) as = (s == ls.tN) ? 1.0 : 0.0;
                else {
#pragma unroll
                  for (int d = 0; d < 3; d++) if (s == ls.tG[d]) as = g[d * 3 + ca - 1];
                }
                if (as == 0.0) continue;
                double inner = 0.0;
                for (int be = 0; be < NA; be++) {
                  const int cb = prm.mc0 + be;
                  double at = 0.0;
                  if (cb == 0) at = (tt == ls.tN) ? 1.0 : 0.0;
                  else {
#pragma unroll
                    for (int d = 0; d < 3; d++) if (tt == ls.tG[d]) at = g[d * 3 + cb - 1];
                  }
                  inner += sp.Cc[al * NA + be] * at;
                }
                acc += as * inner;
              }
              v = acc * g[9];
            }
            D[pr * k3DS + (q & 15) + 24 * (q >> 4)] = v;
          }
        // ---- element vector by the transposed sum factorisation, fix-up, scatter ----
        if (want_vec) {
          if (NV > 0) {
            double* R1 = T1;                                           // [s][q2][q1][a0]
            double* R2 = T2;                                           // [s][q2][a1][a0]
            for (int t = gt; t < NT * 64; t += k3GeoThreads) {
              const int s = t >> 6, r = t & 63, q12 = r >> 2, a0 = r & 3, o = ls.torder[s][0];
              const double* f = Fp + s * 64 + q12 * 4;
              const double* b = gB + o * 16 + a0;
              R1[t] = b[0] * f[0] + b[4] * f[1] + b[8] * f[2] + b[12] * f[3];
            }
            bar_geo();
            for (int t = gt; t < NT * 64; t += k3GeoThreads) {
              const int s = t >> 6, r = t & 63, q2 = r >> 4, a1 = (r >> 2) & 3, a0 = r & 3, o = ls.torder[s][1];
              const double* b = gB + 32 + o * 16 + a1;
              const double* x = R1 + s * 64 + q2 * 16 + a0;
              R2[t] = b[0] * x[0] + b[4] * x[4] + b[8] * x[8] + b[12] * x[12];
            }
            bar_geo();
          }
          if (gt < 64) {
            const int a = gt, a2 = a >> 4, a01 = a & 15;
            double F = 0.0;
            if (NV > 0)
              for (int s = 0; s < NT; s++) {
                const int o = ls.torder[s][2];
                const double* b = gB + 64 + o * 16 + a2;
                const double* x = T2 + s * 64 + a01;
                F += b[0] * x[0] + b[4] * x[16] + b[8] * x[32] + b[12] * x[48];
              }
            if (prm.slot == PETIGA_SLOT_SYSTEM && elem_bc) {          // FixSystem vector part (petigaelem.c:1365-1387)
              F += Fe[64 + a];
              if (reinterpret_cast<int*>(M + kM_fixflag)[a]) F = M[kM_fixval + a];
            }
            if (F != 0.0) atomicAdd(&prm.rhs[reinterpret_cast<int*>(M + kM_rowlr)[a]], F);
          }
        }
        bar_geo();                                                     // every write of this ring slot is done
        if (gt == 0) mbar_arrive(&bars[buf]);
      }
    }
    return;
  }

  // =========================================== ASSEMBLY GROUP ===========================================
  const int lane = tid & 31, warp = tid >> 5, r = lane >> 2, c = lane & 3;
  // flush: a lane owns entries e = lane + 32 k of a row's 112 (c0 fastest, then b1, b2)
  int fc0[4], fbb[4];
#pragma unroll
  for (int k = 0; k < 4; k++) { const int e = lane + 32 * k; fc0[k] = e % 7; fbb[k] = e / 7; }
  uint32_t it = 0;
  for (int pen = blockIdx.x; pen < sp.npencils; pen += gridDim.x) {
    const int e1 = prm.ax[1].es + pen % ew1, e2 = prm.ax[2].es + pen / ew1;
    const int G1 = prm.ax[1].offset[e1] - prm.ax[1].gs, G2 = prm.ax[2].offset[e2] - prm.ax[2].gs;
    bar_asm();                                                         // previous pencil's readers of PP1/PP2/PT are done
    if (want_mat) {
      const double* g1p = sp.pp[1] + (size_t)e1 * 576;
      const double* g2p = sp.pp[2] + (size_t)e2 * 576;
      for (int t = tid; t < 576; t += k3AsmThreads) { PP1[t] = g1p[t]; PP2[t] = g2p[t]; }
      {  // PT[(a1,a2)][(b1,b2)]: pos = P1*W0 + P2*Bi + P3*Si + Li
        const int a12 = tid >> 4, bb = tid & 15, a1 = a12 & 3, a2 = a12 >> 2, b1 = bb & 3, b2 = bb >> 2;
        const int g1 = G1 + a1, g2 = G2 + a2;
        const uint32_t s1 = prm.ax[1].seg[g1 * kMaxW + b1 - a1 + prm.ax[1].lo[g1]], s2 = prm.ax[2].seg[g2 * kMaxW + b2 - a2 + prm.ax[2].lo[g2]];
        const int Bj = s1 & 255, Sj = (s1 >> 8) & 255, Lj = (s1 >> 16) & 255, Bk = s2 & 255, Sk = (s2 >> 8) & 255, Lk = (s2 >> 16) & 255;
        const int W1 = prm.ax[1].W[g1];
        PT[tid] = (uint32_t)(Bk * W1 + Sk * Bj) | ((uint32_t)(Sk * Sj) << 16) | ((uint32_t)(Lk * Sj + Lj) << 24);
      }
    }
    bar_asm();
    for (int le = 0; le < ew0; le++, it++) {
      const int buf = it & 1;
      const double* D = Dd + buf * k3MaxPairs * k3DS;
      const double* P0 = PP0 + buf * 576;
      double* M = Meta + buf * k3Meta;
      mbar_wait(&bars[buf], (it >> 1) & 1);                            // D', metadata and the PP0 bytes have landed
      const int* fl = reinterpret_cast<const int*>(M + kM_flags);
      const int elem_fix = fl[0], nflush = fl[1], G0 = fl[2];
      if (want_mat) {
        // ---- stages A + B: a warp owns the combos (g2, q2) = combo, combo + 8, ...; stage A's accumulators feed stage B ----
        for (int combo = warp; combo < ls.ng2 * 4; combo += 8) {
          const int g2 = combo >> 2, q2 = combo & 3, b = q2 >> 1, x = q2 & 1;
          double cB[2][2][2] = {{{0, 0}, {0, 0}}, {{0, 0}, {0, 0}}};   // [mt][nt][reg]
          for (int g1 = ls.g2_first[g2]; g1 < ls.g2_first[g2 + 1]; g1++) {
            double cA[2][2] = {{0, 0}, {0, 0}};                        // [mt][reg]: U1[ab0 = r + 8 mt][q1 = c][q2 = 2b + reg]
            for (int pr = ls.g1_first[g1]; pr < ls.g1_first[g1 + 1]; pr++) {
              const double bf = D[pr * k3DS + c + 4 * (r >> 1) + 24 * (2 * b + (r & 1))];
              const double* pa = P0 + ls.pair_oo0[pr] * 64 + lane;
              dmma(cA[0][0], cA[0][1], pa[0], bf);
              dmma(cA[1][0], cA[1][1], pa[32], bf);
            }
            const double* pb = PP1 + ls.g1_oo1[g1] * 64 + lane;
            const double b0 = pb[0], b1 = pb[32];
            dmma(cB[0][0][0], cB[0][0][1], cA[0][x], b0);
            dmma(cB[0][1][0], cB[0][1][1], cA[0][x], b1);
            dmma(cB[1][0][0], cB[1][0][1], cA[1][x], b0);
            dmma(cB[1][1][0], cB[1][1][1], cA[1][x], b1);
          }
          double* u = U2 + (g2 * 4 + q2) * k3U2Q + r * 24 + 2 * c;     // U2[g2][q2][ab0 * 24 + ab1]
#pragma unroll
          for (int mt = 0; mt < 2; mt++)
#pragma unroll
            for (int nt = 0; nt < 2; nt++)
              *reinterpret_cast<double2*>(u + mt * 8 * 24 + nt * 8) = make_double2(cB[mt][nt][0], cB[mt][nt][1]);
        }
      }
      bar_asm();
      if (want_mat) {
        // ---- stage C: K[ab2][ab1][ab0] for ab0 in {2 warp, 2 warp + 1}; k index = q2 ----
        double cC[2][2][2][2];                                         // [h][nt][mt][reg]
#pragma unroll
        for (int h = 0; h < 2; h++)
#pragma unroll
          for (int nt = 0; nt < 2; nt++)
#pragma unroll
            for (int mt = 0; mt < 2; mt++) cC[h][nt][mt][0] = cC[h][nt][mt][1] = 0.0;
        for (int g2 = 0; g2 < ls.ng2; g2++) {
          const double* pa = PP2 + ls.g2_oo2[g2] * 64 + lane;
          const double a0f = pa[0], a1f = pa[32];
          const double* ub = U2 + (g2 * 4 + c) * k3U2Q + (2 * warp) * 24 + r;
#pragma unroll
          for (int h = 0; h < 2; h++)
#pragma unroll
            for (int nt = 0; nt < 2; nt++) {
              const double bf = ub[h * 24 + nt * 8];
              dmma(cC[h][nt][0][0], cC[h][nt][0][1], a0f, bf);
              dmma(cC[h][nt][1][0], cC[h][nt][1][1], a1f, bf);
            }
        }
        // ---- Dirichlet fix-up on the fragments (petigaelem.c:1360-1389), then the shared-memory window ----
        const int a0 = warp >> 1;
        const int sl = (G0 + a0) & 3;
        const int* fixflag = reinterpret_cast<const int*>(M + kM_fixflag);
        const int* rowlr = reinterpret_cast<const int*>(M + kM_rowlr);
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int b0 = 2 * (warp & 1) + h, c0 = b0 - a0 + 3;
#pragma unroll
          for (int nt = 0; nt < 2; nt++)
#pragma unroll
            for (int mt = 0; mt < 2; mt++) {
              const int a2 = 2 * mt + (r >> 2), b2 = r & 3, a1 = 2 * nt + (c >> 1);
              double* sp0 = S + ((sl * 4 + a2) * 7 + c0) * k3SRow + a1 * 20 + 8 * (c & 1) + b2;   // b1 = 2 (c & 1) + reg
#pragma unroll
              for (int x = 0; x < 2; x++) {
                double v = cC[h][nt][mt][x];
                if (elem_fix) {
                  const int b1 = 2 * (c & 1) + x, ra = a0 + 4 * a1 + 16 * a2, cb = b0 + 4 * b1 + 16 * b2;
                  const bool fr = fixflag[ra], fcx = fixflag[cb];
                  if (fr || fcx) {
                    if (fcx && !fr) atomicAdd(&prm.rhs[rowlr[ra]], -v * M[kM_fixval + cb]);
                    v = (ra == cb) ? 1.0 : 0.0;
                  }
                }
                sp0[4 * x] += v;
              }
            }
        }
      }
      bar_asm();
      // ---- flush the row slots no later element of the pencil touches: coalesced red.global.add.f64 ----
      if (want_mat) {
        const int64_t* rowoff = reinterpret_cast<const int64_t*>(M + kM_rowoff);
        const uint32_t* seg0 = reinterpret_cast<const uint32_t*>(M + kM_seg0);
        const int* W0s = reinterpret_cast<const int*>(M + kM_W0);
        for (int f = 0; f < nflush; f++) {
          const int sl = (G0 + f) & 3, W0 = W0s[f];
#pragma unroll
          for (int rr = 0; rr < 2; rr++) {
            const int a12 = 2 * warp + rr, a1 = a12 & 3, a2 = a12 >> 2;
            const int64_t ro = rowoff[f * 16 + a12];
            double* dst = ((ro >> 62) & 1) ? prm.ghost_values + (ro & (((int64_t)1 << 62) - 1)) : prm.values + ro;
#pragma unroll
            for (int k = 0; k < 4; k++) {
              if (k == 3 && lane >= 16) break;
              const int c0 = fc0[k], bb = fbb[k];
              const uint32_t s0 = seg0[f * 8 + c0];
              double* sp0 = S + ((sl * 4 + a2) * 7 + c0) * k3SRow + a1 * 20 + (bb & 3) * 4 + (bb >> 2);
              const double v = *sp0;
              if (v != 0.0) {
                const uint32_t ptv = PT[a12 * 16 + bb];
                const int pos = (int)(ptv & 0xFFFF) * W0 + (int)((ptv >> 16) & 255) * (int)(s0 & 255) + (int)(ptv >> 24) * (int)((s0 >> 8) & 255) + (int)((s0 >> 16) & 255);
                atomicAdd(dst + pos, v);
                *sp0 = 0.0;
              }
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[2 + buf]);                       // this warp is done with the ring slot
    }
  }
}

}  // namespace pc
