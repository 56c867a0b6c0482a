/*
 * petiga_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY).
 *
 * A plain-C restatement of the PetIGA element-assembly hot path, written from the reference's
 * algorithm (every function cites the reference file:line it follows).  It exists so that the
 * CUDA product (libpetiga_cuda) can be checked against "what the reference computes" on a box
 * that has no PETSc/MPI/gfortran.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product never does.
 *
 * Pinning: the reference cannot be compiled here (needs PETSc + MPI + a Fortran compiler), so the
 * oracle is pinned by the reference's own known answers (tests/test_oracle_known_answers.py):
 * tutorial pattern sizes (docs/manual/TUTORIAL.rst:113-115,204-205), partition report (:78-80),
 * partition-of-unity mass solve (test/IGACreate.c:105-149), quarter-annulus closed forms
 * (test/IGAGeometryMap.c:18-257), polynomial norms / exact L2 projection (test/IGAErrNorm.c),
 * Gauss literals bit-identical (src/petigarule.c:182-319).  The PETSc value-array level
 * (MPIAIJ diag/off-diag split, BAIJ in-block order) is PARITY UNPINNED: no reference test holds it.
 *
 * Conventions: FP64, 32-bit ints.  Everything is done for 3 axes; unused axes are the
 * reference's "reset" axis (p=0, U=[-1/2,1/2], 1 element, 1-point rule: src/petigaaxis.c:44-66,
 * src/petiga.c:1125-1126,1478-1482).  Multi-rank runs are emulated by looping over ranks and
 * adding into one global matrix in PETSc global numbering -- what MatAssemblyEnd produces.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include <float.h>

#include "gauss_tables.h"

#define MAXP 8
#define MAXBC 64

/* ------------------------------------------------------------------------------------------ */
/* types                                                                                      */
/* ------------------------------------------------------------------------------------------ */

typedef struct { /* include/petiga.h:50-60 */
  int p, m; double *U; int periodic; int nnp, nel; int *span;
} Axis;

typedef struct { /* include/petiga.h:122-141 */
  int nel, nqp, nen;
  int *offset; double *detJac, *weight, *point, *value; /* value[nel][nqp][nen][5] */
  double bnd_value[2][(MAXP+1)*5], bnd_point[2];        /* basis at the two ends of the axis (petigabasis.c:208-217) */
} Basis;

typedef struct { int count; int field[MAXBC]; double value[MAXBC]; } FormBC; /* petiga.h:221-225 */

typedef struct {
  int dim, dof, order;
  Axis axis[3];
  int rule_nqp[3];
  Basis basis[3];
  FormBC value[3][2], load[3][2];
  int visit[3][2];                                   /* IGASetBoundaryForm (petigaform.c): boundary-integral pass per face */
  /* geometry in natural (geom) ordering, i fastest: src/petigaio.c:201-286 */
  int geometry /* = nsd or 0 */, rational;
  double *geomX_nat, *geomW_nat;
  /* fix table: global vector in PETSc ordering (IGASetFixTable, src/petigaform.c) or NULL */
  int fixtable; double *fixtable_glob;
  /* partition state (src/petiga.c:1111-1310) */
  int size, rank;
  int proc_sizes[3], proc_ranks[3];
  int elem_sizes[3], elem_start[3], elem_width[3];
  int geom_sizes[3], geom_lstart[3], geom_lwidth[3], geom_gstart[3], geom_gwidth[3];
  int node_sizes[3], node_lstart[3], node_lwidth[3], node_gstart[3], node_gwidth[3];
  /* per-rank ghost arrays (rebuilt by setup_rank) */
  int *own[3], *box_ls[3], *box_lw[3], *rstart;   /* per-axis owner tables (derived from node_box_1d) */
  int *lgmap;              /* ghost node -> global (PETSc) node */
  double *geometryX, *rationalW, *fixtableU;
  int tables_ready;
  /* third state vector / second shift / second time of the IE and I2 drivers (set by oiga_set_aux before oiga_assemble) */
  double aux_shift2, aux_t0; const double *aux_W;
} OIGA;

enum { SLOT_VECTOR=0, SLOT_MATRIX, SLOT_SYSTEM, SLOT_FUNCTION, SLOT_JACOBIAN, SLOT_IFUNCTION, SLOT_IJACOBIAN,
       SLOT_IEFUNCTION, SLOT_IEJACOBIAN, SLOT_RHSFUNCTION, SLOT_RHSJACOBIAN, SLOT_I2FUNCTION, SLOT_I2JACOBIAN };   /* src/petigats.c:182-477, src/petigats2.c:23-175 */
enum { FORM_POISSON=0, FORM_LAPLACE, FORM_L2PROJECTION, FORM_ELASTICITY3D, FORM_ELASTICITY,
       FORM_CAHNHILLIARD2D, FORM_BRATU, FORM_MASS, FORM_BOUNDARYINTEGRAL, FORM_NEUMANN, FORM_CAHNHILLIARD3D, FORM_CONVTEST,
       FORM_SNES2D, FORM_PATTERNFORMATION, FORM_ELASTICROD, FORM_NITSCHE };

/* ------------------------------------------------------------------------------------------ */
/* axis: src/petigaaxis.c                                                                     */
/* ------------------------------------------------------------------------------------------ */

static int next_knot(int m, const double U[], int k, int direction) /* petigaaxis.c:482-494 */
{
  int j;
  if (direction >= 0) {
    if (k < 0) return 0;
    for (j = k+1; j < m; j++) if (U[j] > U[k]) return j;
    return m;
  } else {
    if (k > m) return m;
    for (j = k-1; j > 0; j--) if (U[j] < U[k]) return j;
    return 0;
  }
}

static void axis_reset(Axis *ax) /* petigaaxis.c:44-66 */
{
  free(ax->U); free(ax->span);
  ax->periodic = 0; ax->p = 0; ax->m = 1;
  ax->U = (double*)malloc(2*sizeof(double)); ax->U[0] = -0.5; ax->U[1] = +0.5;
  ax->span = (int*)malloc(sizeof(int)); ax->nnp = 1; ax->nel = 1; ax->span[0] = 0;
}

static int axis_init_uniform(Axis *ax, int p, int N, double Ui, double Uf, int C, int periodic)
{ /* petigaaxis.c:401-454 */
  int i, j, k, s, n, m, r; double *U;
  if (C < 0 && C != -1) return 1;
  if (C == -1) C = p-1;
  if (p < 1 || N < 1 || !(Ui < Uf) || C < 0 || C >= p) return 1;
  s = p - C; r = N; m = 2*(p+1) + (N-1)*s - 1; n = m - p - 1;
  free(ax->U); free(ax->span);
  ax->p = p; ax->m = m; ax->periodic = periodic;
  U = ax->U = (double*)malloc((size_t)(m+1)*sizeof(double));
  for (k = 0; k <= p; k++) { U[k] = Ui; U[m-k] = Uf; }
  for (i = 1; i <= r-1; i++)
    for (j = 1; j <= s; j++)
      U[k++] = Ui + (double)i/(double)N * (Uf-Ui);
  if (periodic)
    for (k = 0; k <= C; k++) {
      U[C-k]   = U[p] - U[m-p] + U[n-k];
      U[m-C+k] = U[m-p] - U[p] + U[p+1+k];
    }
  ax->nel = r;
  ax->span = (int*)malloc((size_t)r*sizeof(int));
  for (i = 0; i < r; i++) ax->span[i] = p + i*s;
  ax->nnp = periodic ? n-C : n+1;
  return 0;
}

static int axis_set_knots(Axis *ax, int p, int m, const double *Uin, int periodic)
{ /* IGAAxisSetKnots + IGAAxisSetUp: petigaaxis.c:286-310,456-480 */
  int n = m - p - 1, k, count;
  if (p < 1 || m < 2*p+1) return 1;
  free(ax->U); free(ax->span);
  ax->p = p; ax->m = m; ax->periodic = periodic;
  ax->U = (double*)malloc((size_t)(m+1)*sizeof(double));
  memcpy(ax->U, Uin, (size_t)(m+1)*sizeof(double));
  k = p; count = 0;
  while ((k = next_knot(m, ax->U, k, 1)) <= n+1) count++;
  ax->span = (int*)malloc((size_t)count*sizeof(int));
  k = p; count = 0;
  while ((k = next_knot(m, ax->U, k, 1)) <= n+1) ax->span[count++] = k-1;
  ax->nel = count;
  if (periodic) {
    int kk = n+1, j = next_knot(m, ax->U, kk, 1), s = j-kk, C = p-s;
    ax->nnp = n-C;
  } else ax->nnp = n+1;
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* 1-D B-spline basis: src/petigabsb.f90.in:3-63 (Piegl-Tiller A2.3), src/petigabsp.F90:3-15  */
/* ------------------------------------------------------------------------------------------ */

static void bspline_ders(int i, double uu, int p, int n, const double *U, double ders[MAXP+1][5])
{ /* ders[r][k], r = 0..p (function), k = 0..n (derivative); entries k>n zeroed (petigabsp.F90:13) */
  int j, k, r, s1, s2, rk, pk, j1, j2;
  double saved, temp, d;
  double left[MAXP+1], right[MAXP+1], ndu[MAXP+1][MAXP+1], a[2][MAXP+1];
  ndu[0][0] = 1;
  for (j = 1; j <= p; j++) {
    left[j]  = uu - U[i+1-j];
    right[j] = U[i+j] - uu;
    saved = 0;
    for (r = 0; r <= j-1; r++) {
      ndu[j][r] = right[r+1] + left[j-r];
      temp = ndu[r][j-1] / ndu[j][r];
      ndu[r][j] = saved + right[r+1] * temp;
      saved = left[j-r] * temp;
    }
    ndu[j][j] = saved;
  }
  for (r = 0; r <= p; r++) for (k = 0; k < 5; k++) ders[r][k] = 0;
  for (r = 0; r <= p; r++) ders[r][0] = ndu[r][p];
  for (r = 0; r <= p; r++) {
    s1 = 0; s2 = 1;
    a[0][0] = 1;
    for (k = 1; k <= n; k++) {
      d = 0; rk = r-k; pk = p-k;
      if (r >= k) { a[s2][0] = a[s1][0] / ndu[pk+1][rk]; d = a[s2][0] * ndu[rk][pk]; }
      j1 = (rk > -1) ? 1 : -rk;
      j2 = (r-1 <= pk) ? k-1 : p-r;
      for (j = j1; j <= j2; j++) {
        a[s2][j] = (a[s1][j] - a[s1][j-1]) / ndu[pk+1][rk+j];
        d = d + a[s2][j] * ndu[rk+j][pk];
      }
      if (r <= pk) { a[s2][k] = - a[s1][k-1] / ndu[pk+1][r]; d = d + a[s2][k] * ndu[r][pk]; }
      ders[r][k] = d;
      j = s1; s1 = s2; s2 = j;
    }
  }
  r = p;
  for (k = 1; k <= n; k++) {
    for (j = 0; j <= p; j++) ders[j][k] = ders[j][k] * (double)r;
    r = r * (p-k);
  }
}

static void basis_free(Basis *b)
{ free(b->offset); free(b->detJac); free(b->weight); free(b->point); free(b->value); memset(b,0,sizeof(*b)); }

static int basis_init_quadrature(Basis *b, const Axis *ax, int nqp) /* petigabasis.c:83-219 (LEGENDRE rule) */
{
  int p = ax->p, iel, iqp, nel = ax->nel, nen = p+1, ndr = 5, d = p < 4 ? p : 4, a, k;
  const double *U = ax->U, *X, *W;
  if (nqp < 1) nqp = p+1;                       /* petigabasis.c:103 */
  if (nqp > 10 || p > MAXP) return 1;
  X = GAUSS_X[nqp]; W = GAUSS_W[nqp];
  basis_free(b);
  b->nel = nel; b->nqp = nqp; b->nen = nen;
  b->offset = (int*)malloc((size_t)nel*sizeof(int));
  b->detJac = (double*)malloc((size_t)nel*sizeof(double));
  b->weight = (double*)malloc((size_t)nel*nqp*sizeof(double));
  b->point  = (double*)malloc((size_t)nel*nqp*sizeof(double));
  b->value  = (double*)calloc((size_t)nel*nqp*nen*ndr, sizeof(double));
  for (iel = 0; iel < nel; iel++) {
    int kk = ax->span[iel];
    double u0 = U[kk], u1 = U[kk+1], J = (u1-u0)/2;
    double *w = b->weight + iel*nqp, *u = b->point + iel*nqp;
    b->detJac[iel] = J;
    for (iqp = 0; iqp < nqp; iqp++) { w[iqp] = W[iqp]; u[iqp] = (X[iqp] + 1) * J + u0; }
  }
  for (iel = 0; iel < nel; iel++) {
    int kk = ax->span[iel];
    double *w = b->weight + iel*nqp, *u = b->point + iel*nqp, *N = b->value + (size_t)iel*nqp*nen*ndr;
    b->offset[iel] = kk - p;
    for (iqp = 0; iqp < nqp && w[iqp] > 0; iqp++) {
      double ders[MAXP+1][5];
      bspline_ders(kk, u[iqp], p, d, U, ders);
      for (a = 0; a < nen; a++) for (k = 0; k < 5; k++) N[(iqp*nen + a)*ndr + k] = ders[a][k];
    }
  }
  { /* boundary tables: petigabasis.c:208-217 (k0 = p, u0 = U[k0]; k1 = n, u1 = U[k1+1]) */
    int n = ax->m - p - 1, side; int kb[2]; double ub[2];
    kb[0] = p; ub[0] = U[p]; kb[1] = n; ub[1] = U[n+1];
    for (side = 0; side < 2; side++) {
      double ders[MAXP+1][5];
      memset(ders, 0, sizeof(ders));
      bspline_ders(kb[side], ub[side], p, d, U, ders);
      b->bnd_point[side] = ub[side];
      for (a = 0; a < nen; a++) for (k = 0; k < 5; k++) b->bnd_value[side][a*ndr + k] = ders[a][k];
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* partition: src/petigapart.c (whole file)                                                   */
/* ------------------------------------------------------------------------------------------ */

static int cut2d(int M,int N,int m,int n) { return M*(n-1) + N*(m-1); }
static int cut3d(int M,int N,int P,int m,int n,int p) { return N*P*(m-1) + M*P*(n-1) + M*N*(p-1); }

static int part2d_inner(int size,int M,int N,int *_m,int *_n)
{
  int m,n;
  m = (int)(0.5 + sqrt(((double)M)/((double)N)*((double)size)));
  if (m == 0) {m = 1;} while (m > 0 && size % m) m--;
  n = size / m;
  *_m = m; *_n = n;
  return cut2d(M,N,m,n);
}
static void part2d(int size,int M,int N,int *_m,int *_n)
{
  int m,n,m1,n1,a,m2,n2,b;
  a = part2d_inner(size,M,N,&m1,&n1);
  b = part2d_inner(size,N,M,&n2,&m2);
  if (a<b) {m = m1; n = n1;} else {m = m2; n = n2;}
  if (M == N && n < m) {int t = m; m = n; n = t;}
  *_m = m; *_n = n;
}
static int part3d_inner(int size,int M,int N,int P,int *_m,int *_n,int *_p)
{
  int m,n,p,C,mm,nn,pp,CC;
  m = (int)(0.5 + pow(((double)M*(double)M)/((double)N*(double)P)*(double)size,1./3.));
  if (m == 0) {m = 1;} while (m > 0 && size % m) m--;
  part2d(size/m,N,P,&n,&p);
  C = cut3d(M,N,P,m,n,p);
  for (mm=m; mm>=1; mm--) {
    if (size % mm) continue;
    part2d(size/mm,N,P,&nn,&pp);
    CC = cut3d(M,N,P,mm,nn,pp);
    if (CC < C) {m = mm; n = nn; p = pp; C = CC;}
  }
  for (nn=n; nn>=1; nn--) {
    if (size % nn) continue;
    part2d(size/nn,M,P,&mm,&pp);
    CC = cut3d(M,N,P,mm,nn,pp);
    if (CC < C) {m = mm; n = nn; p = pp; C = CC;}
  }
  for (pp=p; pp>=1; pp--) {
    if (size % pp) continue;
    part2d(size/pp,M,N,&mm,&nn);
    CC = cut3d(M,N,P,mm,nn,pp);
    if (CC < C) {m = mm; n = nn; p = pp; C = CC;}
  }
  *_m = m; *_n = n; *_p = p;
  return cut3d(M,N,P,m,n,p);
}
static void part3d(int size,int M,int N,int P,int *_m,int *_n,int *_p)
{
  int m[3],n[3],p[3],C[3],k,i=0,Cmin=INT_MAX,t;
  C[0] = part3d_inner(size,M,N,P,&m[0],&n[0],&p[0]);
  C[1] = part3d_inner(size,N,M,P,&n[1],&m[1],&p[1]);
  C[2] = part3d_inner(size,P,M,N,&p[2],&m[2],&n[2]);
  for (k=0; k<3; k++) if (C[k]<Cmin) {Cmin=C[k]; i=k;}
  if (M == N && n[i] < m[i]) {t = m[i]; m[i] = n[i]; n[i] = t;}
  if (M == P && p[i] < m[i]) {t = m[i]; m[i] = p[i]; p[i] = t;}
  if (N == P && p[i] < n[i]) {t = n[i]; n[i] = p[i]; p[i] = t;}
  *_m = m[i]; *_n = n[i]; *_p = p[i];
}

/* IGA_Partition (petigapart.c:136-168); n[] entries < 1 mean "decide" */
int oiga_partition(int size,int rank,int dim,const int N[],int n[],int i[])
{
  int k,p=1;
  if (size < 1 || rank < 0 || rank >= size) return 1;
  switch (dim) {
  case 3:
    if (n[0]<1 && n[1]<1 && n[2]<1) part3d(size,N[0],N[1],N[2],&n[0],&n[1],&n[2]);
    else if (n[0]<1 && n[1]<1) part2d(size/n[2],N[0],N[1],&n[0],&n[1]);
    else if (n[0]<1 && n[2]<1) part2d(size/n[1],N[0],N[2],&n[0],&n[2]);
    else if (n[1]<1 && n[2]<1) part2d(size/n[0],N[1],N[2],&n[1],&n[2]);
    else if (n[0]<1) n[0] = size/(n[1]*n[2]);
    else if (n[1]<1) n[1] = size/(n[0]*n[2]);
    else if (n[2]<1) n[2] = size/(n[0]*n[1]);
    break;
  case 2:
    if (n[0]<1 && n[1]<1) part2d(size,N[0],N[1],&n[0],&n[1]);
    else if (n[0]<1) n[0] = size/n[1];
    else if (n[1]<1) n[1] = size/n[0];
    break;
  case 1: if (n[0] < 1) n[0] = size; break;
  default: return 1;
  }
  for (k=0; k<dim; k++) p *= n[k];
  if (p != size) return 2;
  for (k=0; k<dim; k++) if (N[k] < n[k]) return 3;
  if (i) for (k=0; k<dim; k++) { i[k] = rank % n[k]; rank -= i[k]; rank /= n[k]; }
  return 0;
}

static void dist1d(int size,int rank,int N,int *n,int *s) /* petigapart.c:170-176 */
{
  *n = N/size + ((N % size) > rank);
  *s = rank * (N/size) + (((N % size) > rank) ? rank : (N % size));
}

/* ------------------------------------------------------------------------------------------ */
/* setup: src/petiga.c:1111-1310 (Stage1), :1450-1493 (Stage3)                                */
/* ------------------------------------------------------------------------------------------ */

/* node box of proc coordinate r along axis i (the Stage1 arithmetic, callable for any rank) */
static void node_box_1d(const OIGA *o, int i, int r, int *lstart, int *lwidth, int *gstart, int *gwidth,
                        int *estart, int *ewidth)
{
  const Axis *ax = &o->axis[i];
  int nel = ax->nel, p = ax->p, *span = ax->span, ew, es, efirst, elast, ls, le, gs, ge;
  dist1d(o->proc_sizes[i], r, nel, &ew, &es);
  efirst = es; elast = es + ew - 1;
  gs = span[efirst] - p; ge = span[elast] + 1; ls = span[efirst] - p;
  le = (elast < nel-1) ? span[elast+1] - p : span[elast] + 1;
  *lstart = ls; *lwidth = le - ls; *gstart = gs; *gwidth = ge - gs;
  if (r == o->proc_sizes[i]-1) *lwidth = ax->nnp - ls;       /* petiga.c:1206-1207 */
  if (estart) *estart = es; if (ewidth) *ewidth = ew;
}

static void free_rank_arrays(OIGA *o)
{
  free(o->lgmap); free(o->geometryX); free(o->rationalW); free(o->fixtableU);
  o->lgmap = NULL; o->geometryX = o->rationalW = o->fixtableU = NULL;
}

/* global (PETSc) node index of natural node (i,j,k): AO of src/petigagrid.c:185-199 =
   rstart(owner) + lexicographic position inside the owner's box */
static void build_owner_tables(OIGA *o)
{
  int d, rr, q, i, size = o->proc_sizes[0]*o->proc_sizes[1]*o->proc_sizes[2], start = 0;
  for (d = 0; d < 3; d++) {
    int P = o->proc_sizes[d], gs, gw;
    free(o->own[d]); free(o->box_ls[d]); free(o->box_lw[d]);
    o->own[d] = (int*)malloc((size_t)o->axis[d].nnp*sizeof(int));
    o->box_ls[d] = (int*)malloc((size_t)P*sizeof(int)); o->box_lw[d] = (int*)malloc((size_t)P*sizeof(int));
    for (i = 0; i < o->axis[d].nnp; i++) o->own[d][i] = -1;
    for (rr = 0; rr < P; rr++) {
      node_box_1d(o, d, rr, &o->box_ls[d][rr], &o->box_lw[d][rr], &gs, &gw, NULL, NULL);
      for (i = o->box_ls[d][rr]; i < o->box_ls[d][rr]+o->box_lw[d][rr]; i++) if (i >= 0 && i < o->axis[d].nnp) o->own[d][i] = rr;
    }
  }
  free(o->rstart); o->rstart = (int*)malloc((size_t)(size+1)*sizeof(int));
  for (q = 0; q < size; q++) {
    int vol = 1, qq = q;
    for (d = 0; d < 3; d++) { int c = qq % o->proc_sizes[d]; qq /= o->proc_sizes[d]; vol *= o->box_lw[d][c]; }
    o->rstart[q] = start; start += vol;
  }
  o->rstart[size] = start;
}
static int natural_to_global(const OIGA *o, const int A[3])
{
  int r[3], d, rank;
  for (d = 0; d < 3; d++) { r[d] = o->own[d][A[d]]; if (r[d] < 0) return -1; }
  rank = r[0] + r[1]*o->proc_sizes[0] + r[2]*o->proc_sizes[0]*o->proc_sizes[1];
  return o->rstart[rank] + (A[0]-o->box_ls[0][r[0]]) + o->box_lw[0][r[0]]*((A[1]-o->box_ls[1][r[1]]) + o->box_lw[1][r[1]]*(A[2]-o->box_ls[2][r[2]]));
}

static int setup_tables(OIGA *o)
{
  int i, dim = o->dim;
  for (i = dim; i < 3; i++) axis_reset(&o->axis[i]);
  if (o->order < 0) {                                   /* petiga.c:1472-1476, IGASetOrder :470 */
    for (i = 0; i < dim; i++) if (o->axis[i].p > o->order) o->order = o->axis[i].p;
    if (o->order < 1) o->order = 1; if (o->order > 4) o->order = 4;
  }
  for (i = 0; i < 3; i++) {
    int q = (i < dim) ? o->rule_nqp[i] : 0;
    if (basis_init_quadrature(&o->basis[i], &o->axis[i], q)) return 1;
  }
  o->tables_ready = 1;
  return 0;
}

/* Stage1 for (size, rank) + the ghost-box arrays of Stage2 (lgmap) and geometry/fixtable gathers */
static int setup_rank(OIGA *o, int size, int rank)
{
  int i, dim = o->dim, N[3] = {1,1,1};
  if (!o->tables_ready && setup_tables(o)) return 1;
  o->size = size; o->rank = rank;
  for (i = 0; i < dim; i++) N[i] = o->axis[i].nel;
  { int n[3] = {0,0,0}, c[3] = {0,0,0};
    if (oiga_partition(size, rank, dim, N, n, c)) return 2;
    for (i = 0; i < 3; i++) { o->proc_sizes[i] = i<dim ? n[i] : 1; o->proc_ranks[i] = i<dim ? c[i] : 0; } }
  for (i = 0; i < 3; i++) {
    o->elem_sizes[i] = N[i];
    node_box_1d(o, i, o->proc_ranks[i], &o->node_lstart[i], &o->node_lwidth[i],
                &o->node_gstart[i], &o->node_gwidth[i], &o->elem_start[i], &o->elem_width[i]);
    o->node_sizes[i] = o->axis[i].nnp;
    o->geom_sizes[i] = o->axis[i].span[N[i]-1] + 1;
    o->geom_gstart[i] = o->node_gstart[i]; o->geom_gwidth[i] = o->node_gwidth[i];
  }
  free_rank_arrays(o);
  build_owner_tables(o);
  { /* lgmap: src/petigagrid.c:132-171 (periodic wrap) then AOApplicationToPetsc :213-228 */
    const int *gs = o->node_gstart, *gw = o->node_gwidth, *sz = o->node_sizes;
    int ii, jj, kk, pos = 0;
    o->lgmap = (int*)malloc((size_t)gw[0]*gw[1]*gw[2]*sizeof(int));
    for (kk = gs[2]; kk < gs[2]+gw[2]; kk++)
      for (jj = gs[1]; jj < gs[1]+gw[1]; jj++)
        for (ii = gs[0]; ii < gs[0]+gw[0]; ii++) {
          int A[3]; A[0] = ii; A[1] = jj; A[2] = kk;
          for (i = 0; i < 3; i++) { if (A[i] < 0) A[i] = sz[i] + A[i]; else if (A[i] >= sz[i]) A[i] = A[i] % sz[i]; }
          o->lgmap[pos++] = natural_to_global(o, A);
        }
  }
  if (o->geometry) { /* ghost-box slices of the natural geometry arrays (src/petigaio.c:255-286) */
    const int *gs = o->geom_gstart, *gw = o->geom_gwidth, *sz = o->geom_sizes;
    int nsd = o->geometry, ii, jj, kk, c, pos = 0;
    o->geometryX = (double*)malloc((size_t)gw[0]*gw[1]*gw[2]*nsd*sizeof(double));
    if (o->rational) o->rationalW = (double*)malloc((size_t)gw[0]*gw[1]*gw[2]*sizeof(double));
    for (kk = gs[2]; kk < gs[2]+gw[2]; kk++)
      for (jj = gs[1]; jj < gs[1]+gw[1]; jj++)
        for (ii = gs[0]; ii < gs[0]+gw[0]; ii++, pos++) {
          size_t nat = (size_t)ii + (size_t)sz[0]*((size_t)jj + (size_t)sz[1]*kk);
          for (c = 0; c < nsd; c++) o->geometryX[(size_t)pos*nsd + c] = o->geomX_nat[nat*nsd + c];
          if (o->rational) o->rationalW[pos] = o->geomW_nat[nat];
        }
  }
  if (o->fixtable) { /* G2L of the fix-table vector: src/petigaform.c IGASetFixTable, petigavec.c:147-169 */
    int ng = o->node_gwidth[0]*o->node_gwidth[1]*o->node_gwidth[2], a, c, dof = o->dof;
    o->fixtableU = (double*)malloc((size_t)ng*dof*sizeof(double));
    for (a = 0; a < ng; a++) for (c = 0; c < dof; c++)
      o->fixtableU[(size_t)a*dof + c] = o->fixtable_glob[(size_t)o->lgmap[a]*dof + c];
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* sparsity pattern: src/petigamat.c:197-267 (Stencil, ColumnIndices), :448-537                */
/* ------------------------------------------------------------------------------------------ */

static void stencil(const OIGA *o, int dir, int i, int *first, int *last) /* petigamat.c:197-233 */
{
  const Axis *ax = &o->axis[dir];
  int p = ax->p, m = ax->m, n = m - p - 1, k; const double *U = ax->U;
  k = next_knot(m,U,i,+1); *first = k - p - 1;
  k = next_knot(m,U,i+p+1,-1); *last = k;
  if (!ax->periodic) {
    if (i <= p)   *first = 0;
    if (i >= n-p) *last  = n;
  } else if (i == 0) {
    int kk = n+1, j = next_knot(m,U,kk,+1), s = j-kk, C = p-s, nnp = n-C;
    kk = next_knot(m,U,nnp,+1) - nnp;
    *first = kk - p - 1;
  }
}

static int cmp_int(const void *a, const void *b) { int x = *(const int*)a, y = *(const int*)b; return (x>y)-(x<y); }

/* Block-row pattern of the rows owned by the current rank, global column ids sorted ascending and
   de-duplicated (what PETSc's preallocated AIJ/BAIJ row holds).  Returns nnz blocks; arrays malloc'ed. */
static long pattern_rank(const OIGA *o, int **rowptr_out, int **colidx_out, int *nrows_out)
{
  int dim = o->dim, i, j, k, d;
  const int *ls = o->node_lstart, *lw = o->node_lwidth, *sz = o->node_sizes;
  int gstart[3] = {0,0,0}, gwidth[3] = {1,1,1};
  int nrows = lw[0]*lw[1]*lw[2], row = 0, maxnnz = 1, *rowptr, *colidx, *tmp; long cap, nnz = 0;
  for (d = 0; d < dim; d++) {            /* petigamat.c:414-420 */
    int gfirst, glast, first = ls[d], last = ls[d] + lw[d] - 1;
    stencil(o,d,first,&gstart[d],&glast);
    stencil(o,d,last,&gfirst,&glast);
    gwidth[d] = glast + 1 - gstart[d];
    maxnnz *= (2*o->axis[d].p + 1);
  }
  rowptr = (int*)malloc((size_t)(nrows+1)*sizeof(int));
  cap = (long)nrows*maxnnz; colidx = (int*)malloc((size_t)cap*sizeof(int));
  tmp = (int*)malloc((size_t)maxnnz*sizeof(int));
  rowptr[0] = 0;
  for (k = ls[2]; k < ls[2]+lw[2]; k++)
    for (j = ls[1]; j < ls[1]+lw[1]; j++)
      for (i = ls[0]; i < ls[0]+lw[0]; i++) {
        int first[3] = {0,0,0}, last[3] = {0,0,0}, A[3], count = 0, ii, jj, kk, c, u;
        A[0] = i; A[1] = j; A[2] = k;
        for (d = 0; d < dim; d++) {      /* ColumnIndices: petigamat.c:243-267 */
          stencil(o,d,A[d],&first[d],&last[d]);
          if (first[d] < gstart[d]) first[d] = gstart[d];
          if (last[d] > gstart[d]+gwidth[d]-1) last[d] = gstart[d]+gwidth[d]-1;
        }
        for (kk = first[2]; kk <= last[2]; kk++)
          for (jj = first[1]; jj <= last[1]; jj++)
            for (ii = first[0]; ii <= last[0]; ii++) {
              int B[3]; B[0] = ii; B[1] = jj; B[2] = kk;   /* ghost wrap: petigagrid.c:160-163 */
              for (d = 0; d < 3; d++) { if (B[d] < 0) B[d] = sz[d] + B[d]; else if (B[d] >= sz[d]) B[d] = B[d] % sz[d]; }
              tmp[count++] = natural_to_global(o, B);
            }
        qsort(tmp, (size_t)count, sizeof(int), cmp_int);
        for (c = 0, u = 0; c < count; c++) if (c == 0 || tmp[c] != tmp[c-1]) tmp[u++] = tmp[c];
        memcpy(colidx + nnz, tmp, (size_t)u*sizeof(int));
        nnz += u; rowptr[++row] = (int)nnz;
      }
  free(tmp);
  *rowptr_out = rowptr; *colidx_out = colidx; *nrows_out = nrows;
  return nnz;
}

/* ------------------------------------------------------------------------------------------ */
/* element tabulation: src/petiga{1,2,3}d.F90 + .f90.in fragments, generic in dim              */
/* ------------------------------------------------------------------------------------------ */

static int ipow(int b, int e) { int r = 1; while (e-- > 0) r *= b; return r; }

typedef struct {
  const OIGA *o;
  int ID[3], nen, nqp, dim, nsd, dof, order;
  int nenA[3], nqpA[3];
  int *mapping;
  double *W, *X;                         /* rationalW[nen], geometryX[nen][nsd] */
  double *weight, *detJac;               /* [nqp] */
  double *basis[4], *shape[4];           /* [nqp][nen][dim^k] */
  double *mapU[4], *mapX[4], *detX;      /* mapU[0]=point[nqp][dim]; mapU[k][nqp][dim][nsd^k]; mapX[k][nqp][nsd][dim^k] */
  int nfix, *ifix; double *vfix, *ufix; int nflux, *iflux; double *vflux;
  int geometry, rational;
  int atboundary, baxis, bside;          /* boundary pass (petigaelem.c:427-447): face = (baxis, bside) */
  double *normal, *detS;                 /* [nqp][nsd], [nqp] (petigaelem.c:1012-1022) */
} Elem;

static void elem_alloc(Elem *e, const OIGA *o)
{
  int k, dim = o->dim, nsd = o->geometry ? o->geometry : dim, nen = 1, nqp = 1, i;
  memset(e, 0, sizeof(*e));
  e->o = o; e->dim = dim; e->nsd = nsd; e->dof = o->dof; e->order = o->order > 3 ? 3 : o->order;
  e->geometry = o->geometry ? 1 : 0; e->rational = o->rational;
  for (i = 0; i < 3; i++) { e->nenA[i] = o->basis[i].nen; e->nqpA[i] = o->basis[i].nqp; nen *= e->nenA[i]; nqp *= e->nqpA[i]; }
  e->nen = nen; e->nqp = nqp;
  e->mapping = (int*)malloc((size_t)nen*sizeof(int));
  e->W = (double*)calloc((size_t)nen, sizeof(double)); e->X = (double*)calloc((size_t)nen*nsd, sizeof(double));
  e->weight = (double*)calloc((size_t)nqp, sizeof(double)); e->detJac = (double*)calloc((size_t)nqp, sizeof(double));
  e->detX = (double*)calloc((size_t)nqp, sizeof(double));
  e->detS = (double*)calloc((size_t)nqp, sizeof(double)); e->normal = (double*)calloc((size_t)nqp*nsd, sizeof(double));
  for (k = 0; k < 4; k++) {
    e->basis[k] = (double*)calloc((size_t)nqp*nen*ipow(dim,k), sizeof(double));
    e->shape[k] = (double*)calloc((size_t)nqp*nen*ipow(nsd,k), sizeof(double));
    e->mapU[k]  = (double*)calloc((size_t)nqp*dim*ipow(nsd,k), sizeof(double));
    e->mapX[k]  = (double*)calloc((size_t)nqp*nsd*ipow(dim,k), sizeof(double));
  }
  e->ifix = (int*)malloc((size_t)nen*o->dof*sizeof(int)); e->vfix = (double*)malloc((size_t)nen*o->dof*sizeof(double));
  e->ufix = (double*)malloc((size_t)nen*o->dof*sizeof(double));
  e->iflux = (int*)malloc((size_t)nen*o->dof*sizeof(int)); e->vflux = (double*)malloc((size_t)nen*o->dof*sizeof(double));
}
static void elem_free(Elem *e)
{
  int k;
  free(e->mapping); free(e->W); free(e->X); free(e->weight); free(e->detJac); free(e->detX); free(e->detS); free(e->normal);
  for (k = 0; k < 4; k++) { free(e->basis[k]); free(e->shape[k]); free(e->mapU[k]); free(e->mapX[k]); }
  free(e->ifix); free(e->vfix); free(e->ufix); free(e->iflux); free(e->vflux);
}

static void elem_closure(Elem *e) /* src/petigaelem.c:693-755 */
{
  const OIGA *o = e->o; const int *ID = e->ID;
  int ia, ja, ka, a = 0, nsd = e->nsd, i;
  int ioff = o->basis[0].offset[ID[0]], joff = o->basis[1].offset[ID[1]], koff = o->basis[2].offset[ID[2]];
  const int *start = o->node_gstart, *width = o->node_gwidth;
  int jstride = width[0], kstride = width[0]*width[1];
  for (ka = 0; ka < e->nenA[2]; ka++)
    for (ja = 0; ja < e->nenA[1]; ja++)
      for (ia = 0; ia < e->nenA[0]; ia++) {
        int iA = (ioff+ia) - start[0], jA = (joff+ja) - start[1], kA = (koff+ka) - start[2];
        e->mapping[a++] = iA + jA*jstride + kA*kstride;
      }
  if (e->rational) for (a = 0; a < e->nen; a++) e->W[a] = o->rationalW[e->mapping[a]];
  if (e->geometry) for (a = 0; a < e->nen; a++) for (i = 0; i < nsd; i++) e->X[i + a*nsd] = o->geometryX[(size_t)e->mapping[a]*nsd + i];
}

/* Determinant / Inverse: src/petigadet.f90.in, src/petigainv.f90.in.  A is Fortran A(dim,dim):
   A(r,c) at A[(c-1)*dim + (r-1)]. */
#define FA(A,r,c) (A)[((c)-1)*dim + ((r)-1)]
static double determinant(int dim, const double *A)
{
  switch (dim) {
  case 1: return FA(A,1,1);
  case 2: return + FA(A,1,1)*FA(A,2,2) - FA(A,2,1)*FA(A,1,2);
  case 3: return + FA(A,1,1) * ( FA(A,2,2)*FA(A,3,3) - FA(A,3,2)*FA(A,2,3) )
                 - FA(A,2,1) * ( FA(A,1,2)*FA(A,3,3) - FA(A,3,2)*FA(A,1,3) )
                 + FA(A,3,1) * ( FA(A,1,2)*FA(A,2,3) - FA(A,2,2)*FA(A,1,3) );
  }
  return 0;
}
static void inverse(int dim, double detA, const double *A, double *invA)
{
  int i;
  switch (dim) {
  case 1: invA[0] = 1/detA; return;
  case 2:
    FA(invA,1,1) = + FA(A,2,2); FA(invA,2,1) = - FA(A,2,1);
    FA(invA,1,2) = - FA(A,1,2); FA(invA,2,2) = + FA(A,1,1);
    break;
  case 3:
    FA(invA,1,1) = + FA(A,2,2)*FA(A,3,3) - FA(A,2,3)*FA(A,3,2);
    FA(invA,2,1) = - FA(A,2,1)*FA(A,3,3) + FA(A,2,3)*FA(A,3,1);
    FA(invA,3,1) = + FA(A,2,1)*FA(A,3,2) - FA(A,2,2)*FA(A,3,1);
    FA(invA,1,2) = - FA(A,1,2)*FA(A,3,3) + FA(A,1,3)*FA(A,3,2);
    FA(invA,2,2) = + FA(A,1,1)*FA(A,3,3) - FA(A,1,3)*FA(A,3,1);
    FA(invA,3,2) = - FA(A,1,1)*FA(A,3,2) + FA(A,1,2)*FA(A,3,1);
    FA(invA,1,3) = + FA(A,1,2)*FA(A,2,3) - FA(A,1,3)*FA(A,2,2);
    FA(invA,2,3) = - FA(A,1,1)*FA(A,2,3) + FA(A,1,3)*FA(A,2,1);
    FA(invA,3,3) = + FA(A,1,1)*FA(A,2,2) - FA(A,1,2)*FA(A,2,1);
    break;
  }
  for (i = 0; i < dim*dim; i++) invA[i] = invA[i]/detA;
}

/* interior tabulation of one element: src/petigaelem.c:794-1033 (atboundary == false branch) */
static void elem_tabulate(Elem *e)
{
  const OIGA *o = e->o; const int *ID = e->ID;
  int dim = e->dim, nsd = e->nsd, nen = e->nen, ord = e->order;
  int NQ[3], i, q, a, k, nqp;
  const double *V[3];
  const int bnd = e->atboundary, bax = e->baxis, bsd = e->bside;
  for (i = 0; i < 3; i++) {              /* IGA_Quadrature_SIZE: petigaelem.c:764-776 */
    const Basis *b = &o->basis[i]; int qq = b->nqp - 1; const double *w = b->weight + ID[i]*b->nqp;
    NQ[i] = 1; while (qq >= 0 && w[qq] <= 0) qq--; NQ[i] += qq;
    V[i] = b->value + (size_t)ID[i]*b->nqp*b->nen*5;
    if (bnd && i == bax) { NQ[i] = 1; V[i] = b->bnd_value[bsd]; }   /* nqp /= NQ[axis]; NQ[axis] = 1 (:813-817); IGA_BasisFuns_BNDR :791 */
  }
  nqp = e->nqp = NQ[0]*NQ[1]*NQ[2];
  { /* IGA_Quadrature_3D: petiga3d.F90:1-29 */
    int iq, jq, kq; q = 0;
    double J = 1;
    for (i = 0; i < dim; i++) { double Ji = (bnd && i == bax) ? 1.0 : o->basis[i].detJac[ID[i]];   /* IGA_Quadrature_BNDR :788: bnd_detJac = 1 */
      J = (i == 0) ? Ji : J * Ji; }
    for (kq = 0; kq < NQ[2]; kq++) for (jq = 0; jq < NQ[1]; jq++) for (iq = 0; iq < NQ[0]; iq++, q++) {
      int qi[3]; double w = 0; qi[0] = iq; qi[1] = jq; qi[2] = kq;
      for (i = 0; i < dim; i++) {
        const Basis *b = &o->basis[i];
        double pt = (bnd && i == bax) ? b->bnd_point[bsd] : b->point[ID[i]*b->nqp + qi[i]];
        double wi = (bnd && i == bax) ? 1.0 : b->weight[ID[i]*b->nqp + qi[i]];                      /* bnd_weight = 1 */
        e->mapU[0][q*dim + i] = pt;
        w = (i == 0) ? wi : w * wi;
      }
      e->weight[q] = w; e->detJac[q] = J;
    }
  }
  { /* TensorBasisFuns: petiga3d.F90:32-233 (generic in dim; derivative multi-index i1 fastest) */
    int iq, jq, kq, ia, ja, ka; q = 0;
    for (kq = 0; kq < NQ[2]; kq++) for (jq = 0; jq < NQ[1]; jq++) for (iq = 0; iq < NQ[0]; iq++, q++) {
      int qi[3]; qi[0] = iq; qi[1] = jq; qi[2] = kq; a = 0;
      for (ka = 0; ka < e->nenA[2]; ka++) for (ja = 0; ja < e->nenA[1]; ja++) for (ia = 0; ia < e->nenA[0]; ia++, a++) {
        int ai[3]; ai[0] = ia; ai[1] = ja; ai[2] = ka;
        for (k = 0; k <= ord; k++) {
          int nk = ipow(dim,k), idx;
          for (idx = 0; idx < nk; idx++) {
            int cnt[3] = {0,0,0}, t = idx, s; double v;
            for (s = 0; s < k; s++) { cnt[t % dim]++; t /= dim; }
            v = V[0][(qi[0]*e->nenA[0] + ai[0])*5 + cnt[0]];
            if (dim > 1) v = v * V[1][(qi[1]*e->nenA[1] + ai[1])*5 + cnt[1]];
            if (dim > 2) v = v * V[2][(qi[2]*e->nenA[2] + ai[2])*5 + cnt[2]];
            e->basis[k][((size_t)q*nen + a)*nk + idx] = v;
          }
        }
      }
    }
  }
  if (e->rational) { /* Rationalize: src/petigarat.f90.in:3-57 */
    int j, l, d1 = dim, d2 = dim*dim, d3 = dim*dim*dim;
    for (q = 0; q < nqp; q++) {
      double *R0 = e->basis[0] + (size_t)q*nen, *R1 = e->basis[1] + (size_t)q*nen*d1;
      double *R2 = e->basis[2] + (size_t)q*nen*d2, *R3 = e->basis[3] + (size_t)q*nen*d3;
      double W0 = 0, W1[3], W2[9], W3[27]; const double *W = e->W;
      for (a = 0; a < nen; a++) R0[a] = W[a] * R0[a];
      for (a = 0; a < nen; a++) W0 += R0[a];
      for (a = 0; a < nen; a++) R0[a] = R0[a] / W0;
      if (ord < 1) continue;
      for (i = 0; i < dim; i++) {
        W1[i] = 0; for (a = 0; a < nen; a++) W1[i] += W[a]*R1[a*d1+i];
        for (a = 0; a < nen; a++) R1[a*d1+i] = W[a]*R1[a*d1+i] - R0[a]*W1[i];
      }
      for (a = 0; a < nen*d1; a++) R1[a] = R1[a] / W0;
      if (ord < 2) continue;
      for (j = 0; j < dim; j++) for (i = 0; i < dim; i++) {
        int ij = j*dim + i; W2[ij] = 0; for (a = 0; a < nen; a++) W2[ij] += W[a]*R2[a*d2+ij];
        for (a = 0; a < nen; a++)
          R2[a*d2+ij] = W[a]*R2[a*d2+ij] - R0[a]*W2[ij] - R1[a*d1+i]*W1[j] - R1[a*d1+j]*W1[i];
      }
      for (a = 0; a < nen*d2; a++) R2[a] = R2[a] / W0;
      if (ord < 3) continue;
      for (l = 0; l < dim; l++) for (j = 0; j < dim; j++) for (i = 0; i < dim; i++) {
        int ijk = (l*dim + j)*dim + i; W3[ijk] = 0; for (a = 0; a < nen; a++) W3[ijk] += W[a]*R3[a*d3+ijk];
        for (a = 0; a < nen; a++)
          R3[a*d3+ijk] = W[a]*R3[a*d3+ijk] - R0[a]*W3[ijk]
            - R1[a*d1+i]*W2[l*dim+j] - R1[a*d1+j]*W2[l*dim+i] - R1[a*d1+l]*W2[j*dim+i]
            - R2[a*d2+l*dim+j]*W1[i] - R2[a*d2+l*dim+i]*W1[j] - R2[a*d2+j*dim+i]*W1[l];
      }
      for (a = 0; a < nen*d3; a++) R3[a] = R3[a] / W0;
    }
  }
  if (!e->geometry) {                    /* identity map: petigaelem.c:350-358; shape aliases basis :549-561 */
    for (q = 0; q < nqp; q++) {
      e->detX[q] = 1.0;
      memset(e->mapX[1] + (size_t)q*nsd*dim, 0, sizeof(double)*nsd*dim);
      memset(e->mapU[1] + (size_t)q*dim*nsd, 0, sizeof(double)*nsd*dim);
      for (i = 0; i < dim; i++) { e->mapX[1][q*nsd*dim + i*(dim+1)] = 1.0; e->mapU[1][q*dim*nsd + i*(dim+1)] = 1.0; }
    }
    for (k = 0; k <= ord; k++) memcpy(e->shape[k], e->basis[k], sizeof(double)*(size_t)nqp*nen*ipow(dim,k));
    if (bnd) for (q = 0; q < nqp; q++) {   /* petigaelem.c:1018-1021: unit normal of the parametric face, detS = 1 */
      double *n = e->normal + (size_t)q*nsd;
      for (i = 0; i < nsd; i++) n[i] = 0.0;
      e->detS[q] = 1.0; n[bax] = bsd ? 1.0 : -1.0;
    }
    return;
  }
  /* GeometryMap: src/petigamapgeo.f90.in:3-71.  X_k(:,i) += X(i,node)*M_k(:,node); C layout mapX[k][q][i][dim^k] */
  for (k = 0; k <= ord; k++) {
    int nk = ipow(dim,k), c;
    for (q = 0; q < nqp; q++) {
      double *Xk = e->mapX[k] + (size_t)q*nsd*nk; const double *Mk = e->basis[k] + (size_t)q*nen*nk;
      for (c = 0; c < nsd*nk; c++) Xk[c] = 0;
      for (a = 0; a < nen; a++) for (i = 0; i < nsd; i++) for (c = 0; c < nk; c++)
        Xk[i*nk + c] = Xk[i*nk + c] + e->X[i + a*nsd]*Mk[a*nk + c];
    }
  }
  if (dim != nsd) return;                /* manifolds: out of scope */
  /* InverseMap: src/petigamapinv.f90.in:3-74.  Fortran X1(a,i)=X1[i*dim+a]; E1(i,a)=E1[a*nsd+i];
     X2(b,a,k)=X2[(k*dim+a)*dim+b]; E2(j,i,c)=E2[(c*nsd+i)*nsd+j]; X3(c,b,a,l); E3(k,j,i,d). */
  for (q = 0; q < nqp; q++) {
    const double *X1 = e->mapX[1] + (size_t)q*nsd*dim, *X2 = e->mapX[2] + (size_t)q*nsd*dim*dim, *X3 = e->mapX[3] + (size_t)q*nsd*dim*dim*dim;
    double *E1 = e->mapU[1] + (size_t)q*dim*nsd, *E2 = e->mapU[2] + (size_t)q*dim*nsd*nsd, *E3 = e->mapU[3] + (size_t)q*dim*nsd*nsd*nsd;
    int j, kk, l, b, c, d;
    if (ord < 1) break;
    e->detX[q] = determinant(dim, X1);
    inverse(dim, e->detX[q], X1, E1);
    if (ord < 2) continue;
#define X2F(b,a,k) X2[((k)*dim+(a))*dim+(b)]
#define X3F(c,b,a,l) X3[(((l)*dim+(a))*dim+(b))*dim+(c)]
#define E1F(i,a) E1[(a)*nsd+(i)]
#define E2F(j,i,c) E2[((c)*nsd+(i))*nsd+(j)]
#define E3F(k,j,i,d) E3[(((d)*nsd+(i))*nsd+(j))*nsd+(k)]
    for (c = 0; c < dim*nsd*nsd; c++) E2[c] = 0;
    for (i = 0; i < nsd; i++) for (j = 0; j < nsd; j++) for (kk = 0; kk < nsd; kk++)
      for (a = 0; a < dim; a++) for (b = 0; b < dim; b++) for (c = 0; c < dim; c++)
        E2F(j,i,c) = E2F(j,i,c) - X2F(b,a,kk)*E1F(i,a)*E1F(j,b)*E1F(kk,c);
    if (ord < 3) continue;
    for (c = 0; c < dim*nsd*nsd*nsd; c++) E3[c] = 0;
    for (d = 0; d < dim; d++) for (i = 0; i < nsd; i++) for (j = 0; j < nsd; j++) for (kk = 0; kk < nsd; kk++)
      for (a = 0; a < dim; a++) for (b = 0; b < dim; b++) for (l = 0; l < nsd; l++) {
        for (c = 0; c < dim; c++)
          E3F(kk,j,i,d) = E3F(kk,j,i,d) - X3F(c,b,a,l)*E1F(i,a)*E1F(j,b)*E1F(kk,c)*E1F(l,d);
        E3F(kk,j,i,d) = E3F(kk,j,i,d) - X2F(b,a,l)*(E1F(i,a)*E2F(kk,j,b)+E1F(j,b)*E2F(kk,i,a)+E1F(kk,b)*E2F(j,i,a))*E1F(l,d);
      }
  }
  /* ShapeFunctions: src/petigamapshf.f90.in:3-83; N0 copied (petigaelem.c:995) */
  memcpy(e->shape[0], e->basis[0], sizeof(double)*(size_t)nqp*nen);
  for (q = 0; q < nqp; q++) {
    const double *E1 = e->mapU[1] + (size_t)q*dim*nsd, *E2 = e->mapU[2] + (size_t)q*dim*nsd*nsd, *E3 = e->mapU[3] + (size_t)q*dim*nsd*nsd*nsd;
    int j, kk, b, c, node;
    if (ord < 1) break;
    for (node = 0; node < nen; node++) {
      const double *N1 = e->basis[1] + ((size_t)q*nen + node)*dim, *N2 = e->basis[2] + ((size_t)q*nen + node)*dim*dim;
      const double *N3 = e->basis[3] + ((size_t)q*nen + node)*dim*dim*dim;
      double *R1 = e->shape[1] + ((size_t)q*nen + node)*nsd, *R2 = e->shape[2] + ((size_t)q*nen + node)*nsd*nsd;
      double *R3 = e->shape[3] + ((size_t)q*nen + node)*nsd*nsd*nsd;
      for (i = 0; i < nsd; i++) { R1[i] = 0; for (a = 0; a < dim; a++) R1[i] = R1[i] + N1[a]*E1F(i,a); }
      if (ord < 2) continue;
      for (i = 0; i < nsd; i++) for (j = 0; j < nsd; j++) {
        double r = 0;
        for (a = 0; a < dim; a++) {
          for (b = 0; b < dim; b++) r = r + N2[a*dim+b]*E1F(i,a)*E1F(j,b);
          r = r + N1[a]*E2F(j,i,a);
        }
        R2[i*nsd+j] = r;
      }
      if (ord < 3) continue;
      for (i = 0; i < nsd; i++) for (j = 0; j < nsd; j++) for (kk = 0; kk < nsd; kk++) {
        double r = 0;
        for (a = 0; a < dim; a++) {
          for (b = 0; b < dim; b++) {
            for (c = 0; c < dim; c++) r = r + N3[(a*dim+b)*dim+c]*E1F(i,a)*E1F(j,b)*E1F(kk,c);
            r = r + N2[a*dim+b]*(E1F(i,a)*E2F(kk,j,b)+E1F(j,b)*E2F(kk,i,a)+E1F(kk,b)*E2F(j,i,a));
          }
          r = r + N1[a]*E3F(kk,j,i,a);
        }
        R3[(i*nsd+j)*nsd+kk] = r;
      }
    }
  }
  if (bnd) {   /* normal and detS: petigaelem.c:1012-1022, IGA_GetNormal src/petigaval.F90:45-99 */
    for (q = 0; q < nqp; q++) {
      double *n = e->normal + (size_t)q*nsd; const double *F = e->mapX[1] + (size_t)q*nsd*dim;   /* F[i*dim+d] = dx_i/du_d */
      if (e->geometry && dim == nsd) {
        double dS = 1;
        if (dim == 1) { n[0] = 1; }
        else if (dim == 2) {   /* t = +F(2,:) (axis 0) or -F(1,:) (axis 1); N = (t2, -t1) */
          double t[2]; int dd = (bax == 0) ? 1 : 0; double sg = (bax == 0) ? 1.0 : -1.0;
          t[0] = sg*F[0*dim+dd]; t[1] = sg*F[1*dim+dd];
          n[0] = +t[1]; n[1] = -t[0];
          dS = sqrt(n[0]*n[0] + n[1]*n[1]); n[0] /= dS; n[1] /= dS;
        } else {               /* s, t = rows (axis+1, axis+2) cyclic; N = s x t */
          int ds = (bax+1)%3, dt = (bax+2)%3; double sv[3], tv[3];
          for (i = 0; i < 3; i++) { sv[i] = F[i*dim+ds]; tv[i] = F[i*dim+dt]; }
          n[0] = sv[1]*tv[2] - sv[2]*tv[1]; n[1] = sv[2]*tv[0] - sv[0]*tv[2]; n[2] = sv[0]*tv[1] - sv[1]*tv[0];
          dS = sqrt(n[0]*n[0] + n[1]*n[1] + n[2]*n[2]); n[0] /= dS; n[1] /= dS; n[2] /= dS;
        }
        if (bsd == 0) for (i = 0; i < dim; i++) n[i] = -n[i];
        e->detS[q] = dS;
      } else {
        for (i = 0; i < nsd; i++) n[i] = 0.0;
        e->detS[q] = 1.0; n[bax] = bsd ? 1.0 : -1.0;
      }
    }
    if (e->geometry) for (q = 0; q < nqp; q++) e->detJac[q] *= e->detS[q];   /* :1027-1028 */
    return;
  }
  for (q = 0; q < nqp; q++) e->detJac[q] *= e->detX[q];   /* petigaelem.c:1024-1029 */
}

/* ------------------------------------------------------------------------------------------ */
/* boundary fix-up lists: src/petigaelem.c:1118-1283                                           */
/* ------------------------------------------------------------------------------------------ */

static double boundary_area(const Elem *e, int dir, int side) /* petigaelem.c:1118-1164 + petiga{2,3}d.F90 BoundaryArea */
{
  const OIGA *o = e->o; const int *ID = e->ID; double A = 1; int i, dim = e->dim;
  if (dim == 1) return A;
  for (i = 0; i < dim; i++) if (i != dir) A *= o->basis[i].detJac[ID[i]]/(double)o->basis[i].nen;
  if (!e->geometry) { A *= (dim == 2) ? 2 : 4; return A; }
  { /* IGA_BoundaryArea_{2,3}D: surface Jacobian integrated over the face */
    int ax[2], n = 0, nq[2] = {1,1}, ne[2] = {1,1}, iq, jq, ia, ja, nsd = e->nsd, sd = dim-1, c, r;
    const double *Wq[2] = {NULL,NULL}, *Nv[2] = {NULL,NULL}; double dS = 0;
    int kfix = side ? e->nenA[dir]-1 : 0;
    for (i = 0; i < dim; i++) if (i != dir) {
      const Basis *b = &o->basis[i]; int qq = b->nqp - 1; const double *w = b->weight + ID[i]*b->nqp;
      while (qq >= 0 && w[qq] <= 0) qq--;
      ax[n] = i; nq[n] = qq+1; ne[n] = b->nen; Wq[n] = w; Nv[n] = b->value + (size_t)ID[i]*b->nqp*b->nen*5; n++;
    }
    for (jq = 0; jq < nq[1]; jq++) for (iq = 0; iq < nq[0]; iq++) {
      double N0[(MAXP+1)*(MAXP+1)], N1[(MAXP+1)*(MAXP+1)][2], F[2][3], M[4], detJ, W0, S1[2], wq;
      int nn = ne[0]*ne[1];
      for (ja = 0; ja < ne[1]; ja++) for (ia = 0; ia < ne[0]; ia++) {
        double i0 = Nv[0][(iq*ne[0]+ia)*5+0], i1 = Nv[0][(iq*ne[0]+ia)*5+1];
        double j0 = sd > 1 ? Nv[1][(jq*ne[1]+ja)*5+0] : 1, j1 = sd > 1 ? Nv[1][(jq*ne[1]+ja)*5+1] : 0;
        N0[ja*ne[0]+ia] = sd > 1 ? i0*j0 : i0;
        N1[ja*ne[0]+ia][0] = sd > 1 ? i1*j0 : i1;
        N1[ja*ne[0]+ia][1] = i0*j1;
      }
      if (e->rational) {
        W0 = 0;
        for (c = 0; c < nn; c++) { int idx[3], a; idx[dir] = kfix; idx[ax[0]] = c % ne[0]; if (sd > 1) idx[ax[1]] = c / ne[0];
          a = idx[0] + e->nenA[0]*((dim>1?idx[1]:0) + e->nenA[1]*(dim>2?idx[2]:0)); N0[c] = e->W[a]*N0[c]; W0 += N0[c]; }
        for (c = 0; c < nn; c++) N0[c] = N0[c]/W0;
        for (r = 0; r < sd; r++) { S1[r] = 0;
          for (c = 0; c < nn; c++) { int idx[3], a; idx[dir] = kfix; idx[ax[0]] = c % ne[0]; if (sd > 1) idx[ax[1]] = c / ne[0];
            a = idx[0] + e->nenA[0]*((dim>1?idx[1]:0) + e->nenA[1]*(dim>2?idx[2]:0)); N1[c][r] = e->W[a]*N1[c][r]; S1[r] += N1[c][r]; }
          for (c = 0; c < nn; c++) N1[c][r] = (N1[c][r] - N0[c]*S1[r])/W0; }
      }
      for (r = 0; r < sd; r++) for (c = 0; c < nsd; c++) F[r][c] = 0;
      for (c = 0; c < nn; c++) { int idx[3], a, s; idx[dir] = kfix; idx[ax[0]] = c % ne[0]; if (sd > 1) idx[ax[1]] = c / ne[0];
        a = idx[0] + e->nenA[0]*((dim>1?idx[1]:0) + e->nenA[1]*(dim>2?idx[2]:0));
        for (r = 0; r < sd; r++) for (s = 0; s < nsd; s++) F[r][s] += N1[c][r]*e->X[s + a*nsd]; }
      for (r = 0; r < sd; r++) for (c = 0; c < sd; c++) { int s; M[c*sd+r] = 0; for (s = 0; s < nsd; s++) M[c*sd+r] += F[r][s]*F[c][s]; }
      detJ = sqrt(fabs(determinant(sd, M)));
      wq = Wq[0][iq]; if (sd > 1) wq = wq*Wq[1][jq];
      dS = dS + detJ*wq;
    }
    A *= dS;
  }
  return A;
}

static void add_fixa(Elem *e, const FormBC *bc, int a) /* petigaelem.c:1166-1189 */
{
  int j, k, dof = e->dof;
  for (k = 0; k < bc->count; k++) {
    int c = bc->field[k], idx = a*dof + c; double val = bc->value[k];
    if (c >= dof) continue;
    if (e->o->fixtable) val = e->o->fixtableU[c + (size_t)e->mapping[a]*dof];
    for (j = 0; j < e->nfix; j++) if (e->ifix[j] == idx) break;
    if (j == e->nfix) e->nfix++;
    e->ifix[j] = idx; e->vfix[j] = val;
  }
}
static void add_flux(Elem *e, const FormBC *bc, int a, double A) /* petigaelem.c:1191-1212 */
{
  int j, k, dof = e->dof;
  for (k = 0; k < bc->count; k++) {
    int c = bc->field[k], idx = a*dof + c; double val = bc->value[k];
    if (c >= dof) continue;
    for (j = 0; j < e->nflux; j++) if (e->iflux[j] == idx) break;
    if (j == e->nflux) e->vflux[e->nflux++] = 0.0;
    e->iflux[j] = idx; e->vflux[j] += val*A;
  }
}
static void build_fix_side(Elem *e, int dir, int side) /* petigaelem.c:1214-1238 */
{
  const FormBC *bcv = &e->o->value[dir][side], *bcl = &e->o->load[dir][side];
  if (bcv->count || bcl->count) {
    double Area = bcl->count ? boundary_area(e,dir,side) : 1;
    int S[3] = {0,0,0}, E[3] = {1,1,1}, ia, ja, ka, i, jstride, kstride;
    for (i = 0; i < e->dim; i++) E[i] = e->nenA[i];
    jstride = E[0]; kstride = E[0]*E[1];
    if (side) S[dir] = E[dir]-1; else E[dir] = S[dir]+1;
    for (ka = S[2]; ka < E[2]; ka++) for (ja = S[1]; ja < E[1]; ja++) for (ia = S[0]; ia < E[0]; ia++) {
      int a = ia + ja*jstride + ka*kstride;
      add_fixa(e,bcv,a); add_flux(e,bcl,a,Area);
    }
  }
}
static void elem_build_fix(Elem *e) /* petigaelem.c:1263-1283 */
{
  int i;
  e->nfix = 0; e->nflux = 0;
  for (i = 0; i < e->dim; i++) {
    int w = e->o->axis[i].periodic, last = e->o->elem_sizes[i]-1;
    if (e->ID[i] == 0 && !w) build_fix_side(e,i,0);
    if (e->ID[i] == last && !w) build_fix_side(e,i,1);
  }
}

/* ------------------------------------------------------------------------------------------ */
/* forms (the user callbacks of the demos), evaluated at one quadrature point                  */
/* ------------------------------------------------------------------------------------------ */

typedef struct { /* what a callback reads from IGAPoint: include/petiga.h:644-703 */
  int nen, dof, dim, nsd;
  const double *N0, *N1, *N2;  /* shape[0..2] of this point */
  const double *x;             /* mapX[0] (or mapU[0] when no geometry) */
  int atboundary, boundary_id; const double *normal;   /* petiga.h:645-647,668 */
  const double *E1;            /* mapU[1] of this point ([dim][nsd]) or NULL without geometry (IGAPointFormInvGradGeomMap) */
  double L[3];                 /* IGAPointFormScale: detJac of the element per axis (petigapoint.c:209-223) */
  int maxdeg;                  /* max_i axis[i]->p (demo/NitscheMethod.c Degree()) */
} Point;

static double l2_function(int choice, int dim, const double x[3]) /* demo/L2Projection.c:3-61 */
{
  int i; double f = 0;
  switch (choice) {
  case 0: for (i=0;i<dim;i++) f += x[i]; return f;
  case 1: for (i=0;i<dim;i++) f += x[i]*x[i]; return f;
  case 2: for (i=0;i<dim;i++) f += x[i]*x[i]*x[i]; return f;
  case 3: for (i=0;i<dim;i++) f += x[i]*x[i]*x[i]*x[i]; return f;
  case 4: { double X = x[0], Y = x[1]; X = 2.5*X+1; Y = 2.0*Y+0;
            return exp(-X*X-Y*Y) + 0.5 * exp(-(X-2)*(X-2)-(Y-0.5)*(Y-0.5)); }
  case 5: { double X = x[0]*3, Y = x[1]*3;
            return 3 * pow(1-X,2) * exp(-pow(X,2) - pow(Y+1,2))
                   - 10 * (X/5 - pow(X,3) - pow(Y,5)) * exp(-pow(X,2) - pow(Y,2))
                   - 1.0/3 * exp(-pow(X+1,2) - pow(Y,2)); }
  case 6: f = 1; for (i=0;i<dim;i++) f *= sin(M_PI*x[i]); return f;
  case 7: for (i=0;i<dim;i++) f += (x[i] < 0.0) ? -1.0 : +1.0; return f;
  }
  return 0;
}

static void get_value(const Point *p, const double *U, double *u) /* petigaval.F90:182-196 */
{ int a, c; for (c=0;c<p->dof;c++) u[c]=0; for (a=0;a<p->nen;a++) for (c=0;c<p->dof;c++) u[c] = u[c] + p->N0[a]*U[a*p->dof+c]; }
static void get_grad(const Point *p, const double *U, double *u) /* petigaval.F90:198-214: V(dim,dof) */
{ int a, c, i, d = p->nsd; for (c=0;c<p->dof*d;c++) u[c]=0;
  for (a=0;a<p->nen;a++) for (c=0;c<p->dof;c++) for (i=0;i<d;i++) u[c*d+i] = u[c*d+i] + p->N1[a*d+i]*U[a*p->dof+c]; }
static void get_del2(const Point *p, const double *U, double *u) /* petigaval.F90:234-251 */
{ int a, c, i, d = p->nsd; for (c=0;c<p->dof;c++) u[c]=0;
  for (a=0;a<p->nen;a++) for (c=0;c<p->dof;c++) for (i=0;i<d;i++) u[c] = u[c] + p->N2[a*d*d+i*d+i]*U[a*p->dof+c]; }

/* K is [nen][dof][nen][dof] row-major, F is [nen][dof]; both pre-zeroed (petigapoint.c:414-449) */
static int form_system(int form, const double *prm, const Point *p, double *K, double *F)
{
  int a, b, i, j, nen = p->nen, dim = p->nsd;
  switch (form) {
  case FORM_POISSON:   /* demo/Poisson{1,2,3}D.c:3-23 */
  case FORM_LAPLACE:   /* demo/Laplace.c:35-48 */
    for (a = 0; a < nen; a++) {
      for (b = 0; b < nen; b++) { double s = 0.0;
        if (form == FORM_POISSON && dim == 3) s = p->N1[a*3]*p->N1[b*3] + p->N1[a*3+1]*p->N1[b*3+1] + p->N1[a*3+2]*p->N1[b*3+2];
        else if (form == FORM_POISSON && dim == 2) s = p->N1[a*2]*p->N1[b*2] + p->N1[a*2+1]*p->N1[b*2+1];
        else for (i = 0; i < dim; i++) s += p->N1[a*dim+i]*p->N1[b*dim+i];
        K[a*nen+b] = s; }
      F[a] = (form == FORM_POISSON) ? p->N0[a] * 1.0 : 0.0;
    }
    return 0;
  case FORM_L2PROJECTION: { /* demo/L2Projection.c:67-88 */
    double xyz[3] = {0,0,0}, f; for (i = 0; i < dim; i++) xyz[i] = p->x[i];
    f = l2_function((int)prm[0], p->dim, xyz);
    for (a = 0; a < nen; a++) { for (b = 0; b < nen; b++) K[a*nen+b] = p->N0[a]*p->N0[b]; F[a] = p->N0[a]*f; }
    return 0; }
  case FORM_MASS: { /* test/IGACreate.c:10-64 (Vector, Matrix, System): block mass, F = N_a */
    int dof = p->dof;
    for (a = 0; a < nen; a++) {
      for (b = 0; b < nen; b++) for (i = 0; i < dof; i++) for (j = 0; j < dof; j++)
        if (i == j) K[a*dof*nen*dof+i*nen*dof+b*dof+j] = p->N0[a]*p->N0[b];
      for (i = 0; i < dof; i++) F[a*dof+i] = p->N0[a] * 1;
    }
    return 0; }
  case FORM_ELASTICITY3D: { /* demo/Elasticity3D.c:13-46 -- literal, including the extra *mu at :37 */
    double lambda = prm[0], mu = prm[1];
#define KL(a,i,b,j) K[(((a)*3+(i))*nen+(b))*3+(j)]
    for (a = 0; a < nen; a++) {
      double Na_x = p->N1[a*3], Na_y = p->N1[a*3+1], Na_z = p->N1[a*3+2];
      for (b = 0; b < nen; b++) {
        double Nb_x = p->N1[b*3], Nb_y = p->N1[b*3+1], Nb_z = p->N1[b*3+2];
        KL(a,0,b,0) = Na_x*Nb_x*(lambda + 2*mu) + mu*(Na_y*Nb_y + Na_z*Nb_z);
        KL(a,0,b,1) = Na_x*Nb_y*lambda + Na_y*Nb_x*mu;
        KL(a,0,b,2) = Na_x*Nb_z*lambda + Na_z*Nb_x*mu;
        KL(a,1,b,0) = Na_x*Nb_y*mu + Na_y*Nb_x*lambda;
        KL(a,1,b,1) = Na_y*Nb_y*(lambda + 2*mu) + mu*(Na_z*Nb_z + Na_x*Nb_x*mu);
        KL(a,1,b,2) = Na_y*Nb_z*lambda + Na_z*Nb_y*mu;
        KL(a,2,b,0) = Na_x*Nb_z*mu + Na_z*Nb_x*lambda;
        KL(a,2,b,1) = Na_y*Nb_z*mu + Na_z*Nb_y*lambda;
        KL(a,2,b,2) = mu*(Na_x*Nb_x + Na_y*Nb_y) + Na_z*Nb_z*(lambda + 2*mu);
      }
      F[a] = 0.0;
    }
#undef KL
    return 0; }
  case FORM_BOUNDARYINTEGRAL: /* demo/BoundaryIntegral.c:27-57: Laplace in the interior, F = N*1.0 on the visited faces */
    if (!p->atboundary) {
      for (a = 0; a < nen; a++) { for (b = 0; b < nen; b++) { double sum = 0.0; for (i = 0; i < dim; i++) sum += p->N1[a*dim+i]*p->N1[b*dim+i]; K[a*nen+b] = sum; } F[a] = 0.0; }
    } else for (a = 0; a < nen; a++) F[a] = p->N0[a] * 1.0;
    return 0;
  case FORM_NEUMANN: { /* demo/Neumann.c:5-45 SystemGalerkin: f = 4 pi^2 (sin 2 pi x + sin 2 pi y + sin 2 pi z) */
    double xx[3] = {0,0,0}, f; for (i = 0; i < dim; i++) xx[i] = p->x[i];
    f = 4*M_PI*M_PI * (sin(2*M_PI*xx[0]) + sin(2*M_PI*xx[1]) + sin(2*M_PI*xx[2]));
    for (a = 0; a < nen; a++) { for (b = 0; b < nen; b++) { double sum = 0.0; for (i = 0; i < dim; i++) sum += p->N1[a*dim+i]*p->N1[b*dim+i]; K[a*nen+b] = sum; } F[a] = p->N0[a]*f; }
    return 0; }
  case FORM_CONVTEST: { /* test/ConvTest.c:30-69 Galerkin: c N_a N_b + k grad N_a . grad N_b, f = (c + k dim pi^2) prod sin(pi x_i) */
    double c = prm[0], k = prm[1], f = c + k*dim*M_PI*M_PI;
    for (i = 0; i < dim; i++) f *= sin(M_PI*p->x[i]);
    for (a = 0; a < nen; a++) { for (b = 0; b < nen; b++) { double sum = 0.0; for (i = 0; i < dim; i++) sum += p->N1[a*dim+i]*p->N1[b*dim+i];
      K[a*nen+b] = c*p->N0[a]*p->N0[b] + k*sum; } F[a] = p->N0[a]*f; }
    return 0; }
  case FORM_NITSCHE: { /* demo/NitscheMethod.c:70-119: Poisson with f = -2 dim inside; Nitsche terms on the visited faces */
    if (!p->atboundary) {
      double f = -2.0*dim;
      for (a = 0; a < nen; a++) { for (b = 0; b < nen; b++) { double sum = 0.0; for (i = 0; i < dim; i++) sum += p->N1[a*dim+i]*p->N1[b*dim+i]; K[a*nen+b] = sum; } F[a] = p->N0[a]*f; }
    } else {
      double g = 0.0, G[3][3], Nn[3], nn = 0.0, h, alpha; const double *n = p->normal;
      for (i = 0; i < dim; i++) g += p->x[i]*p->x[i];
      for (i = 0; i < dim; i++) for (j = 0; j < dim; j++)           /* IGAPointFormInvGradGeomMap (petigapoint.c:269-294) */
        G[i][j] = (p->E1 ? p->E1[i*dim+j] : (i == j ? 1.0 : 0.0)) / p->L[i];
      for (i = 0; i < dim; i++) { Nn[i] = 0.0; for (j = 0; j < dim; j++) Nn[i] += G[i][j]*n[j]; nn += Nn[i]*Nn[i]; }
      h = 2/sqrt(nn);                                               /* NormalMeshSize :58-67 */
      alpha = 5*(p->maxdeg+1)/h;
      for (a = 0; a < nen; a++) {
        double dna = 0.0; for (i = 0; i < dim; i++) dna += p->N1[a*dim+i]*n[i];
        for (b = 0; b < nen; b++) {
          double dnb = 0.0; for (i = 0; i < dim; i++) dnb += p->N1[b*dim+i]*n[i];
          K[a*nen+b] += - p->N0[a] * dnb;
          K[a*nen+b] += - p->N0[b] * dna;
          K[a*nen+b] += + alpha * p->N0[a]*p->N0[b];
        }
        F[a] += - dna*g;
        F[a] += + alpha * p->N0[a]*g;
      }
    }
    return 0; }
  case FORM_ELASTICITY: { /* demo/Elasticity.c:22-52; dof == dim */
    double lambda = prm[0], mu = prm[1];
    for (a = 0; a < nen; a++) for (b = 0; b < nen; b++) {
      double Kabii = 0.0; for (i = 0; i < dim; i++) Kabii += p->N1[a*dim+i]*p->N1[b*dim+i];
      for (i = 0; i < dim; i++) K[((a*dim+i)*nen+b)*dim+i] += mu * Kabii;
      for (i = 0; i < dim; i++) for (j = 0; j < dim; j++)
        K[((a*dim+i)*nen+b)*dim+j] += lambda * p->N1[a*dim+i]*p->N1[b*dim+j] + mu * p->N1[a*dim+j]*p->N1[b*dim+i];
    }
    for (a = 0; a < nen; a++) for (i = 0; i < dim; i++) F[a*dim+i] = p->N0[a] * 0.0;
    return 0; }
  }
  return 1;
}

/* residual-type forms: slot FUNCTION / IFUNCTION */
static int form_function(int form, const double *prm, const Point *p, double shift, const double *V, double t,
                         const double *U, double *R)
{
  int a, i, nen = p->nen, dim = p->nsd; (void)shift; (void)t;
  switch (form) {
  case FORM_CAHNHILLIARD2D: { /* demo/CahnHilliard2D.c:84-133 (Residual), :9-32 */
    double theta = prm[0], alpha = prm[1];
    double c_t, c, M, dM, dmu, c1[2], del2_c, c_x, c_y, t1;
    get_value(p,V,&c_t); get_value(p,U,&c);
    M = c*(1-c); dM = 1-2*c;
    dmu = 0.5/theta*1/(c*(1-c)) - 2; dmu *= 3*alpha;
    get_grad(p,U,c1); get_del2(p,U,&del2_c);
    c_x = c1[0]; c_y = c1[1];
    t1 = M*dmu + dM*del2_c;
    for (a = 0; a < nen; a++) {
      double Na = p->N0[a], Na_x = p->N1[a*2], Na_y = p->N1[a*2+1], Na_xx = p->N2[a*4+0], Na_yy = p->N2[a*4+3], Ra = 0;
      Ra += Na * c_t;
      Ra += (Na_x * c_x + Na_y * c_y) * t1;
      Ra += (Na_xx+Na_yy) * M * del2_c;
      R[a] = Ra;
    }
    return 0; }
  case FORM_CAHNHILLIARD3D: { /* demo/CahnHilliard3D.c:54-107 (Residual), :11-52 */
    double theta = prm[0], L0 = prm[1], lambda = prm[2];
    double c_t, c, M, dM, dmu, c1[3], c2[9], c_x, c_y, c_z, c_xx, c_yy, c_zz, t1; int cc;
    get_value(p,V,&c_t); get_value(p,U,&c);
    M = c*(1-c); dM = 1-2*c;
    dmu = 0.5/theta*1.0/(c*(1-c)) - 2; dmu *= L0*L0/lambda;
    get_grad(p,U,c1);
    for (cc = 0; cc < 9; cc++) c2[cc] = 0;                     /* IGAPointFormHess -> IGA_GetHess (petigaval.F90:216-232) */
    for (a = 0; a < nen; a++) for (cc = 0; cc < 9; cc++) c2[cc] = c2[cc] + p->N2[a*9+cc]*U[a];
    c_x = c1[0]; c_y = c1[1]; c_z = c1[2]; c_xx = c2[0]; c_yy = c2[4]; c_zz = c2[8];
    for (a = 0; a < nen; a++) {
      double Na = p->N0[a], Na_x = p->N1[a*3], Na_y = p->N1[a*3+1], Na_z = p->N1[a*3+2];
      double Na_xx = p->N2[a*9+0], Na_yy = p->N2[a*9+4], Na_zz = p->N2[a*9+8], Ra = 0;
      Ra += Na * c_t;
      t1 = M*dmu + dM*(c_xx+c_yy+c_zz);
      Ra += Na_x * t1 * c_x;
      Ra += Na_y * t1 * c_y;
      Ra += Na_z * t1 * c_z;
      Ra += (Na_xx+Na_yy+Na_zz) * M * (c_xx+c_yy+c_zz);
      R[a] = Ra;
    }
    return 0; }
  case FORM_BRATU: { /* demo/BratuFJ.F90: Function (:22-62), IFunction (:118-150) */
    double lambda = prm[0], u, v = 0, gu[3];
    get_value(p,U,&u); get_grad(p,U,gu);
    if (V) get_value(p,V,&v);
    for (a = 0; a < nen; a++) {
      double dot = 0; for (i = 0; i < dim; i++) dot += p->N1[a*dim+i]*gu[i];
      R[a] = (V ? p->N0[a]*v : 0.0) + dot - p->N0[a] * lambda * exp(u);
    }
    return 0; }
  case FORM_SNES2D: { /* test/Test_SNES_2D.c:12-46 Function: L2 projection of Peaks, Poisson, reaction-diffusion, Bratu (dof 4, dim 2) */
    double u0[4], u1[8], X, Y, peaks;
    get_value(p,U,u0); get_grad(p,U,u1);
    X = p->x[0]*3; Y = p->x[1]*3;
    peaks = 3 * pow(1-X,2) * exp(-pow(X,2) - pow(Y+1,2)) - 10 * (X/5 - pow(X,3) - pow(Y,5)) * exp(-pow(X,2) - pow(Y,2))
            - 1.0/3 * exp(-pow(X+1,2) - pow(Y,2));
    for (a = 0; a < nen; a++) {
      double Na = p->N0[a], Na_x = p->N1[a*2], Na_y = p->N1[a*2+1];
      R[a*4+0] = Na*u0[0] - Na * peaks;
      R[a*4+1] = Na_x*u1[2] + Na_y*u1[3] - Na * 1.0;
      R[a*4+2] = Na*u0[2] + Na_x*u1[4] + Na_y*u1[5] - Na * 1.0;
      R[a*4+3] = Na_x*u1[6] + Na_y*u1[7] - Na * 1.0*exp(u0[3]);
    }
    return 0; }
  case FORM_POISSON: { /* residual of demo/Poisson: R_a = grad N_a . grad u - N_a*1 (linear problem as SNES) */
    double gu[3]; get_grad(p,U,gu);
    for (a = 0; a < nen; a++) { double dot = 0; for (i = 0; i < dim; i++) dot += p->N1[a*dim+i]*gu[i]; R[a] = dot - p->N0[a]*1.0; }
    return 0; }
  }
  return 1;
}

/* tangent-type forms: slot JACOBIAN / IJACOBIAN */
static int form_jacobian(int form, const double *prm, const Point *p, double shift, const double *V, double t,
                         const double *U, double *K, int transient)
{
  int a, b, i, nen = p->nen, dim = p->nsd; (void)t;
  switch (form) {
  case FORM_CAHNHILLIARD2D: { /* demo/CahnHilliard2D.c:135-197 (Tangent) */
    double theta = prm[0], alpha = prm[1];
    double c_t, c, M, dM, d2M, dmu, d2mu, c1[2], del2_c, c_x, c_y, t1, t2;
    get_value(p,V,&c_t); get_value(p,U,&c); (void)c_t;
    M = c*(1-c); dM = 1-2*c; d2M = -2;
    dmu = 0.5/theta*1/(c*(1-c)) - 2; dmu *= 3*alpha;
    d2mu = 0.5/theta*(2*c-1)/(c*c*(1-c)*(1-c)); d2mu *= 3*alpha;
    get_grad(p,U,c1); get_del2(p,U,&del2_c);
    c_x = c1[0]; c_y = c1[1];
    t1 = M*dmu + dM*del2_c;
    t2 = (dM*dmu+M*d2mu+d2M*del2_c);
    for (a = 0; a < nen; a++) {
      double Na = p->N0[a], Na_x = p->N1[a*2], Na_y = p->N1[a*2+1], del2_Na = p->N2[a*4+0] + p->N2[a*4+3];
      for (b = 0; b < nen; b++) {
        double Nb = p->N0[b], Nb_x = p->N1[b*2], Nb_y = p->N1[b*2+1], del2_Nb = p->N2[b*4+0] + p->N2[b*4+3];
        double Kab = 0, t3;
        Kab += shift*Na*Nb;
        Kab += (Na_x * Nb_x + Na_y * Nb_y) * t1;
        t3 = t2*Nb + dM*del2_Nb;
        Kab += (Na_x * c_x + Na_y * c_y) * t3;
        Kab += del2_Na * (dM*del2_c*Nb + M*del2_Nb);
        K[a*nen+b] = Kab;
      }
    }
    return 0; }
  case FORM_CAHNHILLIARD3D: { /* demo/CahnHilliard3D.c:109-169 (Tangent) */
    double theta = prm[0], L0 = prm[1], lambda = prm[2];
    double c, M, dM, d2M, dmu, d2mu, c1[3], c2[9], c_x, c_y, c_z, c_xx, c_yy, c_zz; int cc;
    get_value(p,U,&c);
    M = c*(1-c); dM = 1-2*c; d2M = -2;
    dmu = 0.5/theta*1.0/(c*(1-c)) - 2; dmu *= L0*L0/lambda;
    d2mu = -0.5/theta*(1-2*c)/(c*c*(1-c)*(1-c)); d2mu *= L0*L0/lambda;
    get_grad(p,U,c1);
    for (cc = 0; cc < 9; cc++) c2[cc] = 0;
    for (a = 0; a < nen; a++) for (cc = 0; cc < 9; cc++) c2[cc] = c2[cc] + p->N2[a*9+cc]*U[a];
    c_x = c1[0]; c_y = c1[1]; c_z = c1[2]; c_xx = c2[0]; c_yy = c2[4]; c_zz = c2[8];
    for (a = 0; a < nen; a++) {
      double Na = p->N0[a], Na_x = p->N1[a*3], Na_y = p->N1[a*3+1], Na_z = p->N1[a*3+2];
      double Na_xx = p->N2[a*9+0], Na_yy = p->N2[a*9+4], Na_zz = p->N2[a*9+8];
      for (b = 0; b < nen; b++) {
        double Nb = p->N0[b], Nb_x = p->N1[b*3], Nb_y = p->N1[b*3+1], Nb_z = p->N1[b*3+2];
        double Nb_xx = p->N2[b*9+0], Nb_yy = p->N2[b*9+4], Nb_zz = p->N2[b*9+8];
        double Kab = 0, t1, t2;
        Kab += shift*Na*Nb;
        t1 = M*dmu + dM*(c_xx+c_yy+c_zz);
        Kab += Na_x * t1 * Nb_x;
        Kab += Na_y * t1 * Nb_y;
        Kab += Na_z * t1 * Nb_z;
        t2 = (dM*dmu+M*d2mu+d2M*(c_xx+c_yy+c_zz))*Nb + dM*(Nb_xx+Nb_yy+Nb_zz);
        Kab += Na_x * t2 * c_x;
        Kab += Na_y * t2 * c_y;
        Kab += Na_z * t2 * c_z;
        Kab += (Na_xx+Na_yy+Na_zz) * (dM*(c_xx+c_yy+c_zz)*Nb + M*(Nb_xx+Nb_yy+Nb_zz));
        K[a*nen+b] = Kab;
      }
    }
    return 0; }
  case FORM_BRATU: { /* demo/BratuFJ.F90: Jacobian (:64-114), IJacobian */
    double lambda = prm[0], u; get_value(p,U,&u);
    for (a = 0; a < nen; a++) for (b = 0; b < nen; b++) {
      double dot = 0; for (i = 0; i < dim; i++) dot += p->N1[a*dim+i]*p->N1[b*dim+i];
      K[a*nen+b] = (transient ? shift*p->N0[a]*p->N0[b] : 0.0) + dot - p->N0[a]*p->N0[b]*lambda*exp(u);
    }
    return 0; }
  case FORM_POISSON:
    for (a = 0; a < nen; a++) for (b = 0; b < nen; b++) {
      double dot = 0; for (i = 0; i < dim; i++) dot += p->N1[a*dim+i]*p->N1[b*dim+i]; K[a*nen+b] = dot; }
    return 0;
  case FORM_SNES2D: { /* test/Test_SNES_2D.c:48-72 Jacobian */
    double u0[4], r;
    get_value(p,U,u0); r = u0[3];
    for (a = 0; a < nen; a++) {
      double Na = p->N0[a], Na_x = p->N1[a*2], Na_y = p->N1[a*2+1];
      for (b = 0; b < nen; b++) {
        double Nb = p->N0[b], Nb_x = p->N1[b*2], Nb_y = p->N1[b*2+1];
        K[a*nen*16+0*nen*4+b*4+0] = Na*Nb;
        K[a*nen*16+1*nen*4+b*4+1] = Na_x*Nb_x + Na_y*Nb_y;
        K[a*nen*16+2*nen*4+b*4+2] = Na*Nb + Na_x*Nb_x + Na_y*Nb_y;
        K[a*nen*16+3*nen*4+b*4+3] = Na_x*Nb_x + Na_y*Nb_y - Na*Nb * 1.0*exp(r);
      }
    }
    return 0; }
  }
  return 1;
}

/* callbacks of the IE, RHS and I2 drivers (include/petiga.h:172-197): W is the third vector (U0 for IE, A for I2) */
static int form_ext_function(int form, int slot, const double *prm, const Point *p, double a1, const double *V, double t,
                             const double *U, double a2, const double *W, double t0, double *R)
{
  int a, i, nen = p->nen, dim = p->nsd; (void)a1; (void)a2; (void)t; (void)t0;
  switch (form) {
  case FORM_PATTERNFORMATION: { /* demo/PatternFormation.c:26-77 IEFunction; prm = {IMPLICIT, delta, D1, D2, alpha, beta, gamma, tau1, tau2} */
    int IMPLICIT = prm[0] != 0.0;
    double delta = prm[1], D1 = prm[2], D2 = prm[3], alpha = prm[4], beta = prm[5], gamma = prm[6], tau1 = prm[7], tau2 = prm[8];
    double uv_t[2], uv_0[2], uv_1[4], u_t, v_t, u, v, u_x, v_x, u_y, v_y, f, g;
    if (slot != SLOT_IEFUNCTION || dim != 2 || p->dof != 2) return 1;
    get_value(p,V,uv_t);
    if (IMPLICIT) get_value(p,U,uv_0); else get_value(p,W,uv_0);
    get_grad(p,U,uv_1);
    u_t = uv_t[0]; v_t = uv_t[1]; u = uv_0[0]; v = uv_0[1];
    u_x = uv_1[0]; v_x = uv_1[2]; u_y = uv_1[1]; v_y = uv_1[3];
    f = alpha*u*(1-tau1*v*v) + v*(1-tau2*u);
    g = beta*v*(1+alpha*tau1/beta*u*v) + u*(gamma+tau2*v);
    for (a = 0; a < nen; a++) {
      double Na = p->N0[a], Na_x = p->N1[a*2], Na_y = p->N1[a*2+1];
      R[a*2+0] = Na*u_t + delta*D1*(Na_x*u_x + Na_y*u_y) - Na*f;
      R[a*2+1] = Na*v_t + delta*D2*(Na_x*v_x + Na_y*v_y) - Na*g;
    }
    return 0; }
  case FORM_ELASTICROD: { /* demo/ElasticRodFJ.F90:19-55 I2Function: F = N rho A + E grad N . grad U; prm = {rho, E} */
    double rho = prm[0], E = prm[1], A1, gu[3];
    if (slot != SLOT_I2FUNCTION || p->dof != 1) return 1;
    get_value(p,W,&A1); get_grad(p,U,gu);
    for (a = 0; a < nen; a++) { double dot = 0; for (i = 0; i < dim; i++) dot += p->N1[a*dim+i]*gu[i]; R[a] = p->N0[a]*rho*A1 + E*dot; }
    return 0; }
  case FORM_BRATU: { /* explicit form of the transient Bratu problem u_t = lap u + lambda exp(u) (demo/BratuFJ.F90 terms moved to the
                        right-hand side): G_a = -grad N_a . grad u + N_a lambda exp(u).  No reference demo registers an RHSFunction;
                        this form exercises IGAComputeRHSFunction (src/petigats.c:357-416) */
    double lambda = prm[0], u, gu[3];
    if (slot != SLOT_RHSFUNCTION || p->dof != 1) return 1;
    get_value(p,U,&u); get_grad(p,U,gu);
    for (a = 0; a < nen; a++) { double dot = 0; for (i = 0; i < dim; i++) dot += p->N1[a*dim+i]*gu[i]; R[a] = -dot + p->N0[a]*lambda*exp(u); }
    return 0; }
  }
  return 1;
}
static int form_ext_jacobian(int form, int slot, const double *prm, const Point *p, double a1, const double *V, double t,
                             const double *U, double a2, const double *W, double t0, double *K)
{
  int a, b, i, j, nen = p->nen, dim = p->nsd; (void)V; (void)W; (void)a2; (void)t; (void)t0;
  switch (form) {
  case FORM_PATTERNFORMATION: { /* demo/PatternFormation.c:79-141 IEJacobian */
    int IMPLICIT = prm[0] != 0.0;
    double delta = prm[1], D1 = prm[2], D2 = prm[3], alpha = prm[4], beta = prm[5], gamma = prm[6], tau1 = prm[7], tau2 = prm[8];
    double f_u = 0, f_v = 0, g_u = 0, g_v = 0, shift = a1;
    if (slot != SLOT_IEJACOBIAN || dim != 2 || p->dof != 2) return 1;
    if (IMPLICIT) {
      double uv_0[2], u, v; get_value(p,U,uv_0); u = uv_0[0]; v = uv_0[1];
      f_u = alpha*(1-tau1*v*v) - tau2*v;
      f_v = -2*alpha*tau1*u*v + (1-tau2*u);
      g_u = alpha*tau1*v*v + (gamma+tau2*v);
      g_v = (beta+2*alpha*tau1*u*v) + tau2*u;
    }
    for (a = 0; a < nen; a++) {
      double Na = p->N0[a], Na_x = p->N1[a*2], Na_y = p->N1[a*2+1];
      for (b = 0; b < nen; b++) {
        double Nb = p->N0[b], Nb_x = p->N1[b*2], Nb_y = p->N1[b*2+1], Kab[2][2] = {{0,0},{0,0}};
        Kab[0][0] = shift*Na*Nb + delta*D1*(Na_x*Nb_x + Na_y*Nb_y);
        Kab[1][1] = shift*Na*Nb + delta*D2*(Na_x*Nb_x + Na_y*Nb_y);
        if (IMPLICIT) {
          Kab[0][0] -= Na*f_u*Nb; Kab[0][1] -= Na*f_v*Nb;
          Kab[1][0] -= Na*g_u*Nb; Kab[1][1] -= Na*g_v*Nb;
          for (i = 0; i < 2; i++) for (j = 0; j < 2; j++) K[((a*2+i)*nen+b)*2+j] += Kab[i][j];
        } else {
          K[((a*2+0)*nen+b)*2+0] += Kab[0][0];
          K[((a*2+1)*nen+b)*2+1] += Kab[1][1];
        }
      }
    }
    return 0; }
  case FORM_ELASTICROD: { /* demo/ElasticRodFJ.F90:57-95 I2Jacobian: shiftA rho N_a N_b + E grad N_a . grad N_b */
    double rho = prm[0], E = prm[1], shiftA = a1;
    if (slot != SLOT_I2JACOBIAN || p->dof != 1) return 1;
    for (a = 0; a < nen; a++) for (b = 0; b < nen; b++) {
      double dot = 0; for (i = 0; i < dim; i++) dot += p->N1[a*dim+i]*p->N1[b*dim+i];
      K[a*nen+b] = shiftA*rho*p->N0[a]*p->N0[b] + E*dot;
    }
    return 0; }
  case FORM_BRATU: { /* derivative of the RHSFunction above */
    double lambda = prm[0], u;
    if (slot != SLOT_RHSJACOBIAN || p->dof != 1) return 1;
    get_value(p,U,&u);
    for (a = 0; a < nen; a++) for (b = 0; b < nen; b++) {
      double dot = 0; for (i = 0; i < dim; i++) dot += p->N1[a*dim+i]*p->N1[b*dim+i];
      K[a*nen+b] = -dot + p->N0[a]*lambda*exp(u)*p->N0[b];
    }
    return 0; }
  }
  return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* drivers: src/petigaksp.c:33-202, src/petigasnes.c:23-139, src/petigats.c:23-159             */
/* ------------------------------------------------------------------------------------------ */

typedef struct { int nrows; int *rowptr, *colidx; long nnz; int *rank_rowstart; } GlobalPattern;

/* concatenate the per-rank patterns: global block CSR in PETSc numbering */
static int global_pattern(OIGA *o, int size, GlobalPattern *gp)
{
  int r, ntot = 0; long nnz = 0;
  int **rp = (int**)calloc((size_t)size, sizeof(int*)), **ci = (int**)calloc((size_t)size, sizeof(int*));
  int *nr = (int*)calloc((size_t)size, sizeof(int)); long *nz = (long*)calloc((size_t)size, sizeof(long));
  gp->rank_rowstart = (int*)malloc((size_t)(size+1)*sizeof(int));
  for (r = 0; r < size; r++) {
    if (setup_rank(o, size, r)) return 1;
    nz[r] = pattern_rank(o, &rp[r], &ci[r], &nr[r]);
    gp->rank_rowstart[r] = ntot; ntot += nr[r]; nnz += nz[r];
  }
  gp->rank_rowstart[size] = ntot;
  gp->nrows = ntot; gp->nnz = nnz;
  gp->rowptr = (int*)malloc((size_t)(ntot+1)*sizeof(int)); gp->colidx = (int*)malloc((size_t)nnz*sizeof(int));
  { int row = 0; long off = 0, i;
    gp->rowptr[0] = 0;
    for (r = 0; r < size; r++) {
      for (i = 0; i < nr[r]; i++) gp->rowptr[++row] = (int)(off + rp[r][i+1]);
      memcpy(gp->colidx + off, ci[r], (size_t)nz[r]*sizeof(int)); off += nz[r];
      free(rp[r]); free(ci[r]);
    } }
  free(rp); free(ci); free(nr); free(nz);
  return 0;
}

static long csr_find(const GlobalPattern *gp, int row, int col) /* the sorted-row search of MatSetValues */
{
  int lo = gp->rowptr[row], hi = gp->rowptr[row+1]-1;
  while (lo <= hi) { int mid = (lo+hi)/2, c = gp->colidx[mid]; if (c == col) return mid; if (c < col) lo = mid+1; else hi = mid-1; }
  return -1;
}

/* One full assembly over `size` emulated ranks.
   values: [nnz_blocks][dof][dof] (row-major blocks), rhs: [nrows][dof]; either may be NULL per slot.
   Ug/Vg: global vectors (PETSc ordering) or NULL.  Returns 0, or >0 on error (e.g. entry outside pattern). */
/* PETSc's stash (MatSetValues/VecSetValues on rows owned by another rank, shipped and added in Mat/VecAssemblyEnd,
   src/petigaksp.c:197-200): entries as (flat index, value); vector entries carry index -(i)-1 */
typedef struct { long n, cap; long *idx; double *val; int row0, row1; } Stash;
static void stash_push(Stash *st, long idx, double v)
{
  if (st->n == st->cap) { st->cap = st->cap ? 2*st->cap : 4096; st->idx = (long*)realloc(st->idx, sizeof(long)*(size_t)st->cap);
                          st->val = (double*)realloc(st->val, sizeof(double)*(size_t)st->cap); }
  st->idx[st->n] = idx; st->val[st->n] = v; st->n++;
}

/* idx0/idx1: element index range of the rank's loop (idx1 < 0: all of it); zero: MatZeroEntries/VecZeroEntries first;
   st: NULL, or the stash that receives contributions to rows outside [st->row0, st->row1) */
static int assemble_ex(OIGA *o, int size, int r0, int r1, int idx0, int idx1, int zero, Stash *st, int slot, int form, const double *prm,
                       double shift, const double *Vg, double t, const double *Ug, const GlobalPattern *gp, double *values, double *rhs)
{
  int r, dof = o->dof, err = 0;
  int want_mat = (slot == SLOT_MATRIX || slot == SLOT_SYSTEM || slot == SLOT_JACOBIAN || slot == SLOT_IJACOBIAN ||
                  slot == SLOT_IEJACOBIAN || slot == SLOT_RHSJACOBIAN || slot == SLOT_I2JACOBIAN);
  int want_vec = (slot == SLOT_VECTOR || slot == SLOT_SYSTEM || slot == SLOT_FUNCTION || slot == SLOT_IFUNCTION ||
                  slot == SLOT_IEFUNCTION || slot == SLOT_RHSFUNCTION || slot == SLOT_I2FUNCTION);
  int state = (slot >= SLOT_FUNCTION);
  int third = (slot == SLOT_IEFUNCTION || slot == SLOT_IEJACOBIAN || slot == SLOT_I2FUNCTION || slot == SLOT_I2JACOBIAN);
  int transient = (slot == SLOT_IFUNCTION || slot == SLOT_IJACOBIAN || third);   /* a V vector is gathered and DelValues'ed */
  int i2 = (slot == SLOT_I2FUNCTION || slot == SLOT_I2JACOBIAN);
  int fixsys = (slot == SLOT_SYSTEM), fixfun = (want_vec && state), fixjac = (want_mat && state);
  const double *Wg = o->aux_W; double shift2 = o->aux_shift2, t0 = o->aux_t0; int maxdeg = 0;
  { int dd; for (dd = 0; dd < o->dim; dd++) if (o->axis[dd].p > maxdeg) maxdeg = o->axis[dd].p; }
  if (third && !Wg) return 2;
  if (want_mat && zero) memset(values, 0, sizeof(double)*(size_t)gp->nnz*dof*dof);   /* MatZeroEntries */
  if (want_vec && zero) memset(rhs, 0, sizeof(double)*(size_t)gp->nrows*dof);        /* VecZeroEntries */
  for (r = r0; r < r1 && !err; r++) {
    Elem e; int index, count, N, ng; double *A, *B, *K, *F, *U = NULL, *V = NULL, *W = NULL, *arrayU = NULL, *arrayV = NULL, *arrayW = NULL;
    if (setup_rank(o, size, r)) return 1;
    elem_alloc(&e, o);
    N = e.nen*dof; ng = o->node_gwidth[0]*o->node_gwidth[1]*o->node_gwidth[2];
    A = (double*)malloc(sizeof(double)*(size_t)N*N); K = (double*)malloc(sizeof(double)*(size_t)N*N);
    B = (double*)malloc(sizeof(double)*(size_t)N);   F = (double*)malloc(sizeof(double)*(size_t)N);
    if (state) { /* IGAGetLocalVecArray: G2L scatter (petigavec.c:256-269) */
      int a, c;
      U = (double*)malloc(sizeof(double)*(size_t)N); arrayU = (double*)malloc(sizeof(double)*(size_t)ng*dof);
      for (a = 0; a < ng; a++) for (c = 0; c < dof; c++) arrayU[(size_t)a*dof+c] = Ug[(size_t)o->lgmap[a]*dof+c];
      if (transient) { V = (double*)malloc(sizeof(double)*(size_t)N); arrayV = (double*)malloc(sizeof(double)*(size_t)ng*dof);
        for (a = 0; a < ng; a++) for (c = 0; c < dof; c++) arrayV[(size_t)a*dof+c] = Vg[(size_t)o->lgmap[a]*dof+c]; }
      if (third) { W = (double*)malloc(sizeof(double)*(size_t)N); arrayW = (double*)malloc(sizeof(double)*(size_t)ng*dof);
        for (a = 0; a < ng; a++) for (c = 0; c < dof; c++) arrayW[(size_t)a*dof+c] = Wg[(size_t)o->lgmap[a]*dof+c]; }
    }
    count = o->elem_width[0]*o->elem_width[1]*o->elem_width[2];
    if (idx1 >= 0 && idx1 < count) count = idx1;
    for (index = (idx1 >= 0 ? idx0 : 0); index < count && !err; index++) {   /* IGANextElement: petigaelem.c:375-410 */
      int i, q, a, b, idx = index, f;
      for (i = 0; i < 3; i++) { int coord = idx % o->elem_width[i]; idx = (idx - coord)/o->elem_width[i]; e.ID[i] = coord + o->elem_start[i]; }
      elem_closure(&e);
      elem_build_fix(&e);
      memset(A, 0, sizeof(double)*(size_t)N*N); memset(B, 0, sizeof(double)*(size_t)N);
      if (state) { /* GetValues / DelValues / FixValues: petigaelem.c:1074-1100,1327-1358 */
        for (a = 0; a < e.nen; a++) for (i = 0; i < dof; i++) U[a*dof+i] = arrayU[(size_t)e.mapping[a]*dof+i];
        if (transient) { for (a = 0; a < e.nen; a++) for (i = 0; i < dof; i++) V[a*dof+i] = arrayV[(size_t)e.mapping[a]*dof+i];
          for (f = 0; f < e.nfix; f++) V[e.ifix[f]] = 0.0; }
        if (third) { for (a = 0; a < e.nen; a++) for (i = 0; i < dof; i++) W[a*dof+i] = arrayW[(size_t)e.mapping[a]*dof+i];
          /* I2: DelValues(A) (petigats2.c:68); IE: FixValues(U0) (petigats.c:228) */
          for (f = 0; f < e.nfix; f++) W[e.ifix[f]] = i2 ? 0.0 : e.vfix[f]; }
        for (f = 0; f < e.nfix; f++) { e.ufix[f] = U[e.ifix[f]]; U[e.ifix[f]] = e.vfix[f]; }
      }
      { int pass;   /* IGAElementNextForm (petigaelem.c:427-447): visited boundary faces of this element first, then the interior */
      for (pass = 0; pass <= 2*e.dim && !err; pass++) {
      if (pass < 2*e.dim) {
        int bi = pass/2, bs = pass%2, be = bs ? o->elem_sizes[bi]-1 : 0;
        if (e.ID[bi] != be || !o->visit[bi][bs]) continue;
        e.atboundary = 1; e.baxis = bi; e.bside = bs;
      } else e.atboundary = 0;
      elem_tabulate(&e);
      for (q = 0; q < e.nqp; q++) {      /* quadrature loop + IGAPointAddArray (petigapoint.c:451-465) */
        Point p; double JW = e.detJac[q] * e.weight[q]; int ret = 0;
        p.nen = e.nen; p.dof = dof; p.dim = e.dim; p.nsd = e.nsd;
        p.atboundary = e.atboundary; p.boundary_id = e.atboundary ? 2*e.baxis + e.bside : -1; p.normal = e.normal + (size_t)q*e.nsd;
        p.N0 = e.shape[0] + (size_t)q*e.nen; p.N1 = e.shape[1] + (size_t)q*e.nen*e.nsd; p.N2 = e.shape[2] + (size_t)q*e.nen*e.nsd*e.nsd;
        p.x = e.geometry ? e.mapX[0] + (size_t)q*e.nsd : e.mapU[0] + (size_t)q*e.dim;
        p.E1 = e.geometry ? e.mapU[1] + (size_t)q*e.dim*e.nsd : NULL; p.maxdeg = maxdeg;
        { int dd; for (dd = 0; dd < 3; dd++) p.L[dd] = o->basis[dd].detJac[e.ID[dd]]; }
        if (want_mat) memset(K, 0, sizeof(double)*(size_t)N*N);
        memset(F, 0, sizeof(double)*(size_t)N);
        switch (slot) {
        case SLOT_SYSTEM: case SLOT_MATRIX: case SLOT_VECTOR: {
          double *Kq = want_mat ? K : (double*)calloc((size_t)N*N, sizeof(double));
          ret = form_system(form, prm, &p, Kq, F);
          if (!want_mat) free(Kq);
          break; }
        case SLOT_FUNCTION:  ret = form_function(form, prm, &p, 0.0, NULL, 0.0, U, F); break;
        case SLOT_IFUNCTION: ret = form_function(form, prm, &p, shift, V, t, U, F); break;
        case SLOT_JACOBIAN:  ret = form_jacobian(form, prm, &p, 0.0, NULL, 0.0, U, K, 0); break;
        case SLOT_IJACOBIAN: ret = form_jacobian(form, prm, &p, shift, V, t, U, K, 1); break;
        case SLOT_IEFUNCTION: case SLOT_I2FUNCTION: ret = form_ext_function(form, slot, prm, &p, shift, V, t, U, shift2, W, t0, F); break;
        case SLOT_RHSFUNCTION: ret = form_ext_function(form, slot, prm, &p, 0.0, NULL, t, U, 0.0, NULL, 0.0, F); break;
        case SLOT_IEJACOBIAN: case SLOT_I2JACOBIAN: ret = form_ext_jacobian(form, slot, prm, &p, shift, V, t, U, shift2, W, t0, K); break;
        case SLOT_RHSJACOBIAN: ret = form_ext_jacobian(form, slot, prm, &p, 0.0, NULL, t, U, 0.0, NULL, 0.0, K); break;
        }
        if (ret) { err = 10; break; }
        if (want_mat) for (i = 0; i < N*N; i++) A[i] += K[i] * JW;
        if (want_vec) for (i = 0; i < N; i++) B[i] += F[i] * JW;
      }
      } e.atboundary = 0; }
      if (err) break;
      /* fix-up: petigaelem.c:1360-1389 (System), :1441-1463 (Function), :1483-1501 (Jacobian) */
      if (fixsys) {
        for (f = 0; f < e.nflux; f++) B[e.iflux[f]] += e.vflux[f];
        for (f = 0; f < e.nfix; f++) {
          int k = e.ifix[f]; double v = e.vfix[f];
          for (i = 0; i < N; i++) B[i] -= A[i*N+k] * v;
          for (i = 0; i < N; i++) A[i*N+k] = 0.0;
          for (i = 0; i < N; i++) A[k*N+i] = 0.0;
          A[k*N+k] = 1.0; B[k] = v;
        }
      } else if (fixfun) {
        for (f = 0; f < e.nflux; f++) B[e.iflux[f]] -= e.vflux[f];
        for (f = 0; f < e.nfix; f++) B[e.ifix[f]] = e.ufix[f] - e.vfix[f];
      } else if (fixjac) {
        for (f = 0; f < e.nfix; f++) {
          int k = e.ifix[f];
          for (i = 0; i < N; i++) A[i*N+k] = 0.0;
          for (i = 0; i < N; i++) A[k*N+i] = 0.0;
          A[k*N+k] = 1.0;
        }
      }
      /* scatter: MatSetValues[Blocked]Local / VecSetValues[Blocked]Local with ADD_VALUES (petigaelem.c:1525-1559) */
      for (a = 0; a < e.nen; a++) {
        int grow = o->lgmap[e.mapping[a]], ii, jj;
        int off = st && (grow < st->row0 || grow >= st->row1);   /* row of another rank: goes to the stash */
        if (want_vec) for (ii = 0; ii < dof; ii++) {
          if (off) stash_push(st, -((long)grow*dof+ii)-1, B[a*dof+ii]); else rhs[(size_t)grow*dof+ii] += B[a*dof+ii]; }
        if (want_mat) for (b = 0; b < e.nen; b++) {
          int gcol = o->lgmap[e.mapping[b]]; long pos = csr_find(gp, grow, gcol);
          if (pos < 0) { err = 20; break; }            /* MAT_NEW_NONZERO_LOCATION_ERR (petigamat.c:538) */
          for (ii = 0; ii < dof; ii++) for (jj = 0; jj < dof; jj++) {
            if (off) stash_push(st, ((long)pos*dof+ii)*dof+jj, A[((size_t)(a*dof+ii))*N + b*dof+jj]);
            else values[((size_t)pos*dof+ii)*dof+jj] += A[((size_t)(a*dof+ii))*N + b*dof+jj]; }
        }
        if (err) break;
      }
    }
    free(A); free(B); free(K); free(F); free(U); free(V); free(W); free(arrayU); free(arrayV); free(arrayW);
    elem_free(&e);
  }
  return err;
}

/* ------------------------------------------------------------------------------------------ */
/* exported API (ctypes)                                                                      */
/* ------------------------------------------------------------------------------------------ */

OIGA *oiga_create(int dim, int dof)
{
  OIGA *o = (OIGA*)calloc(1, sizeof(OIGA)); int i;
  o->dim = dim; o->dof = dof; o->order = -1;
  for (i = 0; i < 3; i++) axis_reset(&o->axis[i]);
  return o;
}
void oiga_destroy(OIGA *o)
{
  int i; if (!o) return;
  for (i = 0; i < 3; i++) { free(o->axis[i].U); free(o->axis[i].span); basis_free(&o->basis[i]); }
  for (i = 0; i < 3; i++) { free(o->own[i]); free(o->box_ls[i]); free(o->box_lw[i]); } free(o->rstart);
  free_rank_arrays(o); free(o->geomX_nat); free(o->geomW_nat); free(o->fixtable_glob); free(o);
}
int oiga_axis_init_uniform(OIGA *o, int i, int p, int N, double Ui, double Uf, int C, int periodic)
{ o->tables_ready = 0; return axis_init_uniform(&o->axis[i], p, N, Ui, Uf, C, periodic); }
int oiga_axis_set_knots(OIGA *o, int i, int p, int m, const double *U, int periodic)
{ o->tables_ready = 0; return axis_set_knots(&o->axis[i], p, m, U, periodic); }
void oiga_set_rule_size(OIGA *o, int i, int q) { o->rule_nqp[i] = q; o->tables_ready = 0; }
void oiga_set_order(OIGA *o, int order) { o->order = order < 1 ? 1 : (order > 4 ? 4 : order); }
void oiga_set_boundary_value(OIGA *o, int axis, int side, int field, double v) /* petigaform.c:102-121 */
{ FormBC *bc = &o->value[axis][side]; int k; for (k = 0; k < bc->count; k++) if (bc->field[k] == field) break;
  if (k == bc->count) bc->count++; bc->field[k] = field; bc->value[k] = v; }
void oiga_set_boundary_form(OIGA *o, int axis, int side, int flag) { o->visit[axis][side] = flag ? 1 : 0; }   /* petigaform.c IGAFormSetBoundaryForm */
void oiga_set_boundary_load(OIGA *o, int axis, int side, int field, double v)
{ FormBC *bc = &o->load[axis][side]; int k; for (k = 0; k < bc->count; k++) if (bc->field[k] == field) break;
  if (k == bc->count) bc->count++; bc->field[k] = field; bc->value[k] = v; }
/* geometry in natural ordering over geom sizes (n_d+1 per axis), X[...][nsd]; W may be NULL.
   rational iff max(w)-min(w) > 100*eps (src/petigaio.c:251-253) */
int oiga_set_geometry(OIGA *o, int nsd, const double *X, const double *W)
{
  size_t n = 1; int i; double wmin = DBL_MAX, wmax = -DBL_MAX; size_t k;
  for (i = 0; i < 3; i++) { const Axis *ax = &o->axis[i]; n *= (size_t)(i < o->dim ? ax->m - ax->p : 1); }
  free(o->geomX_nat); free(o->geomW_nat); o->geomW_nat = NULL;
  o->geometry = nsd; o->rational = 0;
  o->geomX_nat = (double*)malloc(n*nsd*sizeof(double)); memcpy(o->geomX_nat, X, n*nsd*sizeof(double));
  if (W) { o->geomW_nat = (double*)malloc(n*sizeof(double)); memcpy(o->geomW_nat, W, n*sizeof(double));
    for (k = 0; k < n; k++) { if (W[k] < wmin) wmin = W[k]; if (W[k] > wmax) wmax = W[k]; }
    o->rational = (fabs(wmax-wmin) > 100*DBL_EPSILON); }
  return 0;
}
void oiga_set_fixtable(OIGA *o, const double *Uglobal, long n)
{ free(o->fixtable_glob); o->fixtable_glob = NULL; o->fixtable = 0;
  if (Uglobal) { o->fixtable_glob = (double*)malloc((size_t)n*sizeof(double)); memcpy(o->fixtable_glob, Uglobal, (size_t)n*sizeof(double)); o->fixtable = 1; } }

int oiga_setup(OIGA *o, int size, int rank) { return setup_rank(o, size, rank); }
/* third vector (U0 of the IE drivers, A of the I2 drivers; global PETSc ordering), second shift, second time */
void oiga_set_aux(OIGA *o, double shift2, double t0, const double *Wg) { o->aux_shift2 = shift2; o->aux_t0 = t0; o->aux_W = Wg; }

/* info[0..]: order, then per axis: p, m, nnp, nel, nqp, nen, proc_size, proc_rank, elem_start, elem_width,
   node_lstart, node_lwidth, node_gstart, node_gwidth, geom_size  (15 per axis) */
void oiga_get_info(const OIGA *o, int *info)
{
  int i, k = 0; info[k++] = o->order;
  for (i = 0; i < 3; i++) {
    info[k++] = o->axis[i].p; info[k++] = o->axis[i].m; info[k++] = o->axis[i].nnp; info[k++] = o->axis[i].nel;
    info[k++] = o->basis[i].nqp; info[k++] = o->basis[i].nen; info[k++] = o->proc_sizes[i]; info[k++] = o->proc_ranks[i];
    info[k++] = o->elem_start[i]; info[k++] = o->elem_width[i]; info[k++] = o->node_lstart[i]; info[k++] = o->node_lwidth[i];
    info[k++] = o->node_gstart[i]; info[k++] = o->node_gwidth[i]; info[k++] = o->geom_sizes[i];
  }
}
const double *oiga_knots(const OIGA *o, int i) { return o->axis[i].U; }
const int    *oiga_spans(const OIGA *o, int i) { return o->axis[i].span; }
const int    *oiga_basis_offset(const OIGA *o, int i) { return o->basis[i].offset; }
const double *oiga_basis_detJac(const OIGA *o, int i) { return o->basis[i].detJac; }
const double *oiga_basis_weight(const OIGA *o, int i) { return o->basis[i].weight; }
const double *oiga_basis_point(const OIGA *o, int i)  { return o->basis[i].point; }
const double *oiga_basis_value(const OIGA *o, int i)  { return o->basis[i].value; }
const int    *oiga_lgmap(const OIGA *o) { return o->lgmap; }

/* global pattern handle */
GlobalPattern *oiga_pattern_create(OIGA *o, int size)
{ GlobalPattern *gp = (GlobalPattern*)calloc(1, sizeof(GlobalPattern)); if (global_pattern(o, size, gp)) { free(gp); return NULL; } return gp; }
void oiga_pattern_destroy(GlobalPattern *gp) { if (!gp) return; free(gp->rowptr); free(gp->colidx); free(gp->rank_rowstart); free(gp); }
int  oiga_pattern_nrows(const GlobalPattern *gp) { return gp->nrows; }
long oiga_pattern_nnz(const GlobalPattern *gp) { return gp->nnz; }
const int *oiga_pattern_rowptr(const GlobalPattern *gp) { return gp->rowptr; }
const int *oiga_pattern_colidx(const GlobalPattern *gp) { return gp->colidx; }
const int *oiga_pattern_rank_rowstart(const GlobalPattern *gp) { return gp->rank_rowstart; }

static int assemble(OIGA *o, int size, int r0, int r1, int slot, int form, const double *prm, double shift, const double *Vg,
                    double t, const double *Ug, const GlobalPattern *gp, double *values, double *rhs)
{ return assemble_ex(o, size, r0, r1, 0, -1, 1, NULL, slot, form, prm, shift, Vg, t, Ug, gp, values, rhs); }
int oiga_assemble(OIGA *o, int size, int slot, int form, const double *prm, double shift, const double *Vg,
                  double t, const double *Ug, const GlobalPattern *gp, double *values, double *rhs)
{ return assemble(o, size, 0, size, slot, form, prm, shift, Vg, t, Ug, gp, values, rhs); }
/* elements [idx0, idx1) of the ONE-rank element loop, ADDED into caller-zeroed arrays: lets the test harness split the loop
   over host threads (tests/par_oracle.py) while keeping the one-rank numbering */
int oiga_assemble_range(OIGA *o, int idx0, int idx1, int slot, int form, const double *prm, double shift, const double *Vg,
                        double t, const double *Ug, const GlobalPattern *gp, double *values, double *rhs)
{ return assemble_ex(o, 1, 0, 1, idx0, idx1, 0, NULL, slot, form, prm, shift, Vg, t, Ug, gp, values, rhs); }
/* One emulated MPI rank the way PETSc runs it: rows it owns are added straight into the (shared, caller-zeroed) global
   arrays, rows of other ranks go to its stash; oiga_stash_apply is the receiving side of Mat/VecAssemblyEnd.  With one
   thread per rank nothing races: a rank only writes its own row range, and applies only entries inside it. */
void *oiga_stash_create(void) { return calloc(1, sizeof(Stash)); }
void  oiga_stash_destroy(void *p) { Stash *st = (Stash*)p; if (!st) return; free(st->idx); free(st->val); free(st); }
long  oiga_stash_count(const void *p) { return ((const Stash*)p)->n; }
int oiga_assemble_rank_stash(OIGA *o, int size, int rank, int slot, int form, const double *prm, double shift, const double *Vg,
                             double t, const double *Ug, const GlobalPattern *gp, double *values, double *rhs, void *stash)
{
  Stash *st = (Stash*)stash;
  st->n = 0; st->row0 = gp->rank_rowstart[rank]; st->row1 = gp->rank_rowstart[rank+1];
  return assemble_ex(o, size, rank, rank+1, 0, -1, 0, st, slot, form, prm, shift, Vg, t, Ug, gp, values, rhs);
}
void oiga_stash_apply(const void *stash, int dof, const GlobalPattern *gp, int rank, double *values, double *rhs)
{
  const Stash *st = (const Stash*)stash; long k;
  const long r0 = gp->rank_rowstart[rank], r1 = gp->rank_rowstart[rank+1];
  const long v0 = (long)gp->rowptr[r0]*dof*dof, v1 = (long)gp->rowptr[r1]*dof*dof;
  for (k = 0; k < st->n; k++) {
    long i = st->idx[k];
    if (i >= 0) { if (i >= v0 && i < v1) values[i] += st->val[k]; }
    else { i = -i-1; if (i >= r0*dof && i < r1*dof) rhs[i] += st->val[k]; }
  }
}
/* the element loop of ONE emulated rank (its contributions only, into the global arrays): lets the CPU baseline run
   the ranks in parallel processes the way mpiexec -n size would */
int oiga_assemble_rank(OIGA *o, int size, int rank, int slot, int form, const double *prm, double shift, const double *Vg,
                       double t, const double *Ug, const GlobalPattern *gp, double *values, double *rhs)
{ return assemble(o, size, rank, rank+1, slot, form, prm, shift, Vg, t, Ug, gp, values, rhs); }

/* ------------------------------------------------------------------------------------------ */
/* IGAComputeScalar / IGAComputeErrorNorm: src/petigacomp.c:35-186                              */
/* ------------------------------------------------------------------------------------------ */
enum { SCALAR_ERRNORM=0, SCALAR_CH_STATS=1 };

/* IGAPointEvaluate (petigapoint.c:387-412) -> IGA_GetValue/Grad/Hess (petigaval.F90:182-232): V(dim^k,dof) */
static void point_evaluate(const Point *p, int k, const double *U, double *u)
{
  int a, i, c, n = ipow(p->dim, k);
  const double *N = (k == 0) ? p->N0 : (k == 1) ? p->N1 : p->N2;
  for (c = 0; c < p->dof*n; c++) u[c] = 0;
  for (a = 0; a < p->nen; a++) for (i = 0; i < p->dof; i++) for (c = 0; c < n; c++)
    u[i*n+c] = u[i*n+c] + N[(size_t)a*n+c] * U[a*p->dof+i];
}

/* exact solutions usable as the Exact callback.  id 1: test/IGAErrNorm.c:26-52 (dof 4: 1, sum x, sum x^2, prod x);
   id 2: demo/L2Projection.c:3-61 value only (choice = prm2), same function for every field */
static int exact_eval(int id, int choice, int dim, int dof, const double *x, int k, double *value)
{
  int i, j, c, n = ipow(dim, k);
  if (id == 1) {
    double prod = 1; for (i = 0; i < dim; i++) prod *= x[i];
    if (dof != 4 || k > 2) return 1;
    if (k == 0) { double s1 = 0, s2 = 0; for (i = 0; i < dim; i++) { s1 += x[i]; s2 += x[i]*x[i]; }
      value[0] = 1; value[1] = s1; value[2] = s2; value[3] = prod; }
    else if (k == 1) { for (i = 0; i < dim; i++) { value[0*dim+i] = 0; value[1*dim+i] = 1; value[2*dim+i] = 2*x[i]; value[3*dim+i] = prod/x[i]; } }
    else { for (i = 0; i < dim; i++) for (j = 0; j < dim; j++) { value[0*n+i*dim+j] = 0; value[1*n+i*dim+j] = 0;
      value[2*n+i*dim+j] = (i==j) ? 2 : 0; value[3*n+i*dim+j] = (i==j) ? 0 : (prod/(x[i]*x[j])); } }
    return 0;
  }
  if (id == 2) { double xx[3] = {0,0,0}; if (k != 0) return 1; for (i = 0; i < dim; i++) xx[i] = x[i];
    for (c = 0; c < dof; c++) value[c] = l2_function(choice, dim, xx); return 0; }
  if (id == 4) { /* test/ConvTest.c:8-28,104-111: prod sin(pi x_i), value (k = 0) or gradient (k = 1) */
    if (k > 1) return 1;
    for (c = 0; c < dof; c++) {
      if (k == 0) { double v = 1; for (i = 0; i < dim; i++) v *= sin(M_PI*x[i]); value[c] = v; }
      else for (i = 0; i < dim; i++) { double g = 1; for (j = 0; j < dim; j++) g *= (i == j) ? M_PI*cos(M_PI*x[j]) : sin(M_PI*x[j]); value[c*dim+i] = g; }
    }
    return 0; }
  if (id == 3) { double xx[3] = {0,0,0}; if (k != 0) return 1; for (i = 0; i < dim; i++) xx[i] = x[i];   /* demo/Neumann.c:5-8,80-86 Solution */
    for (c = 0; c < dof; c++) value[c] = sin(2*M_PI*xx[0]) + sin(2*M_PI*xx[1]) + sin(2*M_PI*xx[2]); return 0; }
  return 1;
}

/* the Scalar callbacks: ErrorSqr (petigacomp.c:103-124) and the CahnHilliard monitor Stats (demo/CahnHilliard2D.c:36-58) */
static int scalar_point(int sid, const double *prm, const Point *p, const double *U, int n, double *S, double *w0, double *w1)
{
  int i, j;
  if (sid == SCALAR_ERRNORM) {
    int k = (int)prm[0], ex = (int)prm[1], nn = ipow(p->dim, k);
    if (k < 0 || k > 2 || n != p->dof) return 1;
    if (U) point_evaluate(p, k, U, w0); else for (i = 0; i < p->dof*nn; i++) w0[i] = 0;   /* vecU == NULL: zero state */
    for (i = 0; i < p->dof*nn; i++) w1[i] = 0;
    if (ex) { if (exact_eval(ex, (int)prm[2], p->dim, p->dof, p->x, k, w1)) return 1; }
    for (i = 0; i < p->dof; i++) for (j = 0; j < nn; j++) { double e = fabs(w1[i*nn+j] - w0[i*nn+j]); S[i] += e*e; }
    return 0;
  }
  if (sid == SCALAR_CH_STATS) {
    double theta = prm[0], alpha = prm[1], cbar = prm[2], c, c1[3], diff;
    if (p->dim != 2 || p->dof != 1 || n != 3 || !U) return 1;
    get_value(p, U, &c); get_grad(p, U, c1);
    diff = c - cbar;
    S[0] = c*log(c) + (1-c)*log(1-c) + 2*theta*c*(1-c) + theta/(3*alpha)*(c1[0]*c1[0]+c1[1]*c1[1]);
    S[1] = diff*diff;
    S[2] = S[1]*diff;
    return 0;
  }
  return 1;
}

/* IGAComputeScalar (petigacomp.c:35-96) over `size` emulated ranks; the final loop over ranks is the MPI_Allreduce */
int oiga_compute_scalar(OIGA *o, int size, int sid, const double *prm, const double *Ug, int n, double *S)
{
  int r, err = 0, i;
  double *localS = (double*)calloc((size_t)n*size, sizeof(double)), *workS = (double*)malloc(sizeof(double)*(size_t)n);
  for (r = 0; r < size && !err; r++) {
    Elem e; int index, count, N, ng, dof = o->dof; double *U = NULL, *arrayU = NULL, *w0, *w1;
    if (setup_rank(o, size, r)) { err = 1; break; }
    elem_alloc(&e, o);
    N = e.nen*dof; ng = o->node_gwidth[0]*o->node_gwidth[1]*o->node_gwidth[2];
    w0 = (double*)malloc(sizeof(double)*(size_t)dof*81); w1 = (double*)malloc(sizeof(double)*(size_t)dof*81);
    if (Ug) { int a, c; U = (double*)malloc(sizeof(double)*(size_t)N); arrayU = (double*)malloc(sizeof(double)*(size_t)ng*dof);
      for (a = 0; a < ng; a++) for (c = 0; c < dof; c++) arrayU[(size_t)a*dof+c] = Ug[(size_t)o->lgmap[a]*dof+c]; }
    count = o->elem_width[0]*o->elem_width[1]*o->elem_width[2];
    for (index = 0; index < count && !err; index++) {
      int q, a, idx = index;
      for (i = 0; i < 3; i++) { int coord = idx % o->elem_width[i]; idx = (idx - coord)/o->elem_width[i]; e.ID[i] = coord + o->elem_start[i]; }
      elem_closure(&e);
      if (Ug) for (a = 0; a < e.nen; a++) for (i = 0; i < dof; i++) U[a*dof+i] = arrayU[(size_t)e.mapping[a]*dof+i];   /* IGAElementGetValues */
      elem_tabulate(&e);
      for (q = 0; q < e.nqp; q++) {
        Point p; double JW = e.detJac[q] * e.weight[q];
        p.nen = e.nen; p.dof = dof; p.dim = e.dim; p.nsd = e.nsd;
        p.N0 = e.shape[0] + (size_t)q*e.nen; p.N1 = e.shape[1] + (size_t)q*e.nen*e.nsd; p.N2 = e.shape[2] + (size_t)q*e.nen*e.nsd*e.nsd;
        p.x = e.geometry ? e.mapX[0] + (size_t)q*e.nsd : e.mapU[0] + (size_t)q*e.dim;
        p.atboundary = 0; p.boundary_id = -1; p.normal = NULL;
        memset(workS, 0, sizeof(double)*(size_t)n);
        if (scalar_point(sid, prm, &p, U, n, workS, w0, w1)) { err = 10; break; }
        for (i = 0; i < n; i++) localS[(size_t)r*n+i] += workS[i] * JW;    /* IGAPointAddArray (petigapoint.c:451-465) */
      }
    }
    free(U); free(arrayU); free(w0); free(w1);
    elem_free(&e);
  }
  for (i = 0; i < n; i++) { S[i] = 0; for (r = 0; r < size; r++) S[i] += localS[(size_t)r*n+i]; }
  free(localS); free(workS);
  return err;
}

/* Boundary tabulation of one element face (after oiga_setup): out arrays [nqp_face]: detJac (already *detS), detS, normal[..][nsd],
   X0[..][nsd], shape0[..][nen] -- for the known answers of test/IGAGeometryMap.c:275-389 */
int oiga_tabulate_boundary(OIGA *o, const int ID[3], int axis, int side, int *nqp, double *weight, double *detJac, double *detS,
                           double *normal, double *X0, double *shape0)
{
  Elem e; int k;
  elem_alloc(&e, o);
  for (k = 0; k < 3; k++) e.ID[k] = ID[k];
  e.atboundary = 1; e.baxis = axis; e.bside = side;
  elem_closure(&e); elem_tabulate(&e);
  *nqp = e.nqp;
  if (weight) memcpy(weight, e.weight, sizeof(double)*e.nqp);
  if (detJac) memcpy(detJac, e.detJac, sizeof(double)*e.nqp);
  if (detS)   memcpy(detS, e.detS, sizeof(double)*e.nqp);
  if (normal) memcpy(normal, e.normal, sizeof(double)*e.nqp*e.nsd);
  if (X0) memcpy(X0, e.geometry ? e.mapX[0] : e.mapU[0], sizeof(double)*e.nqp*e.nsd);
  if (shape0) memcpy(shape0, e.shape[0], sizeof(double)*e.nqp*e.nen);
  elem_free(&e);
  return 0;
}

/* Tabulate one element of the current rank (after oiga_setup) for the geometry known-answer tests.
   out arrays sized by the caller: weight[nqp], detJac[nqp] (already *detX), detX[nqp], point[nqp][dim],
   X0[nqp][nsd], X1[nqp][nsd][dim], shape0[nqp][nen], shape1[nqp][nen][nsd], shape2[nqp][nen][nsd][nsd] */
int oiga_tabulate_element(OIGA *o, const int ID[3], int *nqp, int *nen, double *weight, double *detJac, double *detX,
                          double *point, double *X0, double *X1, double *X2, double *X3,
                          double *shape0, double *shape1, double *shape2, double *shape3)
{
  Elem e; int dim, nsd, k;
  elem_alloc(&e, o); dim = e.dim; nsd = e.nsd;
  for (k = 0; k < 3; k++) e.ID[k] = ID[k];
  elem_closure(&e); elem_tabulate(&e);
  *nqp = e.nqp; *nen = e.nen;
  if (weight) memcpy(weight, e.weight, sizeof(double)*e.nqp);
  if (detJac) memcpy(detJac, e.detJac, sizeof(double)*e.nqp);
  if (detX)   memcpy(detX, e.detX, sizeof(double)*e.nqp);
  if (point)  memcpy(point, e.mapU[0], sizeof(double)*e.nqp*dim);
  if (X0) memcpy(X0, e.geometry ? e.mapX[0] : e.mapU[0], sizeof(double)*e.nqp*nsd);
  if (X1) memcpy(X1, e.mapX[1], sizeof(double)*e.nqp*nsd*dim);
  if (X2 && e.order >= 2) memcpy(X2, e.mapX[2], sizeof(double)*e.nqp*nsd*dim*dim);
  if (X3 && e.order >= 3) memcpy(X3, e.mapX[3], sizeof(double)*e.nqp*nsd*dim*dim*dim);
  if (shape0) memcpy(shape0, e.shape[0], sizeof(double)*e.nqp*e.nen);
  if (shape1) memcpy(shape1, e.shape[1], sizeof(double)*e.nqp*e.nen*nsd);
  if (shape2 && e.order >= 2) memcpy(shape2, e.shape[2], sizeof(double)*e.nqp*e.nen*nsd*nsd);
  if (shape3 && e.order >= 3) memcpy(shape3, e.shape[3], sizeof(double)*e.nqp*e.nen*nsd*nsd*nsd);
  elem_free(&e);
  return 0;
}
