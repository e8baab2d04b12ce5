/* rkfd_oracle.c - CPU oracle for the RoKi-FD step path.  TEST INFRASTRUCTURE, NOT PRODUCT.
 * See rkfd_oracle.h for the status of this file (parity unpinned, what may load it).
 *
 * Every function cites the reference (mi-lib/roki-fd v1.7.9) file:line it restates, or the
 * [EXT] assumption (DESIGN.md "EXT assumptions", tags A-k as in SURVEY.md Appendix A) it
 * fixes when the arithmetic lives in RoKi/ZM/Zeo, which are not in the reference tree.
 *
 * Conventions: 3x3 matrices row-major m[3*r+c]; 6-D vectors (linear[3], angular[3]) (A-1);
 * link velocities/accelerations are those of the link origin expressed in the link frame,
 * accelerations are classical (A-2); gravity is applied as a force m*g at each COM (A-4).
 */
#include "rkfd_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------ */
/* small dense helpers */
static void v3_set(double *a, double x, double y, double z){ a[0]=x; a[1]=y; a[2]=z; }
static void v3_copy(const double *a, double *b){ b[0]=a[0]; b[1]=a[1]; b[2]=a[2]; }
static void v3_add(const double *a, const double *b, double *c){ c[0]=a[0]+b[0]; c[1]=a[1]+b[1]; c[2]=a[2]+b[2]; }
static void v3_sub(const double *a, const double *b, double *c){ c[0]=a[0]-b[0]; c[1]=a[1]-b[1]; c[2]=a[2]-b[2]; }
static void v3_cat(double *a, double k, const double *b){ a[0]+=k*b[0]; a[1]+=k*b[1]; a[2]+=k*b[2]; }
static double v3_dot(const double *a, const double *b){ return a[0]*b[0]+a[1]*b[1]+a[2]*b[2]; }
static void v3_cross(const double *a, const double *b, double *c)
{
  double x = a[1]*b[2]-a[2]*b[1], y = a[2]*b[0]-a[0]*b[2], z = a[0]*b[1]-a[1]*b[0];
  c[0]=x; c[1]=y; c[2]=z;
}
static double v3_norm(const double *a){ return sqrt(v3_dot(a,a)); }
/* y = M x */
static void m3_mulv(const double *m, const double *x, double *y)
{
  double a = m[0]*x[0]+m[1]*x[1]+m[2]*x[2];
  double b = m[3]*x[0]+m[4]*x[1]+m[5]*x[2];
  double c = m[6]*x[0]+m[7]*x[1]+m[8]*x[2];
  y[0]=a; y[1]=b; y[2]=c;
}
/* y = M^T x */
static void m3_tmulv(const double *m, const double *x, double *y)
{
  double a = m[0]*x[0]+m[3]*x[1]+m[6]*x[2];
  double b = m[1]*x[0]+m[4]*x[1]+m[7]*x[2];
  double c = m[2]*x[0]+m[5]*x[1]+m[8]*x[2];
  y[0]=a; y[1]=b; y[2]=c;
}
static void m3_mul(const double *a, const double *b, double *c)
{
  double t[9]; int i,j,k;
  for(i=0;i<3;i++) for(j=0;j<3;j++){ t[3*i+j]=0; for(k=0;k<3;k++) t[3*i+j]+=a[3*i+k]*b[3*k+j]; }
  memcpy(c,t,sizeof t);
}
static void m3_ident(double *m){ memset(m,0,9*sizeof(double)); m[0]=m[4]=m[8]=1.0; }
static void m3_skew(const double *p, double *s)
{
  s[0]=0; s[1]=-p[2]; s[2]=p[1]; s[3]=p[2]; s[4]=0; s[5]=-p[0]; s[6]=-p[1]; s[7]=p[0]; s[8]=0;
}

/* [EXT A-3] angle-axis vector -> rotation matrix (Rodrigues; Zeo zMat3DFromAA) */
static void aa_to_mat(const double *aa, double *R)
{
  double th2 = v3_dot(aa,aa), A, B, K[9], K2[9]; int i;
  if( th2 < 1.0e-24 ){ A = 1.0; B = 0.5; }
  else { double th = sqrt(th2); A = sin(th)/th; B = (1.0-cos(th))/th2; }
  m3_skew(aa,K); m3_mul(K,K,K2); m3_ident(R);
  for(i=0;i<9;i++) R[i] += A*K[i] + B*K2[i];
}
/* [EXT A-3/A-9] aa <- log( R(w) R(aa) ): world(org)-frame angular increment (Zeo zAACascade),
 * evaluated through unit quaternions */
static void aa_cascade(double *aa, const double *w)
{
  double q1[4], q2[4], q[4], th, s, n;
  if( w[0] == 0.0 && w[1] == 0.0 && w[2] == 0.0 ) return;      /* no increment: the displacement is kept bit for bit (a held joint) */
  th = v3_norm(aa);
  if( th < 1.0e-12 ){ q1[0]=1.0; q1[1]=0.5*aa[0]; q1[2]=0.5*aa[1]; q1[3]=0.5*aa[2]; }
  else { s = sin(0.5*th)/th; q1[0]=cos(0.5*th); q1[1]=s*aa[0]; q1[2]=s*aa[1]; q1[3]=s*aa[2]; }
  th = v3_norm(w);
  if( th < 1.0e-12 ){ q2[0]=1.0; q2[1]=0.5*w[0]; q2[2]=0.5*w[1]; q2[3]=0.5*w[2]; }
  else { s = sin(0.5*th)/th; q2[0]=cos(0.5*th); q2[1]=s*w[0]; q2[2]=s*w[1]; q2[3]=s*w[2]; }
  /* q = q2 * q1 */
  q[0] = q2[0]*q1[0] - q2[1]*q1[1] - q2[2]*q1[2] - q2[3]*q1[3];
  q[1] = q2[0]*q1[1] + q2[1]*q1[0] + q2[2]*q1[3] - q2[3]*q1[2];
  q[2] = q2[0]*q1[2] - q2[1]*q1[3] + q2[2]*q1[0] + q2[3]*q1[1];
  q[3] = q2[0]*q1[3] + q2[1]*q1[2] - q2[2]*q1[1] + q2[3]*q1[0];
  if( q[0] < 0 ){ q[0]=-q[0]; q[1]=-q[1]; q[2]=-q[2]; q[3]=-q[3]; }
  n = sqrt(q[1]*q[1]+q[2]*q[2]+q[3]*q[3]);
  if( n < 1.0e-12 ){ aa[0]=2.0*q[1]; aa[1]=2.0*q[2]; aa[2]=2.0*q[3]; }
  else { th = 2.0*atan2(n,q[0]); aa[0]=th*q[1]/n; aa[1]=th*q[2]/n; aa[2]=th*q[3]/n; }
}

/* ------------------------------------------------------------------------------------ */
typedef struct {
  int parent, jtype, mtype, stuff;
  double org_R[9], org_p[3];
  double mass, com[3], inertia[9];
  double stiffness, viscosity, coulomb, sfriction;
  double mk, madm, mgear, mrotor, mgearin, mmin, mmax;
  int qofs, ndof;
  double brk_f, brk_t;      /* breakable float: thresholds */
  double M[36];   /* rigid-body 6x6 inertia at the link origin (A-4) */
} ork_link;

/* slide mode of a collision cell ("fake crawler", rkfd_sim.c:386-440): mode, belt speed, axis in the frame of the cell's link;
 * ord = registration order of the shape in rkCD (decides which cell of a pair is pd->cell[0]) */
typedef struct { int mode, ord; double vel, axis[3]; } ork_slide;
typedef struct { int link, nvert, vofs, volbox; ork_slide sl; } ork_cell;   /* volbox: the 8 corners of a parallelepiped in sign-bit order (Volume solver, A-15) */
/* link < 0: a static box, (R, p) its world frame; link >= 0: a box carried by that moving link, (R, p) its frame in the link */
typedef struct { double R[9], p[3], half[3]; int stuff, link; ork_slide sl; double lR[9], lp[3]; /* static box: world frame of its link */ } ork_box;
typedef struct { int type; double K, L, E, V, SF, KF; } ork_cinfo;
typedef struct { int sa, sb; ork_cinfo ci; } ork_cinfo_ent;
typedef struct { int cell, box, sofs; ork_cinfo ci; } ork_pair;    /* vertices of `cell` against `box` (static, or on another moving link) */

struct ork_world {
  int nl, nq;
  ork_link *link;
  int ncell, nvert; ork_cell *cell; double *vert;
  int nbox; ork_box *box;
  int nci; ork_cinfo_ent *ci; ork_cinfo cidef;
  int npair, nslot; ork_pair *pair;
  int *self_col;      /* per link: 1 when pairs between cells of ITS chain are registered (rkCDPairChainUnreg clears it) */
  double dt, friction_weight; int pyramid, max_iter, solver, integrator;
  double sc_table[2][64];   /* sin/cos pyramid table (rkfd_util.c:199-214) */
};

/* per-link working set of one dynamics evaluation */
typedef struct {
  double Rrel[9], prel[3];  /* link frame w.r.t. parent frame */
  double Rw[9], pw[3];      /* world frame */
  double v[6];              /* link velocity (lin, ang) in link frame */
  double zeta[6];           /* velocity-product acceleration */
  double S[36];             /* motion subspace, 6 x ndof, row-major with row stride 6 */
  double X[36];             /* acceleration transform parent -> link */
  double IA[36], pA[6];     /* articulated inertia / bias */
  double U[36], Dinv[36], u[6];
  double wext[6];           /* external wrench at origin, link frame */
  double a[6];              /* link acceleration */
  double a0[6];             /* saved acceleration (rkChainSaveABIAccBias) */
  double du[6];             /* probe increments of u */
  double dp[6];             /* probe increments of pA */
} ork_lw;

struct ork_env {
  const ork_world *w;
  double t;
  double *q, *qd, *qdd;       /* committed state (fd->dis, fd->vel, fd->acc) */
  double *min;                /* motor input per link */
  int *piv_type; double *piv_prev;   /* joint friction pivot (A-8) per dof */
  double *tf, *tdrive, *jm;   /* per-dof friction torque, driving torque, motor inertia */
  int *c_active, *c_type; double *c_ref, *c_f, *c_pro;   /* per contact slot */
  double *c_norm, *c_axis, *c_vert, *c_refw, *c_vel;     /* per slot, current evaluation */
  ork_lw *lw;
  int *broken, *nd;           /* per link: breakable float broken? ; effective joint dofs of this evaluation */
  /* rigid system of the last evaluation */
  int rn; double *rA, *rb, *rf;
  /* test hook: the last Vert QP (ork_env_get_qp) */
  int qp_n, qp_m, qp_iter, qp_term; double *qp_Q, *qp_c, *qp_nf, *qp_x; int *qp_idx;
  /* Volume solver, per pair: friction type (rkCDPairDat.type), planes of the last evaluation (-1: no volume), wrench, center */
  int *v_type, *v_np; double *v_wrench, *v_center;
  double *v_qc;      /* per pair: Q6 (36, row-major), c6 (6), norm (3) of the last evaluation (test hook) */
  /* RKG workspace */
  double *k[4][2], *xs[2];
};

/* ------------------------------------------------------------------------------------ */
static int jtype_ndof(int jt)
{
  switch(jt){ case ORK_JOINT_REVOL: case ORK_JOINT_PRISM: return 1;
              case ORK_JOINT_CYLIN: case ORK_JOINT_HOOKE: return 2;
              case ORK_JOINT_SPHER: return 3; case ORK_JOINT_FLOAT: case ORK_JOINT_BRFLOAT: return 6; default: return 0; }
}

static void link_build_inertia(ork_link *l)
{
  /* [EXT A-4] M = [[m E, -m[c x]],[m[c x], Ic - m[c x][c x]]] acting on (lin acc, ang acc) */
  double C[9], C2[9]; int r,c;
  memset(l->M,0,sizeof l->M);
  m3_skew(l->com,C); m3_mul(C,C,C2);
  for(r=0;r<3;r++){
    l->M[6*r+r] = l->mass;
    for(c=0;c<3;c++){
      l->M[6*r+3+c]     = -l->mass*C[3*r+c];
      l->M[6*(3+r)+c]   =  l->mass*C[3*r+c];
      l->M[6*(3+r)+3+c] =  l->inertia[3*r+c] - l->mass*C2[3*r+c];
    }
  }
}

ork_world *ork_world_new(int nl, const int *li, const double *ld)
{
  ork_world *w = (ork_world*)calloc(1,sizeof *w); int i;
  w->nl = nl; w->link = (ork_link*)calloc(nl,sizeof(ork_link));
  w->self_col = (int*)calloc(nl>0?nl:1,sizeof(int)); for(i=0;i<nl;i++) w->self_col[i] = 1;
  w->nq = 0;
  for(i=0;i<nl;i++){
    ork_link *l = &w->link[i]; const double *d = ld + ORK_LINK_ND*i;
    l->parent = li[4*i]; l->jtype = li[4*i+1]; l->mtype = li[4*i+2]; l->stuff = li[4*i+3];
    memcpy(l->org_R,d,9*sizeof(double)); memcpy(l->org_p,d+9,3*sizeof(double));
    l->mass = d[12]; memcpy(l->com,d+13,3*sizeof(double)); memcpy(l->inertia,d+16,9*sizeof(double));
    l->stiffness=d[25]; l->viscosity=d[26]; l->coulomb=d[27]; l->sfriction=d[28];
    l->mk=d[29]; l->madm=d[30]; l->mgear=d[31]; l->mrotor=d[32]; l->mgearin=d[33]; l->mmin=d[34]; l->mmax=d[35];
    l->ndof = jtype_ndof(l->jtype); l->qofs = w->nq; w->nq += l->ndof;
    link_build_inertia(l);
  }
  /* reference defaults: rkfd_property.c:10-18, rkfd_defs.h:15-23 */
  ork_world_set_prp(w, 0.001, 8, 100.0, 10);
  ork_world_set_solver(w, ORK_SOLVER_VERT);   /* rkfd_sim.c:52 */
  return w;
}
void ork_world_free(ork_world *w)
{
  if(!w) return;
  free(w->link); free(w->cell); free(w->vert); free(w->box); free(w->ci); free(w->pair); free(w->self_col); free(w);
}
int ork_world_add_cell(ork_world *w, int link, int nvert, const double *verts)
{
  w->cell = (ork_cell*)realloc(w->cell,(w->ncell+1)*sizeof(ork_cell));
  w->vert = (double*)realloc(w->vert,3*(w->nvert+nvert)*sizeof(double));
  w->cell[w->ncell].link = link; w->cell[w->ncell].nvert = nvert; w->cell[w->ncell].vofs = w->nvert; w->cell[w->ncell].volbox = 0;
  memset(&w->cell[w->ncell].sl,0,sizeof(ork_slide)); w->cell[w->ncell].sl.ord = w->ncell;
  memcpy(w->vert+3*w->nvert, verts, 3*nvert*sizeof(double));
  w->nvert += nvert;
  return w->ncell++;
}
int ork_world_add_box(ork_world *w, const double *R, const double *p, const double *half, int stuff)
{
  ork_box *b;
  w->box = (ork_box*)realloc(w->box,(w->nbox+1)*sizeof(ork_box));
  b = &w->box[w->nbox];
  memcpy(b->R,R,9*sizeof(double)); memcpy(b->p,p,3*sizeof(double)); memcpy(b->half,half,3*sizeof(double));
  b->stuff = stuff; b->link = -1;
  memset(&b->sl,0,sizeof(ork_slide)); b->sl.ord = 1000000 + w->nbox;      /* default: static shapes registered after the moving ones */
  m3_ident(b->lR); b->lp[0]=b->lp[1]=b->lp[2]=0.0;
  return w->nbox++;
}
/* a box primitive on a moving link (frame in the link): a collision TARGET for the vertices of cells on other links; its own
 * 8 corners are registered as a cell by the caller.  [EXT A-10] moving-vs-moving collision = vertex of one cell inside a box of
 * another link (both directions when both carry boxes). */
int ork_world_add_link_box(ork_world *w, int link, const double *R, const double *p, const double *half)
{
  int b = ork_world_add_box(w,R,p,half,w->link[link].stuff);
  w->box[b].link = link;
  return b;
}
/* slide mode (rkFDCDCellSetSlideMode/-Vel/-Axis, rkfd_sim.c:386-400) of a cell (is_box = 0) or of a box target (is_box = 1);
 * ord: registration order of the shape; lR/lp: world frame of the link of a STATIC box (NULL: identity) */
void ork_world_set_slide(ork_world *w, int is_box, int idx, int mode, double vel, const double *axis, int ord, const double *lR, const double *lp)
{
  ork_slide *sl = is_box ? &w->box[idx].sl : &w->cell[idx].sl;
  sl->mode = mode; sl->vel = vel; memcpy(sl->axis,axis,3*sizeof(double)); sl->ord = ord;
  if( is_box && lR ){ memcpy(w->box[idx].lR,lR,9*sizeof(double)); memcpy(w->box[idx].lp,lp,3*sizeof(double)); }
}
/* [EXT] rkCDPairChainUnreg: drops the pairs between cells of the chain that `link` belongs to (registered by default) */
void ork_world_unreg_self_collision(ork_world *w, int link)
{
  int i, r = link, ri;
  while( w->link[r].parent >= 0 ) r = w->link[r].parent;
  for(i=0;i<w->nl;i++){ ri = i; while( w->link[ri].parent >= 0 ) ri = w->link[ri].parent; if( ri == r ) w->self_col[i] = 0; }
}
void ork_world_add_contact_info(ork_world *w, int sa, int sb, int type,
                                double K, double L, double E, double V, double SF, double KF)
{
  ork_cinfo_ent *e;
  w->ci = (ork_cinfo_ent*)realloc(w->ci,(w->nci+1)*sizeof(ork_cinfo_ent));
  e = &w->ci[w->nci++];
  e->sa=sa; e->sb=sb; e->ci.type=type; e->ci.K=K; e->ci.L=L; e->ci.E=E; e->ci.V=V; e->ci.SF=SF; e->ci.KF=KF;
}
/* [EXT] zODE2AssignRegular( ode, RKG | RK4 | Euler | Heun ) (reference rkfd_sim.h:86): 0 Runge-Kutta-Gill (the
 * default, rkfd_sim.c:46-47), 1 classical Runge-Kutta, 2 Euler, 3 Heun */
void ork_world_set_integrator(ork_world *w, int integrator){ w->integrator = integrator; }
void ork_world_set_prp(ork_world *w, double dt, int pyramid, double fw, int max_iter)
{
  int i; double th, dth, off;
  w->dt=dt; w->pyramid=pyramid; w->friction_weight=fw; w->max_iter=max_iter;
  /* rkFDCrateSinCosTable (rkfd_util.c:199-214) with the Vert offset -pi/pyramid (rkfd_vert.c:369) */
  off = -M_PI / pyramid; dth = 2.0*M_PI / pyramid;
  for( i=0,th=0.0; i<pyramid && i<64; i++,th+=dth ){
    w->sc_table[0][i] = sin(th+off); w->sc_table[1][i] = cos(th+off);
  }
}
void ork_world_set_solver(ork_world *w, int solver)
{
  w->solver = solver;
  /* default contact info: rkfd_vert.c:340-348, rkfd_mlcp.c:301-310, rkfd_volume.c:961-969 */
  w->cidef.type = ORK_CONTACT_RIGID;
  w->cidef.K = 1000.0; w->cidef.L = solver == ORK_SOLVER_VOLUME ? 0.001 : 1.0; w->cidef.SF = 0.5; w->cidef.KF = 0.3;
  w->cidef.E = 0.0; w->cidef.V = 0.0;
}
/* [EXT A-15] the Volume solver takes cells with the 8 corners of a parallelepiped; their vertices are put into sign-bit
 * order (vertex k = v0 + (k&1) ea + (k>>1&1) eb + (k>>2&1) ec, v0 = the first vertex, (a,b,c) the first index triple in
 * lexicographic order that spans the cell) so that a polyhedron-described box (mighty.ztk's soles: bottom ring, top ring)
 * is accepted.  Returns 0 when the 8 points are not a parallelepiped (order left as it is). */
static int box_sign_bit_order(double *v)
{
  int a, b, c, k, j, i; double scale = 0, out[24];
  for(k=1;k<8;k++) for(i=0;i<3;i++) if( fabs(v[3*k+i]-v[i]) > scale ) scale = fabs(v[3*k+i]-v[i]);
  for(a=1;a<8;a++) for(b=a+1;b<8;b++) for(c=b+1;c<8;c++){
    int used = 0, ok = 1;
    for(k=0;k<8 && ok;k++){
      double p[3]; int found = -1;
      for(i=0;i<3;i++) p[i] = v[i] + ((k&1)?v[3*a+i]-v[i]:0.0) + ((k&2)?v[3*b+i]-v[i]:0.0) + ((k&4)?v[3*c+i]-v[i]:0.0);
      for(j=0;j<8;j++) if( !(used>>j & 1) && fabs(v[3*j]-p[0]) <= 1e-9*(1+scale) && fabs(v[3*j+1]-p[1]) <= 1e-9*(1+scale) && fabs(v[3*j+2]-p[2]) <= 1e-9*(1+scale) ){ found = j; break; }
      if( found < 0 ){ ok = 0; break; }
      used |= 1<<found; memcpy(out+3*k,v+3*found,24);
    }
    if( ok ){ memcpy(v,out,sizeof out); return 1; }
  }
  return 0;
}
void ork_world_finalize(ork_world *w)
{
  /* pairs in registration order: (moving cell) x (static box); contact info by stuff pair with
   * fallback to the solver default (rkfd_sim.c:200-207, :266-271) */
  int c,b,k,sofs=0;
  if( w->solver == ORK_SOLVER_VOLUME ) for(c=0;c<w->ncell;c++) w->cell[c].volbox = w->cell[c].nvert == 8 && box_sign_bit_order(w->vert+3*w->cell[c].vofs);
  /* static pairs first (cell order x box order), then the pairs with boxes on moving links: cells of OTHER links - of other
   * chains, and of the same chain while its self-collision pairs are registered ([EXT] rkCDPairChainUnreg) */
  { int pass, n = 0;
    free(w->pair); w->pair = NULL;
    for(pass=0;pass<2;pass++){
      k = 0;
      for(c=0;c<w->ncell;c++) for(b=0;b<w->nbox;b++) if( w->box[b].link < 0 ){ if( pass ){ w->pair[k].cell = c; w->pair[k].box = b; } k++; }
      for(c=0;c<w->ncell;c++) for(b=0;b<w->nbox;b++) if( w->box[b].link >= 0 ){
        int la = w->cell[c].link, lb = w->box[b].link, ra = la, rb = lb;
        if( la == lb ) continue;
        while( w->link[ra].parent >= 0 ) ra = w->link[ra].parent;
        while( w->link[rb].parent >= 0 ) rb = w->link[rb].parent;
        if( ra == rb && !w->self_col[la] ) continue;
        if( pass ){ w->pair[k].cell = c; w->pair[k].box = b; } k++; }
      if( !pass ){ n = k; w->pair = (ork_pair*)calloc(n>0?n:1,sizeof(ork_pair)); }
    }
    w->npair = n; }
  { int n = 0, dropped = 0;
    for(k=0;k<w->npair;k++){
      ork_pair p = w->pair[k]; int i, sa, sb;
      c = p.cell; b = p.box; sa = w->link[w->cell[c].link].stuff; sb = w->box[b].stuff;
      p.ci = w->cidef;
      for(i=0;i<w->nci;i++)
        if( (w->ci[i].sa==sa && w->ci[i].sb==sb) || (w->ci[i].sa==sb && w->ci[i].sb==sa) ){ p.ci = w->ci[i].ci; break; }
      /* rigid contact between two MOVING links: the vertex solvers couple the two links / chains through A
       * (rkfd_vert.c:125-185, rkfd_mlcp.c:76-142); the Volume solver's contact volume ([EXT A-15]) is formed against a static
       * box only - such pairs are not formed under it (the device side does the same and says so) */
      if( w->box[b].link >= 0 && p.ci.type != ORK_CONTACT_ELASTIC && w->solver == ORK_SOLVER_VOLUME ){ dropped++; continue; }
      p.sofs = sofs; sofs += w->cell[c].nvert;
      w->pair[n++] = p;
    }
    w->npair = n; (void)dropped; }
  w->nslot = sofs;
}
void ork_world_set_break(ork_world *w, int link, double fth, double tth){ w->link[link].brk_f = fth; w->link[link].brk_t = tth; }
void ork_env_get_broken(const ork_env *e, int *broken){ int i; for(i=0;i<e->w->nl;i++) broken[i] = e->broken[i]; }
int ork_world_nq(const ork_world *w){ return w->nq; }
int ork_world_nslot(const ork_world *w){ return w->nslot; }
int ork_world_nl(const ork_world *w){ return w->nl; }

/* ------------------------------------------------------------------------------------ */
ork_env *ork_env_new(const ork_world *w)
{
  ork_env *e = (ork_env*)calloc(1,sizeof *e); int nq = w->nq>0?w->nq:1, ns = w->nslot>0?w->nslot:1, i, j;
  e->w = w; e->t = 0.0;
  e->q=(double*)calloc(nq,8); e->qd=(double*)calloc(nq,8); e->qdd=(double*)calloc(nq,8);
  e->min=(double*)calloc(w->nl,8);
  e->broken=(int*)calloc(w->nl>0?w->nl:1,sizeof(int)); e->nd=(int*)calloc(w->nl>0?w->nl:1,sizeof(int));
  /* friction pivot init {SF, q, 0} (rkfd_sim.c:157-175) */
  e->piv_type=(int*)calloc(nq,sizeof(int)); e->piv_prev=(double*)calloc(nq,8);
  e->tf=(double*)calloc(nq,8); e->tdrive=(double*)calloc(nq,8); e->jm=(double*)calloc(nq,8);
  e->c_active=(int*)calloc(ns,sizeof(int)); e->c_type=(int*)calloc(ns,sizeof(int));
  e->c_ref=(double*)calloc(3*ns,8); e->c_f=(double*)calloc(3*ns,8); e->c_pro=(double*)calloc(3*ns,8);
  e->c_norm=(double*)calloc(3*ns,8); e->c_axis=(double*)calloc(9*ns,8); e->c_vert=(double*)calloc(3*ns,8);
  e->c_refw=(double*)calloc(3*ns,8); e->c_vel=(double*)calloc(3*ns,8);
  e->lw=(ork_lw*)calloc(w->nl,sizeof(ork_lw));
  e->rn=0; e->rA=(double*)calloc(9*ns*ns,8); e->rb=(double*)calloc(3*ns,8); e->rf=(double*)calloc(3*ns,8);
  { int np = w->npair>0?w->npair:1; e->v_type=(int*)calloc(np,sizeof(int)); e->v_np=(int*)calloc(np,sizeof(int));
    e->v_wrench=(double*)calloc(6*np,8); e->v_center=(double*)calloc(3*np,8); e->v_qc=(double*)calloc(45*np,8); }
  for(i=0;i<4;i++) for(j=0;j<2;j++) e->k[i][j]=(double*)calloc(nq,8);
  e->xs[0]=(double*)calloc(nq,8); e->xs[1]=(double*)calloc(nq,8);
  return e;
}
void ork_env_free(ork_env *e)
{
  int i,j; if(!e) return;
  free(e->q); free(e->qd); free(e->qdd); free(e->min); free(e->piv_type); free(e->piv_prev);
  free(e->tf); free(e->tdrive); free(e->jm);
  free(e->c_active); free(e->c_type); free(e->c_ref); free(e->c_f); free(e->c_pro);
  free(e->c_norm); free(e->c_axis); free(e->c_vert); free(e->c_refw); free(e->c_vel);
  free(e->broken); free(e->nd); free(e->lw); free(e->rA); free(e->rb); free(e->rf);
  free(e->qp_Q); free(e->qp_c); free(e->qp_nf); free(e->qp_x); free(e->qp_idx);
  free(e->v_type); free(e->v_np); free(e->v_wrench); free(e->v_center); free(e->v_qc);
  for(i=0;i<4;i++) for(j=0;j<2;j++) free(e->k[i][j]);
  free(e->xs[0]); free(e->xs[1]); free(e);
}
void ork_env_set_state(ork_env *e, const double *q, const double *qd)
{ /* rkFDChainSetDis / SetVel (rkfd_sim.c:277-287) */
  if(q)  memcpy(e->q, q, e->w->nq*8);
  if(qd) memcpy(e->qd,qd,e->w->nq*8);
}
void ork_env_get_state(const ork_env *e, double *q, double *qd, double *qdd)
{
  if(q)   memcpy(q,  e->q,  e->w->nq*8);
  if(qd)  memcpy(qd, e->qd, e->w->nq*8);
  if(qdd) memcpy(qdd,e->qdd,e->w->nq*8);
}
void ork_env_set_motor_input(ork_env *e, const double *u){ memcpy(e->min,u,e->w->nl*8); }
void ork_env_get_pivot(const ork_env *e, int *type, double *prev)
{ memcpy(type,e->piv_type,e->w->nq*sizeof(int)); memcpy(prev,e->piv_prev,e->w->nq*8); }
void ork_env_set_pivot(ork_env *e, const int *type, const double *prev)
{ memcpy(e->piv_type,type,e->w->nq*sizeof(int)); memcpy(e->piv_prev,prev,e->w->nq*8); }
void ork_env_get_contact(const ork_env *e, int *active, int *type, double *ref, double *f)
{
  int ns = e->w->nslot;
  if(active) memcpy(active,e->c_active,ns*sizeof(int));
  if(type)   memcpy(type,e->c_type,ns*sizeof(int));
  if(ref)    memcpy(ref,e->c_ref,3*ns*8);
  if(f)      memcpy(f,e->c_f,3*ns*8);
}
void ork_env_set_contact(ork_env *e, const int *active, const int *type, const double *ref)
{
  int ns = e->w->nslot;
  memcpy(e->c_active,active,ns*sizeof(int)); memcpy(e->c_type,type,ns*sizeof(int)); memcpy(e->c_ref,ref,3*ns*8);
}
double ork_env_time(const ork_env *e){ return e->t; }

/* ------------------------------------------------------------------------------------ */
/* outward pass 1: rkChainFK + rkChainSetJointVelAll + rkChainUpdateVel
 * (call site rkfd_sim.c:298-300; [EXT A-2, A-3]) */
static void eval_kinematics(ork_env *e, const double *q, const double *qd)
{
  const ork_world *w = e->w; int i, r, c;
  for(i=0;i<w->nl;i++){
    const ork_link *l = &w->link[i]; ork_lw *x = &e->lw[i];
    const double *qi = q + l->qofs, *qdi = qd + l->qofs;
    double RJ[9], pJ[3] = {0,0,0}, vJ[6] = {0,0,0,0,0,0};
    double wp[3] = {0,0,0}, vp[3] = {0,0,0};     /* parent velocity in parent frame */
    double wpl[3], t1[3], t2[3];
    m3_ident(RJ); memset(x->S,0,sizeof x->S);
    e->nd[i] = l->ndof;
    if( l->jtype == ORK_JOINT_BRFLOAT && !e->broken[i] ) e->nd[i] = 0;      /* rigid: the displacement stays where it is */
    switch(l->jtype){
    case ORK_JOINT_REVOL: { double s = sin(qi[0]), co = cos(qi[0]);
      RJ[0]=co; RJ[1]=-s; RJ[3]=s; RJ[4]=co; x->S[6*5+0]=1.0; vJ[5]=qdi[0]; } break;
    case ORK_JOINT_PRISM: pJ[2]=qi[0]; x->S[6*2+0]=1.0; vJ[2]=qdi[0]; break;
    case ORK_JOINT_SPHER: aa_to_mat(qi,RJ);
      for(r=0;r<3;r++) for(c=0;c<3;c++) x->S[6*(3+r)+c] = RJ[3*c+r];   /* S = RJ^T (angular) */
      m3_tmulv(RJ,qdi,vJ+3); break;
    case ORK_JOINT_BRFLOAT:
      if( !e->broken[i] ){ v3_copy(qi,pJ); aa_to_mat(qi+3,RJ); break; }
      /* fall through: broken = float */
    case ORK_JOINT_FLOAT: v3_copy(qi,pJ); aa_to_mat(qi+3,RJ);
      for(r=0;r<3;r++) for(c=0;c<3;c++){ x->S[6*r+c] = RJ[3*c+r]; x->S[6*(3+r)+3+c] = RJ[3*c+r]; }
      m3_tmulv(RJ,qdi,vJ); m3_tmulv(RJ,qdi+3,vJ+3); break;
    /* [EXT] RoKi rk_joint_cylin: slides along and turns about the joint's z axis; both motion axes are constant in the link */
    case ORK_JOINT_CYLIN: { double s = sin(qi[1]), co = cos(qi[1]);
      pJ[2]=qi[0]; RJ[0]=co; RJ[1]=-s; RJ[3]=s; RJ[4]=co;
      x->S[6*2+0]=1.0; x->S[6*5+1]=1.0; vJ[2]=qdi[0]; vJ[5]=qdi[1]; } break;
    /* [EXT] RoKi rk_joint_hooke: R = Rz(q0) Ry(q1); in the link frame the first axis is Ry(q1)^T z = (-sin q1, 0, cos q1), the
     * second is y; the first axis moves in the link: its rate (-cos q1, 0, -sin q1) q1' q0' enters the bias acceleration */
    case ORK_JOINT_HOOKE: { double s0 = sin(qi[0]), c0 = cos(qi[0]), s1 = sin(qi[1]), c1 = cos(qi[1]);
      double Rz[9] = {c0,-s0,0, s0,c0,0, 0,0,1}, Ry[9] = {c1,0,s1, 0,1,0, -s1,0,c1};
      m3_mul(Rz,Ry,RJ);
      x->S[6*3+0]=-s1; x->S[6*5+0]=c1; x->S[6*4+1]=1.0;
      vJ[3]=-s1*qdi[0]; vJ[4]=qdi[1]; vJ[5]=c1*qdi[0]; } break;
    default: break;
    }
    m3_mul(l->org_R,RJ,x->Rrel);
    m3_mulv(l->org_R,pJ,x->prel); v3_add(x->prel,l->org_p,x->prel);
    if( l->parent >= 0 ){
      const ork_lw *p = &e->lw[l->parent];
      m3_mul(p->Rw,x->Rrel,x->Rw); m3_mulv(p->Rw,x->prel,x->pw); v3_add(x->pw,p->pw,x->pw);
      v3_copy(p->v,vp); v3_copy(p->v+3,wp);
    } else { memcpy(x->Rw,x->Rrel,sizeof x->Rw); v3_copy(x->prel,x->pw); }
    /* velocity: w_i = R^T w_p + w_J ; v_i = R^T (v_p + w_p x p) + v_J */
    m3_tmulv(x->Rrel,wp,wpl);
    v3_cross(wp,x->prel,t1); v3_add(vp,t1,t1); m3_tmulv(x->Rrel,t1,x->v);
    v3_add(x->v,vJ,x->v); v3_add(wpl,vJ+3,x->v+3);
    /* velocity-product acceleration: lin = R^T( w_p x (w_p x p) ) + 2 (R^T w_p) x v_J ; ang = (R^T w_p) x w_J */
    v3_cross(wp,x->prel,t1); v3_cross(wp,t1,t2); m3_tmulv(x->Rrel,t2,x->zeta);
    v3_cross(wpl,vJ,t1); v3_cat(x->zeta,2.0,t1);
    v3_cross(wpl,vJ+3,x->zeta+3);
    if( l->jtype == ORK_JOINT_HOOKE ){ x->zeta[3] += -cos(qi[1])*qdi[0]*qdi[1]; x->zeta[5] += -sin(qi[1])*qdi[0]*qdi[1]; }   /* S' q' */
    /* acceleration transform X = [[R^T, -R^T [p x]],[0, R^T]] */
    { double P[9], RtP[9], Rt[9];
      for(r=0;r<3;r++) for(c=0;c<3;c++) Rt[3*r+c] = x->Rrel[3*c+r];
      m3_skew(x->prel,P); m3_mul(Rt,P,RtP); memset(x->X,0,sizeof x->X);
      for(r=0;r<3;r++) for(c=0;c<3;c++){
        x->X[6*r+c] = Rt[3*r+c]; x->X[6*r+3+c] = -RtP[3*r+c]; x->X[6*(3+r)+3+c] = Rt[3*r+c]; } }
    memset(x->wext,0,sizeof x->wext);
  }
}

/* ------------------------------------------------------------------------------------ */
/* rkFDLinkPointWldVel (rkfd_util.c:14-24): world velocity of world point p carried by link */
static void link_point_wld_vel(const ork_lw *x, const double *p, double *v)
{
  double lin[3], ang[3], r[3], t[3];
  m3_mulv(x->Rw,x->v,lin); m3_mulv(x->Rw,x->v+3,ang);
  v3_sub(p,x->pw,r); v3_cross(ang,r,t); v3_add(lin,t,v);
}
/* rkFDLinkPointWldAcc (rkfd_util.c:92-101) with [EXT] rkLinkPointAcc = a + alpha x r + w x (w x r) */
static void link_point_wld_acc(const ork_lw *x, const double *p, double *a)
{
  double r[3], t[3], t2[3], al[3];
  v3_sub(p,x->pw,r); m3_tmulv(x->Rw,r,r);
  v3_copy(x->a,al);
  v3_cross(x->a+3,r,t); v3_add(al,t,al);
  v3_cross(x->v+3,r,t); v3_cross(x->v+3,t,t2); v3_add(al,t2,al);
  m3_mulv(x->Rw,al,a);
}

/* [EXT A-10] rkCDColChkVert restricted to (moving vertex cloud) x (static box), followed by
 * the elastic/rigid partition of rkFDCDUpdate (rkfd_cd.c:33-49).  Persistent per-vertex state:
 * active / type / anchor _ref (box frame); new contacts start {SF, _ref=_pro}. */
static void eval_collision(ork_env *e)
{
  const ork_world *w = e->w; int pi, k, a;
  for(pi=0;pi<w->npair;pi++){
    const ork_pair *p = &w->pair[pi]; const ork_cell *cl = &w->cell[p->cell]; const ork_box *bx0 = &w->box[p->box];
    const ork_lw *x = &e->lw[cl->link];
    ork_box bw = *bx0; const ork_box *bx = &bw;
    if( bx0->link >= 0 ){      /* a box on a moving link: its world frame of this evaluation */
      const ork_lw *xb = &e->lw[bx0->link];
      m3_mul(xb->Rw,bx0->R,bw.R); m3_mulv(xb->Rw,bx0->p,bw.p); v3_add(bw.p,xb->pw,bw.p);
    }
    for(k=0;k<cl->nvert;k++){
      int s = p->sofs + k, amin = 0, inside = 1; double vw[3], vb[3], d[3], dep, depmin = 0, sg;
      double nb[3] = {0,0,0}, t1b[3] = {0,0,0}, t2b[3] = {0,0,0}, prob[3];
      m3_mulv(x->Rw,w->vert+3*(cl->vofs+k),vw); v3_add(vw,x->pw,vw);
      v3_sub(vw,bx->p,d); m3_tmulv(bx->R,d,vb);
      for(a=0;a<3;a++){
        dep = bx->half[a] - fabs(vb[a]);
        if( dep <= -ORK_TOL ) inside = 0;
        if( a==0 || dep < depmin ){ depmin = dep; amin = a; }
      }
      v3_copy(vw,e->c_vert+3*s);
      if( !inside ){ e->c_active[s] = 0; continue; }
      sg = vb[amin] >= 0 ? 1.0 : -1.0;
      nb[amin] = sg; t1b[(amin+1)%3] = 1.0; t2b[(amin+2)%3] = sg;
      v3_copy(vb,prob); prob[amin] = sg*bx->half[amin];
      m3_mulv(bx->R,nb,e->c_norm+3*s);
      m3_mulv(bx->R,nb,e->c_axis+9*s); m3_mulv(bx->R,t1b,e->c_axis+9*s+3); m3_mulv(bx->R,t2b,e->c_axis+9*s+6);
      v3_copy(prob,e->c_pro+3*s);
      if( !e->c_active[s] ){ e->c_active[s] = 1; e->c_type[s] = ORK_SF; v3_copy(prob,e->c_ref+3*s); }
      m3_mulv(bx->R,e->c_ref+3*s,e->c_refw+3*s); v3_add(e->c_refw+3*s,bx->p,e->c_refw+3*s);
    }
  }
}

/* rkFDKineticFrictionWeight (rkfd_util.c:193-196) */
static double kinetic_friction_weight(double w, double fs){ return 1.0 - exp(-1.0*w*fs); }

/* rkFDContactForcePushWrench (rkfd_util.c:268-282); the static partner's wrench is dropped */
static void push_wrench(ork_env *e, int link, const double *vert, const double *f)
{
  ork_lw *x = &e->lw[link]; double pos[3], fl[3], n[3];
  v3_sub(vert,x->pw,pos); m3_tmulv(x->Rw,pos,pos);     /* zXform3DInv */
  m3_tmulv(x->Rw,f,fl);
  v3_cross(pos,fl,n);                                   /* [EXT A-4] wrench at origin: (f, pos x f) */
  v3_add(x->wext,fl,x->wext); v3_add(x->wext+3,n,x->wext+3);
}

/* frame (R, p) of the link that carries a cell / a box in this evaluation (static box: the fixed frame of its link) */
static void slide_frame(const ork_env *e, int is_box, int idx, const double **R, const double **pw)
{
  const ork_world *w = e->w; int link = is_box ? w->box[idx].link : w->cell[idx].link;
  if( link >= 0 ){ *R = e->lw[link].Rw; *pw = e->lw[link].pw; } else { *R = w->box[idx].lR; *pw = w->box[idx].lp; }
}
/* belt direction at world point p: (R axis) x (p - p_link) without its normal component; 0 when it vanishes.  Returns its norm */
static double slide_dir(const ork_env *e, int is_box, int idx, const double *p, const double *n, double *sv)
{
  const ork_world *w = e->w; const ork_slide *sl = is_box ? &w->box[idx].sl : &w->cell[idx].sl;
  const double *R, *pw; double r[3], ax[3];
  slide_frame(e,is_box,idx,&R,&pw);
  v3_sub(p,pw,r); m3_mulv(R,sl->axis,ax); v3_cross(ax,r,sv); v3_cat(sv,-v3_dot(sv,n),n);
  return v3_norm(sv);
}
/* rkFDLinkAddSlideVel (rkfd_util.c:26-40) for the two cells of pair pi; v = relative velocity of the vertex's cell (in: without slide) */
static void add_slide_vel(const ork_env *e, int pi, const double *p, const double *n, double *v)
{
  const ork_world *w = e->w; const ork_pair *pr = &w->pair[pi]; double sv[3], nv;
  if( w->cell[pr->cell].sl.mode ){ nv = slide_dir(e,0,pr->cell,p,n,sv); if( !(fabs(nv) < ORK_TOL) ) v3_cat(v, w->cell[pr->cell].sl.vel/nv, sv); }
  if( w->box[pr->box].sl.mode ){ nv = slide_dir(e,1,pr->box,p,n,sv); if( !(fabs(nv) < ORK_TOL) ) v3_cat(v, -w->box[pr->box].sl.vel/nv, sv); }
}
/* rkFDUpdateRefSlide (rkfd_util.c:218-237): the anchor of a sticking contact rides on the belt.  The reference adds
 * R_k^T sv to _ref with k = pd->cell[1]'s link when the sliding cell is the vertex's cell, else pd->cell[0]'s link; _ref lives in
 * the partner's frame, so the world shift is R_partner R_k^T sv (= sv whenever the sliding cell is pd->cell[0]) - mirrored.
 * Our anchor is kept in the box frame (A-10): c_ref += Rbox_world^T (world shift). */
static void update_ref_slide(ork_env *e, int pi, int s)
{
  const ork_world *w = e->w; const ork_pair *pr = &w->pair[pi]; int i;
  const int vfirst = w->cell[pr->cell].sl.ord < w->box[pr->box].sl.ord;     /* the vertex's cell is pd->cell[0] */
  const double *Rv, *pv, *Rp, *pp; double Rbw[9];
  slide_frame(e,0,pr->cell,&Rv,&pv); slide_frame(e,1,pr->box,&Rp,&pp);
  if( w->box[pr->box].link >= 0 ) m3_mul(Rp,w->box[pr->box].R,Rbw); else memcpy(Rbw,w->box[pr->box].R,sizeof Rbw);
  for(i=0;i<2;i++){
    const int is_vcell = vfirst ? i == 0 : i == 1;         /* pd->cell[i] is the vertex's cell */
    const ork_slide *sl = is_vcell ? &w->cell[pr->cell].sl : &w->box[pr->box].sl;
    double sv[3], nv, t[3], dw[3], db[3]; const double *Rk;
    if( !sl->mode ) continue;
    nv = slide_dir(e, is_vcell ? 0 : 1, is_vcell ? pr->cell : pr->box, e->c_vert+3*s, e->c_norm+3*s, sv);
    if( fabs(nv) < ORK_TOL ) continue;
    { double k = ( is_vcell ? -1.0 : 1.0 ) * w->dt * sl->vel / nv; sv[0]*=k; sv[1]*=k; sv[2]*=k; }
    /* k-link: pd->cell[ is_vcell ? 1 : 0 ] */
    { const int kcell_is_v = is_vcell ? !vfirst : vfirst; Rk = kcell_is_v ? Rv : Rp; }
    m3_tmulv(Rk,sv,t); m3_mulv(Rp,t,dw);
    m3_tmulv(Rbw,dw,db); v3_add(e->c_ref+3*s,db,e->c_ref+3*s);
  }
}

/* rkFDContactForceModifyFriction (rkfd_util.c:239-266); v passed by value */
static void modify_friction(ork_env *e, int pi, const ork_cinfo *ci, int s, const double *vin, int do_up_ref)
{
  double *f = e->c_f+3*s, *ax = e->c_axis+9*s, v[3], fn, fs, vs, mu;
  v3_copy(vin,v);
  fn = v3_dot(f,ax);
  fs = sqrt( v3_dot(f,ax+3)*v3_dot(f,ax+3) + v3_dot(f,ax+6)*v3_dot(f,ax+6) );
  mu = e->c_type[s]==ORK_SF ? ci->SF : ci->KF;
  if( !(fabs(fs) < ORK_TOL) && fs > mu*fn ){
    v3_cat(v,-v3_dot(v,ax),ax);
    vs = v3_norm(v);
    f[0]=fn*ax[0]; f[1]=fn*ax[1]; f[2]=fn*ax[2];
    if( !(fabs(vs) < ORK_TOL) ){
      v[0]/=vs; v[1]/=vs; v[2]/=vs;
      v3_cat(f, -kinetic_friction_weight(e->w->friction_weight,vs)*ci->KF*fn, v);
    }
    if( do_up_ref ){ e->c_type[s] = ORK_KF; v3_copy(e->c_pro+3*s,e->c_ref+3*s); }
  } else {
    if( do_up_ref ){ e->c_type[s] = ORK_SF; update_ref_slide(e,pi,s); }
  }
}

/* rkFDSolverPenalty (rkfd_penalty.c:11-31) */
static void solver_penalty(ork_env *e, int do_up_ref)
{
  const ork_world *w = e->w; int pi, k;
  for(pi=0;pi<w->npair;pi++){
    const ork_pair *p = &w->pair[pi]; const ork_cell *cl = &w->cell[p->cell];
    if( p->ci.type != ORK_CONTACT_ELASTIC ) continue;
    for(k=0;k<cl->nvert;k++){
      int s = p->sofs+k; double d[3], vr[3], *f = e->c_f+3*s;
      if( !e->c_active[s] ) continue;
      v3_sub(e->c_vert+3*s,e->c_refw+3*s,d);
      link_point_wld_vel(&e->lw[cl->link],e->c_vert+3*s,vr);    /* rkFDChainPointRelativeVel (rkfd_util.c:42-60), STAT partner = 0 */
      if( w->box[p->box].link >= 0 ){ double vb[3]; link_point_wld_vel(&e->lw[w->box[p->box].link],e->c_vert+3*s,vb); v3_sub(vr,vb,vr); }
      add_slide_vel(e,pi,e->c_vert+3*s,e->c_norm+3*s,vr);
      f[0] = -p->ci.E*d[0]; f[1] = -p->ci.E*d[1]; f[2] = -p->ci.E*d[2];
      v3_cat(f, -1.0*(p->ci.V + p->ci.E*w->dt), vr);
      if( v3_dot(f,e->c_axis+9*s) < 0.0 ) continue;
      modify_friction(e,pi,&p->ci,s,vr,do_up_ref);
      push_wrench(e,cl->link,e->c_vert+3*s,f);
      if( w->box[p->box].link >= 0 ){        /* the partner takes the opposite force at the same point (rkfd_util.c:276-278) */
        double fr[3] = { -f[0], -f[1], -f[2] }; push_wrench(e,w->box[p->box].link,e->c_vert+3*s,fr); }
    }
  }
}

/* [EXT A-6] DC / torque motor */
static void motor_eval(const ork_link *l, double in, double qd, double *tin, double *treg, double *jm)
{
  double ecl = in < l->mmin ? l->mmin : ( in > l->mmax ? l->mmax : in );
  switch(l->mtype){
  case ORK_MOTOR_DC:
    *tin = l->mgear*l->mk*l->madm*ecl;
    *treg = (l->mgear*l->mk)*(l->mgear*l->mk)*l->madm*qd;
    *jm = l->mgear*l->mgear*(l->mrotor+l->mgearin); break;
  case ORK_MOTOR_TRQ: *tin = ecl; *treg = 0; *jm = 0; break;
  default: *tin = 0; *treg = 0; *jm = 0; break;
  }
}

/* rkFDJointFriction (rkfd_util.c:366-387) + rkFDJointFrictionRevolDC (:330-364) +
 * rkFDJointFrictionAll (:318-328); also evaluates the motor driving torque used by the ABA */
static void joint_friction(ork_env *e, const double *q, const double *qd, int do_up_ref)
{
  const ork_world *w = e->w; int i, j;
  for(i=0;i<w->nl;i++){
    const ork_link *l = &w->link[i];
    if( l->ndof == 1 ){
      int o = l->qofs; double tin, treg, jm, tf, fmax, v = qd[o];
      motor_eval(l,e->min[i],v,&tin,&treg,&jm);
      e->tdrive[o] = tin - treg; e->jm[o] = jm;
      if( l->mtype == ORK_MOTOR_DC ){
        tf = jm; tf *= -v / w->dt; tf -= tin; tf += treg; tf += e->piv_prev[o];
        if( e->piv_type[o] == ORK_SF ) fmax = l->sfriction;
        else{ /* [EXT A-7] rkJointGetKFriction */
          double sg = v > 0 ? 1.0 : ( v < 0 ? -1.0 : 0.0 );
          fmax = -l->stiffness*q[o] - l->viscosity*v - l->coulomb*sg;
        }
        fmax = fabs(fmax);
        if( fabs(tf) > fmax ){
          tf = tf > 0 ? fmax : -fmax;
          if( do_up_ref ) e->piv_type[o] = ORK_KF;
        } else if( do_up_ref ) e->piv_type[o] = ORK_SF;
        e->tf[o] = tf;
      }
      /* 1-DoF joints without a DC motor: friction is never set (rkfd_util.c:379-383) -> stays 0 */
    } else {
      /* multi-DoF: kf_i (1 - exp(-w |v_i|)); spherical/float joints carry no passive torque -> 0 */
      for(j=0;j<l->ndof;j++){ e->tf[l->qofs+j] = 0.0*kinetic_friction_weight(w->friction_weight,fabs(qd[l->qofs+j]));
                              e->tdrive[l->qofs+j] = 0.0; e->jm[l->qofs+j] = 0.0; }
    }
  }
}

/* rkFDUpdateJointPrevDrivingTrq (rkfd_util.c:289-311) */
static void update_prev_driving_trq(ork_env *e)
{
  int i; for(i=0;i<e->w->nq;i++) e->piv_prev[i] = e->tdrive[i] + e->tf[i];
}

/* ------------------------------------------------------------------------------------ */
/* dense 6x6 helpers for the ABA */
static void m6_mulv(const double *m, const double *x, double *y)
{ double t[6]; int r,c; for(r=0;r<6;r++){ t[r]=0; for(c=0;c<6;c++) t[r]+=m[6*r+c]*x[c]; } memcpy(y,t,sizeof t); }
static void m6_tmulv(const double *m, const double *x, double *y)
{ double t[6]; int r,c; for(r=0;r<6;r++){ t[r]=0; for(c=0;c<6;c++) t[r]+=m[6*c+r]*x[c]; } memcpy(y,t,sizeof t); }

/* in-place inverse of a symmetric positive definite n x n (n<=6) matrix stored with stride 6 */
static void spd_inverse(double *a, int n)
{
  double L[36], Li[36]; int i,j,k;
  memset(L,0,sizeof L); memset(Li,0,sizeof Li);
  for(j=0;j<n;j++){
    double s = a[6*j+j];
    for(k=0;k<j;k++) s -= L[6*j+k]*L[6*j+k];
    L[6*j+j] = sqrt(s);
    for(i=j+1;i<n;i++){
      s = a[6*i+j];
      for(k=0;k<j;k++) s -= L[6*i+k]*L[6*j+k];
      L[6*i+j] = s / L[6*j+j];
    }
  }
  for(j=0;j<n;j++){
    Li[6*j+j] = 1.0 / L[6*j+j];
    for(i=j+1;i<n;i++){
      double s = 0; for(k=j;k<i;k++) s -= L[6*i+k]*Li[6*k+j];
      Li[6*i+j] = s / L[6*i+i];
    }
  }
  for(i=0;i<n;i++) for(j=0;j<n;j++){
    double s = 0; for(k=(i>j?i:j);k<n;k++) s += Li[6*k+i]*Li[6*k+j];
    a[6*i+j] = s;
  }
}

/* [EXT A-4] rkChainUpdateABI: init + inward articulated-inertia pass + outward acceleration pass
 * (call sites rkfd_sim.c:509-520, rkfd_util.c:156) */
static void aba_backward(ork_env *e)
{
  const ork_world *w = e->w; int i, r, c, k, j;
  for(i=0;i<w->nl;i++){
    const ork_link *l = &w->link[i]; ork_lw *x = &e->lw[i];
    double *om = x->v+3, t1[3], t2[3], Io[9], fg[3], ng[3], gl[3], gw[3] = {0,0,-ORK_G};
    memcpy(x->IA,l->M,sizeof x->IA);
    /* bias: ( m w x (w x c) ; w x (Io w) ) */
    v3_cross(om,l->com,t1); v3_cross(om,t1,t2);
    x->pA[0]=l->mass*t2[0]; x->pA[1]=l->mass*t2[1]; x->pA[2]=l->mass*t2[2];
    for(r=0;r<3;r++) for(c=0;c<3;c++) Io[3*r+c] = l->M[6*(3+r)+3+c];
    m3_mulv(Io,om,t1); v3_cross(om,t1,x->pA+3);
    /* gravity as a force at the COM */
    m3_tmulv(x->Rw,gw,gl); fg[0]=l->mass*gl[0]; fg[1]=l->mass*gl[1]; fg[2]=l->mass*gl[2];
    v3_cross(l->com,fg,ng);
    for(r=0;r<3;r++){ x->pA[r] -= fg[r] + x->wext[r]; x->pA[3+r] -= ng[r] + x->wext[3+r]; }
  }
  for(i=w->nl-1;i>=0;i--){
    const ork_link *l = &w->link[i]; ork_lw *x = &e->lw[i]; int nd = e->nd[i];
    double pz[6], Ia[36], pa[6], t6[6];
    m6_mulv(x->IA,x->zeta,pz); for(r=0;r<6;r++) pz[r] += x->pA[r];     /* p' = pA + IA zeta */
    memcpy(Ia,x->IA,sizeof Ia); memcpy(pa,pz,sizeof pa);
    if( nd > 0 ){
      /* U = IA S (6 x nd) ; D = S^T U + Jm ; u = tau - S^T p' */
      for(r=0;r<6;r++) for(j=0;j<nd;j++){ double s=0; for(k=0;k<6;k++) s += x->IA[6*r+k]*x->S[6*k+j]; x->U[6*r+j]=s; }
      memset(x->Dinv,0,sizeof x->Dinv);
      for(r=0;r<nd;r++) for(j=0;j<nd;j++){ double s=0; for(k=0;k<6;k++) s += x->S[6*k+r]*x->U[6*k+j]; x->Dinv[6*r+j]=s; }
      for(j=0;j<nd;j++) x->Dinv[6*j+j] += e->jm[l->qofs+j];
      if( nd == 1 ) x->Dinv[0] = 1.0/x->Dinv[0]; else spd_inverse(x->Dinv,nd);
      for(j=0;j<nd;j++){ double s=0; for(k=0;k<6;k++) s += x->S[6*k+j]*pz[k];
        x->u[j] = e->tdrive[l->qofs+j] + e->tf[l->qofs+j] - s; }
      /* Ia = IA - U Dinv U^T ; pa = p' + U Dinv u */
      for(j=0;j<nd;j++){ t6[j]=0; for(k=0;k<nd;k++) t6[j] += x->Dinv[6*j+k]*x->u[k]; }
      for(r=0;r<6;r++) for(j=0;j<nd;j++) pa[r] += x->U[6*r+j]*t6[j];
      for(r=0;r<6;r++) for(c=0;c<6;c++){ double s=0;
        for(j=0;j<nd;j++) for(k=0;k<nd;k++) s += x->U[6*r+j]*x->Dinv[6*j+k]*x->U[6*c+k];
        Ia[6*r+c] -= s; }
    }
    if( l->parent >= 0 ){
      ork_lw *p = &e->lw[l->parent]; double T[36];
      /* IA_p += X^T Ia X ; pA_p += X^T pa */
      for(r=0;r<6;r++) for(c=0;c<6;c++){ double s=0; for(k=0;k<6;k++) s += Ia[6*r+k]*x->X[6*k+c]; T[6*r+c]=s; }
      for(r=0;r<6;r++) for(c=0;c<6;c++){ double s=0; for(k=0;k<6;k++) s += x->X[6*k+r]*T[6*k+c]; p->IA[6*r+c]+=s; }
      m6_tmulv(x->X,pa,t6); for(r=0;r<6;r++) p->pA[r] += t6[r];
    }
  }
}
/* outward acceleration pass; du = optional per-dof increment of u (cached-ABA probes) */
static void aba_forward(ork_env *e, double *qdd, int use_du)
{
  const ork_world *w = e->w; int i, j, k, r;
  for(i=0;i<w->nl;i++){
    const ork_link *l = &w->link[i]; ork_lw *x = &e->lw[i]; int nd = e->nd[i];
    double ap[6] = {0,0,0,0,0,0}, rhs[6], qa[6] = {0,0,0,0,0,0};
    if( l->parent >= 0 ) m6_mulv(x->X,e->lw[l->parent].a,ap);
    for(j=0;j<nd;j++){ double s=0; for(k=0;k<6;k++) s += x->U[6*k+j]*ap[k];
      rhs[j] = x->u[j] + ( use_du ? x->du[j] : 0.0 ) - s; }
    for(j=0;j<nd;j++){ qa[j]=0; for(k=0;k<nd;k++) qa[j] += x->Dinv[6*j+k]*rhs[k]; }
    for(r=0;r<6;r++){ double s = ap[r] + x->zeta[r]; for(j=0;j<nd;j++) s += x->S[6*r+j]*qa[j]; x->a[r] = s; }
    if( qdd ) for(j=0;j<l->ndof;j++) qdd[l->qofs+j] = j < nd ? qa[j] : 0.0;
  }
}

/* [EXT A-5] rkChainUpdateCachedABIPair: unit test wrench (f at world point vert) on `link`;
 * bias-only inward propagation with the cached articulated inertias, then the outward pass */
static void aba_probe2(ork_env *e, int link, int link2, const double *vert, const double *fw)
{
  const ork_world *w = e->w; int i, j, k, r, t;
  for(i=0;i<w->nl;i++){ memset(e->lw[i].du,0,sizeof e->lw[i].du); memset(e->lw[i].dp,0,sizeof e->lw[i].dp); }
  /* the test wrench on `link`, its opposite on the partner `link2` (a moving partner: rkfd_vert.c:166-175) */
  for(t=0;t<2;t++){ int lk = t ? link2 : link; ork_lw *x; double pos[3], fl[3], n[3], sg = t ? -1.0 : 1.0;
    if( lk < 0 ) continue;
    x = &e->lw[lk];
    v3_sub(vert,x->pw,pos); m3_tmulv(x->Rw,pos,pos); m3_tmulv(x->Rw,fw,fl); v3_cross(pos,fl,n);
    for(r=0;r<3;r++){ x->dp[r] += -sg*fl[r]; x->dp[3+r] += -sg*n[r]; } }
  /* bias-only inward pass: links in descending index order (parents precede their children), so that a link that
   * collects the increments of two paths (the two links of one chain) is processed once, with their sum */
  for(i=w->nl-1;i>=0;i--){
    const ork_link *l = &w->link[i]; ork_lw *y = &e->lw[i]; int nd = e->nd[i]; double pa[6], t6[6], nz = 0;
    for(r=0;r<6;r++) nz += fabs(y->dp[r]);
    if( nz == 0.0 ) continue;
    memcpy(pa,y->dp,sizeof pa);
    for(j=0;j<nd;j++){ double s=0; for(k=0;k<6;k++) s += y->S[6*k+j]*y->dp[k]; y->du[j] = -s; }
    for(j=0;j<nd;j++){ t6[j]=0; for(k=0;k<nd;k++) t6[j] += y->Dinv[6*j+k]*y->du[k]; }
    for(r=0;r<6;r++) for(j=0;j<nd;j++) pa[r] += y->U[6*r+j]*t6[j];
    if( l->parent < 0 ) continue;
    m6_tmulv(y->X,pa,t6); for(r=0;r<6;r++) e->lw[l->parent].dp[r] += t6[r];
  }
  aba_forward(e,NULL,1);
}
static void aba_probe(ork_env *e, int link, const double *vert, const double *fw){ aba_probe2(e,link,-1,vert,fw); }

/* ------------------------------------------------------------------------------------ */
/* zLESolveMP restatement for a symmetric matrix: x = pinv(a) b through a cyclic Jacobi
 * eigen-decomposition ([EXT A-14]; the KKT matrix of rkfd_opt_qp.c:82-106 is symmetric) */
void ork_le_solve_mp_sym(int n, const double *a, const double *b, double *x)
{
  double *A = (double*)malloc(n*n*8), *V = (double*)malloc(n*n*8); int i,j,k,sweep; double lmax = 0;
  memcpy(A,a,n*n*8);
  for(i=0;i<n;i++) for(j=0;j<n;j++) V[n*i+j] = i==j;
  for(sweep=0;sweep<60;sweep++){
    double off = 0; for(i=0;i<n;i++) for(j=i+1;j<n;j++) off += A[n*i+j]*A[n*i+j];
    if( off < 1e-300 ) break;
    for(i=0;i<n;i++) for(j=i+1;j<n;j++){
      double apq = A[n*i+j], th, t, c, s;
      if( fabs(apq) < 1e-300 ) continue;
      th = (A[n*j+j]-A[n*i+i])/(2.0*apq);
      t = (th>=0?1.0:-1.0)/(fabs(th)+sqrt(th*th+1.0)); c = 1.0/sqrt(t*t+1.0); s = t*c;
      for(k=0;k<n;k++){ double akp=A[n*k+i], akq=A[n*k+j]; A[n*k+i]=c*akp-s*akq; A[n*k+j]=s*akp+c*akq; }
      for(k=0;k<n;k++){ double apk=A[n*i+k], aqk=A[n*j+k]; A[n*i+k]=c*apk-s*aqk; A[n*j+k]=s*apk+c*aqk; }
      for(k=0;k<n;k++){ double vkp=V[n*k+i], vkq=V[n*k+j]; V[n*k+i]=c*vkp-s*vkq; V[n*k+j]=s*vkp+c*vkq; }
    }
  }
  for(i=0;i<n;i++) if( fabs(A[n*i+i]) > lmax ) lmax = fabs(A[n*i+i]);
  for(i=0;i<n;i++) x[i] = 0;
  /* x = pinv(a) b, then iterative refinement x += pinv(a) (b - a x) with the residual accumulated in long double: the
   * active-set loop of rkfd_opt_qp.c decides with ABSOLUTE 1e-12 thresholds on x (:108-110), while one pseudo-inverse
   * application in double carries an error of eps * cond(a) * |x| (1e-16 * 1e7 * 1e2 on the KKT matrices of contactinfo.ztk's
   * relaxation 1e-4).  [EXT A-14] specifies zLESolveMP as THE minimum-norm least-squares solution; refinement makes this
   * implementation deliver it to rounding, so that the loop's decisions are those of exact arithmetic
   * (tests/test_oracle_physics.py pins it against a 60-digit evaluation).  Two solves whose exact solutions coincide (an
   * active row released from a redundant set) then return the same doubles up to an ulp. */
  { int round; double *r = (double*)malloc(n*8); long double *xl = (long double*)calloc(n,sizeof(long double));
    for(round=0;round<5;round++){
      for(i=0;i<n;i++){ long double acc = b[i]; for(j=0;j<n;j++) acc -= (long double)a[n*i+j]*xl[j]; r[i] = (double)acc; }
      for(k=0;k<n;k++){
        double lam = A[n*k+k], s = 0;
        if( fabs(lam) <= 1.0e-11*lmax ) continue;
        for(i=0;i<n;i++) s += V[n*i+k]*r[i];
        s /= lam;
        for(i=0;i<n;i++) xl[i] += (long double)(s*V[n*i+k]);
      }
    }
    for(i=0;i<n;i++) x[i] = (double)xl[i];      /* the iterate is kept in long double and rounded once */
    free(xl);
    free(r); }
  free(A); free(V);
}

/* rkFDQPSolveASM (rkfd_opt_qp.c:43-181): min 1/2 x^T Q x + c^T x  s.t.  a x >= b.
 * `cond` is always row.x here: the Vert callback (rkfd_vert.c:246-250) is the same dot product
 * restricted to the only three non-zero entries of the row. */
static double qp_cond(int n, const double *a, const double *x, int i)
{ double s = 0; int j; for(j=0;j<n;j++) s += a[n*i+j]*x[j]; return s; }
#define ORK_QP_ASM_TOL 1.0e-8
#define ORK_QP_MAX_ITER 10000
/* term (may be NULL): how the loop ended - 0 optimal (rkfd_opt_qp.c:113), 1 anti-cycling exit (:165), 2 iteration cap (not in the reference) */
int ork_qp_solve_asm_ex(int n, int m, const double *q, const double *c, const double *a,
                        const double *b, const double *init, double *ans, int *idx, int *term);
int ork_qp_solve_asm(int n, int m, const double *q, const double *c, const double *a,
                     const double *b, const double *init, double *ans, int *idx)
{ return ork_qp_solve_asm_ex(n,m,q,c,a,b,init,ans,idx,NULL); }
int ork_qp_solve_asm_ex(int n, int m, const double *q, const double *c, const double *a,
                        const double *b, const double *init, double *ans, int *idx, int *term)
{
  int nmax = n+m, i, j, k, ma, nm, iter = 0, nhist = 0, caph = 16;
  double *qa = (double*)malloc(nmax*nmax*8), *xy = (double*)malloc(nmax*8), *cb = (double*)malloc(nmax*8);
  double *d = (double*)malloc(n*8), tempd, tempd2, objv;
  int *hist = (int*)malloc(caph*(m>0?m:1)*sizeof(int)); double *hmin = (double*)malloc(caph*8);
  for(i=0;i<n;i++) ans[i] = init ? init[i] : 0.0;
  /* _rkFDQPSolveASMInitIndex (rkfd_opt_qp.c:27-40) */
  for(ma=0,i=0;i<m;i++){ idx[i] = fabs(qp_cond(n,a,ans,i)-b[i]) < ORK_TOL; ma += idx[i]; }
  for(;;){
    int stepped = 0;
    if( ++iter > ORK_QP_MAX_ITER ){ if(term) *term = 2; break; }    /* safety net; the reference loop is unbounded */
    nm = n + ma;
    for(i=0;i<n;i++) for(j=0;j<n;j++) qa[nm*i+j] = -q[n*i+j];
    for(k=0,j=n;j<nm;j++){ while( k<m && !idx[k] ) k++; for(i=0;i<n;i++){ qa[nm*i+j] = a[n*k+i]; qa[nm*j+i] = a[n*k+i]; } k++; }
    for(i=n;i<nm;i++) for(j=n;j<nm;j++) qa[nm*i+j] = 0.0;
    for(i=0;i<n;i++) cb[i] = c[i];
    for(k=0,i=n;i<nm;i++){ while( k<m && !idx[k] ) k++; cb[i] = b[k]; k++; }
    ork_le_solve_mp_sym(nm,qa,cb,xy);
    for(i=0;i<n;i++) if( !(fabs(xy[i]-ans[i]) < ORK_TOL) ){ stepped = 1; break; }
    if( !stepped ){
      int neg = 0;
      for(i=0;i<n;i++) ans[i] = xy[i];
      for(i=0;i<ma;i++) if( xy[n+i] < 0 ){ neg = 1; break; }
      if( !neg ){ if(term) *term = 0; break; }                        /* optimal */
      tempd = xy[n]; for(i=1;i<ma;i++) if( xy[n+i] < tempd ) tempd = xy[n+i];
      for(k=0,i=0;i<m;i++) if( idx[i] ){
        if( fabs(xy[k+n]-tempd) < ORK_QP_ASM_TOL ){ idx[i] = 0; }
        k++; }
      for(ma=0,i=0;i<m;i++) ma += idx[i];
      continue;
    }
    /* STEP2: feasible direction and step length */
    for(i=0;i<n;i++) d[i] = xy[i]-ans[i];
    tempd = 1.0;
    for(i=0;i<m;i++){
      tempd2 = qp_cond(n,a,d,i);
      if( idx[i]==0 && tempd2 < 0 ){ tempd2 = ( b[i] - qp_cond(n,a,ans,i) ) / tempd2; if( tempd2 < tempd ) tempd = tempd2; }
    }
    for(i=0;i<n;i++) ans[i] += tempd*d[i];
    for(i=0;i<m;i++) if( idx[i]==0 && fabs(qp_cond(n,a,ans,i)-b[i]) < ORK_TOL ){ idx[i] = 1; ma++; }
    /* anti-cycling (rkfd_opt_qp.c:152-171) */
    objv = 0; for(i=0;i<n;i++){ double s=0; for(j=0;j<n;j++) s += q[n*i+j]*ans[j]; objv += 0.5*ans[i]*s + c[i]*ans[i]; }
    { int end = 0;
      for(k=0;k<nhist;k++){ int same = 1; for(i=0;i<m;i++) if( idx[i] != hist[m*k+i] ){ same = 0; break; }
        if( !same ) continue;
        if( fabs(hmin[k]/objv - 1.0) > ORK_QP_ASM_TOL ) continue;
        end = 1; break; }
      if( end ){ if(term) *term = 1; break; } }
    if( nhist == caph ){ caph *= 2; hist = (int*)realloc(hist,caph*(m>0?m:1)*sizeof(int)); hmin = (double*)realloc(hmin,caph*8); }
    for(i=0;i<m;i++) hist[m*nhist+i] = idx[i];
    hmin[nhist++] = objv;
  }
  free(qa); free(xy); free(cb); free(d); free(hist); free(hmin);
  return iter;
}

/* ------------------------------------------------------------------------------------ */
/* rigid contact: Vert (rkfd_vert.c:327-336) and MLCP (rkfd_mlcp.c:287-297) */
static void solver_rigid(ork_env *e, int do_up_ref)
{
  const ork_world *w = e->w; int pi, k, N = 0, n, i, j, col;
  int *slot = (int*)malloc((w->nslot>0?w->nslot:1)*sizeof(int)), *plink = (int*)malloc((w->nslot>0?w->nslot:1)*sizeof(int));
  int *blink = (int*)malloc((w->nslot>0?w->nslot:1)*sizeof(int));      /* the partner's link when it moves, else -1 */
  int *ppair = (int*)malloc((w->nslot>0?w->nslot:1)*sizeof(int));
  const ork_cinfo **pci = (const ork_cinfo**)malloc((w->nslot>0?w->nslot:1)*sizeof(void*));
  double *A = e->rA, *b = e->rb, *f = e->rf, dt = w->dt;
  /* _rkFDSolverCountContacts (rkfd_vert.c:22-29): contacts in (pair, vertex) order */
  for(pi=0;pi<w->npair;pi++){
    const ork_pair *p = &w->pair[pi]; const ork_cell *cl = &w->cell[p->cell];
    if( p->ci.type != ORK_CONTACT_RIGID ) continue;
    for(k=0;k<cl->nvert;k++) if( e->c_active[p->sofs+k] ){ slot[N]=p->sofs+k; plink[N]=cl->link; blink[N]=w->box[p->box].link; ppair[N]=pi; pci[N]=&p->ci; N++; }
  }
  e->rn = 0;
  if( N == 0 ){ free(slot); free(plink); free(blink); free(ppair); free(pci); return; }
  n = 3*N; e->rn = n;
  /* rkFDUpdateAccBias (rkfd_util.c:149-161): full ABA with the friction / penalty wrenches, save */
  aba_backward(e); aba_forward(e,NULL,0);
  for(i=0;i<w->nl;i++) memcpy(e->lw[i].a0,e->lw[i].a,sizeof e->lw[i].a);
  /* _rkFDSolverBiasAcc (rkfd_vert.c:107-123 / rkfd_mlcp.c:58-74) */
  /* rkFDChainPointRelativeAcc (rkfd_util.c:103-118): the vertex's link minus the partner's (0 for a static partner) */
  for(k=0;k<N;k++){ double av[3]; link_point_wld_acc(&e->lw[plink[k]],e->c_vert+3*slot[k],av);
    if( blink[k] >= 0 ){ double ab[3]; link_point_wld_acc(&e->lw[blink[k]],e->c_vert+3*slot[k],ab); v3_sub(av,ab,av); }
    for(i=0;i<3;i++) b[3*k+i] = v3_dot(e->c_axis+9*slot[k]+3*i,av); }
  /* _rkFDSolverRelationAccForce (rkfd_vert.c:153-185 / rkfd_mlcp.c:104-142): 3N cached-ABA probes */
  for(k=0;k<N;k++) for(i=0;i<3;i++){
    col = 3*k+i;
    aba_probe2(e,plink[k],blink[k],e->c_vert+3*slot[k],e->c_axis+9*slot[k]+3*i);
    for(j=0;j<N;j++){ double av[3]; int ii; link_point_wld_acc(&e->lw[plink[j]],e->c_vert+3*slot[j],av);
      if( blink[j] >= 0 ){ double ab[3]; link_point_wld_acc(&e->lw[blink[j]],e->c_vert+3*slot[j],ab); v3_sub(av,ab,av); }
      for(ii=0;ii<3;ii++) A[n*(3*j+ii)+col] = v3_dot(e->c_axis+9*slot[j]+3*ii,av) - b[3*j+ii]; }
  }
  /* rkFDChainRestoreABIAccBiasPair */
  for(i=0;i<w->nl;i++) memcpy(e->lw[i].a,e->lw[i].a0,sizeof e->lw[i].a);
  /* _rkFDSolverBiasVel (rkfd_vert.c:189-206 / rkfd_mlcp.c:146-162) */
  for(i=0;i<n;i++) b[i] *= dt;
  for(k=0;k<N;k++){ double *vel = e->c_vel+3*slot[k];
    link_point_wld_vel(&e->lw[plink[k]],e->c_vert+3*slot[k],vel);
    if( blink[k] >= 0 ){ double vb[3]; link_point_wld_vel(&e->lw[blink[k]],e->c_vert+3*slot[k],vb); v3_sub(vel,vb,vel); }
    add_slide_vel(e,ppair[k],e->c_vert+3*slot[k],e->c_norm+3*slot[k],vel);
    for(i=0;i<3;i++) b[3*k+i] += v3_dot(vel,e->c_axis+9*slot[k]+3*i); }

  if( w->solver == ORK_SOLVER_MLCP ){
    int cnt; double ff[2], fs, fnorm;
    /* _rkFDSolverRelaxationCompensation (rkfd_mlcp.c:164-188) */
    for(k=0;k<N;k++){ double d[3], mu; int s = slot[k];
      v3_sub(e->c_vert+3*s,e->c_refw+3*s,d);
      for(i=0;i<3;i++) A[n*(3*k+i)+3*k+i] += pci[k]->L;
      mu = e->c_type[s]==ORK_SF ? pci[k]->SF : pci[k]->KF;
      b[3*k  ] += pci[k]->K      * v3_dot(d,e->c_axis+9*s);
      b[3*k+1] += pci[k]->K * mu * v3_dot(d,e->c_axis+9*s+3);
      b[3*k+2] += pci[k]->K * mu * v3_dot(d,e->c_axis+9*s+6); }
    /* _rkFDSolverMLCP (rkfd_mlcp.c:190-249): projected Gauss-Seidel.  The friction sweep reads
     * rows offset+0 and offset+1 (NOT +1,+2) exactly as the reference does (:219-225). */
    for(i=0;i<n;i++) f[i] = 0.0;
    for(cnt=0;cnt<w->max_iter;cnt++){
      for(k=0;k<N;k++){ int o = 3*k; double s = 0; for(j=0;j<n;j++) s += A[n*o+j]*f[j];
        ff[0] = -( b[o] + s - A[n*o+o]*f[o] ) / A[n*o+o];
        f[o] = ff[0] < ORK_TOL ? 0.0 : ff[0]; }
      for(k=0;k<N;k++){ int o = 3*k;
        for(i=0;i<2;i++){
          if( fabs(A[n*(o+i)+o+i]) < ORK_TOL ) ff[i] = 0;
          else{ double s = 0; for(j=0;j<n;j++) s += A[n*(o+i)+j]*f[j];
            ff[i] = -( b[o+i] + s - A[n*(o+i)+o+i]*f[o+i] ) / A[n*(o+i)+o+i]; }
        }
        fnorm = ff[0]*ff[0] + ff[1]*ff[1];
        fs = e->c_type[slot[k]]==ORK_SF ? (pci[k]->SF*f[o])*(pci[k]->SF*f[o]) : (pci[k]->KF*f[o])*(pci[k]->KF*f[o]);
        if( fnorm < ORK_TOL || fs < ORK_TOL ){ f[o+1] = 0.0; f[o+2] = 0.0; }
        else if( fnorm > fs ){ fs /= fnorm; f[o+1] = ff[0]*fs; f[o+2] = ff[1]*fs; }
        else { f[o+1] = ff[0]; f[o+2] = ff[1]; }
      }
    }
    for(i=0;i<n;i++) f[i] /= dt;
    /* _rkFDSolverSetForce (rkfd_mlcp.c:252-284): commits the friction type regardless of doUpRef and
     * tests WORLD components of f (f.e[0] as "normal") - both mirrored */
    for(k=0;k<N;k++){ int s = slot[k]; double *fw = e->c_f+3*s, fn, fss, mu;
      fw[0]=fw[1]=fw[2]=0; for(i=0;i<3;i++) v3_cat(fw,f[3*k+i],e->c_axis+9*s+3*i);
      push_wrench(e,plink[k],e->c_vert+3*s,fw);
      if( blink[k] >= 0 ){ double fr[3] = { -fw[0], -fw[1], -fw[2] }; push_wrench(e,blink[k],e->c_vert+3*s,fr); }
      fn = fw[0]; fss = sqrt(fw[1]*fw[1]+fw[2]*fw[2]);
      mu = e->c_type[s]==ORK_SF ? pci[k]->SF : pci[k]->KF;
      if( fss > mu*fn - ORK_TOL ){ e->c_type[s] = ORK_KF; v3_copy(e->c_pro+3*s,e->c_ref+3*s); }
      else { e->c_type[s] = ORK_SF; update_ref_slide(e,ppair[k],s); } }       /* in EVERY evaluation, as the reference does (rkfd_mlcp.c:275-280) */
  } else {
    /* Vert: _rkFDSolverFrictionConstraint (rkfd_vert.c:73-103), _rkFDSolverCompensateDepth (:208-232),
     * _rkFDSolverQP (:258-283) */
    int pyr = w->pyramid, m = pyr*N; double *nf = (double*)calloc(m*n,8), *dz = (double*)calloc(m,8);
    double *Q = (double*)malloc(n*n*8), *c = (double*)malloc(n*8), *c2 = (double*)malloc(n*8), *init = (double*)calloc(n,8);
    int *idx = (int*)calloc(m,sizeof(int));
    for(k=0;k<N;k++){ double fric = ( e->c_type[slot[k]]==ORK_KF ? pci[k]->KF : pci[k]->SF ) * w->sc_table[1][0];
      for(i=0;i<pyr;i++){ nf[n*(pyr*k+i)+3*k] = fric; nf[n*(pyr*k+i)+3*k+1] = w->sc_table[0][i]; nf[n*(pyr*k+i)+3*k+2] = w->sc_table[1][i]; } }
    for(k=0;k<N;k++){ double d[3], fric, K; int s = slot[k];
      v3_sub(e->c_vert+3*s,e->c_refw+3*s,d);
      fric = e->c_type[s]==ORK_KF ? pci[k]->KF : pci[k]->SF; K = pci[k]->K;
      c[3*k  ] = b[3*k  ] + K        * v3_dot(d,e->c_axis+9*s);
      c[3*k+1] = b[3*k+1] + K * fric * v3_dot(d,e->c_axis+9*s+3);
      c[3*k+2] = b[3*k+2] + K * fric * v3_dot(d,e->c_axis+9*s+6); }
    for(i=0;i<n;i++) for(j=0;j<n;j++){ double s=0; int kk; for(kk=0;kk<n;kk++) s += A[n*kk+i]*A[n*kk+j]; Q[n*i+j]=s; }
    for(i=0;i<n;i++){ double s=0; for(j=0;j<n;j++) s += A[n*j+i]*c[j]; c2[i]=s; }
    for(k=0;k<N;k++) for(i=0;i<3;i++) Q[n*(3*k+i)+3*k+i] += pci[k]->L;
    for(k=0;k<N;k++) init[3*k] = 1.0;          /* _rkFDSolverQPASMInit (rkfd_vert.c:234-244) */
    { int term = 0, it = ork_qp_solve_asm_ex(n,m,Q,c2,nf,dz,init,f,idx,&term);
      e->qp_n = n; e->qp_m = m; e->qp_iter = it; e->qp_term = term;
      e->qp_Q = (double*)realloc(e->qp_Q,n*n*8); e->qp_c = (double*)realloc(e->qp_c,n*8); e->qp_nf = (double*)realloc(e->qp_nf,m*n*8);
      e->qp_x = (double*)realloc(e->qp_x,n*8); e->qp_idx = (int*)realloc(e->qp_idx,m*sizeof(int));
      memcpy(e->qp_Q,Q,n*n*8); memcpy(e->qp_c,c2,n*8); memcpy(e->qp_nf,nf,m*n*8); memcpy(e->qp_x,f,n*8); memcpy(e->qp_idx,idx,m*sizeof(int)); }
    for(i=0;i<n;i++) f[i] /= dt;
    /* _rkFDSolverSetForce (rkfd_vert.c:286-324) */
    for(k=0;k<N;k++){ int s = slot[k], flag = 0; double *fw = e->c_f+3*s;
      fw[0]=fw[1]=fw[2]=0; for(i=0;i<3;i++) v3_cat(fw,f[3*k+i],e->c_axis+9*s+3*i);
      push_wrench(e,plink[k],e->c_vert+3*s,fw);
      if( blink[k] >= 0 ){ double fr[3] = { -fw[0], -fw[1], -fw[2] }; push_wrench(e,blink[k],e->c_vert+3*s,fr); }
      if( do_up_ref ){
        for(i=0;i<pyr;i++) if( idx[pyr*k+i] ){ flag = 1; break; }
        if( flag ){ e->c_type[s] = ORK_KF; v3_copy(e->c_pro+3*s,e->c_ref+3*s); }
        else { e->c_type[s] = ORK_SF; update_ref_slide(e,ppair[k],s); } } }
    free(nf); free(dz); free(Q); free(c); free(c2); free(init); free(idx);
  }
  free(slot); free(plink); free(blink); free(ppair); free(pci);
}

/* ------------------------------------------------------------------------------------ */
/* Volume solver (rkfd_volume.c).  [EXT A-15] rkCDColVolBREPVert (RoKi, absent) is restricted to
 * (moving cell with the 8 corners of a parallelepiped, vertex k = sign bits x:k&1 y:k&2 z:k&4) x
 * (ONE face half-space of a static box, the face of the first vertex found inside):
 *   colvol = cell polyhedron clipped by that half-space (triangulated: every clipped triangle as a fan,
 *            the cut polygon as a fan over its angularly sorted corners),
 *   norm   = outward normal of that face, axis[] = the face frame of the vertex test (A-10),
 *   center = barycentre of colvol.
 * Every integral of rkfd_volume.c:397-491 is a midpoint rule on polynomials of degree <= 2, i.e.
 * independent of how the faces are triangulated. */
#define VOL_MAXTRI 96
#define VOL_MAXPL 48
typedef struct { double v[3], norm[3], r[2], s[2]; } ork_vplane;
typedef struct {
  int pair, link; const ork_cinfo *ci;
  double center[3], norm[3], axis[9];
  int ntri; double tri[VOL_MAXTRI][9], tn[VOL_MAXTRI][3];
  int npl; ork_vplane pl[VOL_MAXPL];      /* in zListForEach order (tail first) */
  double wrench[6];
} ork_vpair;

static const int ork_box_quads[6][4] = { {1,3,7,5}, {0,4,6,2}, {2,6,7,3}, {0,1,5,4}, {4,5,7,6}, {0,2,3,1} };

static void vol_emit(ork_vpair *vp, const double *a, const double *b, const double *c, const double *n)
{
  double e1[3], e2[3], cr[3]; double *t;
  if( vp->ntri >= VOL_MAXTRI ) return;
  t = vp->tri[vp->ntri];
  v3_sub(b,a,e1); v3_sub(c,a,e2); v3_cross(e1,e2,cr);
  v3_copy(a,t);
  if( v3_dot(cr,n) >= 0 ){ v3_copy(b,t+3); v3_copy(c,t+6); } else { v3_copy(c,t+3); v3_copy(b,t+6); }
  v3_copy(n,vp->tn[vp->ntri]);
  vp->ntri++;
}

/* colvol, center of one colliding pair; returns 0 when the clipped volume is empty */
static int vol_build(ork_env *e, int pi, ork_vpair *vp)
{
  const ork_world *w = e->w; const ork_pair *p = &w->pair[pi]; const ork_cell *cl = &w->cell[p->cell];
  const ork_box *bx = &w->box[p->box]; const ork_lw *x = &e->lw[cl->link];
  double vw[8][3], cen[3] = {0,0,0}, p0[3], d[8], cut[64][3], vol = 0, bc[3] = {0,0,0}; int ncut = 0, k, s0 = -1, f, i, j;
  if( !cl->volbox ) return 0;      /* [EXT A-15]: contact volumes of box cells only; other rigid cells are watched, not solved */
  for(k=0;k<8;k++) if( e->c_active[p->sofs+k] ){ s0 = p->sofs+k; break; }
  if( s0 < 0 ) return 0;
  vp->pair = pi; vp->link = cl->link; vp->ci = &p->ci; vp->ntri = 0; vp->npl = 0;
  v3_copy(e->c_norm+3*s0,vp->norm); memcpy(vp->axis,e->c_axis+9*s0,9*sizeof(double));
  m3_mulv(bx->R,e->c_pro+3*s0,p0); v3_add(p0,bx->p,p0);          /* a point of the face plane */
  for(k=0;k<8;k++){ m3_mulv(x->Rw,w->vert+3*(cl->vofs+k),vw[k]); v3_add(vw[k],x->pw,vw[k]); v3_cat(cen,0.125,vw[k]);
    { double t[3]; v3_sub(vw[k],p0,t); d[k] = v3_dot(vp->norm,t); if( fabs(d[k]) <= ORK_TOL ) d[k] = 0.0; } }
  for(f=0;f<6;f++){
    const int *qd = ork_box_quads[f]; double e1[3], e2[3], fn[3], t[3], nn; int tr;
    v3_sub(vw[qd[1]],vw[qd[0]],e1); v3_sub(vw[qd[3]],vw[qd[0]],e2); v3_cross(e1,e2,fn);
    nn = v3_norm(fn); if( nn == 0 ) continue;
    fn[0]/=nn; fn[1]/=nn; fn[2]/=nn;
    v3_sub(vw[qd[0]],cen,t); if( v3_dot(fn,t) < 0 ){ fn[0]=-fn[0]; fn[1]=-fn[1]; fn[2]=-fn[2]; }
    for(tr=0;tr<2;tr++){
      int id[3] = { qd[0], qd[1+tr], qd[2+tr] }; double poly[4][3]; int np = 0;
      /* Sutherland-Hodgman against d <= 0 */
      for(i=0;i<3;i++){
        int a = id[i], b = id[(i+1)%3];
        if( d[a] <= 0 ){ v3_copy(vw[a],poly[np++]); if( d[a] == 0 && ncut < 64 ) v3_copy(vw[a],cut[ncut++]); }
        if( (d[a] < 0 && d[b] > 0) || (d[a] > 0 && d[b] < 0) ){
          double tt = d[a]/(d[a]-d[b]);
          for(j=0;j<3;j++) poly[np][j] = vw[a][j] + tt*(vw[b][j]-vw[a][j]);
          if( ncut < 64 ) v3_copy(poly[np],cut[ncut++]);
          np++;
        }
      }
      for(i=1;i+1<np;i++) vol_emit(vp,poly[0],poly[i],poly[i+1],fn);
    }
  }
  /* the cut polygon: distinct cut points sorted by angle about their mean in the face frame */
  { double u[64][3]; double ang[64], m[3] = {0,0,0}; int nu = 0;
    for(i=0;i<ncut;i++){ int dup = 0;
      for(j=0;j<nu;j++){ double t[3]; v3_sub(cut[i],u[j],t); if( v3_norm(t) < 1.0e-10 ){ dup = 1; break; } }
      if( !dup ) v3_copy(cut[i],u[nu++]); }
    if( nu >= 3 ){
      for(i=0;i<nu;i++) v3_cat(m,1.0/nu,u[i]);
      for(i=0;i<nu;i++){ double t[3]; v3_sub(u[i],m,t); ang[i] = atan2(v3_dot(t,vp->axis+6),v3_dot(t,vp->axis+3)); }
      for(i=1;i<nu;i++){ double a = ang[i], t[3]; v3_copy(u[i],t);
        for(j=i-1;j>=0 && ang[j] > a;j--){ ang[j+1] = ang[j]; v3_copy(u[j],u[j+1]); }
        ang[j+1] = a; v3_copy(t,u[j+1]); }
      for(i=1;i+1<nu;i++) vol_emit(vp,u[0],u[i],u[i+1],vp->norm);
    } }
  /* barycentre (signed tetrahedra about a point of the clipping plane: a thin volume is then a sum of thin tetrahedra
   * instead of the small difference of large ones; triangles are outward oriented) */
  for(i=0;i<vp->ntri;i++){ double a[3], b[3], c[3], cr[3], v6;
    v3_sub(vp->tri[i],p0,a); v3_sub(vp->tri[i]+3,p0,b); v3_sub(vp->tri[i]+6,p0,c);
    v3_cross(b,c,cr); v6 = v3_dot(a,cr)/6.0; vol += v6;
    for(j=0;j<3;j++) bc[j] += v6*0.25*(a[j]+b[j]+c[j]); }
  if( !(vol > 1.0e-18) ) return 0;
  for(j=0;j<3;j++) vp->center[j] = p0[j] + bc[j]/vol;
  return 1;
}

/* _rkFDSolverSetContactPlane (rkfd_volume.c:350-374) */
static void vol_set_plane(ork_vpair *vp, const double *p, const double *fnorm)
{
  double tmpv[3], nn, cand_n[3], t[3]; int i;
  v3_copy(fnorm,tmpv); v3_cat(tmpv,-v3_dot(fnorm,vp->norm),vp->norm);
  if( fabs(tmpv[0]) < ORK_TOL && fabs(tmpv[1]) < ORK_TOL && fabs(tmpv[2]) < ORK_TOL ) return;   /* zVec3DIsTiny */
  nn = -v3_norm(tmpv); cand_n[0]=tmpv[0]/nn; cand_n[1]=tmpv[1]/nn; cand_n[2]=tmpv[2]/nn;
  for(i=0;i<vp->npl;i++){ ork_vplane *q = &vp->pl[i];
    v3_sub(cand_n,q->norm,t);
    if( fabs(t[0]) < 1e-8 && fabs(t[1]) < 1e-8 && fabs(t[2]) < 1e-8 ){
      v3_sub(q->v,p,t);
      if( fabs(v3_dot(cand_n,t)) < 1e-8 ){
        double cr[3]; v3_cross(cand_n,t,cr);
        if( v3_dot(vp->norm,cr) > 0.0 ) v3_copy(p,q->v);
        return; } } }
  if( vp->npl >= VOL_MAXPL ) return;
  v3_copy(p,vp->pl[vp->npl].v); v3_copy(cand_n,vp->pl[vp->npl].norm); vp->npl++;   /* zListInsertHead: last in zListForEach order */
}

static void vol_cc(const double *pm_in, const double *h, double K, double s, const double *norm, double *cc)
{ /* _rkFDSolverConstraintMidDepth + _rkFDSolverConstraintDepth (rkfd_volume.c:232-241, 296-310) */
  double k = K*s/6.0, hm[3], hc, t[3], pmh[3]; int i, j;
  hm[0] = k*(h[0]+h[1]); hm[1] = k*(h[1]+h[2]); hm[2] = k*(h[0]+h[2]); hc = k*(h[0]+h[1]+h[2])*2;
  for(j=0;j<3;j++){ cc[j] = -hc*norm[j]; cc[3+j] = 0; }
  for(i=0;i<3;i++){ for(j=0;j<3;j++) pmh[j] = pm_in[3*i+j]*hm[i]; v3_cross(norm,pmh,t); v3_add(cc+3,t,cc+3); }
}
static double vol_area(const double *p)
{ double e1[3], e2[3], cr[3]; v3_sub(p+3,p,e1); v3_sub(p+6,p,e2); v3_cross(e1,e2,cr); return 0.5*v3_norm(cr); }
static void vol_mid(const double *p, double *pm)
{ int j; for(j=0;j<3;j++){ pm[j] = 0.5*(p[j]+p[3+j]); pm[3+j] = 0.5*(p[3+j]+p[6+j]); pm[6+j] = 0.5*(p[6+j]+p[j]); } }
/* _rkFDSolverConstraintInnerPoint (rkfd_volume.c:331-348) */
static void vol_inner(const double *p1, const double *p2, double h1, double h2, double *pp)
{ int j;
  if( fabs(h1) < ORK_TOL ){ v3_copy(p1,pp); return; }
  if( fabs(h2) < ORK_TOL ){ v3_copy(p2,pp); return; }
  if( fabs(h2-h1) < ORK_TOL ){ for(j=0;j<3;j++) pp[j] = 0.5*(p1[j]+p2[j]); return; }
  for(j=0;j<3;j++) pp[j] = p1[j]*(h2/(h2-h1)) + (h1/(h1-h2))*p2[j];
}
/* _rkFDSolverConstraint (rkfd_volume.c:397-491): q (6x6 row-major, element (i,j) = _zMat6DElem), c */
static void vol_constraint(ork_vpair *vp, double *q, double *c)
{
  int i, j, a, b;
  memset(q,0,36*8); memset(c,0,6*8);
  for(i=0;i<vp->ntri;i++){
    double pf[9], p[9], pm[9], pp[6], h[3], s, cc[6], pc[3], S[9]; int st = 0, stp[3] = {0,0,0};
    for(j=0;j<3;j++){ v3_sub(vp->tri[i]+3*j,vp->center,pf+3*j); h[j] = v3_dot(vp->norm,pf+3*j);
      v3_copy(pf+3*j,p+3*j); v3_cat(p+3*j,-h[j],vp->norm); }
    vol_mid(p,pm); s = vol_area(p);
    /* _rkFDSolverConstraintAddQ (:279-294) */
    for(j=0;j<3;j++) pc[j] = (s/3.0)*(p[j]+p[3+j]+p[6+j]);
    for(j=0;j<3;j++) q[6*j+j] += s;
    m3_skew(pc,S);                      /* [pc x] */
    for(a=0;a<3;a++) for(b=0;b<3;b++){ q[6*(3+a)+b] += S[3*a+b]; q[6*a+3+b] -= S[3*a+b]; }
    for(j=0;j<3;j++){ double M[9], M2[9]; m3_skew(pm+3*j,M); m3_mul(M,M,M2);     /* [pm x][pm x] */
      for(a=0;a<3;a++) for(b=0;b<3;b++) q[6*(3+a)+3+b] -= (s/3.0)*M2[3*a+b]; }
    vol_cc(pm,h,vp->ci->K,s,vp->norm,cc);
    /* _rkFDSolverConstraintSignDepth (:312-329) */
    for(j=0;j<3;j++){
      if( h[j] > ORK_TOL ){ st += 1<<(j*2); stp[1] = j; }
      else if( h[j] < -ORK_TOL ){ st += 1<<(j*2+1); stp[2] = j; }
      else stp[0] = j; }
    switch( st ){
    case 0x01: case 0x04: case 0x10: case 0x05: case 0x11: case 0x14:
      vol_set_plane(vp,pf+3*stp[0],vp->tn[i]);  /* fall through */
    case 0x15:
      for(j=0;j<6;j++) c[j] += cc[j];
      continue;
    case 0x02: case 0x08: case 0x20: case 0x0a: case 0x22: case 0x28:
      vol_set_plane(vp,pf+3*stp[0],vp->tn[i]);  /* fall through */
    case 0x2a:
      for(j=0;j<6;j++) c[j] -= cc[j];
      continue;
    case 0x24: case 0x12: case 0x09:
      for(j=0;j<6;j++) c[j] += cc[j];
      vol_inner(pf+3*stp[1],pf+3*stp[2],h[stp[1]],h[stp[2]],p+3*stp[1]);
      h[stp[1]] = 0.0;
      v3_copy(p+3*stp[0],pp); v3_copy(p+3*stp[1],pp+3);
      break;
    case 0x06: case 0x21: case 0x18:
      for(j=0;j<6;j++) c[j] += cc[j];
      vol_inner(pf+3*stp[1],pf+3*stp[2],h[stp[1]],h[stp[2]],p+3*stp[2]);
      h[stp[2]] = 0.0;
      v3_copy(p+3*stp[2],pp); v3_copy(p+3*stp[0],pp+3);
      break;
    case 0x16: case 0x19: case 0x25:
      stp[0] = (stp[2]+1)%3; stp[1] = (stp[0]+1)%3;
      for(j=0;j<6;j++) c[j] += cc[j];
      vol_inner(pf+3*stp[2],pf+3*stp[0],h[stp[2]],h[stp[0]],p+3*stp[0]);
      vol_inner(pf+3*stp[2],pf+3*stp[1],h[stp[2]],h[stp[1]],p+3*stp[1]);
      h[stp[0]] = h[stp[1]] = 0.0;
      v3_copy(p+3*stp[0],pp); v3_copy(p+3*stp[1],pp+3);
      break;
    case 0x1a: case 0x26: case 0x29:
      stp[0] = (stp[1]+1)%3; stp[2] = (stp[0]+1)%3;
      for(j=0;j<6;j++) c[j] -= cc[j];
      vol_inner(pf+3*stp[1],pf+3*stp[0],h[stp[1]],h[stp[0]],p+3*stp[0]);
      vol_inner(pf+3*stp[1],pf+3*stp[2],h[stp[1]],h[stp[2]],p+3*stp[2]);
      h[stp[0]] = h[stp[2]] = 0.0;
      v3_copy(p+3*stp[2],pp); v3_copy(p+3*stp[0],pp+3);
      break;
    default:
      continue;
    }
    s = vol_area(p); vol_mid(p,pm);
    vol_cc(pm,h,vp->ci->K,s,vp->norm,cc);
    switch( st ){
    case 0x1a: case 0x26: case 0x29: for(j=0;j<6;j++) c[j] += 2.0*cc[j]; break;
    default: for(j=0;j<6;j++) c[j] -= 2.0*cc[j];
    }
    vol_set_plane(vp,pp,vp->tn[i]);
  }
  /* rkCDPlaneListQuickSort with __rk_fd_plane_cmp (:376-395): ascending angle key along zListForEach
   * order ([EXT] sort direction assumed); insertion sort, ties (|dth| tiny) keep their order */
  { double th[VOL_MAXPL]; const double *n = vp->axis, *a1 = vp->axis+3;
    for(i=0;i<vp->npl;i++){ double t[3]; v3_cross(a1,vp->pl[i].norm,t);
      th[i] = v3_dot(t,n) > 0 ? atan2(-v3_norm(t),v3_dot(a1,vp->pl[i].norm)) : atan2(v3_norm(t),v3_dot(a1,vp->pl[i].norm)); }
    for(i=1;i<vp->npl;i++){ ork_vplane t = vp->pl[i]; double a = th[i];
      for(j=i-1;j>=0 && !(fabs(th[j]-a) < ORK_TOL) && th[j] > a;j--){ vp->pl[j+1] = vp->pl[j]; th[j+1] = th[j]; }
      vp->pl[j+1] = t; th[j+1] = a; } }
}

/* [EXT A-16] zLPSolveSimplex / zLPFeasibleBase (ZM, absent): min c^T x  s.t.  A x = b, x >= 0 by the two-phase
 * tableau simplex with Bland's rule.  c == NULL: feasibility only.  Returns 1 on success. */
static int vol_lp(int m, int n, const double *A, const double *b, const double *c, double *x)
{
  int nt = n + m, i, j, r, it, ok = 1, phase; int *basis = (int*)malloc(m*sizeof(int));
  double *T = (double*)malloc((size_t)m*(nt+1)*8), *cost = (double*)malloc(nt*8), bmax = 0, eps;
  for(i=0;i<m;i++){ double sg = b[i] < 0 ? -1.0 : 1.0;
    for(j=0;j<n;j++) T[(nt+1)*i+j] = sg*A[n*i+j];
    for(j=0;j<m;j++) T[(nt+1)*i+n+j] = i==j ? 1.0 : 0.0;
    T[(nt+1)*i+nt] = sg*b[i]; basis[i] = n+i; if( fabs(b[i]) > bmax ) bmax = fabs(b[i]); }
  eps = 1.0e-10*(1.0+bmax);
  for(phase=1;phase<=2 && ok;phase++){
    int ncol = phase==1 ? nt : n;
    if( phase == 2 && !c ) break;
    for(j=0;j<nt;j++) cost[j] = phase==1 ? ( j>=n ? 1.0 : 0.0 ) : ( j<n ? c[j] : 0.0 );
    for(it=0;it<20000;it++){
      int enter = -1, leave = -1; double best = 0;
      for(j=0;j<ncol && enter<0;j++){ double rc = cost[j]; int bas = 0;
        for(i=0;i<m;i++){ if( basis[i]==j ) bas = 1; rc -= cost[basis[i]]*T[(nt+1)*i+j]; }
        if( !bas && rc < -1.0e-11 ) enter = j; }
      if( enter < 0 ) break;
      for(i=0;i<m;i++){ double a = T[(nt+1)*i+enter];
        if( a > 1.0e-11 ){ double ratio = T[(nt+1)*i+nt]/a;
          if( leave < 0 || ratio < best - 1.0e-13 || ( fabs(ratio-best) <= 1.0e-13 && basis[i] < basis[leave] ) ){ leave = i; best = ratio; } } }
      if( leave < 0 ){ ok = 0; break; }          /* unbounded */
      { double pv = T[(nt+1)*leave+enter];
        for(j=0;j<=nt;j++) T[(nt+1)*leave+j] /= pv;
        for(r=0;r<m;r++) if( r != leave ){ double fct = T[(nt+1)*r+enter]; if( fct != 0 ) for(j=0;j<=nt;j++) T[(nt+1)*r+j] -= fct*T[(nt+1)*leave+j]; }
        basis[leave] = enter; }
    }
    if( phase == 1 ){ double art = 0;
      for(i=0;i<m;i++) if( basis[i] >= n ) art += T[(nt+1)*i+nt];
      if( art > eps ) ok = 0;
      else for(i=0;i<m;i++) if( basis[i] >= n ){   /* drive a degenerate artificial out of the basis if possible */
        for(j=0;j<n;j++) if( fabs(T[(nt+1)*i+j]) > 1.0e-9 ) break;
        if( j < n ){ double pv = T[(nt+1)*i+j]; int jj;
          for(jj=0;jj<=nt;jj++) T[(nt+1)*i+jj] /= pv;
          for(r=0;r<m;r++) if( r != i ){ double fct = T[(nt+1)*r+j]; if( fct != 0 ) for(jj=0;jj<=nt;jj++) T[(nt+1)*r+jj] -= fct*T[(nt+1)*i+jj]; }
          basis[i] = j; } } }
  }
  if( ok && x ){ for(j=0;j<n;j++) x[j] = 0; for(i=0;i<m;i++) if( basis[i] < n ) x[basis[i]] = T[(nt+1)*i+nt]; }
  free(T); free(cost); free(basis);
  return ok;
}

/* cached-ABA probe with a 6-D wrench (f, t world axes) at world point pos of `link` (rkfd_volume.c:176-211) */
static void aba_probe6(ork_env *e, int link, const double *pos_w, const double *f6)
{
  const ork_world *w = e->w; int i, j, k, r; ork_lw *x = &e->lw[link]; double pos[3], fl[3], tl[3], n[3];
  for(i=0;i<w->nl;i++){ memset(e->lw[i].du,0,sizeof e->lw[i].du); memset(e->lw[i].dp,0,sizeof e->lw[i].dp); }
  v3_sub(pos_w,x->pw,pos); m3_tmulv(x->Rw,pos,pos); m3_tmulv(x->Rw,f6,fl); m3_tmulv(x->Rw,f6+3,tl); v3_cross(pos,fl,n);
  for(r=0;r<3;r++){ x->dp[r] = -fl[r]; x->dp[3+r] = -(tl[r]+n[r]); }
  for(i=link;i>=0;i=w->link[i].parent){
    const ork_link *l = &w->link[i]; ork_lw *y = &e->lw[i]; int nd = e->nd[i]; double pa[6], t6[6];
    memcpy(pa,y->dp,sizeof pa);
    for(j=0;j<nd;j++){ double s=0; for(k=0;k<6;k++) s += y->S[6*k+j]*y->dp[k]; y->du[j] = -s; }
    for(j=0;j<nd;j++){ t6[j]=0; for(k=0;k<nd;k++) t6[j] += y->Dinv[6*j+k]*y->du[k]; }
    for(r=0;r<6;r++) for(j=0;j<nd;j++) pa[r] += y->U[6*r+j]*t6[j];
    if( l->parent < 0 ) break;
    m6_tmulv(y->X,pa,t6); for(r=0;r<6;r++) e->lw[l->parent].dp[r] += t6[r];
  }
  aba_forward(e,NULL,1);
}
int ork_lp_solve(int m, int n, const double *A, const double *b, const double *c, double *x){ return vol_lp(m,n,A,b,c,x); }
static int link_root(const ork_world *w, int l){ while( w->link[l].parent >= 0 ) l = w->link[l].parent; return l; }

static void solver_volume(ork_env *e, int do_up_ref)
{
  const ork_world *w = e->w; int pi, P = 0, n, m, i, j, k, col, off, colnum = 0, pyr = w->pyramid;
  ork_vpair *vp = (ork_vpair*)malloc(sizeof(ork_vpair)*(w->npair>0?w->npair:1));
  double dt = w->dt, *A, *b, *Q, *c, *nf, *dz, *init, *f, sn[64], cs[64]; int *idx;
  for(pi=0;pi<w->npair;pi++) e->v_np[pi] = -1;
  for(pi=0;pi<w->npair;pi++){
    if( w->pair[pi].ci.type != ORK_CONTACT_RIGID ) continue;
    if( vol_build(e,pi,&vp[P]) ) P++;
  }
  if( P == 0 ){ free(vp); return; }
  n = 6*P;
  A = (double*)calloc(n*n,8); b = (double*)calloc(n,8); Q = (double*)calloc(n*n,8); c = (double*)calloc(n,8);
  init = (double*)calloc(n,8); f = (double*)calloc(n,8);
  for(i=0;i<pyr && i<64;i++){ sn[i] = sin(2.0*M_PI/pyr*i); cs[i] = cos(2.0*M_PI/pyr*i); }   /* offset 0 (rkfd_volume.c:1000) */
  /* _rkFDSolverRelationAccForce (rkfd_volume.c:176-211) */
  aba_backward(e); aba_forward(e,NULL,0);
  for(i=0;i<w->nl;i++) memcpy(e->lw[i].a0,e->lw[i].a,sizeof e->lw[i].a);
  for(k=0;k<P;k++){ const ork_lw *x = &e->lw[vp[k].link];
    link_point_wld_acc(x,vp[k].center,b+6*k); m3_mulv(x->Rw,x->a+3,b+6*k+3); }
  for(k=0;k<P;k++) for(i=0;i<6;i++){
    double f6[6] = {0,0,0,0,0,0}; f6[i] = 1.0; col = 6*k+i;
    aba_probe6(e,vp[k].link,vp[k].center,f6);
    for(j=0;j<P;j++){
      if( link_root(w,vp[j].link) != link_root(w,vp[k].link) ){ int r; for(r=0;r<6;r++) A[n*(6*j+r)+col] = 0; }
      else { const ork_lw *x = &e->lw[vp[j].link]; double av[6]; int r;
        link_point_wld_acc(x,vp[j].center,av); m3_mulv(x->Rw,x->a+3,av+3);
        for(r=0;r<6;r++) A[n*(6*j+r)+col] = av[r] - b[6*j+r]; } }
  }
  for(i=0;i<w->nl;i++) memcpy(e->lw[i].a,e->lw[i].a0,sizeof e->lw[i].a);
  /* _rkFDSolverBiasVel (:214-226) */
  for(i=0;i<n;i++) b[i] *= dt;
  for(k=0;k<P;k++){ const ork_lw *x = &e->lw[vp[k].link]; double v6[6];
    link_point_wld_vel(x,vp[k].center,v6); m3_mulv(x->Rw,x->v+3,v6+3);
    for(i=0;i<6;i++) b[6*k+i] += v6[i]; }
  /* _rkFDSolverQPCreate (:496-528) */
  for(k=0;k<P;k++){ double qv[36], cv[6], t6[6]; int r, s;
    vol_constraint(&vp[k],qv,cv);
    memcpy(e->v_qc+45*vp[k].pair,qv,36*8); memcpy(e->v_qc+45*vp[k].pair+36,cv,6*8); memcpy(e->v_qc+45*vp[k].pair+42,vp[k].norm,3*8);
    for(i=0;i<6;i++) for(j=0;j<6;j++){ double qe = qv[6*i+j];
      for(r=0;r<n;r++) for(s=0;s<n;s++) Q[n*r+s] += qe*A[n*(6*k+i)+r]*A[n*(6*k+j)+s]; }
    m6_mulv(qv,b+6*k,t6);
    for(i=0;i<6;i++){ cv[i] += t6[i]; for(r=0;r<n;r++) c[r] += cv[i]*A[n*(6*k+i)+r]; }
  }
  for(k=0;k<P;k++) for(i=0;i<6;i++) Q[n*(6*k+i)+6*k+i] += vp[k].ci->L;
  /* _rkFDSolverCountContacts, _rkFDSolverFrictionConstraint (:21-29, 121-138) */
  for(k=0;k<P;k++) colnum += vp[k].npl;
  m = P + colnum;
  nf = (double*)calloc((size_t)m*n,8); dz = (double*)calloc(m,8); idx = (int*)calloc(m,sizeof(int));
  { int io = 0, jo = 0;
    for(k=0;k<P;k++){ ork_vpair *v = &vp[k];
      for(j=0;j<3;j++) nf[n*io+jo+j] = v->norm[j];
      io++;
      for(i=0;i<v->npl;i++){ ork_vplane *pl = &v->pl[i]; double a = -v3_dot(pl->norm,pl->v), b1 = v3_dot(pl->norm,v->axis+6), b2 = -v3_dot(pl->norm,v->axis+3);
        for(j=0;j<3;j++){ nf[n*io+jo+j] = a*v->norm[j]; nf[n*io+jo+3+j] = b1*v->axis[3+j] + b2*v->axis[6+j]; }
        io++; }
      jo += 6; } }
  /* _rkFDSolverQP (:530-548) */
  for(k=0;k<P;k++) for(j=0;j<3;j++) init[6*k+j] = vp[k].norm[j];
  ork_qp_solve_asm(n,m,Q,c,nf,dz,init,f,idx);
  for(i=0;i<n;i++) f[i] /= dt;
  /* _rkFDSolverSetForce (:552-568); the offset is NOT advanced for a pair without planes - mirrored */
  off = 0;
  for(k=0;k<P;k++){ ork_vpair *v = &vp[k];
    if( v->npl == 0 ){ memset(v->wrench,0,sizeof v->wrench); continue; }
    memcpy(v->wrench,f+off,6*8);
    if( ( fabs(v->wrench[0]) < ORK_TOL && fabs(v->wrench[1]) < ORK_TOL && fabs(v->wrench[2]) < ORK_TOL ) || v3_dot(v->wrench,v->norm) < ORK_TOL )
      memset(v->wrench,0,sizeof v->wrench);
    off += 6; }
  /* _rkFDSolverModifyNormalForceCenter (:580-631) */
  for(k=0;k<P;k++){ ork_vpair *v = &vp[k]; double fn = v3_dot(v->norm,v->wrench), r0[3], r[3], dir[3], tmp[3], d, s; int flag = 0, np = v->npl, i0, i1, i2, i3, it;
    if( fn < ORK_TOL || np < 3 ) continue;
    for(j=0;j<3;j++) r0[j] = v->axis[3+j]*( -v3_dot(v->axis+6,v->wrench+3)/fn ) + v->axis[6+j]*( v3_dot(v->axis+3,v->wrench+3)/fn );
    i2 = np-1; i1 = np-2; i0 = np-3;
    for(it=0,i3=0;it<np;it++,i0=i1,i1=i2,i2=i3,i3=i3+1){
      int mod = 0;
      v3_sub(v->pl[i2].v,v->pl[i1].v,dir); d = v3_dot(dir,dir);
      if( fabs(d) < ORK_TOL ) continue;
      v3_sub(r0,v->pl[i1].v,tmp);
      if( v3_dot(tmp,v->pl[i1].norm) > ORK_TOL ) continue;
      s = v3_dot(dir,tmp)/d;
      if( s < ORK_TOL ){
        if( flag ) break;
        v3_sub(v->pl[i0].v,v->pl[i1].v,tmp); v3_add(tmp,dir,tmp);
        v3_copy(v->pl[i1].v,r); v3_cat(r,ORK_TOL/v3_norm(tmp),tmp); mod = 1;
      } else if( s < 1.0-ORK_TOL ){
        v3_copy(v->pl[i1].v,r); v3_cat(r,s,dir); v3_cat(r,ORK_TOL,v->pl[i1].norm); mod = 1;
      } else {
        const double *v3p = v->pl[i3 < np ? i3 : 0].v;      /* zListCellNext of the head is the root cell: index 0 assumed */
        v3_sub(v3p,v->pl[i2].v,tmp); v3_sub(tmp,dir,tmp);
        v3_copy(v->pl[i2].v,r); v3_cat(r,ORK_TOL/v3_norm(tmp),tmp); mod = 2;
      }
      { /* _rkFDSolverModifyNormForceCenterTrq (:572-578) */
        double na = v3_dot(v->norm,v->wrench+3), k1 = fn*v3_dot(v->axis+6,r), k2 = -fn*v3_dot(v->axis+3,r);
        for(j=0;j<3;j++) v->wrench[3+j] = na*v->norm[j] + k1*v->axis[3+j] + k2*v->axis[6+j]; }
      if( mod == 1 ) break;
      flag = 1;
    } }
  /* _rkFDSolverModifyWrench (:869-916) */
  for(k=0;k<P;k++){ ork_vpair *v = &vp[k]; double wv[6], fn, fs, tl = 0; int np = v->npl, kinetic = 0, setforce = 0;
    if( np == 0 ) continue;
    if( fabs(v3_dot(v->wrench,v->axis)) < ORK_TOL ) continue;
    for(i=0;i<3;i++){ wv[i] = v3_dot(v->wrench,v->axis+3*i); wv[i+3] = v3_dot(v->wrench+3,v->axis+3*i); }
    fn = wv[0]; fs = sqrt(wv[1]*wv[1]+wv[2]*wv[2]);
    for(i=0;i<np;i++){ ork_vplane *pl = &v->pl[i]; double rl;      /* _rkFDSolverPlaneVertPos (:700-713) */
      pl->r[0] = v3_dot(pl->v,v->axis+3); pl->r[1] = v3_dot(pl->v,v->axis+6);
      rl = sqrt(pl->r[0]*pl->r[0]+pl->r[1]*pl->r[1]); if( tl < rl ) tl = rl; }
    if( fabs(tl) < ORK_TOL ){
      wv[3] = wv[4] = wv[5] = 0;
      if( !(fabs(fs) < ORK_TOL) && fs > v->ci->SF*fn ){   /* _rkFDSolverModifyWrenchKineticCenter (:715-731) */
        double vel[3], nv; link_point_wld_vel(&e->lw[v->link],v->center,vel);
        v3_cat(vel,-v3_dot(v->norm,vel),v->norm); nv = v3_norm(vel);
        if( fabs(nv) < ORK_TOL ){ wv[1] = 0; wv[2] = 0; }
        else { double t = kinetic_friction_weight(w->friction_weight,nv)*v->ci->KF*wv[0]/nv;
          wv[1] = -t*v3_dot(vel,v->axis+3); wv[2] = -t*v3_dot(vel,v->axis+6); }
        kinetic = 1;
      }
      setforce = 1;
    } else if( ( !(fabs(fs) < ORK_TOL) && fs > v->ci->SF*fn ) || fabs(wv[3]) > tl*wv[0] ){
      kinetic = 2;
    } else {
      /* _rkFDSolverModifyWrenchStatic (:643-688): feasibility of the wrench inside the friction pyramids at the polygon corners */
      int fnum = pyr*np; double *ma = (double*)malloc(6*fnum*8), mb[6];
      for(j=0;j<np;j++) for(i=0;i<pyr;i++){ int cc = pyr*j+i; ork_vplane *pl = &v->pl[j];
        ma[0*fnum+cc] = 1.0; ma[1*fnum+cc] = pl->r[1]; ma[2*fnum+cc] = -pl->r[0];
        ma[3*fnum+cc] = v->ci->SF*cs[i]; ma[4*fnum+cc] = v->ci->SF*sn[i];
        ma[5*fnum+cc] = -( ma[2*fnum+cc]*ma[4*fnum+cc] + ma[1*fnum+cc]*ma[3*fnum+cc] ); }
      mb[0] = wv[0]; mb[1] = wv[4]; mb[2] = wv[5]; mb[3] = wv[1]; mb[4] = wv[2]; mb[5] = wv[3];
      if( !vol_lp(6,fnum,ma,mb,NULL,NULL) ) kinetic = 2;
      free(ma);
    }
    if( kinetic == 2 ){
      /* _rkFDSolverModifyWrenchKinetic (:733-843) */
      double *ma = (double*)malloc(3*np*8), mb[3], *mc = (double*)malloc(np*8), *mf = (double*)calloc(np,8), wn[3];
      for(j=0;j<np;j++){ ma[j] = 1.0; ma[np+j] = v->pl[j].r[1]; ma[2*np+j] = -v->pl[j].r[0]; }
      mb[0] = wv[0]; mb[1] = wv[4]; mb[2] = wv[5];
      for(i=0;i<3;i++) wn[i] = fabs(wv[i+1]) < ORK_TOL ? 0.0 : 1.0/wv[i+1];
      for(j=0;j<np;j++){ ork_vplane *pl = &v->pl[j]; double p[3], vel[3], nv;     /* _rkFDSolverPlaneVertSlideDir (:759-776) */
        v3_add(v->center,pl->v,p); link_point_wld_vel(&e->lw[v->link],p,vel);
        v3_cat(vel,-v3_dot(v->norm,vel),v->norm); nv = v3_norm(vel);
        if( fabs(nv) < ORK_TOL ){ pl->s[0] = pl->s[1] = 0; }
        else { double ww = kinetic_friction_weight(w->friction_weight,nv)*v->ci->KF/nv;
          pl->s[0] = -ww*v3_dot(vel,v->axis+3); pl->s[1] = -ww*v3_dot(vel,v->axis+6); }
        mc[j] = -wn[0]*pl->s[0] - wn[1]*pl->s[1] - wn[2]*( pl->r[0]*pl->s[1] - pl->r[1]*pl->s[0] ); }
      if( !vol_lp(3,np,ma,mb,mc,mf) ){
        /* _rkFDSolverModifyWrenchKineticEvalFuncSafety (:795-812): one row, costs accumulated */
        double wn2[2]; for(i=0;i<2;i++) wn2[i] = fabs(wv[i+3]) < ORK_TOL ? 0.0 : 1.0/wv[i+3];
        for(j=0;j<np;j++) mc[j] += wn2[0]*v->pl[j].r[0] - wn2[1]*v->pl[j].r[1];
        for(j=0;j<np;j++) mf[j] = 0;
        vol_lp(1,np,ma,mb,mc,mf);
      }
      wv[1] = wv[2] = wv[3] = 0;                          /* _rkFDSolverModifyWrenchKineticTotalWrench (:814-828) */
      for(j=0;j<np;j++){ double fx = v->pl[j].s[0]*mf[j], fy = v->pl[j].s[1]*mf[j];
        wv[1] += fx; wv[2] += fy; wv[3] += v->pl[j].r[0]*fy - v->pl[j].r[1]*fx; }
      free(ma); free(mc); free(mf);
      setforce = 1;
    }
    if( do_up_ref ) e->v_type[v->pair] = kinetic ? ORK_KF : ORK_SF;
    if( setforce ){ memset(v->wrench,0,sizeof v->wrench);     /* _rkFDSolverModifyWrenchSetForce (:858-867) */
      for(i=0;i<3;i++){ v3_cat(v->wrench,wv[i],v->axis+3*i); v3_cat(v->wrench+3,wv[i+3],v->axis+3*i); } }
  }
  /* _rkFDSolverPushWrench (:919-936) */
  for(k=0;k<P;k++){ ork_vpair *v = &vp[k]; ork_lw *x = &e->lw[v->link]; double pos[3], fl[3], tl[3], nn[3];
    v3_sub(v->center,x->pw,pos); m3_tmulv(x->Rw,pos,pos);
    m3_tmulv(x->Rw,v->wrench,fl); m3_tmulv(x->Rw,v->wrench+3,tl); v3_cross(pos,fl,nn);
    v3_add(x->wext,fl,x->wext); v3_add(x->wext+3,tl,x->wext+3); v3_add(x->wext+3,nn,x->wext+3);
    e->v_np[v->pair] = v->npl; memcpy(e->v_wrench+6*v->pair,v->wrench,6*8); memcpy(e->v_center+3*v->pair,v->center,3*8); }
  free(A); free(b); free(Q); free(c); free(nf); free(dz); free(init); free(f); free(idx); free(vp);
}

/* ------------------------------------------------------------------------------------ */
/* one dynamics evaluation: _rkFDUpdate body / _rkFDUpdateRef (rkfd_sim.c:525-549) */
static void eval_dynamics(ork_env *e, const double *q, const double *qd, double *qdd, int do_up_ref)
{
  const ork_world *w = e->w; int pi, has_elastic = 0, has_rigid = 0, k;
  memset(qdd,0,(w->nq>0?w->nq:1)*8);                  /* zVecZero(acc) */
  eval_kinematics(e,q,qd);                            /* _rkFDConnectJointState (:290-302) */
  /* _rkFDUpdateReset (:445-453): wrench lists cleared in eval_kinematics */
  eval_collision(e);                                  /* _rkFDUpdateCD (:466-471) */
  for(pi=0;pi<w->npair;pi++){ const ork_pair *p = &w->pair[pi]; int col = 0;
    for(k=0;k<w->cell[p->cell].nvert;k++) col |= e->c_active[p->sofs+k];
    if( col ){ if( p->ci.type==ORK_CONTACT_ELASTIC ) has_elastic = 1; else has_rigid = 1; } }
  /* solver->_update (rkfd_vert.c:380-388 / rkfd_mlcp.c:335-343) */
  joint_friction(e,q,qd,do_up_ref);
  if( has_elastic ) solver_penalty(e,do_up_ref);
  e->rn = 0;
  if( has_rigid ){ if( w->solver == ORK_SOLVER_VOLUME ) solver_volume(e,do_up_ref); else solver_rigid(e,do_up_ref); }
  /* _rkFDUpdateAcc (:502-523) */
  aba_backward(e); aba_forward(e,qdd,0);
  if( do_up_ref ){
    int i, r, c;
    update_prev_driving_trq(e);         /* solver->_update_ref (:548) */
    /* [EXT A-17] breakable float: rkChainUpdateABIWrench (the committing evaluation only, rkfd_sim.c:511-514) forms the wrench
     * every joint transmits, IA a + pA at the link origin in link axes; a rigid breakable joint whose force or torque
     * exceeds its threshold is a float joint from the next evaluation on */
    for(i=0;i<w->nl;i++) if( w->link[i].jtype == ORK_JOINT_BRFLOAT && !e->broken[i] ){
      const ork_lw *x = &e->lw[i]; double wr[6];
      for(r=0;r<6;r++){ wr[r] = x->pA[r]; for(c=0;c<6;c++) wr[r] += x->IA[6*r+c]*x->a[c]; }
      if( v3_norm(wr) > w->link[i].brk_f || v3_norm(wr+3) > w->link[i].brk_t ) e->broken[i] = 1;
    }
  }
}
/* a rigid (unbroken) breakable-float joint holds its displacement: the velocity slope of its dofs is zero */
static void hold_unbroken(const ork_env *e, double *vel)
{ int i, j; for(i=0;i<e->w->nl;i++) if( e->w->link[i].jtype == ORK_JOINT_BRFLOAT && !e->broken[i] ) for(j=0;j<6;j++) vel[e->w->link[i].qofs+j] = 0.0; }

void ork_env_eval(ork_env *e, int do_up_ref){ eval_dynamics(e,e->q,e->qd,e->qdd,do_up_ref); }
void ork_env_update_init(ork_env *e){ eval_dynamics(e,e->q,e->qd,e->qdd,1); }

/* rkFDODECatDefault (rkfd_sim.c:306-320) + [EXT] rkChainCatJointDisAll: q <- q (+) k v */
static void cat_dis(const ork_world *w, double *q, double k, const double *v)
{
  int i, j;
  for(i=0;i<w->nl;i++){ const ork_link *l = &w->link[i]; double *qi = q+l->qofs; const double *vi = v+l->qofs;
    switch(l->jtype){
    case ORK_JOINT_SPHER: { double dw[3] = {k*vi[0],k*vi[1],k*vi[2]}; aa_cascade(qi,dw); } break;
    case ORK_JOINT_FLOAT: case ORK_JOINT_BRFLOAT: { double dw[3] = {k*vi[3],k*vi[4],k*vi[5]};
      for(j=0;j<3;j++){ qi[j] += k*vi[j]; }
      aa_cascade(qi+3,dw); } break;
    default: for(j=0;j<l->ndof;j++) qi[j] += k*vi[j]; break; } }
}

/* rkFDUpdate (rkfd_sim.c:560-566): zODE2Update with the assigned explicit Runge-Kutta scheme (default
 * Runge-Kutta-Gill) on the regularised system x = (dis, vel) ([EXT A-9]), t += dt, then the committing
 * reference evaluation.  Stage states are built from the committed state by successive manifold increments in
 * the order k1, k2, ...; zero tableau entries are skipped. */
void ork_env_update(ork_env *e)
{
  const ork_world *w = e->w; int nq = w->nq, i, s, j, ns; double dt = w->dt;
  const double r2 = sqrt(2.0);
  double a[4][4] = {{0}}, b[4] = {0};
  double *xq = e->xs[0], *xv = e->xs[1];
  switch(w->integrator){
  case 1: ns = 4; a[1][0] = 0.5; a[2][1] = 0.5; a[3][2] = 1.0; b[0] = 1.0/6.0; b[1] = 2.0/6.0; b[2] = 2.0/6.0; b[3] = 1.0/6.0; break;
  case 2: ns = 1; b[0] = 1.0; break;
  case 3: ns = 2; a[1][0] = 1.0; b[0] = 0.5; b[1] = 0.5; break;
  default: ns = 4; a[1][0] = 0.5; a[2][0] = (r2-1.0)/2.0; a[2][1] = 1.0-1.0/r2; a[3][1] = -1.0/r2; a[3][2] = 1.0+1.0/r2;
    b[0] = 1.0/6.0; b[1] = (2.0-r2)/6.0; b[2] = (2.0+r2)/6.0; b[3] = 1.0/6.0; break;
  }
  for(s=0;s<ns;s++){
    if( s == 0 ){ memcpy(e->k[0][0],e->qd,nq*8); hold_unbroken(e,e->k[0][0]); eval_dynamics(e,e->q,e->qd,e->k[0][1],0); continue; }
    memcpy(xq,e->q,nq*8); memcpy(xv,e->qd,nq*8);
    for(j=0;j<s;j++) if( a[s][j] != 0.0 ){
      cat_dis(w,xq,a[s][j]*dt,e->k[j][0]);
      for(i=0;i<nq;i++) xv[i] += a[s][j]*dt*e->k[j][1][i]; }
    memcpy(e->k[s][0],xv,nq*8); hold_unbroken(e,e->k[s][0]); eval_dynamics(e,xq,xv,e->k[s][1],0);
  }
  /* combination */
  for(s=0;s<ns;s++) cat_dis(w,e->q,b[s]*dt,e->k[s][0]);
  for(i=0;i<nq;i++) for(s=0;s<ns;s++) e->qd[i] += b[s]*dt*e->k[s][1][i];
  e->t += dt;
  eval_dynamics(e,e->q,e->qd,e->qdd,1);               /* _rkFDUpdateRef */
}

/* ------------------------------------------------------------------------------------ */
void ork_env_get_link_frames(const ork_env *e, double *fr)
{ int i; for(i=0;i<e->w->nl;i++){ memcpy(fr+12*i,e->lw[i].Rw,72); memcpy(fr+12*i+9,e->lw[i].pw,24); } }
void ork_env_get_link_vel(const ork_env *e, double *v)
{ int i; for(i=0;i<e->w->nl;i++) memcpy(v+6*i,e->lw[i].v,48); }
void ork_env_get_link_acc(const ork_env *e, double *a)
{ int i; for(i=0;i<e->w->nl;i++) memcpy(a+6*i,e->lw[i].a,48); }
double ork_env_energy(const ork_env *e)
{
  const ork_world *w = e->w; int i, j; double E = 0; ork_env *m = (ork_env*)e;
  eval_kinematics(m,e->q,e->qd);
  for(i=0;i<w->nl;i++){ const ork_link *l = &w->link[i]; const ork_lw *x = &e->lw[i];
    double vc[3], t[3], Iw[3], cw[3], Ic[9]; int r, c;
    v3_cross(x->v+3,l->com,t); v3_add(x->v,t,vc);
    for(r=0;r<3;r++) for(c=0;c<3;c++) Ic[3*r+c] = l->inertia[3*r+c];
    m3_mulv(Ic,x->v+3,Iw);
    E += 0.5*l->mass*v3_dot(vc,vc) + 0.5*v3_dot(x->v+3,Iw);
    m3_mulv(x->Rw,l->com,cw); E += l->mass*ORK_G*(x->pw[2]+cw[2]);
    if( l->ndof==1 ){ double tin,treg,jm; motor_eval(l,0.0,e->qd[l->qofs],&tin,&treg,&jm); E += 0.5*jm*e->qd[l->qofs]*e->qd[l->qofs]; }
    (void)j; }
  return E;
}
/* Volume solver results of the last evaluation per pair: np (-1: no contact volume), type, wrench (world, at center), center */
void ork_env_get_volume(const ork_env *e, int *np, int *type, double *wrench, double *center)
{
  int n = e->w->npair;
  if(np) memcpy(np,e->v_np,n*sizeof(int)); if(type) memcpy(type,e->v_type,n*sizeof(int));
  if(wrench) memcpy(wrench,e->v_wrench,6*n*8); if(center) memcpy(center,e->v_center,3*n*8);
}
int ork_world_npair(const ork_world *w){ return w->npair; }
/* test hook: Q6 (36, row-major), c6 (6), norm (3) per pair of the last Volume evaluation (rkfd_volume.c:397-491) */
void ork_env_get_volume_constraint(const ork_env *e, double *qc){ memcpy(qc,e->v_qc,45*(e->w->npair>0?e->w->npair:1)*8); }
/* test hook: the Vert QP of the last evaluation (min 1/2 x^T Q x + c^T x, nf x >= 0), its answer x (= f dt), the final
 * active set and how the active-set loop ended: info = {n, m, iterations, term (0 optimal, 1 anti-cycling exit, 2 cap)} */
int ork_env_get_qp(const ork_env *e, double *Q, double *c, double *nf, double *x, int *idx, int *info, int capn, int capm)
{
  int n = e->qp_n, m = e->qp_m;
  if( info ){ info[0] = n; info[1] = m; info[2] = e->qp_iter; info[3] = e->qp_term; }
  if( n > capn || m > capm || !e->qp_Q ) return 0;
  if(Q) memcpy(Q,e->qp_Q,n*n*8);
  if(c) memcpy(c,e->qp_c,n*8);
  if(nf) memcpy(nf,e->qp_nf,m*n*8);
  if(x) memcpy(x,e->qp_x,n*8);
  if(idx) memcpy(idx,e->qp_idx,m*sizeof(int));
  return n;
}

int ork_env_get_rigid_system(const ork_env *e, double *A, double *b, double *f, int cap)
{
  int n = e->rn; if( n > cap ) return -n;
  if(A) memcpy(A,e->rA,n*n*8);
  if(b) memcpy(b,e->rb,n*8);
  if(f) memcpy(f,e->rf,n*8);
  return n;
}

/* ------------------------------------------------------------------------------------ */
/* as ork_batch_run, plus the final contact / pivot state of every env (any pointer may be NULL):
 * act, type [B][nslot], cf [B][nslot][3] contact forces of the last committing evaluation, piv [B][nq] */
int ork_batch_run_state(const ork_world *w, int B, double *q, double *qd, const double *u,
                        int nsteps, int nthreads, double *qdd_out, int *act, int *type, double *cf, int *piv)
{
  int used = 1, nq = w->nq, nl = w->nl, ns = w->nslot;
#ifdef _OPENMP
  if( nthreads > 0 ) omp_set_num_threads(nthreads);
  used = omp_get_max_threads();
#endif
#pragma omp parallel
  {
    ork_env *e = ork_env_new(w); int b, s;
#pragma omp for schedule(dynamic,16)
    for(b=0;b<B;b++){
      int i;
      e->t = 0;
      for(i=0;i<nq;i++){ e->piv_type[i] = ORK_SF; e->piv_prev[i] = 0; e->tf[i] = 0; }
      for(i=0;i<ns;i++) e->c_active[i] = 0;
      for(i=0;i<nl;i++) e->broken[i] = 0;
      ork_env_set_state(e,q+(size_t)nq*b,qd+(size_t)nq*b);
      if( u ) ork_env_set_motor_input(e,u+(size_t)nl*b); else memset(e->min,0,nl*8);
      ork_env_update_init(e);
      for(s=0;s<nsteps;s++) ork_env_update(e);
      ork_env_get_state(e,q+(size_t)nq*b,qd+(size_t)nq*b,qdd_out?qdd_out+(size_t)nq*b:NULL);
      if( act ) memcpy(act+(size_t)ns*b,e->c_active,ns*sizeof(int));
      if( type ) memcpy(type+(size_t)ns*b,e->c_type,ns*sizeof(int));
      if( cf ) memcpy(cf+(size_t)3*ns*b,e->c_f,3*ns*8);
      if( piv ) memcpy(piv+(size_t)nq*b,e->piv_type,nq*sizeof(int));
    }
    ork_env_free(e);
  }
  return used;
}
int ork_batch_run(const ork_world *w, int B, double *q, double *qd, const double *u,
                  int nsteps, int nthreads, double *qdd_out)
{
  int used = 1, nq = w->nq, nl = w->nl;
#ifdef _OPENMP
  if( nthreads > 0 ) omp_set_num_threads(nthreads);
  used = omp_get_max_threads();
#endif
#pragma omp parallel
  {
    ork_env *e = ork_env_new(w); int b, s;
#pragma omp for schedule(static)
    for(b=0;b<B;b++){
      int i;
      e->t = 0;
      for(i=0;i<nq;i++){ e->piv_type[i] = ORK_SF; e->piv_prev[i] = 0; e->tf[i] = 0; }
      for(i=0;i<w->nslot;i++) e->c_active[i] = 0;
      for(i=0;i<nl;i++) e->broken[i] = 0;
      ork_env_set_state(e,q+(size_t)nq*b,qd+(size_t)nq*b);
      if( u ) ork_env_set_motor_input(e,u+(size_t)nl*b); else memset(e->min,0,nl*8);
      ork_env_update_init(e);
      for(s=0;s<nsteps;s++) ork_env_update(e);
      ork_env_get_state(e,q+(size_t)nq*b,qd+(size_t)nq*b,qdd_out?qdd_out+(size_t)nq*b:NULL);
    }
    ork_env_free(e);
  }
  return used;
}
