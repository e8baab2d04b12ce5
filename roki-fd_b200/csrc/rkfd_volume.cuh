/* rkfd_volume.cuh - Volume contact solver (reference src/rkfd_volume.c), member functions of Core (included inside
 * the struct body of rkfd_core.cuh).
 *
 * One environment per lane: every lane whose environment has a rigid pair in contact runs the whole solve on its
 * own scratch column and on thread-local workspaces (the matrices are at most 12 x 12 / 18 x 12):
 *   contact volume   [EXT A-15] cell parallelepiped (8 corners, sign-bit order) clipped by the half-space of ONE
 *                    face of the static box; two streaming passes over the 12 cell triangles - (0) volume and
 *                    barycentre (tetrahedra about a point of the clipping plane, so the cut polygon contributes
 *                    nothing), (1) the surface integrals and the contact-polygon planes; the cut polygon is
 *                    integrated as a fan of its boundary segments.  No triangle list is stored.
 *   A, b             6 cached-ABA probes per pair (unit force / torque at the volume centre), responses read at the
 *                    centres of all pairs of the same tree                         rkfd_volume.c:141-226
 *   Q6, c6, planes   per triangle: midpoint-rule integrals, sign cases             rkfd_volume.c:232-491
 *   QP               Q = sum A_p^T Q6 A_p + L, rows: normal force >= 0 and one centre-of-pressure row per polygon
 *                    edge; dense active-set method (rkfd_opt_qp.c:43-181) with the KKT solve through the Schur
 *                    complement of the Cholesky-factored Q (minimum-norm multipliers by iterated Tikhonov
 *                    regularisation when the active rows are dependent)            rkfd_volume.c:496-548
 *   post-processing  non-pushing wrenches zeroed, centre of pressure projected into the polygon, static-friction
 *                    feasibility LP / kinetic-friction redistribution LP ([EXT A-16] two-phase simplex with Bland's
 *                    rule), wrench on the link                                     rkfd_volume.c:552-936
 * Limits: VOL_P pairs in contact per environment, VOL_PL polygon planes per pair, pyramid * planes <= VOL_LPN
 * (status bit 2 is raised otherwise and the surplus is ignored). */

#ifdef RKFD_VOL_STATS      /* host harness only: work counters */
#define VOL_STAT(i, n) (rkfd_vol_stats[i] += (n))
#else
#define VOL_STAT(i, n)
#endif
/* loops over the unknowns of the QP: fully unrolled (compile-time trip counts), so that the thread-local loads of a dot
 * product or a substitution are issued together instead of one per dependent multiply-add.  Measured on C4 (131,072 envs, ms
 * per step): rolled 68, unrolled 76 while the warps of an SM still went their own ways (58 KB more code: instruction
 * fetch was the limit then); with the block barriers between the phases rolled 58, unrolled 39.  RKFD_VOL_ROLL rolls them. */
#ifdef RKFD_VOL_ROLL
#define RKFD_VOL_U _Pragma("unroll 1")
#else
#define RKFD_VOL_U _Pragma("unroll")
#endif
#ifdef __CUDACC__
#define RKFD_VOL_NI __host__ __device__ __noinline__
#else
#define RKFD_VOL_NI inline
#endif
  static constexpr int VOL_P = 2, VOL_PL = 8, VOL_N = 6*VOL_P, VOL_M = VOL_P*(1+VOL_PL), VOL_MA = VOL_M, VOL_LPN = VOL_PL, VOL_LPS = VOL_LPN + 4;
  struct VolPair {
    int pair, link, fsl, wsl, npl, sofs, fofs;
    V3 center, norm, a1, a2;
    double K, L, SF, KF;
    V3 plv[VOL_PL], pln[VOL_PL]; double r[VOL_PL][2], s[VOL_PL][2];
    double q6[36], c6[6], w[6];
  };
  static RKFD_HD bool vtiny(double x){ return fabs(x) < ZTOL; }

  /* _rkFDSolverSetContactPlane (rkfd_volume.c:350-374) */
  RKFD_VOL_NI void vol_set_plane(VolPair &vp, V3 p, V3 fnorm){
    const V3 t = fnorm - dot(fnorm, vp.norm)*vp.norm;
    if( vtiny(t.x) && vtiny(t.y) && vtiny(t.z) ) return;
    const V3 cn = (-1.0/norm(t))*t;
    for(int i=0;i<vp.npl;i++){
      const V3 dn = cn - vp.pln[i];
      if( fabs(dn.x) < 1e-8 && fabs(dn.y) < 1e-8 && fabs(dn.z) < 1e-8 ){
        const V3 dv = vp.plv[i] - p;
        if( fabs(dot(cn, dv)) < 1e-8 ){
          if( dot(vp.norm, cross(cn, dv)) > 0.0 ) vp.plv[i] = p;
          return;
        }
      }
    }
    if( vp.npl >= VOL_PL ){ bad |= 4; return; }
    vp.plv[vp.npl] = p; vp.pln[vp.npl] = cn; vp.npl++;
  }
  /* _rkFDSolverConstraintMidDepth + _rkFDSolverConstraintDepth (rkfd_volume.c:232-241, 296-310) */
  static RKFD_VOL_NI void vol_cc(const V3 (&p)[3], const double (&h)[3], double K, V3 nrm, double (&cc)[6]){
    const double s = 0.5*norm(cross(p[1]-p[0], p[2]-p[0])), k = K*s/6.0;
    const double hm0 = k*(h[0]+h[1]), hm1 = k*(h[1]+h[2]), hm2 = k*(h[0]+h[2]), hc = k*(h[0]+h[1]+h[2])*2;
    const V3 m0 = 0.5*(p[0]+p[1]), m1 = 0.5*(p[1]+p[2]), m2 = 0.5*(p[2]+p[0]);
    const V3 a = cross(nrm, hm0*m0) + cross(nrm, hm1*m1) + cross(nrm, hm2*m2);
    cc[0] = -hc*nrm.x; cc[1] = -hc*nrm.y; cc[2] = -hc*nrm.z; cc[3] = a.x; cc[4] = a.y; cc[5] = a.z;
  }
  /* _rkFDSolverConstraintInnerPoint (rkfd_volume.c:331-348) */
  static RKFD_VOL_NI V3 vol_inner(V3 p1, V3 p2, double h1, double h2){
    if( vtiny(h1) ) return p1;
    if( vtiny(h2) ) return p2;
    if( vtiny(h2-h1) ) return 0.5*(p1+p2);
    return (h2/(h2-h1))*p1 + (h1/(h1-h2))*p2;
  }
  /* one triangle of the contact volume (outward face normal fn): body of the loop of _rkFDSolverConstraint
   * (rkfd_volume.c:397-491) */
  RKFD_VOL_NI void vol_tri(VolPair &vp, V3 ta, V3 tb, V3 tc, V3 fn){
    if( dot(cross(tb-ta, tc-ta), fn) < 0 ){ const V3 t = tb; tb = tc; tc = t; }
    V3 pf[3] = { ta - vp.center, tb - vp.center, tc - vp.center }, p[3], pp0;
    double h[3], cc[6];
    for(int j=0;j<3;j++){ h[j] = dot(vp.norm, pf[j]); p[j] = pf[j] - h[j]*vp.norm; }
    { /* _rkFDSolverConstraintAddQ (:279-294) */
      const double s = 0.5*norm(cross(p[1]-p[0], p[2]-p[0]));
      const V3 pc = (s/3.0)*(p[0]+p[1]+p[2]);
      double *q = vp.q6;
      q[0] += s; q[7] += s; q[14] += s;
      /* [pc x] into the lower-left block, its negative into the upper-right one */
      q[6*3+1] += -pc.z; q[6*3+2] +=  pc.y; q[6*4+0] +=  pc.z; q[6*4+2] += -pc.x; q[6*5+0] += -pc.y; q[6*5+1] +=  pc.x;
      q[6*0+4] -= -pc.z; q[6*0+5] -=  pc.y; q[6*1+3] -=  pc.z; q[6*1+5] -= -pc.x; q[6*2+3] -= -pc.y; q[6*2+4] -=  pc.x;
      const V3 pm[3] = { 0.5*(p[0]+p[1]), 0.5*(p[1]+p[2]), 0.5*(p[2]+p[0]) };
      for(int j=0;j<3;j++){   /* [m x][m x] = m m^T - |m|^2 I */
        const V3 m_ = pm[j]; const double mm_ = dot(m_, m_), k = s/3.0;
        q[6*3+3] -= k*(m_.x*m_.x - mm_); q[6*3+4] -= k*m_.x*m_.y; q[6*3+5] -= k*m_.x*m_.z;
        q[6*4+3] -= k*m_.y*m_.x; q[6*4+4] -= k*(m_.y*m_.y - mm_); q[6*4+5] -= k*m_.y*m_.z;
        q[6*5+3] -= k*m_.z*m_.x; q[6*5+4] -= k*m_.z*m_.y; q[6*5+5] -= k*(m_.z*m_.z - mm_);
      }
    }
    vol_cc(p, h, vp.K, vp.norm, cc);
    int st = 0, stp[3] = {0,0,0};
    for(int j=0;j<3;j++){
      if( h[j] > ZTOL ){ st += 1<<(j*2); stp[1] = j; }
      else if( h[j] < -ZTOL ){ st += 1<<(j*2+1); stp[2] = j; }
      else stp[0] = j;
    }
    double sgn2 = -2.0;
    switch( st ){
    case 0x01: case 0x04: case 0x10: case 0x05: case 0x11: case 0x14:
      vol_set_plane(vp, pf[stp[0]], fn);
    case 0x15:
      for(int j=0;j<6;j++) vp.c6[j] += cc[j];
      return;
    case 0x02: case 0x08: case 0x20: case 0x0a: case 0x22: case 0x28:
      vol_set_plane(vp, pf[stp[0]], fn);
    case 0x2a:
      for(int j=0;j<6;j++) vp.c6[j] -= cc[j];
      return;
    case 0x24: case 0x12: case 0x09:
      for(int j=0;j<6;j++) vp.c6[j] += cc[j];
      p[stp[1]] = vol_inner(pf[stp[1]], pf[stp[2]], h[stp[1]], h[stp[2]]);
      h[stp[1]] = 0.0;
      pp0 = p[stp[0]];
      break;
    case 0x06: case 0x21: case 0x18:
      for(int j=0;j<6;j++) vp.c6[j] += cc[j];
      p[stp[2]] = vol_inner(pf[stp[1]], pf[stp[2]], h[stp[1]], h[stp[2]]);
      h[stp[2]] = 0.0;
      pp0 = p[stp[2]];
      break;
    case 0x16: case 0x19: case 0x25:
      stp[0] = (stp[2]+1)%3; stp[1] = (stp[0]+1)%3;
      for(int j=0;j<6;j++) vp.c6[j] += cc[j];
      p[stp[0]] = vol_inner(pf[stp[2]], pf[stp[0]], h[stp[2]], h[stp[0]]);
      p[stp[1]] = vol_inner(pf[stp[2]], pf[stp[1]], h[stp[2]], h[stp[1]]);
      h[stp[0]] = h[stp[1]] = 0.0;
      pp0 = p[stp[0]];
      break;
    case 0x1a: case 0x26: case 0x29:
      stp[0] = (stp[1]+1)%3; stp[2] = (stp[0]+1)%3;
      for(int j=0;j<6;j++) vp.c6[j] -= cc[j];
      p[stp[0]] = vol_inner(pf[stp[1]], pf[stp[0]], h[stp[1]], h[stp[0]]);
      p[stp[2]] = vol_inner(pf[stp[1]], pf[stp[2]], h[stp[1]], h[stp[2]]);
      h[stp[0]] = h[stp[2]] = 0.0;
      pp0 = p[stp[2]];
      sgn2 = 2.0;
      break;
    default:
      return;
    }
    vol_cc(p, h, vp.K, vp.norm, cc);
    for(int j=0;j<6;j++) vp.c6[j] += sgn2*cc[j];
    vol_set_plane(vp, pp0, fn);
  }

  /* the 12 triangles of the cell clipped by {x : nn.(x - p0) <= 0}.  pass 0: volume and barycentre; pass 1:
   * integration (vol_tri) incl. the cut polygon.  Returns false when the volume is empty. */
  RKFD_VOL_NI bool vol_clip(VolPair &vp, const V3 (&vw)[8], const double (&d)[8], V3 cen, V3 p0, int pass){
    const int quads[6][4] = { {1,3,7,5}, {0,4,6,2}, {2,6,7,3}, {0,1,5,4}, {4,5,7,6}, {0,2,3,1} };
    double vol = 0; V3 bc = v3(0,0,0), o = v3(0,0,0); bool have_o = false;
    for(int f=0;f<6;f++){
      const int *qd = quads[f];
      V3 fn = cross(vw[qd[1]]-vw[qd[0]], vw[qd[3]]-vw[qd[0]]);
      const double nn = norm(fn); if( nn == 0 ) continue;
      fn = (1.0/nn)*fn;
      if( dot(fn, vw[qd[0]]-cen) < 0 ) fn = -fn;
      for(int tr=0;tr<2;tr++){
        const int id[3] = { qd[0], qd[1+tr], qd[2+tr] };
        V3 poly[4]; bool onp[4]; int np = 0;
        for(int i=0;i<3;i++){
          const int a = id[i], b = id[(i+1)%3];
          if( d[a] <= 0 ){ poly[np] = vw[a]; onp[np] = d[a] == 0; np++; }
          if( (d[a] < 0 && d[b] > 0) || (d[a] > 0 && d[b] < 0) ){
            const double tt = d[a]/(d[a]-d[b]);
            poly[np] = vw[a] + tt*(vw[b]-vw[a]); onp[np] = true; np++;
          }
        }
        if( np < 3 ) continue;
        /* pass 1: the boundary segment of the cut polygon carried by this clipped triangle closes the loop below as
         * one more triangle (o, segment) of the cut polygon, o = the first cut point found */
        int n_on = 0, i0 = 0, i1 = 0;
        for(int i=0;i<np;i++) if( onp[i] ){ if( n_on == 0 ) i0 = i; else i1 = i; n_on++; }
        const bool seg = pass == 1 && n_on == 2;
        if( seg && !have_o ){ o = poly[i0]; have_o = true; }
        const V3 sa = poly[i0], sb = poly[i1];
        for(int i=1;i+1<np+(seg?1:0);i++){
          V3 a = i+1 < np ? poly[0] : o, b = i+1 < np ? poly[i] : sa, c2 = i+1 < np ? poly[i+1] : sb;
          if( pass == 0 ){
            if( dot(cross(b-a, c2-a), fn) < 0 ){ const V3 t = b; b = c2; c2 = t; }
            const V3 ra = a-p0, rb = b-p0, rc = c2-p0;
            const double v6 = dot(ra, cross(rb, rc))/6.0;
            vol += v6; bc = bc + (0.25*v6)*(ra+rb+rc);
          } else vol_tri(vp, a, b, c2, i+1 < np ? fn : vp.norm);
        }
      }
    }
    if( pass == 0 ){
      if( !(vol > 1.0e-18) ) return false;
      vp.center = p0 + (1.0/vol)*bc;
    }
    return true;
  }

  /* acceleration response (frame of link Lt) of link Lt to the bias change (dpf, dpn) on link Lc of the same tree
   * ([EXT A-5]; probe_link with a separate target) */
  RKFD_VOL_NI void vol_probe(const ModelDev &m, int Lc, V3 dpf, V3 dpn, int Lt, V3 &ral, V3 &raa){
    double du[6*MAX_LINKS]; int pth[MAX_LINKS]; int np = 0;
    for(int i=Lc;;){
      const LinkDev &L = m.link[i]; const int sl = Spec::slot(i,L), jt = eff_jt(Spec::jtype(i,L), Spec::slot(i,L));
      pth[np] = i;
      V3 paf = dpf, pan = dpn;
      for(int k=0;k<6;k++) du[6*np+k] = 0.0;
      switch(jt){
      case J_REVOL: case J_PRISM: {
        const double d = jt == J_REVOL ? -dpn.z : -dpf.z;
        du[6*np] = d;
        const double k = Q(Spec::sc(i,L)+2)*d;
        paf = dpf + k*ld3(sl); pan = dpn + k*ld3(sl+3);
      } break;
      case J_SPHER: {
        const M3 E = tmm(ldm(sl+27), org_R(L));
        const V3 d = -tmul(E, dpn);
        du[6*np] = d.x; du[6*np+1] = d.y; du[6*np+2] = d.z;
        const V3 k = mul(lds(sl+18), d);
        paf = dpf + mul(ldm(sl), k); pan = dpn + mul(ldm(sl+9), k);
      } break;
      case J_FLOAT: du[6*np]=dpf.x; du[6*np+1]=dpf.y; du[6*np+2]=dpf.z; du[6*np+3]=dpn.x; du[6*np+4]=dpn.y; du[6*np+5]=dpn.z; break;
      case J_CYLIN: case J_HOOKE: { double d2[2]; probe2_in(jt, sl, dpf, dpn, d2, paf, pan); du[6*np] = d2[0]; du[6*np+1] = d2[1]; } break;
      default: break;
      }
      np++;
      if( Spec::parent(i,L) < 0 || jt == J_FLOAT ) break;
      V3 vJ, wJ; const XF x = joint_xform<TagRT,false>(m, L, i, vJ, wJ);
      dpf = xf_mul(x, paf); dpn = xf_mul(x, pan) + cross(x.p, dpf);
      i = Spec::parent(i,L);
    }
    int pt[MAX_LINKS]; int nt = 0;
    for(int i=Lt;;){ const LinkDev &L = m.link[i]; pt[nt++] = i; if( Spec::parent(i,L) < 0 || eff_jt(Spec::jtype(i,L), Spec::slot(i,L)) == J_FLOAT ) break; i = Spec::parent(i,L); }
    V3 al = v3(0,0,0), aa = v3(0,0,0);
    for(int q=nt-1;q>=0;q--){
      const int i = pt[q]; const LinkDev &L = m.link[i]; const int sl = Spec::slot(i,L), jt = eff_jt(Spec::jtype(i,L), Spec::slot(i,L));
      double d[6] = {0,0,0,0,0,0};
      for(int k=0;k<np;k++) if( pth[k] == i ) for(int r=0;r<6;r++) d[r] = du[6*k+r];
      V3 vJ, wJ; const XF x = joint_xform<TagRT,false>(m, L, i, vJ, wJ);
      V3 xl = xf_tmul(x, al + cross(aa, x.p)), xa = xf_tmul(x, aa);
      switch(jt){
      case J_REVOL: case J_PRISM: {
        const double acc = Q(Spec::sc(i,L)+2)*( d[0] - (dot(ld3(sl),xl) + dot(ld3(sl+3),xa)) );
        if( jt == J_REVOL ) xa.z += acc; else xl.z += acc;
      } break;
      case J_SPHER: {
        const V3 rhs = v3(d[0],d[1],d[2]) - (tmul(ldm(sl), xl) + tmul(ldm(sl+9), xa));
        xa = xa + mul(tmm(ldm(sl+27), org_R(L)), mul(lds(sl+18), rhs));
      } break;
      case J_FLOAT: {   /* da = -IA^-1 dp */
        double r[6] = {0,0,0,0,0,0}; int k = 0;
        for(int a=0;a<6;a++) for(int b=a;b<6;b++){ const double iv = c.S(sl+18+k); r[a] -= iv*d[b]; if( b != a ) r[b] -= iv*d[a]; k++; }
        xl = v3(r[0],r[1],r[2]); xa = v3(r[3],r[4],r[5]);
      } break;
      case J_CYLIN: case J_HOOKE: { const double d2[2] = { d[0], d[1] }; probe2_out(jt, sl, d2, xl, xa); } break;
      default: break;
      }
      al = xl; aa = xa;
    }
    ral = al; raa = aa;
  }

  /* rkFDQPSolveASM (rkfd_opt_qp.c:43-181) on dense data: min 1/2 x^T Q x + c^T x  s.t.  A x >= 0 (m rows, stride
   * VOL_N).  Q is positive definite (relaxation L > 0 on its diagonal): Q = G G^T once, z = G^-1 c and the rows
   * B_i = G^-1 a_i of all constraints once (forward substitutions); an iteration then needs, with the active rows W,
   *   S = B_W B_W^T,  S lambda = B_W z,  x* = G^-T ( B_W^T lambda - z )
   * which is the KKT system [[-Q, A_W^T],[A_W, 0]] [x; lambda] = [c; 0] of the reference (rkfd_opt_qp.c:82-106).  The
   * reference solves it with zLESolveMP: when active rows are dependent (an unloaded sole has ALL its rows active) the
   * multipliers are the minimum-norm ones.  Here: lambda = S^+ (B_W z) by iterated Tikhonov regularisation -
   * lambda += (S + eps I)^-1 (B_W z - S lambda), one Cholesky factorisation and refinements to rounding - which converges to the
   * minimum-norm solution of the (always consistent) system and is the plain solve when S is regular: one code path, no
   * eigen-decomposition.  x holds the initial point on entry.  idx: active flags. */
  RKFD_VOL_NI void vol_asm(int mrows, const double *Qm, const double *cv, const double *A, double *x, unsigned &idx_out){
    /* The problem always has VOL_N unknowns (a single pair is padded with an identity block by the caller): every loop
     * over the unknowns has a compile-time trip count (RKFD_VOL_U above decides whether it is unrolled). */
    constexpr int N = VOL_N;
    const int QP_HIST = 32, QP_MAXIT = 256;
    double G[N*N], z[N], Bm[VOL_M*N];
RKFD_VOL_U
    for(int i=0;i<N;i++){
RKFD_VOL_U
      for(int j=0;j<=i;j++){
        double s = Qm[N*i+j];
RKFD_VOL_U
        for(int k=0;k<j;k++) s -= G[N*i+k]*G[N*j+k];
        if( i == j ){ if( !(s > 0) ){ bad |= 2; s = 1.0; } G[N*i+i] = sqrt(s); } else G[N*i+j] = s/G[N*j+j];
      }
    }
    double gi[N];                        /* reciprocals of the diagonal */
RKFD_VOL_U
    for(int i=0;i<N;i++) gi[i] = 1.0/G[N*i+i];
RKFD_VOL_U
    for(int i=0;i<N;i++){ double s = cv[i];
RKFD_VOL_U
      for(int k=0;k<i;k++) s -= G[N*i+k]*z[k];
      z[i] = s*gi[i]; }
#pragma unroll 1
    for(int r=0;r<mrows;r++){
      double br[N];
RKFD_VOL_U
      for(int i=0;i<N;i++){ double s = A[N*r+i];
RKFD_VOL_U
        for(int k=0;k<i;k++) s -= G[N*i+k]*br[k];
        br[i] = s*gi[i]; }
RKFD_VOL_U
      for(int i=0;i<N;i++) Bm[N*r+i] = br[i];
    }
    unsigned idx = 0;
#pragma unroll 1
    for(int i=0;i<mrows;i++){ double s = 0;
RKFD_VOL_U
      for(int j=0;j<N;j++) s += A[N*i+j]*x[j];
      if( fabs(s) < ZTOL ) idx |= 1u << i; }
    unsigned hist_idx[QP_HIST]; double hist_obj[QP_HIST]; int nhist = 0;
#pragma unroll 1
    for(int iter=0; iter<QP_MAXIT; iter++){
      int act[VOL_M]; int ma = 0; VOL_STAT(0, 1);
      for(int i=0;i<mrows;i++) if( idx >> i & 1u ) act[ma++] = i;
      double xs[N], lam[VOL_MA], wv[N];
RKFD_VOL_U
      for(int i=0;i<N;i++) wv[i] = -z[i];
      if( ma > 0 ){
        double S[VOL_MA*VOL_MA], C[VOL_MA*VOL_MA], rhs[VOL_MA], res[VOL_MA];
        double smax = 0;
#pragma unroll 1
        for(int a=0;a<ma;a++){
          double ba[N];
RKFD_VOL_U
          for(int j=0;j<N;j++) ba[j] = Bm[N*act[a]+j];
#pragma unroll 1
          for(int b=0;b<=a;b++){ double s = 0;
RKFD_VOL_U
            for(int j=0;j<N;j++) s += ba[j]*Bm[N*act[b]+j];
            S[VOL_MA*a+b] = s; S[VOL_MA*b+a] = s; }
          double s = 0;
RKFD_VOL_U
          for(int j=0;j<N;j++) s += ba[j]*z[j];
          rhs[a] = s; lam[a] = 0.0;
          if( S[VOL_MA*a+a] > smax ) smax = S[VOL_MA*a+a];
        }
        const double eps = 1.0e-9*smax + 1.0e-300;
        for(int i=0;i<ma;i++) for(int j=0;j<=i;j++){
          double s = S[VOL_MA*i+j] + ( i == j ? eps : 0.0 ); for(int k=0;k<j;k++) s -= C[VOL_MA*i+k]*C[VOL_MA*j+k];
          if( i == j ) C[VOL_MA*i+i] = sqrt(s > 0 ? s : eps); else C[VOL_MA*i+j] = s/C[VOL_MA*j+j];
        }
        /* at most six refinements, fewer when the update is below rounding (2 when S is regular and well conditioned; a
         * singular value s of S converges with the factor eps / (s + eps) per step) */
        for(int ref=0; ref<6; ref++){
          for(int i=0;i<ma;i++){ double s = rhs[i]; for(int k=0;k<ma;k++) s -= S[VOL_MA*i+k]*lam[k]; res[i] = s; }
          for(int i=0;i<ma;i++){ double s = res[i]; for(int k=0;k<i;k++) s -= C[VOL_MA*i+k]*res[k]; res[i] = s/C[VOL_MA*i+i]; }
          for(int i=ma-1;i>=0;i--){ double s = res[i]; for(int k=i+1;k<ma;k++) s -= C[VOL_MA*k+i]*res[k]; res[i] = s/C[VOL_MA*i+i]; }
          double dmax = 0, lmax = 0;
          for(int i=0;i<ma;i++){ lam[i] += res[i]; if( fabs(res[i]) > dmax ) dmax = fabs(res[i]); if( fabs(lam[i]) > lmax ) lmax = fabs(lam[i]); }
          if( dmax <= 1.0e-14*lmax ) break;
        }
#pragma unroll 1
        for(int k=0;k<ma;k++){ const double lk = lam[k];
RKFD_VOL_U
          for(int i=0;i<N;i++) wv[i] += lk*Bm[N*act[k]+i]; }
      }
RKFD_VOL_U
      for(int i=N-1;i>=0;i--){ double s = wv[i];
RKFD_VOL_U
        for(int k=i+1;k<N;k++) s -= G[N*k+i]*xs[k];
        xs[i] = s*gi[i]; }
      bool stepped = false;
RKFD_VOL_U
      for(int i=0;i<N;i++) stepped = stepped || !(fabs(xs[i]-x[i]) < ZTOL);
      if( !stepped ){
RKFD_VOL_U
        for(int i=0;i<N;i++) x[i] = xs[i];
        bool neg = false; for(int k=0;k<ma;k++) if( lam[k] < 0 ){ neg = true; break; }
        if( !neg ) break;
        double lmin = lam[0]; for(int k=1;k<ma;k++) if( lam[k] < lmin ) lmin = lam[k];
        for(int k=0;k<ma;k++) if( fabs(lam[k]-lmin) < 1.0e-8 ) idx &= ~(1u << act[k]);
        continue;
      }
      double dx[N], alpha = 1.0;
RKFD_VOL_U
      for(int j=0;j<N;j++) dx[j] = xs[j]-x[j];
#pragma unroll 1
      for(int i=0;i<mrows;i++){
        if( idx >> i & 1u ) continue;
        double ad = 0, ax = 0;
RKFD_VOL_U
        for(int j=0;j<N;j++){ const double aij = A[N*i+j]; ad += aij*dx[j]; ax += aij*x[j]; }
        if( ad < 0 ){ const double t = (0.0 - ax)/ad; if( t < alpha ) alpha = t; }
      }
RKFD_VOL_U
      for(int i=0;i<N;i++) x[i] += alpha*dx[i];
#pragma unroll 1
      for(int i=0;i<mrows;i++){
        if( idx >> i & 1u ) continue;
        double ax = 0;
RKFD_VOL_U
        for(int j=0;j<N;j++) ax += A[N*i+j]*x[j];
        if( fabs(ax) < ZTOL ) idx |= 1u << i;
      }
      double objv = 0;
RKFD_VOL_U
      for(int i=0;i<N;i++){ double s = 0;
RKFD_VOL_U
        for(int j=0;j<N;j++) s += Qm[N*i+j]*x[j];
        objv += 0.5*x[i]*s + cv[i]*x[i]; }
      bool endflag = false;
      for(int h=0;h<nhist && !endflag;h++) if( hist_idx[h] == idx && !(fabs(hist_obj[h]/objv - 1.0) > 1.0e-8) ) endflag = true;
      if( endflag ) break;
      if( nhist < QP_HIST ){ hist_idx[nhist] = idx; hist_obj[nhist] = objv; nhist++; }
      if( iter == QP_MAXIT-1 ) bad |= 2;
    }
    wk = (unsigned)nhist;
    idx_out = idx;
  }

  /* [EXT A-16] min cost^T x  s.t.  A x = b, x >= 0: two-phase tableau simplex, Bland's rule (cost == nullptr:
   * feasibility only).  A: mr x nc (mr <= 3: the kinetic-friction LPs) with stride VOL_LPN.  Same pivoting rules and tolerances as the oracle's vol_lp. */
  static RKFD_VOL_NI bool vol_lp(int mr, int nc, const double *A, const double *b, const double *cost_in, double *x){
    const int nt = nc + mr; double T[3*VOL_LPS]; int basis[3]; double bmax = 0;      /* mr <= 3, nc <= VOL_LPN */
    for(int i=0;i<mr;i++){ const double sg = b[i] < 0 ? -1.0 : 1.0;
      for(int j=0;j<nc;j++) T[VOL_LPS*i+j] = sg*A[VOL_LPN*i+j];
      for(int j=0;j<mr;j++) T[VOL_LPS*i+nc+j] = i == j ? 1.0 : 0.0;
      T[VOL_LPS*i+nt] = sg*b[i]; basis[i] = nc+i; if( fabs(b[i]) > bmax ) bmax = fabs(b[i]); }
    const double eps = 1.0e-10*(1.0+bmax);
    bool ok = true;
    for(int phase=1;phase<=2 && ok;phase++){
      const int ncol = phase == 1 ? nt : nc;
      if( phase == 2 && !cost_in ) break;
      for(int it=0;it<20000;it++){
        int enter = -1, leave = -1; double best = 0; VOL_STAT(5, 1);
        for(int j=0;j<ncol && enter<0;j++){
          double rc = phase == 1 ? ( j >= nc ? 1.0 : 0.0 ) : cost_in[j]; bool bas = false;
          for(int i=0;i<mr;i++){ if( basis[i] == j ) bas = true;
            const double cb = phase == 1 ? ( basis[i] >= nc ? 1.0 : 0.0 ) : ( basis[i] < nc ? cost_in[basis[i]] : 0.0 );
            rc -= cb*T[VOL_LPS*i+j]; }
          if( !bas && rc < -1.0e-11 ) enter = j;
        }
        if( enter < 0 ) break;
        for(int i=0;i<mr;i++){ const double a = T[VOL_LPS*i+enter];
          if( a > 1.0e-11 ){ const double ratio = T[VOL_LPS*i+nt]/a;
            if( leave < 0 || ratio < best - 1.0e-13 || ( fabs(ratio-best) <= 1.0e-13 && basis[i] < basis[leave] ) ){ leave = i; best = ratio; } } }
        if( leave < 0 ){ ok = false; break; }
        const double pv = T[VOL_LPS*leave+enter];
        for(int j=0;j<=nt;j++) T[VOL_LPS*leave+j] /= pv;
        for(int r=0;r<mr;r++) if( r != leave ){ const double fct = T[VOL_LPS*r+enter]; if( fct != 0 ) for(int j=0;j<=nt;j++) T[VOL_LPS*r+j] -= fct*T[VOL_LPS*leave+j]; }
        basis[leave] = enter;
      }
      if( phase == 1 ){
        double art = 0;
        for(int i=0;i<mr;i++) if( basis[i] >= nc ) art += T[VOL_LPS*i+nt];
        if( art > eps ) ok = false;
        else for(int i=0;i<mr;i++) if( basis[i] >= nc ){
          int j = 0; for(;j<nc;j++) if( fabs(T[VOL_LPS*i+j]) > 1.0e-9 ) break;
          if( j < nc ){ const double pv = T[VOL_LPS*i+j];
            for(int jj=0;jj<=nt;jj++) T[VOL_LPS*i+jj] /= pv;
            for(int r=0;r<mr;r++) if( r != i ){ const double fct = T[VOL_LPS*r+j]; if( fct != 0 ) for(int jj=0;jj<=nt;jj++) T[VOL_LPS*r+jj] -= fct*T[VOL_LPS*i+jj]; }
            basis[i] = j; }
        }
      }
    }
    if( ok && x ){ for(int j=0;j<nc;j++) x[j] = 0; for(int i=0;i<mr;i++) if( basis[i] < nc ) x[basis[i]] = T[VOL_LPS*i+nt]; }
    return ok;
  }

  /* static friction (rkfd_volume.c:643-688): is b = (fn, t1, t2, f1, f2, tn) a non-negative combination of the columns
   * g_j = (1, r2, -r1, SF cos_i, SF sin_i, r1 SF sin_i - r2 SF cos_i), j = pyramid * corner + i, i.e. can pyramid forces
   * at the polygon corners carry the wrench?  Phase 1 of the simplex in revised form: the 6 x 6 basis inverse instead of the 6 x (columns + 7) tableau, columns
   * generated on the fly - the whole state stays in registers. */
  static RKFD_HD void vol_static_col(const ModelDev &m, const VolPair &v, int pyr, int nc, const double (&sg)[6], int j, double (&col)[6]){
    if( j < nc ){
      const int k = j/pyr, i = j - pyr*k; const double r1 = v.r[k][0], r2 = v.r[k][1], fc = v.SF*m.sc_cos[i], fs = v.SF*m.sc_sin[i];
      col[0] = sg[0]; col[1] = sg[1]*r2; col[2] = -sg[2]*r1; col[3] = sg[3]*fc; col[4] = sg[4]*fs; col[5] = -sg[5]*( (-r1)*fs + r2*fc );
    } else {
#pragma unroll
      for(int i=0;i<6;i++) col[i] = i == j-nc ? 1.0 : 0.0;
    }
  }
  RKFD_VOL_NI bool vol_static_feasible(const ModelDev &m, const VolPair &v, int np, const double (&b)[6]){
    const int pyr = m.pyramid, nc = pyr*np, nt = nc + 6;
    double Bi[36], xb[6], sg[6]; int basis[6]; double bmax = 0;
#pragma unroll
    for(int i=0;i<6;i++){ sg[i] = b[i] < 0 ? -1.0 : 1.0; xb[i] = sg[i]*b[i]; basis[i] = nc+i; if( fabs(b[i]) > bmax ) bmax = fabs(b[i]);
#pragma unroll
      for(int j=0;j<6;j++) Bi[6*i+j] = i == j ? 1.0 : 0.0; }
    const double eps = 1.0e-10*(1.0+bmax);
    VOL_STAT(3, 1);
    for(int it=0; it<20000; it++){
      double y[6], col[6]; VOL_STAT(4, 1);
#pragma unroll
      for(int j=0;j<6;j++){ double s = 0;
#pragma unroll
        for(int i=0;i<6;i++) s += basis[i] >= nc ? Bi[6*i+j] : 0.0; y[j] = s; }
      /* entering column: the most negative reduced cost (Dantzig; every lane scans all columns: no divergence, a third of
       * Bland's pivots); Bland's first negative column after 48 pivots (anti-cycling: the start is degenerate whenever
       * a friction component of b is zero).  The verdict does not depend on the rule. */
      int enter = -1; double rcmin = -1.0e-11;
      for(int j=0;j<nt;j++){
        bool bas = false;
#pragma unroll
        for(int i=0;i<6;i++) bas = bas || basis[i] == j;
        if( bas ) continue;
        double cj[6];
        vol_static_col(m, v, pyr, nc, sg, j, cj);
        double rc = j >= nc ? 1.0 : 0.0;
#pragma unroll
        for(int i=0;i<6;i++) rc -= y[i]*cj[i];
        if( rc < rcmin ){ rcmin = rc; enter = j; if( it >= 48 ) break; }
      }
      if( enter >= 0 ) vol_static_col(m, v, pyr, nc, sg, enter, col);
      if( enter < 0 ) break;
      double d[6]; int leave = -1, lbas = 0; double best = 0;
#pragma unroll
      for(int i=0;i<6;i++){ double s = 0;
#pragma unroll
        for(int j=0;j<6;j++) s += Bi[6*i+j]*col[j]; d[i] = s; }
#pragma unroll
      for(int i=0;i<6;i++) if( d[i] > 1.0e-11 ){ const double ratio = xb[i]/d[i];
        if( leave < 0 || ratio < best - 1.0e-13 || ( fabs(ratio-best) <= 1.0e-13 && basis[i] < lbas ) ){ leave = i; lbas = basis[i]; best = ratio; } }
      if( leave < 0 ) return false;
      double prow[6], pv = 0, px = 0;
#pragma unroll
      for(int i=0;i<6;i++) if( i == leave ){ pv = d[i]; px = xb[i];
#pragma unroll
        for(int j=0;j<6;j++) prow[j] = Bi[6*i+j]; }
      px /= pv;
#pragma unroll
      for(int j=0;j<6;j++) prow[j] /= pv;
#pragma unroll
      for(int i=0;i<6;i++){
        if( i == leave ){ xb[i] = px; basis[i] = enter;
#pragma unroll
          for(int j=0;j<6;j++) Bi[6*i+j] = prow[j]; }
        else { xb[i] -= d[i]*px;
#pragma unroll
          for(int j=0;j<6;j++) Bi[6*i+j] -= d[i]*prow[j]; }
      }
    }
    double art = 0;
#pragma unroll
    for(int i=0;i<6;i++) if( basis[i] >= nc ) art += xb[i];
    return !(art > eps);
  }

  /* rkFDKineticFrictionWeight (rkfd_util.c:193-196) times KF / |v|, tangential slip velocity of the point in `vel` */
  RKFD_VOL_NI double vol_slip(const ModelDev &m, const VolPair &v, V3 p, V3 &vel){
    vel = vol_point_vel(v, p); vel = vel - dot(v.norm, vel)*v.norm;
    const double nv = norm(vel);
    return vtiny(nv) ? 0.0 : (1.0 - exp(-1.0*m.friction_weight*nv))*v.KF/nv;
  }
  RKFD_HD V3 vol_point_vel(const VolPair &vp, V3 p){     /* rkFDLinkPointWldVel (rkfd_util.c:14-24), static partner */
    const M3 Rw = ldm(vp.fsl); const V3 pw = ld3(vp.fsl+9), vl = ld3(vp.fsl+12), om = ld3(vp.fsl+15);
    return mul(Rw, vl) + cross(mul(Rw, om), p - pw);
  }

  /* _rkFDSolverVolume (rkfd_volume.c:939-957) for this lane's environment */
  RKFD_VOL_NI void rigid_volume(const ModelDev &m, bool ref){
    /* Every thread of the block walks through the phases below (lanes and warps without a contact volume with P = 0, i.e.
     * empty loops), with a block barrier between phases: the warps of an SM then execute the same few KB of code at any
     * time instead of eight different parts of the 350 KB solver (instruction-fetch stalls dominated the profile). */
    VolPair vp[VOL_P]; int P = 0;
    /* ---- contact volumes (rkFDSolverColChk_Volume, [EXT A-15]).  First the pairs of this environment that touch (16 bits
     * each: pair, first vertex inside), then their volumes candidate by candidate with the lanes of the warp in lockstep: a
     * lane whose left sole touches and a lane whose right sole touches clip their cells together, not one after the other. */
    constexpr int VOL_CAND = 4;
    unsigned long long cand = 0; int ncand = 0;
    for(int pi=0;pi<m.npair;pi++){
      const PairDev &pr = m.pair[pi]; if( pr.type != C_RIGID ) continue;
      const CellDev &cl = m.cell[pr.cell];
      if( !pr.volbox ){
        /* a rigid cell that is not a box (the non-sole shapes of a humanoid): its vertices are watched, a contact volume is
         * not formed - [EXT A-15] covers box cells; the environment is flagged when such a cell touches */
        for(int c0=0; c0<cl.nvert; c0+=32){ flag_select((pr.fofs + c0) >> 5);
          const int sh = (pr.fofs + c0) & 31, nvc = cl.nvert - c0 < 32 ? cl.nvert - c0 : 32;
          for(int k=0;k<nvc;k++) if( cfl >> (2*(sh+k)) & 1ull ) bad |= 8; }
        continue;
      }
      flag_select(pr.fofs >> 5);
      int k0 = -1; for(int k=0;k<8;k++) if( cfl >> (2*((pr.fofs & 31)+k)) & 1ull ){ k0 = k; break; }
#ifdef RKFD_VOL_DEBUG
      if( k0 >= 0 ) printf("VOLDBG pair %d k0 %d volbox %d P %d\n", pi, k0, pr.volbox, P);
#endif
      if( k0 < 0 ) continue;
      if( ncand < VOL_CAND ){ cand |= (unsigned long long)((pi << 4) | k0) << (16*ncand); ncand++; } else bad |= 4;
    }
#pragma unroll 1
    for(int ci=0; ci<VOL_CAND; ci++){
      if( !c.hany(ci < ncand) ) break;
      if( ci < ncand && P >= VOL_P ){ bad |= 4; ncand = 0; }
      if( ci < ncand ){
      const int pi = (int)(cand >> (16*ci) & 0xffffull) >> 4, k0 = (int)(cand >> (16*ci) & 0xfull);
      const PairDev &pr = m.pair[pi]; const CellDev &cl = m.cell[pr.cell];
      VolPair &v = vp[P];
      const LinkDev &L = m.link[cl.link]; const BoxDev &bx = m.box[pr.box];
      v.pair = pi; v.link = cl.link; v.fsl = Spec::frame_slot(cl.link, L); v.wsl = Spec::wext_slot(cl.link, L); v.npl = 0; v.sofs = pr.sofs; v.fofs = pr.fofs;
      v.K = pr.K; v.L = pr.L; v.SF = pr.SF; v.KF = pr.KF;
      const M3 Rw = ldm(v.fsl); const V3 pw = ld3(v.fsl+9);
      const M3 Rb = box_R(bx); const V3 pb = v3(bx.p[0],bx.p[1],bx.p[2]);
      V3 vw[8], cen = v3(0,0,0); double d[8];
      for(int k=0;k<8;k++){ const int vi = cl.vofs + k; vw[k] = pw + mul(Rw, v3(m.vert[3*vi], m.vert[3*vi+1], m.vert[3*vi+2])); cen = cen + 0.125*vw[k]; }
      V3 prob;
      { const V3 vb = tmul(Rb, vw[k0] - pb);
        box_face(bx, Rb, vb, bx.half[0]-fabs(vb.x), bx.half[1]-fabs(vb.y), bx.half[2]-fabs(vb.z), v.norm, v.a1, v.a2, prob); }
      const V3 p0 = pb + mul(Rb, prob);
      for(int k=0;k<8;k++){ d[k] = dot(v.norm, vw[k] - p0); if( fabs(d[k]) <= ZTOL ) d[k] = 0.0; }
      for(int i=0;i<36;i++) v.q6[i] = 0.0;
      for(int i=0;i<6;i++){ v.c6[i] = 0.0; v.w[i] = 0.0; }
      bool empty = false;
#pragma unroll 1
      for(int pass=0; pass<2 && !empty; pass++) empty = !vol_clip(v, vw, d, cen, p0, pass);
#ifdef RKFD_VOL_DEBUG
      printf("VOLDBG pair %d empty %d npl %d\n", pi, (int)empty, v.npl);
#endif
      /* rkCDPlaneListQuickSort with __rk_fd_plane_cmp (:376-395): ascending angle key, ties keep their order */
      if( !empty ){ double th[VOL_PL];
        for(int i=0;i<v.npl;i++){ const V3 t = cross(v.a1, v.pln[i]); const double y = norm(t); th[i] = atan2(dot(t, v.norm) > 0 ? -y : y, dot(v.a1, v.pln[i])); }
        for(int i=1;i<v.npl;i++){ const V3 tv = v.plv[i], tn = v.pln[i]; const double a = th[i]; int j = i-1;
          for(;j>=0 && !(fabs(th[j]-a) < ZTOL) && th[j] > a;j--){ v.plv[j+1] = v.plv[j]; v.pln[j+1] = v.pln[j]; th[j+1] = th[j]; }
          v.plv[j+1] = tv; v.pln[j+1] = tn; th[j+1] = a; }
        P++; }
      }
    }
    c.phase_sync(1);
    const int n = 6*P;
    /* ---- A (6P x 6P), b: probes at the volume centres (rkfd_volume.c:141-226) */
    double A[VOL_N*VOL_N], b[VOL_N];
    for(int k=0;k<P;k++){
      const VolPair &v = vp[k];
      const M3 Rw = ldm(v.fsl); const V3 pw = ld3(v.fsl+9), vl = ld3(v.fsl+12), om = ld3(v.fsl+15), al = ld3(v.fsl+18), aa = ld3(v.fsl+21);
      const V3 r = tmul(Rw, v.center - pw);
      const V3 accp = mul(Rw, al + cross(aa, r) + cross(om, cross(om, r))), acca = mul(Rw, aa);
      const V3 velp = mul(Rw, vl) + cross(mul(Rw, om), v.center - pw), vela = mul(Rw, om);
      b[6*k] = accp.x*m.dt + velp.x; b[6*k+1] = accp.y*m.dt + velp.y; b[6*k+2] = accp.z*m.dt + velp.z;
      b[6*k+3] = acca.x*m.dt + vela.x; b[6*k+4] = acca.y*m.dt + vela.y; b[6*k+5] = acca.z*m.dt + vela.z;
    }
    for(int k=0;k<P;k++){
      const VolPair &v = vp[k];
      const M3 Rw = ldm(v.fsl); const V3 pos = tmul(Rw, v.center - ld3(v.fsl+9));
      int rootk = v.link; while( Spec::parent(rootk, m.link[rootk]) >= 0 ) rootk = Spec::parent(rootk, m.link[rootk]);
      for(int i=0;i<6;i++){
        const V3 e = v3(i%3 == 0 ? 1.0 : 0.0, i%3 == 1 ? 1.0 : 0.0, i%3 == 2 ? 1.0 : 0.0);
        const V3 el = tmul(Rw, e);
        const V3 dpf = i < 3 ? -el : v3(0,0,0), dpn = i < 3 ? -cross(pos, el) : -el;
        for(int j=0;j<P;j++){
          const VolPair &u = vp[j];
          int rootj = u.link; while( Spec::parent(rootj, m.link[rootj]) >= 0 ) rootj = Spec::parent(rootj, m.link[rootj]);
          V3 rl = v3(0,0,0), ra = v3(0,0,0);
          if( rootj == rootk ){
            V3 ral, raa; vol_probe(m, v.link, dpf, dpn, u.link, ral, raa);
            const M3 Ru = ldm(u.fsl); const V3 ru = tmul(Ru, u.center - ld3(u.fsl+9));
            rl = mul(Ru, ral + cross(raa, ru)); ra = mul(Ru, raa);
          }
          A[VOL_N*(6*j)+6*k+i] = rl.x; A[VOL_N*(6*j+1)+6*k+i] = rl.y; A[VOL_N*(6*j+2)+6*k+i] = rl.z;
          A[VOL_N*(6*j+3)+6*k+i] = ra.x; A[VOL_N*(6*j+4)+6*k+i] = ra.y; A[VOL_N*(6*j+5)+6*k+i] = ra.z;
        }
      }
    }
    c.phase_sync(1);
    /* ---- QP (rkfd_volume.c:496-548) */
    double Qm[VOL_N*VOL_N], cv[VOL_N], nf[VOL_M*VOL_N], x[VOL_N];
    for(int i=0;i<n;i++){ cv[i] = 0.0; x[i] = 0.0; for(int j=0;j<n;j++) Qm[VOL_N*i+j] = 0.0; }
    for(int k=0;k<P;k++){
      const VolPair &v = vp[k];
      double T[6*VOL_N];       /* T = Q6 A_k (6 x n) */
      for(int i=0;i<6;i++) for(int s=0;s<n;s++){ double t = 0;
RKFD_VOL_U
        for(int j=0;j<6;j++) t += v.q6[6*i+j]*A[VOL_N*(6*k+j)+s];
        T[VOL_N*i+s] = t; }
      for(int r=0;r<n;r++) for(int s=0;s<n;s++){ double t = 0;
RKFD_VOL_U
        for(int i=0;i<6;i++) t += A[VOL_N*(6*k+i)+r]*T[VOL_N*i+s];
        Qm[VOL_N*r+s] += t; }
      for(int i=0;i<6;i++){ double t = v.c6[i];
RKFD_VOL_U
        for(int j=0;j<6;j++) t += v.q6[6*i+j]*b[6*k+j];
        for(int r=0;r<n;r++) cv[r] += t*A[VOL_N*(6*k+i)+r]; }
    }
    for(int k=0;k<P;k++) for(int i=0;i<6;i++) Qm[VOL_N*(6*k+i)+6*k+i] += vp[k].L;
    int mrows = 0;
    for(int k=0;k<P;k++){
      const VolPair &v = vp[k];
      for(int j=0;j<n;j++) nf[VOL_N*mrows+j] = 0.0;
      nf[VOL_N*mrows+6*k] = v.norm.x; nf[VOL_N*mrows+6*k+1] = v.norm.y; nf[VOL_N*mrows+6*k+2] = v.norm.z;
      mrows++;
      for(int i=0;i<v.npl;i++){
        const double a = -dot(v.pln[i], v.plv[i]), b1 = dot(v.pln[i], v.a2), b2 = -dot(v.pln[i], v.a1);
        for(int j=0;j<n;j++) nf[VOL_N*mrows+j] = 0.0;
        const V3 f = a*v.norm, t = b1*v.a1 + b2*v.a2;
        nf[VOL_N*mrows+6*k] = f.x; nf[VOL_N*mrows+6*k+1] = f.y; nf[VOL_N*mrows+6*k+2] = f.z;
        nf[VOL_N*mrows+6*k+3] = t.x; nf[VOL_N*mrows+6*k+4] = t.y; nf[VOL_N*mrows+6*k+5] = t.z;
        mrows++;
      }
      x[6*k] = v.norm.x; x[6*k+1] = v.norm.y; x[6*k+2] = v.norm.z;
    }
    /* a single pair: identity block for the unused unknowns (they stay zero) */
    for(int i=n;i<VOL_N;i++){ cv[i] = 0.0; x[i] = 0.0;
      for(int j=0;j<VOL_N;j++){ Qm[VOL_N*i+j] = i == j ? 1.0 : 0.0; Qm[VOL_N*j+i] = i == j ? 1.0 : 0.0; }
      for(int r=0;r<mrows;r++) nf[VOL_N*r+i] = 0.0; }
    unsigned idx = 0; VOL_STAT(6, 1);
    c.phase_sync(1);
    unsigned wkey = 0;      /* the re-sort key: final active set, iterations, pairs, friction paths of the pairs (hashed) */
    if( P > 0 ){ vol_asm(mrows, Qm, cv, nf, x, idx); wkey = idx ^ (wk << 20) ^ ((unsigned)P << 27); }
#ifdef RKFD_VOL_DEBUG
    for(int k=0;k<P;k++){ printf("VOLDBG pair %d npl %d center %.6e %.6e %.6e x", vp[k].pair, vp[k].npl, vp[k].center.x, vp[k].center.y, vp[k].center.z); for(int i=0;i<6;i++) printf(" %.6e", x[6*k+i]); printf(" idx %x mrows %d bad %d c6", idx, mrows, bad); for(int i=0;i<6;i++) printf(" %.4e", vp[k].c6[i]); printf("\n"); }
#endif
    c.phase_sync(1);
    /* ---- f /= dt, _rkFDSolverSetForce (:552-568; the offset is not advanced for a pair without planes - mirrored) */
    { int off = 0;
      for(int k=0;k<P;k++){ VolPair &v = vp[k];
        if( v.npl == 0 ){ for(int i=0;i<6;i++) v.w[i] = 0.0; continue; }
        for(int i=0;i<6;i++) v.w[i] = x[off+i]/m.dt;
        const V3 f = v3(v.w[0],v.w[1],v.w[2]);
        if( ( vtiny(f.x) && vtiny(f.y) && vtiny(f.z) ) || dot(f, v.norm) < ZTOL ) for(int i=0;i<6;i++) v.w[i] = 0.0;
        off += 6; } }
    /* The pairs in lockstep (trip count VOL_P for every lane) with the lanes of a warp brought together before each of the
     * two linear programs: reached from the nested decisions below lane by lane they ran with 3 of 32 lanes (ncu). */
#pragma unroll 1
    for(int k=0;k<VOL_P;k++){
      const bool in = k < P;
      VolPair &v = vp[in ? k : 0]; const int np = in ? v.npl : 0;
      V3 wf = v3(0,0,0), wt = v3(0,0,0);
      double wv[6] = {0,0,0,0,0,0};
      int kinetic = 0; bool setforce = false, try_static = false, mod_w = false;
      if( in ){
      wf = v3(v.w[0],v.w[1],v.w[2]); wt = v3(v.w[3],v.w[4],v.w[5]);
      /* ---- _rkFDSolverModifyNormalForceCenter (:580-631) */
      { const double fn = dot(v.norm, wf);
        if( !(fn < ZTOL) && np >= 3 ){
          const V3 r0 = (-dot(v.a2, wt)/fn)*v.a1 + (dot(v.a1, wt)/fn)*v.a2;
          bool flag = false; int i0 = np-3, i1 = np-2, i2 = np-1, i3 = 0;
          for(int it=0; it<np; it++, i0=i1, i1=i2, i2=i3, i3=i3+1){
            const V3 dir = v.plv[i2] - v.plv[i1]; const double d = dot(dir, dir);
            if( vtiny(d) ) continue;
            const V3 tmp = r0 - v.plv[i1];
            if( dot(tmp, v.pln[i1]) > ZTOL ) continue;
            const double s = dot(dir, tmp)/d;
            V3 r; int mod;
            if( s < ZTOL ){
              if( flag ) break;
              const V3 t2 = (v.plv[i0] - v.plv[i1]) + dir;
              r = v.plv[i1] + (ZTOL/norm(t2))*t2; mod = 1;
            } else if( s < 1.0-ZTOL ){
              r = v.plv[i1] + s*dir + ZTOL*v.pln[i1]; mod = 1;
            } else {
              const V3 t2 = (v.plv[i3 < np ? i3 : 0] - v.plv[i2]) - dir;
              r = v.plv[i2] + (ZTOL/norm(t2))*t2; mod = 2;
            }
            wt = dot(v.norm, wt)*v.norm + (fn*dot(v.a2, r))*v.a1 + (-fn*dot(v.a1, r))*v.a2;
            if( mod == 1 ) break;
            flag = true;
          }
        } }
      /* ---- _rkFDSolverModifyWrench (:869-916) */
      if( np > 0 && !vtiny(dot(wf, v.norm)) ){
        mod_w = true;
        wv[0] = dot(wf, v.norm); wv[1] = dot(wf, v.a1); wv[2] = dot(wf, v.a2); wv[3] = dot(wt, v.norm); wv[4] = dot(wt, v.a1); wv[5] = dot(wt, v.a2);
        const double fn = wv[0], fs = sqrt(wv[1]*wv[1] + wv[2]*wv[2]);
        double tl = 0;
        for(int i=0;i<np;i++){ v.r[i][0] = dot(v.plv[i], v.a1); v.r[i][1] = dot(v.plv[i], v.a2);
          const double rl = sqrt(v.r[i][0]*v.r[i][0] + v.r[i][1]*v.r[i][1]); if( tl < rl ) tl = rl; }
        if( vtiny(tl) ){
          wv[3] = wv[4] = wv[5] = 0;
          if( !vtiny(fs) && fs > v.SF*fn ){
            V3 vel; const double t = vol_slip(m, v, v.center, vel)*wv[0];
            wv[1] = -t*dot(vel, v.a1); wv[2] = -t*dot(vel, v.a2);
            kinetic = 1;
          }
          setforce = true;
        } else if( ( !vtiny(fs) && fs > v.SF*fn ) || fabs(wv[3]) > tl*wv[0] ){
          kinetic = 2;
        } else try_static = true;
      } }
      wkey ^= ( try_static ? 1u : ( kinetic == 2 ? 2u : 0u ) ) << (29 + 2*(k & 1));
      c.hsync();
      if( try_static ){
        /* static friction: the wrench inside the friction pyramids at the polygon corners? (:643-688) */
        const double mb[6] = { wv[0], wv[4], wv[5], wv[1], wv[2], wv[3] };
        if( !vol_static_feasible(m, v, np, mb) ) kinetic = 2;
      }
      c.hsync();
      if( mod_w ){
        if( kinetic == 2 ){
          /* kinetic friction: normal force redistributed over the polygon corners by an LP (:733-843) */
          double ma[3*VOL_LPN], mb[3], mc[VOL_PL], mf[VOL_PL], wn[3];
          for(int j=0;j<np;j++){ ma[j] = 1.0; ma[VOL_LPN+j] = v.r[j][1]; ma[2*VOL_LPN+j] = -v.r[j][0]; mf[j] = 0.0; }
          mb[0] = wv[0]; mb[1] = wv[4]; mb[2] = wv[5];
          for(int i=0;i<3;i++) wn[i] = vtiny(wv[i+1]) ? 0.0 : 1.0/wv[i+1];
          for(int j=0;j<np;j++){
            V3 vel; const double ww = vol_slip(m, v, v.center + v.plv[j], vel);
            v.s[j][0] = -ww*dot(vel, v.a1); v.s[j][1] = -ww*dot(vel, v.a2);
            mc[j] = -wn[0]*v.s[j][0] - wn[1]*v.s[j][1] - wn[2]*( v.r[j][0]*v.s[j][1] - v.r[j][1]*v.s[j][0] );
          }
          if( !vol_lp(3, np, ma, mb, mc, mf) ){
            double wn2[2]; for(int i=0;i<2;i++) wn2[i] = vtiny(wv[i+3]) ? 0.0 : 1.0/wv[i+3];
            for(int j=0;j<np;j++){ mc[j] += wn2[0]*v.r[j][0] - wn2[1]*v.r[j][1]; mf[j] = 0.0; }
            vol_lp(1, np, ma, mb, mc, mf);
          }
          wv[1] = wv[2] = wv[3] = 0;
          for(int j=0;j<np;j++){ const double fx = v.s[j][0]*mf[j], fy = v.s[j][1]*mf[j];
            wv[1] += fx; wv[2] += fy; wv[3] += v.r[j][0]*fy - v.r[j][1]*fx; }
          setforce = true;
        }
        if( ref ){ flag_select(v.fofs >> 5); if( kinetic ) cfl |= 2ull << (2*(v.fofs & 31)); else cfl &= ~(2ull << (2*(v.fofs & 31))); }
        if( setforce ){ wf = wv[0]*v.norm + wv[1]*v.a1 + wv[2]*v.a2; wt = wv[3]*v.norm + wv[4]*v.a1 + wv[5]*v.a2; }
      }
      /* ---- _rkFDSolverPushWrench (:919-936) */
      if( in ){ const M3 Rw = ldm(v.fsl); const V3 pos = tmul(Rw, v.center - ld3(v.fsl+9));
        const V3 fl_ = tmul(Rw, wf), tl_ = tmul(Rw, wt) + cross(pos, fl_);
        c.S(v.wsl) += fl_.x; c.S(v.wsl+1) += fl_.y; c.S(v.wsl+2) += fl_.z;
        c.S(v.wsl+3) += tl_.x; c.S(v.wsl+4) += tl_.y; c.S(v.wsl+5) += tl_.z;
        if( ref ){   /* results of the pair in its first three contact slots: force, torque, centre */
          const int s = v.sofs;
          c.gst(c.st.cf,3*s,wf.x); c.gst(c.st.cf,3*s+1,wf.y); c.gst(c.st.cf,3*s+2,wf.z);
          c.gst(c.st.cf,3*s+3,wt.x); c.gst(c.st.cf,3*s+4,wt.y); c.gst(c.st.cf,3*s+5,wt.z);
          c.gst(c.st.cf,3*s+6,v.center.x); c.gst(c.st.cf,3*s+7,v.center.y); c.gst(c.st.cf,3*s+8,v.center.z);
        } }
    }
    wk = P > 0 ? 1u + (wkey*2654435761u >> 24)%255u : 0u;
    c.phase_sync(1);
  }
