/* rkfd_core.cuh - one environment's rkFDUpdate, written once for the sm_100a kernel.
 *
 * Thread-per-environment formulation of the reference step (reference src/rkfd_sim.c:560-566):
 * 4 Runge-Kutta-Gill stage evaluations + 1 committing reference evaluation, each evaluation =
 *   pass 1 (outward)  FK + link velocities          <- _rkFDConnectJointState   rkfd_sim.c:290-302
 *                     vertex/box collision          <- _rkFDUpdateCD            rkfd_sim.c:466-471
 *                     penalty contact + Coulomb     <- rkFDSolverPenalty        rkfd_penalty.c:11-31
 *   pass 2 (inward)   joint friction / motor torque <- rkFDJointFriction        rkfd_util.c:366-387
 *                     articulated inertia + bias    <- rkChainUpdateABI [EXT]   rkfd_sim.c:509-520
 *   pass 3 (outward)  accelerations, q''            <- rkChainGetJointAccAll    rkfd_sim.c:521
 *                     RKG stage bookkeeping         <- zODE2Update [EXT]        rkfd_sim.c:562
 * The per-link quantities that must survive between passes live in a per-thread scratch column
 * (shared memory, element k of thread t at [k*BLOCK + t]: conflict-free 64-bit accesses); the
 * running articulated inertia of a serial chain segment stays in registers.
 *
 * The same source is compiled (a) by nvcc into the product kernel and (b) by g++ into a host
 * harness that only tests/ loads, to debug the arithmetic where no GPU exists.  The product
 * library contains (a) only.
 */
#ifndef RKFD_CORE_CUH
#define RKFD_CORE_CUH

#include "rkfd_types.h"
#include "rkfd_math.cuh"

#if defined(__CUDACC__)
#define RKFD_NOINLINE __host__ __device__ __noinline__
#else
#define RKFD_NOINLINE
#endif

namespace rkfd {

/* scratch slots per link by joint type */
RKFD_HD int link_slot_count(int jtype, int has_rigid){
  switch(jtype){
    case J_REVOL: case J_PRISM: return 10;    /* w,gd->U (6), sin, cos, Dinv, u */
    case J_SPHER: return 36;                  /* U (18, w/gd aliased), Dinv (6), u (3), Rrel (9) */
    case J_FLOAT: return has_rigid ? 39 : 18; /* a0 (6, w/gd aliased), Rrel (9), prel (3) [, IA^-1 (21)] */
    default: return 6;                        /* w, gd */
  }
}
constexpr int BRANCH_SLOTS = 15;   /* pass 1: Rw(9) pw(3) vl(3); pass 3: a(6) w(3) */
constexpr int ACCUM_SLOTS = 27;    /* A(6) B(9) C(6) pf(3) pn(3) */
constexpr int WEXT_SLOTS = 6;

enum StageMode : int { ST_K1 = 0, ST_K2 = 1, ST_K3 = 2, ST_K4 = 3, ST_REF = 4, ST_EVAL = 5, ST_EVAL_REF = 6 };

/* symmetric 6x6 inverse through Cholesky; a is the full row-major matrix, overwritten */
RKFD_NOINLINE inline void spd6_inverse(double (&a)[36]){
  double L[36], Li[36];
#pragma unroll
  for(int k=0;k<36;k++){ L[k]=0; Li[k]=0; }
#pragma unroll
  for(int j=0;j<6;j++){
    double s = a[6*j+j];
#pragma unroll
    for(int k=0;k<6;k++) if(k<j) s -= L[6*j+k]*L[6*j+k];
    L[6*j+j] = sqrt(s);
    double inv = 1.0/L[6*j+j];
#pragma unroll
    for(int i=0;i<6;i++) if(i>j){
      double t = a[6*i+j];
#pragma unroll
      for(int k=0;k<6;k++) if(k<j) t -= L[6*i+k]*L[6*j+k];
      L[6*i+j] = t*inv;
    }
  }
#pragma unroll
  for(int j=0;j<6;j++){
    Li[6*j+j] = 1.0/L[6*j+j];
#pragma unroll
    for(int i=0;i<6;i++) if(i>j){
      double s = 0;
#pragma unroll
      for(int k=0;k<6;k++) if(k>=j && k<i) s -= L[6*i+k]*Li[6*k+j];
      Li[6*i+j] = s/L[6*i+i];
    }
  }
#pragma unroll
  for(int i=0;i<6;i++)
#pragma unroll
    for(int j=0;j<6;j++){
      double s = 0;
#pragma unroll
      for(int k=0;k<6;k++) if(k>=i && k>=j) s += Li[6*k+i]*Li[6*k+j];
      a[6*i+j] = s;
    }
}

/* symmetric 3x3 inverse (adjugate) */
RKFD_HD S3 sym3_inverse(const S3 &d){
  S3 c;
  c.xx = d.yy*d.zz - d.yz*d.yz; c.xy = d.xz*d.yz - d.xy*d.zz; c.xz = d.xy*d.yz - d.xz*d.yy;
  c.yy = d.xx*d.zz - d.xz*d.xz; c.yz = d.xy*d.xz - d.xx*d.yz; c.zz = d.xx*d.yy - d.xy*d.xy;
  double det = d.xx*c.xx + d.xy*c.xy + d.xz*c.xz, inv = 1.0/det;
  c.xx*=inv; c.xy*=inv; c.xz*=inv; c.yy*=inv; c.yz*=inv; c.zz*=inv;
  return c;
}

template <class Ctx>
struct Core {
  Ctx &c;
  unsigned int piv;             /* joint friction pivot: bit j = dof j kinetic */
  unsigned long long cfl;       /* contact flags: bit 2s active, bit 2s+1 kinetic */
  int bad;

  RKFD_HD explicit Core(Ctx &ctx) : c(ctx), piv(0), cfl(0), bad(0) {}

  RKFD_HD V3 ld3(int k){ return v3(c.S(k), c.S(k+1), c.S(k+2)); }
  RKFD_HD void st3(int k, V3 v){ c.S(k)=v.x; c.S(k+1)=v.y; c.S(k+2)=v.z; }
  RKFD_HD M3 ldm(int k){ M3 m; m.xx=c.S(k); m.xy=c.S(k+1); m.xz=c.S(k+2); m.yx=c.S(k+3); m.yy=c.S(k+4); m.yz=c.S(k+5); m.zx=c.S(k+6); m.zy=c.S(k+7); m.zz=c.S(k+8); return m; }
  RKFD_HD void stm(int k, const M3 &m){ c.S(k)=m.xx; c.S(k+1)=m.xy; c.S(k+2)=m.xz; c.S(k+3)=m.yx; c.S(k+4)=m.yy; c.S(k+5)=m.yz; c.S(k+6)=m.zx; c.S(k+7)=m.zy; c.S(k+8)=m.zz; }
  RKFD_HD S3 lds(int k){ S3 s; s.xx=c.S(k); s.xy=c.S(k+1); s.xz=c.S(k+2); s.yy=c.S(k+3); s.yz=c.S(k+4); s.zz=c.S(k+5); return s; }

  static RKFD_HD M3 org_R(const LinkDev &L){ M3 m; m.xx=L.Ro[0]; m.xy=L.Ro[1]; m.xz=L.Ro[2]; m.yx=L.Ro[3]; m.yy=L.Ro[4]; m.yz=L.Ro[5]; m.zx=L.Ro[6]; m.zy=L.Ro[7]; m.zz=L.Ro[8]; return m; }
  static RKFD_HD V3 org_p(const LinkDev &L){ return v3(L.po[0],L.po[1],L.po[2]); }

  /* link frame w.r.t. parent (R,p) and joint velocity (vJ,wJ, link frame) from the stage state and
   * the joint data cached by pass 1 ([EXT A-3]) */
  RKFD_HD void joint_xform(const ModelDev &m, const LinkDev &L, M3 &R, V3 &p, V3 &vJ, V3 &wJ){
    const int sl = L.slot, qs = m.rk_slot + L.qofs, qds = m.rk_slot + m.nq + L.qofs;
    vJ = v3(0,0,0); wJ = v3(0,0,0);
    switch(L.jtype){
    case J_REVOL: {
      const double s = c.S(sl+6), co = c.S(sl+7);
      const M3 Ro = org_R(L); const V3 o0 = col0(Ro), o1 = col1(Ro);
      R = from_cols(co*o0 + s*o1, co*o1 - s*o0, col2(Ro)); p = org_p(L);
      wJ.z = c.S(qds);
    } break;
    case J_PRISM: {
      R = org_R(L); p = org_p(L) + c.S(qs)*col2(R); vJ.z = c.S(qds);
    } break;
    case J_SPHER: {
      R = ldm(sl+27); p = org_p(L);
      wJ = tmul(R, mul(org_R(L), ld3(qds)));
    } break;
    case J_FLOAT: {
      R = ldm(sl+6); p = ld3(sl+15);
      const M3 Ro = org_R(L);
      vJ = tmul(R, mul(Ro, ld3(qds))); wJ = tmul(R, mul(Ro, ld3(qds+3)));
    } break;
    default: R = org_R(L); p = org_p(L); break;
    }
  }

  /* ---- contact of the cells carried by link i: vertex-in-box detection ([EXT A-10]), elastic pairs:
   * penalty force + Coulomb clamp + wrench accumulation (rkfd_penalty.c:11-31, rkfd_util.c:239-282) */
  RKFD_HD V6 contacts(const ModelDev &m, const LinkDev &L, const M3 &Rw, V3 pw, V3 vl, V3 om, bool ref){
    V6 w; w.l = v3(0,0,0); w.a = v3(0,0,0);
    for(int ci=L.cell_begin; ci<L.cell_end; ci++){
      const CellDev &cl = m.cell[ci];
      for(int pi=cl.pair_begin; pi<cl.pair_end; pi++){
        const PairDev &pr = m.pair[pi]; const BoxDev &bx = m.box[pr.box];
        M3 Rb; Rb.xx=bx.R[0]; Rb.xy=bx.R[1]; Rb.xz=bx.R[2]; Rb.yx=bx.R[3]; Rb.yy=bx.R[4]; Rb.yz=bx.R[5]; Rb.zx=bx.R[6]; Rb.zy=bx.R[7]; Rb.zz=bx.R[8];
        const V3 pb = v3(bx.p[0],bx.p[1],bx.p[2]);
        for(int k=0;k<cl.nvert;k++){
          const int s = pr.sofs + k;
          const unsigned long long abit = 1ull << (2*s), kbit = 2ull << (2*s);
          const V3 vloc = v3(m.vert[3*(cl.vofs+k)], m.vert[3*(cl.vofs+k)+1], m.vert[3*(cl.vofs+k)+2]);
          const V3 vw = pw + mul(Rw, vloc);
          const V3 vb = tmul(Rb, vw - pb);
          const double dx = bx.half[0]-fabs(vb.x), dy = bx.half[1]-fabs(vb.y), dz = bx.half[2]-fabs(vb.z);
          const bool inside = (dx > -ZTOL) && (dy > -ZTOL) && (dz > -ZTOL);
          if( !inside ){ cfl &= ~(abit|kbit); if( ref ){ c.gst(c.st.cf, 3*s, 0.0); c.gst(c.st.cf, 3*s+1, 0.0); c.gst(c.st.cf, 3*s+2, 0.0); } continue; }
          /* closest face: first minimum over x,y,z */
          int amin = 0; double dmin = dx;
          if( dy < dmin ){ dmin = dy; amin = 1; }
          if( dz < dmin ){ dmin = dz; amin = 2; }
          const double vba = amin==0 ? vb.x : ( amin==1 ? vb.y : vb.z );
          const double sg = vba >= 0 ? 1.0 : -1.0;
          const V3 b0 = col0(Rb), b1 = col1(Rb), b2 = col2(Rb);
          const V3 n  = sg*( amin==0 ? b0 : ( amin==1 ? b1 : b2 ) );
          const V3 t1 =      amin==0 ? b1 : ( amin==1 ? b2 : b0 );
          const V3 t2 = sg*( amin==0 ? b2 : ( amin==1 ? b0 : b1 ) );
          V3 prob = vb;
          if( amin==0 ) prob.x = sg*bx.half[0]; else if( amin==1 ) prob.y = sg*bx.half[1]; else prob.z = sg*bx.half[2];
          V3 refb;
          if( !(cfl & abit) ){           /* new contact: {SF, _ref = _pro} */
            cfl = (cfl | abit) & ~kbit; refb = prob;
            c.gst(c.st.cref, 3*s, refb.x); c.gst(c.st.cref, 3*s+1, refb.y); c.gst(c.st.cref, 3*s+2, refb.z);
          } else refb = v3(c.gld(c.st.cref,3*s), c.gld(c.st.cref,3*s+1), c.gld(c.st.cref,3*s+2));
          if( pr.type != C_ELASTIC ) continue;     /* rigid pairs are solved by the rigid path */
          const V3 refw = pb + mul(Rb, refb);
          const V3 d = vw - refw;
          /* rkFDLinkPointWldVel (rkfd_util.c:14-24); the static partner contributes 0 */
          const V3 vr = mul(Rw, vl) + cross(mul(Rw, om), vw - pw);
          V3 f = (-pr.E)*d + (-1.0*(pr.V + pr.E*m.dt))*vr;
          if( dot(f,n) < 0.0 ){ if( ref ){ c.gst(c.st.cf,3*s,f.x); c.gst(c.st.cf,3*s+1,f.y); c.gst(c.st.cf,3*s+2,f.z); } continue; }
          /* rkFDContactForceModifyFriction */
          const double fn = dot(f,n), f1 = dot(f,t1), f2 = dot(f,t2);
          const double fs = sqrt(f1*f1 + f2*f2);
          const double mu = (cfl & kbit) ? pr.KF : pr.SF;
          if( !(fabs(fs) < ZTOL) && fs > mu*fn ){
            V3 v = vr + (-dot(vr,n))*n;
            const double vs = norm(v);
            f = fn*n;
            if( !(fabs(vs) < ZTOL) ){
              v = v3(v.x/vs, v.y/vs, v.z/vs);
              f = f + (-(1.0 - exp(-1.0*m.friction_weight*vs))*pr.KF*fn)*v;
            }
            if( ref ){ cfl |= kbit; c.gst(c.st.cref,3*s,prob.x); c.gst(c.st.cref,3*s+1,prob.y); c.gst(c.st.cref,3*s+2,prob.z); }
          } else if( ref ) cfl &= ~kbit;
          /* rkFDContactForcePushWrench: (f, pos x f) at the link origin, link axes */
          const V3 pos = tmul(Rw, vw - pw);
          const V3 fl = tmul(Rw, f);
          w.l = w.l + fl; w.a = w.a + cross(pos, fl);
          if( ref ){ c.gst(c.st.cf,3*s,f.x); c.gst(c.st.cf,3*s+1,f.y); c.gst(c.st.cf,3*s+2,f.z); }
        }
      }
    }
    return w;
  }

  /* ---- pass 1: outward kinematics + collision + penalty */
  RKFD_HD void pass1(const ModelDev &m, bool ref){
    M3 Rw = ident3(); V3 pw = v3(0,0,0), vl = v3(0,0,0), om = v3(0,0,0), gd = v3(0,0,-GRAVITY);
    const int qs = m.rk_slot, qds = m.rk_slot + m.nq;
    for(int i=0;i<m.nl;i++){
      const LinkDev &L = m.link[i]; const int sl = L.slot;
      if( !L.serial ){
        if( L.parent < 0 ){ Rw = ident3(); pw = v3(0,0,0); vl = v3(0,0,0); om = v3(0,0,0); gd = v3(0,0,-GRAVITY); }
        else {
          const LinkDev &P = m.link[L.parent];
          om = ld3(P.slot); gd = ld3(P.slot+3);
          if( m.need_world ){ Rw = ldm(P.branch_slot); pw = ld3(P.branch_slot+9); vl = ld3(P.branch_slot+12); }
        }
      }
      M3 R; V3 p, vJ = v3(0,0,0), wJ = v3(0,0,0);
      const M3 Ro = org_R(L);
      switch(L.jtype){
      case J_REVOL: {
        double s, co; sincos(c.S(qs+L.qofs), &s, &co);
        c.S(sl+6) = s; c.S(sl+7) = co;
        const V3 o0 = col0(Ro), o1 = col1(Ro);
        R = from_cols(co*o0 + s*o1, co*o1 - s*o0, col2(Ro)); p = org_p(L);
        wJ.z = c.S(qds+L.qofs);
      } break;
      case J_PRISM: R = Ro; p = org_p(L) + c.S(qs+L.qofs)*col2(Ro); vJ.z = c.S(qds+L.qofs); break;
      case J_SPHER: {
        R = mm(Ro, aa_to_mat(ld3(qs+L.qofs))); p = org_p(L);
        stm(sl+27, R);
        wJ = tmul(R, mul(Ro, ld3(qds+L.qofs)));
      } break;
      case J_FLOAT: {
        R = mm(Ro, aa_to_mat(ld3(qs+L.qofs+3))); p = org_p(L) + mul(Ro, ld3(qs+L.qofs));
        stm(sl+6, R); st3(sl+15, p);
        vJ = tmul(R, mul(Ro, ld3(qds+L.qofs))); wJ = tmul(R, mul(Ro, ld3(qds+L.qofs+3)));
      } break;
      default: R = Ro; p = org_p(L); break;
      }
      const V3 om_n = tmul(R, om) + wJ;
      const V3 gd_n = tmul(R, gd);
      if( m.need_world ){
        const V3 vl_n = tmul(R, vl + cross(om, p)) + vJ;
        pw = pw + mul(Rw, p); Rw = mm(Rw, R); vl = vl_n;
      }
      om = om_n; gd = gd_n;
      st3(sl, om); st3(sl+3, gd);
      if( L.accum_slot >= 0 ) for(int k=0;k<ACCUM_SLOTS;k++) c.S(L.accum_slot+k) = 0.0;
      if( L.wext_slot >= 0 ){
        const V6 w = contacts(m, L, Rw, pw, vl, om, ref);
        st3(L.wext_slot, w.l); st3(L.wext_slot+3, w.a);
      }
      if( L.branch_slot >= 0 && m.need_world ){ stm(L.branch_slot, Rw); st3(L.branch_slot+9, pw); st3(L.branch_slot+12, vl); }
    }
  }

  /* motor + joint friction of a 1-DoF joint: returns tau = driving torque + friction, jm = rotor inertia
   * (rkfd_util.c:330-364, [EXT A-6, A-7]); at the reference stage commits pivot type and prev_trq */
  RKFD_HD double joint_torque(const ModelDev &m, const LinkDev &L, int i, bool ref, double &jm){
    const int j = L.qofs;
    const double v = c.S(m.rk_slot + m.nq + j);
    double tdrive = 0.0, tf = 0.0; jm = 0.0;
    if( L.mtype != M_NONE ){
      double e = c.gld(c.st.u, i);
      e = e < L.m_min ? L.m_min : ( e > L.m_max ? L.m_max : e );
      if( L.mtype == M_DC ){
        const double tin = L.m_tin*e, treg = L.m_reg*v;
        jm = L.m_jm; tdrive = tin - treg;
        tf = jm; tf *= -v / m.dt; tf -= tin; tf += treg; tf += c.gld(c.st.piv_prev, j);
        double fmax;
        if( !(piv & (1u<<j)) ) fmax = L.sfriction;
        else {
          const double sg = v > 0 ? 1.0 : ( v < 0 ? -1.0 : 0.0 );
          fmax = -L.stiffness*c.S(m.rk_slot + j) - L.viscosity*v - L.coulomb*sg;
        }
        fmax = fabs(fmax);
        if( fabs(tf) > fmax ){ tf = tf > 0 ? fmax : -fmax; if( ref ) piv |= (1u<<j); }
        else if( ref ) piv &= ~(1u<<j);
      } else tdrive = e;
    }
    if( ref ) c.gst(c.st.piv_prev, j, tdrive + tf);     /* rkFDUpdateJointPrevDrivingTrq (rkfd_util.c:289-311) */
    return tdrive + tf;
  }

  /* ---- pass 2: inward articulated-inertia pass */
  RKFD_HD void pass2(const ModelDev &m, bool ref){
    S3 kA, kC; M3 kB; V3 kf, kn;          /* contribution carried to link i from its serial child */
    kA.xx=kA.xy=kA.xz=kA.yy=kA.yz=kA.zz=0; kC = kA; kB.xx=kB.xy=kB.xz=kB.yx=kB.yy=kB.yz=kB.zx=kB.zy=kB.zz=0; kf = v3(0,0,0); kn = kf;
    for(int i=m.nl-1;i>=0;i--){
      const LinkDev &L = m.link[i]; const int sl = L.slot;
      const V3 om = ld3(sl), gd = ld3(sl+3);
      const V3 mc = v3(L.mc[0],L.mc[1],L.mc[2]);
      S3 A, C; M3 B;
      A.xx = L.mass; A.xy = 0; A.xz = 0; A.yy = L.mass; A.yz = 0; A.zz = L.mass;
      B.xx = 0; B.xy = mc.z; B.xz = -mc.y; B.yx = -mc.z; B.yy = 0; B.yz = mc.x; B.zx = mc.y; B.zy = -mc.x; B.zz = 0;
      C.xx = L.Io[0]; C.xy = L.Io[1]; C.xz = L.Io[2]; C.yy = L.Io[3]; C.yz = L.Io[4]; C.zz = L.Io[5];
      /* bias ( w x (w x mc) ; w x (Io w) ) minus gravity (m gd ; mc x gd) minus external wrench */
      V3 pf = cross(om, cross(om, mc)) - L.mass*gd;
      V3 pn = cross(om, mul(C, om)) - cross(mc, gd);
      if( L.wext_slot >= 0 ){ pf = pf - ld3(L.wext_slot); pn = pn - ld3(L.wext_slot+3); }
      if( L.accum_slot >= 0 ){
        const int a = L.accum_slot; const S3 aA = lds(a), aC = lds(a+15); const M3 aB = ldm(a+6);
        A.xx+=aA.xx; A.xy+=aA.xy; A.xz+=aA.xz; A.yy+=aA.yy; A.yz+=aA.yz; A.zz+=aA.zz;
        C.xx+=aC.xx; C.xy+=aC.xy; C.xz+=aC.xz; C.yy+=aC.yy; C.yz+=aC.yz; C.zz+=aC.zz;
        B.xx+=aB.xx; B.xy+=aB.xy; B.xz+=aB.xz; B.yx+=aB.yx; B.yy+=aB.yy; B.yz+=aB.yz; B.zx+=aB.zx; B.zy+=aB.zy; B.zz+=aB.zz;
        pf = pf + ld3(a+21); pn = pn + ld3(a+24);
      }
      if( i+1 < m.nl && m.link[i+1].serial ){
        A.xx+=kA.xx; A.xy+=kA.xy; A.xz+=kA.xz; A.yy+=kA.yy; A.yz+=kA.yz; A.zz+=kA.zz;
        C.xx+=kC.xx; C.xy+=kC.xy; C.xz+=kC.xz; C.yy+=kC.yy; C.yz+=kC.yz; C.zz+=kC.zz;
        B.xx+=kB.xx; B.xy+=kB.xy; B.xz+=kB.xz; B.yx+=kB.yx; B.yy+=kB.yy; B.yz+=kB.yz; B.zx+=kB.zx; B.zy+=kB.zy; B.zz+=kB.zz;
        pf = pf + kf; pn = pn + kn;
      }
      M3 R; V3 p, vJ, wJ;
      joint_xform(m, L, R, p, vJ, wJ);
      /* velocity-product acceleration (link frame): parent angular velocity in link axes = om - wJ */
      const V3 omp = om - wJ;
      const V3 zl = cross(omp, cross(omp, tmul(R, p))) + 2.0*cross(omp, vJ);
      const V3 za = cross(omp, wJ);
      /* p' = pA + IA zeta */
      if( L.jtype != J_FLOAT ){
        pf = pf + mul(A, zl) + mul(B, za);
        pn = pn + tmul(B, zl) + mul(C, za);
      }
      switch(L.jtype){
      case J_REVOL: {
        double jm; const double tau = joint_torque(m, L, i, ref, jm);
        const V3 Ul = col2(B), Ua = v3(C.xz, C.yz, C.zz);
        const double Dinv = 1.0/(C.zz + jm), u = tau - pn.z;
        st3(sl, Ul); st3(sl+3, Ua); c.S(sl+8) = Dinv; c.S(sl+9) = u;
        const V3 Wl = Dinv*Ul, Wa = Dinv*Ua;
        A.xx-=Wl.x*Ul.x; A.xy-=Wl.x*Ul.y; A.xz-=Wl.x*Ul.z; A.yy-=Wl.y*Ul.y; A.yz-=Wl.y*Ul.z; A.zz-=Wl.z*Ul.z;
        C.xx-=Wa.x*Ua.x; C.xy-=Wa.x*Ua.y; C.xz-=Wa.x*Ua.z; C.yy-=Wa.y*Ua.y; C.yz-=Wa.y*Ua.z; C.zz-=Wa.z*Ua.z;
        B.xx-=Wl.x*Ua.x; B.xy-=Wl.x*Ua.y; B.xz-=Wl.x*Ua.z; B.yx-=Wl.y*Ua.x; B.yy-=Wl.y*Ua.y; B.yz-=Wl.y*Ua.z; B.zx-=Wl.z*Ua.x; B.zy-=Wl.z*Ua.y; B.zz-=Wl.z*Ua.z;
        pf = pf + u*Wl; pn = pn + u*Wa;
      } break;
      case J_PRISM: {
        double jm; const double tau = joint_torque(m, L, i, ref, jm);
        const V3 Ul = v3(A.xz, A.yz, A.zz), Ua = v3(B.zx, B.zy, B.zz);
        const double Dinv = 1.0/(A.zz + jm), u = tau - pf.z;
        st3(sl, Ul); st3(sl+3, Ua); c.S(sl+8) = Dinv; c.S(sl+9) = u;
        const V3 Wl = Dinv*Ul, Wa = Dinv*Ua;
        A.xx-=Wl.x*Ul.x; A.xy-=Wl.x*Ul.y; A.xz-=Wl.x*Ul.z; A.yy-=Wl.y*Ul.y; A.yz-=Wl.y*Ul.z; A.zz-=Wl.z*Ul.z;
        C.xx-=Wa.x*Ua.x; C.xy-=Wa.x*Ua.y; C.xz-=Wa.x*Ua.z; C.yy-=Wa.y*Ua.y; C.yz-=Wa.y*Ua.z; C.zz-=Wa.z*Ua.z;
        B.xx-=Wl.x*Ua.x; B.xy-=Wl.x*Ua.y; B.xz-=Wl.x*Ua.z; B.yx-=Wl.y*Ua.x; B.yy-=Wl.y*Ua.y; B.yz-=Wl.y*Ua.z; B.zx-=Wl.z*Ua.x; B.zy-=Wl.z*Ua.y; B.zz-=Wl.z*Ua.z;
        pf = pf + u*Wl; pn = pn + u*Wa;
      } break;
      case J_SPHER: {
        /* S = [0; E], E = R^T Ro (= RJ^T); U = [B E; C E]; D = E^T C E; tau = 0 */
        const M3 E = tmm(R, org_R(L));
        const M3 Ul = mm(B, E); M3 Cf; Cf.xx=C.xx; Cf.xy=C.xy; Cf.xz=C.xz; Cf.yx=C.xy; Cf.yy=C.yy; Cf.yz=C.yz; Cf.zx=C.xz; Cf.zy=C.yz; Cf.zz=C.zz;
        const M3 Ua = mm(Cf, E);
        const M3 Df = tmm(E, Ua); S3 D; D.xx=Df.xx; D.xy=Df.xy; D.xz=Df.xz; D.yy=Df.yy; D.yz=Df.yz; D.zz=Df.zz;
        const S3 Di = sym3_inverse(D);
        const V3 u = -tmul(E, pn);
        stm(sl, Ul); stm(sl+9, Ua);
        c.S(sl+18)=Di.xx; c.S(sl+19)=Di.xy; c.S(sl+20)=Di.xz; c.S(sl+21)=Di.yy; c.S(sl+22)=Di.yz; c.S(sl+23)=Di.zz;
        st3(sl+24, u);
        M3 Dif; Dif.xx=Di.xx; Dif.xy=Di.xy; Dif.xz=Di.xz; Dif.yx=Di.xy; Dif.yy=Di.yy; Dif.yz=Di.yz; Dif.zx=Di.xz; Dif.zy=Di.yz; Dif.zz=Di.zz;
        const M3 Wl = mm(Ul, Dif), Wa = mm(Ua, Dif);
        const M3 dA = mm(Wl, transpose(Ul)), dB = mm(Wl, transpose(Ua)), dC = mm(Wa, transpose(Ua));
        A.xx-=dA.xx; A.xy-=dA.xy; A.xz-=dA.xz; A.yy-=dA.yy; A.yz-=dA.yz; A.zz-=dA.zz;
        C.xx-=dC.xx; C.xy-=dC.xy; C.xz-=dC.xz; C.yy-=dC.yy; C.yz-=dC.yz; C.zz-=dC.zz;
        B.xx-=dB.xx; B.xy-=dB.xy; B.xz-=dB.xz; B.yx-=dB.yx; B.yy-=dB.yy; B.yz-=dB.yz; B.zx-=dB.zx; B.zy-=dB.zy; B.zz-=dB.zz;
        pf = pf + mul(Wl, u); pn = pn + mul(Wa, u);
      } break;
      case J_FLOAT: {
        /* free 6-DoF joint: a = -IA^-1 pA, nothing is transmitted to the parent */
        double a[36];
        a[0]=A.xx; a[1]=A.xy; a[2]=A.xz; a[6]=A.xy; a[7]=A.yy; a[8]=A.yz; a[12]=A.xz; a[13]=A.yz; a[14]=A.zz;
        a[3]=B.xx; a[4]=B.xy; a[5]=B.xz; a[9]=B.yx; a[10]=B.yy; a[11]=B.yz; a[15]=B.zx; a[16]=B.zy; a[17]=B.zz;
        a[18]=B.xx; a[19]=B.yx; a[20]=B.zx; a[24]=B.xy; a[25]=B.yy; a[26]=B.zy; a[30]=B.xz; a[31]=B.yz; a[32]=B.zz;
        a[21]=C.xx; a[22]=C.xy; a[23]=C.xz; a[27]=C.xy; a[28]=C.yy; a[29]=C.yz; a[33]=C.xz; a[34]=C.yz; a[35]=C.zz;
        spd6_inverse(a);
        const double b[6] = {pf.x,pf.y,pf.z,pn.x,pn.y,pn.z};
#pragma unroll
        for(int r=0;r<6;r++){ double s = 0;
#pragma unroll
          for(int k=0;k<6;k++) s -= a[6*r+k]*b[k];
          c.S(sl+r) = s; }
        if( m.has_rigid ){ int k = 0;
#pragma unroll
          for(int r=0;r<6;r++)
#pragma unroll
            for(int q=0;q<6;q++) if(q>=r){ c.S(sl+18+k) = a[6*r+q]; k++; } }
      } break;
      default: break;
      }
      if( L.parent < 0 || L.jtype == J_FLOAT ) continue;
      /* X^T Ia X and X^T pa into the parent frame */
      const S3 Ar = rot_sym(R, A), Cr = rot_sym(R, C); const M3 Br = rot_gen(R, B);
      const M3 T = mul_skew(Ar, p);                       /* A' [p x] */
      M3 Bp; Bp.xx=Br.xx-T.xx; Bp.xy=Br.xy-T.xy; Bp.xz=Br.xz-T.xz; Bp.yx=Br.yx-T.yx; Bp.yy=Br.yy-T.yy; Bp.yz=Br.yz-T.yz; Bp.zx=Br.zx-T.zx; Bp.zy=Br.zy-T.zy; Bp.zz=Br.zz-T.zz;
      const M3 Z1 = skew_mul(p, Bp), Z2 = skew_mul(p, Br);  /* C_p = C' + [p x] B_p + ([p x] B')^T */
      S3 Cp;
      Cp.xx = Cr.xx + Z1.xx + Z2.xx; Cp.xy = Cr.xy + Z1.xy + Z2.yx; Cp.xz = Cr.xz + Z1.xz + Z2.zx;
      Cp.yy = Cr.yy + Z1.yy + Z2.yy; Cp.yz = Cr.yz + Z1.yz + Z2.zy; Cp.zz = Cr.zz + Z1.zz + Z2.zz;
      const V3 fp = mul(R, pf); const V3 np = mul(R, pn) + cross(p, fp);
      if( L.serial ){ kA = Ar; kB = Bp; kC = Cp; kf = fp; kn = np; }
      else {
        const int a = m.link[L.parent].accum_slot;
        c.S(a)+=Ar.xx; c.S(a+1)+=Ar.xy; c.S(a+2)+=Ar.xz; c.S(a+3)+=Ar.yy; c.S(a+4)+=Ar.yz; c.S(a+5)+=Ar.zz;
        c.S(a+6)+=Bp.xx; c.S(a+7)+=Bp.xy; c.S(a+8)+=Bp.xz; c.S(a+9)+=Bp.yx; c.S(a+10)+=Bp.yy; c.S(a+11)+=Bp.yz; c.S(a+12)+=Bp.zx; c.S(a+13)+=Bp.zy; c.S(a+14)+=Bp.zz;
        c.S(a+15)+=Cp.xx; c.S(a+16)+=Cp.xy; c.S(a+17)+=Cp.xz; c.S(a+18)+=Cp.yy; c.S(a+19)+=Cp.yz; c.S(a+20)+=Cp.zz;
        c.S(a+21)+=fp.x; c.S(a+22)+=fp.y; c.S(a+23)+=fp.z; c.S(a+24)+=np.x; c.S(a+25)+=np.y; c.S(a+26)+=np.z;
      }
    }
  }

  /* Runge-Kutta-Gill bookkeeping of one scalar state pair (x, x') with slope (kq, kv) ([EXT A-9]):
   * stage states are built from the committed state by successive increments, the combination is
   * accumulated in the output buffer */
  struct RK { double c21, c31, c32, c42, c43, b1, b2, b3, b4; };
  RKFD_HD RK rk_coef(double dt){
    const double r2 = sqrt(2.0); RK k;
    k.c21 = 0.5*dt; k.c31 = ((r2-1.0)/2.0)*dt; k.c32 = (1.0-1.0/r2)*dt; k.c42 = (-1.0/r2)*dt; k.c43 = (1.0+1.0/r2)*dt;
    k.b1 = (1.0/6.0)*dt; k.b2 = ((2.0-r2)/6.0)*dt; k.b3 = ((2.0+r2)/6.0)*dt; k.b4 = (1.0/6.0)*dt;
    return k;
  }
  /* velocity-like (vector-space) component j */
  RKFD_HD void rk_lin(const ModelDev &m, const RK &k, int stage, int slotS, int slotP, double *gin, double *gout, int j, double slope){
    switch(stage){
    case ST_K1: { const double x0 = c.S(slotS); c.gst(gout, j, x0 + k.b1*slope); c.S(slotP) = x0 + k.c31*slope; c.S(slotS) = x0 + k.c21*slope; } break;
    case ST_K2: { const double x0 = c.gld(gin, j); c.gst(gout, j, c.gld(gout, j) + k.b2*slope); c.S(slotS) = c.S(slotP) + k.c32*slope; c.S(slotP) = x0 + k.c42*slope; } break;
    case ST_K3: { c.gst(gout, j, c.gld(gout, j) + k.b3*slope); c.S(slotS) = c.S(slotP) + k.c43*slope; } break;
    case ST_K4: { const double x = c.gld(gout, j) + k.b4*slope; c.gst(gout, j, x); c.S(slotS) = x; } break;
    default: break;
    }
  }
  /* rotation (angle-axis) component triple starting at j: increments compose on SO(3) */
  RKFD_HD void rk_rot(const ModelDev &m, const RK &k, int stage, int slotS, int slotP, double *gin, double *gout, int j, V3 w){
    switch(stage){
    case ST_K1: { const V3 x0 = ld3(slotS); const V3 F = aa_cascade(x0, k.b1*w);
      c.gst(gout,j,F.x); c.gst(gout,j+1,F.y); c.gst(gout,j+2,F.z);
      st3(slotP, aa_cascade(x0, k.c31*w)); st3(slotS, aa_cascade(x0, k.c21*w)); } break;
    case ST_K2: { const V3 x0 = v3(c.gld(gin,j), c.gld(gin,j+1), c.gld(gin,j+2));
      const V3 F = aa_cascade(v3(c.gld(gout,j), c.gld(gout,j+1), c.gld(gout,j+2)), k.b2*w);
      c.gst(gout,j,F.x); c.gst(gout,j+1,F.y); c.gst(gout,j+2,F.z);
      st3(slotS, aa_cascade(ld3(slotP), k.c32*w)); st3(slotP, aa_cascade(x0, k.c42*w)); } break;
    case ST_K3: { const V3 F = aa_cascade(v3(c.gld(gout,j), c.gld(gout,j+1), c.gld(gout,j+2)), k.b3*w);
      c.gst(gout,j,F.x); c.gst(gout,j+1,F.y); c.gst(gout,j+2,F.z);
      st3(slotS, aa_cascade(ld3(slotP), k.c43*w)); } break;
    case ST_K4: { const V3 F = aa_cascade(v3(c.gld(gout,j), c.gld(gout,j+1), c.gld(gout,j+2)), k.b4*w);
      c.gst(gout,j,F.x); c.gst(gout,j+1,F.y); c.gst(gout,j+2,F.z); st3(slotS, F); } break;
    default: break;
    }
  }
  /* one dof: displacement uses the stage velocity as slope, velocity uses the acceleration */
  RKFD_HD void rk_dof(const ModelDev &m, const RK &k, int stage, int j, double acc){
    const int qs = m.rk_slot + j, qds = m.rk_slot + m.nq + j, pq = m.rk_slot + 2*m.nq + j, pqd = m.rk_slot + 3*m.nq + j;
    if( stage >= ST_REF ){ c.gst(c.st.qdd, j, acc); if( !(fabs(acc) < 1.0e300) ) bad = 1; return; }
    const double vel = c.S(qds);
    rk_lin(m, k, stage, qs, pq, c.st.q[c.cur], c.st.q[c.cur^1], j, vel);
    rk_lin(m, k, stage, qds, pqd, c.st.qd[c.cur], c.st.qd[c.cur^1], j, acc);
  }

  /* ---- pass 3: outward acceleration pass + integrator bookkeeping */
  RKFD_HD void pass3(const ModelDev &m, int stage){
    const RK k = rk_coef(m.dt);
    V3 al = v3(0,0,0), aa = v3(0,0,0), om = v3(0,0,0);
    for(int i=0;i<m.nl;i++){
      const LinkDev &L = m.link[i]; const int sl = L.slot;
      if( !L.serial ){
        if( L.parent < 0 ){ al = v3(0,0,0); aa = v3(0,0,0); om = v3(0,0,0); }
        else { const int b = m.link[L.parent].branch_slot; al = ld3(b); aa = ld3(b+3); om = ld3(b+6); }
      }
      M3 R; V3 p, vJ, wJ;
      joint_xform(m, L, R, p, vJ, wJ);
      const V3 omp = tmul(R, om);
      const V3 zl = cross(omp, cross(omp, tmul(R, p))) + 2.0*cross(omp, vJ);
      const V3 za = cross(omp, wJ);
      const V3 xl = tmul(R, al + cross(aa, p)), xa = tmul(R, aa);
      switch(L.jtype){
      case J_REVOL: case J_PRISM: {
        const V3 Ul = ld3(sl), Ua = ld3(sl+3);
        const double acc = c.S(sl+8)*( c.S(sl+9) - (dot(Ul,xl) + dot(Ua,xa)) );
        al = xl + zl; aa = xa + za;
        if( L.jtype == J_REVOL ) aa.z += acc; else al.z += acc;
        rk_dof(m, k, stage, L.qofs, acc);
      } break;
      case J_SPHER: {
        const M3 Ul = ldm(sl), Ua = ldm(sl+9); const S3 Di = lds(sl+18); const V3 u = ld3(sl+24);
        const V3 rhs = u - (tmul(Ul, xl) + tmul(Ua, xa));
        const V3 acc = mul(Di, rhs);
        const M3 E = tmm(R, org_R(L));
        al = xl + zl; aa = xa + za + mul(E, acc);
        if( stage >= ST_REF ){ c.gst(c.st.qdd,L.qofs,acc.x); c.gst(c.st.qdd,L.qofs+1,acc.y); c.gst(c.st.qdd,L.qofs+2,acc.z);
          if( !(fabs(acc.x)+fabs(acc.y)+fabs(acc.z) < 1.0e300) ) bad = 1; }
        else {
          const int qs = m.rk_slot + L.qofs, qds = qs + m.nq, pq = qs + 2*m.nq, pqd = qs + 3*m.nq;
          const V3 w = ld3(qds);
          rk_rot(m, k, stage, qs, pq, c.st.q[c.cur], c.st.q[c.cur^1], L.qofs, w);
          rk_lin(m, k, stage, qds,   pqd,   c.st.qd[c.cur], c.st.qd[c.cur^1], L.qofs,   acc.x);
          rk_lin(m, k, stage, qds+1, pqd+1, c.st.qd[c.cur], c.st.qd[c.cur^1], L.qofs+1, acc.y);
          rk_lin(m, k, stage, qds+2, pqd+2, c.st.qd[c.cur], c.st.qd[c.cur^1], L.qofs+2, acc.z);
        }
      } break;
      case J_FLOAT: {
        const V3 a0l = ld3(sl), a0a = ld3(sl+3);
        const M3 RJ = mm(transpose(org_R(L)), R);       /* S^-1 = blockdiag(RJ, RJ) */
        const V3 accl = mul(RJ, a0l - xl - zl), acca = mul(RJ, a0a - xa - za);
        al = a0l; aa = a0a;
        if( stage >= ST_REF ){
          c.gst(c.st.qdd,L.qofs,accl.x); c.gst(c.st.qdd,L.qofs+1,accl.y); c.gst(c.st.qdd,L.qofs+2,accl.z);
          c.gst(c.st.qdd,L.qofs+3,acca.x); c.gst(c.st.qdd,L.qofs+4,acca.y); c.gst(c.st.qdd,L.qofs+5,acca.z);
          if( !(fabs(accl.x)+fabs(accl.y)+fabs(accl.z)+fabs(acca.x)+fabs(acca.y)+fabs(acca.z) < 1.0e300) ) bad = 1;
        } else {
          const int qs = m.rk_slot + L.qofs, qds = qs + m.nq, pq = qs + 2*m.nq, pqd = qs + 3*m.nq;
          const V3 v = ld3(qds), w = ld3(qds+3);
          rk_lin(m, k, stage, qs,   pq,   c.st.q[c.cur], c.st.q[c.cur^1], L.qofs,   v.x);
          rk_lin(m, k, stage, qs+1, pq+1, c.st.q[c.cur], c.st.q[c.cur^1], L.qofs+1, v.y);
          rk_lin(m, k, stage, qs+2, pq+2, c.st.q[c.cur], c.st.q[c.cur^1], L.qofs+2, v.z);
          rk_rot(m, k, stage, qs+3, pq+3, c.st.q[c.cur], c.st.q[c.cur^1], L.qofs+3, w);
          rk_lin(m, k, stage, qds,   pqd,   c.st.qd[c.cur], c.st.qd[c.cur^1], L.qofs,   accl.x);
          rk_lin(m, k, stage, qds+1, pqd+1, c.st.qd[c.cur], c.st.qd[c.cur^1], L.qofs+1, accl.y);
          rk_lin(m, k, stage, qds+2, pqd+2, c.st.qd[c.cur], c.st.qd[c.cur^1], L.qofs+2, accl.z);
          rk_lin(m, k, stage, qds+3, pqd+3, c.st.qd[c.cur], c.st.qd[c.cur^1], L.qofs+3, acca.x);
          rk_lin(m, k, stage, qds+4, pqd+4, c.st.qd[c.cur], c.st.qd[c.cur^1], L.qofs+4, acca.y);
          rk_lin(m, k, stage, qds+5, pqd+5, c.st.qd[c.cur], c.st.qd[c.cur^1], L.qofs+5, acca.z);
        }
      } break;
      default: al = xl + zl; aa = xa + za; break;
      }
      om = omp + wJ;
      if( L.branch_slot >= 0 ){ st3(L.branch_slot, al); st3(L.branch_slot+3, aa); st3(L.branch_slot+6, om); }
    }
  }

  RKFD_HD void load_flags(){ piv = c.st.piv_type[c.e]; cfl = c.st.cflags[c.e]; }
  RKFD_HD void store_flags(){ c.st.piv_type[c.e] = piv; c.st.cflags[c.e] = cfl; if( bad ) c.st.status[c.e] |= 1; }
  /* committed state (buffer `cur`) -> stage state */
  RKFD_HD void load_stage_state(const ModelDev &m){
    for(int j=0;j<m.nq;j++){ c.S(m.rk_slot+j) = c.gld(c.st.q[c.cur], j); c.S(m.rk_slot+m.nq+j) = c.gld(c.st.qd[c.cur], j); }
  }
  RKFD_HD void evaluate(const ModelDev &m, int stage){
    const bool ref = (stage == ST_REF) || (stage == ST_EVAL_REF);
    pass1(m, ref); pass2(m, ref); pass3(m, stage);
  }
  /* mode 0: rkFDUpdate x nsteps (reference rkfd_sim.c:560-566); mode 1 / 2: a single non-committing /
   * committing evaluation on the committed state (2 = rkFDUpdateInit's t=0 evaluation).  One stage loop so
   * that the evaluation body is instantiated once. */
  RKFD_HD void run(const ModelDev &m, int mode, int nsteps){
    load_flags();
    const int first = mode == 0 ? ST_K1 : ( mode == 2 ? ST_EVAL_REF : ST_EVAL );
    const int last  = mode == 0 ? ST_REF : first;
    if( mode != 0 ) nsteps = 1;
#pragma unroll 1
    for(int s=0;s<nsteps;s++){
      load_stage_state(m);
#pragma unroll 1
      for(int stage=first; stage<=last; stage++){
        if( stage == ST_REF ) c.cur ^= 1;   /* the output buffer now holds the committed state */
        evaluate(m, stage);
      }
    }
    store_flags();
  }
};

}  // namespace rkfd
#endif
