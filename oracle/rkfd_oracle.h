/* rkfd_oracle.h - CPU oracle for the RoKi-FD step path (TEST INFRASTRUCTURE, NOT PRODUCT).
 *
 * This is a plain-C, fp64, single-environment restatement of the algorithm that
 * mi-lib/roki-fd v1.7.9 runs inside rkFDUpdate() (reference src/rkfd_sim.c:560-566).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it; the product library (librokifd_b200.so) never links or calls it.
 *
 * PARITY UNPINNED: the reference ships no golden vectors / known-answer tests
 * (reference test/test.sh:5-12 globs *test.c, none exist) and cannot be built here
 * (ZEDA/ZM/Zeo/RoKi are un-vendored dependencies, reference libinfo:3).  The in-tree
 * parts (rkfd_sim.c, rkfd_util.c, rkfd_penalty.c, rkfd_cd.c, rkfd_mlcp.c, rkfd_vert.c,
 * rkfd_opt_qp.c) are restated line by line; the parts that live in RoKi/ZM/Zeo
 * (ABA, FK, DC motor, vertex collision, RKG, zLESolveMP) follow their published
 * algorithms with the conventions fixed in DESIGN.md ("EXT assumptions").
 */
#ifndef RKFD_ORACLE_H
#define RKFD_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum { ORK_JOINT_FIXED = 0, ORK_JOINT_REVOL = 1, ORK_JOINT_PRISM = 2,
       ORK_JOINT_SPHER = 3, ORK_JOINT_FLOAT = 4,
       ORK_JOINT_CYLIN = 5,   /* [EXT] cylindrical: dis = (translation along z, rotation about z) */
       ORK_JOINT_HOOKE = 6,
       ORK_JOINT_BRFLOAT = 7 };  /* [EXT A-17] breakable float (wall.ztk:51-95): rigid until the wrench it transmits exceeds its force or torque
                                  * threshold in a committing evaluation, a float joint from then on */ /* [EXT] hooke / universal: dis = (rotation about z, then about the new y): R = Rz(q0) Ry(q1) */
enum { ORK_MOTOR_NONE = 0, ORK_MOTOR_DC = 1, ORK_MOTOR_TRQ = 2 };
enum { ORK_CONTACT_RIGID = 0, ORK_CONTACT_ELASTIC = 1 };
enum { ORK_SF = 0, ORK_KF = 1 };                     /* static / kinetic friction state */
enum { ORK_SOLVER_VERT = 0, ORK_SOLVER_MLCP = 1, ORK_SOLVER_VOLUME = 2 };

#define ORK_G 9.80665          /* RoKi RK_G */
#define ORK_TOL 1.0e-12        /* ZM zTOL   */

/* number of doubles per link in the flat description passed to ork_world_new() */
#define ORK_LINK_ND 37
/* link_d layout (per link):
 *  [0..8]  org_R row-major (link frame w.r.t. parent frame, ZTK "frame:" rotation part)
 *  [9..11] org_p
 *  [12]    mass   [13..15] COM (link frame)   [16..24] inertia about COM, row-major
 *  [25] stiffness [26] viscosity [27] coulomb [28] staticfriction      (1-DoF joints)
 *  [29] motor constant k [30] admittance [31] gear ratio [32] rotor inertia
 *  [33] gear inertia [34] min input [35] max input   [36] reserved
 * link_i layout (per link, 4 ints): parent, jtype, mtype, stuff id
 */
#define ORK_LINK_NI 4

typedef struct ork_world ork_world;   /* model + properties (shared by all envs) */
typedef struct ork_env ork_env;       /* one environment's state */

/* ---- world construction ------------------------------------------------------- */
ork_world *ork_world_new(int nl, const int *link_i, const double *link_d);
void ork_world_free(ork_world *w);
/* a convex vertex cloud fixed to moving link `link` (RoKi: one rkCDCell of type MOVE) */
int ork_world_add_cell(ork_world *w, int link, int nvert, const double *verts /*3*nvert, link frame*/);
/* a static box (RoKi: rkCDCell of type STAT): world pose R (row-major), p = centre; half extents */
int ork_world_add_box(ork_world *w, const double *R, const double *p, const double *half, int stuff);
/* contact-info table (reference rkfd_sim.c:259-273); key = unordered stuff pair */
void ork_world_add_contact_info(ork_world *w, int stuff_a, int stuff_b, int type,
                                double K, double L, double E, double V, double SF, double KF);
/* properties (reference rkfd_property.c:10-18) and solver choice (rkfd_sim.h:89-93);
 * choosing the solver also installs its default contact info (rkfd_vert.c:340-348 ...) */
void ork_world_set_prp(ork_world *w, double dt, int pyramid, double friction_weight, int max_iter);
void ork_world_set_integrator(ork_world *w, int integrator);    /* 0 RKG (default), 1 RK4, 2 Euler, 3 Heun */
void ork_world_set_solver(ork_world *w, int solver);
/* a box carried by a moving link (frame in the link): a collision target for the vertices of cells on other links */
int ork_world_add_link_box(ork_world *w, int link, const double *R, const double *p, const double *half);
/* [EXT] rkCDPairChainUnreg: drops the pairs between cells of the chain `link` belongs to (registered by default) */
void ork_world_unreg_self_collision(ork_world *w, int link);
/* finish: builds (cell x box) pairs in registration order and associates contact info */
void ork_world_finalize(ork_world *w);
/* force / torque threshold of a breakable-float link */
void ork_world_set_break(ork_world *w, int link, double fth, double tth);
/* per link: 1 when its breakable-float joint has broken */
void ork_env_get_broken(const ork_env *e, int *broken);
int ork_world_nq(const ork_world *w);       /* total joint size */
int ork_world_nslot(const ork_world *w);    /* contact slots = sum over pairs of nvert */
int ork_world_nl(const ork_world *w);

/* ---- environment -------------------------------------------------------------- */
ork_env *ork_env_new(const ork_world *w);
void ork_env_free(ork_env *e);
void ork_env_set_state(ork_env *e, const double *q, const double *qd);
void ork_env_get_state(const ork_env *e, double *q, double *qd, double *qdd);
void ork_env_set_motor_input(ork_env *e, const double *u /* per link */);
/* persistent friction/contact state, for teacher-forced parity */
void ork_env_get_pivot(const ork_env *e, int *type, double *prev_trq);
void ork_env_set_pivot(ork_env *e, const int *type, const double *prev_trq);
void ork_env_get_contact(const ork_env *e, int *active, int *type, double *ref, double *f);
void ork_env_set_contact(ork_env *e, const int *active, const int *type, const double *ref);
double ork_env_time(const ork_env *e);

/* rkFDUpdateInit: committing evaluation at t=0 (reference rkfd_sim.c:552-558) */
void ork_env_update_init(ork_env *e);
/* rkFDUpdate: one step (reference rkfd_sim.c:560-566) */
void ork_env_update(ork_env *e);
/* one dynamics evaluation on the CURRENT state: _rkFDUpdateRef if do_up_ref, else the
 * body of _rkFDUpdate (reference rkfd_sim.c:533-549).  Writes qdd. */
void ork_env_eval(ork_env *e, int do_up_ref);

/* diagnostics for tests: world frames (12 per link: R row-major, p), link velocities (6 per
 * link: lin, ang in link frame), link accelerations (6 per link) of the last evaluation */
void ork_env_get_link_frames(const ork_env *e, double *frames);
void ork_env_get_link_vel(const ork_env *e, double *vel);
void ork_env_get_link_acc(const ork_env *e, double *acc);
/* total mechanical energy (kinetic incl. motor rotor inertia, gravity potential) */
double ork_env_energy(const ork_env *e);
/* last rigid-contact system (Delassus matrix A (n x n, row-major), bias b, solution f); returns n */
int ork_env_get_rigid_system(const ork_env *e, double *A, double *b, double *f, int cap);

/* test hook: the Vert QP of the last evaluation and how its active-set loop ended (see rkfd_oracle.c) */
int ork_env_get_qp(const ork_env *e, double *Q, double *c, double *nf, double *x, int *idx, int *info, int capn, int capm);
int ork_qp_solve_asm_ex(int n, int m, const double *q, const double *c, const double *a,
                        const double *b, const double *init, double *ans, int *idx, int *term);

/* Volume solver (rkfd_volume.c) results of the last evaluation, per pair: np = contact-polygon planes (-1: no contact
 * volume), type = pair friction type, wrench (6, world axes, at center), center (3) */
void ork_env_get_volume(const ork_env *e, int *np, int *type, double *wrench, double *center);
int ork_world_npair(const ork_world *w);
/* test hook: per pair Q6 (36, row-major), c6 (6), norm (3) of the last Volume evaluation (rkfd_volume.c:397-491) */
void ork_env_get_volume_constraint(const ork_env *e, double *qc);
/* [EXT A-16] LP by the two-phase simplex (Bland): min c^T x s.t. A x = b (m x n, row-major), x >= 0; c NULL: feasibility */
int ork_lp_solve(int m, int n, const double *A, const double *b, const double *c, double *x);

/* ---- batch driver (CPU baseline): B independent envs, OpenMP over envs ---------- */
/* q, qd, u: env-major (B x nq / B x nl); steps every env `nsteps` times; returns threads used */
int ork_batch_run(const ork_world *w, int B, double *q, double *qd, const double *u,
                  int nsteps, int nthreads, double *qdd_out);

int ork_batch_run_state(const ork_world *w, int B, double *q, double *qd, const double *u,
                        int nsteps, int nthreads, double *qdd_out, int *act, int *type, double *cf, int *piv);

/* ---- standalone pieces exported for unit tests ---------------------------------- */
/* rkFDQPSolveASM restatement (reference rkfd_opt_qp.c:43-181) with the default cond
 * (row . x).  q: n x n, c: n, a: m x n, b: m, ans: n (out), idx: m (out).  Returns iterations. */
int ork_qp_solve_asm(int n, int m, const double *q, const double *c, const double *a,
                     const double *b, const double *init, double *ans, int *idx);
/* zLESolveMP restatement: minimum-norm least-squares solution of a (n x n, symmetric) x = b */
void ork_le_solve_mp_sym(int n, const double *a, const double *b, double *x);

#ifdef __cplusplus
}
#endif
#endif
