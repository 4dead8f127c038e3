// mjoracle.cpp — CPU ORACLE (test infrastructure, NOT product code).
//
// Scalar fp64, one environment at a time, straight-line restatement of the forward-dynamics step
// that dm_control reaches through `mujoco.mj_step / mj_step1 / mj_step2 / mj_forward`
// (call sites: dm_control/mujoco/engine.py:156-176 (step ordering), :306-343 (reset/forward)).
//
// PARITY UNPINNED: the arithmetic of that path lives in the third-party PyPI package `mujoco`
// (pinned ==3.11.0, /root/reference/requirements.txt:9), which is absent from /root/reference and from
// this image. The stages below restate MuJoCo's published "Computation" chapter + SURVEY.md §8a /
// Appendix C from memory; they are anchored only on the reference's own analytic known-answers
// (tests/test_oracle_kat.py: mujoco/README.md:25-50 rest depth 0.19996362, wrapper/core_test.py:393-416
// weight = sum of contact normal forces, suite/lqr_solver.py:44-66 closed-form Euler update, ...).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load this.
//
// Build: see oracle/Makefile  ->  oracle/_build/libmjoracle.so

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>

#include "../include/b200mj_model_fields.h"
#include "../include/b200mj_convex.h"   // MPR / capsule-box: the one narrow-phase source shared with the engine (see its header)

typedef double real;
typedef std::vector<real> vec;
typedef std::vector<int> ivec;

// ------------------------------------------------------------------------------------------------
// model
// ------------------------------------------------------------------------------------------------
struct OModel {
#define DECL_I(name) ivec name;
#define DECL_R(name) vec name;
  B200MJ_MODEL_FIELDS(DECL_I, DECL_R)
#undef DECL_I
#undef DECL_R
  int nq, nv, nu, na, nbody, njnt, ngeom, nsite, ntendon, neq, nsensor, nsensordata, npair, nlevel, nconmax, njmax;
  real timestep, gravity[3], tolerance, ls_tolerance, impratio, meaninertia;
  int integrator, solver, iterations, ls_iterations, disableflags;
};

static OModel* model_from_blob(const int* idata, const double* rdata) {
  OModel* m = new OModel();
  int k = 0;
#define LOAD_I(name) { int off = idata[2*k], len = idata[2*k+1]; m->name.assign(idata+off, idata+off+len); k++; }
#define LOAD_R(name) { int off = idata[2*k], len = idata[2*k+1]; m->name.assign(rdata+off, rdata+off+len); k++; }
  B200MJ_MODEL_FIELDS(LOAD_I, LOAD_R)
#undef LOAD_I
#undef LOAD_R
  const ivec& s = m->sizes;
  m->nq = s[BMJ_NQ]; m->nv = s[BMJ_NV]; m->nu = s[BMJ_NU]; m->na = s[BMJ_NA]; m->nbody = s[BMJ_NBODY];
  m->njnt = s[BMJ_NJNT]; m->ngeom = s[BMJ_NGEOM]; m->nsite = s[BMJ_NSITE]; m->ntendon = s[BMJ_NTENDON];
  m->neq = s[BMJ_NEQ]; m->nsensor = s[BMJ_NSENSOR]; m->nsensordata = s[BMJ_NSENSORDATA];
  m->npair = s[BMJ_NPAIR]; m->nlevel = s[BMJ_NLEVEL]; m->nconmax = s[BMJ_NCONMAX]; m->njmax = s[BMJ_NJMAX];
  m->timestep = m->opt_real[BMJ_OPT_TIMESTEP];
  for (int i = 0; i < 3; i++) m->gravity[i] = m->opt_real[BMJ_OPT_GRAVITY_X + i];
  m->tolerance = m->opt_real[BMJ_OPT_TOLERANCE]; m->ls_tolerance = m->opt_real[BMJ_OPT_LS_TOLERANCE];
  m->impratio = m->opt_real[BMJ_OPT_IMPRATIO]; m->meaninertia = m->opt_real[BMJ_OPT_MEANINERTIA];
  m->integrator = m->opt_int[BMJ_OPT_INTEGRATOR]; m->solver = m->opt_int[BMJ_OPT_SOLVER];
  m->iterations = m->opt_int[BMJ_OPT_ITERATIONS]; m->ls_iterations = m->opt_int[BMJ_OPT_LS_ITERATIONS];
  m->disableflags = m->opt_int[BMJ_OPT_DISABLEFLAGS];
  return m;
}

// ------------------------------------------------------------------------------------------------
// data
// ------------------------------------------------------------------------------------------------
struct OContact {
  real dist, pos[3], frame[9], includemargin, friction[5], solref[2], solimp[5], mu;
  int dim, geom1, geom2, efc_address;
};

struct OData {
  real time;
  vec qpos, qvel, act, ctrl, qacc, qacc_warmstart, act_dot, qfrc_applied, xfrc_applied;
  vec xpos, xquat, xmat, xipos, ximat, xanchor, xaxis, geom_xpos, geom_xmat, site_xpos, site_xmat;
  vec subtree_com, cinert, crb, cdof, cdof_dot, cvel, cacc, cfrc_int, cfrc_ext, subtree_linvel;
  vec M, L, Lh;  // dense nv*nv: inertia, chol(M), chol(M + h*B)
  vec ten_length, ten_J, ten_velocity, actuator_length, actuator_velocity, actuator_moment, actuator_force;
  vec qfrc_bias, qfrc_passive, qfrc_actuator, qfrc_smooth, qacc_smooth, qfrc_constraint;
  std::vector<OContact> contact;
  int ncon;
  int nefc;
  vec efc_J, efc_pos, efc_margin, efc_D, efc_R, efc_aref, efc_force, efc_vel, efc_diagApprox;
  ivec efc_type, efc_id, efc_state;
  vec sensordata;
  int warning[BMJ_NWARNING];
  int solver_niter;
  real energy[2];
};

static OData* data_create(const OModel* m) {
  OData* d = new OData();
  int nv = m->nv, nb = m->nbody;
  d->qpos.resize(m->nq); d->qvel.resize(nv); d->act.resize(m->na); d->ctrl.resize(m->nu);
  d->qacc.resize(nv); d->qacc_warmstart.resize(nv); d->act_dot.resize(m->na); d->qfrc_applied.resize(nv);
  d->xfrc_applied.resize(6 * nb);
  d->xpos.resize(3 * nb); d->xquat.resize(4 * nb); d->xmat.resize(9 * nb); d->xipos.resize(3 * nb);
  d->ximat.resize(9 * nb); d->xanchor.resize(3 * m->njnt); d->xaxis.resize(3 * m->njnt);
  d->geom_xpos.resize(3 * m->ngeom); d->geom_xmat.resize(9 * m->ngeom);
  d->site_xpos.resize(3 * m->nsite); d->site_xmat.resize(9 * m->nsite);
  d->subtree_com.resize(3 * nb); d->cinert.resize(10 * nb); d->crb.resize(10 * nb);
  d->cdof.resize(6 * nv); d->cdof_dot.resize(6 * nv); d->cvel.resize(6 * nb); d->cacc.resize(6 * nb);
  d->cfrc_int.resize(6 * nb); d->cfrc_ext.resize(6 * nb); d->subtree_linvel.resize(3 * nb);
  d->M.resize(nv * nv); d->L.resize(nv * nv); d->Lh.resize(nv * nv);
  d->ten_length.resize(m->ntendon); d->ten_J.resize(m->ntendon * nv); d->ten_velocity.resize(m->ntendon);
  d->actuator_length.resize(m->nu); d->actuator_velocity.resize(m->nu); d->actuator_moment.resize(m->nu * nv);
  d->actuator_force.resize(m->nu);
  d->qfrc_bias.resize(nv); d->qfrc_passive.resize(nv); d->qfrc_actuator.resize(nv); d->qfrc_smooth.resize(nv);
  d->qacc_smooth.resize(nv); d->qfrc_constraint.resize(nv);
  d->contact.resize(m->nconmax);
  int nj = m->njmax;
  d->efc_J.resize(nj * nv); d->efc_pos.resize(nj); d->efc_margin.resize(nj); d->efc_D.resize(nj); d->efc_R.resize(nj);
  d->efc_aref.resize(nj); d->efc_force.resize(nj); d->efc_vel.resize(nj); d->efc_diagApprox.resize(nj);
  d->efc_type.resize(nj); d->efc_id.resize(nj); d->efc_state.resize(nj);
  d->sensordata.resize(m->nsensordata);
  return d;
}

static void reset_data(const OModel* m, OData* d, int key) {
  std::fill(d->qvel.begin(), d->qvel.end(), 0.0);
  std::fill(d->act.begin(), d->act.end(), 0.0);
  std::fill(d->ctrl.begin(), d->ctrl.end(), 0.0);
  std::fill(d->qacc.begin(), d->qacc.end(), 0.0);
  std::fill(d->qacc_warmstart.begin(), d->qacc_warmstart.end(), 0.0);
  std::fill(d->qfrc_applied.begin(), d->qfrc_applied.end(), 0.0);
  std::fill(d->xfrc_applied.begin(), d->xfrc_applied.end(), 0.0);
  std::fill(d->sensordata.begin(), d->sensordata.end(), 0.0);
  std::fill(d->actuator_force.begin(), d->actuator_force.end(), 0.0);
  std::fill(d->efc_force.begin(), d->efc_force.end(), 0.0);
  if (key >= 0 && key < m->sizes[BMJ_NKEY]) for (int i = 0; i < m->nq; i++) d->qpos[i] = m->key_qpos[key * m->nq + i];
  else d->qpos = m->qpos0;
  d->time = 0; d->ncon = 0; d->nefc = 0; d->solver_niter = 0;
  for (int i = 0; i < BMJ_NWARNING; i++) d->warning[i] = 0;
}

// ------------------------------------------------------------------------------------------------
// 3-vector / quaternion helpers
// ------------------------------------------------------------------------------------------------
static inline real dot3(const real* a, const real* b) { return a[0]*b[0] + a[1]*b[1] + a[2]*b[2]; }
static inline void cross3(real* r, const real* a, const real* b) {
  real x = a[1]*b[2] - a[2]*b[1], y = a[2]*b[0] - a[0]*b[2], z = a[0]*b[1] - a[1]*b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline real norm3(const real* a) { return std::sqrt(dot3(a, a)); }
static inline real normalize3(real* a) {
  real n = norm3(a);
  if (n < BMJ_MINVAL) { a[0] = 1; a[1] = 0; a[2] = 0; } else { a[0] /= n; a[1] /= n; a[2] /= n; }
  return n;
}
static inline void mul_quat(real* r, const real* a, const real* b) {
  real w = a[0]*b[0] - a[1]*b[1] - a[2]*b[2] - a[3]*b[3];
  real x = a[0]*b[1] + a[1]*b[0] + a[2]*b[3] - a[3]*b[2];
  real y = a[0]*b[2] - a[1]*b[3] + a[2]*b[0] + a[3]*b[1];
  real z = a[0]*b[3] + a[1]*b[2] - a[2]*b[1] + a[3]*b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
static inline void normalize4(real* q) {
  real n = std::sqrt(q[0]*q[0] + q[1]*q[1] + q[2]*q[2] + q[3]*q[3]);
  if (n < BMJ_MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; }
  else if (std::fabs(n - 1) > BMJ_MINVAL) { real inv = 1 / n; for (int i = 0; i < 4; i++) q[i] *= inv; }   // idempotent, as mju_normalize4
}
static inline void quat2mat(real* m, const real* q) {
  real q00 = q[0]*q[0], q11 = q[1]*q[1], q22 = q[2]*q[2], q33 = q[3]*q[3];
  m[0] = q00 + q11 - q22 - q33; m[4] = q00 - q11 + q22 - q33; m[8] = q00 - q11 - q22 + q33;
  m[1] = 2*(q[1]*q[2] - q[0]*q[3]); m[2] = 2*(q[1]*q[3] + q[0]*q[2]);
  m[3] = 2*(q[1]*q[2] + q[0]*q[3]); m[5] = 2*(q[2]*q[3] - q[0]*q[1]);
  m[6] = 2*(q[1]*q[3] - q[0]*q[2]); m[7] = 2*(q[2]*q[3] + q[0]*q[1]);
}
static inline void rot_vec_quat(real* r, const real* v, const real* q) {
  real m[9]; quat2mat(m, q);
  real x = m[0]*v[0] + m[1]*v[1] + m[2]*v[2], y = m[3]*v[0] + m[4]*v[1] + m[5]*v[2], z = m[6]*v[0] + m[7]*v[1] + m[8]*v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void mul_mat_vec3(real* r, const real* m, const real* v) {
  real x = m[0]*v[0] + m[1]*v[1] + m[2]*v[2], y = m[3]*v[0] + m[4]*v[1] + m[5]*v[2], z = m[6]*v[0] + m[7]*v[1] + m[8]*v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void mul_matT_vec3(real* r, const real* m, const real* v) {
  real x = m[0]*v[0] + m[3]*v[1] + m[6]*v[2], y = m[1]*v[0] + m[4]*v[1] + m[7]*v[2], z = m[2]*v[0] + m[5]*v[1] + m[8]*v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void axis_angle2quat(real* q, const real* axis, real angle) {
  if (angle == 0) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  real s = std::sin(angle * 0.5);
  q[0] = std::cos(angle * 0.5); q[1] = axis[0]*s; q[2] = axis[1]*s; q[3] = axis[2]*s;
}
// quaternion integration: q <- q * exp(h*omega/2), omega in the local frame
static inline void quat_integrate(real* q, const real* w, real h) {
  real ax[3] = {w[0], w[1], w[2]};
  real n = norm3(ax);
  real ang = h * n;
  if (n < BMJ_MINVAL) return;
  ax[0] /= n; ax[1] /= n; ax[2] /= n;
  real dq[4], r[4];
  axis_angle2quat(dq, ax, ang);
  normalize4(q);
  mul_quat(r, q, dq);
  for (int i = 0; i < 4; i++) q[i] = r[i];
}

// spatial algebra (rot[3], lin[3])
static inline void mul_inert_vec(real* r, const real* i, const real* v) {
  r[0] = i[0]*v[0] + i[3]*v[1] + i[4]*v[2] - i[8]*v[4] + i[7]*v[5];
  r[1] = i[3]*v[0] + i[1]*v[1] + i[5]*v[2] + i[8]*v[3] - i[6]*v[5];
  r[2] = i[4]*v[0] + i[5]*v[1] + i[2]*v[2] - i[7]*v[3] + i[6]*v[4];
  r[3] = i[8]*v[1] - i[7]*v[2] + i[9]*v[3];
  r[4] = i[6]*v[2] - i[8]*v[0] + i[9]*v[4];
  r[5] = i[7]*v[0] - i[6]*v[1] + i[9]*v[5];
}
static inline void cross_motion(real* r, const real* vel, const real* v) {
  real a[3], b[3], c[3];
  cross3(a, vel, v); cross3(b, vel, v + 3); cross3(c, vel + 3, v);
  r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; r[3] = b[0] + c[0]; r[4] = b[1] + c[1]; r[5] = b[2] + c[2];
}
static inline void cross_force(real* r, const real* vel, const real* f) {
  real a[3], b[3], c[3];
  cross3(a, vel, f); cross3(b, vel + 3, f + 3); cross3(c, vel, f + 3);
  r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; r[3] = c[0]; r[4] = c[1]; r[5] = c[2];
}

// dense Cholesky (lower, row-major n*n): A = L L^T. Returns smallest pivot.
static real chol_factor(real* L, const real* A, int n) {
  real minpiv = 1e300;
  for (int i = 0; i < n * n; i++) L[i] = 0;
  for (int j = 0; j < n; j++) {
    real s = A[j * n + j];
    for (int k = 0; k < j; k++) s -= L[j * n + k] * L[j * n + k];
    if (s < BMJ_MINVAL) s = BMJ_MINVAL;
    minpiv = std::min(minpiv, s);
    real piv = std::sqrt(s);
    L[j * n + j] = piv;
    for (int i = j + 1; i < n; i++) {
      real t = A[i * n + j];
      for (int k = 0; k < j; k++) t -= L[i * n + k] * L[j * n + k];
      L[i * n + j] = t / piv;
    }
  }
  return minpiv;
}
static void chol_solve(real* x, const real* L, const real* b, int n) {
  for (int i = 0; i < n; i++) {
    real s = b[i];
    for (int k = 0; k < i; k++) s -= L[i * n + k] * x[k];
    x[i] = s / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    real s = x[i];
    for (int k = n - 1; k > i; k--) s -= L[k * n + i] * x[k];
    x[i] = s / L[i * n + i];
  }
}

// ------------------------------------------------------------------------------------------------
// position stage
// ------------------------------------------------------------------------------------------------
static void kinematics(const OModel* m, OData* d) {
  real* xpos = d->xpos.data(); real* xquat = d->xquat.data(); real* xmat = d->xmat.data();
  xpos[0] = xpos[1] = xpos[2] = 0; xquat[0] = 1; xquat[1] = xquat[2] = xquat[3] = 0;
  quat2mat(xmat, xquat);
  // normalise quaternions stored in qpos
  for (int j = 0; j < m->njnt; j++) {
    if (m->jnt_type[j] == BMJ_JNT_FREE) normalize4(&d->qpos[m->jnt_qposadr[j] + 3]);
    else if (m->jnt_type[j] == BMJ_JNT_BALL) normalize4(&d->qpos[m->jnt_qposadr[j]]);
  }
  for (int b = 1; b < m->nbody; b++) {
    int p = m->body_parentid[b];
    real pos[3], quat[4], tmp[3];
    mul_mat_vec3(tmp, xmat + 9 * p, &m->body_pos[3 * b]);
    for (int i = 0; i < 3; i++) pos[i] = xpos[3 * p + i] + tmp[i];
    mul_quat(quat, xquat + 4 * p, &m->body_quat[4 * b]);
    for (int j = m->body_jntadr[b]; j < m->body_jntadr[b] + m->body_jntnum[b]; j++) {
      int qa = m->jnt_qposadr[j];
      real* anchor = &d->xanchor[3 * j]; real* axis = &d->xaxis[3 * j];
      if (m->jnt_type[j] == BMJ_JNT_FREE) {
        for (int i = 0; i < 3; i++) pos[i] = d->qpos[qa + i];
        for (int i = 0; i < 4; i++) quat[i] = d->qpos[qa + 3 + i];
        for (int i = 0; i < 3; i++) anchor[i] = pos[i];
        rot_vec_quat(axis, &m->jnt_axis[3 * j], quat);
        continue;
      }
      rot_vec_quat(tmp, &m->jnt_pos[3 * j], quat);
      for (int i = 0; i < 3; i++) anchor[i] = pos[i] + tmp[i];
      rot_vec_quat(axis, &m->jnt_axis[3 * j], quat);
      if (m->jnt_type[j] == BMJ_JNT_SLIDE) {
        real q = d->qpos[qa] - m->qpos0[qa];
        for (int i = 0; i < 3; i++) pos[i] += axis[i] * q;
      } else {
        real ql[4], r[4];
        if (m->jnt_type[j] == BMJ_JNT_HINGE) axis_angle2quat(ql, &m->jnt_axis[3 * j], d->qpos[qa] - m->qpos0[qa]);
        else for (int i = 0; i < 4; i++) ql[i] = d->qpos[qa + i];
        mul_quat(r, quat, ql);
        for (int i = 0; i < 4; i++) quat[i] = r[i];
        rot_vec_quat(tmp, &m->jnt_pos[3 * j], quat);
        for (int i = 0; i < 3; i++) pos[i] = anchor[i] - tmp[i];
      }
    }
    normalize4(quat);
    for (int i = 0; i < 3; i++) xpos[3 * b + i] = pos[i];
    for (int i = 0; i < 4; i++) xquat[4 * b + i] = quat[i];
    quat2mat(xmat + 9 * b, quat);
  }
  for (int b = 0; b < m->nbody; b++) {
    real tmp[3], q[4];
    mul_mat_vec3(tmp, xmat + 9 * b, &m->body_ipos[3 * b]);
    for (int i = 0; i < 3; i++) d->xipos[3 * b + i] = xpos[3 * b + i] + tmp[i];
    mul_quat(q, xquat + 4 * b, &m->body_iquat[4 * b]);
    quat2mat(&d->ximat[9 * b], q);
  }
  for (int g = 0; g < m->ngeom; g++) {
    int b = m->geom_bodyid[g]; real tmp[3], q[4];
    mul_mat_vec3(tmp, xmat + 9 * b, &m->geom_pos[3 * g]);
    for (int i = 0; i < 3; i++) d->geom_xpos[3 * g + i] = xpos[3 * b + i] + tmp[i];
    mul_quat(q, xquat + 4 * b, &m->geom_quat[4 * g]);
    quat2mat(&d->geom_xmat[9 * g], q);
  }
  for (int s = 0; s < m->nsite; s++) {
    int b = m->site_bodyid[s]; real tmp[3], q[4];
    mul_mat_vec3(tmp, xmat + 9 * b, &m->site_pos[3 * s]);
    for (int i = 0; i < 3; i++) d->site_xpos[3 * s + i] = xpos[3 * b + i] + tmp[i];
    mul_quat(q, xquat + 4 * b, &m->site_quat[4 * s]);
    quat2mat(&d->site_xmat[9 * s], q);
  }
}

static void com_pos(const OModel* m, OData* d) {
  int nb = m->nbody;
  real* sc = d->subtree_com.data();
  for (int b = 0; b < nb; b++) for (int i = 0; i < 3; i++) sc[3 * b + i] = m->body_mass[b] * d->xipos[3 * b + i];
  for (int b = nb - 1; b > 0; b--) { int p = m->body_parentid[b]; for (int i = 0; i < 3; i++) sc[3 * p + i] += sc[3 * b + i]; }
  for (int b = 0; b < nb; b++) {
    if (m->body_subtreemass[b] < BMJ_MINVAL) for (int i = 0; i < 3; i++) sc[3 * b + i] = d->xipos[3 * b + i];
    else for (int i = 0; i < 3; i++) sc[3 * b + i] /= m->body_subtreemass[b];
  }
  // body inertia in the frame centred at the root subtree's COM, world orientation
  for (int i = 0; i < 10; i++) d->cinert[i] = 0;
  for (int b = 1; b < nb; b++) {
    const real* mat = &d->ximat[9 * b]; const real* in = &m->body_inertia[3 * b];
    real dif[3]; int root = m->body_rootid[b];
    for (int i = 0; i < 3; i++) dif[i] = d->xipos[3 * b + i] - sc[3 * root + i];
    real mass = m->body_mass[b];
    real t[9];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++)
      t[3 * r + c] = mat[3 * r + 0] * in[0] * mat[3 * c + 0] + mat[3 * r + 1] * in[1] * mat[3 * c + 1] + mat[3 * r + 2] * in[2] * mat[3 * c + 2];
    real* ci = &d->cinert[10 * b];
    ci[0] = t[0] + mass * (dif[1]*dif[1] + dif[2]*dif[2]);
    ci[1] = t[4] + mass * (dif[0]*dif[0] + dif[2]*dif[2]);
    ci[2] = t[8] + mass * (dif[0]*dif[0] + dif[1]*dif[1]);
    ci[3] = t[1] - mass * dif[0]*dif[1];
    ci[4] = t[2] - mass * dif[0]*dif[2];
    ci[5] = t[5] - mass * dif[1]*dif[2];
    ci[6] = mass * dif[0]; ci[7] = mass * dif[1]; ci[8] = mass * dif[2]; ci[9] = mass;
  }
  // motion axes of every dof about the same point
  for (int j = 0; j < m->njnt; j++) {
    int b = m->jnt_bodyid[j], root = m->body_rootid[b], da = m->jnt_dofadr[j];
    real off[3];
    for (int i = 0; i < 3; i++) off[i] = sc[3 * root + i] - d->xanchor[3 * j + i];
    real* cd = &d->cdof[6 * da];
    int t = m->jnt_type[j];
    if (t == BMJ_JNT_FREE) {
      for (int k = 0; k < 3; k++) { for (int i = 0; i < 6; i++) cd[6 * k + i] = 0; cd[6 * k + 3 + k] = 1; }
      cd += 18;
    }
    if (t == BMJ_JNT_FREE || t == BMJ_JNT_BALL) {
      const real* xm = &d->xmat[9 * b];
      for (int k = 0; k < 3; k++) {
        real ax[3] = {xm[k], xm[3 + k], xm[6 + k]};
        for (int i = 0; i < 3; i++) cd[6 * k + i] = ax[i];
        cross3(cd + 6 * k + 3, ax, off);
      }
    } else if (t == BMJ_JNT_SLIDE) {
      for (int i = 0; i < 3; i++) { cd[i] = 0; cd[3 + i] = d->xaxis[3 * j + i]; }
    } else {
      for (int i = 0; i < 3; i++) cd[i] = d->xaxis[3 * j + i];
      cross3(cd + 3, &d->xaxis[3 * j], off);
    }
  }
}

static void tendon_and_transmission(const OModel* m, OData* d) {
  int nv = m->nv;
  for (int t = 0; t < m->ntendon; t++) {
    real len = 0;
    for (int i = 0; i < nv; i++) d->ten_J[t * nv + i] = 0;
    for (int w = m->tendon_adr[t]; w < m->tendon_adr[t] + m->tendon_num[t]; w++) {
      int j = m->wrap_objid[w];
      len += m->wrap_prm[w] * d->qpos[m->jnt_qposadr[j]];
      d->ten_J[t * nv + m->jnt_dofadr[j]] += m->wrap_prm[w];
    }
    d->ten_length[t] = len;
  }
  for (int a = 0; a < m->nu; a++) {
    real* mom = &d->actuator_moment[a * nv];
    for (int i = 0; i < nv; i++) mom[i] = 0;
    real gear = m->actuator_gear[a];
    if (m->actuator_trntype[a] == BMJ_TRN_JOINT) {
      int j = m->actuator_trnid[a];
      d->actuator_length[a] = gear * d->qpos[m->jnt_qposadr[j]];
      mom[m->jnt_dofadr[j]] = gear;
    } else {
      int t = m->actuator_trnid[a];
      d->actuator_length[a] = gear * d->ten_length[t];
      for (int i = 0; i < nv; i++) mom[i] = gear * d->ten_J[t * nv + i];
    }
  }
}

static void crb_and_factor(const OModel* m, OData* d) {
  int nv = m->nv, nb = m->nbody;
  d->crb = d->cinert;
  for (int b = nb - 1; b > 0; b--) {
    int p = m->body_parentid[b];
    if (p > 0) for (int i = 0; i < 10; i++) d->crb[10 * p + i] += d->crb[10 * b + i];
  }
  std::fill(d->M.begin(), d->M.end(), 0.0);
  for (int i = 0; i < nv; i++) {
    real buf[6];
    mul_inert_vec(buf, &d->crb[10 * m->dof_bodyid[i]], &d->cdof[6 * i]);
    for (int j = i; j >= 0; j = m->dof_parentid[j]) {
      real s = 0;
      for (int k = 0; k < 6; k++) s += d->cdof[6 * j + k] * buf[k];
      d->M[i * nv + j] = s; d->M[j * nv + i] = s;
    }
    d->M[i * nv + i] += m->dof_armature[i];
  }
  chol_factor(d->L.data(), d->M.data(), nv);
}

// ------------------------------------------------------------------------------------------------
// collision
// ------------------------------------------------------------------------------------------------
struct RawCon { real dist, pos[3], normal[3], tangent[3]; };

static int raw_plane_sphere(RawCon* c, real margin, const real* ppos, const real* pmat, const real* spos, real radius) {
  real n[3] = {pmat[2], pmat[5], pmat[8]};
  real dif[3] = {spos[0] - ppos[0], spos[1] - ppos[1], spos[2] - ppos[2]};
  real cdist = dot3(dif, n);
  if (cdist > margin + radius) return 0;
  c->dist = cdist - radius;
  for (int i = 0; i < 3; i++) { c->normal[i] = n[i]; c->pos[i] = spos[i] - n[i] * (radius + 0.5 * c->dist); c->tangent[i] = 0; }
  return 1;
}
static int raw_sphere_sphere(RawCon* c, real margin, const real* p1, real r1, const real* p2, real r2) {
  real dif[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
  real cdist = norm3(dif);
  if (cdist > margin + r1 + r2) return 0;
  c->dist = cdist - r1 - r2;
  if (cdist < BMJ_MINVAL) { c->normal[0] = 1; c->normal[1] = c->normal[2] = 0; }
  else for (int i = 0; i < 3; i++) c->normal[i] = dif[i] / cdist;
  for (int i = 0; i < 3; i++) { c->pos[i] = p1[i] + c->normal[i] * (r1 + 0.5 * c->dist); c->tangent[i] = 0; }
  return 1;
}
static inline real clampr(real x, real lo, real hi) { return x < lo ? lo : (x > hi ? hi : x); }

static int collide_capsule_capsule(RawCon* c, real margin, const real* p1, const real* m1, const real* s1,
                                   const real* p2, const real* m2, const real* s2) {
  real a1[3] = {m1[2], m1[5], m1[8]}, a2[3] = {m2[2], m2[5], m2[8]};
  real dif[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
  real ma = dot3(a1, a1), mb = -dot3(a1, a2), mc = dot3(a2, a2);
  real u = -dot3(a1, dif), v = dot3(a2, dif);
  real det = ma * mc - mb * mb;
  real len1 = s1[1], len2 = s2[1];
  if (std::fabs(det) >= BMJ_MINVAL) {
    real x1 = (mc * u - mb * v) / det, x2 = (ma * v - mb * u) / det;
    if (x1 > len1) { x1 = len1; x2 = (v - mb * len1) / mc; }
    else if (x1 < -len1) { x1 = -len1; x2 = (v + mb * len1) / mc; }
    if (x2 > len2) { x2 = len2; x1 = clampr((u - mb * len2) / ma, -len1, len1); }
    else if (x2 < -len2) { x2 = -len2; x1 = clampr((u + mb * len2) / ma, -len1, len1); }
    real v1[3], v2[3];
    for (int i = 0; i < 3; i++) { v1[i] = p1[i] + a1[i] * x1; v2[i] = p2[i] + a2[i] * x2; }
    return raw_sphere_sphere(c, margin, v1, s1[0], v2, s2[0]);
  }
  // parallel axes: test both ends of segment 1 against segment 2, then the ends of 2 against 1 (max 2 contacts)
  int n = 0;
  for (int e = 0; e < 2 && n < 2; e++) {
    real x1 = e == 0 ? len1 : -len1;
    real v1[3]; for (int i = 0; i < 3; i++) v1[i] = p1[i] + a1[i] * x1;
    real w[3] = {v1[0] - p2[0], v1[1] - p2[1], v1[2] - p2[2]};
    real x2 = dot3(w, a2);
    if (x2 < -len2 || x2 > len2) continue;
    real v2[3]; for (int i = 0; i < 3; i++) v2[i] = p2[i] + a2[i] * x2;
    n += raw_sphere_sphere(c + n, margin, v1, s1[0], v2, s2[0]);
  }
  for (int e = 0; e < 2 && n < 2; e++) {
    real x2 = e == 0 ? len2 : -len2;
    real v2[3]; for (int i = 0; i < 3; i++) v2[i] = p2[i] + a2[i] * x2;
    real w[3] = {v2[0] - p1[0], v2[1] - p1[1], v2[2] - p1[2]};
    real x1 = dot3(w, a1);
    if (x1 <= -len1 || x1 >= len1) continue;
    real v1[3]; for (int i = 0; i < 3; i++) v1[i] = p1[i] + a1[i] * x1;
    n += raw_sphere_sphere(c + n, margin, v1, s1[0], v2, s2[0]);
  }
  if (n == 0) {  // ends do not overlap: closest end pair
    real x1 = clampr(u / ma, -len1, len1);
    real v1[3]; for (int i = 0; i < 3; i++) v1[i] = p1[i] + a1[i] * x1;
    real w[3] = {v1[0] - p2[0], v1[1] - p2[1], v1[2] - p2[2]};
    real x2 = clampr(dot3(w, a2), -len2, len2);
    real v2[3]; for (int i = 0; i < 3; i++) v2[i] = p2[i] + a2[i] * x2;
    n = raw_sphere_sphere(c, margin, v1, s1[0], v2, s2[0]);
  }
  return n;
}

// returns number of raw contacts (normal points from geom1 to geom2)
static int narrowphase(RawCon* c, int t1, int t2, real margin, const real* p1, const real* m1, const real* s1,
                       const real* p2, const real* m2, const real* s2) {
  if (t1 == BMJ_GEOM_PLANE) {
    if (t2 == BMJ_GEOM_SPHERE) return raw_plane_sphere(c, margin, p1, m1, p2, s2[0]);
    if (t2 == BMJ_GEOM_CAPSULE) {
      real ax[3] = {m2[2], m2[5], m2[8]};
      real e[3]; int n = 0;
      for (int i = 0; i < 3; i++) e[i] = p2[i] + ax[i] * s2[1];
      int n1 = raw_plane_sphere(c, margin, p1, m1, e, s2[0]);
      if (n1) for (int i = 0; i < 3; i++) c[0].tangent[i] = ax[i];
      n += n1;
      for (int i = 0; i < 3; i++) e[i] = p2[i] - ax[i] * s2[1];
      int n2 = raw_plane_sphere(c + n, margin, p1, m1, e, s2[0]);
      if (n2) for (int i = 0; i < 3; i++) c[n].tangent[i] = ax[i];
      return n + n2;
    }
    if (t2 == BMJ_GEOM_ELLIPSOID) {
      // support point of the ellipsoid along -n
      real n[3] = {m1[2], m1[5], m1[8]};
      real nl[3]; mul_matT_vec3(nl, m2, n);
      real sv[3] = {nl[0] * s2[0], nl[1] * s2[1], nl[2] * s2[2]};
      real len = norm3(sv);
      if (len < BMJ_MINVAL) return 0;
      real loc[3] = {-s2[0] * sv[0] / len, -s2[1] * sv[1] / len, -s2[2] * sv[2] / len};
      real pt[3]; mul_mat_vec3(pt, m2, loc);
      for (int i = 0; i < 3; i++) pt[i] += p2[i];
      real dif[3] = {pt[0] - p1[0], pt[1] - p1[1], pt[2] - p1[2]};
      real dist = dot3(dif, n);
      if (dist > margin) return 0;
      c->dist = dist;
      for (int i = 0; i < 3; i++) { c->normal[i] = n[i]; c->pos[i] = pt[i] - n[i] * dist * 0.5; c->tangent[i] = 0; }
      return 1;
    }
    if (t2 == BMJ_GEOM_BOX) {
      real n[3] = {m1[2], m1[5], m1[8]};
      real dif[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
      real dist0 = dot3(dif, n);
      int cnt = 0;
      for (int k = 0; k < 8 && cnt < 4; k++) {
        real loc[3] = {(k & 1 ? s2[0] : -s2[0]), (k & 2 ? s2[1] : -s2[1]), (k & 4 ? s2[2] : -s2[2])};
        real corner[3]; mul_mat_vec3(corner, m2, loc);
        real ldist = dot3(n, corner);
        if (dist0 + ldist > margin || ldist > 0) continue;
        real dist = dist0 + ldist;
        RawCon* cc = c + cnt;
        cc->dist = dist;
        for (int i = 0; i < 3; i++) { cc->normal[i] = n[i]; cc->pos[i] = p2[i] + corner[i] - n[i] * dist * 0.5; cc->tangent[i] = 0; }
        cnt++;
      }
      return cnt;
    }
    if (t2 == BMJ_GEOM_CYLINDER) {
      real n[3] = {m1[2], m1[5], m1[8]};
      real ax[3] = {m2[2], m2[5], m2[8]};
      real dif[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
      real dist0 = dot3(dif, n);
      real prjaxis = dot3(n, ax);
      if (prjaxis > 0) { for (int i = 0; i < 3; i++) ax[i] = -ax[i]; prjaxis = -prjaxis; }
      // radial direction pointing most into the plane
      real vec[3]; real len_sq = 0;
      for (int i = 0; i < 3; i++) { vec[i] = ax[i] * prjaxis - n[i]; len_sq += vec[i] * vec[i]; }
      real len = std::sqrt(len_sq);
      if (len < 1e-12) {  // disk parallel to plane: pick x axis of the cylinder
        vec[0] = m2[0]; vec[1] = m2[3]; vec[2] = m2[6];
        for (int i = 0; i < 3; i++) vec[i] *= s2[0];
      } else for (int i = 0; i < 3; i++) vec[i] *= s2[0] / len;
      real prjvec = dot3(vec, n);
      real axl[3]; for (int i = 0; i < 3; i++) axl[i] = ax[i] * s2[1];
      real prjax = prjaxis * s2[1];
      int cnt = 0;
      // deepest point on the near disk
      if (dist0 + prjax + prjvec <= margin) {
        RawCon* cc = c + cnt; cc->dist = dist0 + prjax + prjvec;
        for (int i = 0; i < 3; i++) { cc->normal[i] = n[i]; cc->pos[i] = p2[i] + vec[i] + axl[i] - n[i] * cc->dist * 0.5; cc->tangent[i] = 0; }
        cnt++;
      } else return 0;
      // far-disk point on the same radial line
      if (dist0 - prjax + prjvec <= margin) {
        RawCon* cc = c + cnt; cc->dist = dist0 - prjax + prjvec;
        for (int i = 0; i < 3; i++) { cc->normal[i] = n[i]; cc->pos[i] = p2[i] + vec[i] - axl[i] - n[i] * cc->dist * 0.5; cc->tangent[i] = 0; }
        cnt++;
      }
      // two side points on the near disk (triangle with the first point)
      real prjvec1 = -prjvec * 0.5;
      if (dist0 + prjax + prjvec1 <= margin) {
        real v1[3]; cross3(v1, vec, ax);
        real l1 = norm3(v1);
        if (l1 > BMJ_MINVAL) {
          for (int i = 0; i < 3; i++) v1[i] *= s2[0] * std::sqrt(3.0) * 0.5 / l1;
          for (int sgn = -1; sgn <= 1 && cnt < 4; sgn += 2) {
            RawCon* cc = c + cnt; cc->dist = dist0 + prjax + prjvec1;
            for (int i = 0; i < 3; i++) {
              cc->normal[i] = n[i];
              cc->pos[i] = p2[i] + sgn * v1[i] + axl[i] - vec[i] * 0.5 - n[i] * cc->dist * 0.5;
              cc->tangent[i] = 0;
            }
            cnt++;
          }
        }
      }
      return cnt;
    }
    return -1;
  }
  if (t1 == BMJ_GEOM_SPHERE) {
    if (t2 == BMJ_GEOM_SPHERE) return raw_sphere_sphere(c, margin, p1, s1[0], p2, s2[0]);
    if (t2 == BMJ_GEOM_CAPSULE) {
      real ax[3] = {m2[2], m2[5], m2[8]};
      real w[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
      real x = clampr(dot3(ax, w), -s2[1], s2[1]);
      real near[3]; for (int i = 0; i < 3; i++) near[i] = p2[i] + ax[i] * x;
      return raw_sphere_sphere(c, margin, p1, s1[0], near, s2[0]);
    }
    if (t2 == BMJ_GEOM_BOX) {
      // closest point on the box to the sphere centre (centre outside the box)
      real w[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
      real loc[3]; mul_matT_vec3(loc, m2, w);
      real cl[3]; bool inside = true;
      for (int i = 0; i < 3; i++) { cl[i] = clampr(loc[i], -s2[i], s2[i]); if (cl[i] != loc[i]) inside = false; }
      real nl[3], dist;
      if (!inside) {
        real dl[3] = {loc[0] - cl[0], loc[1] - cl[1], loc[2] - cl[2]};
        real dn = norm3(dl);
        if (dn - s1[0] > margin) return 0;
        dist = dn - s1[0];
        for (int i = 0; i < 3; i++) nl[i] = -dl[i] / dn;   // from sphere toward box
      } else {
        int best = 0; real bd = 1e300;
        for (int i = 0; i < 3; i++) { real dd = s2[i] - std::fabs(loc[i]); if (dd < bd) { bd = dd; best = i; } }
        nl[0] = nl[1] = nl[2] = 0; nl[best] = loc[best] > 0 ? -1 : 1;
        dist = -bd - s1[0];
        for (int i = 0; i < 3; i++) cl[i] = loc[i];
      }
      real nw[3]; mul_mat_vec3(nw, m2, nl);
      for (int i = 0; i < 3; i++) { c->normal[i] = nw[i]; c->pos[i] = p1[i] + nw[i] * (s1[0] + 0.5 * dist); c->tangent[i] = 0; }
      c->dist = dist;
      return 1;
    }
  }
  if (t1 == BMJ_GEOM_CAPSULE && t2 == BMJ_GEOM_CAPSULE) return collide_capsule_capsule(c, margin, p1, m1, s1, p2, m2, s2);
  if (t1 == BMJ_GEOM_CAPSULE && t2 == BMJ_GEOM_BOX) {
    real out[14];
    int n = cvx_capsule_box(p1, m1, s1, p2, m2, s2, margin, out);
    for (int k = 0; k < n; k++) {
      c[k].dist = out[7 * k];
      for (int i = 0; i < 3; i++) { c[k].pos[i] = out[7 * k + 1 + i]; c[k].normal[i] = out[7 * k + 4 + i]; c[k].tangent[i] = 0; }
    }
    return n;
  }
  if (t1 >= BMJ_GEOM_SPHERE && t2 <= BMJ_GEOM_BOX) {
    // every remaining pair of convex primitives (an ellipsoid, a cylinder or two boxes involved): MPR, one contact
    int n = cvx_pair(t1, p1, m1, s1, t2, p2, m2, s2, margin, &c->dist, c->pos, c->normal);
    if (n) for (int i = 0; i < 3; i++) c->tangent[i] = 0;
    return n;
  }
  return -1;
}

static void make_frame(real* frame, const real* normal, const real* tangent) {
  real x[3] = {normal[0], normal[1], normal[2]};
  normalize3(x);
  real y[3] = {tangent[0], tangent[1], tangent[2]};
  if (norm3(y) < 0.5) {
    y[0] = y[1] = y[2] = 0;
    if (x[1] < 0.5 && x[1] > -0.5) y[1] = 1; else y[2] = 1;
  }
  real dp = dot3(x, y);
  for (int i = 0; i < 3; i++) y[i] -= dp * x[i];
  normalize3(y);
  real z[3]; cross3(z, x, y);
  for (int i = 0; i < 3; i++) { frame[i] = x[i]; frame[3 + i] = y[i]; frame[6 + i] = z[i]; }
}

static void collision(const OModel* m, OData* d) {
  d->ncon = 0;
  if (m->disableflags & (BMJ_DSBL_CONTACT | BMJ_DSBL_CONSTRAINT)) return;
  for (int p = 0; p < m->npair; p++) {
    int g1 = m->pair_geom1[p], g2 = m->pair_geom2[p];
    int t1 = m->geom_type[g1], t2 = m->geom_type[g2];
    real margin = std::max(m->geom_margin[g1], m->geom_margin[g2]);
    real gap = std::max(m->geom_gap[g1], m->geom_gap[g2]);
    const real* p1 = &d->geom_xpos[3 * g1]; const real* p2 = &d->geom_xpos[3 * g2];
    const real* m1 = &d->geom_xmat[9 * g1]; const real* m2 = &d->geom_xmat[9 * g2];
    // bounding-sphere rejection
    if (t1 == BMJ_GEOM_PLANE) {
      real n[3] = {m1[2], m1[5], m1[8]};
      real dif[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
      if (dot3(dif, n) > m->geom_rbound[g2] + margin) continue;
    } else {
      real dif[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
      real bound = m->geom_rbound[g1] + m->geom_rbound[g2] + margin;
      if (dot3(dif, dif) > bound * bound) continue;
    }
    RawCon raw[8];
    int n = narrowphase(raw, t1, t2, margin, p1, m1, &m->geom_size[3 * g1], p2, m2, &m->geom_size[3 * g2]);
    if (n <= 0) continue;   // (-1: pair function outside the supported subset)
    // contact parameter mixing
    int pr1 = m->geom_priority[g1], pr2 = m->geom_priority[g2];
    int condim; real fr[3], solref[2], solimp[5];
    if (pr1 != pr2) {
      int gp = pr1 > pr2 ? g1 : g2;
      condim = m->geom_condim[gp];
      for (int i = 0; i < 3; i++) fr[i] = m->geom_friction[3 * gp + i];
      for (int i = 0; i < 2; i++) solref[i] = m->geom_solref[2 * gp + i];
      for (int i = 0; i < 5; i++) solimp[i] = m->geom_solimp[5 * gp + i];
    } else {
      condim = std::max(m->geom_condim[g1], m->geom_condim[g2]);
      for (int i = 0; i < 3; i++) fr[i] = std::max(m->geom_friction[3 * g1 + i], m->geom_friction[3 * g2 + i]);
      real s1 = m->geom_solmix[g1], s2 = m->geom_solmix[g2], mix;
      if (s1 >= BMJ_MINVAL && s2 >= BMJ_MINVAL) mix = s1 / (s1 + s2);
      else if (s1 < BMJ_MINVAL && s2 < BMJ_MINVAL) mix = 0.5;
      else if (s1 < BMJ_MINVAL) mix = 0.0; else mix = 1.0;
      if (m->geom_solref[2 * g1] > 0 && m->geom_solref[2 * g2] > 0)
        for (int i = 0; i < 2; i++) solref[i] = mix * m->geom_solref[2 * g1 + i] + (1 - mix) * m->geom_solref[2 * g2 + i];
      else
        for (int i = 0; i < 2; i++) solref[i] = std::min(m->geom_solref[2 * g1 + i], m->geom_solref[2 * g2 + i]);
      for (int i = 0; i < 5; i++) solimp[i] = mix * m->geom_solimp[5 * g1 + i] + (1 - mix) * m->geom_solimp[5 * g2 + i];
    }
    for (int k = 0; k < n; k++) {
      if (d->ncon >= m->nconmax) { d->warning[BMJ_WARN_CONTACTFULL]++; return; }
      OContact* c = &d->contact[d->ncon++];
      c->dist = raw[k].dist;
      for (int i = 0; i < 3; i++) c->pos[i] = raw[k].pos[i];
      make_frame(c->frame, raw[k].normal, raw[k].tangent);
      c->includemargin = margin - gap;
      c->friction[0] = fr[0]; c->friction[1] = fr[0]; c->friction[2] = fr[1]; c->friction[3] = fr[2]; c->friction[4] = fr[2];
      for (int i = 0; i < 2; i++) c->solref[i] = solref[i];
      for (int i = 0; i < 5; i++) c->solimp[i] = solimp[i];
      c->dim = condim; c->geom1 = g1; c->geom2 = g2; c->efc_address = -1; c->mu = fr[0];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// constraint assembly
// ------------------------------------------------------------------------------------------------
// translational (jp) and rotational (jr) Jacobian rows of a world point attached to `body`
static void jac_point(const OModel* m, const OData* d, real* jp, real* jr, const real* point, int body) {
  int nv = m->nv;
  for (int i = 0; i < 3 * nv; i++) { jp[i] = 0; if (jr) jr[i] = 0; }
  if (body <= 0) return;
  int root = m->body_rootid[body];
  real off[3];
  for (int i = 0; i < 3; i++) off[i] = point[i] - d->subtree_com[3 * root + i];
  // walk to the first body with dofs
  int b = body;
  while (b > 0 && m->body_dofnum[b] == 0) b = m->body_parentid[b];
  if (b <= 0) return;
  int i = m->body_dofadr[b] + m->body_dofnum[b] - 1;
  while (i >= 0) {
    const real* cd = &d->cdof[6 * i];
    real tmp[3]; cross3(tmp, cd, off);
    for (int k = 0; k < 3; k++) { jp[k * nv + i] = cd[3 + k] + tmp[k]; if (jr) jr[k * nv + i] = cd[k]; }
    i = m->dof_parentid[i];
  }
}

static real impedance(const real* solimp, real pos, real margin) {
  real d0 = solimp[0], dmax = solimp[1], width = solimp[2], mid = solimp[3], power = solimp[4];
  d0 = clampr(d0, BMJ_MINIMP, BMJ_MAXIMP); dmax = clampr(dmax, BMJ_MINIMP, BMJ_MAXIMP);
  width = std::max(BMJ_MINVAL, width); mid = clampr(mid, BMJ_MINIMP, BMJ_MAXIMP); power = std::max(1.0, power);
  real x = std::fabs(pos - margin) / width;
  if (x >= 1) return dmax;
  if (x <= 0) return d0;
  real y;
  if (power == 1) y = x;
  else if (x <= mid) y = std::pow(x / mid, power) * mid;       // == x^p / mid^(p-1)
  else y = 1 - std::pow((1 - x) / (1 - mid), power) * (1 - mid);
  return d0 + y * (dmax - d0);
}

static bool add_row(const OModel* m, OData* d, int type, int id, real pos, real margin, real diag) {
  if (d->nefc >= m->njmax) { d->warning[BMJ_WARN_CNSTRFULL]++; return false; }
  int r = d->nefc++;
  d->efc_type[r] = type; d->efc_id[r] = id; d->efc_pos[r] = pos; d->efc_margin[r] = margin; d->efc_diagApprox[r] = diag;
  return true;
}

static void make_constraint(const OModel* m, OData* d) {
  int nv = m->nv;
  d->nefc = 0;
  if (m->disableflags & BMJ_DSBL_CONSTRAINT) return;
  // ---- equality -------------------------------------------------------------------------------
  if (!(m->disableflags & BMJ_DSBL_EQUALITY)) {
    for (int e = 0; e < m->neq; e++) {
      if (!m->eq_active0[e]) continue;
      const real* data = &m->eq_data[11 * e];
      int o1 = m->eq_obj1id[e], o2 = m->eq_obj2id[e];
      if (m->eq_type[e] == BMJ_EQ_TENDON) {
        real pos = d->ten_length[o1] - m->tendon_length0[o1];
        real diag = m->tendon_invweight0[o1];
        int r = d->nefc;
        if (o2 >= 0) {
          real dif = d->ten_length[o2] - m->tendon_length0[o2];
          pos -= data[0] + data[1]*dif + data[2]*dif*dif + data[3]*dif*dif*dif + data[4]*dif*dif*dif*dif;
          diag += m->tendon_invweight0[o2];
          if (!add_row(m, d, BMJ_CNSTR_EQUALITY, e, pos, 0, diag)) return;
          real deriv = data[1] + 2*data[2]*dif + 3*data[3]*dif*dif + 4*data[4]*dif*dif*dif;
          for (int i = 0; i < nv; i++) d->efc_J[r * nv + i] = d->ten_J[o1 * nv + i] - deriv * d->ten_J[o2 * nv + i];
        } else {
          pos -= data[0];
          if (!add_row(m, d, BMJ_CNSTR_EQUALITY, e, pos, 0, diag)) return;
          for (int i = 0; i < nv; i++) d->efc_J[r * nv + i] = d->ten_J[o1 * nv + i];
        }
      } else if (m->eq_type[e] == BMJ_EQ_JOINT) {
        int qa1 = m->jnt_qposadr[o1], da1 = m->jnt_dofadr[o1];
        real pos = d->qpos[qa1] - m->qpos0[qa1];
        real diag = m->dof_invweight0[da1];
        int r = d->nefc;
        if (o2 >= 0) {
          int qa2 = m->jnt_qposadr[o2], da2 = m->jnt_dofadr[o2];
          real dif = d->qpos[qa2] - m->qpos0[qa2];
          pos -= data[0] + data[1]*dif + data[2]*dif*dif + data[3]*dif*dif*dif + data[4]*dif*dif*dif*dif;
          diag += m->dof_invweight0[da2];
          if (!add_row(m, d, BMJ_CNSTR_EQUALITY, e, pos, 0, diag)) return;
          real deriv = data[1] + 2*data[2]*dif + 3*data[3]*dif*dif + 4*data[4]*dif*dif*dif;
          for (int i = 0; i < nv; i++) d->efc_J[r * nv + i] = 0;
          d->efc_J[r * nv + da1] = 1; d->efc_J[r * nv + da2] = -deriv;
        } else {
          pos -= data[0];
          if (!add_row(m, d, BMJ_CNSTR_EQUALITY, e, pos, 0, diag)) return;
          for (int i = 0; i < nv; i++) d->efc_J[r * nv + i] = 0;
          d->efc_J[r * nv + da1] = 1;
        }
      }
    }
  }
  // ---- joint limits ---------------------------------------------------------------------------
  if (!(m->disableflags & BMJ_DSBL_LIMIT)) {
    for (int j = 0; j < m->njnt; j++) {
      if (!m->jnt_limited[j]) continue;
      int t = m->jnt_type[j];
      if (t != BMJ_JNT_SLIDE && t != BMJ_JNT_HINGE) continue;
      real value = d->qpos[m->jnt_qposadr[j]], margin = m->jnt_margin[j];
      for (int side = -1; side <= 1; side += 2) {
        real dist = side * (m->jnt_range[2 * j + (side + 1) / 2] - value);
        if (dist < margin) {
          int r = d->nefc;
          if (!add_row(m, d, BMJ_CNSTR_LIMIT_JOINT, j, dist, margin, m->dof_invweight0[m->jnt_dofadr[j]])) return;
          for (int i = 0; i < nv; i++) d->efc_J[r * nv + i] = 0;
          d->efc_J[r * nv + m->jnt_dofadr[j]] = -side;
        }
      }
    }
  }
  // ---- contacts -------------------------------------------------------------------------------
  static thread_local vec jp1, jp2, jr1, jr2, jf;
  jp1.resize(3 * nv); jp2.resize(3 * nv); jr1.resize(3 * nv); jr2.resize(3 * nv); jf.resize(6 * nv);
  for (int ci = 0; ci < d->ncon; ci++) {
    OContact* c = &d->contact[ci];
    c->efc_address = -1;
    if (c->dist >= c->includemargin) continue;   // inside the gap buffer: no constraint rows
    int b1 = m->geom_bodyid[c->geom1], b2 = m->geom_bodyid[c->geom2];
    jac_point(m, d, jp1.data(), jr1.data(), c->pos, b1);
    jac_point(m, d, jp2.data(), jr2.data(), c->pos, b2);
    int dim = c->dim;
    // contact-frame Jacobian rows: translational 0..2, rotational 3..5
    for (int r = 0; r < 3; r++) for (int i = 0; i < nv; i++) {
      real st = 0, sr = 0;
      for (int k = 0; k < 3; k++) {
        st += c->frame[3 * r + k] * (jp2[k * nv + i] - jp1[k * nv + i]);
        sr += c->frame[3 * r + k] * (jr2[k * nv + i] - jr1[k * nv + i]);
      }
      jf[r * nv + i] = st; jf[(3 + r) * nv + i] = sr;
    }
    real tran = m->body_invweight0[2 * b1] + m->body_invweight0[2 * b2];
    real rot = m->body_invweight0[2 * b1 + 1] + m->body_invweight0[2 * b2 + 1];
    if (dim == 1) {
      int r = d->nefc;
      if (!add_row(m, d, BMJ_CNSTR_CONTACT_FRICTIONLESS, ci, c->dist, c->includemargin, tran)) return;
      c->efc_address = r;
      for (int i = 0; i < nv; i++) d->efc_J[r * nv + i] = jf[i];
    } else {
      if (d->nefc + 2 * (dim - 1) > m->njmax) { d->warning[BMJ_WARN_CNSTRFULL]++; return; }
      c->efc_address = d->nefc;
      for (int k = 1; k < dim; k++) {
        real fri = c->friction[k - 1];
        real diag = tran + fri * fri * (k < 3 ? tran : rot);
        for (int sgn = 1; sgn >= -1; sgn -= 2) {
          int r = d->nefc;
          add_row(m, d, BMJ_CNSTR_CONTACT_PYRAMIDAL, ci, c->dist, c->includemargin, diag);
          for (int i = 0; i < nv; i++) d->efc_J[r * nv + i] = jf[i] + sgn * fri * jf[k * nv + i];
        }
      }
    }
  }
}

static void make_impedance(const OModel* m, OData* d) {
  int nv = m->nv;
  bool refsafe = !(m->disableflags & BMJ_DSBL_REFSAFE);
  for (int r = 0; r < d->nefc; r++) {
    const real *solref, *solimp;
    int id = d->efc_id[r];
    switch (d->efc_type[r]) {
      case BMJ_CNSTR_EQUALITY: solref = &m->eq_solref[2 * id]; solimp = &m->eq_solimp[5 * id]; break;
      case BMJ_CNSTR_LIMIT_JOINT: solref = &m->jnt_solref[2 * id]; solimp = &m->jnt_solimp[5 * id]; break;
      default: solref = d->contact[id].solref; solimp = d->contact[id].solimp; break;
    }
    real pos = d->efc_pos[r], margin = d->efc_margin[r];
    real imp = impedance(solimp, pos, margin);
    real dmax = clampr(solimp[1], BMJ_MINIMP, BMJ_MAXIMP);
    real K, B;
    if (solref[0] > 0) {
      real tc = solref[0], dr = solref[1];
      if (refsafe) tc = std::max(tc, 2 * m->timestep);
      K = 1 / std::max(BMJ_MINVAL, dmax * dmax * tc * tc * dr * dr);
      B = 2 / std::max(BMJ_MINVAL, dmax * tc);
    } else { K = -solref[0] / std::max(BMJ_MINVAL, dmax * dmax); B = -solref[1] / std::max(BMJ_MINVAL, dmax); }
    d->efc_R[r] = std::max(BMJ_MINVAL, (1 - imp) * d->efc_diagApprox[r] / imp);
    real vel = 0;
    for (int i = 0; i < nv; i++) vel += d->efc_J[r * nv + i] * d->qvel[i];
    d->efc_vel[r] = vel;
    d->efc_aref[r] = -B * vel - K * imp * (pos - margin);
  }
  // pyramidal contacts share one regulariser: R_py = 2 mu^2 R(first edge)
  for (int ci = 0; ci < d->ncon; ci++) {
    OContact* c = &d->contact[ci];
    if (c->efc_address < 0 || c->dim == 1) continue;
    int a = c->efc_address, nrow = 2 * (c->dim - 1);
    real mu = c->friction[0] / std::sqrt(std::max(BMJ_MINVAL, m->impratio));
    c->mu = mu;
    real Rpy = std::max(BMJ_MINVAL, 2 * mu * mu * d->efc_R[a]);
    for (int k = 0; k < nrow; k++) d->efc_R[a + k] = Rpy;
  }
  for (int r = 0; r < d->nefc; r++) d->efc_D[r] = 1 / d->efc_R[r];
}

// ------------------------------------------------------------------------------------------------
// velocity stage
// ------------------------------------------------------------------------------------------------
static void com_vel(const OModel* m, OData* d) {
  for (int i = 0; i < 6; i++) d->cvel[i] = 0;
  for (int b = 1; b < m->nbody; b++) {
    real cv[6]; int p = m->body_parentid[b];
    for (int i = 0; i < 6; i++) cv[i] = d->cvel[6 * p + i];
    for (int j = m->body_jntadr[b]; j < m->body_jntadr[b] + m->body_jntnum[b]; j++) {
      int da = m->jnt_dofadr[j], t = m->jnt_type[j];
      if (t == BMJ_JNT_FREE) {
        for (int i = 0; i < 18; i++) d->cdof_dot[6 * da + i] = 0;
        for (int k = 0; k < 3; k++) for (int i = 0; i < 6; i++) cv[i] += d->cdof[6 * (da + k) + i] * d->qvel[da + k];
        da += 3;
      }
      if (t == BMJ_JNT_FREE || t == BMJ_JNT_BALL) {
        for (int k = 0; k < 3; k++) cross_motion(&d->cdof_dot[6 * (da + k)], cv, &d->cdof[6 * (da + k)]);
        for (int k = 0; k < 3; k++) for (int i = 0; i < 6; i++) cv[i] += d->cdof[6 * (da + k) + i] * d->qvel[da + k];
      } else {
        cross_motion(&d->cdof_dot[6 * da], cv, &d->cdof[6 * da]);
        for (int i = 0; i < 6; i++) cv[i] += d->cdof[6 * da + i] * d->qvel[da];
      }
    }
    for (int i = 0; i < 6; i++) d->cvel[6 * b + i] = cv[i];
  }
}

static void passive(const OModel* m, OData* d) {
  int nv = m->nv;
  for (int i = 0; i < nv; i++) d->qfrc_passive[i] = 0;
  if (m->disableflags & BMJ_DSBL_PASSIVE) return;
  for (int j = 0; j < m->njnt; j++) {
    int t = m->jnt_type[j];
    if ((t == BMJ_JNT_SLIDE || t == BMJ_JNT_HINGE) && m->jnt_stiffness[j] != 0) {
      int qa = m->jnt_qposadr[j];
      d->qfrc_passive[m->jnt_dofadr[j]] -= m->jnt_stiffness[j] * (d->qpos[qa] - m->qpos_spring[qa]);
    }
  }
  for (int i = 0; i < nv; i++) d->qfrc_passive[i] -= m->dof_damping[i] * d->qvel[i];
}

// recursive Newton-Euler: qfrc_bias (flg_acc=0) — gravity enters as base acceleration
static void rne(const OModel* m, OData* d, real* result) {
  int nb = m->nbody, nv = m->nv;
  static thread_local vec loc_cacc, loc_cfrc;
  loc_cacc.assign(6 * nb, 0.0); loc_cfrc.assign(6 * nb, 0.0);
  if (!(m->disableflags & BMJ_DSBL_GRAVITY)) for (int i = 0; i < 3; i++) loc_cacc[3 + i] = -m->gravity[i];
  for (int b = 1; b < nb; b++) {
    int p = m->body_parentid[b];
    real* ca = &loc_cacc[6 * b];
    for (int i = 0; i < 6; i++) ca[i] = loc_cacc[6 * p + i];
    for (int k = m->body_dofadr[b]; k < m->body_dofadr[b] + m->body_dofnum[b]; k++)
      for (int i = 0; i < 6; i++) ca[i] += d->cdof_dot[6 * k + i] * d->qvel[k];
    real t1[6], t2[6], t3[6];
    mul_inert_vec(t1, &d->cinert[10 * b], ca);
    mul_inert_vec(t2, &d->cinert[10 * b], &d->cvel[6 * b]);
    cross_force(t3, &d->cvel[6 * b], t2);
    for (int i = 0; i < 6; i++) loc_cfrc[6 * b + i] = t1[i] + t3[i];
  }
  for (int b = nb - 1; b > 0; b--) { int p = m->body_parentid[b]; if (p > 0) for (int i = 0; i < 6; i++) loc_cfrc[6 * p + i] += loc_cfrc[6 * b + i]; }
  for (int k = 0; k < nv; k++) {
    real s = 0; int b = m->dof_bodyid[k];
    for (int i = 0; i < 6; i++) s += d->cdof[6 * k + i] * loc_cfrc[6 * b + i];
    result[k] = s;
  }
}

static void fwd_velocity(const OModel* m, OData* d) {
  int nv = m->nv;
  for (int t = 0; t < m->ntendon; t++) { real s = 0; for (int i = 0; i < nv; i++) s += d->ten_J[t * nv + i] * d->qvel[i]; d->ten_velocity[t] = s; }
  for (int a = 0; a < m->nu; a++) { real s = 0; for (int i = 0; i < nv; i++) s += d->actuator_moment[a * nv + i] * d->qvel[i]; d->actuator_velocity[a] = s; }
  com_vel(m, d);
  passive(m, d);
  rne(m, d, d->qfrc_bias.data());
}

static void subtree_vel(const OModel* m, OData* d) {
  int nb = m->nbody;
  for (int b = 0; b < nb; b++) {
    int root = m->body_rootid[b];
    real dif[3], tmp[3];
    for (int i = 0; i < 3; i++) dif[i] = d->xipos[3 * b + i] - d->subtree_com[3 * root + i];
    cross3(tmp, dif, &d->cvel[6 * b]);      // lin_new = lin - dif x ang
    for (int i = 0; i < 3; i++) d->subtree_linvel[3 * b + i] = m->body_mass[b] * (d->cvel[6 * b + 3 + i] - tmp[i]);
  }
  for (int b = nb - 1; b > 0; b--) { int p = m->body_parentid[b]; for (int i = 0; i < 3; i++) d->subtree_linvel[3 * p + i] += d->subtree_linvel[3 * b + i]; }
  for (int b = 0; b < nb; b++) {
    real sm = std::max(BMJ_MINVAL, m->body_subtreemass[b]);
    for (int i = 0; i < 3; i++) d->subtree_linvel[3 * b + i] /= sm;
  }
}

// ------------------------------------------------------------------------------------------------
// sensors
// ------------------------------------------------------------------------------------------------
static void object_velocity_site(const OModel* m, const OData* d, int site, real* res /*ang,lin local*/) {
  int b = m->site_bodyid[site], root = m->body_rootid[b];
  real dif[3], tmp[3], lin[3];
  for (int i = 0; i < 3; i++) dif[i] = d->site_xpos[3 * site + i] - d->subtree_com[3 * root + i];
  cross3(tmp, dif, &d->cvel[6 * b]);
  for (int i = 0; i < 3; i++) lin[i] = d->cvel[6 * b + 3 + i] - tmp[i];
  mul_matT_vec3(res, &d->site_xmat[9 * site], &d->cvel[6 * b]);
  mul_matT_vec3(res + 3, &d->site_xmat[9 * site], lin);
}


// ---- ray / convex zone test used by the touch sensor: does p + t v (t >= 0) meet the shape? ----------------
struct Interval { real lo, hi; bool ok; };
static inline Interval iv_all() { Interval r = {-1e300, 1e300, true}; return r; }
static inline void iv_clip(Interval& a, real lo, real hi) { if (lo > hi) { real t = lo; lo = hi; hi = t; } a.lo = std::max(a.lo, lo); a.hi = std::min(a.hi, hi); if (a.lo > a.hi) a.ok = false; }
// quadratic a t^2 + 2 b t + c <= 0
static inline void iv_quadric(Interval& r, real a, real b, real c) {
  if (a < BMJ_MINVAL) { if (c > 0) r.ok = false; return; }
  real det = b * b - a * c;
  if (det < 0) { r.ok = false; return; }
  real sq = std::sqrt(det);
  iv_clip(r, (-b - sq) / a, (-b + sq) / a);
}
static inline void iv_slab(Interval& r, real p, real v, real half) {
  if (std::fabs(v) < BMJ_MINVAL) { if (std::fabs(p) > half) r.ok = false; return; }
  iv_clip(r, (-half - p) / v, (half - p) / v);
}
static bool ray_hits_zone(int type, const real* sz, const real* p, const real* v) {
  Interval r = iv_all();
  switch (type) {
    case BMJ_GEOM_SPHERE: iv_quadric(r, dot3(v, v), dot3(p, v), dot3(p, p) - sz[0] * sz[0]); break;
    case BMJ_GEOM_ELLIPSOID: {
      real ps[3] = {p[0] / sz[0], p[1] / sz[1], p[2] / sz[2]}, vs[3] = {v[0] / sz[0], v[1] / sz[1], v[2] / sz[2]};
      iv_quadric(r, dot3(vs, vs), dot3(ps, vs), dot3(ps, ps) - 1);
    } break;
    case BMJ_GEOM_BOX: for (int i = 0; i < 3; i++) iv_slab(r, p[i], v[i], sz[i]); break;
    case BMJ_GEOM_CYLINDER:
      iv_quadric(r, v[0]*v[0] + v[1]*v[1], p[0]*v[0] + p[1]*v[1], p[0]*p[0] + p[1]*p[1] - sz[0]*sz[0]);
      iv_slab(r, p[2], v[2], sz[1]);
      break;
    case BMJ_GEOM_CAPSULE: {
      iv_quadric(r, v[0]*v[0] + v[1]*v[1], p[0]*v[0] + p[1]*v[1], p[0]*p[0] + p[1]*p[1] - sz[0]*sz[0]);
      iv_slab(r, p[2], v[2], sz[1]);
      if (r.ok && r.hi >= 0) return true;
      for (int s = -1; s <= 1; s += 2) {
        Interval q = iv_all();
        real pc[3] = {p[0], p[1], p[2] - s * sz[1]};
        iv_quadric(q, dot3(v, v), dot3(pc, v), dot3(pc, pc) - sz[0] * sz[0]);
        if (q.ok && q.hi >= 0) return true;
      }
      return false;
    }
    default: return false;
  }
  return r.ok && r.hi >= 0;
}

static void sensors(const OModel* m, OData* d, int stage) {
  if (m->disableflags & BMJ_DSBL_SENSOR) return;
  bool did_subtree = false;
  for (int s = 0; s < m->nsensor; s++) {
    if (m->sensor_needstage[s] != stage) continue;
    real* out = &d->sensordata[m->sensor_adr[s]];
    int id = m->sensor_objid[s];
    switch (m->sensor_type[s]) {
      case BMJ_SENS_JOINTPOS: out[0] = d->qpos[m->jnt_qposadr[id]]; break;
      case BMJ_SENS_JOINTVEL: out[0] = d->qvel[m->jnt_dofadr[id]]; break;
      case BMJ_SENS_ACTUATORFRC: out[0] = d->actuator_force[id]; break;
      case BMJ_SENS_SUBTREECOM: for (int i = 0; i < 3; i++) out[i] = d->subtree_com[3 * id + i]; break;
      case BMJ_SENS_SUBTREELINVEL:
        if (!did_subtree) { subtree_vel(m, d); did_subtree = true; }
        for (int i = 0; i < 3; i++) out[i] = d->subtree_linvel[3 * id + i];
        break;
      case BMJ_SENS_FRAMEPOS: {
        const real* p = m->sensor_objtype[s] == BMJ_OBJ_SITE ? &d->site_xpos[3 * id]
                      : m->sensor_objtype[s] == BMJ_OBJ_GEOM ? &d->geom_xpos[3 * id]
                      : m->sensor_objtype[s] == BMJ_OBJ_BODY ? &d->xipos[3 * id] : &d->xpos[3 * id];
        int rid = m->sensor_refid[s];
        if (rid < 0) { for (int i = 0; i < 3; i++) out[i] = p[i]; break; }
        const real *rp, *rm;
        switch (m->sensor_reftype[s]) {
          case BMJ_OBJ_SITE: rp = &d->site_xpos[3 * rid]; rm = &d->site_xmat[9 * rid]; break;
          case BMJ_OBJ_GEOM: rp = &d->geom_xpos[3 * rid]; rm = &d->geom_xmat[9 * rid]; break;
          case BMJ_OBJ_BODY: rp = &d->xipos[3 * rid]; rm = &d->ximat[9 * rid]; break;
          default: rp = &d->xpos[3 * rid]; rm = &d->xmat[9 * rid]; break;
        }
        real dif[3] = {p[0] - rp[0], p[1] - rp[1], p[2] - rp[2]};
        mul_matT_vec3(out, rm, dif);
      } break;
      case BMJ_SENS_VELOCIMETER: { real v[6]; object_velocity_site(m, d, id, v); for (int i = 0; i < 3; i++) out[i] = v[3 + i]; } break;
      case BMJ_SENS_GYRO: { real v[6]; object_velocity_site(m, d, id, v); for (int i = 0; i < 3; i++) out[i] = v[i]; } break;
      case BMJ_SENS_ACCELEROMETER: {
        int b = m->site_bodyid[id], root = m->body_rootid[b];
        real dif[3], tmp[3], lin[3], v[6], acc[3], corr[3];
        for (int i = 0; i < 3; i++) dif[i] = d->site_xpos[3 * id + i] - d->subtree_com[3 * root + i];
        cross3(tmp, dif, &d->cacc[6 * b]);
        for (int i = 0; i < 3; i++) lin[i] = d->cacc[6 * b + 3 + i] - tmp[i];
        mul_matT_vec3(acc, &d->site_xmat[9 * id], lin);
        object_velocity_site(m, d, id, v);
        cross3(corr, v, v + 3);
        for (int i = 0; i < 3; i++) out[i] = acc[i] + corr[i];
      } break;
      case BMJ_SENS_FORCE: {
        int b = m->site_bodyid[id];
        mul_matT_vec3(out, &d->site_xmat[9 * id], &d->cfrc_int[6 * b + 3]);
      } break;
      case BMJ_SENS_TORQUE: {
        int b = m->site_bodyid[id], root = m->body_rootid[b];
        real dif[3], tmp[3], tq[3];
        for (int i = 0; i < 3; i++) dif[i] = d->site_xpos[3 * id + i] - d->subtree_com[3 * root + i];
        cross3(tmp, dif, &d->cfrc_int[6 * b + 3]);
        for (int i = 0; i < 3; i++) tq[i] = d->cfrc_int[6 * b + i] - tmp[i];
        mul_matT_vec3(out, &d->site_xmat[9 * id], tq);
      } break;
      case BMJ_SENS_TOUCH: {
        int b = m->site_bodyid[id];
        real total = 0;
        for (int ci = 0; ci < d->ncon; ci++) {
          const OContact* c = &d->contact[ci];
          if (c->efc_address < 0) continue;
          int b1 = m->geom_bodyid[c->geom1], b2 = m->geom_bodyid[c->geom2];
          if (b1 != b && b2 != b) continue;
          real nf = 0;
          if (c->dim == 1) nf = d->efc_force[c->efc_address];
          else for (int k = 0; k < 2 * (c->dim - 1); k++) nf += d->efc_force[c->efc_address + k];
          if (nf <= 0) continue;
          // MuJoCo counts a contact when the ray from the contact point along the (body-oriented) normal meets
          // the site volume (mju_rayGeom >= 0), which also admits points exactly on the zone boundary
          real dif[3] = {c->pos[0] - d->site_xpos[3 * id], c->pos[1] - d->site_xpos[3 * id + 1], c->pos[2] - d->site_xpos[3 * id + 2]};
          real loc[3]; mul_matT_vec3(loc, &d->site_xmat[9 * id], dif);
          real ray[3] = {c->frame[0], c->frame[1], c->frame[2]};
          if (b2 == b) for (int i = 0; i < 3; i++) ray[i] = -ray[i];
          real vloc[3]; mul_matT_vec3(vloc, &d->site_xmat[9 * id], ray);
          bool in = ray_hits_zone(m->site_type[id], &m->site_size[3 * id], loc, vloc);
          if (in) total += nf;
        }
        out[0] = total;
      } break;
      default: break;
    }
  }
}

// cacc / cfrc_int / cfrc_ext for acceleration-stage sensors
static void rne_post_constraint(const OModel* m, OData* d) {
  int nb = m->nbody;
  std::fill(d->cfrc_ext.begin(), d->cfrc_ext.end(), 0.0);
  for (int b = 1; b < nb; b++) {
    const real* xf = &d->xfrc_applied[6 * b];
    int root = m->body_rootid[b];
    real dif[3], tq[3];
    for (int i = 0; i < 3; i++) dif[i] = d->xipos[3 * b + i] - d->subtree_com[3 * root + i];
    cross3(tq, dif, xf);
    for (int i = 0; i < 3; i++) { d->cfrc_ext[6 * b + i] += xf[3 + i] + tq[i]; d->cfrc_ext[6 * b + 3 + i] += xf[i]; }
  }
  for (int ci = 0; ci < d->ncon; ci++) {
    const OContact* c = &d->contact[ci];
    if (c->efc_address < 0) continue;
    // contact force in the contact frame (normal, t1, t2)
    real f[3] = {0, 0, 0};
    if (c->dim == 1) f[0] = d->efc_force[c->efc_address];
    else for (int k = 1; k < c->dim && k < 3; k++) {
      real fp = d->efc_force[c->efc_address + 2 * (k - 1)], fn = d->efc_force[c->efc_address + 2 * (k - 1) + 1];
      f[0] += fp + fn; f[k] += (fp - fn) * c->friction[k - 1];
    }
    real fw[3]; mul_matT_vec3(fw, c->frame, f);   // frame rows are axes: world = frame^T f
    int bb[2] = {m->geom_bodyid[c->geom1], m->geom_bodyid[c->geom2]};
    for (int side = 0; side < 2; side++) {
      int b = bb[side]; if (b <= 0) continue;
      real sgn = side == 0 ? -1 : 1;
      int root = m->body_rootid[b];
      real dif[3], tq[3];
      for (int i = 0; i < 3; i++) dif[i] = c->pos[i] - d->subtree_com[3 * root + i];
      cross3(tq, dif, fw);
      for (int i = 0; i < 3; i++) { d->cfrc_ext[6 * b + i] += sgn * tq[i]; d->cfrc_ext[6 * b + 3 + i] += sgn * fw[i]; }
    }
  }
  for (int i = 0; i < 6; i++) { d->cacc[i] = 0; d->cfrc_int[i] = 0; }
  if (!(m->disableflags & BMJ_DSBL_GRAVITY)) for (int i = 0; i < 3; i++) d->cacc[3 + i] = -m->gravity[i];
  for (int b = 1; b < nb; b++) {
    int p = m->body_parentid[b];
    real* ca = &d->cacc[6 * b];
    for (int i = 0; i < 6; i++) ca[i] = d->cacc[6 * p + i];
    for (int k = m->body_dofadr[b]; k < m->body_dofadr[b] + m->body_dofnum[b]; k++)
      for (int i = 0; i < 6; i++) ca[i] += d->cdof_dot[6 * k + i] * d->qvel[k] + d->cdof[6 * k + i] * d->qacc[k];
    real t1[6], t2[6], t3[6];
    mul_inert_vec(t1, &d->cinert[10 * b], ca);
    mul_inert_vec(t2, &d->cinert[10 * b], &d->cvel[6 * b]);
    cross_force(t3, &d->cvel[6 * b], t2);
    for (int i = 0; i < 6; i++) d->cfrc_int[6 * b + i] = t1[i] + t3[i] - d->cfrc_ext[6 * b + i];
  }
  for (int b = nb - 1; b > 0; b--) { int p = m->body_parentid[b]; for (int i = 0; i < 6; i++) d->cfrc_int[6 * p + i] += d->cfrc_int[6 * b + i]; }
}

// ------------------------------------------------------------------------------------------------
// acceleration stage
// ------------------------------------------------------------------------------------------------
static void fwd_actuation(const OModel* m, OData* d) {
  int nv = m->nv;
  for (int i = 0; i < nv; i++) d->qfrc_actuator[i] = 0;
  for (int a = 0; a < m->nu; a++) d->actuator_force[a] = 0;
  for (int i = 0; i < m->na; i++) d->act_dot[i] = 0;
  if (m->disableflags & BMJ_DSBL_ACTUATION) return;
  for (int a = 0; a < m->nu; a++) {
    real ctrl = d->ctrl[a];
    if (m->actuator_ctrllimited[a] && !(m->disableflags & BMJ_DSBL_CLAMPCTRL))
      ctrl = clampr(ctrl, m->actuator_ctrlrange[2 * a], m->actuator_ctrlrange[2 * a + 1]);
    real input = ctrl;
    int aa = m->actuator_actadr[a];
    if (m->actuator_dyntype[a] == BMJ_DYN_INTEGRATOR) { d->act_dot[aa] = ctrl; input = d->act[aa]; }
    else if (m->actuator_dyntype[a] == BMJ_DYN_FILTER) {
      real tau = std::max(BMJ_MINVAL, m->actuator_dynprm[a]);
      d->act_dot[aa] = (ctrl - d->act[aa]) / tau; input = d->act[aa];
    }
    const real* gp = &m->actuator_gainprm[3 * a]; const real* bp = &m->actuator_biasprm[3 * a];
    real gain = gp[0];
    if (m->actuator_gaintype[a] == BMJ_GAIN_AFFINE) gain += gp[1] * d->actuator_length[a] + gp[2] * d->actuator_velocity[a];
    real bias = 0;
    if (m->actuator_biastype[a] == BMJ_BIAS_AFFINE) bias = bp[0] + bp[1] * d->actuator_length[a] + bp[2] * d->actuator_velocity[a];
    real force = gain * input + bias;
    if (m->actuator_forcelimited[a]) force = clampr(force, m->actuator_forcerange[2 * a], m->actuator_forcerange[2 * a + 1]);
    d->actuator_force[a] = force;
    for (int i = 0; i < nv; i++) d->qfrc_actuator[i] += d->actuator_moment[a * nv + i] * force;
  }
}

static void fwd_acceleration(const OModel* m, OData* d) {
  int nv = m->nv;
  for (int i = 0; i < nv; i++) d->qfrc_smooth[i] = d->qfrc_passive[i] - d->qfrc_bias[i] + d->qfrc_applied[i] + d->qfrc_actuator[i];
  // Cartesian forces applied at body COMs
  static thread_local vec jp, jr;
  jp.resize(3 * nv); jr.resize(3 * nv);
  for (int b = 1; b < m->nbody; b++) {
    const real* xf = &d->xfrc_applied[6 * b];
    bool any = false; for (int i = 0; i < 6; i++) if (xf[i] != 0) any = true;
    if (!any) continue;
    jac_point(m, d, jp.data(), jr.data(), &d->xipos[3 * b], b);
    for (int i = 0; i < nv; i++) for (int k = 0; k < 3; k++) d->qfrc_smooth[i] += jp[k * nv + i] * xf[k] + jr[k * nv + i] * xf[3 + k];
  }
  chol_solve(d->qacc_smooth.data(), d->L.data(), d->qfrc_smooth.data(), nv);
}

// --- primal (Newton) solver -----------------------------------------------------------------------
struct PrimalCtx {
  const OModel* m; OData* d; int nv, nefc;
  vec jar, Ma, grad, Mgrad, search, Mv, jv, H, LH;
  real cost, gauss;
};

static inline bool row_active(int type, real jar) { return type == BMJ_CNSTR_EQUALITY || jar < 0; }

// cost, force, qfrc_constraint from jar; returns total cost incl. Gauss term
static void constraint_update(PrimalCtx& c) {
  OData* d = c.d; int nv = c.nv;
  real cost = 0;
  for (int r = 0; r < c.nefc; r++) {
    bool act = row_active(d->efc_type[r], c.jar[r]);
    d->efc_state[r] = act;
    d->efc_force[r] = act ? -d->efc_D[r] * c.jar[r] : 0;
    if (act) cost += 0.5 * d->efc_D[r] * c.jar[r] * c.jar[r];
  }
  for (int i = 0; i < nv; i++) {
    real s = 0;
    for (int r = 0; r < c.nefc; r++) s += d->efc_J[r * nv + i] * d->efc_force[r];
    d->qfrc_constraint[i] = s;
  }
  real gauss = 0;
  for (int i = 0; i < nv; i++) gauss += (c.Ma[i] - d->qfrc_smooth[i]) * (d->qacc[i] - d->qacc_smooth[i]);
  c.gauss = 0.5 * gauss;
  c.cost = cost + c.gauss;
}

static void newton_direction(PrimalCtx& c) {
  OData* d = c.d; int nv = c.nv;
  for (int i = 0; i < nv; i++) c.grad[i] = c.Ma[i] - d->qfrc_smooth[i] - d->qfrc_constraint[i];
  for (int i = 0; i < nv * nv; i++) c.H[i] = d->M[i];
  for (int r = 0; r < c.nefc; r++) {
    if (!d->efc_state[r]) continue;
    const real* Jr = &d->efc_J[r * nv]; real D = d->efc_D[r];
    for (int i = 0; i < nv; i++) { if (Jr[i] == 0) continue; real s = D * Jr[i]; for (int j = 0; j <= i; j++) c.H[i * nv + j] += s * Jr[j]; }
  }
  for (int i = 0; i < nv; i++) for (int j = i + 1; j < nv; j++) c.H[i * nv + j] = c.H[j * nv + i];
  chol_factor(c.LH.data(), c.H.data(), nv);
  chol_solve(c.Mgrad.data(), c.LH.data(), c.grad.data(), nv);
  for (int i = 0; i < nv; i++) c.search[i] = -c.Mgrad[i];
}

// derivatives of the 1-D cost along the search direction at step alpha
static void ls_eval(const PrimalCtx& c, real alpha, const real* quadGauss, real* cost, real* d1, real* d2) {
  const OData* d = c.d;
  real q0 = quadGauss[0], q1 = quadGauss[1], q2 = quadGauss[2];
  for (int r = 0; r < c.nefc; r++) {
    real x = c.jar[r] + alpha * c.jv[r];
    if (row_active(d->efc_type[r], x)) {
      real D = d->efc_D[r];
      q0 += 0.5 * D * c.jar[r] * c.jar[r]; q1 += D * c.jar[r] * c.jv[r]; q2 += 0.5 * D * c.jv[r] * c.jv[r];
    }
  }
  *cost = alpha * alpha * q2 + alpha * q1 + q0;
  *d1 = 2 * alpha * q2 + q1;
  *d2 = 2 * q2;
}

// exact line search on a convex piecewise-quadratic: safeguarded Newton with bracketing
static real line_search(PrimalCtx& c) {
  const OModel* m = c.m; OData* d = c.d; int nv = c.nv;
  real snorm = 0;
  for (int i = 0; i < nv; i++) snorm += c.search[i] * c.search[i];
  snorm = std::sqrt(snorm);
  if (snorm < BMJ_MINVAL) return 0;
  real scale = m->meaninertia * std::max(1, nv);
  real gtol = m->tolerance * m->ls_tolerance * snorm * scale;
  // Mv, jv
  for (int i = 0; i < nv; i++) { real s = 0; for (int j = 0; j < nv; j++) s += d->M[i * nv + j] * c.search[j]; c.Mv[i] = s; }
  for (int r = 0; r < c.nefc; r++) { real s = 0; for (int i = 0; i < nv; i++) s += d->efc_J[r * nv + i] * c.search[i]; c.jv[r] = s; }
  real quadGauss[3] = {c.gauss, 0, 0};
  for (int i = 0; i < nv; i++) { quadGauss[1] += c.search[i] * (c.Ma[i] - d->qfrc_smooth[i]); quadGauss[2] += 0.5 * c.search[i] * c.Mv[i]; }
  real cost0, d1, d2;
  ls_eval(c, 0, quadGauss, &cost0, &d1, &d2);
  if (d1 >= 0 || d2 <= 0) return 0;            // not a descent direction
  real lo = 0, dlo = d1;                        // derivative < 0 at lo
  real hi = 0, dhi = 0; bool have_hi = false;
  real alpha = -d1 / d2, best = 0;
  for (int it = 0; it < m->ls_iterations; it++) {
    real cst, e1, e2;
    ls_eval(c, alpha, quadGauss, &cst, &e1, &e2);
    best = alpha;
    if (std::fabs(e1) < gtol) break;
    if (e1 < 0) { lo = alpha; dlo = e1; } else { hi = alpha; dhi = e1; have_hi = true; }
    real next = alpha - e1 / e2;                // Newton step on the local quadratic
    if (have_hi) { if (!(next > lo && next < hi)) next = lo + (hi - lo) * (-dlo) / (dhi - dlo); }
    else if (next <= lo) next = 2 * alpha + 1e-12;
    if (next == alpha) break;
    alpha = next;
  }
  return best;
}

static void solve_newton(const OModel* m, OData* d) {
  int nv = m->nv, nefc = d->nefc;
  static thread_local PrimalCtx c; c.m = m; c.d = d; c.nv = nv; c.nefc = nefc;
  c.jar.resize(nefc); c.Ma.resize(nv); c.grad.resize(nv); c.Mgrad.resize(nv); c.search.resize(nv); c.Mv.resize(nv);
  c.jv.resize(nefc); c.H.resize(nv * nv); c.LH.resize(nv * nv);
  auto compute_Ma_jar = [&]() {
    for (int i = 0; i < nv; i++) { real s = 0; for (int j = 0; j < nv; j++) s += d->M[i * nv + j] * d->qacc[j]; c.Ma[i] = s; }
    for (int r = 0; r < nefc; r++) { real s = 0; for (int i = 0; i < nv; i++) s += d->efc_J[r * nv + i] * d->qacc[i]; c.jar[r] = s - d->efc_aref[r]; }
  };
  // warm start: keep qacc_warmstart only if it has lower cost than the unconstrained acceleration
  if (!(m->disableflags & BMJ_DSBL_WARMSTART)) {
    d->qacc = d->qacc_warmstart; compute_Ma_jar(); constraint_update(c);
    real cost_warm = c.cost;
    d->qacc = d->qacc_smooth; compute_Ma_jar(); constraint_update(c);
    if (cost_warm < c.cost) { d->qacc = d->qacc_warmstart; compute_Ma_jar(); constraint_update(c); }
  } else { d->qacc = d->qacc_smooth; compute_Ma_jar(); constraint_update(c); }
  newton_direction(c);
  real scale = 1 / (m->meaninertia * std::max(1, nv));
  int iter = 0;
  while (iter < m->iterations) {
    real alpha = line_search(c);
    if (alpha == 0) break;
    for (int i = 0; i < nv; i++) { d->qacc[i] += alpha * c.search[i]; c.Ma[i] += alpha * c.Mv[i]; }
    for (int r = 0; r < nefc; r++) c.jar[r] += alpha * c.jv[r];
    real oldcost = c.cost;
    constraint_update(c);
    newton_direction(c);
    iter++;
    real improvement = scale * (oldcost - c.cost);
    real gn = 0; for (int i = 0; i < nv; i++) gn += c.grad[i] * c.grad[i];
    real gradient = scale * std::sqrt(gn);
    if (improvement < m->tolerance || gradient < m->tolerance) break;
  }
  d->solver_niter = iter;
}

// --- dual solver: projected Gauss-Seidel (mj_solPGS) ------------------------------------------------
// min over force of  1/2 f' A f + f' b,  A = J M^-1 J' + diag(R),  b = J qacc_smooth - aref,  with f_i >= 0 for every
// row that is not an equality (limits, pyramidal contact edges). One sweep updates the rows in order,
// f_i <- project(f_i - (A_i f + b_i) / A_ii); the sweep's cost decrease, scaled like the Newton solver's, stops the
// iteration. Warm start: the forces the previous acceleration implies (primal constraint update at qacc_warmstart),
// kept only if their dual cost is negative (zero forces cost 0). The testing humanoid of the reference asks for this
// solver (dm_control/mujoco/testing/assets/humanoid.xml:9, solver="PGS" iterations="50").
static void solve_pgs(const OModel* m, OData* d) {
  const int nv = m->nv, nefc = d->nefc;
  static thread_local vec A, B, X, f;
  A.assign((size_t)nefc * nefc, 0); B.assign(nefc, 0); X.assign((size_t)nefc * nv, 0); f.assign(nefc, 0);
  for (int r = 0; r < nefc; r++) chol_solve(&X[(size_t)r * nv], d->L.data(), &d->efc_J[(size_t)r * nv], nv);    // M^-1 J_r'
  for (int i = 0; i < nefc; i++) {
    for (int j = 0; j < nefc; j++) {
      real s = 0;
      for (int k = 0; k < nv; k++) s += d->efc_J[(size_t)i * nv + k] * X[(size_t)j * nv + k];
      A[(size_t)i * nefc + j] = s;
    }
    A[(size_t)i * nefc + i] += d->efc_R[i];
    real s = 0;
    for (int k = 0; k < nv; k++) s += d->efc_J[(size_t)i * nv + k] * d->qacc_smooth[k];
    B[i] = s - d->efc_aref[i];
  }
  auto dual_cost = [&]() {
    real c = 0;
    for (int i = 0; i < nefc; i++) { real s = 0; for (int j = 0; j < nefc; j++) s += A[(size_t)i * nefc + j] * f[j]; c += f[i] * (0.5 * s + B[i]); }
    return c;
  };
  if (!(m->disableflags & BMJ_DSBL_WARMSTART)) {
    for (int r = 0; r < nefc; r++) {
      real s = 0;
      for (int k = 0; k < nv; k++) s += d->efc_J[(size_t)r * nv + k] * d->qacc_warmstart[k];
      const real jar = s - d->efc_aref[r];
      f[r] = row_active(d->efc_type[r], jar) ? -d->efc_D[r] * jar : 0;
    }
    if (dual_cost() > 0) f.assign(nefc, 0);
  }
  const real scale = 1 / (m->meaninertia * std::max(1, nv));
  int iter = 0;
  while (iter < m->iterations) {
    real improvement = 0;
    for (int i = 0; i < nefc; i++) {
      real res = B[i];
      for (int j = 0; j < nefc; j++) res += A[(size_t)i * nefc + j] * f[j];
      const real old = f[i];
      f[i] -= res / A[(size_t)i * nefc + i];
      if (d->efc_type[i] != BMJ_CNSTR_EQUALITY && f[i] < 0) f[i] = 0;
      const real delta = f[i] - old;
      improvement -= 0.5 * delta * delta * A[(size_t)i * nefc + i] + delta * res;
    }
    iter++;
    if (improvement * scale < m->tolerance) break;
  }
  for (int r = 0; r < nefc; r++) { d->efc_force[r] = f[r]; d->efc_state[r] = f[r] != 0 || d->efc_type[r] == BMJ_CNSTR_EQUALITY; }
  for (int i = 0; i < nv; i++) {
    real s = 0;
    for (int r = 0; r < nefc; r++) s += d->efc_J[(size_t)r * nv + i] * f[r];
    d->qfrc_constraint[i] = s;
  }
  vec rhs(nv);
  for (int i = 0; i < nv; i++) rhs[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i];
  chol_solve(d->qacc.data(), d->L.data(), rhs.data(), nv);
  d->solver_niter = iter;
}

static void fwd_constraint(const OModel* m, OData* d) {
  int nv = m->nv;
  if (d->nefc == 0) {
    d->qacc = d->qacc_smooth;
    for (int i = 0; i < nv; i++) d->qfrc_constraint[i] = 0;
    d->solver_niter = 0;
    return;
  }
  if (m->solver == BMJ_SOL_PGS) solve_pgs(m, d);
  else solve_newton(m, d);
}

// ------------------------------------------------------------------------------------------------
// checks, integration, top level
// ------------------------------------------------------------------------------------------------
static inline bool bad(real x) { return std::isnan(x) || x > BMJ_MAXVAL || x < -BMJ_MAXVAL; }
static void fwd_position(const OModel* m, OData* d);
static void forward_skip(const OModel* m, OData* d, int skipsensor);

static bool check_vec(const OModel* m, OData* d, const vec& v, int warn) {
  for (size_t i = 0; i < v.size(); i++) if (bad(v[i])) {
    d->warning[warn]++;
    // MuJoCo resets the offending mjData in place and keeps the warning counters
    int w[BMJ_NWARNING]; memcpy(w, d->warning, sizeof(w));
    reset_data(m, d, -1);
    memcpy(d->warning, w, sizeof(w));
    return true;
  }
  return false;
}

static void fwd_position(const OModel* m, OData* d) {
  kinematics(m, d);
  com_pos(m, d);
  tendon_and_transmission(m, d);
  crb_and_factor(m, d);
  collision(m, d);
  make_constraint(m, d);
}

static void forward_skip(const OModel* m, OData* d, int skipsensor) {
  fwd_position(m, d);
  if (!skipsensor) sensors(m, d, 1);
  fwd_velocity(m, d);
  make_impedance(m, d);     // reference acceleration needs J*qvel (computed lazily in MuJoCo's makeConstraint chain)
  if (!skipsensor) sensors(m, d, 2);
  fwd_actuation(m, d);
  fwd_acceleration(m, d);
  fwd_constraint(m, d);
  if (!skipsensor) { rne_post_constraint(m, d); sensors(m, d, 3); }
}

static void integrate_pos(const OModel* m, real* qpos, const real* qvel, real h) {
  for (int j = 0; j < m->njnt; j++) {
    int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    switch (m->jnt_type[j]) {
      case BMJ_JNT_FREE:
        for (int i = 0; i < 3; i++) qpos[qa + i] += h * qvel[da + i];
        quat_integrate(qpos + qa + 3, qvel + da + 3, h);
        break;
      case BMJ_JNT_BALL: quat_integrate(qpos + qa, qvel + da, h); break;
      default: qpos[qa] += h * qvel[da];
    }
  }
}

static void advance(const OModel* m, OData* d, const real* act_dot, const real* qacc, const real* qvel_for_pos) {
  real h = m->timestep;
  for (int a = 0; a < m->nu; a++) {
    int aa = m->actuator_actadr[a];
    if (aa < 0) continue;
    d->act[aa] += h * act_dot[aa];
    if (m->actuator_actlimited[a]) d->act[aa] = clampr(d->act[aa], m->actuator_actrange[2 * a], m->actuator_actrange[2 * a + 1]);
  }
  for (int i = 0; i < m->nv; i++) d->qvel[i] += h * qacc[i];
  integrate_pos(m, d->qpos.data(), qvel_for_pos ? qvel_for_pos : d->qvel.data(), h);
  d->time += h;
  d->qacc_warmstart = d->qacc;
}

static void euler(const OModel* m, OData* d) {
  int nv = m->nv; real h = m->timestep;
  bool damped = false;
  for (int i = 0; i < nv; i++) if (m->dof_damping[i] > 0) damped = true;
  if (damped && !(m->disableflags & BMJ_DSBL_EULERDAMP)) {
    static thread_local vec A, rhs, qacc;
    A = d->M; rhs.resize(nv); qacc.resize(nv);
    for (int i = 0; i < nv; i++) { A[i * nv + i] += h * m->dof_damping[i]; rhs[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i]; }
    chol_factor(d->Lh.data(), A.data(), nv);
    chol_solve(qacc.data(), d->Lh.data(), rhs.data(), nv);
    advance(m, d, d->act_dot.data(), qacc.data(), nullptr);
  } else advance(m, d, d->act_dot.data(), d->qacc.data(), nullptr);
}

static void runge_kutta4(const OModel* m, OData* d) {
  static const real A[9] = {0.5, 0, 0, 0, 0.5, 0, 0, 0, 1}, B[4] = {1.0/6, 1.0/3, 1.0/3, 1.0/6}, C[3] = {0.5, 0.5, 1.0};
  int nq = m->nq, nv = m->nv, na = m->na; real h = m->timestep, t0 = d->time;
  vec X0q(d->qpos), X0v(d->qvel), X0a(d->act);
  vec Fv[4], Fa[4], Fd[4];
  Fv[0] = d->qvel; Fa[0] = d->qacc; Fd[0] = d->act_dot;
  for (int i = 1; i < 4; i++) {
    vec dv(nv, 0.0), da(nv, 0.0), dd(na, 0.0);
    for (int j = 0; j < i; j++) {
      real a = A[(i - 1) * 3 + j];
      for (int k = 0; k < nv; k++) { dv[k] += a * Fv[j][k]; da[k] += a * Fa[j][k]; }
      for (int k = 0; k < na; k++) dd[k] += a * Fd[j][k];
    }
    d->qpos = X0q; integrate_pos(m, d->qpos.data(), dv.data(), h);
    for (int k = 0; k < nv; k++) d->qvel[k] = X0v[k] + h * da[k];
    for (int k = 0; k < na; k++) d->act[k] = X0a[k] + h * dd[k];
    d->time = t0 + h * C[i - 1];
    forward_skip(m, d, 1);
    Fv[i] = d->qvel; Fa[i] = d->qacc; Fd[i] = d->act_dot;
  }
  vec dv(nv, 0.0), da(nv, 0.0), dd(na, 0.0);
  for (int j = 0; j < 4; j++) {
    for (int k = 0; k < nv; k++) { dv[k] += B[j] * Fv[j][k]; da[k] += B[j] * Fa[j][k]; }
    for (int k = 0; k < na; k++) dd[k] += B[j] * Fd[j][k];
  }
  d->qpos = X0q; d->qvel = X0v; d->act = X0a; d->time = t0;
  advance(m, d, dd.data(), da.data(), dv.data());
  (void)nq;
}

static void check_ctrl(const OModel* m, OData* d) {
  for (int a = 0; a < m->nu; a++) if (bad(d->ctrl[a])) { d->warning[BMJ_WARN_BADCTRL]++; for (int k = 0; k < m->nu; k++) d->ctrl[k] = 0; break; }
}

static void step1(const OModel* m, OData* d) {
  check_vec(m, d, d->qpos, BMJ_WARN_BADQPOS);
  check_vec(m, d, d->qvel, BMJ_WARN_BADQVEL);
  fwd_position(m, d);
  sensors(m, d, 1);
  fwd_velocity(m, d);
  make_impedance(m, d);
  sensors(m, d, 2);
}

static void step2(const OModel* m, OData* d) {
  check_ctrl(m, d);
  fwd_actuation(m, d);
  fwd_acceleration(m, d);
  fwd_constraint(m, d);
  rne_post_constraint(m, d);
  sensors(m, d, 3);
  check_vec(m, d, d->qacc, BMJ_WARN_BADQACC);
  euler(m, d);   // step2 cannot do RK4 (the reference special-cases it, engine.py:155-160)
}

static void step(const OModel* m, OData* d) {
  check_vec(m, d, d->qpos, BMJ_WARN_BADQPOS);
  check_vec(m, d, d->qvel, BMJ_WARN_BADQVEL);
  check_ctrl(m, d);
  forward_skip(m, d, 0);
  check_vec(m, d, d->qacc, BMJ_WARN_BADQACC);
  if (m->integrator == BMJ_INT_RK4) runge_kutta4(m, d); else euler(m, d);
}

// ------------------------------------------------------------------------------------------------
// C API (ctypes)
// ------------------------------------------------------------------------------------------------
extern "C" {

void* bmjo_model_create(const int* idata, const double* rdata) { return model_from_blob(idata, rdata); }
void bmjo_model_destroy(void* m) { delete (OModel*)m; }
void* bmjo_data_create(void* m) { OData* d = data_create((OModel*)m); reset_data((OModel*)m, d, -1); return d; }
void bmjo_data_destroy(void* d) { delete (OData*)d; }
void bmjo_reset(void* m, void* d, int key) { reset_data((OModel*)m, (OData*)d, key); }
void bmjo_forward(void* m, void* d) { check_ctrl((OModel*)m, (OData*)d); forward_skip((OModel*)m, (OData*)d, 0); }
void bmjo_step(void* m, void* d, int n) { for (int i = 0; i < n; i++) step((OModel*)m, (OData*)d); }
void bmjo_step1(void* m, void* d) { step1((OModel*)m, (OData*)d); }
void bmjo_step2(void* m, void* d) { step2((OModel*)m, (OData*)d); }
void bmjo_subtree_vel(void* m, void* d) { subtree_vel((OModel*)m, (OData*)d); }
void bmjo_set_disableflags(void* m, int flags) { ((OModel*)m)->disableflags = flags; }
int bmjo_get_disableflags(void* m) { return ((OModel*)m)->disableflags; }

// field access by name -> pointer + length (doubles)
double* bmjo_field(void* dv, const char* name, int* n) {
  OData* d = (OData*)dv;
#define F(x) if (!strcmp(name, #x)) { *n = (int)d->x.size(); return d->x.data(); }
  F(qpos) F(qvel) F(act) F(ctrl) F(qacc) F(qacc_warmstart) F(act_dot) F(qfrc_applied) F(xfrc_applied)
  F(xpos) F(xquat) F(xmat) F(xipos) F(ximat) F(xanchor) F(xaxis) F(geom_xpos) F(geom_xmat) F(site_xpos) F(site_xmat)
  F(subtree_com) F(cinert) F(crb) F(cdof) F(cdof_dot) F(cvel) F(cacc) F(cfrc_int) F(cfrc_ext) F(subtree_linvel)
  F(M) F(L) F(ten_length) F(ten_J) F(ten_velocity) F(actuator_length) F(actuator_velocity) F(actuator_moment)
  F(actuator_force) F(qfrc_bias) F(qfrc_passive) F(qfrc_actuator) F(qfrc_smooth) F(qacc_smooth) F(qfrc_constraint)
  F(efc_J) F(efc_pos) F(efc_margin) F(efc_D) F(efc_R) F(efc_aref) F(efc_force) F(efc_vel) F(efc_diagApprox) F(sensordata)
#undef F
  if (!strcmp(name, "time")) { *n = 1; return &d->time; }
  *n = -1; return nullptr;
}
int bmjo_ncon(void* d) { return ((OData*)d)->ncon; }
int bmjo_nefc(void* d) { return ((OData*)d)->nefc; }
int bmjo_solver_niter(void* d) { return ((OData*)d)->solver_niter; }
int* bmjo_warning(void* d) { return ((OData*)d)->warning; }
int* bmjo_efc_int(void* dv, const char* name, int* n) {
  OData* d = (OData*)dv;
  if (!strcmp(name, "efc_type")) { *n = (int)d->efc_type.size(); return d->efc_type.data(); }
  if (!strcmp(name, "efc_id")) { *n = (int)d->efc_id.size(); return d->efc_id.data(); }
  if (!strcmp(name, "efc_state")) { *n = (int)d->efc_state.size(); return d->efc_state.data(); }
  *n = -1; return nullptr;
}
// contact i -> out[0..] = dist, pos[3], frame[9], includemargin, friction[5], solref[2], solimp[5], dim, geom1, geom2, efc_address
void bmjo_contact(void* dv, int i, double* out) {
  OContact* c = &((OData*)dv)->contact[i];
  int k = 0;
  out[k++] = c->dist; for (int j = 0; j < 3; j++) out[k++] = c->pos[j]; for (int j = 0; j < 9; j++) out[k++] = c->frame[j];
  out[k++] = c->includemargin; for (int j = 0; j < 5; j++) out[k++] = c->friction[j];
  for (int j = 0; j < 2; j++) out[k++] = c->solref[j];
  for (int j = 0; j < 5; j++) out[k++] = c->solimp[j];
  out[k++] = c->dim; out[k++] = c->geom1; out[k++] = c->geom2; out[k++] = c->efc_address;
}

// One geom pair through the narrow phase (test hook, tests/test_convex_pairs.py): geoms given as type, pos[3], mat[9]
// (row-major), size[3]; the pair is type-sorted as the compiler would emit it. out[k*10 ..] = dist, pos[3], normal[3],
// tangent hint[3]. Returns the contact count (-1: unsupported pair).
int bmjo_narrowphase(int t1, const double* p1, const double* m1, const double* s1, int t2, const double* p2, const double* m2,
                     const double* s2, double margin, double* out) {
  RawCon raw[8];
  int n = narrowphase(raw, t1, t2, margin, p1, m1, s1, p2, m2, s2);
  for (int k = 0; k < n; k++) {
    out[10 * k] = raw[k].dist;
    for (int i = 0; i < 3; i++) { out[10 * k + 1 + i] = raw[k].pos[i]; out[10 * k + 4 + i] = raw[k].normal[i]; out[10 * k + 7 + i] = raw[k].tangent[i]; }
  }
  return n;
}

// Batched rollout helper for the CPU baseline: env e uses state rows e of the [B, n] arrays.
// One physics "control step" in the reference's legacy ordering (engine.py:147-162):
//   step2, (step1+step2)*(nstep-1), step1   (Euler)   |   step*nstep, step1   (RK4)
void bmjo_control_step(void* mv, void* dv, int nstep) {
  OModel* m = (OModel*)mv; OData* d = (OData*)dv;
  if (m->integrator != BMJ_INT_RK4) { step2(m, d); for (int i = 1; i < nstep; i++) step(m, d); }
  else for (int i = 0; i < nstep; i++) step(m, d);
  step1(m, d);
}

// Rollout helper for the CPU baseline: `nsteps` control steps with a pre-generated action tape [nsteps, nu],
// entirely in C so that Python threads only wait (ctypes drops the GIL for the whole call).
void bmjo_rollout(void* mv, void* dv, const double* tape, int nsteps, int nsub) {
  OModel* m = (OModel*)mv; OData* d = (OData*)dv;
  for (int k = 0; k < nsteps; k++) {
    for (int a = 0; a < m->nu; a++) d->ctrl[a] = tape[(size_t)k * m->nu + a];
    bmjo_control_step(mv, dv, nsub);
  }
}

}  // extern "C"
