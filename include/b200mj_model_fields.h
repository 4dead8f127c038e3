/* b200mj_model_fields.h — the compiled-model blob layout (single source of truth).
 *
 * A compiled model travels across the C ABI as two flat arrays (int32 `idata`, float64 `rdata`)
 * plus a directory: for field number k (position in the X-macro list below) the int pair
 * idata[2*k], idata[2*k+1] = (offset, length) into idata (int fields) or rdata (real fields).
 * The Python MJCF compiler (dm_control_b200/model.py) parses THIS FILE to learn the order, so the
 * host packer, the CUDA engine (dm_control_b200/csrc) and the CPU oracle (oracle/) cannot drift.
 *
 * Field meaning follows MuJoCo's mjModel naming wherever a field of that name exists there
 * (the reference reaches them as `physics.model.<field>`, dm_control/mujoco/wrapper/core.py:253-432,
 * shape table consumed at dm_control/mujoco/index.py:177-297). Extra fields (level_*, pair_*, sizes)
 * are this engine's own compile-time tables.
 *
 * BMJ_I(name)  int32 array      BMJ_R(name)  float64 array
 */
#ifndef B200MJ_MODEL_FIELDS_H_
#define B200MJ_MODEL_FIELDS_H_

#define B200MJ_MODEL_FIELDS(BMJ_I, BMJ_R)                                                        \
  /* ---- sizes[..]: see enum b200mj_size below ---- */                                          \
  BMJ_I(sizes)                                                                                   \
  /* ---- options (opt_real: see enum b200mj_optr; opt_int: enum b200mj_opti) ---- */            \
  BMJ_R(opt_real)                                                                                \
  BMJ_I(opt_int)                                                                                 \
  /* ---- bodies ---- */                                                                         \
  BMJ_I(body_parentid) BMJ_I(body_rootid) BMJ_I(body_weldid) BMJ_I(body_jntnum)                  \
  BMJ_I(body_jntadr) BMJ_I(body_dofnum) BMJ_I(body_dofadr) BMJ_I(body_geomnum) BMJ_I(body_geomadr)\
  BMJ_R(body_pos) BMJ_R(body_quat) BMJ_R(body_ipos) BMJ_R(body_iquat) BMJ_R(body_mass)           \
  BMJ_R(body_subtreemass) BMJ_R(body_inertia) BMJ_R(body_invweight0)                             \
  /* tree levels: bodies sorted by depth; level_adr has nlevel+1 entries into level_body */      \
  BMJ_I(level_adr) BMJ_I(level_body)                                                             \
  /* body_dofmask[2*b], [2*b+1]: bit i (lo word: dofs 0-31, hi word: 32-63) set iff dof i moves body b */ \
  BMJ_I(body_dofmask)                                                             \
  /* ---- joints ---- */                                                                         \
  BMJ_I(jnt_type) BMJ_I(jnt_qposadr) BMJ_I(jnt_dofadr) BMJ_I(jnt_bodyid) BMJ_I(jnt_limited)      \
  BMJ_R(jnt_pos) BMJ_R(jnt_axis) BMJ_R(jnt_stiffness) BMJ_R(jnt_range) BMJ_R(jnt_margin)         \
  BMJ_R(jnt_solref) BMJ_R(jnt_solimp)                                                            \
  BMJ_R(qpos0) BMJ_R(qpos_spring)                                                                \
  /* ---- dofs ---- */                                                                           \
  BMJ_I(dof_bodyid) BMJ_I(dof_jntid) BMJ_I(dof_parentid)                                         \
  BMJ_R(dof_armature) BMJ_R(dof_damping) BMJ_R(dof_invweight0) BMJ_R(dof_frictionloss)           \
  BMJ_R(dof_solref) BMJ_R(dof_solimp)                                                            \
  /* ---- geoms ---- */                                                                          \
  BMJ_I(geom_type) BMJ_I(geom_bodyid) BMJ_I(geom_condim) BMJ_I(geom_contype)                     \
  BMJ_I(geom_conaffinity) BMJ_I(geom_priority)                                                   \
  BMJ_R(geom_size) BMJ_R(geom_pos) BMJ_R(geom_quat) BMJ_R(geom_rbound) BMJ_R(geom_friction)      \
  BMJ_R(geom_solmix) BMJ_R(geom_solref) BMJ_R(geom_solimp) BMJ_R(geom_margin) BMJ_R(geom_gap)    \
  /* candidate geom pairs that survive the static filters, in emission order (type-swapped) */  \
  BMJ_I(pair_geom1) BMJ_I(pair_geom2)                                                            \
  /* ---- sites ---- */                                                                          \
  BMJ_I(site_bodyid) BMJ_I(site_type) BMJ_R(site_pos) BMJ_R(site_quat) BMJ_R(site_size)          \
  /* ---- actuators ---- */                                                                      \
  BMJ_I(actuator_trntype) BMJ_I(actuator_trnid) BMJ_I(actuator_dyntype) BMJ_I(actuator_gaintype) \
  BMJ_I(actuator_biastype) BMJ_I(actuator_ctrllimited) BMJ_I(actuator_forcelimited)              \
  BMJ_I(actuator_actlimited) BMJ_I(actuator_actadr)                                              \
  BMJ_R(actuator_gear) BMJ_R(actuator_gainprm) BMJ_R(actuator_biasprm) BMJ_R(actuator_dynprm)    \
  BMJ_R(actuator_ctrlrange) BMJ_R(actuator_forcerange) BMJ_R(actuator_actrange)                  \
  /* ---- fixed tendons ---- */                                                                  \
  BMJ_I(tendon_adr) BMJ_I(tendon_num) BMJ_I(wrap_objid) BMJ_R(wrap_prm)                          \
  BMJ_R(tendon_length0) BMJ_R(tendon_invweight0)                                                 \
  /* ---- equality constraints ---- */                                                           \
  BMJ_I(eq_type) BMJ_I(eq_obj1id) BMJ_I(eq_obj2id) BMJ_I(eq_active0)                             \
  BMJ_R(eq_data) BMJ_R(eq_solref) BMJ_R(eq_solimp)                                               \
  /* ---- sensors ---- */                                                                        \
  BMJ_I(sensor_type) BMJ_I(sensor_objtype) BMJ_I(sensor_objid) BMJ_I(sensor_reftype)             \
  BMJ_I(sensor_refid) BMJ_I(sensor_dim) BMJ_I(sensor_adr) BMJ_I(sensor_needstage)                \
  /* ---- keyframes (qpos only) ---- */                                                          \
  BMJ_R(key_qpos)

/* indices into `sizes` */
enum b200mj_size {
  BMJ_NQ = 0, BMJ_NV, BMJ_NU, BMJ_NA, BMJ_NBODY, BMJ_NJNT, BMJ_NGEOM, BMJ_NSITE, BMJ_NTENDON,
  BMJ_NWRAP, BMJ_NEQ, BMJ_NSENSOR, BMJ_NSENSORDATA, BMJ_NPAIR, BMJ_NLEVEL, BMJ_NKEY,
  BMJ_NCONMAX,  /* per-env contact capacity (CONTACTFULL warning when exceeded) */
  BMJ_NJMAX,    /* per-env constraint-row capacity (CNSTRFULL warning when exceeded) */
  BMJ_NSIZES
};

/* indices into `opt_real` */
enum b200mj_optr {
  BMJ_OPT_TIMESTEP = 0, BMJ_OPT_GRAVITY_X, BMJ_OPT_GRAVITY_Y, BMJ_OPT_GRAVITY_Z,
  BMJ_OPT_TOLERANCE, BMJ_OPT_LS_TOLERANCE, BMJ_OPT_IMPRATIO, BMJ_OPT_MEANINERTIA,
  BMJ_OPT_O_MARGIN, BMJ_NOPTR
};

/* indices into `opt_int` */
enum b200mj_opti {
  BMJ_OPT_INTEGRATOR = 0, BMJ_OPT_SOLVER, BMJ_OPT_ITERATIONS, BMJ_OPT_LS_ITERATIONS,
  BMJ_OPT_DISABLEFLAGS, BMJ_OPT_CONE, BMJ_OPT_ENABLEFLAGS, BMJ_NOPTI
};

/* enums mirrored from MuJoCo's mjmodel.h (values are MuJoCo's so `physics.model.jnt_type` etc.
 * read the same through the facade; reference use: dm_control/mujoco/wrapper/mjbindings enums,
 * dm_control/mujoco/engine.py:50-60) */
enum { BMJ_JNT_FREE = 0, BMJ_JNT_BALL = 1, BMJ_JNT_SLIDE = 2, BMJ_JNT_HINGE = 3 };
enum { BMJ_GEOM_PLANE = 0, BMJ_GEOM_HFIELD = 1, BMJ_GEOM_SPHERE = 2, BMJ_GEOM_CAPSULE = 3,
       BMJ_GEOM_ELLIPSOID = 4, BMJ_GEOM_CYLINDER = 5, BMJ_GEOM_BOX = 6, BMJ_GEOM_MESH = 7 };
enum { BMJ_INT_EULER = 0, BMJ_INT_RK4 = 1, BMJ_INT_IMPLICIT = 2, BMJ_INT_IMPLICITFAST = 3 };
enum { BMJ_SOL_PGS = 0, BMJ_SOL_CG = 1, BMJ_SOL_NEWTON = 2 };
enum { BMJ_TRN_JOINT = 0, BMJ_TRN_JOINTINPARENT = 1, BMJ_TRN_SLIDERCRANK = 2, BMJ_TRN_TENDON = 3,
       BMJ_TRN_SITE = 4 };
enum { BMJ_DYN_NONE = 0, BMJ_DYN_INTEGRATOR = 1, BMJ_DYN_FILTER = 2, BMJ_DYN_FILTEREXACT = 3 };
enum { BMJ_GAIN_FIXED = 0, BMJ_GAIN_AFFINE = 1 };
enum { BMJ_BIAS_NONE = 0, BMJ_BIAS_AFFINE = 1 };
enum { BMJ_EQ_CONNECT = 0, BMJ_EQ_WELD = 1, BMJ_EQ_JOINT = 2, BMJ_EQ_TENDON = 3 };
/* mjtDisableBit */
enum { BMJ_DSBL_CONSTRAINT = 1 << 0, BMJ_DSBL_EQUALITY = 1 << 1, BMJ_DSBL_FRICTIONLOSS = 1 << 2,
       BMJ_DSBL_LIMIT = 1 << 3, BMJ_DSBL_CONTACT = 1 << 4, BMJ_DSBL_PASSIVE = 1 << 5,
       BMJ_DSBL_GRAVITY = 1 << 6, BMJ_DSBL_CLAMPCTRL = 1 << 7, BMJ_DSBL_WARMSTART = 1 << 8,
       BMJ_DSBL_FILTERPARENT = 1 << 9, BMJ_DSBL_ACTUATION = 1 << 10, BMJ_DSBL_REFSAFE = 1 << 11,
       BMJ_DSBL_SENSOR = 1 << 12, BMJ_DSBL_MIDPHASE = 1 << 13, BMJ_DSBL_EULERDAMP = 1 << 14 };
/* mjtSensor subset (MuJoCo numbering) */
enum { BMJ_SENS_TOUCH = 0, BMJ_SENS_ACCELEROMETER = 1, BMJ_SENS_VELOCIMETER = 2, BMJ_SENS_GYRO = 3,
       BMJ_SENS_FORCE = 4, BMJ_SENS_TORQUE = 5, BMJ_SENS_JOINTPOS = 8, BMJ_SENS_JOINTVEL = 9,
       BMJ_SENS_ACTUATORFRC = 14, BMJ_SENS_FRAMEPOS = 25, BMJ_SENS_SUBTREECOM = 34,
       BMJ_SENS_SUBTREELINVEL = 35, BMJ_SENS_SUBTREEANGMOM = 36 };
/* mjtObj subset */
enum { BMJ_OBJ_UNKNOWN = 0, BMJ_OBJ_BODY = 1, BMJ_OBJ_XBODY = 2, BMJ_OBJ_JOINT = 3, BMJ_OBJ_GEOM = 5,
       BMJ_OBJ_SITE = 6, BMJ_OBJ_ACTUATOR = 19 };
/* mjtConstraint */
enum { BMJ_CNSTR_EQUALITY = 0, BMJ_CNSTR_FRICTION_DOF = 1, BMJ_CNSTR_FRICTION_TENDON = 2,
       BMJ_CNSTR_LIMIT_JOINT = 3, BMJ_CNSTR_LIMIT_TENDON = 4, BMJ_CNSTR_CONTACT_FRICTIONLESS = 5,
       BMJ_CNSTR_CONTACT_PYRAMIDAL = 6, BMJ_CNSTR_CONTACT_ELLIPTIC = 7 };
/* mjtWarning (bit positions in the per-env warning word) */
enum { BMJ_WARN_INERTIA = 0, BMJ_WARN_CONTACTFULL = 1, BMJ_WARN_CNSTRFULL = 2, BMJ_WARN_VGEOMFULL = 3,
       BMJ_WARN_BADQPOS = 4, BMJ_WARN_BADQVEL = 5, BMJ_WARN_BADQACC = 6, BMJ_WARN_BADCTRL = 7,
       BMJ_NWARNING = 8 };

#define BMJ_MINVAL 1e-15
#define BMJ_MAXVAL 1e10
#define BMJ_MINIMP 0.0001
#define BMJ_MAXIMP 0.9999

#endif /* B200MJ_MODEL_FIELDS_H_ */
