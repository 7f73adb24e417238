/* b200sim model blob: on-disk/in-memory layout of the constant tables produced by
 * gymnasium_robotics_b200/mjcf.py (the stand-in for mujoco.MjModel.from_xml_path, reference
 * gymnasium_robotics/envs/robot_env.py:293).  Plain C, header-only reader; used by the CUDA
 * library (csrc/) and by the CPU oracle (oracle/) -- it is a data format, not an algorithm.
 *
 * Blob = { u32 magic, u32 version, u32 nentries, u32 pad } + nentries * { char name[32]; u32 dtype;
 * u32 count; u64 offset } + payload.  dtype 0 = int32, 1 = float64, 2 = bytes (JSON name tables).
 * Offsets are relative to the start of the payload; float64 arrays are 8-byte aligned.
 * Optional entries (absent = empty): geom_hull [ngeom][2] = (first vertex, count) into hull_vert [n][3], the reduced convex hull of a
 * MESH geom in its geom frame (support-map narrow phase); models compiled without `mesh_hull` carry meshes as BOX proxies instead.
 */
#ifndef B200SIM_MODEL_H
#define B200SIM_MODEL_H

#include <stdint.h>
#include <string.h>
#include <stddef.h>

#define B200M_MAGIC 0x4D303242u
#define B200M_VERSION 5u

enum { B200_JNT_FREE = 0, B200_JNT_BALL = 1, B200_JNT_SLIDE = 2, B200_JNT_HINGE = 3 };
enum { B200_GEOM_PLANE = 0, B200_GEOM_HFIELD = 1, B200_GEOM_SPHERE = 2, B200_GEOM_CAPSULE = 3,
       B200_GEOM_ELLIPSOID = 4, B200_GEOM_CYLINDER = 5, B200_GEOM_BOX = 6, B200_GEOM_MESH = 7 };
enum { B200_EQ_CONNECT = 0, B200_EQ_WELD = 1, B200_EQ_JOINT = 2 };
enum { B200_INT_EULER = 0, B200_INT_RK4 = 1 };

/* indices into `sizes` */
enum { B200_NBODY = 0, B200_NJNT, B200_NQ, B200_NV, B200_NU, B200_NGEOM, B200_NSITE, B200_NMOCAP, B200_NEQ,
       B200_NPAIR, B200_NTENDON, B200_NWRAP, B200_NSENSOR, B200_NM, B200_NSIZES };
/* indices into `opt` (float64) and `opt_int` */
enum { B200_OPT_TIMESTEP = 0, B200_OPT_GRAVITY = 1, B200_OPT_TOLERANCE = 4, B200_OPT_IMPRATIO = 5,
       B200_OPT_MEANINERTIA = 6, B200_OPT_LS_TOLERANCE = 7 };
enum { B200_OPTI_ITERATIONS = 0, B200_OPTI_LS_ITERATIONS = 1, B200_OPTI_INTEGRATOR = 2,
       B200_OPTI_NOSLIP_ITERATIONS = 3, B200_OPTI_WARMSTART = 4 };

#define B200M_INT_FIELDS(X) \
  X(sizes) X(opt_int) X(body_parent) X(body_jntadr) X(body_jntnum) X(body_dofadr) X(body_dofnum) \
  X(body_mocapid) X(body_rootid) X(jnt_type) X(jnt_body) X(jnt_qposadr) X(jnt_dofadr) X(jnt_limited) \
  X(dof_body) X(dof_jnt) X(dof_parent) X(geom_type) X(geom_body) X(pair_geom1) X(pair_geom2) X(pair_condim) \
  X(site_body) X(act_trnid) X(act_ctrllimited) X(act_forcelimited) X(eq_type) X(eq_obj1) X(eq_obj2) \
  X(eq_active) X(mocap_body) X(ten_adr) X(ten_num) X(ten_limited) X(wrap_dof) X(sensor_site) X(sensor_body) X(sensor_type) \
  X(pair_grid) X(grid_dims) X(grid_walls) X(geom_mjbody) X(mjbody_rt) X(geom_hull)

#define B200M_FLT_FIELDS(X) \
  X(opt) X(body_pos) X(body_quat) X(body_ipos) X(body_iquat) X(body_mass) X(body_inertia) X(jnt_pos) \
  X(jnt_axis) X(jnt_range) X(jnt_margin) X(jnt_stiffness) X(jnt_solref) X(jnt_solimp) X(qpos0) \
  X(qpos_spring) X(dof_armature) X(dof_damping) X(dof_frictionloss) X(dof_invweight0) X(dof_solref_fri) \
  X(dof_solimp_fri) X(geom_pos) X(geom_quat) X(geom_size) X(geom_rbound) X(pair_friction) X(pair_margin) \
  X(pair_gap) X(pair_solref) X(pair_solimp) X(pair_invweight) X(site_pos) X(site_quat) X(act_gear) \
  X(act_gainprm) X(act_biasprm) X(act_ctrlrange) X(act_forcerange) X(eq_data) X(eq_solref) X(eq_solimp) \
  X(eq_invweight) X(ten_range) X(ten_margin) X(ten_solref) X(ten_solimp) X(ten_invweight0) X(wrap_coef) \
  X(sensor_size) X(key_qpos) X(grid_param) X(hull_vert)

typedef struct b200_model_view {
  int nbody, njnt, nq, nv, nu, ngeom, nsite, nmocap, neq, npair, ntendon, nwrap, nsensor, nM;
#define X(f) const int32_t* f; int n_##f;
  B200M_INT_FIELDS(X)
#undef X
#define X(f) const double* f; int n_##f;
  B200M_FLT_FIELDS(X)
#undef X
} b200_model_view;

typedef struct b200m_entry { char name[32]; uint32_t dtype; uint32_t count; uint64_t offset; } b200m_entry;

/* Parse `blob` (must stay alive while the view is used).  Returns 0 on success, negative on error. */
static inline int b200_model_parse(const void* blob, size_t nbytes, b200_model_view* v) {
  const uint8_t* p = (const uint8_t*)blob;
  uint32_t head[4];
  if (nbytes < 16) return -1;
  memcpy(head, p, 16);
  if (head[0] != B200M_MAGIC || head[1] != B200M_VERSION) return -2;
  uint32_t n = head[2];
  size_t base = 16 + (size_t)n * sizeof(b200m_entry);
  if (base > nbytes) return -3;
  memset(v, 0, sizeof(*v));
  for (uint32_t i = 0; i < n; i++) {
    b200m_entry e;
    memcpy(&e, p + 16 + (size_t)i * sizeof(b200m_entry), sizeof(e));
    size_t bytes = (size_t)e.count * (e.dtype == 0 ? 4 : e.dtype == 1 ? 8 : 1);
    if (base + e.offset + bytes > nbytes) return -4;
    const void* data = p + base + e.offset;
#define X(f) if (e.dtype == 0 && strncmp(e.name, #f, 32) == 0) { v->f = (const int32_t*)data; v->n_##f = (int)e.count; continue; }
    B200M_INT_FIELDS(X)
#undef X
#define X(f) if (e.dtype == 1 && strncmp(e.name, #f, 32) == 0) { v->f = (const double*)data; v->n_##f = (int)e.count; continue; }
    B200M_FLT_FIELDS(X)
#undef X
  }
  if (!v->sizes || v->n_sizes < B200_NSIZES) return -5;
  v->nbody = v->sizes[B200_NBODY]; v->njnt = v->sizes[B200_NJNT]; v->nq = v->sizes[B200_NQ];
  v->nv = v->sizes[B200_NV]; v->nu = v->sizes[B200_NU]; v->ngeom = v->sizes[B200_NGEOM];
  v->nsite = v->sizes[B200_NSITE]; v->nmocap = v->sizes[B200_NMOCAP]; v->neq = v->sizes[B200_NEQ];
  v->npair = v->sizes[B200_NPAIR]; v->ntendon = v->sizes[B200_NTENDON]; v->nwrap = v->sizes[B200_NWRAP];
  v->nsensor = v->sizes[B200_NSENSOR]; v->nM = v->sizes[B200_NM];
  return 0;
}

#endif /* B200SIM_MODEL_H */
