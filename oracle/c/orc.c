/* orc.c -- C restatement of the oracle's kinematics (TEST INFRASTRUCTURE, see oracle/__init__.py).
 *
 * Same arithmetic as oracle/robot.py (which restates robot_wrapper.py:82-95 of the reference, i.e. pinocchio's
 * forwardKinematics / updateFramePlacement / computeFrameJacobian rotated to world axes): one 4x4 product per
 * URDF joint, fixed joints included, float64.  It exists so that the CPU baseline (the reference-faithful SLSQP
 * path) is not dominated by Python loops -- the reference runs this part in C++ (pinocchio) too -- and so that
 * the float64 polish used by the parity tests is quick.  Checked against the pure-Python implementation in
 * tests/test_oracle_c.py.  Build: gcc -O2 -shared -fPIC -o oracle/_build/liborc.so oracle/c/orc.c -lm
 */
#include <math.h>
#include <string.h>

typedef struct {
  int n_links, n_joints, dof;
  const int* parent_link;  /* [n_joints] link index of the joint's parent, joints in topological order */
  const int* child_link;   /* [n_joints] */
  const int* type;         /* [n_joints] 0 fixed, 1 revolute, 2 prismatic */
  const int* dof_index;    /* [n_joints] DoF index or -1 */
  const double* origin;    /* [n_joints][16] row-major 4x4 */
  const double* axis;      /* [n_joints][3] unit axis in the joint frame */
  const int* chain_start;  /* [n_links+1] movable joints between the root and each link (CSR) */
  const int* chain_joint;  /* joint indices */
} orc_model;

static void mat4_mul(const double* a, const double* b, double* c) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0.0;
      for (int k = 0; k < 4; ++k) s += a[4 * i + k] * b[4 * k + j];
      c[4 * i + j] = s;
    }
}

static void cross(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

/* poses [n_links][16]; axis_w, origin_w [n_joints][3] (movable joints only) */
void orc_fk(const orc_model* m, const double* q, double* poses, double* axis_w, double* origin_w) {
  memset(poses, 0, sizeof(double) * 16);
  poses[0] = poses[5] = poses[10] = poses[15] = 1.0; /* link 0 is the root */
  for (int j = 0; j < m->n_joints; ++j) {
    double T[16], M[16], TM[16];
    mat4_mul(poses + 16 * m->parent_link[j], m->origin + 16 * j, T);
    double* out = poses + 16 * m->child_link[j];
    if (m->type[j] == 0) {
      memcpy(out, T, sizeof(T));
      continue;
    }
    const double* a = m->axis + 3 * j;
    for (int r = 0; r < 3; ++r) {
      axis_w[3 * j + r] = T[4 * r] * a[0] + T[4 * r + 1] * a[1] + T[4 * r + 2] * a[2];
      origin_w[3 * j + r] = T[4 * r + 3];
    }
    const double v = q[m->dof_index[j]];
    memset(M, 0, sizeof(M));
    M[0] = M[5] = M[10] = M[15] = 1.0;
    if (m->type[j] == 1) { /* Rodrigues: I + sin K + (1 - cos) K^2 */
      const double s = sin(v), c1 = 1.0 - cos(v);
      const double K[9] = {0, -a[2], a[1], a[2], 0, -a[0], -a[1], a[0], 0};
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
          double k2 = 0.0;
          for (int k = 0; k < 3; ++k) k2 += K[3 * r + k] * K[3 * k + c];
          M[4 * r + c] += s * K[3 * r + c] + c1 * k2;
        }
    } else {
      M[3] = a[0] * v; M[7] = a[1] * v; M[11] = a[2] * v;
    }
    mat4_mul(T, M, TM);
    memcpy(out, TM, sizeof(TM));
  }
}

/* positions [n][3] and world-aligned linear Jacobians [n][3][dof] of the requested links */
void orc_positions_jacobians(const orc_model* m, const double* poses, const double* axis_w, const double* origin_w,
                             int n, const int* link_ids, double* pos, double* jac) {
  if (jac) memset(jac, 0, sizeof(double) * (size_t)n * 3 * m->dof);
  for (int r = 0; r < n; ++r) {
    const int l = link_ids[r];
    const double* P = poses + 16 * l;
    const double p[3] = {P[3], P[7], P[11]};
    pos[3 * r] = p[0]; pos[3 * r + 1] = p[1]; pos[3 * r + 2] = p[2];
    if (!jac) continue;
    for (int c = m->chain_start[l]; c < m->chain_start[l + 1]; ++c) {
      const int j = m->chain_joint[c];
      const int i = m->dof_index[j];
      double col[3];
      if (m->type[j] == 1) {
        const double d[3] = {p[0] - origin_w[3 * j], p[1] - origin_w[3 * j + 1], p[2] - origin_w[3 * j + 2]};
        cross(axis_w + 3 * j, d, col);
      } else {
        col[0] = axis_w[3 * j]; col[1] = axis_w[3 * j + 1]; col[2] = axis_w[3 * j + 2];
      }
      for (int k = 0; k < 3; ++k) jac[((size_t)r * 3 + k) * m->dof + i] = col[k];
    }
  }
}

/* S[dof][dof] = sum_r sum_c g[r][c] * d2 p_r[c] / dq dq   (see oracle/robot.py link_position_hessian_contraction) */
void orc_hessian_contraction(const orc_model* m, const double* poses, const double* axis_w, const double* origin_w,
                             int n, const int* link_ids, const double* g, double* S) {
  memset(S, 0, sizeof(double) * (size_t)m->dof * m->dof);
  for (int r = 0; r < n; ++r) {
    const int l = link_ids[r];
    const double* P = poses + 16 * l;
    const double p[3] = {P[3], P[7], P[11]};
    for (int u = m->chain_start[l]; u < m->chain_start[l + 1]; ++u) {
      const int ji = m->chain_joint[u];
      if (m->type[ji] != 1) continue;
      const int i = m->dof_index[ji];
      for (int w = u; w < m->chain_start[l + 1]; ++w) {
        const int jj = m->chain_joint[w];
        const int j = m->dof_index[jj];
        double inner[3], d[3];
        if (m->type[jj] == 1) {
          const double e[3] = {p[0] - origin_w[3 * jj], p[1] - origin_w[3 * jj + 1], p[2] - origin_w[3 * jj + 2]};
          cross(axis_w + 3 * jj, e, inner);
        } else {
          inner[0] = axis_w[3 * jj]; inner[1] = axis_w[3 * jj + 1]; inner[2] = axis_w[3 * jj + 2];
        }
        cross(axis_w + 3 * ji, inner, d);
        const double v = g[3 * r] * d[0] + g[3 * r + 1] * d[1] + g[3 * r + 2] * d[2];
        S[(size_t)i * m->dof + j] += v;
        if (i != j) S[(size_t)j * m->dof + i] += v;
      }
    }
  }
}
