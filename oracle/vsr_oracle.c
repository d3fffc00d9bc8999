/*
 * vsr_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).  See vsr_oracle.h.
 *
 * PARITY UNPINNED (no reference source / JVM / golden vectors for this path).
 *
 * What is restated, and which upstream file each block follows (all un-vendored, recalled, [U]):
 *   build_robot()        2dhmsr core/objects/Voxel.java assemble(), core/objects/Robot.java assemble()/join()
 *   place_robot()        2dhmsr tasks/locomotion/Locomotion.java apply(): translate to x-gap / y-gap
 *   world_step()         dyn4j dynamics/World.java step() + dynamics/Island.java solve()
 *   spring_*             dyn4j dynamics/joint/DistanceJoint.java (frequency>0 soft constraint)
 *   rope_*               dyn4j dynamics/joint/RopeJoint.java (lower limit only)
 *   weld_*               dyn4j dynamics/joint/WeldJoint.java (rigid, frequency 0)
 *   collide_box_poly()   dyn4j collision/narrowphase + collision/manifold/ClippingManifoldSolver.java
 *                        (restated as SAT + reference-face clipping, which yields the same normal,
 *                        depth and clipped points as GJK/EPA + clipping for convex polygons)
 *   contacts_*           dyn4j dynamics/contact/SequentialImpulses.java
 *   sense()/control()    2dhmsr core/sensors/{AreaRatio,Touch,Velocity,Angle,Average,SoftNormalization}.java,
 *                        core/controllers/{PhaseSin,CentralizedSensing,DistributedSensing,MultiLayerPerceptron}.java
 *   episode loop/fitness 2dhmsr Locomotion.apply(), Outcome.subOutcome()/getVelocity();
 *                        call site /root/reference/src/main/java/it/units/erallab/locomotion/Starter.java:427-435,174
 *
 * Constraint order ("canonical order"): per velocity iteration
 *   for every voxel in grid order: its springs then its ropes; then all horizontal welds; then all
 *   vertical welds; then contacts body by body.  Position iterations: contacts, then ropes, then welds.
 * Not modelled (documented in DESIGN.md): body sleeping, continuous collision detection, robot
 * self-collision (filtered out in 2dhmsr).
 */
#include "vsr_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define MAXM 3 /* max ground polygons one corner mass can have a manifold with */

/* ------------------------------------------------------------------------------------------- */
/* types                                                                                        */
/* ------------------------------------------------------------------------------------------- */
typedef struct {
  double x, y, a;    /* centre, angle */
  double vx, vy, w;  /* velocities */
  double c, s;       /* cos(a), sin(a) */
} Body;

enum { J_SPRING = 0, J_ROPE = 1, J_WELD = 2 };
enum { R_PARALLEL = 0, R_CROSS = 1, R_CENTRAL = 2 };

typedef struct {
  int type, b1, b2;
  double a1x, a1y, a2x, a2y; /* local anchors (body frames) */
  int voxel, range_kind;     /* spring: owner voxel, which SpringRange */
  double rest;               /* spring: current target distance */
  double lower;              /* rope: lower limit */
  double ref_angle;          /* weld */
  /* per-step solver data */
  double nx, ny, r1x, r1y, r2x, r2y;
  double invK, bias, gamma, impulse;
  int active;
  double K[6];   /* weld: symmetric 3x3 (00,01,02,11,12,22) */
  double imp3[3];
} Joint;

typedef struct {
  double px, py;  /* world point at detection (on the incident polygon) */
  double lx, ly;  /* same point in the corner mass' local frame */
  double depth;
  uint32_t id;
  double jn, jt;
  double rx, ry, massN, massT, vb;
} CPoint;

typedef struct {
  int poly, npts;
  double nx, ny; /* ground -> mass */
  CPoint pt[2];
  int block;
  double K11, K12, K22;
} Manifold;

typedef struct {
  vo_params p;
  /* derived */
  double m_side, half, inv_m, inv_i, mu, rest_e;
  double range_min[3], range_rest[3], range_max[3];
  int w, h, n_vox;
  int* vox_of_cell; /* [w*h] -> voxel index or -1 */
  int* cell_x;      /* [n_vox] */
  int* cell_y;
  int n_bodies;
  Body* bodies;
  int n_joints;
  Joint* joints;
  /* ground */
  int n_t;
  double *tx, *ty;
  double base_y;
  Manifold* man; /* [n_bodies][MAXM] */
  int* n_man;    /* [n_bodies] */
  /* control */
  double* ctrl_f; /* [n_vox] last applied control */
  uint32_t frozen; /* bit v: actuator of voxel v frozen at 0 (vo_extras.frozen_mask) */
  vo_result* tally;
} World;

/* ------------------------------------------------------------------------------------------- */
/* defaults                                                                                     */
/* ------------------------------------------------------------------------------------------- */
void vo_default_params(vo_params* p) {
  p->dt = 1.0 / 60.0;
  p->vel_iters = 10;
  p->pos_iters = 10;
  p->gravity_y = -9.8;
  p->baumgarte = 0.2;
  p->linear_tol = 0.005;
  p->angular_tol = 2.0 * M_PI / 180.0;
  p->max_linear_corr = 0.2;
  p->max_translation = 2.0;
  p->max_rotation = 0.5 * M_PI;
  p->restitution_velocity = 1.0;
  p->side_length = 3.0;
  p->mass_side_ratio = 0.3;
  p->spring_f = 8.0;
  p->spring_d = 0.3;
  p->lin_damping = 1.0;
  p->ang_damping = 1.0;
  p->voxel_mass = 1.0;
  p->area_ratio_max_delta = 0.225;
  p->voxel_friction = 100.0;
  p->voxel_restitution = 0.1;
  p->limit_contraction = 1;
  p->scaffolding_mask = 0xF;
  p->ground_friction = 0.2;
  p->ground_restitution = 0.0;
  p->ground_thickness = 50.0;
  p->place_x_gap = 1.0;
  p->place_y_gap = 1.0;
  p->final_t = 30.0;
  p->transient_t = 1.0;
}

/* ------------------------------------------------------------------------------------------- */
/* robot assembly                                                                               */
/* ------------------------------------------------------------------------------------------- */
static void body_set_angle(Body* b, double a) {
  b->a = a;
  b->c = cos(a);
  b->s = sin(a);
}

static Joint* add_joint(World* W, int type, int b1, int b2, double a1x, double a1y, double a2x, double a2y) {
  Joint* j = &W->joints[W->n_joints++];
  memset(j, 0, sizeof(*j));
  j->type = type;
  j->b1 = b1;
  j->b2 = b2;
  j->a1x = a1x;
  j->a1y = a1y;
  j->a2x = a2x;
  j->a2y = a2y;
  return j;
}

/* pair order inside every scaffolding group: (0,1) (2,3) (1,2) (3,0) */
static const int PAIR_A[4] = {0, 2, 1, 3};
static const int PAIR_B[4] = {1, 3, 2, 0};

static void count_joints(const vo_params* p, int w, int h, const uint8_t* mask, int* n_vox, int* n_springs, int* n_ropes, int* n_welds) {
  int nv = 0, nw = 0;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++)
      if (mask[y * w + x]) {
        nv++;
        if (x > 0 && mask[y * w + x - 1]) nw += 2;
        if (y > 0 && mask[(y - 1) * w + x]) nw += 2;
      }
  int per = 0;
  if (p->scaffolding_mask & 1) per += 4;
  if (p->scaffolding_mask & 2) per += 4;
  if (p->scaffolding_mask & 4) per += 8;
  if (p->scaffolding_mask & 8) per += 2;
  *n_vox = nv;
  *n_springs = nv * per;
  *n_ropes = p->limit_contraction ? nv * 4 : 0;
  *n_welds = nw;
}

int vo_counts(const vo_params* p, int w, int h, const uint8_t* mask, int* n_vox, int* n_springs, int* n_ropes, int* n_welds) {
  if (w <= 0 || h <= 0 || !mask) return -1;
  count_joints(p, w, h, mask, n_vox, n_springs, n_ropes, n_welds);
  return 0;
}

static void add_spring(World* W, int v, int k1, int k2, double sx1, double sy1, double sx2, double sy2, int kind) {
  double hh = W->half;
  Joint* j = add_joint(W, J_SPRING, v * 4 + k1, v * 4 + k2, sx1 * hh, sy1 * hh, sx2 * hh, sy2 * hh);
  j->voxel = v;
  j->range_kind = kind;
  j->rest = W->range_rest[kind];
}

/* Voxel.assemble(): bodies NW(0) NE(1) SE(2) SW(3); spring anchors on the facing corners. */
static void build_robot(World* W, const uint8_t* mask) {
  const vo_params* p = &W->p;
  double L = p->side_length, m = W->m_side, off = L / 2 - m / 2;
  static const double OX[4] = {-1, +1, +1, -1}, OY[4] = {+1, +1, -1, -1};
  int v = 0;
  for (int y = 0; y < W->h; y++)
    for (int x = 0; x < W->w; x++) {
      int cell = y * W->w + x;
      W->vox_of_cell[cell] = -1;
      if (!mask[cell]) continue;
      W->vox_of_cell[cell] = v;
      W->cell_x[v] = x;
      W->cell_y[v] = y;
      for (int k = 0; k < 4; k++) {
        Body* b = &W->bodies[v * 4 + k];
        memset(b, 0, sizeof(*b));
        b->x = x * L + OX[k] * off;
        b->y = y * L + OY[k] * off;
        body_set_angle(b, 0.0);
      }
      v++;
    }
  /* anchors as sign pairs (multiples of half mass side), per pair (a,b) listed for bodies a then b */
  /* SIDE_INTERNAL */
  static const double SI[4][4] = {/*01*/ {+1, -1, -1, -1}, /*23*/ {-1, +1, +1, +1}, /*12*/ {-1, -1, -1, +1}, /*30*/ {+1, +1, +1, -1}};
  /* SIDE_EXTERNAL */
  static const double SE[4][4] = {/*01*/ {+1, +1, -1, +1}, /*23*/ {-1, -1, +1, -1}, /*12*/ {+1, -1, +1, +1}, /*30*/ {-1, +1, -1, -1}};
  /* SIDE_CROSS a and b */
  static const double SCA[4][4] = {/*01*/ {+1, +1, -1, -1}, /*23*/ {-1, +1, +1, -1}, /*12*/ {+1, -1, -1, +1}, /*30*/ {-1, +1, +1, -1}};
  static const double SCB[4][4] = {/*01*/ {+1, -1, -1, +1}, /*23*/ {-1, -1, +1, +1}, /*12*/ {-1, -1, +1, +1}, /*30*/ {+1, +1, -1, -1}};
  for (v = 0; v < W->n_vox; v++) {
    if (p->scaffolding_mask & 1)
      for (int q = 0; q < 4; q++) add_spring(W, v, PAIR_A[q], PAIR_B[q], SI[q][0], SI[q][1], SI[q][2], SI[q][3], R_PARALLEL);
    if (p->scaffolding_mask & 2)
      for (int q = 0; q < 4; q++) add_spring(W, v, PAIR_A[q], PAIR_B[q], SE[q][0], SE[q][1], SE[q][2], SE[q][3], R_PARALLEL);
    if (p->scaffolding_mask & 4) {
      for (int q = 0; q < 4; q++) add_spring(W, v, PAIR_A[q], PAIR_B[q], SCA[q][0], SCA[q][1], SCA[q][2], SCA[q][3], R_CROSS);
      for (int q = 0; q < 4; q++) add_spring(W, v, PAIR_A[q], PAIR_B[q], SCB[q][0], SCB[q][1], SCB[q][2], SCB[q][3], R_CROSS);
    }
    if (p->scaffolding_mask & 8) {
      add_spring(W, v, 0, 2, 0, 0, 0, 0, R_CENTRAL);
      add_spring(W, v, 1, 3, 0, 0, 0, 0, R_CENTRAL);
    }
    if (p->limit_contraction) {
      double min_side = sqrt(L * L * (1.0 - p->area_ratio_max_delta));
      for (int q = 0; q < 4; q++) {
        Joint* j = add_joint(W, J_ROPE, v * 4 + PAIR_A[q], v * 4 + PAIR_B[q], 0, 0, 0, 0);
        j->voxel = v;
        j->lower = min_side - m;
      }
    }
  }
  /* Robot.assemble(): welds; join(body1, body2) anchors at body1's centre (the upstream quirk). */
  for (int pass = 0; pass < 2; pass++) /* pass 0: horizontal (left neighbour), pass 1: vertical (below) */
    for (v = 0; v < W->n_vox; v++) {
      int x = W->cell_x[v], y = W->cell_y[v];
      int nb = -1, k1[2], k2[2];
      if (pass == 0 && x > 0) {
        nb = W->vox_of_cell[y * W->w + x - 1];
        k1[0] = 0; k2[0] = 1; k1[1] = 3; k2[1] = 2;
      }
      if (pass == 1 && y > 0) {
        nb = W->vox_of_cell[(y - 1) * W->w + x];
        k1[0] = 3; k2[0] = 0; k1[1] = 2; k2[1] = 1;
      }
      if (nb < 0) continue;
      for (int q = 0; q < 2; q++) {
        Body* b1 = &W->bodies[v * 4 + k1[q]];
        Body* b2 = &W->bodies[nb * 4 + k2[q]];
        /* anchor (world) = b1 centre: local1 = 0, local2 = b1 - b2 (angles are 0 at assembly) */
        Joint* j = add_joint(W, J_WELD, v * 4 + k1[q], nb * 4 + k2[q], 0, 0, b1->x - b2->x, b1->y - b2->y);
        j->ref_angle = 0.0;
      }
    }
}

/* ------------------------------------------------------------------------------------------- */
/* ground                                                                                       */
/* ------------------------------------------------------------------------------------------- */
static double ground_y_at(const World* W, double x) {
  /* Ground.yAt(): linear interpolation on the polyline; outside -> -inf (never hit in practice) */
  for (int i = 1; i < W->n_t; i++)
    if (x >= W->tx[i - 1] && x <= W->tx[i]) {
      double u = (x - W->tx[i - 1]) / (W->tx[i] - W->tx[i - 1]);
      return W->ty[i - 1] + u * (W->ty[i] - W->ty[i - 1]);
    }
  return -INFINITY;
}

static void box_vertices(const Body* b, double hh, double v[4][2]) {
  /* dyn4j Rectangle vertex order: (-,-) (+,-) (+,+) (-,+), CCW */
  static const double SX[4] = {-1, +1, +1, -1}, SY[4] = {-1, -1, +1, +1};
  for (int i = 0; i < 4; i++) {
    double lx = SX[i] * hh, ly = SY[i] * hh;
    v[i][0] = b->x + b->c * lx - b->s * ly;
    v[i][1] = b->y + b->s * lx + b->c * ly;
  }
}

/* Locomotion.apply(): translate so bbox.min.x = xs[1] + x_gap, then so the lowest voxel is y_gap above ground */
static void place_robot(World* W) {
  double L = W->p.side_length;
  double minx = INFINITY;
  for (int i = 0; i < W->n_bodies; i++) minx = fmin(minx, W->bodies[i].x - W->half);
  double dx = (W->tx[1] + W->p.place_x_gap) - minx;
  for (int i = 0; i < W->n_bodies; i++) W->bodies[i].x += dx;
  double min_gap = INFINITY;
  for (int v = 0; v < W->n_vox; v++) {
    double cx = 0, miny = INFINITY;
    for (int k = 0; k < 4; k++) {
      cx += W->bodies[v * 4 + k].x;
      miny = fmin(miny, W->bodies[v * 4 + k].y - W->half);
    }
    cx /= 4.0;
    min_gap = fmin(min_gap, miny - ground_y_at(W, cx));
  }
  (void)L;
  double dy = W->p.place_y_gap - min_gap;
  for (int i = 0; i < W->n_bodies; i++) W->bodies[i].y += dy;
}

/* ------------------------------------------------------------------------------------------- */
/* narrowphase: box vs convex ground polygon (SAT + clipping)                                   */
/* ------------------------------------------------------------------------------------------- */
static void poly_normals(const double v[4][2], double n[4][2]) {
  for (int i = 0; i < 4; i++) {
    int j = (i + 1) & 3;
    double ex = v[j][0] - v[i][0], ey = v[j][1] - v[i][1];
    double len = sqrt(ex * ex + ey * ey);
    n[i][0] = ey / len; /* outward for CCW */
    n[i][1] = -ex / len;
  }
}

static double max_separation(const double a[4][2], const double an[4][2], const double b[4][2], int* edge) {
  double best = -INFINITY;
  int be = 0;
  for (int i = 0; i < 4; i++) {
    double s = INFINITY;
    for (int j = 0; j < 4; j++) {
      double d = an[i][0] * (b[j][0] - a[i][0]) + an[i][1] * (b[j][1] - a[i][1]);
      if (d < s) s = d;
    }
    if (s > best) {
      best = s;
      be = i;
    }
  }
  *edge = be;
  return best;
}

/* returns number of contact points (0 = no manifold) */
static int collide_box_poly(const double g[4][2], const double bx[4][2], double lin_tol, Manifold* m) {
  double gn[4][2], bn[4][2];
  poly_normals(g, gn);
  poly_normals(bx, bn);
  int eA, eB;
  double sA = max_separation(g, gn, bx, &eA);
  if (sA > 0) return 0;
  double sB = max_separation(bx, bn, g, &eB);
  if (sB > 0) return 0;
  const double(*ref)[2], (*refn)[2], (*inc)[2], (*incn)[2];
  int e, flip;
  if (sB > sA + 0.1 * lin_tol) {
    ref = bx; refn = bn; inc = g; incn = gn; e = eB; flip = 1;
  } else {
    ref = g; refn = gn; inc = bx; incn = bn; e = eA; flip = 0;
  }
  double nx = refn[e][0], ny = refn[e][1];
  /* incident edge: most anti-parallel normal */
  int ie = 0;
  double mind = INFINITY;
  for (int j = 0; j < 4; j++) {
    double d = nx * incn[j][0] + ny * incn[j][1];
    if (d < mind) {
      mind = d;
      ie = j;
    }
  }
  double c[2][2] = {{inc[ie][0], inc[ie][1]}, {inc[(ie + 1) & 3][0], inc[(ie + 1) & 3][1]}};
  const double* v1 = ref[e];
  const double* v2 = ref[(e + 1) & 3];
  double tx = v2[0] - v1[0], ty = v2[1] - v1[1];
  double tl = sqrt(tx * tx + ty * ty);
  tx /= tl;
  ty /= tl;
  /* clip against t.p >= t.v1 */
  {
    double o = tx * v1[0] + ty * v1[1];
    double d0 = tx * c[0][0] + ty * c[0][1] - o, d1 = tx * c[1][0] + ty * c[1][1] - o;
    if (d0 < 0 && d1 < 0) return 0;
    if (d0 < 0 || d1 < 0) {
      double u = d0 / (d0 - d1);
      double ix = c[0][0] + u * (c[1][0] - c[0][0]), iy = c[0][1] + u * (c[1][1] - c[0][1]);
      if (d0 < 0) { c[0][0] = ix; c[0][1] = iy; } else { c[1][0] = ix; c[1][1] = iy; }
    }
  }
  /* clip against t.p <= t.v2 */
  {
    double o = tx * v2[0] + ty * v2[1];
    double d0 = o - (tx * c[0][0] + ty * c[0][1]), d1 = o - (tx * c[1][0] + ty * c[1][1]);
    if (d0 < 0 && d1 < 0) return 0;
    if (d0 < 0 || d1 < 0) {
      double u = d0 / (d0 - d1);
      double ix = c[0][0] + u * (c[1][0] - c[0][0]), iy = c[0][1] + u * (c[1][1] - c[0][1]);
      if (d0 < 0) { c[0][0] = ix; c[0][1] = iy; } else { c[1][0] = ix; c[1][1] = iy; }
    }
  }
  int np = 0;
  for (int q = 0; q < 2; q++) {
    double sep = nx * (c[q][0] - v1[0]) + ny * (c[q][1] - v1[1]);
    if (sep <= 0) {
      CPoint* cp = &m->pt[np++];
      memset(cp, 0, sizeof(*cp));
      cp->px = c[q][0];
      cp->py = c[q][1];
      cp->depth = -sep;
      cp->id = ((uint32_t)flip << 24) | ((uint32_t)e << 16) | ((uint32_t)ie << 8) | (uint32_t)q;
    }
  }
  m->npts = np;
  m->nx = flip ? -nx : nx;
  m->ny = flip ? -ny : ny;
  return np;
}

/* World.detect(): every corner mass against every ground polygon whose AABB it overlaps. */
static void detect(World* W) {
  double hh = W->half;
  for (int i = 0; i < W->n_bodies; i++) {
    Body* b = &W->bodies[i];
    Manifold old[MAXM];
    int n_old = W->n_man[i];
    memcpy(old, &W->man[i * MAXM], sizeof(old));
    int nm = 0;
    double ext = hh * (fabs(b->c) + fabs(b->s));
    double ax0 = b->x - ext, ax1 = b->x + ext, ay0 = b->y - ext;
    double bv[4][2];
    box_vertices(b, hh, bv);
    for (int k = 1; k < W->n_t && nm < MAXM; k++) {
      double x0 = W->tx[k - 1], x1 = W->tx[k], y0 = W->ty[k - 1], y1 = W->ty[k];
      if (ax1 < x0 || ax0 > x1) continue;
      if (ay0 > fmax(y0, y1)) continue;
      W->tally->n_detect_pairs++;
      double g[4][2] = {{x0, W->base_y}, {x1, W->base_y}, {x1, y1}, {x0, y0}};
      Manifold* m = &W->man[i * MAXM + nm];
      memset(m, 0, sizeof(*m));
      if (!collide_box_poly(g, bv, W->p.linear_tol, m)) continue;
      m->poly = k;
      for (int q = 0; q < m->npts; q++) {
        CPoint* cp = &m->pt[q];
        double dx = cp->px - b->x, dy = cp->py - b->y;
        cp->lx = b->c * dx + b->s * dy;
        cp->ly = -b->s * dx + b->c * dy;
        /* warm start: same polygon, same feature id */
        for (int o = 0; o < n_old; o++)
          if (old[o].poly == k)
            for (int r = 0; r < old[o].npts; r++)
              if (old[o].pt[r].id == cp->id) {
                cp->jn = old[o].pt[r].jn;
                cp->jt = old[o].pt[r].jt;
              }
      }
      nm++;
      W->tally->n_manifolds++;
    }
    W->n_man[i] = nm;
  }
}

/* ------------------------------------------------------------------------------------------- */
/* joints                                                                                       */
/* ------------------------------------------------------------------------------------------- */
static inline void rot(const Body* b, double lx, double ly, double* ox, double* oy) {
  *ox = b->c * lx - b->s * ly;
  *oy = b->s * lx + b->c * ly;
}

static void spring_init(World* W, Joint* j) {
  Body *b1 = &W->bodies[j->b1], *b2 = &W->bodies[j->b2];
  double im = W->inv_m, ii = W->inv_i, dt = W->p.dt;
  rot(b1, j->a1x, j->a1y, &j->r1x, &j->r1y);
  rot(b2, j->a2x, j->a2y, &j->r2x, &j->r2y);
  double nx = (b1->x + j->r1x) - (b2->x + j->r2x), ny = (b1->y + j->r1y) - (b2->y + j->r2y);
  double len = sqrt(nx * nx + ny * ny);
  if (len < W->p.linear_tol) { nx = 0; ny = 0; } else { nx /= len; ny /= len; }
  j->nx = nx; j->ny = ny;
  double cr1 = j->r1x * ny - j->r1y * nx, cr2 = j->r2x * ny - j->r2y * nx;
  double inv_mass = im + ii * cr1 * cr1 + im + ii * cr2 * cr2;
  double k0 = inv_mass <= 1e-300 ? 0.0 : 1.0 / inv_mass;
  /* soft constraint (frequency > 0) */
  double x = len - j->rest;
  double w = 2.0 * M_PI * W->p.spring_f;
  double d = 2.0 * k0 * W->p.spring_d * w;
  double k = k0 * w * w;
  double gamma = dt * (d + dt * k);
  gamma = gamma <= 1e-300 ? 0.0 : 1.0 / gamma;
  j->gamma = gamma;
  j->bias = x * dt * k * gamma;
  inv_mass += gamma;
  j->invK = inv_mass <= 1e-300 ? 0.0 : 1.0 / inv_mass;
  /* warm start (dt ratio = 1) */
  double Jx = nx * j->impulse, Jy = ny * j->impulse;
  b1->vx += Jx * im; b1->vy += Jy * im; b1->w += ii * (j->r1x * Jy - j->r1y * Jx);
  b2->vx -= Jx * im; b2->vy -= Jy * im; b2->w -= ii * (j->r2x * Jy - j->r2y * Jx);
}

static void spring_solve_vel(World* W, Joint* j) {
  Body *b1 = &W->bodies[j->b1], *b2 = &W->bodies[j->b2];
  double im = W->inv_m, ii = W->inv_i;
  double v1x = b1->vx - b1->w * j->r1y, v1y = b1->vy + b1->w * j->r1x;
  double v2x = b2->vx - b2->w * j->r2y, v2y = b2->vy + b2->w * j->r2x;
  double Jv = j->nx * (v1x - v2x) + j->ny * (v1y - v2y);
  double l = -j->invK * (Jv + j->bias + j->gamma * j->impulse);
  j->impulse += l;
  double Jx = j->nx * l, Jy = j->ny * l;
  b1->vx += Jx * im; b1->vy += Jy * im; b1->w += ii * (j->r1x * Jy - j->r1y * Jx);
  b2->vx -= Jx * im; b2->vy -= Jy * im; b2->w -= ii * (j->r2x * Jy - j->r2y * Jx);
  W->tally->n_spring_vel++;
}

/* Rope joint with only the lower limit enabled: unilateral  len >= lower. */
static void rope_init(World* W, Joint* j) {
  Body *b1 = &W->bodies[j->b1], *b2 = &W->bodies[j->b2];
  double im = W->inv_m, ii = W->inv_i;
  rot(b1, j->a1x, j->a1y, &j->r1x, &j->r1y);
  rot(b2, j->a2x, j->a2y, &j->r2x, &j->r2y);
  double nx = (b1->x + j->r1x) - (b2->x + j->r2x), ny = (b1->y + j->r1y) - (b2->y + j->r2y);
  double len = sqrt(nx * nx + ny * ny);
  if (len < W->p.linear_tol) { nx = 0; ny = 0; } else { nx /= len; ny /= len; }
  j->nx = nx; j->ny = ny;
  j->active = len < j->lower;
  if (!j->active) { j->impulse = 0; return; }
  double cr1 = j->r1x * ny - j->r1y * nx, cr2 = j->r2x * ny - j->r2y * nx;
  double inv_mass = im + ii * cr1 * cr1 + im + ii * cr2 * cr2;
  j->invK = inv_mass <= 1e-300 ? 0.0 : 1.0 / inv_mass;
  double Jx = nx * j->impulse, Jy = ny * j->impulse;
  b1->vx += Jx * im; b1->vy += Jy * im; b1->w += ii * (j->r1x * Jy - j->r1y * Jx);
  b2->vx -= Jx * im; b2->vy -= Jy * im; b2->w -= ii * (j->r2x * Jy - j->r2y * Jx);
}

static void rope_solve_vel(World* W, Joint* j) {
  if (!j->active) return;
  Body *b1 = &W->bodies[j->b1], *b2 = &W->bodies[j->b2];
  double im = W->inv_m, ii = W->inv_i;
  double v1x = b1->vx - b1->w * j->r1y, v1y = b1->vy + b1->w * j->r1x;
  double v2x = b2->vx - b2->w * j->r2y, v2y = b2->vy + b2->w * j->r2x;
  double Jv = j->nx * (v1x - v2x) + j->ny * (v1y - v2y);
  double l = -j->invK * Jv;
  double old = j->impulse;
  j->impulse = fmax(old + l, 0.0); /* only push apart */
  l = j->impulse - old;
  double Jx = j->nx * l, Jy = j->ny * l;
  b1->vx += Jx * im; b1->vy += Jy * im; b1->w += ii * (j->r1x * Jy - j->r1y * Jx);
  b2->vx -= Jx * im; b2->vy -= Jy * im; b2->w -= ii * (j->r2x * Jy - j->r2y * Jx);
  W->tally->n_rope_vel++;
}

static int rope_solve_pos(World* W, Joint* j) {
  if (!j->active) return 1;
  Body *b1 = &W->bodies[j->b1], *b2 = &W->bodies[j->b2];
  double im = W->inv_m, ii = W->inv_i;
  double r1x, r1y, r2x, r2y;
  rot(b1, j->a1x, j->a1y, &r1x, &r1y);
  rot(b2, j->a2x, j->a2y, &r2x, &r2y);
  double nx = (b1->x + r1x) - (b2->x + r2x), ny = (b1->y + r1y) - (b2->y + r2y);
  double len = sqrt(nx * nx + ny * ny);
  if (len < W->p.linear_tol) { nx = 0; ny = 0; } else { nx /= len; ny /= len; }
  double C = len - j->lower;
  int solved = -C < W->p.linear_tol;
  if (C > 0) C = 0;
  if (C < -W->p.max_linear_corr) C = -W->p.max_linear_corr;
  double cr1 = r1x * ny - r1y * nx, cr2 = r2x * ny - r2y * nx;
  double inv_mass = im + ii * cr1 * cr1 + im + ii * cr2 * cr2;
  double k = inv_mass <= 1e-300 ? 0.0 : 1.0 / inv_mass;
  double l = -k * C;
  double Jx = nx * l, Jy = ny * l;
  b1->x += Jx * im; b1->y += Jy * im; body_set_angle(b1, b1->a + ii * (r1x * Jy - r1y * Jx));
  b2->x -= Jx * im; b2->y -= Jy * im; body_set_angle(b2, b2->a - ii * (r2x * Jy - r2y * Jx));
  W->tally->n_rope_pos++;
  return solved;
}

/* symmetric 3x3 solve K x = b, K = (00,01,02,11,12,22) -- cofactor form like dyn4j Matrix33.solve33 */
static void solve33(const double K[6], const double b[3], double x[3]) {
  double a = K[0], bb = K[1], c = K[2], d = K[3], e = K[4], f = K[5];
  double c00 = d * f - e * e, c01 = c * e - bb * f, c02 = bb * e - c * d;
  double det = a * c00 + bb * c01 + c * c02;
  if (det != 0.0) det = 1.0 / det;
  double c11 = a * f - c * c, c12 = bb * c - a * e, c22 = a * d - bb * bb;
  x[0] = det * (c00 * b[0] + c01 * b[1] + c02 * b[2]);
  x[1] = det * (c01 * b[0] + c11 * b[1] + c12 * b[2]);
  x[2] = det * (c02 * b[0] + c12 * b[1] + c22 * b[2]);
}

static void weld_K(const World* W, double r1x, double r1y, double r2x, double r2y, double K[6]) {
  double im = W->inv_m, ii = W->inv_i;
  K[0] = im + im + r1y * r1y * ii + r2y * r2y * ii;
  K[1] = -r1y * r1x * ii - r2y * r2x * ii;
  K[2] = -r1y * ii - r2y * ii;
  K[3] = im + im + r1x * r1x * ii + r2x * r2x * ii;
  K[4] = r1x * ii + r2x * ii;
  K[5] = ii + ii;
}

static void weld_apply_vel(World* W, Joint* j, const double l[3]) {
  Body *b1 = &W->bodies[j->b1], *b2 = &W->bodies[j->b2];
  double im = W->inv_m, ii = W->inv_i;
  b1->vx += im * l[0]; b1->vy += im * l[1]; b1->w += ii * (j->r1x * l[1] - j->r1y * l[0] + l[2]);
  b2->vx -= im * l[0]; b2->vy -= im * l[1]; b2->w -= ii * (j->r2x * l[1] - j->r2y * l[0] + l[2]);
}

static void weld_init(World* W, Joint* j) {
  Body *b1 = &W->bodies[j->b1], *b2 = &W->bodies[j->b2];
  rot(b1, j->a1x, j->a1y, &j->r1x, &j->r1y);
  rot(b2, j->a2x, j->a2y, &j->r2x, &j->r2y);
  weld_K(W, j->r1x, j->r1y, j->r2x, j->r2y, j->K);
  weld_apply_vel(W, j, j->imp3);
}

static void weld_solve_vel(World* W, Joint* j) {
  Body *b1 = &W->bodies[j->b1], *b2 = &W->bodies[j->b2];
  double C[3];
  C[0] = -((b1->vx - b1->w * j->r1y) - (b2->vx - b2->w * j->r2y));
  C[1] = -((b1->vy + b1->w * j->r1x) - (b2->vy + b2->w * j->r2x));
  C[2] = -(b1->w - b2->w);
  double l[3];
  solve33(j->K, C, l);
  j->imp3[0] += l[0]; j->imp3[1] += l[1]; j->imp3[2] += l[2];
  weld_apply_vel(W, j, l);
  W->tally->n_weld_vel++;
}

static int weld_solve_pos(World* W, Joint* j) {
  Body *b1 = &W->bodies[j->b1], *b2 = &W->bodies[j->b2];
  double im = W->inv_m, ii = W->inv_i;
  double r1x, r1y, r2x, r2y, K[6];
  rot(b1, j->a1x, j->a1y, &r1x, &r1y);
  rot(b2, j->a2x, j->a2y, &r2x, &r2y);
  double C[3];
  C[0] = (b1->x + r1x) - (b2->x + r2x);
  C[1] = (b1->y + r1y) - (b2->y + r2y);
  C[2] = b1->a - b2->a - j->ref_angle;
  double lin_err = sqrt(C[0] * C[0] + C[1] * C[1]), ang_err = fabs(C[2]);
  weld_K(W, r1x, r1y, r2x, r2y, K);
  double nb[3] = {-C[0], -C[1], -C[2]}, l[3];
  solve33(K, nb, l);
  b1->x += im * l[0]; b1->y += im * l[1]; body_set_angle(b1, b1->a + ii * (r1x * l[1] - r1y * l[0] + l[2]));
  b2->x -= im * l[0]; b2->y -= im * l[1]; body_set_angle(b2, b2->a - ii * (r2x * l[1] - r2y * l[0] + l[2]));
  W->tally->n_weld_pos++;
  return lin_err <= W->p.linear_tol && ang_err <= W->p.angular_tol;
}

/* ------------------------------------------------------------------------------------------- */
/* contacts (ground is static: only the corner mass moves)                                      */
/* ------------------------------------------------------------------------------------------- */
static void contacts_init(World* W) {
  double im = W->inv_m, ii = W->inv_i;
  for (int i = 0; i < W->n_bodies; i++) {
    Body* b = &W->bodies[i];
    for (int k = 0; k < W->n_man[i]; k++) {
      Manifold* m = &W->man[i * MAXM + k];
      double nx = m->nx, ny = m->ny, tx = ny, ty = -nx;
      for (int q = 0; q < m->npts; q++) {
        CPoint* c = &m->pt[q];
        c->rx = c->px - b->x;
        c->ry = c->py - b->y;
        double rn = c->rx * ny - c->ry * nx, rt = c->rx * ty - c->ry * tx;
        c->massN = 1.0 / (im + ii * rn * rn);
        c->massT = 1.0 / (im + ii * rt * rt);
        double vx = b->vx - b->w * c->ry, vy = b->vy + b->w * c->rx;
        double vn = nx * vx + ny * vy;
        c->vb = vn < -W->p.restitution_velocity ? -W->rest_e * vn : 0.0;
      }
      m->block = 0;
      if (m->npts == 2) {
        double rn1 = m->pt[0].rx * ny - m->pt[0].ry * nx, rn2 = m->pt[1].rx * ny - m->pt[1].ry * nx;
        m->K11 = im + ii * rn1 * rn1;
        m->K22 = im + ii * rn2 * rn2;
        m->K12 = im + ii * rn1 * rn2;
        if (m->K11 * m->K11 < 100.0 * (m->K11 * m->K22 - m->K12 * m->K12)) m->block = 1;
        else m->npts = 1; /* ill conditioned: keep the first point only */
      }
    }
  }
  /* warm start */
  for (int i = 0; i < W->n_bodies; i++) {
    Body* b = &W->bodies[i];
    for (int k = 0; k < W->n_man[i]; k++) {
      Manifold* m = &W->man[i * MAXM + k];
      double nx = m->nx, ny = m->ny, tx = ny, ty = -nx;
      for (int q = 0; q < m->npts; q++) {
        CPoint* c = &m->pt[q];
        double Jx = nx * c->jn + tx * c->jt, Jy = ny * c->jn + ty * c->jt;
        b->vx += im * Jx; b->vy += im * Jy; b->w += ii * (c->rx * Jy - c->ry * Jx);
      }
    }
  }
}

static void contacts_solve_vel(World* W) {
  double im = W->inv_m, ii = W->inv_i, mu = W->mu;
  for (int i = 0; i < W->n_bodies; i++) {
    Body* b = &W->bodies[i];
    for (int k = 0; k < W->n_man[i]; k++) {
      Manifold* m = &W->man[i * MAXM + k];
      double nx = m->nx, ny = m->ny, tx = ny, ty = -nx;
      /* friction first */
      for (int q = 0; q < m->npts; q++) {
        CPoint* c = &m->pt[q];
        double vx = b->vx - b->w * c->ry, vy = b->vy + b->w * c->rx;
        double l = -c->massT * (tx * vx + ty * vy);
        double maxf = mu * c->jn;
        double nj = fmax(-maxf, fmin(c->jt + l, maxf));
        l = nj - c->jt;
        c->jt = nj;
        double Jx = tx * l, Jy = ty * l;
        b->vx += im * Jx; b->vy += im * Jy; b->w += ii * (c->rx * Jy - c->ry * Jx);
        W->tally->n_contact_pt_vel++;
      }
      if (!m->block) {
        for (int q = 0; q < m->npts; q++) {
          CPoint* c = &m->pt[q];
          double vx = b->vx - b->w * c->ry, vy = b->vy + b->w * c->rx;
          double l = -c->massN * (nx * vx + ny * vy - c->vb);
          double nj = fmax(c->jn + l, 0.0);
          l = nj - c->jn;
          c->jn = nj;
          double Jx = nx * l, Jy = ny * l;
          b->vx += im * Jx; b->vy += im * Jy; b->w += ii * (c->rx * Jy - c->ry * Jx);
        }
      } else {
        /* 2-point block solver (Box2D b2ContactSolver / dyn4j SequentialImpulses) */
        CPoint *c1 = &m->pt[0], *c2 = &m->pt[1];
        double a1 = c1->jn, a2 = c2->jn;
        double v1x = b->vx - b->w * c1->ry, v1y = b->vy + b->w * c1->rx;
        double v2x = b->vx - b->w * c2->ry, v2y = b->vy + b->w * c2->rx;
        double vn1 = nx * v1x + ny * v1y, vn2 = nx * v2x + ny * v2y;
        double b1 = vn1 - c1->vb - (m->K11 * a1 + m->K12 * a2);
        double b2 = vn2 - c2->vb - (m->K12 * a1 + m->K22 * a2);
        double x1, x2;
        double det = m->K11 * m->K22 - m->K12 * m->K12;
        if (det != 0.0) det = 1.0 / det;
        x1 = -det * (m->K22 * b1 - m->K12 * b2);
        x2 = -det * (m->K11 * b2 - m->K12 * b1);
        if (!(x1 >= 0.0 && x2 >= 0.0)) {
          x1 = -b1 / m->K11; x2 = 0.0;
          double vn2n = m->K12 * x1 + b2;
          if (!(x1 >= 0.0 && vn2n >= 0.0)) {
            x1 = 0.0; x2 = -b2 / m->K22;
            double vn1n = m->K12 * x2 + b1;
            if (!(x2 >= 0.0 && vn1n >= 0.0)) {
              x1 = 0.0; x2 = 0.0;
              if (!(b1 >= 0.0 && b2 >= 0.0)) { x1 = a1; x2 = a2; } /* no solution: leave unchanged */
            }
          }
        }
        double d1 = x1 - a1, d2 = x2 - a2;
        c1->jn = x1; c2->jn = x2;
        double J1x = nx * d1, J1y = ny * d1, J2x = nx * d2, J2y = ny * d2;
        b->vx += im * (J1x + J2x); b->vy += im * (J1y + J2y);
        b->w += ii * ((c1->rx * J1y - c1->ry * J1x) + (c2->rx * J2y - c2->ry * J2x));
        W->tally->n_block_vel++;
      }
    }
  }
}

static int contacts_solve_pos(World* W) {
  double im = W->inv_m, ii = W->inv_i;
  double min_sep = 0.0;
  for (int i = 0; i < W->n_bodies; i++) {
    Body* b = &W->bodies[i];
    for (int k = 0; k < W->n_man[i]; k++) {
      Manifold* m = &W->man[i * MAXM + k];
      double nx = m->nx, ny = m->ny;
      for (int q = 0; q < m->npts; q++) {
        CPoint* c = &m->pt[q];
        double rx, ry;
        rot(b, c->lx, c->ly, &rx, &ry);
        double sep = nx * (b->x + rx - c->px) + ny * (b->y + ry - c->py) - c->depth;
        if (sep < min_sep) min_sep = sep;
        double C = W->p.baumgarte * (sep + W->p.linear_tol);
        if (C > 0) C = 0;
        if (C < -W->p.max_linear_corr) C = -W->p.max_linear_corr;
        double rn = rx * ny - ry * nx;
        double K = im + ii * rn * rn;
        double l = K > 0 ? -C / K : 0.0;
        b->x += im * nx * l; b->y += im * ny * l;
        body_set_angle(b, b->a + ii * rn * l);
        W->tally->n_contact_pt_pos++;
      }
    }
  }
  return min_sep >= -3.0 * W->p.linear_tol;
}

/* ------------------------------------------------------------------------------------------- */
/* World.step(1)                                                                                */
/* ------------------------------------------------------------------------------------------- */
static void world_step(World* W) {
  const vo_params* p = &W->p;
  double dt = p->dt;
  /* integrate velocities: gravity then damping (Island.solve) */
  double lin = 1.0 - dt * p->lin_damping, ang = 1.0 - dt * p->ang_damping;
  lin = fmin(fmax(lin, 0.0), 1.0);
  ang = fmin(fmax(ang, 0.0), 1.0);
  for (int i = 0; i < W->n_bodies; i++) {
    Body* b = &W->bodies[i];
    b->vy += p->gravity_y * dt;
    b->vx *= lin; b->vy *= lin; b->w *= ang;
  }
  contacts_init(W);
  for (int i = 0; i < W->n_joints; i++) {
    Joint* j = &W->joints[i];
    if (j->type == J_SPRING) spring_init(W, j);
    else if (j->type == J_ROPE) rope_init(W, j);
    else weld_init(W, j);
  }
  for (int it = 0; it < p->vel_iters; it++) {
    for (int i = 0; i < W->n_joints; i++) {
      Joint* j = &W->joints[i];
      if (j->type == J_SPRING) spring_solve_vel(W, j);
      else if (j->type == J_ROPE) rope_solve_vel(W, j);
      else weld_solve_vel(W, j);
    }
    contacts_solve_vel(W);
  }
  /* integrate positions with the per-step clamps */
  for (int i = 0; i < W->n_bodies; i++) {
    Body* b = &W->bodies[i];
    double tx = b->vx * dt, ty = b->vy * dt, r = b->w * dt;
    double t2 = tx * tx + ty * ty;
    if (t2 > p->max_translation * p->max_translation) {
      double ratio = p->max_translation / sqrt(t2);
      b->vx *= ratio; b->vy *= ratio;
      tx = b->vx * dt; ty = b->vy * dt;
    }
    if (fabs(r) > p->max_rotation) {
      double ratio = p->max_rotation / fabs(r);
      b->w *= ratio;
      r = b->w * dt;
    }
    b->x += tx; b->y += ty;
    body_set_angle(b, b->a + r);
  }
  for (int it = 0; it < p->pos_iters; it++) {
    W->tally->n_pos_iters++;
    int cs = contacts_solve_pos(W);
    int js = 1;
    for (int i = 0; i < W->n_joints; i++) {
      Joint* j = &W->joints[i];
      if (j->type == J_ROPE) js &= rope_solve_pos(W, j);
      else if (j->type == J_WELD) js &= weld_solve_pos(W, j);
    }
    if (cs && js) break;
  }
  detect(W);
}

/* ------------------------------------------------------------------------------------------- */
/* voxel geometry, sensors, controllers                                                         */
/* ------------------------------------------------------------------------------------------- */
/* Voxel.getVoxelPoly(): outer corners (body k, rectangle vertex 3-k) */
static void voxel_vertices(const World* W, int v, double out[4][2]) {
  static const double SX[4] = {-1, +1, +1, -1}, SY[4] = {+1, +1, -1, -1}; /* NW NE SE SW outer corner */
  for (int k = 0; k < 4; k++) {
    const Body* b = &W->bodies[v * 4 + k];
    double lx = SX[k] * W->half, ly = SY[k] * W->half;
    out[k][0] = b->x + b->c * lx - b->s * ly;
    out[k][1] = b->y + b->s * lx + b->c * ly;
  }
}

static double voxel_area_ratio(const World* W, int v) {
  double q[4][2];
  voxel_vertices(W, v, q);
  double a = 0;
  for (int i = 0; i < 4; i++) {
    int j = (i + 1) & 3;
    a += q[i][0] * q[j][1] - q[j][0] * q[i][1];
  }
  return fabs(a) * 0.5 / (W->p.side_length * W->p.side_length);
}

static double voxel_angle(const World* W, int v) {
  const Body* b = &W->bodies[v * 4];
  double up = atan2(b[1].y - b[0].y, b[1].x - b[0].x);
  double dn = atan2(b[2].y - b[3].y, b[2].x - b[3].x);
  return (up + dn) / 2.0;
}

static double robot_center_x(const World* W) {
  /* mean over voxels of the voxel polygon centre (mean of its 4 vertices) */
  double s = 0;
  for (int v = 0; v < W->n_vox; v++) {
    double q[4][2];
    voxel_vertices(W, v, q);
    s += (q[0][0] + q[1][0] + q[2][0] + q[3][0]) / 4.0;
  }
  return s / W->n_vox;
}

/* java.util.Random (LCG + polar nextGaussian): the Noisy sensor wrapper draws from `new Random(seed)`; 2dhmsr's
 * RobotUtils.buildSensorizingFunction builds every wrapper as `new Noisy(sensor, sigma, 0)` [U], so all voxels see the
 * same sequence: one stream per sensor of the per-voxel list. */
typedef struct { uint64_t seed; int have; double next_g; } JRandom;
static void jr_init(JRandom* r, int64_t seed) { r->seed = ((uint64_t)seed ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1); r->have = 0; r->next_g = 0; }
static int32_t jr_next(JRandom* r, int bits) {
  r->seed = (r->seed * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
  return (int32_t)((int64_t)r->seed >> (48 - bits));
}
static double jr_double(JRandom* r) { return (double)(((int64_t)jr_next(r, 26) << 27) + jr_next(r, 27)) * (1.0 / (double)(1LL << 53)); }
static double jr_gaussian(JRandom* r) {
  if (r->have) { r->have = 0; return r->next_g; }
  double v1, v2, s;
  do { v1 = 2 * jr_double(r) - 1; v2 = 2 * jr_double(r) - 1; s = v1 * v1 + v2 * v2; } while (s >= 1 || s == 0);
  double mul = sqrt(-2 * log(s) / s);
  r->next_g = v2 * mul; r->have = 1;
  return v1 * mul;
}

typedef struct {
  int n_vox, n_inst;       /* sensor instances over all voxels */
  vo_sensor* list;         /* [n_inst] the voxels' sensor lists, concatenated in voxel (grid) order */
  int* voff;               /* [n_vox + 1] first instance of each voxel */
  int* roff;               /* [n_inst + 1] first reading of each instance */
  int* vroff;              /* [n_vox + 1] first reading of each voxel */
  int total;               /* readings of the whole robot */
  JRandom* rng;            /* Noisy: every wrapper owns `new Random(0)` [U]: one stream per instance */
  double* readings;        /* [total] */
  /* Average windows: ring of (t, value) per reading */
  int hist_cap;
  double* hist_t;          /* [hist_cap] */
  double* hist_v;          /* [hist_cap][total] raw readings */
  int hist_n;              /* number of steps stored so far */
} Sensing;

static int sensor_dim(const vo_sensor* s) {
  if (s->kind == VO_SENSOR_VELOCITY) return ((s->axes & 1) ? 1 : 0) + ((s->axes & 2) ? 1 : 0);
  if (s->kind == VO_SENSOR_LIDAR) return s->n_rays;
  return 1;
}

static double soft_norm(double r, double lo, double hi) {
  double v = (r - lo) / (hi - lo);
  return tanh(((v * 2.0) - 1.0) * 2.0) / 2.0 + 0.5;
}

/* Lidar [U]: Side sectors N (pi/4, 3pi/4), E (-pi/4, pi/4), S (5pi/4, 7pi/4), W (3pi/4, 5pi/4); one ray sits at (end - start) / 2
 * (the upstream sampleRangeWithRays quirk), n > 1 rays are spread from start to end inclusive */
static double lidar_direction(int side, int n, int i) {
  static const double START[4] = {M_PI * 1.0 / 4.0, M_PI * -1.0 / 4.0, M_PI * 5.0 / 4.0, M_PI * 3.0 / 4.0};
  static const double END[4] = {M_PI * 3.0 / 4.0, M_PI * 1.0 / 4.0, M_PI * 7.0 / 4.0, M_PI * 5.0 / 4.0};
  if (n == 1) return (END[side] - START[side]) / 2.0;
  return START[side] + (END[side] - START[side]) / (double)(n - 1) * (double)i;
}

/* distance along the ray (ox, oy) + t (dx, dy), |d| = 1, to the ground's surface (the polyline, closed by the two outer walls), or
 * max_len when nothing is hit within max_len.  The rays see the ground only (the robot's own voxels are filtered out [U]). */
static double raycast_ground(const World* W, double ox, double oy, double dx, double dy, double max_len) {
  double best = max_len;
  for (int k = 1; k < W->n_t; k++) {
    const double x0 = W->tx[k - 1], y0 = W->ty[k - 1], ex = W->tx[k] - x0, ey = W->ty[k] - y0;
    if (fmax(x0, x0 + ex) < fmin(ox, ox + dx * max_len) || fmin(x0, x0 + ex) > fmax(ox, ox + dx * max_len)) continue;
    const double den = dx * ey - dy * ex;
    if (den == 0.0) continue;
    const double t = ((x0 - ox) * ey - (y0 - oy) * ex) / den, u = ((x0 - ox) * dy - (y0 - oy) * dx) / den;
    if (t >= 0.0 && t < best && u >= 0.0 && u <= 1.0) best = t;
  }
  for (int side = 0; side < 2; side++) { /* outer walls of the first / last ground polygon, from the surface down to the base */
    const double xw = side == 0 ? W->tx[0] : W->tx[W->n_t - 1], yt = side == 0 ? W->ty[0] : W->ty[W->n_t - 1];
    if (dx == 0.0) continue;
    const double t = (xw - ox) / dx, y = oy + t * dy;
    if (t >= 0.0 && t < best && y <= yt && y >= W->base_y) best = t;
  }
  return best;
}

static void sense(World* W, Sensing* S, double t) {
  /* raw readings of this step go in the ring */
  int slot = S->hist_n % S->hist_cap;
  S->hist_t[slot] = t;
  double* raw = &S->hist_v[(size_t)slot * S->total];
  for (int v = 0; v < W->n_vox; v++)
    for (int si = S->voff[v]; si < S->voff[v + 1]; si++) {
      const vo_sensor* s = &S->list[si];
      double* r = &raw[S->roff[si]];
      switch (s->kind) {
        case VO_SENSOR_AREA_RATIO: r[0] = voxel_area_ratio(W, v); break;
        case VO_SENSOR_TOUCH: {
          int touch = 0;
          for (int k = 0; k < 4; k++) touch |= W->n_man[v * 4 + k] > 0;
          r[0] = touch ? 1.0 : 0.0;
        } break;
        case VO_SENSOR_ANGLE: r[0] = voxel_angle(W, v); break;
        case VO_SENSOR_VELOCITY: {
          double vx = 0, vy = 0;
          for (int k = 0; k < 4; k++) { vx += W->bodies[v * 4 + k].vx; vy += W->bodies[v * 4 + k].vy; }
          vx /= 4.0; vy /= 4.0;
          if (s->rotated) {
            double a = -voxel_angle(W, v), c = cos(a), sn = sin(a);
            double rx = c * vx - sn * vy, ry = sn * vx + c * vy;
            vx = rx; vy = ry;
          }
          int o = 0;
          if (s->axes & 1) r[o++] = vx;
          if (s->axes & 2) r[o++] = vy;
        } break;
        case VO_SENSOR_CONSTANT: r[0] = s->param; break;
        case VO_SENSOR_TIME_SIN: r[0] = sin(2.0 * M_PI * s->param * t); break;
        case VO_SENSOR_LIDAR: {
          double cx = 0, cy = 0;
          for (int k = 0; k < 4; k++) { cx += W->bodies[v * 4 + k].x; cy += W->bodies[v * 4 + k].y; }
          cx /= 4.0; cy /= 4.0;
          const double va = voxel_angle(W, v);
          for (int i = 0; i < s->n_rays; i++) {
            const double dir = lidar_direction(s->side, s->n_rays, i) + va;
            r[i] = raycast_ground(W, cx, cy, cos(dir), sin(dir), s->param);
          }
        } break;
      }
    }
  S->hist_n++;
  for (int si = 0; si < S->n_inst; si++) {
    const vo_sensor* s = &S->list[si];
    const int dim = sensor_dim(s);
    for (int d = 0; d < dim; d++) {
      const int idx = S->roff[si] + d;
      double val = raw[idx];
      if (s->avg_window > 0) {
        /* Average: mean of stored readings with t_j >= t - window */
        double sum = 0; int n = 0;
        int have = S->hist_n < S->hist_cap ? S->hist_n : S->hist_cap;
        for (int back = 0; back < have; back++) {
          int sl = (S->hist_n - 1 - back) % S->hist_cap;
          if (S->hist_t[sl] < t - s->avg_window) break;
          sum += S->hist_v[(size_t)sl * S->total + idx];
          n++;
        }
        val = sum / n;
      }
      if (s->soft_norm == 1) val = soft_norm(val, s->dom_min, s->dom_max);
      else if (s->soft_norm == 2) val = fmin(fmax((val - s->dom_min) / (s->dom_max - s->dom_min), 0.0), 1.0);
      /* Noisy.sense: value + nextGaussian() * sigma * |domain| of the wrapped sensor */
      if (s->noise_sigma > 0) val += jr_gaussian(&S->rng[si]) * s->noise_sigma * (s->soft_norm ? 1.0 : fabs(s->dom_max - s->dom_min));
      S->readings[idx] = val;
    }
  }
}

static void mlp_apply(const vo_ctrl* c, const double* wts, const double* in, double* out) {
  double a[2][256];
  int cur = 0;
  for (int i = 0; i < c->layers[0]; i++) a[0][i] = c->input_activation ? tanh(in[i]) : in[i];
  int o = 0;
  for (int l = 1; l < c->n_layers; l++) {
    int nf = c->layers[l - 1], nt = c->layers[l];
    for (int j = 0; j < nt; j++) {
      double sum = wts[o++]; /* bias first */
      for (int k = 0; k < nf; k++) sum += a[cur][k] * wts[o++];
      a[1 - cur][j] = tanh(sum);
    }
    cur = 1 - cur;
  }
  for (int j = 0; j < c->layers[c->n_layers - 1]; j++) out[j] = a[cur][j];
}

static int mlp_n_weights(const vo_ctrl* c) {
  int n = 0;
  for (int l = 1; l < c->n_layers; l++) n += c->layers[l] * (c->layers[l - 1] + 1);
  return n;
}

static int genome_len_dims(int nv, int total_dim, int dim, const vo_ctrl* ctrl);

int vo_genome_len(int w, int h, const uint8_t* mask, const vo_sensor* sensors, int n_sensors, const vo_ctrl* ctrl) {
  int nv = 0;
  for (int i = 0; i < w * h; i++) nv += mask[i] ? 1 : 0;
  int dim = 0;
  for (int i = 0; i < n_sensors; i++) dim += sensor_dim(&sensors[i]);
  return genome_len_dims(nv, nv * dim, dim, ctrl);
}

int vo_genome_len_vox(int w, int h, const uint8_t* mask, const vo_sensor* vox_sensors, const int* vox_sensor_off, const vo_ctrl* ctrl) {
  int nv = 0;
  for (int i = 0; i < w * h; i++) nv += mask[i] ? 1 : 0;
  int total = 0, dim0 = -1, uniform = 1;
  for (int v = 0; v < nv; v++) {
    int d = 0;
    for (int i = vox_sensor_off[v]; i < vox_sensor_off[v + 1]; i++) d += sensor_dim(&vox_sensors[i]);
    if (dim0 < 0) dim0 = d;
    if (d != dim0) uniform = 0;
    total += d;
  }
  return genome_len_dims(nv, total, uniform ? dim0 : -1, ctrl);
}

/* dim = readings per voxel, or -1 when the voxels differ (distributed controllers then have no single MLP input size) */
static int genome_len_dims(int nv, int total_dim, int dim, const vo_ctrl* ctrl) {
  if ((ctrl->kind == VO_CTRL_MLP_DISTRIBUTED_HOMO || ctrl->kind == VO_CTRL_MLP_DISTRIBUTED_HETERO) && dim < 0) return -2;
  switch (ctrl->kind) {
    case VO_CTRL_PHASE_SIN: return nv;
    case VO_CTRL_MLP_CENTRALIZED:
      if (ctrl->layers[0] != total_dim || ctrl->layers[ctrl->n_layers - 1] != nv) return -2;
      return mlp_n_weights(ctrl);
    case VO_CTRL_MLP_DISTRIBUTED_HOMO:
      if (ctrl->layers[0] != dim + 4 * ctrl->signals || ctrl->layers[ctrl->n_layers - 1] != 1 + 4 * ctrl->signals) return -2;
      return mlp_n_weights(ctrl);
    case VO_CTRL_MLP_DISTRIBUTED_HETERO:
      if (ctrl->layers[0] != dim + 4 * ctrl->signals || ctrl->layers[ctrl->n_layers - 1] != 1 + 4 * ctrl->signals) return -2;
      return nv * mlp_n_weights(ctrl);
    case VO_CTRL_TIME_SIN: return 3 * nv;
  }
  return -1;
}

/* Voxel.applyForce(f), ForceMethod.DISTANCE */
static void apply_control(World* W, int v, double f) {
  if ((W->frozen >> v) & 1u) f = 0.0; /* BreakableVoxel, ACTUATOR FROZEN since t = 0: the last applied force (0) is applied instead */
  if (fabs(f) > 1.0) f = f > 0 ? 1.0 : -1.0;
  W->ctrl_f[v] = f;
  for (int i = 0; i < W->n_joints; i++) {
    Joint* j = &W->joints[i];
    if (j->type != J_SPRING || j->voxel != v) continue;
    int k = j->range_kind;
    if (f >= 0) j->rest = W->range_rest[k] - (W->range_rest[k] - W->range_min[k]) * f;
    else j->rest = W->range_rest[k] + (W->range_max[k] - W->range_rest[k]) * -f;
  }
}

/* ------------------------------------------------------------------------------------------- */
/* episode                                                                                      */
/* ------------------------------------------------------------------------------------------- */
static void derive(World* W) {
  const vo_params* p = &W->p;
  double L = p->side_length;
  W->m_side = L * p->mass_side_ratio;
  W->half = W->m_side / 2;
  double m = W->m_side;
  double density = (p->voxel_mass / 4) / (m * m);
  double mass = density * m * m;
  double inertia = mass * (m * m + m * m) / 12.0;
  W->inv_m = 1.0 / mass;
  W->inv_i = 1.0 / inertia;
  W->mu = sqrt(p->voxel_friction * p->ground_friction);   /* dyn4j CoefficientMixer default */
  W->rest_e = fmax(p->voxel_restitution, p->ground_restitution);
  double min_side = sqrt(L * L * (1.0 - p->area_ratio_max_delta));
  double max_side = sqrt(L * L * (1.0 + p->area_ratio_max_delta));
  W->range_min[R_PARALLEL] = min_side - 2 * m;
  W->range_rest[R_PARALLEL] = L - 2 * m;
  W->range_max[R_PARALLEL] = max_side - 2 * m;
  for (int q = 0; q < 1; q++) {
    W->range_min[R_CROSS] = sqrt(m * m + W->range_min[R_PARALLEL] * W->range_min[R_PARALLEL]);
    W->range_rest[R_CROSS] = sqrt(m * m + W->range_rest[R_PARALLEL] * W->range_rest[R_PARALLEL]);
    W->range_max[R_CROSS] = sqrt(m * m + W->range_max[R_PARALLEL] * W->range_max[R_PARALLEL]);
  }
  W->range_min[R_CENTRAL] = (min_side - m) * sqrt(2.0);
  W->range_rest[R_CENTRAL] = (L - m) * sqrt(2.0);
  W->range_max[R_CENTRAL] = (max_side - m) * sqrt(2.0);
}

static void world_free(World* W) {
  free(W->vox_of_cell); free(W->cell_x); free(W->cell_y); free(W->bodies); free(W->joints);
  free(W->tx); free(W->ty); free(W->man); free(W->n_man); free(W->ctrl_f);
}

/* ------------------------------------------------------------------------------------------- */
/* per-observation summaries (BehaviorUtils [U]; see vo_extras in the header)                    */
/* ------------------------------------------------------------------------------------------- */
int vo_central_voxel(int w, int h, const uint8_t* mask) {
  double mx = 0, my = 0;
  int n = 0;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++)
      if (mask[y * w + x]) { mx += x; my += y; n++; }
  if (!n) return -1;
  mx /= n; my /= n;
  double best = INFINITY;
  int bi = -1, v = 0;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++)
      if (mask[y * w + x]) {
        double d = (x - mx) * (x - mx) + (y - my) * (y - my);
        if (d < best) { best = d; bi = v; } /* first minimum in grid iteration order */
        v++;
      }
  return bi;
}

/* crossing-number point-in-polygon test (what java.awt Path2D.contains decides for a simple polygon) */
static int quad_contains(const double q[4][2], double px, double py) {
  int in = 0;
  for (int i = 0, j = 3; i < 4; j = i++) {
    const double xi = q[i][0], yi = q[i][1], xj = q[j][0], yj = q[j][1];
    if (((yi > py) != (yj > py)) && (px < (xj - xi) * (py - yi) / (yj - yi) + xi)) in = !in;
  }
  return in;
}

/* BehaviorUtils.computePosture(polies, n): cells of the squared bounding box whose centre lies inside a voxel polygon */
static uint64_t posture_mask(const World* W, int n) {
  double minx = INFINITY, miny = INFINITY, maxx = -INFINITY, maxy = -INFINITY;
  for (int v = 0; v < W->n_vox; v++) {
    double q[4][2];
    voxel_vertices(W, v, q);
    for (int k = 0; k < 4; k++) {
      minx = fmin(minx, q[k][0]); maxx = fmax(maxx, q[k][0]);
      miny = fmin(miny, q[k][1]); maxy = fmax(maxy, q[k][1]);
    }
  }
  /* adjust the box to make it squared (centred) */
  const double wdt = maxx - minx, hgt = maxy - miny;
  if (hgt < wdt) { const double d = wdt - hgt; miny -= d / 2; maxy += d / 2; }
  else if (hgt > wdt) { const double d = hgt - wdt; minx -= d / 2; maxx += d / 2; }
  const double side = maxx - minx;
  uint64_t m = 0;
  for (int v = 0; v < W->n_vox; v++) {
    double q[4][2];
    voxel_vertices(W, v, q);
    for (int iy = 0; iy < n; iy++)
      for (int ix = 0; ix < n; ix++) {
        const double px = minx + (ix + 0.5) / n * side, py = miny + (iy + 0.5) / n * side;
        if (quad_contains(q, px, py)) m |= 1ULL << (iy * n + ix);
      }
  }
  return m;
}

/* BehaviorUtils.computeFootprint(polies, n) */
static uint8_t footprint_mask(const World* W, int n) {
  double rmin = INFINITY, rmax = -INFINITY;
  for (int v = 0; v < W->n_vox; v++) {
    double q[4][2];
    voxel_vertices(W, v, q);
    for (int k = 0; k < 4; k++) { rmin = fmin(rmin, q[k][0]); rmax = fmax(rmax, q[k][0]); }
  }
  uint8_t m = 0;
  for (int v = 0; v < W->n_vox; v++) {
    int touch = 0;
    for (int k = 0; k < 4; k++) touch |= W->n_man[v * 4 + k] > 0;
    if (!touch) continue;
    double q[4][2];
    voxel_vertices(W, v, q);
    double cmin = INFINITY, cmax = -INFINITY;
    for (int k = 0; k < 4; k++) { cmin = fmin(cmin, q[k][0]); cmax = fmax(cmax, q[k][0]); }
    const int i0 = (int)floor((cmin - rmin) / (rmax - rmin) * (double)(n - 1) + 0.5); /* Math.round */
    int i1 = (int)floor((cmax - rmin) / (rmax - rmin) * (double)(n - 1) + 0.5);
    if (i1 > n - 1) i1 = n - 1;
    for (int x = i0 < 0 ? 0 : i0; x <= i1; x++) m |= (uint8_t)(1u << x);
  }
  return m;
}

/* VoxelPoly.getAngle(): mean direction of the polygon's upper (NW -> NE) and lower (SW -> SE) sides */
static double voxel_poly_angle(const World* W, int v) {
  double q[4][2];
  voxel_vertices(W, v, q);
  return (atan2(q[1][1] - q[0][1], q[1][0] - q[0][0]) + atan2(q[2][1] - q[3][1], q[2][0] - q[3][0])) / 2.0;
}

static double robot_center_y(const World* W) {
  double s = 0;
  for (int v = 0; v < W->n_vox; v++) {
    double q[4][2];
    voxel_vertices(W, v, q);
    s += (q[0][1] + q[1][1] + q[2][1] + q[3][1]) / 4.0;
  }
  return s / W->n_vox;
}

int vo_run(const vo_params* p, int w, int h, const uint8_t* mask,
           const vo_sensor* sensors, int n_sensors, const vo_ctrl* ctrl,
           const double* terrain_x, const double* terrain_y, int n_terrain,
           const double* genome, int genome_len,
           int trace_steps, double* trace_vertices, double* trace_bodies, double* trace_ctrl,
           vo_result* out) {
  return vo_run_ex(p, w, h, mask, sensors, n_sensors, ctrl, terrain_x, terrain_y, n_terrain, genome, genome_len, trace_steps,
                   trace_vertices, trace_bodies, trace_ctrl, NULL, out);
}

int vo_run_ex(const vo_params* p, int w, int h, const uint8_t* mask,
              const vo_sensor* sensors, int n_sensors, const vo_ctrl* ctrl,
              const double* terrain_x, const double* terrain_y, int n_terrain,
              const double* genome, int genome_len,
              int trace_steps, double* trace_vertices, double* trace_bodies, double* trace_ctrl,
              const vo_extras* ex, vo_result* out) {
  if (!p || !mask || !ctrl || !out || w <= 0 || h <= 0 || n_terrain < 2) return -1;
  if (n_sensors > 8) return -3;
  const int per_voxel = ex && ex->vox_sensor_off && ex->vox_sensors;
  if (ex && (ex->posture_n < 0 || ex->posture_n > 8 || ex->footprint_n[0] < 0 || ex->footprint_n[0] > 8 || ex->footprint_n[1] < 0 ||
             ex->footprint_n[1] > 8)) return -6;
  int glen = per_voxel ? vo_genome_len_vox(w, h, mask, ex->vox_sensors, ex->vox_sensor_off, ctrl)
                       : vo_genome_len(w, h, mask, sensors, n_sensors, ctrl);
  if (glen < 0) return glen;
  if (glen != genome_len) return -4;
  memset(out, 0, sizeof(*out));
  World Wd, *W = &Wd;
  memset(W, 0, sizeof(*W));
  W->p = *p;
  W->tally = out;
  W->frozen = ex ? ex->frozen_mask : 0u;
  derive(W);
  W->w = w; W->h = h;
  int ns, nr, nw;
  count_joints(p, w, h, mask, &W->n_vox, &ns, &nr, &nw);
  if (W->n_vox == 0) return -5;
  W->n_bodies = W->n_vox * 4;
  W->vox_of_cell = (int*)malloc(sizeof(int) * w * h);
  W->cell_x = (int*)malloc(sizeof(int) * W->n_vox);
  W->cell_y = (int*)malloc(sizeof(int) * W->n_vox);
  W->bodies = (Body*)calloc(W->n_bodies, sizeof(Body));
  W->joints = (Joint*)calloc(ns + nr + nw, sizeof(Joint));
  W->n_t = n_terrain;
  W->tx = (double*)malloc(sizeof(double) * n_terrain);
  W->ty = (double*)malloc(sizeof(double) * n_terrain);
  memcpy(W->tx, terrain_x, sizeof(double) * n_terrain);
  memcpy(W->ty, terrain_y, sizeof(double) * n_terrain);
  double miny = INFINITY;
  for (int i = 0; i < n_terrain; i++) miny = fmin(miny, terrain_y[i]);
  W->base_y = miny - p->ground_thickness;
  W->man = (Manifold*)calloc((size_t)W->n_bodies * MAXM, sizeof(Manifold));
  W->n_man = (int*)calloc(W->n_bodies, sizeof(int));
  W->ctrl_f = (double*)calloc(W->n_vox, sizeof(double));
  build_robot(W, mask);
  place_robot(W);
  detect(W); /* World.step's first-step detect() */

  /* sensing state: the voxels' sensor lists (the uniform list repeated for every voxel, or the per-voxel lists of vo_extras) */
  Sensing S;
  memset(&S, 0, sizeof(S));
  S.n_vox = W->n_vox;
  S.voff = (int*)malloc(sizeof(int) * (W->n_vox + 1));
  S.vroff = (int*)malloc(sizeof(int) * (W->n_vox + 1));
  for (int v = 0; v <= W->n_vox; v++) S.voff[v] = per_voxel ? ex->vox_sensor_off[v] - ex->vox_sensor_off[0] : v * n_sensors;
  S.n_inst = S.voff[W->n_vox];
  S.list = (vo_sensor*)malloc(sizeof(vo_sensor) * (S.n_inst > 0 ? S.n_inst : 1));
  for (int v = 0; v < W->n_vox; v++)
    for (int i = S.voff[v]; i < S.voff[v + 1]; i++)
      S.list[i] = per_voxel ? ex->vox_sensors[ex->vox_sensor_off[0] + i] : sensors[i - S.voff[v]];
  S.roff = (int*)malloc(sizeof(int) * (S.n_inst + 1));
  S.rng = (JRandom*)malloc(sizeof(JRandom) * (S.n_inst > 0 ? S.n_inst : 1));
  double max_window = 0;
  S.total = 0;
  for (int i = 0; i < S.n_inst; i++) {
    S.roff[i] = S.total;
    S.total += sensor_dim(&S.list[i]);
    jr_init(&S.rng[i], 0);
    if (S.list[i].avg_window > max_window) max_window = S.list[i].avg_window;
  }
  S.roff[S.n_inst] = S.total;
  for (int v = 0; v <= W->n_vox; v++) S.vroff[v] = S.roff[S.voff[v]];
  S.hist_cap = (int)ceil(max_window / p->dt) + 4;
  int D = S.total > 0 ? S.total : 1;
  S.readings = (double*)calloc((size_t)D, sizeof(double));
  S.hist_t = (double*)calloc(S.hist_cap, sizeof(double));
  S.hist_v = (double*)calloc((size_t)S.hist_cap * D, sizeof(double));
  /* distributed controller signal buffers */
  int sig = (ctrl->kind == VO_CTRL_MLP_DISTRIBUTED_HOMO || ctrl->kind == VO_CTRL_MLP_DISTRIBUTED_HETERO) ? ctrl->signals : 0;
  const int w_stride = ctrl->kind == VO_CTRL_MLP_DISTRIBUTED_HETERO ? mlp_n_weights(ctrl) : 0; /* per-voxel weight block */
  double* last_sig = (double*)calloc((size_t)W->n_vox * 4 * (sig > 0 ? sig : 1), sizeof(double));
  double* cur_sig = (double*)calloc((size_t)W->n_vox * 4 * (sig > 0 ? sig : 1), sizeof(double));

  /* episode loop: Locomotion.apply() */
  double t = 0.0, first_t = 0, last_t = 0, first_x = 0, last_x = 0;
  int k = 0, have_first = 0;
  double in[256], o[256];
  double* last_area = (double*)calloc((size_t)W->n_vox, sizeof(double));
  out->control_energy = out->area_ratio_abs = out->area_ratio_sq = 0.0;
  int post_cnt[64] = {0}, n_kept = 0;
  const int central = vo_central_voxel(w, h, mask);
  while (t < p->final_t) {
    t += p->dt;
    k++;
    world_step(W);
    /* robot.act(t): sensors, then controller, then applyForce */
    sense(W, &S, t);
    switch (ctrl->kind) {
      case VO_CTRL_PHASE_SIN:
        for (int v = 0; v < W->n_vox; v++)
          apply_control(W, v, ctrl->amp * sin(2.0 * M_PI * ctrl->freq * t + genome[v]));
        break;
      case VO_CTRL_MLP_CENTRALIZED: {
        for (int i = 0; i < S.total; i++) in[i] = S.readings[i];
        mlp_apply(ctrl, genome, in, o);
        for (int v = 0; v < W->n_vox; v++) apply_control(W, v, o[v]);
      } break;
      case VO_CTRL_TIME_SIN: /* TimeFunctions: per-voxel (frequency, phase, amplitude) */
        for (int v = 0; v < W->n_vox; v++)
          apply_control(W, v, genome[3 * v + 2] * sin(2.0 * M_PI * genome[3 * v] * t + genome[3 * v + 1]));
        break;
      case VO_CTRL_MLP_DISTRIBUTED_HOMO:
      case VO_CTRL_MLP_DISTRIBUTED_HETERO: {
        /* Dir: N(0,-1) E(1,0) S(0,1) W(-1,0); index N0 E1 S2 W3; input from neighbour's opposite side */
        static const int DX[4] = {0, 1, 0, -1}, DY[4] = {-1, 0, 1, 0}, OPP[4] = {2, 3, 0, 1};
        for (int v = 0; v < W->n_vox; v++) {
          int ni = 0;
          for (int i = S.vroff[v]; i < S.vroff[v + 1]; i++) in[ni++] = S.readings[i];
          for (int d = 0; d < 4; d++) {
            int ax = W->cell_x[v] + DX[d], ay = W->cell_y[v] + DY[d];
            int nb = (ax >= 0 && ax < w && ay >= 0 && ay < h) ? W->vox_of_cell[ay * w + ax] : -1;
            for (int s = 0; s < sig; s++) in[ni++] = nb >= 0 ? last_sig[(nb * 4 + OPP[d]) * sig + s] : 0.0;
          }
          mlp_apply(ctrl, genome + (size_t)v * w_stride, in, o);
          apply_control(W, v, o[0]);
          for (int i = 0; i < 4 * sig; i++) cur_sig[v * 4 * sig + i] = o[1 + i];
        }
        memcpy(last_sig, cur_sig, sizeof(double) * W->n_vox * 4 * (sig > 0 ? sig : 1));
      } break;
    }
    /* observation */
    double cx = robot_center_x(W);
    if (t >= p->transient_t && t < p->final_t) { /* Outcome.subOutcome: subMap(start, end) */
      /* energy increments between consecutive kept observations (the first one only sets the reference) */
      for (int v = 0; v < W->n_vox; v++) {
        const double a = voxel_area_ratio(W, v), f = W->ctrl_f[v];
        if (have_first) {
          out->control_energy += f * f * p->dt;
          out->area_ratio_abs += fabs(a - last_area[v]);
          out->area_ratio_sq += fabs(a * a - last_area[v] * last_area[v]);
        }
        last_area[v] = a;
      }
      if (!have_first) { have_first = 1; first_t = t; first_x = cx; out->k_first = k; }
      last_t = t; last_x = cx; out->k_last = k;
      if (ex) {
        if (ex->posture_n > 0) {
          const uint64_t pm = posture_mask(W, ex->posture_n);
          for (int c = 0; c < ex->posture_n * ex->posture_n; c++) post_cnt[c] += (int)((pm >> c) & 1ULL);
        }
        for (int q = 0; q < 2; q++)
          if (ex->footprint_n[q] > 0 && ex->footprints_out[q]) ex->footprints_out[q][n_kept] = footprint_mask(W, ex->footprint_n[q]);
        if (ex->centers_out) {
          ex->centers_out[3 * (size_t)n_kept] = cx;
          ex->centers_out[3 * (size_t)n_kept + 1] = robot_center_y(W);
          ex->centers_out[3 * (size_t)n_kept + 2] = voxel_poly_angle(W, central);
        }
      }
      n_kept++;
    }
    if (k <= trace_steps) {
      if (trace_vertices)
        for (int v = 0; v < W->n_vox; v++) {
          double q[4][2];
          voxel_vertices(W, v, q);
          memcpy(&trace_vertices[((size_t)(k - 1) * W->n_vox + v) * 8], q, sizeof(q));
        }
      if (trace_bodies)
        for (int i = 0; i < W->n_bodies; i++) {
          double* d = &trace_bodies[((size_t)(k - 1) * W->n_bodies + i) * 6];
          Body* b = &W->bodies[i];
          d[0] = b->x; d[1] = b->y; d[2] = b->a; d[3] = b->vx; d[4] = b->vy; d[5] = b->w;
        }
      if (trace_ctrl)
        for (int v = 0; v < W->n_vox; v++) trace_ctrl[(size_t)(k - 1) * W->n_vox + v] = W->ctrl_f[v];
    }
  }
  out->n_steps = k;
  out->first_x = first_x;
  out->last_x = last_x;
  out->distance = last_x - first_x;
  out->time = last_t - first_t;
  out->velocity = out->distance / out->time;
  if (ex && ex->posture_n > 0)
    for (int c = 0; c < ex->posture_n * ex->posture_n; c++) {
      if (ex->posture_out) ex->posture_out[c] = (uint8_t)(2 * post_cnt[c] > n_kept); /* average over the observations > 0.5 */
      if (ex->posture_counts) ex->posture_counts[c] = post_cnt[c];
    }
  free(S.list); free(S.voff); free(S.vroff); free(S.roff); free(S.rng);
  free(S.readings); free(S.hist_t); free(S.hist_v); free(last_sig); free(cur_sig); free(last_area);
  world_free(W);
  return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* batch on host threads (the CPU baseline)                                                     */
/* ------------------------------------------------------------------------------------------- */
typedef struct {
  const vo_params* p; int w, h; const uint8_t* mask; const vo_sensor* sensors; int n_sensors; const vo_ctrl* ctrl;
  const double *tx, *ty; int nt; const double* genomes; int glen, n; double* fit; vo_result* res;
  int* next; pthread_mutex_t* mu; int err;
  const vo_sensor* vox_sensors; const int* vox_sensor_off; /* per-voxel sensor lists (vo_eval_batch_vox) or NULL */
} BatchJob;

static void* batch_worker(void* arg) {
  BatchJob* J = (BatchJob*)arg;
  for (;;) {
    pthread_mutex_lock(J->mu);
    int i = (*J->next)++;
    pthread_mutex_unlock(J->mu);
    if (i >= J->n) break;
    vo_result r;
    vo_extras ex;
    memset(&ex, 0, sizeof(ex));
    ex.vox_sensors = J->vox_sensors; ex.vox_sensor_off = J->vox_sensor_off;
    int rc = vo_run_ex(J->p, J->w, J->h, J->mask, J->sensors, J->n_sensors, J->ctrl, J->tx, J->ty, J->nt,
                       J->genomes + (size_t)i * J->glen, J->glen, 0, NULL, NULL, NULL, J->vox_sensor_off ? &ex : NULL, &r);
    if (rc) { J->err = rc; J->fit[i] = NAN; continue; }
    J->fit[i] = r.velocity;
    if (J->res) J->res[i] = r;
  }
  return NULL;
}

static int eval_batch_impl(const vo_params* p, int w, int h, const uint8_t* mask,
                           const vo_sensor* sensors, int n_sensors, const vo_sensor* vox_sensors, const int* vox_sensor_off,
                           const vo_ctrl* ctrl, const double* terrain_x, const double* terrain_y, int n_terrain,
                           const double* genomes, int genome_len, int n, int n_threads, double* fitness_out, vo_result* results_out);

int vo_eval_batch(const vo_params* p, int w, int h, const uint8_t* mask,
                  const vo_sensor* sensors, int n_sensors, const vo_ctrl* ctrl,
                  const double* terrain_x, const double* terrain_y, int n_terrain,
                  const double* genomes, int genome_len, int n, int n_threads,
                  double* fitness_out, vo_result* results_out) {
  return eval_batch_impl(p, w, h, mask, sensors, n_sensors, NULL, NULL, ctrl, terrain_x, terrain_y, n_terrain, genomes, genome_len, n,
                         n_threads, fitness_out, results_out);
}

int vo_eval_batch_vox(const vo_params* p, int w, int h, const uint8_t* mask,
                      const vo_sensor* vox_sensors, const int* vox_sensor_off, const vo_ctrl* ctrl,
                      const double* terrain_x, const double* terrain_y, int n_terrain,
                      const double* genomes, int genome_len, int n, int n_threads,
                      double* fitness_out, vo_result* results_out) {
  return eval_batch_impl(p, w, h, mask, NULL, 0, vox_sensors, vox_sensor_off, ctrl, terrain_x, terrain_y, n_terrain, genomes, genome_len, n,
                         n_threads, fitness_out, results_out);
}

static int eval_batch_impl(const vo_params* p, int w, int h, const uint8_t* mask,
                           const vo_sensor* sensors, int n_sensors, const vo_sensor* vox_sensors, const int* vox_sensor_off,
                           const vo_ctrl* ctrl, const double* terrain_x, const double* terrain_y, int n_terrain,
                           const double* genomes, int genome_len, int n, int n_threads, double* fitness_out, vo_result* results_out) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  pthread_t th[256];
  BatchJob jobs[256];
  int next = 0;
  pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
  for (int i = 0; i < n_threads; i++) {
    BatchJob j = {p, w, h, mask, sensors, n_sensors, ctrl, terrain_x, terrain_y, n_terrain, genomes, genome_len, n,
                  fitness_out, results_out, &next, &mu, 0, vox_sensors, vox_sensor_off};
    jobs[i] = j;
    pthread_create(&th[i], NULL, batch_worker, &jobs[i]);
  }
  int err = 0;
  for (int i = 0; i < n_threads; i++) {
    pthread_join(th[i], NULL);
    if (jobs[i].err) err = jobs[i].err;
  }
  return err;
}

/* ------------------------------------------------------------------------------------------- */
/* known-answer hooks                                                                           */
/* ------------------------------------------------------------------------------------------- */
static void mini_world(World* W, const vo_params* p, int n_bodies, int n_joints, vo_result* tally) {
  memset(W, 0, sizeof(*W));
  W->p = *p;
  W->tally = tally;
  memset(tally, 0, sizeof(*tally));
  derive(W);
  W->n_bodies = n_bodies;
  W->bodies = (Body*)calloc(n_bodies, sizeof(Body));
  W->joints = (Joint*)calloc(n_joints > 0 ? n_joints : 1, sizeof(Joint));
  W->man = (Manifold*)calloc((size_t)n_bodies * MAXM, sizeof(Manifold));
  W->n_man = (int*)calloc(n_bodies, sizeof(int));
  for (int i = 0; i < n_bodies; i++) body_set_angle(&W->bodies[i], 0);
  /* flat ground far below unless the test installs one */
  W->n_t = 2;
  W->tx = (double*)malloc(sizeof(double) * 4);
  W->ty = (double*)malloc(sizeof(double) * 4);
  W->tx[0] = -1000; W->tx[1] = 1000; W->ty[0] = -1e6; W->ty[1] = -1e6;
  W->base_y = -1e6 - 50;
}

int vo_test_free_fall(const vo_params* p, int n_steps, double* y_out, double* vy_out) {
  World W; vo_result tally;
  mini_world(&W, p, 1, 0, &tally);
  for (int k = 0; k < n_steps; k++) {
    world_step(&W);
    y_out[k] = W.bodies[0].y;
    vy_out[k] = W.bodies[0].vy;
  }
  world_free(&W);
  return 0;
}

/* two corner masses joined centre-to-centre by one spring, no gravity/damping unless set in p */
int vo_test_spring_pair(const vo_params* p, double rest, double initial_len, int n_steps, double* len_out) {
  World W; vo_result tally;
  mini_world(&W, p, 2, 1, &tally);
  W.bodies[0].x = initial_len;
  Joint* j = add_joint(&W, J_SPRING, 0, 1, 0, 0, 0, 0);
  j->rest = rest;
  for (int k = 0; k < n_steps; k++) {
    world_step(&W);
    len_out[k] = W.bodies[0].x - W.bodies[1].x;
  }
  world_free(&W);
  return 0;
}

/* one corner mass dropped on flat ground y=0 */
int vo_test_box_drop(const vo_params* p, double angle0, double height, double vx0, int n_steps, double* state_out) {
  World W; vo_result tally;
  mini_world(&W, p, 1, 0, &tally);
  W.ty[0] = 0; W.ty[1] = 0; W.base_y = -50;
  W.bodies[0].y = height;
  W.bodies[0].vx = vx0;
  body_set_angle(&W.bodies[0], angle0);
  detect(&W);
  for (int k = 0; k < n_steps; k++) {
    world_step(&W);
    Body* b = &W.bodies[0];
    double* d = &state_out[k * 6];
    d[0] = b->x; d[1] = b->y; d[2] = b->a; d[3] = b->vx; d[4] = b->vy; d[5] = b->w;
  }
  world_free(&W);
  return 0;
}
