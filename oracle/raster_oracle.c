/*
 * oracle/raster_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference rasterizer (pytorch3d v0.7.9) used as the
 * parity checker for the CUDA kernels in pytorch3d_b200/csrc.  It is imported only
 * by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs.  The product path never calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks this file
 *   (a) bit-for-bit against the reference's own C++ CPU implementation compiled from
 *       /root/reference into oracle/_ref/ref_raster_cpu.so (arith = ARITH_CPU), and
 *   (b) against the committed golden fixtures in tests/golden/ that were produced by
 *       running the reference (tests/golden/make_golden.py),
 *   and tests/test_gpu_parity.py checks ARITH_CUDA bit-for-bit against the reference's
 *   CUDA kernels recompiled for sm_100a (oracle/_ref/ref_raster_cuda.so) on the B200.
 *
 * Two arithmetic flavours exist because the reference's CPU and CUDA builds round
 * differently (gcc -O2 never contracts; nvcc -fmad=true contracts a*b+c to FMA):
 *   ARITH_CPU  (0): every * and + individually rounded; association as in
 *                   pytorch3d/csrc/utils/geometry_utils.h
 *   ARITH_CUDA (1): FMA placement as nvcc 12.9 -arch=sm_100a compiles
 *                   pytorch3d/csrc/utils/geometry_utils.cuh (read from the SASS of the
 *                   reference build; see DESIGN.md "Arithmetic spec").
 *
 * Functions and the reference lines they restate:
 *   pix_to_ndc            rasterize_points/rasterization_utils.h:15-39 / .cuh:15-41
 *   edge_fn               utils/geometry_utils.h:52-55   / .cuh:37-40
 *   bary_coords           utils/geometry_utils.h:95-105  / .cuh:76-86
 *   bary_persp            utils/geometry_utils.h:193-206 / .cuh:172-185
 *   bary_clip             utils/geometry_utils.h:267-280 / .cuh:246-259
 *   point_line_dist       utils/geometry_utils.h:375-389 / .cuh:340-352
 *   point_tri_dist        utils/geometry_utils.h:499-512 / .cuh:397-408
 *   eval_face             rasterize_meshes/rasterize_meshes_cpu.cpp:173-238 / rasterize_meshes.cu:108-177
 *   mesh top-K (CPU form) rasterize_meshes_cpu.cpp:249-298  (sorted deque of tuples, pop largest)
 *   mesh top-K (CUDA form) rasterize_meshes.cu:179-237 + BubbleSort rasterization_utils.cuh:52-66
 *   meshes backward       rasterize_meshes_cpu.cpp:391-532 / rasterize_meshes.cu:433-564
 *   points forward        rasterize_points/rasterize_points_cpu.cpp:14-96 / rasterize_points.cu:38-81
 *   points backward       rasterize_points_cpu.cpp:196-251 / rasterize_points.cu:366-411
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared -o oracle/_build/libraster_oracle.so
 *            oracle/raster_oracle.c -lm -lpthread          (see oracle/__init__.py)
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ARITH_CPU 0
#define ARITH_CUDA 1
#define SELECT_CPU 0  /* lexicographic (z, idx, ...) top-K, rasterize_meshes_cpu.cpp:275-285 */
#define SELECT_CUDA 1 /* unsorted array + tracked max, rasterize_meshes.cu:216-236 */

#define K_EPS 1e-8 /* double, geometry_utils.h:17 / .cuh:18 */
#define MAX_K 150  /* kMaxPointsPerPixel, rasterization_utils.cuh:48 */

typedef struct {
  float z;
  int64_t idx;
  float dist;
  float b0, b1, b2;
} hit_t;

/* ------------------------------------------------------------------ scalars */

static inline float f_min3(float a, float b, float c) { return fminf(a, fminf(b, c)); }
static inline float f_max3(float a, float b, float c) { return fmaxf(a, fmaxf(b, c)); }

static float ndc_range(int S1, int S2) {
  float range = 2.0f;
  if (S1 > S2) range = ((float)S1 * range) / (float)S2;
  return range;
}

static float pix_to_ndc(int arith, int i, int S1, int S2) {
  const float range = ndc_range(S1, S2);
  const float offset = range / 2.0f;
  float t;
  if (arith == ARITH_CUDA)
    t = fmaf(range, (float)i, offset);
  else
    t = range * (float)i + offset;
  return t / (float)S1 - offset;
}

/* (p.x-a.x)*(b.y-a.y) - (p.y-a.y)*(b.x-a.x) */
static inline float edge_fn(int arith, float px, float py, float ax, float ay, float bx, float by) {
  if (arith == ARITH_CUDA) {
    const float t = (py - ay) * (bx - ax);
    return fmaf(px - ax, by - ay, -t);
  }
  return (px - ax) * (by - ay) - (py - ay) * (bx - ax);
}

/* a.x*b.x + a.y*b.y  (CUDA build: fma(a.x, b.x, rn(a.y*b.y))) */
static inline float dot2(int arith, float ax, float ay, float bx, float by) {
  if (arith == ARITH_CUDA) return fmaf(ax, bx, ay * by);
  return ax * bx + ay * by;
}

/* d.x*d.x + d.y*d.y for the final squared distances: the CUDA build fuses the OTHER product here,
 * fma(d.y, d.y, rn(d.x*d.x)) (SASS of RasterizeMeshesFine/NaiveCudaKernel and of the points kernels). */
static inline float sqnorm2(int arith, float dx, float dy) {
  if (arith == ARITH_CUDA) return fmaf(dy, dy, dx * dx);
  return dx * dx + dy * dy;
}

static void bary_coords(int arith, float px, float py, const float* v, float* w) {
  /* v = x0 y0 z0 x1 y1 z1 x2 y2 z2 */
  const float e = edge_fn(arith, v[6], v[7], v[0], v[1], v[3], v[4]);
  const float area = (float)((double)e + K_EPS);
  w[0] = edge_fn(arith, px, py, v[3], v[4], v[6], v[7]) / area;
  w[1] = edge_fn(arith, px, py, v[6], v[7], v[0], v[1]) / area;
  w[2] = edge_fn(arith, px, py, v[0], v[1], v[3], v[4]) / area;
}

static void bary_persp(int arith, const float* b, float z0, float z1, float z2, float* w) {
  float t0, t1, t2;
  if (arith == ARITH_CUDA) {
    t0 = b[0] * z1 * z2;
    t1 = z0 * b[1] * z2;
    t2 = z0 * z1 * b[2];
  } else {
    t0 = b[0] * z1 * z2;
    t1 = b[1] * z0 * z2;
    t2 = b[2] * z0 * z1;
  }
  const float denom = fmaxf(t0 + t1 + t2, (float)K_EPS);
  w[0] = t0 / denom;
  w[1] = t1 / denom;
  w[2] = t2 / denom;
}

static void bary_clip(const float* b, float* w) {
  float c0 = fmaxf(b[0], 0.0f), c1 = fmaxf(b[1], 0.0f), c2 = fmaxf(b[2], 0.0f);
  float s = c0 + c1 + c2;
  s = fmaxf(s, (float)1e-5);
  w[0] = c0 / s;
  w[1] = c1 / s;
  w[2] = c2 / s;
}

static float point_line_dist(int arith, float px, float py, float ax, float ay, float bx, float by) {
  const float bax = bx - ax, bay = by - ay;
  const float l2 = dot2(arith, bax, bay, bax, bay);
  if ((double)l2 <= K_EPS) {
    const float dx = px - bx, dy = py - by;
    return sqnorm2(arith, dx, dy);
  }
  float t = dot2(arith, bax, bay, px - ax, py - ay) / l2;
  t = fminf(fmaxf(t, 0.0f), 1.0f);
  float qx, qy;
  if (arith == ARITH_CUDA) {
    qx = fmaf(t, bax, ax);
    qy = fmaf(t, bay, ay);
  } else {
    qx = ax + t * bax;
    qy = ay + t * bay;
  }
  const float dx = qx - px, dy = qy - py;
  return sqnorm2(arith, dx, dy);
}

static float point_tri_dist(int arith, float px, float py, const float* v) {
  const float e01 = point_line_dist(arith, px, py, v[0], v[1], v[3], v[4]);
  const float e02 = point_line_dist(arith, px, py, v[0], v[1], v[6], v[7]);
  const float e12 = point_line_dist(arith, px, py, v[3], v[4], v[6], v[7]);
  return fminf(fminf(e01, e02), e12);
}

/* Per-(pixel, face) predicate + values.  Returns 1 on hit. */
static int eval_face(int arith, const float* v, float blur_radius, float sqrt_blur, float px, float py,
                     int persp, int clip, int cull, hit_t* out) {
  const float xmin = f_min3(v[0], v[3], v[6]) - sqrt_blur;
  const float xmax = f_max3(v[0], v[3], v[6]) + sqrt_blur;
  const float ymin = f_min3(v[1], v[4], v[7]) - sqrt_blur;
  const float ymax = f_max3(v[1], v[4], v[7]) + sqrt_blur;
  const float zmin = f_min3(v[2], v[5], v[8]);
  const float zmax = f_max3(v[2], v[5], v[8]);
  if (zmax < 0.0f) return 0;
  if ((double)zmin < K_EPS) return 0;
  if (px > xmax || px < xmin || py > ymax || py < ymin) return 0;
  const float area = edge_fn(arith, v[0], v[1], v[3], v[4], v[6], v[7]); /* EdgeFunctionForward(v0, v1, v2) */
  if (cull && area < 0.0f) return 0;
  if ((double)area <= K_EPS && (double)area >= -1.0f * K_EPS) return 0;

  float b0[3], b[3], bc[3];
  bary_coords(arith, px, py, v, b0);
  if (persp)
    bary_persp(arith, b0, v[2], v[5], v[8], b);
  else
    memcpy(b, b0, sizeof(b));
  if (clip)
    bary_clip(b, bc);
  else
    memcpy(bc, b, sizeof(bc));

  float pz;
  if (arith == ARITH_CUDA)
    pz = fmaf(bc[2], v[8], fmaf(bc[0], v[2], bc[1] * v[5]));
  else
    pz = bc[0] * v[2] + bc[1] * v[5] + bc[2] * v[8];
  if (pz < 0.0f) return 0;

  const float dist = point_tri_dist(arith, px, py, v);
  const int inside = b[0] > 0.0f && b[1] > 0.0f && b[2] > 0.0f;
  if (!inside && dist >= blur_radius) return 0;
  out->z = pz;
  out->dist = inside ? -dist : dist;
  out->b0 = bc[0];
  out->b1 = bc[1];
  out->b2 = bc[2];
  return 1;
}

/* lexicographic compare of the CPU tuple (z, idx, dist, b0, b1, b2): rasterize_meshes_cpu.cpp:281 */
static int hit_less(const hit_t* a, const hit_t* b) {
  if (a->z != b->z) return a->z < b->z;
  if (a->idx != b->idx) return a->idx < b->idx;
  if (a->dist != b->dist) return a->dist < b->dist;
  if (a->b0 != b->b0) return a->b0 < b->b0;
  if (a->b1 != b->b1) return a->b1 < b->b1;
  return a->b2 < b->b2;
}

static void sort_hits_cpu(hit_t* q, int n) { /* insertion sort == std::sort result for a strict total order */
  for (int i = 1; i < n; ++i) {
    hit_t t = q[i];
    int j = i - 1;
    while (j >= 0 && hit_less(&t, &q[j])) {
      q[j + 1] = q[j];
      --j;
    }
    q[j + 1] = t;
  }
}

/* ------------------------------------------------------------------ meshes forward */

typedef struct {
  const float* face_verts;
  const int64_t* first;
  const int64_t* num;
  const int64_t* neighbor;
  int N, H, W, K;
  float blur;
  int persp, clip, cull, arith, select;
  int y0, y1; /* output-row range */
  int64_t* pix_to_face;
  float* zbuf;
  float* bary;
  float* dists;
} mesh_fwd_args;

static void mesh_fwd_rows(const mesh_fwd_args* a) {
  const int H = a->H, W = a->W, K = a->K;
  const float sqrt_blur = sqrtf(a->blur);
  hit_t q[MAX_K + 1];
  for (int n = 0; n < a->N; ++n) {
    const int64_t f0 = a->first[n], f1 = f0 + a->num[n];
    for (int yi = a->y0; yi < a->y1; ++yi) {
      const float yf = pix_to_ndc(a->arith, H - 1 - yi, H, W);
      for (int xi = 0; xi < W; ++xi) {
        const float xf = pix_to_ndc(a->arith, W - 1 - xi, W, H);
        int qn = 0;
        float q_max_z = -1000.0f;
        int q_max_idx = -1;
        for (int64_t f = f0; f < f1; ++f) {
          hit_t h;
          if (!eval_face(a->arith, a->face_verts + 9 * f, a->blur, sqrt_blur, xf, yf, a->persp, a->clip,
                         a->cull, &h))
            continue;
          h.idx = f;
          const float dist = fabsf(h.dist);
          const int nb = a->neighbor ? (int)a->neighbor[f] : -1;
          int nb_k = -1;
          if (nb != -1)
            for (int i = 0; i < qn; ++i)
              if (q[i].idx == nb) {
                nb_k = i;
                break;
              }
          if (a->select == SELECT_CPU) {
            if (nb_k != -1) {
              if (dist < fabsf(q[nb_k].dist)) q[nb_k] = h;
            } else {
              q[qn++] = h;
            }
            sort_hits_cpu(q, qn);
            if (qn > K) qn = K;
          } else {
            if (nb_k != -1) {
              if (dist < fabsf(q[nb_k].dist)) {
                q[nb_k] = h;
                if (h.z > q_max_z) {
                  q_max_z = h.z;
                  q_max_idx = nb_k;
                }
              }
            } else if (qn < K) {
              q[qn] = h;
              if (h.z > q_max_z) {
                q_max_z = h.z;
                q_max_idx = qn;
              }
              qn++;
            } else if (h.z < q_max_z) {
              q[q_max_idx] = h;
              q_max_z = h.z;
              for (int i = 0; i < K; ++i)
                if (q[i].z > q_max_z) {
                  q_max_z = q[i].z;
                  q_max_idx = i;
                }
            }
          }
        }
        if (a->select == SELECT_CUDA) { /* BubbleSort with operator< on (z, idx) */
          for (int i = 0; i < qn - 1; ++i)
            for (int j = 0; j < qn - i - 1; ++j) {
              const hit_t *x = &q[j + 1], *y = &q[j];
              if (x->z < y->z || (x->z == y->z && x->idx < y->idx)) {
                hit_t t = q[j];
                q[j] = q[j + 1];
                q[j + 1] = t;
              }
            }
        }
        const int64_t base = (((int64_t)n * H + yi) * W + xi) * K;
        for (int k = 0; k < K; ++k) {
          if (k < qn) {
            a->pix_to_face[base + k] = q[k].idx;
            a->zbuf[base + k] = q[k].z;
            a->dists[base + k] = q[k].dist;
            a->bary[(base + k) * 3 + 0] = q[k].b0;
            a->bary[(base + k) * 3 + 1] = q[k].b1;
            a->bary[(base + k) * 3 + 2] = q[k].b2;
          } else {
            a->pix_to_face[base + k] = -1;
            a->zbuf[base + k] = -1.0f;
            a->dists[base + k] = -1.0f;
            a->bary[(base + k) * 3 + 0] = -1.0f;
            a->bary[(base + k) * 3 + 1] = -1.0f;
            a->bary[(base + k) * 3 + 2] = -1.0f;
          }
        }
      }
    }
  }
}

static void* mesh_fwd_thread(void* p) {
  mesh_fwd_rows((const mesh_fwd_args*)p);
  return NULL;
}

/* Rows [row_begin, row_end) of every image are computed (full image: 0, H); other rows are untouched. */
int oracle_rasterize_meshes_forward(const float* face_verts, const int64_t* first, const int64_t* num,
                                    const int64_t* neighbor, int N, int H, int W, float blur_radius, int K,
                                    int persp, int clip, int cull, int arith, int select, int row_begin,
                                    int row_end, int nthreads, int64_t* pix_to_face, float* zbuf, float* bary,
                                    float* dists) {
  if (K > MAX_K || K < 0) return 1;
  if (row_begin < 0) row_begin = 0;
  if (row_end > H) row_end = H;
  if (nthreads < 1) nthreads = 1;
  const int rows = row_end - row_begin;
  if (rows <= 0 || N == 0 || W == 0 || K == 0) return 0;
  if (nthreads > rows) nthreads = rows;
  mesh_fwd_args* args = (mesh_fwd_args*)malloc(sizeof(mesh_fwd_args) * nthreads);
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
  /* interleave small row blocks over threads so that load is balanced */
  const int chunk = (rows + nthreads - 1) / nthreads;
  int started = 0;
  for (int t = 0; t < nthreads; ++t) {
    const int y0 = row_begin + t * chunk;
    const int y1 = y0 + chunk < row_end ? y0 + chunk : row_end;
    if (y0 >= y1) break;
    mesh_fwd_args a = {face_verts, first, num, neighbor, N,   H,  W,  K,           blur_radius, persp,
                       clip,       cull,  arith, select, y0,  y1, pix_to_face, zbuf, bary, dists};
    args[t] = a;
    if (nthreads == 1)
      mesh_fwd_rows(&args[t]);
    else
      pthread_create(&th[t], NULL, mesh_fwd_thread, &args[t]);
    started++;
  }
  if (nthreads > 1)
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
  free(args);
  free(th);
  return 0;
}

/* ------------------------------------------------------------------ meshes backward */

static void edge_bwd(float px, float py, float ax, float ay, float bx, float by, float g, float* dp, float* da,
                     float* db) {
  /* EdgeFunctionBackward geometry_utils.h:70-80 */
  dp[0] = g * (by - ay);
  dp[1] = g * (ax - bx);
  da[0] = g * (py - by);
  da[1] = g * (bx - px);
  db[0] = g * (ay - py);
  db[1] = g * (px - ax);
}

static void bary_coords_bwd(int arith, float px, float py, const float* v, const float* gw, float* gv0,
                            float* gv1, float* gv2) {
  /* BarycentricCoordsBackward geometry_utils.h:121-181 */
  const float area = (float)((double)edge_fn(arith, v[6], v[7], v[0], v[1], v[3], v[4]) + K_EPS);
  const float area2 = area * area;
  const float e0 = edge_fn(arith, px, py, v[3], v[4], v[6], v[7]);
  const float e1 = edge_fn(arith, px, py, v[6], v[7], v[0], v[1]);
  const float e2 = edge_fn(arith, px, py, v[0], v[1], v[3], v[4]);
  float dp[2], da[2], db[2], ap[2], aa[2], ab[2];
  gv0[0] = gv0[1] = gv1[0] = gv1[1] = gv2[0] = gv2[1] = 0.0f;
  /* w0 = E(p, v1, v2) / area ; area = E(v2, v0, v1) */
  {
    const float g_area = gw[0] * (-e0 / area2), g_e = gw[0] * (1.0f / area);
    edge_bwd(px, py, v[3], v[4], v[6], v[7], g_e, dp, da, db);       /* (p, v1, v2) */
    edge_bwd(v[6], v[7], v[0], v[1], v[3], v[4], g_area, ap, aa, ab); /* (v2, v0, v1) */
    gv0[0] += aa[0]; gv0[1] += aa[1];
    gv1[0] += da[0] + ab[0]; gv1[1] += da[1] + ab[1];
    gv2[0] += db[0] + ap[0]; gv2[1] += db[1] + ap[1];
  }
  /* w1 = E(p, v2, v0) / area */
  {
    const float g_area = gw[1] * (-e1 / area2), g_e = gw[1] * (1.0f / area);
    edge_bwd(px, py, v[6], v[7], v[0], v[1], g_e, dp, da, db);       /* (p, v2, v0) */
    edge_bwd(v[6], v[7], v[0], v[1], v[3], v[4], g_area, ap, aa, ab);
    gv0[0] += db[0] + aa[0]; gv0[1] += db[1] + aa[1];
    gv1[0] += ab[0]; gv1[1] += ab[1];
    gv2[0] += da[0] + ap[0]; gv2[1] += da[1] + ap[1];
  }
  /* w2 = E(p, v0, v1) / area */
  {
    const float g_area = gw[2] * (-e2 / area2), g_e = gw[2] * (1.0f / area);
    edge_bwd(px, py, v[0], v[1], v[3], v[4], g_e, dp, da, db);       /* (p, v0, v1) */
    edge_bwd(v[6], v[7], v[0], v[1], v[3], v[4], g_area, ap, aa, ab);
    gv0[0] += da[0] + aa[0]; gv0[1] += da[1] + aa[1];
    gv1[0] += db[0] + ab[0]; gv1[1] += db[1] + ab[1];
    gv2[0] += ap[0]; gv2[1] += ap[1];
  }
}

static void bary_persp_bwd(const float* b, float z0, float z1, float z2, const float* go, float* gb, float* gz) {
  /* BarycentricPerspectiveCorrectionBackward geometry_utils.h:223-252 */
  const float t0 = b[0] * z1 * z2, t1 = b[1] * z0 * z2, t2 = b[2] * z0 * z1;
  const float denom = fmaxf(t0 + t1 + t2, (float)K_EPS);
  const float gdt = -t0 * go[0] - t1 * go[1] - t2 * go[2];
  const float gd = gdt / (denom * denom);
  const float g0 = gd + go[0] / denom, g1 = gd + go[1] / denom, g2 = gd + go[2] / denom;
  gb[0] = g0 * z1 * z2;
  gb[1] = g1 * z0 * z2;
  gb[2] = g2 * z0 * z1;
  gz[0] = g1 * b[1] * z2 + g2 * b[2] * z1;
  gz[1] = g0 * b[0] * z2 + g2 * b[2] * z0;
  gz[2] = g0 * b[0] * z1 + g1 * b[1] * z0;
}

static void bary_clip_bwd(const float* b, const float* gu, float* gb) {
  /* BarycentricClipBackward geometry_utils.h:294-350 */
  float w0 = fmaxf(b[0], 0.0f), w1 = fmaxf(b[1], 0.0f), w2 = fmaxf(b[2], 0.0f);
  float s = w0 + w1 + w2;
  float gsc = 1.0f;
  if (s < (float)1e-5) {
    gsc = 0.0f;
    s = (float)1e-5;
  }
  const float c0 = b[0] < 0.0f ? 0.0f : 1.0f, c1 = b[1] < 0.0f ? 0.0f : 1.0f, c2 = b[2] < 0.0f ? 0.0f : 1.0f;
  const float s2 = s * s;
  const float gs0 = -w0 / s2 * gsc, gs1 = -w1 / s2 * gsc, gs2 = -w2 / s2 * gsc;
  gb[0] = c0 * (gu[0] * (1.0f / s + gs0) + gu[1] * gs1 + gu[2] * gs2);
  gb[1] = c1 * (gu[1] * (1.0f / s + gs1) + gu[0] * gs0 + gu[2] * gs2);
  gb[2] = c2 * (gu[2] * (1.0f / s + gs2) + gu[0] * gs0 + gu[1] * gs1);
}

static void point_line_bwd(float px, float py, float ax, float ay, float bx, float by, float g, float* ga,
                           float* gb) {
  /* PointLineDistanceBackward geometry_utils.h:420-440 */
  const float bax = bx - ax, bay = by - ay, pax = px - ax, pay = py - ay;
  const float tb = bax * bax + bay * bay, tt0 = bax * pax + bay * pay;
  float t = tt0 / tb;
  t = fminf(fmaxf(t, 0.0f), 1.0f);
  const float qx = (1.0f - t) * ax + t * bx, qy = (1.0f - t) * ay + t * by;
  ga[0] = g * (1.0f - t) * 2.0f * (qx - px);
  ga[1] = g * (1.0f - t) * 2.0f * (qy - py);
  gb[0] = g * t * 2.0f * (qx - px);
  gb[1] = g * t * 2.0f * (qy - py);
}

/* arith also selects which bary feeds BarycentricClipBackward: the CPU passes the perspective-corrected
 * bary (rasterize_meshes_cpu.cpp:498-500), the CUDA kernel the uncorrected one (rasterize_meshes.cu:527-529). */
int oracle_rasterize_meshes_backward(const float* face_verts, int64_t F, const int64_t* pix_to_face,
                                     const float* grad_zbuf, const float* grad_bary, const float* grad_dists,
                                     int N, int H, int W, int K, int persp, int clip, int arith,
                                     int clip_bwd_uncorrected, int row_begin, int row_end,
                                     float* grad_face_verts) {
  memset(grad_face_verts, 0, sizeof(float) * 9 * (size_t)F);
  if (row_begin < 0) row_begin = 0;
  if (row_end > H) row_end = H;
  for (int n = 0; n < N; ++n)
    for (int y = row_begin; y < row_end; ++y) {
      const float yf = pix_to_ndc(arith, H - 1 - y, H, W);
      for (int x = 0; x < W; ++x) {
        const float xf = pix_to_ndc(arith, W - 1 - x, W, H);
        for (int k = 0; k < K; ++k) {
          const int64_t i = (((int64_t)n * H + y) * W + x) * K + k;
          const int64_t f = pix_to_face[i];
          if (f < 0) continue;
          const float* v = face_verts + 9 * f;
          float* g = grad_face_verts + 9 * f;
          const float gd = grad_dists[i], gz = grad_zbuf[i];
          float b0[3], b[3], bc[3];
          bary_coords(arith, xf, yf, v, b0);
          if (persp)
            bary_persp(arith, b0, v[2], v[5], v[8], b);
          else
            memcpy(b, b0, sizeof(b));
          if (clip)
            bary_clip(b, bc);
          else
            memcpy(bc, b, sizeof(bc));
          const int inside = b[0] > 0.0f && b[1] > 0.0f && b[2] > 0.0f;
          const float sgd = (inside ? -1.0f : 1.0f) * gd;

          /* PointTriangleDistanceBackward geometry_utils.h:531-571 */
          float dv0[2] = {0, 0}, dv1[2] = {0, 0}, dv2[2] = {0, 0};
          const float e01 = point_line_dist(arith, xf, yf, v[0], v[1], v[3], v[4]);
          const float e02 = point_line_dist(arith, xf, yf, v[0], v[1], v[6], v[7]);
          const float e12 = point_line_dist(arith, xf, yf, v[3], v[4], v[6], v[7]);
          if (e01 <= e02 && e01 <= e12)
            point_line_bwd(xf, yf, v[0], v[1], v[3], v[4], sgd, dv0, dv1);
          else if (e02 <= e01 && e02 <= e12)
            point_line_bwd(xf, yf, v[0], v[1], v[6], v[7], sgd, dv0, dv2);
          else if (e12 <= e01 && e12 <= e02)
            point_line_bwd(xf, yf, v[3], v[4], v[6], v[7], sgd, dv1, dv2);

          float gsum[3] = {grad_bary[i * 3 + 0] + gz * v[2], grad_bary[i * 3 + 1] + gz * v[5],
                           grad_bary[i * 3 + 2] + gz * v[8]};
          float gb[3] = {gsum[0], gsum[1], gsum[2]};
          if (clip) bary_clip_bwd(clip_bwd_uncorrected ? b0 : b, gsum, gb);
          float gzp[3] = {0, 0, 0};
          if (persp) {
            float gb2[3];
            bary_persp_bwd(b0, v[2], v[5], v[8], gb, gb2, gzp);
            memcpy(gb, gb2, sizeof(gb));
            /* the CPU adds the perspective z-gradients first (rasterize_meshes_cpu.cpp:506-508) */
            g[2] += gzp[0];
            g[5] += gzp[1];
            g[8] += gzp[2];
          }
          float bv0[2], bv1[2], bv2[2];
          bary_coords_bwd(arith, xf, yf, v, gb, bv0, bv1, bv2);
          g[0] += bv0[0] + dv0[0];
          g[1] += bv0[1] + dv0[1];
          g[2] += gz * bc[0];
          g[3] += bv1[0] + dv1[0];
          g[4] += bv1[1] + dv1[1];
          g[5] += gz * bc[1];
          g[6] += bv2[0] + dv2[0];
          g[7] += bv2[1] + dv2[1];
          g[8] += gz * bc[2];
        }
      }
    }
  return 0;
}

/* ------------------------------------------------------------------ points */

typedef struct {
  float z;
  int32_t idx;
  float d2;
} phit_t;

static int phit_less(const phit_t* a, const phit_t* b) {
  if (a->z != b->z) return a->z < b->z;
  if (a->idx != b->idx) return a->idx < b->idx;
  return a->d2 < b->d2;
}

typedef struct {
  const float* points;
  const int64_t* first;
  const int64_t* num;
  const float* radius;
  int N, H, W, K, arith, select, y0, y1;
  int32_t* idx;
  float* zbuf;
  float* dists;
} pts_fwd_args;

static void pts_fwd_rows(const pts_fwd_args* a) {
  const int H = a->H, W = a->W, K = a->K;
  phit_t q[MAX_K + 1];
  for (int n = 0; n < a->N; ++n) {
    const int64_t p0 = a->first[n], p1 = p0 + a->num[n];
    for (int yi = a->y0; yi < a->y1; ++yi) {
      const float yf = pix_to_ndc(a->arith, H - 1 - yi, H, W);
      for (int xi = 0; xi < W; ++xi) {
        const float xf = pix_to_ndc(a->arith, W - 1 - xi, W, H);
        int qn = 0;
        float q_max_z = -1000.0f;
        int q_max_idx = -1;
        for (int64_t p = p0; p < p1; ++p) {
          const float px = a->points[3 * p], py = a->points[3 * p + 1], pz = a->points[3 * p + 2];
          const float r = a->radius[p];
          const float r2 = r * r;
          if (pz < 0.0f) continue;
          const float dx = xf - px, dy = yf - py;
          const float d2 = sqnorm2(a->arith, dx, dy);
          if (!(d2 < r2)) continue;
          phit_t h = {pz, (int32_t)p, d2};
          if (a->select == SELECT_CPU) {
            /* std::priority_queue of tuples, pop the largest: rasterize_points_cpu.cpp:56-76 */
            int j = qn++;
            while (j > 0 && phit_less(&h, &q[j - 1])) {
              q[j] = q[j - 1];
              --j;
            }
            q[j] = h;
            if (qn > K) qn = K;
          } else {
            if (qn < K) {
              q[qn] = h;
              if (pz > q_max_z) {
                q_max_z = pz;
                q_max_idx = qn;
              }
              qn++;
            } else if (pz < q_max_z) {
              q[q_max_idx] = h;
              q_max_z = pz;
              for (int i = 0; i < K; ++i)
                if (q[i].z > q_max_z) {
                  q_max_z = q[i].z;
                  q_max_idx = i;
                }
            }
          }
        }
        if (a->select == SELECT_CUDA) { /* BubbleSort on z only (stable) rasterize_points.cu:26-28 */
          for (int i = 0; i < qn - 1; ++i)
            for (int j = 0; j < qn - i - 1; ++j)
              if (q[j + 1].z < q[j].z) {
                phit_t t = q[j];
                q[j] = q[j + 1];
                q[j + 1] = t;
              }
        }
        const int64_t base = (((int64_t)n * H + yi) * W + xi) * K;
        for (int k = 0; k < K; ++k) {
          a->idx[base + k] = k < qn ? q[k].idx : -1;
          a->zbuf[base + k] = k < qn ? q[k].z : -1.0f;
          a->dists[base + k] = k < qn ? q[k].d2 : -1.0f;
        }
      }
    }
  }
}

static void* pts_fwd_thread(void* p) {
  pts_fwd_rows((const pts_fwd_args*)p);
  return NULL;
}

int oracle_rasterize_points_forward(const float* points, const int64_t* first, const int64_t* num,
                                    const float* radius, int N, int H, int W, int K, int arith, int select,
                                    int row_begin, int row_end, int nthreads, int32_t* idx, float* zbuf,
                                    float* dists) {
  if (K > MAX_K || K < 0) return 1;
  if (row_begin < 0) row_begin = 0;
  if (row_end > H) row_end = H;
  if (nthreads < 1) nthreads = 1;
  const int rows = row_end - row_begin;
  if (rows <= 0 || N == 0 || W == 0 || K == 0) return 0;
  if (nthreads > rows) nthreads = rows;
  pts_fwd_args* args = (pts_fwd_args*)malloc(sizeof(pts_fwd_args) * nthreads);
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
  const int chunk = (rows + nthreads - 1) / nthreads;
  int started = 0;
  for (int t = 0; t < nthreads; ++t) {
    const int y0 = row_begin + t * chunk;
    const int y1 = y0 + chunk < row_end ? y0 + chunk : row_end;
    if (y0 >= y1) break;
    pts_fwd_args a = {points, first, num, radius, N, H, W, K, arith, select, y0, y1, idx, zbuf, dists};
    args[t] = a;
    if (nthreads == 1)
      pts_fwd_rows(&args[t]);
    else
      pthread_create(&th[t], NULL, pts_fwd_thread, &args[t]);
    started++;
  }
  if (nthreads > 1)
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
  free(args);
  free(th);
  return 0;
}

int oracle_rasterize_points_backward(const float* points, int64_t P, const int32_t* idx, const float* grad_zbuf,
                                     const float* grad_dists, int N, int H, int W, int K, int arith,
                                     float* grad_points) {
  memset(grad_points, 0, sizeof(float) * 3 * (size_t)P);
  for (int n = 0; n < N; ++n)
    for (int y = 0; y < H; ++y) {
      const float yf = pix_to_ndc(arith, H - 1 - y, H, W);
      for (int x = 0; x < W; ++x) {
        const float xf = pix_to_ndc(arith, W - 1 - x, W, H);
        for (int k = 0; k < K; ++k) {
          const int64_t i = (((int64_t)n * H + y) * W + x) * K + k;
          const int32_t p = idx[i];
          if (p < 0) continue;
          const float gd = grad_dists[i];
          const float dx = points[3 * p] - xf, dy = points[3 * p + 1] - yf;
          grad_points[3 * p + 0] += 2.0f * gd * dx;
          grad_points[3 * p + 1] += 2.0f * gd * dy;
          grad_points[3 * p + 2] += grad_zbuf[i];
        }
      }
    }
  return 0;
}

/* ------------------------------------------------------------------ alpha compositing */
/* Restates pytorch3d/csrc/compositing/alpha_composite_cpu.cpp:17-60 (forward) and :62-130 (backward), and the
 * CUDA kernels' product order (alpha_composite.cu:63-64: features * cum_alpha * alpha; CPU: cum_alpha * alpha *
 * features).  Layouts: features (C,P), alphas / points_idx (N,K,H,W) contiguous, result (N,C,H,W). */
int oracle_alpha_composite_forward(const float* features, int64_t C, int64_t P, const float* alphas,
                                   const int64_t* points_idx, int N, int K, int H, int W, int arith, float* result) {
  const int64_t plane = (int64_t)H * W;
  for (int b = 0; b < N; ++b)
    for (int64_t c = 0; c < C; ++c)
      for (int64_t px = 0; px < plane; ++px) {
        float cum = 1.0f, acc = 0.0f;
        for (int k = 0; k < K; ++k) {
          const int64_t id = points_idx[((int64_t)b * K + k) * plane + px];
          if (id < 0) continue;
          const float a = alphas[((int64_t)b * K + k) * plane + px];
          const float f = features[c * P + id];
          const float term = (arith == ARITH_CUDA) ? (f * cum) * a : (cum * a) * f;
          acc += term;
          cum = cum * (1 - a);
        }
        result[((int64_t)b * C + c) * plane + px] = acc;
      }
  return 0;
}

int oracle_alpha_composite_backward(const float* grad_out, const float* features, int64_t C, int64_t P,
                                    const float* alphas, const int64_t* points_idx, int N, int K, int H, int W,
                                    float* grad_features, float* grad_alphas) {
  const int64_t plane = (int64_t)H * W;
  const float eps = 1e-9f;
  memset(grad_features, 0, sizeof(float) * (size_t)(C * P));
  memset(grad_alphas, 0, sizeof(float) * (size_t)N * K * plane);
  for (int b = 0; b < N; ++b)
    for (int64_t c = 0; c < C; ++c)
      for (int64_t px = 0; px < plane; ++px) {
        const float g = grad_out[((int64_t)b * C + c) * plane + px];
        float cum = 1.0f;
        for (int k = 0; k < K; ++k) {
          const int64_t id = points_idx[((int64_t)b * K + k) * plane + px];
          if (id < 0) continue;
          const float a = alphas[((int64_t)b * K + k) * plane + px];
          grad_alphas[((int64_t)b * K + k) * plane + px] += g * features[c * P + id] * cum;
          grad_features[c * P + id] += g * cum * a;
          for (int t = 0; t < k; ++t) {
            if (points_idx[((int64_t)b * K + t) * plane + px] < 0) continue;
            const float at = alphas[((int64_t)b * K + t) * plane + px];
            grad_alphas[((int64_t)b * K + t) * plane + px] -= g * features[c * P + id] * cum * a / (1 - at + eps);
          }
          cum = cum * (1 - a);
        }
      }
  return 0;
}

/* ------------------------------------------------------------------ weighted sums */
/* Restates pytorch3d/csrc/compositing/weighted_sum_cpu.cpp:17-55 / :57-103 and norm_weighted_sum_cpu.cpp:19-68 /
 * :70-137 (the CUDA kernels weighted_sum.cu:22-103, norm_weighted_sum.cu:24-160 perform the same operations in the
 * same order: (alpha * f) [/ total], summed over ascending k -- one arithmetic flavour).  norm != 0: divide by
 * max(sum of the valid alphas, 1e-4). */
int oracle_weighted_sum_forward(const float* features, int64_t C, int64_t P, const float* alphas,
                                const int64_t* points_idx, int N, int K, int H, int W, int norm, float* result) {
  const int64_t plane = (int64_t)H * W;
  for (int b = 0; b < N; ++b)
    for (int64_t c = 0; c < C; ++c)
      for (int64_t px = 0; px < plane; ++px) {
        float t_alpha = 0.0f;
        for (int k = 0; k < K; ++k) {
          if (points_idx[((int64_t)b * K + k) * plane + px] < 0) continue;
          t_alpha += alphas[((int64_t)b * K + k) * plane + px];
        }
        if (t_alpha < 1e-4f) t_alpha = 1e-4f;
        float acc = 0.0f;
        for (int k = 0; k < K; ++k) {
          const int64_t id = points_idx[((int64_t)b * K + k) * plane + px];
          if (id < 0) continue;
          const float a = alphas[((int64_t)b * K + k) * plane + px];
          const float t = a * features[c * P + id];
          acc += norm ? t / t_alpha : t;
        }
        result[((int64_t)b * C + c) * plane + px] = acc;
      }
  return 0;
}

int oracle_weighted_sum_backward(const float* grad_out, const float* features, int64_t C, int64_t P,
                                 const float* alphas, const int64_t* points_idx, int N, int K, int H, int W, int norm,
                                 float* grad_features, float* grad_alphas) {
  const int64_t plane = (int64_t)H * W;
  memset(grad_features, 0, sizeof(float) * (size_t)(C * P));
  memset(grad_alphas, 0, sizeof(float) * (size_t)N * K * plane);
  for (int b = 0; b < N; ++b)
    for (int64_t c = 0; c < C; ++c)
      for (int64_t px = 0; px < plane; ++px) {
        const float g = grad_out[((int64_t)b * C + c) * plane + px];
        float t_alpha = 0.0f, t_alphafs = 0.0f;
        for (int k = 0; k < K; ++k) {
          const int64_t id = points_idx[((int64_t)b * K + k) * plane + px];
          if (id < 0) continue;
          const float a = alphas[((int64_t)b * K + k) * plane + px];
          t_alpha += a;
          t_alphafs += a * features[c * P + id];
        }
        if (t_alpha < 1e-4f) t_alpha = 1e-4f;
        for (int k = 0; k < K; ++k) {
          const int64_t id = points_idx[((int64_t)b * K + k) * plane + px];
          if (id < 0) continue;
          const float a = alphas[((int64_t)b * K + k) * plane + px];
          if (norm) {
            grad_alphas[((int64_t)b * K + k) * plane + px] +=
                g * (features[c * P + id] * t_alpha - t_alphafs) / (t_alpha * t_alpha);
            grad_features[c * P + id] += g * a / t_alpha;
          } else {
            grad_alphas[((int64_t)b * K + k) * plane + px] += g * features[c * P + id];
            grad_features[c * P + id] += g * a;
          }
        }
      }
  return 0;
}

/* ------------------------------------------------------------------ face attribute interpolation */
/* Restates InterpFaceAttrsForwardKernel / BackwardKernel (pytorch3d/csrc/interp_face_attrs/interp_face_attrs.cu:15-49,
 * 86-124) and the CPU path interpolate_face_attributes_python (pytorch3d/ops/interp_face_attrs.py:83-102).
 * ARITH_CUDA: fma(w2,a2, fma(w1,a1, fma(w0,a0,0))) as compiled; ARITH_CPU: (w0*a0 + w1*a1) + w2*a2. */
int oracle_interp_face_attrs_forward(const int64_t* pix_to_face, const float* bary, const float* attrs, int64_t P,
                                     int64_t D, int arith, float* out) {
  for (int64_t p = 0; p < P; ++p) {
    const int64_t f = pix_to_face[p];
    for (int64_t d = 0; d < D; ++d) {
      float v = 0.0f;
      if (f >= 0) {
        const float* a = attrs + f * 3 * D + d;
        const float w0 = bary[p * 3], w1 = bary[p * 3 + 1], w2 = bary[p * 3 + 2];
        if (arith == ARITH_CUDA)
          v = fmaf(w2, a[2 * D], fmaf(w1, a[D], fmaf(w0, a[0], 0.0f)));
        else
          v = (w0 * a[0] + w1 * a[D]) + w2 * a[2 * D];
      }
      out[p * D + d] = v;
    }
  }
  return 0;
}

int oracle_interp_face_attrs_backward(const int64_t* pix_to_face, const float* bary, const float* attrs,
                                      const float* grad_out, int64_t P, int64_t F, int64_t D, float* grad_bary,
                                      float* grad_attrs) {
  memset(grad_bary, 0, sizeof(float) * 3 * (size_t)P);
  memset(grad_attrs, 0, sizeof(float) * 3 * (size_t)(F * D));
  for (int64_t p = 0; p < P; ++p) {
    const int64_t f = pix_to_face[p];
    if (f < 0) continue;
    for (int64_t d = 0; d < D; ++d) {
      const float u = grad_out[p * D + d];
      for (int i = 0; i < 3; ++i) {
        grad_bary[p * 3 + i] += attrs[f * 3 * D + i * D + d] * u;
        grad_attrs[f * 3 * D + i * D + d] += bary[p * 3 + i] * u;
      }
    }
  }
  return 0;
}
