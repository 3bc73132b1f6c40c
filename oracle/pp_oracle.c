/*
 * pp_oracle.c -- CPU restatement of the reference's native hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (3d-object-detection_b200/) never links, imports or falls back
 * to it.
 *
 * What it restates (all paths under /root/reference):
 *   pp_oracle_create_pillars  <- data/pillars.cpp:236-398 (create_pillars), :18-59 (feature order),
 *                                :311-328 (sequential running mean)
 *   pp_oracle_iou             <- data/pillars.cpp:132-172 (iou) with ring conventions :15-16
 *   pp_oracle_make_ious       <- data/pillars.cpp:400-427 (make_ious)
 *
 * Third-party arithmetic that is NOT under /root/reference and NOT installed here:
 *   Boost.Geometry intersection/area (unpinned version, install_mods.sh:5) -> restated as a
 *     double-precision Sutherland-Hodgman convex clip (exact for convex quads up to rounding;
 *     checked against exact rational arithmetic in tests/test_oracle_iou.py).
 *   boost::unordered_map iteration order (pillar slot order, data/pillars.cpp:332-335) ->
 *     restated as INSERTION order (first in-range point of each pillar, ascending).
 * PARITY PINNING: this file is checked bit-for-bit against the reference's own pillars.cpp
 * compiled with a Boost stand-in (oracle/Makefile target `ref`, oracle/boost_shim/) and against
 * golden fixtures that build produced (tests/golden/).  Boost's own hash order and last-bit IoU
 * behaviour remain UNPINNED (no Boost in the image, no tests or vectors in the reference).
 *
 * Build: gcc -O3 -std=c11 -fPIC -shared -ffp-contract=off pp_oracle.c -o libpp_oracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PP_ORACLE_OK 0
#define PP_ORACLE_ENOMEM 1
#define PP_ORACLE_ENEGIOU 2 /* reference would print "IOU < 0" and exit(1): pillars.cpp:166-169 */

/* ------------------------------------------------------------------------------------------
 * create_pillars
 * ---------------------------------------------------------------------------------------- */

typedef struct {
  double cx, cy;      /* canvas cell (integers held as doubles, like the reference's map key) */
  double mean[4];     /* running mean x,y,z and count: pillars.cpp:313-327 */
  int64_t head, tail; /* linked list of member points, in input order */
  int64_t count;
} oracle_pillar;

static uint64_t key_hash(double cx, double cy) {
  uint64_t a, b;
  memcpy(&a, &cx, 8);
  memcpy(&b, &cy, 8);
  uint64_t h = a * 0x9E3779B97F4A7C15ull ^ (b + 0x7F4A7C15ull) * 0xC2B2AE3D27D4EB4Full;
  h ^= h >> 29;
  h *= 0xBF58476D1CE4E5B9ull;
  h ^= h >> 32;
  return h;
}

/*
 * points: element (i, c) at points[i*stride_pt + c*stride_col]  (the reference receives a
 * transposed view of a [4,N] array, data/dataset.py:88, so both strides are arbitrary).
 * tensor:  caller-zeroed [max_pillars, max_points, 9] doubles, only touched slots written.
 * indices: caller-zeroed [max_pillars, 3] doubles, rows of kept pillars set to [1, cx, cy].
 * Returns the number of pillars written through *n_pillars_out (may be NULL).
 */
int pp_oracle_create_pillars(const double* points, int64_t npts, int64_t stride_pt,
                             int64_t stride_col, double* tensor, double* indices,
                             int max_points_per_pillar, int max_pillars, double x_step,
                             double y_step, double x_min, double y_min, double z_min,
                             double x_max, double y_max, double z_max, double canvas_height,
                             int64_t* n_pillars_out) {
  int64_t cap = 16;
  while (cap < 2 * npts + 16) cap <<= 1;
  int64_t* table = (int64_t*)malloc((size_t)cap * sizeof(int64_t));
  oracle_pillar* pil = (oracle_pillar*)malloc((size_t)(npts > 0 ? npts : 1) * sizeof(oracle_pillar));
  int64_t* next = (int64_t*)malloc((size_t)(npts > 0 ? npts : 1) * sizeof(int64_t));
  if (!table || !pil || !next) {
    free(table); free(pil); free(next);
    return PP_ORACLE_ENOMEM;
  }
  for (int64_t i = 0; i < cap; ++i) table[i] = -1;
  int64_t npil = 0;

  for (int64_t i = 0; i < npts; ++i) {
    const double x = points[i * stride_pt + 0 * stride_col];
    const double y = points[i * stride_pt + 1 * stride_col];
    const double z = points[i * stride_pt + 2 * stride_col];
    /* pillars.cpp:271-275 -- half-open range filter; NaN passes it exactly as in the reference */
    if ((x >= x_max) || (x < x_min) || (y >= y_max) || (y < y_min) || (z >= z_max) || (z < z_min))
      continue;
    /* pillars.cpp:278-280 */
    double canvas_x = floor((x - x_min) / x_step);
    double canvas_y = floor((y - y_min) / y_step);
    canvas_y = (canvas_height - 1) - canvas_y;

    uint64_t h = key_hash(canvas_x, canvas_y) & (uint64_t)(cap - 1);
    int64_t p = -1;
    for (;;) {
      int64_t e = table[h];
      if (e < 0) break;
      if (pil[e].cx == canvas_x && pil[e].cy == canvas_y) { p = e; break; }
      h = (h + 1) & (uint64_t)(cap - 1);
    }
    next[i] = -1;
    if (p < 0) {
      /* pillars.cpp:292-297 and :311-318 -- new pillar, mean initialised with this point */
      p = npil++;
      table[h] = p;
      pil[p].cx = canvas_x;
      pil[p].cy = canvas_y;
      pil[p].head = pil[p].tail = i;
      pil[p].count = 1;
      pil[p].mean[0] = x;
      pil[p].mean[1] = y;
      pil[p].mean[2] = z;
      pil[p].mean[3] = 1;
    } else {
      /* pillars.cpp:299-303 and :320-328 -- append, sequential running mean */
      next[pil[p].tail] = i;
      pil[p].tail = i;
      pil[p].count += 1;
      double n = pil[p].mean[3];
      pil[p].mean[0] = pil[p].mean[0] * (n / (n + 1)) + x / (n + 1);
      pil[p].mean[1] = pil[p].mean[1] * (n / (n + 1)) + y / (n + 1);
      pil[p].mean[2] = pil[p].mean[2] * (n / (n + 1)) + z / (n + 1);
      pil[p].mean[3] = n + 1;
    }
  }

  /* pillars.cpp:335-396 -- pillars in canonical (insertion) order, first max_pillars kept,
     first max_points_per_pillar points of each kept */
  int64_t kept = 0;
  for (int64_t p = 0; p < npil; ++p) {
    if (kept >= max_pillars) break;
    int num_points = 0;
    for (int64_t i = pil[p].head; i >= 0; i = next[i]) {
      if (num_points >= max_points_per_pillar) break;
      const double x = points[i * stride_pt + 0 * stride_col];
      const double y = points[i * stride_pt + 1 * stride_col];
      const double z = points[i * stride_pt + 2 * stride_col];
      const double r = points[i * stride_pt + 3 * stride_col];
      double* f = tensor + ((size_t)kept * (size_t)max_points_per_pillar + (size_t)num_points) * 9;
      f[0] = x;                   /* pillars.cpp:48-56 feature order */
      f[1] = y;
      f[2] = z;
      f[3] = r;
      f[4] = pil[p].cx - x;       /* xp = canvas_x - x : pillars.cpp:30 */
      f[5] = pil[p].cy - y;       /* yp = canvas_y - y : pillars.cpp:31 */
      f[6] = pil[p].mean[0] - x;  /* pillars.cpp:381-383 */
      f[7] = pil[p].mean[1] - y;
      f[8] = pil[p].mean[2] - z;
      num_points++;
    }
    indices[kept * 3 + 0] = 1; /* pillars.cpp:390-392 */
    indices[kept * 3 + 1] = pil[p].cx;
    indices[kept * 3 + 2] = pil[p].cy;
    kept++;
  }
  if (n_pillars_out) *n_pillars_out = kept;
  free(table); free(pil); free(next);
  return PP_ORACLE_OK;
}

/* ------------------------------------------------------------------------------------------
 * iou / make_ious
 * ---------------------------------------------------------------------------------------- */

static double ring_area_ccw(const double* v, int n) {
  /* shoelace, counter-clockwise positive, open ring of n (x,y) pairs */
  if (n < 3) return 0.0;
  double s = 0.0;
  for (int i = 0; i < n; ++i) {
    int j = (i + 1 == n) ? 0 : i + 1;
    s += v[2 * i] * v[2 * j + 1] - v[2 * j] * v[2 * i + 1];
  }
  return 0.5 * s;
}

/*
 * a: anchor ring, 4 corners, counter-clockwise, open  (Polygon_cc, pillars.cpp:16,149-152)
 * g: GT ring, 4 corners, clockwise, open              (Polygon,    pillars.cpp:15,154-157)
 * Returns intersection / union in double; 0 when the intersection has no area (Boost returns
 * no output polygon, pillars.cpp:161-163).  A negative value means wrong corner winding (the
 * reference exits the process, pillars.cpp:166-169).
 */
double pp_oracle_iou(const double* a, const double* g) {
  double buf0[16 * 2], buf1[16 * 2];
  double* subj = buf0;
  double* nxt = buf1;
  int n = 4;
  memcpy(subj, a, 8 * sizeof(double));
  /* clip polygon = GT ring walked backwards so that it is counter-clockwise */
  for (int e = 0; e < 4 && n > 0; ++e) {
    const double* c0 = g + 2 * (3 - e);
    const double* c1 = g + 2 * ((3 - e + 3) % 4); /* previous corner in storage = next in CCW walk */
    const double ex = c1[0] - c0[0], ey = c1[1] - c0[1];
    int m = 0;
    for (int i = 0; i < n; ++i) {
      const double* cur = subj + 2 * i;
      const double* prv = subj + 2 * ((i + n - 1) % n);
      const double dc = ex * (cur[1] - c0[1]) - ey * (cur[0] - c0[0]); /* >= 0: inside (left) */
      const double dp = ex * (prv[1] - c0[1]) - ey * (prv[0] - c0[0]);
      const int cin = dc >= 0.0, pin = dp >= 0.0;
      if (cin != pin) {
        const double t = dp / (dp - dc);
        nxt[2 * m] = prv[0] + t * (cur[0] - prv[0]);
        nxt[2 * m + 1] = prv[1] + t * (cur[1] - prv[1]);
        ++m;
      }
      if (cin) {
        nxt[2 * m] = cur[0];
        nxt[2 * m + 1] = cur[1];
        ++m;
      }
    }
    double* t2 = subj; subj = nxt; nxt = t2;
    n = m;
  }
  if (n < 3) return 0.0;
  /* The reference's output vector holds clockwise-typed polygons (pillars.cpp:15,159), so the
     intersection ring is presented clockwise and bg::area (pillars.cpp:164) walks it that way:
     reverse the ring, take the shoelace sum, flip the sign. */
  for (int i = 0; i < n; ++i) {
    nxt[2 * i] = subj[2 * (n - 1 - i)];
    nxt[2 * i + 1] = subj[2 * (n - 1 - i) + 1];
  }
  const double inter = -ring_area_ccw(nxt, n);
  if (!(inter > 0.0)) return 0.0;
  const double area_a = ring_area_ccw(a, 4);  /* bg::area of a CCW-typed polygon */
  const double area_g = -ring_area_ccw(g, 4); /* bg::area of a CW-typed polygon */
  return inter / (area_a + area_g - inter);
}

/*
 * a_corners [A,4,2], g_corners [G,4,2], a_centers [A,3], g_centers [G,3], ious [A,G], all
 * contiguous doubles.  Every entry of ious is written (pillars.cpp:416-426).
 */
int pp_oracle_make_ious(const double* a_corners, const double* g_corners, const double* a_centers,
                        const double* g_centers, double* ious, int64_t A, int64_t G) {
  for (int64_t i = 0; i < A; ++i) {
    for (int64_t j = 0; j < G; ++j) {
      if ((fabs(a_centers[i * 3 + 0] - g_centers[j * 3 + 0]) > 10) ||
          (fabs(a_centers[i * 3 + 1] - g_centers[j * 3 + 1]) > 10)) {
        ious[i * G + j] = 0;
        continue;
      }
      const double v = pp_oracle_iou(a_corners + i * 8, g_corners + j * 8);
      if (v < 0) return PP_ORACLE_ENEGIOU;
      ious[i * G + j] = v;
    }
  }
  return PP_ORACLE_OK;
}
