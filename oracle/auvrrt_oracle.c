/*
 * TEST INFRASTRUCTURE ONLY -- fp64 CPU restatement ("oracle") of the RRT planning hot path of
 * hmc-lair-shark-tracking/auv-sim.  NOT product code: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference leg may load it, as the checker or the CPU baseline.
 *
 * Every function cites the reference lines it restates (paths relative to /root/reference).
 * The reference is CPython: floats are IEEE doubles, `a ** 2` is libm pow(a, 2.0), math.sin/cos/
 * sqrt/atan2 are libm, random.uniform(a, b) is `a + (b - a) * random()`.  This file is compiled
 * with -ffp-contract=off and calls the same libm, so it is expected to agree with the reference
 * BIT FOR BIT; tests/test_oracle_golden.py pins it against tests/golden/ (outputs of the
 * unmodified reference under oracle/harness.py).
 *
 * Parity status: PINNED for nn / arc steer / collision / cost / exploring (golden vectors from
 * the reference itself).  PARITY UNPINNED for the six-word Dubins solver (orc_dubins_*): the
 * reference only hints at it in a commented-out call into the third-party PyPI `dubins` module
 * (path_planning/rrt_dubins.py:238-251; module absent, unpinned); it restates the published
 * Shkel & Lumelsky normalised formulation and is validated by closure/symmetry properties.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
#include <unistd.h>

#define ORC_OK 0
#define ORC_NO_PATH 1      /* TypeError: opt_path is None (rrt_dubins.py:174) */
#define ORC_ZERO_DIV 2     /* ZeroDivisionError (rrt_dubins.py:270, :281) */
#define ORC_KEY_ERROR 3    /* KeyError / IndexError when uniform() returns its end point (:123-126) */
#define ORC_STREAM_END 4   /* explicit uniform stream exhausted */

typedef struct {
    int K; const double *circles;   /* [K][3] x, y, size, in obstacle_list order */
    int E; const double *poly;      /* [E][2] boundary polygon vertices (open ring) */
    int H; const double *habitats;  /* [H][3] x, y, size */
    int T; const double *bins;      /* [T][2] shark-grid time bins in dict order */
    int C; const double *cells;     /* [C][4] minx, miny, maxx, maxy in dict order */
    const double *probs;            /* [T][C] */
} orc_world_t;

/* ---------------------------------------------------------------- uniform stream ----------- */
typedef struct {
    const double *ext; int64_t n_ext;   /* explicit pre-generated stream, or NULL */
    uint64_t key; int f32u;           /* else: the counter-hash stream; f32u != 0: the fp32 build's 23-bit u */
    int64_t pos; int exhausted;
} orc_stream_t;

static uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
/* The pre-generated sample sequence (restated from DESIGN.md section 3, independently of the product's
 * csrc/common.cuh; tests check that harness, oracle and product agree): position k of seed's stream is
 * the 64-bit word hi:lo with  hi = M32a(c * 0x9E3779B9 + key_lo), lo = M32b(c * 0x85EBCA77 + key_hi),
 * c = k + 1 mod 2^32, key = SplitMix64 finaliser of (seed + 1) * golden.  u = (word >> 11) * 2^-53, or
 * (f32 != 0, what the fp32 build consumes) the top 23 bits, (word >> 41) * 2^-23. */
uint64_t orc_stream_key(uint64_t seed) { return mix64((seed + 1) * 0x9E3779B97F4A7C15ULL); }
static uint64_t stream_word(uint64_t key, int64_t k) {
    uint32_t c = (uint32_t)(k + 1);
    uint32_t a = c * 0x9E3779B9u + (uint32_t)key;
    a ^= a >> 16; a *= 0x21F0AAADu; a ^= a >> 15; a *= 0x735A2D97u;
    uint32_t b = c * 0x85EBCA77u + (uint32_t)(key >> 32);
    b ^= b >> 15; b *= 0xD168AAADu; b ^= b >> 15; b *= 0xAF723597u;
    return ((uint64_t)a << 32) | (uint64_t)b;
}
double orc_stream_u(uint64_t seed, int64_t k, int f32) {
    uint64_t z = stream_word(orc_stream_key(seed), k);
    return f32 ? (double)(z >> 41) * 0x1.0p-23 : (double)(z >> 11) * 0x1.0p-53;
}
static double next_u(orc_stream_t *s) {
    int64_t k = s->pos++;
    if (s->ext) {
        if (k >= s->n_ext) { s->exhausted = 1; return 0.5; }
        return s->ext[k];
    }
    uint64_t z = stream_word(s->key, k);
    return s->f32u ? (double)(z >> 41) * 0x1.0p-23 : (double)(z >> 11) * 0x1.0p-53;
}
/* random.uniform(a, b): CPython Lib/random.py */
static double uniform(orc_stream_t *s, double a, double b) { return a + (b - a) * next_u(s); }


/* ---------------------------------------------------------------- tiny parallel-for -------- */
/* pthreads, dynamic chunks off an atomic counter (no OpenMP runtime needed on the GPU box). */
typedef void (*orc_body_t)(int64_t i, void *ctx);
typedef struct { orc_body_t body; void *ctx; int64_t n, chunk; int64_t next; } orc_pf_t;
static void *orc_pf_worker(void *arg) {
    orc_pf_t *pf = (orc_pf_t *)arg;
    for (;;) {
        int64_t b = __atomic_fetch_add(&pf->next, pf->chunk, __ATOMIC_RELAXED);
        if (b >= pf->n) break;
        int64_t e = b + pf->chunk < pf->n ? b + pf->chunk : pf->n;
        for (int64_t i = b; i < e; i++) pf->body(i, pf->ctx);
    }
    return NULL;
}
int orc_num_threads(void) { long n = sysconf(_SC_NPROCESSORS_ONLN); return n > 0 ? (int)n : 1; }
static void orc_parallel_for(int64_t n, int64_t chunk, int nthreads, orc_body_t body, void *ctx) {
    if (nthreads <= 0) nthreads = orc_num_threads();
    if (nthreads > 256) nthreads = 256;
    orc_pf_t pf = {body, ctx, n, chunk > 0 ? chunk : 1, 0};
    if (nthreads == 1) { orc_pf_worker(&pf); return; }
    pthread_t th[256];
    int started = 0;
    for (int t = 0; t < nthreads; t++) if (pthread_create(&th[started], NULL, orc_pf_worker, &pf) == 0) started++;
    if (started == 0) orc_pf_worker(&pf);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
}

/* ---------------------------------------------------------------- get_closest_mps ---------- */
/* path_planning/rrt_dubins.py:505-513 with get_distance_angle :558-564: strict <, seeded with
 * index 0, compares sqrt(dx**2 + dy**2). */
int orc_nn(const double *tx, const double *ty, int n, double qx, double qy) {
    int best = 0;
    double dx = qx - tx[0], dy = qy - ty[0];
    double min_dist = sqrt(pow(dx, 2.0) + pow(dy, 2.0));
    for (int i = 0; i < n; i++) {
        dx = qx - tx[i]; dy = qy - ty[i];
        double d = sqrt(pow(dx, 2.0) + pow(dy, 2.0));
        if (d < min_dist) { min_dist = d; best = i; }
    }
    return best;
}
void orc_nn_batch(const double *tx, const double *ty, int n, const double *q, int nq, int *out) {
    for (int k = 0; k < nq; k++) out[k] = orc_nn(tx, ty, n, q[2 * k], q[2 * k + 1]);
}

/* ---------------------------------------------------------------- RRT.steer (arc rollout) -- */
/* path_planning/rrt_dubins.py:252-295.  parent/leaf = (x, y, theta, traj_time_stamp, length);
 * wp rows = (x, y, theta, v, traj_time_stamp, length) for the APPENDED waypoints (path[1:]). */
typedef struct { double dist_to_end, diff_max, freq, min_dist, velocity; } orc_steer_params_t;

int orc_steer_arc(const double parent[5], orc_stream_t *rng, const orc_steer_params_t *p,
                  double leaf[5], double *wp, int *nwp) {
    double x = parent[0], y = parent[1], th = parent[2], t = parent[3], len = parent[4];
    int n = 0, status = ORC_OK;
    double n_expand = floor(uniform(rng, 0.0, p->freq) / 1.0);                      /* :259-260 */
    for (int i = 0; i < (int)n_expand; i++) {
        double dist = uniform(rng, 0.0, p->dist_to_end);                            /* :264 */
        double diff = uniform(rng, -p->diff_max, p->diff_max);                      /* :265 */
        if (fabs(dist) > fabs(diff)) {                                              /* :266 */
            double s1 = dist + diff, s2 = dist - diff;
            double den = -s1 + s2;
            if (den == 0.0) { status = ORC_ZERO_DIV; break; }
            double radius = (s1 + s2) / den;                                        /* :270 */
            if (2.0 * radius == 0.0) { status = ORC_ZERO_DIV; break; }
            double phi = (s1 + s2) / (2.0 * radius);                                /* :271 */
            double ori = th;
            th += phi;                                                              /* :274 */
            double dx = radius * (sin(th) - sin(ori));                              /* :275 */
            double dy = radius * (-cos(th) + cos(ori));                             /* :276 */
            x += dx; y += dy;
            double vt = uniform(rng, 0.0, 2.0 * p->velocity);                       /* :279 */
            double movement = sqrt(pow(dx, 2.0) + pow(dy, 2.0));                    /* :280 */
            if (vt == 0.0) { status = ORC_ZERO_DIV; break; }
            t += movement / vt;                                                     /* :281 */
            len += movement;
            if (movement >= p->min_dist) {                                          /* :283 */
                double *w = wp + 6 * n++;
                w[0] = x; w[1] = y; w[2] = th; w[3] = vt; w[4] = t; w[5] = len;
            }
        } else if (!rng->ext) {
            /* slot addressing of the counter stream: a primitive owns three positions
             * (dist, diff, velocity_temp); the reference never draws velocity_temp here, so the slot is left
             * unread.  An explicit (recorded) stream is flat: nothing to skip. */
            rng->pos++;
        }
    }
    leaf[0] = x; leaf[1] = y; leaf[2] = th; leaf[3] = t; leaf[4] = len;
    *nwp = n;
    if (rng->exhausted) status = ORC_STREAM_END;
    return status;
}

/* steer on an explicit uniform array (golden replay); returns status, *nused = uniforms eaten */
int orc_steer_arc_ext(const double parent[5], const double *u, int64_t nu,
                      const orc_steer_params_t *p, double leaf[5], double *wp, int *nwp,
                      int64_t *nused) {
    orc_stream_t s = {u, nu, 0, 0, 0, 0};
    int st = orc_steer_arc(parent, &s, p, leaf, wp, nwp);
    *nused = s.pos;
    return st;
}

/* ---------------------------------------------------------------- exact point-in-polygon --- */
/* shapely Point.within(Polygon) (rrt_dubins.py:544-547): strictly interior, boundary excluded,
 * decided with exact orientation predicates (GEOS).  Float filter + exact expansion fallback. */
static void two_sum(double a, double b, double *s, double *e) {
    double x = a + b, bv = x - a, av = x - bv;
    *s = x; *e = (a - av) + (b - bv);
}
static int grow_expansion(int n, double *e, double b) {
    double h[40]; int m = 0; double q = b;
    for (int i = 0; i < n; i++) {
        double s, err; two_sum(q, e[i], &s, &err);
        if (err != 0.0) h[m++] = err;
        q = s;
    }
    if (q != 0.0 || m == 0) h[m++] = q;
    memcpy(e, h, sizeof(double) * (size_t)m);
    return m;
}
/* sign of (ax-cx)(by-cy) - (ay-cy)(bx-cx), exactly */
int orc_orient2d(double ax, double ay, double bx, double by, double cx, double cy) {
    double detl = (ax - cx) * (by - cy), detr = (ay - cy) * (bx - cx), det = detl - detr, detsum;
    if (detl > 0.0) { if (detr <= 0.0) return (det > 0) - (det < 0); detsum = detl + detr; }
    else if (detl < 0.0) { if (detr >= 0.0) return (det > 0) - (det < 0); detsum = -detl - detr; }
    else detsum = fabs(detr);
    const double errbound = (3.0 + 16.0 * 0x1.0p-53) * 0x1.0p-53;
    if (fabs(det) > errbound * detsum) return (det > 0) - (det < 0);
    /* det = ax*by - ax*cy - cx*by - ay*bx + ay*cx + cy*bx  (the cx*cy terms cancel) */
    const double pa[6] = {ax, -ax, -cx, -ay, ay, cy};
    const double pb[6] = {by, cy, by, bx, cx, bx};
    double e[40]; int n = 0;
    for (int i = 0; i < 6; i++) {
        double hi = pa[i] * pb[i], lo = fma(pa[i], pb[i], -hi);
        n = grow_expansion(n, e, lo);
        n = grow_expansion(n, e, hi);
    }
    double top = e[n - 1];
    return (top > 0) - (top < 0);
}
/* +1 strictly inside, 0 on the boundary, -1 outside */
int orc_point_in_polygon(const double *poly, int E, double px, double py) {
    int inside = 0;
    for (int i = 0; i < E; i++) {
        double ax = poly[2 * i], ay = poly[2 * i + 1];
        int j = (i + 1) % E;
        double bx = poly[2 * j], by = poly[2 * j + 1];
        if (px == ax && py == ay) return 0;
        if (ay == py && by == py) {
            if (fmin(ax, bx) <= px && px <= fmax(ax, bx)) return 0;
            continue;
        }
        if ((ay > py) != (by > py)) {
            int s = orc_orient2d(ax, ay, bx, by, px, py);
            if (s == 0) return 0;
            if ((s > 0) == (by > ay)) inside = !inside;
        }
    }
    return inside ? 1 : -1;
}

/* ---------------------------------------------------------------- RRT.check_collision ------ */
/* path_planning/rrt_dubins.py:530-549.  pts = mps.path as [n][2].  dList is NEVER reset between
 * obstacles (:535-542), so obstacle k is tested against the running minimum over obstacles 0..k.
 * Returns 1 = safe (True), 0 = collision / outside. */
int orc_check_collision(const double *pts, int n, const orc_world_t *w) {
    double running = INFINITY;
    for (int k = 0; k < w->K; k++) {
        double ox = w->circles[3 * k], oy = w->circles[3 * k + 1], sz = w->circles[3 * k + 2];
        for (int i = 0; i < n; i++) {
            double dx = pts[2 * i] - ox, dy = pts[2 * i + 1] - oy;
            double d = sqrt(pow(dx, 2.0) + pow(dy, 2.0));
            if (d < running) running = d;
        }
        if (n == 0) return -1; /* min([]) raises ValueError in the reference */
        if (running <= sz) return 0;
    }
    for (int i = 0; i < n; i++)
        if (orc_point_in_polygon(w->poly, w->E, pts[2 * i], pts[2 * i + 1]) <= 0) return 0;
    return 1;
}
/* check_collision_obstacle (:551-556): one point, per-obstacle test, no running minimum */
int orc_check_collision_obstacle(double x, double y, const orc_world_t *w) {
    for (int k = 0; k < w->K; k++) {
        double dx = x - w->circles[3 * k], dy = y - w->circles[3 * k + 1];
        if (sqrt(pow(dx, 2.0) + pow(dy, 2.0)) <= w->circles[3 * k + 2]) return 0;
    }
    return 1;
}

/* ---------------------------------------------------------------- habitat_shark_cost_func -- */
/* builtin sum() over [c0, c1, c2] as CPython >= 3.12 evaluates it (Python/bltinmodule.c,
 * builtin_sum_impl): the int start 0 absorbs c0 exactly, then the float fast path adds the rest
 * with Neumaier-compensated summation and folds the compensation in at the end.  (CPython < 3.12
 * would give ((0 + c0) + c1) + c2; the build and GPU-box interpreter is 3.12.3.) */
static double py312_sum3(double c0, double c1, double c2) {
    double f = 0.0 + c0, c = 0.0;
    const double xs[2] = {c1, c2};
    for (int i = 0; i < 2; i++) {
        double x = xs[i], t = f + x;
        if (fabs(f) >= fabs(x)) c += (f - t) + x; else c += (x - t) + f;
        f = t;
    }
    if (c != 0.0 && isfinite(c)) f += c;
    return f;
}

/* path_planning/cost.py:145-207.  pts = [n][3] (x, y, traj_time_stamp) in the order the caller
 * passes the path.  bin_mask (may be NULL) selects the bins the planner kept
 * (rrt_dubins.py:161-166).  out = [sum, c0, c1, c2].  The cell test reproduces the reference's
 * typo `mps.x <= cell_bound[3]` (cost.py:182). */
void orc_cost(const double *pts, int n, double total_traj_time, const orc_world_t *w,
              const uint8_t *bin_mask, const double weight[3], double out[4]) {
    double w1 = weight[0], w2 = weight[1], w3 = weight[2];
    double c0 = 0.0, c1 = 0.0, c2 = 0.0;
    uint8_t *visited = (uint8_t *)calloc((size_t)(w->H > 0 ? w->H : 1), 1);
    for (int i = 0; i < n; i++) {
        double x = pts[3 * i], y = pts[3 * i + 1], t = pts[3 * i + 2];
        int tb = -1;
        for (int b = 0; b < w->T; b++) {
            if (bin_mask && !bin_mask[b]) continue;
            if (t >= w->bins[2 * b] && t <= w->bins[2 * b + 1]) { tb = b; break; }
        }
        if (tb < 0) continue;                                                  /* cost.py:178-179 */
        for (int c = 0; c < w->C; c++) {
            const double *cb = w->cells + 4 * c;
            if (x >= cb[0] && x <= cb[2] && y >= cb[1] && x <= cb[3]) {           /* cost.py:182 */
                c2 += w3 * w->probs[(size_t)tb * w->C + c];
                break;
            }
        }
        for (int h = 0; h < w->H; h++) {
            double dx = w->habitats[3 * h] - x, dy = w->habitats[3 * h + 1] - y;
            double dist = sqrt(pow(dx, 2.0) + pow(dy, 2.0));
            if (dist <= w->habitats[3 * h + 2]) { visited[h] = 1; c1 += w2; break; }
        }
    }
    if (total_traj_time > 0) { c1 = c1 / total_traj_time; c2 = c2 / total_traj_time; }
    int count = 0;
    for (int h = 0; h < w->H; h++) count += visited[h];
    if (w->H != 0) c0 = w1 * count / w->H;
    free(visited);
    out[1] = c0; out[2] = c1; out[3] = c2;
    out[0] = py312_sum3(c0, c1, c2);                                          /* sum(cost) */
}

/* habitat_shark_cost_point (cost.py:209-241): `visited[i] == True` is a comparison, not an
 * assignment (:234), so visited never changes.  AUV grid = row `tb` of probs. */
double orc_cost_point(double x, double y, const orc_world_t *w, const uint8_t *visited, int tb,
                      const double weight[3]) {
    double c0 = 0.0, c1 = 0.0, c2 = 0.0;
    for (int h = 0; h < w->H; h++) {
        double dx = w->habitats[3 * h] - x, dy = w->habitats[3 * h + 1] - y;
        if (sqrt(pow(dx, 2.0) + pow(dy, 2.0)) <= w->habitats[3 * h + 2]) {
            if (!visited[h]) c0 += weight[0] / w->H;
            c1 += weight[1] / w->H;
        }
    }
    for (int c = 0; c < w->C; c++) {
        const double *cb = w->cells + 4 * c;
        if (x >= cb[0] && x <= cb[2] && y >= cb[1] && x <= cb[3]) {
            c2 += weight[2] * w->probs[(size_t)tb * w->C + c];
            break;
        }
    }
    return py312_sum3(c0, c1, c2);
}

/* ---------------------------------------------------------------- RRT.exploring ------------ */
typedef struct {
    int iterations;          /* budget: number of steer calls */
    int mode;                /* 0: traj_time_stamp and plan_time (time-bin pick, :122-127)
                                1: plan_time False (get_random_mps + get_closest_mps, :136-139)
                                2: plan_time and not traj_time_stamp (get_closest_mps_time, :129-132) with
                                   time.time() - t_start == (steer calls completed) * max_plan_time / iterations */
    double bin_interval, v, max_traj_time;
    double dist_to_end, diff_max, freq, min_dist;   /* RRT.__init__ defaults 2, .5, 30; 0.5 (:141) */
    double weights[3];
    double max_plan_time;
    /* mode 3 (Dubins-RRT with best-parent selection, see orc_exploring_dubins) */
    double dubins_rho, dubins_eta, near_radius;
    int dubins_w;
} orc_plan_params_t;

typedef struct {
    /* per steer call (caller allocates `iterations` rows; any pointer may be NULL) */
    int32_t *parent; uint8_t *safe; int32_t *nwp; double *leaf; int64_t *upos;
    double *cost_evals; int32_t n_cost_evals;   /* rows: iter, sum, c0, c1, c2, len(path) */
    int32_t best_iter, best_node, n_nodes; int64_t n_uniforms;
    double result[5];                           /* path length, sum, c0, c1, c2 */
    double *path; int32_t path_cap, n_path;     /* best path root->leaf, rows (x,y,theta,v,t,len) */
    int64_t n_waypoints_total;                  /* sum over steer calls of len(new.path) */
} orc_trace_t;

typedef struct { double x, y, th, v, t, len; } wp_t;
typedef struct { wp_t s; int parent; int npath; int path_off; double plan_stamp; } node_t;
typedef struct { int *v; int n, cap; } ilist_t;

static void il_push(ilist_t *l, int x) {
    if (l->n == l->cap) { l->cap = l->cap ? 2 * l->cap : 8; l->v = (int *)realloc(l->v, sizeof(int) * (size_t)l->cap); }
    l->v[l->n++] = x;
}
/* Python float `//` (Objects/floatobject.c float_floor_div / float_divmod) */
static double py_floordiv(double vx, double wx) {
    double mod = fmod(vx, wx), div = (vx - mod) / wx, fd;
    if (mod) { if ((wx < 0) != (mod < 0)) { mod += wx; div -= 1.0; } }
    if (div) { fd = floor(div); if (div - fd > 0.5) fd += 1.0; }
    else fd = copysign(0.0, vx / wx);
    return fd;
}

/* generate_final_course (:321-331): [leaf] + reversed(leaf.path) + reversed(parent.path) + ... */
static int final_course(const node_t *nodes, const wp_t *wps, int leaf, wp_t *out) {
    int n = 0, cur = leaf;
    out[n++] = nodes[leaf].s;
    while (nodes[cur].parent >= 0) {
        for (int k = nodes[cur].npath - 1; k >= 0; k--) out[n++] = wps[nodes[cur].path_off + k];
        out[n++] = nodes[nodes[cur].parent].s;           /* path[0] is the parent node object */
        cur = nodes[cur].parent;
    }
    return n;
}

int orc_exploring_dubins(const orc_world_t *w, const double start[5], orc_stream_t *rng,
                         const orc_plan_params_t *p, orc_trace_t *out);

int orc_exploring(const orc_world_t *w, const double start[5], orc_stream_t *rng,
                  const orc_plan_params_t *p, orc_trace_t *out) {
    if (p->mode == 3) return orc_exploring_dubins(w, start, rng, p, out);
    const int I = p->iterations;
    const int maxp = (int)ceil(p->freq) + 2;          /* most waypoints one steer call can append */
    node_t *nodes = (node_t *)malloc(sizeof(node_t) * (size_t)(I + 1));
    wp_t *wps = (wp_t *)malloc(sizeof(wp_t) * ((size_t)I + 1) * (size_t)maxp);
    wp_t *course = (wp_t *)malloc(sizeof(wp_t) * ((size_t)I * (size_t)(maxp + 1) + 2));
    double *cpts = (double *)malloc(sizeof(double) * 3 * ((size_t)I * (size_t)(maxp + 1) + 2));
    double *pts_buf = (double *)malloc(sizeof(double) * 2 * (size_t)(maxp + 2));
    uint8_t *bin_mask = (uint8_t *)malloc((size_t)(w->T > 0 ? w->T : 1));
    int n_nodes = 0, n_wps = 0, status = ORC_OK;
    nodes[n_nodes++] = (node_t){{start[0], start[1], start[2], 0.0, start[3], start[4]}, -1, 0, 0, 0.0};

    int time_expand = (int)ceil(p->max_traj_time / p->bin_interval);               /* :111 */
    ilist_t *bins = (ilist_t *)calloc((size_t)time_expand + 2, sizeof(ilist_t));
    il_push(&bins[1], 0);                                                          /* :114 */

    double minx = INFINITY, miny = INFINITY, maxx = -INFINITY, maxy = -INFINITY;
    for (int i = 0; i < w->E; i++) {
        minx = fmin(minx, w->poly[2 * i]); maxx = fmax(maxx, w->poly[2 * i]);
        miny = fmin(miny, w->poly[2 * i + 1]); maxy = fmax(maxy, w->poly[2 * i + 1]);
    }
    orc_steer_params_t sp = {p->dist_to_end, p->diff_max, p->freq, p->min_dist, p->v};
    double opt_cost[4] = {INFINITY, 0, 0, 0}, opt_len = 0.0;
    int opt_node = -1, opt_iter = -1;
    out->n_cost_evals = 0; out->n_waypoints_total = 0;
    int it = 0;
    int64_t upos0 = 0;   /* stream position after the previous steer call's iteration */
    int64_t guard = 0, guard_max = 64LL * I + 1024;
    while (it < I && guard++ < guard_max) {
        int parent;
        if (p->mode == 0) {
            int ran_bin = (int)uniform(rng, 1.0, (double)(time_expand + 1));       /* :123 */
            if (ran_bin > time_expand) { status = ORC_KEY_ERROR; break; }
            while (bins[ran_bin].n == 0) {
                ran_bin = (int)uniform(rng, 1.0, (double)(time_expand + 1));       /* :125 */
                if (ran_bin > time_expand || rng->exhausted) break;
            }
            if (ran_bin > time_expand) { status = ORC_KEY_ERROR; break; }
            if (rng->exhausted) { status = ORC_STREAM_END; break; }
            int ran_index = (int)uniform(rng, 0.0, (double)bins[ran_bin].n);       /* :126 */
            if (ran_index >= bins[ran_bin].n) { status = ORC_KEY_ERROR; break; }
            parent = bins[ran_bin].v[ran_index];
        } else if (p->mode == 2) {
            double ran_time = uniform(rng, 0.0, p->max_plan_time * p->freq);       /* :129 */
            int lo = 0, len = n_nodes;                                             /* get_closest_mps_time :515-528 */
            while (len > 3) {
                double left_diff = fabs(nodes[lo + len / 2 - 1].plan_stamp - ran_time);
                double right_diff = fabs(nodes[lo + len / 2 + 1].plan_stamp - ran_time);
                int index = len / 2;
                if (left_diff >= right_diff) { lo += index; len -= index; } else len = index;
            }
            parent = lo;
            if (nodes[parent].s.t > p->max_traj_time) continue;                    /* :131-132 */
        } else {
            double rx = uniform(rng, minx, maxx);                                  /* :336-339 */
            double ry = uniform(rng, miny, maxy);
            (void)uniform(rng, -M_PI, M_PI);
            (void)uniform(rng, 0.0, 15.0);
            int best = 0;                                                          /* :505-513 */
            double dx = rx - nodes[0].s.x, dy = ry - nodes[0].s.y;
            double md = sqrt(pow(dx, 2.0) + pow(dy, 2.0));
            for (int i = 0; i < n_nodes; i++) {
                dx = rx - nodes[i].s.x; dy = ry - nodes[i].s.y;
                double d = sqrt(pow(dx, 2.0) + pow(dy, 2.0));
                if (d < md) { md = d; best = i; }
            }
            parent = best;
            if (nodes[parent].s.t > p->max_traj_time) continue;                    /* :138-139 */
        }
        /* steer (:141) */
        double par[5] = {nodes[parent].s.x, nodes[parent].s.y, nodes[parent].s.th, nodes[parent].s.t, nodes[parent].s.len};
        double leaf[5]; int nwp;
        int st = orc_steer_arc(par, rng, &sp, leaf, (double *)(wps + n_wps), &nwp);
        if (st != ORC_OK) { status = st; break; }
        /* check_collision over path = [parent] + appended (:143) */
        pts_buf[0] = par[0]; pts_buf[1] = par[1];
        for (int k = 0; k < nwp; k++) { pts_buf[2 * k + 2] = wps[n_wps + k].x; pts_buf[2 * k + 3] = wps[n_wps + k].y; }
        int safe = orc_check_collision(pts_buf, nwp + 1, w);
        if (out->parent) out->parent[it] = parent;
        if (out->safe) out->safe[it] = (uint8_t)safe;
        if (out->nwp) out->nwp[it] = nwp + 1;
        if (out->leaf) memcpy(out->leaf + 5 * (size_t)it, leaf, sizeof(double) * 5);
        if (out->upos) out->upos[it] = upos0;
        out->n_waypoints_total += nwp + 1;
        if (safe) {
            int id = n_nodes++;
            nodes[id] = (node_t){{leaf[0], leaf[1], leaf[2], 0.0, leaf[3], leaf[4]}, parent, nwp, n_wps,
                                 (double)it * (p->max_plan_time / (double)I)};   /* time.time() - t_start at steer entry (:252) */
            n_wps += nwp;
            /* time-bin insert (:147-151), only `if traj_time_stamp` */
            double fd = p->mode == 2 ? 0.0 : py_floordiv(leaf[3], p->bin_interval);
            double curr_bin = (fd + 1.0) * p->bin_interval;
            double fidx = fd + 1.0;
            if (p->mode == 2) {
            } else if (curr_bin > p->max_traj_time) {
                if (fidx >= 1.0 && fidx <= (double)time_expand) {      /* resets a live key */
                    int bi = (int)fidx;
                    if (bi * p->bin_interval == curr_bin) { bins[bi].n = 0; il_push(&bins[bi], id); }
                }
            } else {
                if (!(fidx >= 1.0 && fidx <= (double)time_expand)) { status = ORC_KEY_ERROR; break; }
                il_push(&bins[(int)fidx], id);
            }
            if (leaf[3] >= p->max_traj_time - 30) {                                /* :158 */
                int n = final_course(nodes, wps, id, course);
                double t0 = start[3], t1 = leaf[3];
                for (int b = 0; b < w->T; b++) {                                   /* :164-166 */
                    double b0 = w->bins[2 * b], b1 = w->bins[2 * b + 1];
                    bin_mask[b] = (uint8_t)((t0 >= b0 && t0 <= b1) || (b0 >= t0 && b1 <= t1) || (t1 >= b0 && t1 <= b1));
                }
                for (int k = 0; k < n; k++) { cpts[3 * k] = course[k].x; cpts[3 * k + 1] = course[k].y; cpts[3 * k + 2] = course[k].t; }
                double c[4];
                orc_cost(cpts, n, leaf[3], w, bin_mask, p->weights, c);            /* :168 */
                if (out->cost_evals) {
                    double *r = out->cost_evals + 6 * (size_t)out->n_cost_evals;
                    r[0] = it; r[1] = c[0]; r[2] = c[1]; r[3] = c[2]; r[4] = c[3]; r[5] = n;
                }
                out->n_cost_evals++;
                if (c[0] < opt_cost[0]) {                                          /* :169 */
                    memcpy(opt_cost, c, sizeof(c)); opt_len = leaf[4]; opt_node = id; opt_iter = it;
                }
            }
        }
        it++;
        upos0 = rng->pos;
        if (rng->exhausted) { status = ORC_STREAM_END; break; }
    }
    out->n_nodes = n_nodes; out->n_uniforms = rng->pos;
    out->best_iter = opt_iter; out->best_node = opt_node; out->n_path = 0;
    if (status == ORC_OK && opt_node < 0) status = ORC_NO_PATH;
    if (opt_node >= 0) {
        out->result[0] = opt_len; memcpy(out->result + 1, opt_cost, sizeof(opt_cost));
        int n = final_course(nodes, wps, opt_node, course);
        out->n_path = n;
        if (out->path) for (int k = 0; k < n && k < out->path_cap; k++)            /* :174 reverse */
            memcpy(out->path + 6 * (size_t)k, &course[n - 1 - k], sizeof(wp_t));
    }
    for (int i = 0; i < time_expand + 2; i++) free(bins[i].v);
    free(bins); free(nodes); free(wps); free(course); free(cpts); free(bin_mask); free(pts_buf);
    return status;
}

/* Many independent queries (configs 2/5), OpenMP over queries: the CPU baseline.
 * starts [Q][5]; seeds [Q]; out_result [Q][5]; out_counts [Q][3] = nodes, cost evals, waypoints */
typedef struct {
    const orc_world_t *w; const double *starts; const uint64_t *seeds; const orc_plan_params_t *p;
    int f32u; double *out_result; int64_t *out_counts; int32_t *out_status;
} orc_eb_t;
static void orc_eb_body(int64_t q, void *vc) {
    orc_eb_t *c = (orc_eb_t *)vc;
    orc_stream_t rng = {NULL, 0, orc_stream_key(c->seeds[q]), c->f32u, 0, 0};
    orc_trace_t tr; memset(&tr, 0, sizeof(tr));
    int st = orc_exploring(c->w, c->starts + 5 * (size_t)q, &rng, c->p, &tr);
    if (c->out_status) c->out_status[q] = st;
    if (c->out_result) memcpy(c->out_result + 5 * (size_t)q, tr.result, sizeof(double) * 5);
    if (c->out_counts) { c->out_counts[3 * q] = tr.n_nodes; c->out_counts[3 * q + 1] = tr.n_cost_evals; c->out_counts[3 * q + 2] = tr.n_waypoints_total; }
}
int orc_exploring_batch(const orc_world_t *w, const double *starts, const uint64_t *seeds, int Q,
                        const orc_plan_params_t *p, int f32u, double *out_result,
                        int64_t *out_counts, int32_t *out_status, int nthreads) {
    orc_eb_t c = {w, starts, seeds, p, f32u, out_result, out_counts, out_status};
    orc_parallel_for(Q, 1, nthreads, orc_eb_body, &c);
    return 0;
}
/* ---------------------------------------------------------------- Dubins six-word steer ---- */
/* PARITY UNPINNED (see file header).  Words in evaluation order LSL, LSR, RSL, RSR, RLR, LRL;
 * strict < on t+p+q so the first word wins ties. */
static double mod2pi(double t) { return t - 2.0 * M_PI * floor(t / (2.0 * M_PI)); }

static int dubins_word(int word, double alpha, double beta, double d, double o[3]) {
    double sa = sin(alpha), sb = sin(beta), ca = cos(alpha), cb = cos(beta), cab = cos(alpha - beta);
    double d2 = d * d, p2, p, t0, w;
    switch (word) {
    case 0: p2 = 2 + d2 - 2 * cab + 2 * d * (sa - sb); if (p2 < 0) return 0;
        t0 = atan2(cb - ca, d + sa - sb); o[0] = mod2pi(t0 - alpha); o[1] = sqrt(p2); o[2] = mod2pi(beta - t0); return 1;
    case 1: p2 = -2 + d2 + 2 * cab + 2 * d * (sa + sb); if (p2 < 0) return 0;
        p = sqrt(p2); t0 = atan2(-ca - cb, d + sa + sb) - atan2(-2.0, p);
        o[0] = mod2pi(t0 - alpha); o[1] = p; o[2] = mod2pi(t0 - mod2pi(beta)); return 1;
    case 2: p2 = -2 + d2 + 2 * cab - 2 * d * (sa + sb); if (p2 < 0) return 0;
        p = sqrt(p2); t0 = atan2(ca + cb, d - sa - sb) - atan2(2.0, p);
        o[0] = mod2pi(alpha - t0); o[1] = p; o[2] = mod2pi(beta - t0); return 1;
    case 3: p2 = 2 + d2 - 2 * cab + 2 * d * (sb - sa); if (p2 < 0) return 0;
        t0 = atan2(ca - cb, d - sa + sb); o[0] = mod2pi(alpha - t0); o[1] = sqrt(p2); o[2] = mod2pi(t0 - beta); return 1;
    case 4: w = (6. - d2 + 2 * cab + 2 * d * (sa - sb)) / 8.; if (fabs(w) > 1) return 0;
        t0 = atan2(ca - cb, d - sa + sb); p = mod2pi(2 * M_PI - acos(w));
        o[0] = mod2pi(alpha - t0 + mod2pi(p / 2.)); o[1] = p; o[2] = mod2pi(alpha - beta - o[0] + mod2pi(p)); return 1;
    default: w = (6. - d2 + 2 * cab + 2 * d * (sb - sa)) / 8.; if (fabs(w) > 1) return 0;
        t0 = atan2(ca - cb, d + sa - sb); p = mod2pi(2 * M_PI - acos(w));
        o[0] = mod2pi(-alpha - t0 + p / 2.); o[1] = p; o[2] = mod2pi(mod2pi(beta) - alpha - o[0] + mod2pi(p)); return 1;
    }
}
/* returns the word id (0..5) or -1; params = (t, p, q) in units of rho; *length = rho (t+p+q) */
int orc_dubins_shortest(const double q0[3], const double q1[3], double rho, double params[3],
                        double *length, double all_len[6]) {
    double dx = q1[0] - q0[0], dy = q1[1] - q0[1], D = sqrt(dx * dx + dy * dy), d = D / rho;
    double th = d > 0 ? mod2pi(atan2(dy, dx)) : 0.0;
    double alpha = mod2pi(q0[2] - th), beta = mod2pi(q1[2] - th);
    int best = -1; double bestc = INFINITY;
    for (int wd = 0; wd < 6; wd++) {
        double o[3];
        if (all_len) all_len[wd] = -1.0;
        if (!dubins_word(wd, alpha, beta, d, o)) continue;
        double c = o[0] + o[1] + o[2];
        if (all_len) all_len[wd] = c * rho;
        if (c < bestc) { bestc = c; best = wd; params[0] = o[0]; params[1] = o[1]; params[2] = o[2]; }
    }
    *length = best >= 0 ? bestc * rho : INFINITY;
    return best;
}
static const char DUBINS_SEG[6][3] = {{'L','S','L'},{'L','S','R'},{'R','S','L'},{'R','S','R'},{'R','L','R'},{'L','R','L'}};
static void dubins_segment(double t, const double qi[3], char type, double qt[3]) {
    double st = sin(qi[2]), ct = cos(qi[2]);
    if (type == 'L') { qt[0] = sin(qi[2] + t) - st; qt[1] = -cos(qi[2] + t) + ct; qt[2] = t; }
    else if (type == 'R') { qt[0] = -sin(qi[2] - t) + st; qt[1] = cos(qi[2] - t) - ct; qt[2] = -t; }
    else { qt[0] = ct * t; qt[1] = st * t; qt[2] = 0.0; }
    qt[0] += qi[0]; qt[1] += qi[1]; qt[2] += qi[2];
}
/* configuration at arclength s along the path */
void orc_dubins_sample(const double q0[3], double rho, int word, const double params[3], double s,
                       double q[3]) {
    double tp = s / rho, qi[3] = {0.0, 0.0, q0[2]}, a[3], b[3];
    const char *ty = DUBINS_SEG[word];
    dubins_segment(params[0], qi, ty[0], a);
    dubins_segment(params[1], a, ty[1], b);
    if (tp < params[0]) dubins_segment(tp, qi, ty[0], q);
    else if (tp < params[0] + params[1]) dubins_segment(tp - params[0], a, ty[1], q);
    else dubins_segment(tp - params[0] - params[1], b, ty[2], q);
    q[0] = q[0] * rho + q0[0]; q[1] = q[1] * rho + q0[1]; q[2] = mod2pi(q[2]);
}
/* One Dubins edge of the config-4 micro-bench: W waypoints = samples at s = k L/(W-1),
 * k = 0..W-2, then q1 itself (the commented block appends to_mps, rrt_dubins.py:248);
 * then check_collision on those points.  wp [W][3].  Returns safe flag; fills word/length. */
int orc_edge_dubins(const orc_world_t *w, const double q0[3], const double q1[3], double rho, int W,
                    double *wp, int *word, double params[3], double *length) {
    double pts[2 * 64];
    *word = orc_dubins_shortest(q0, q1, rho, params, length, NULL);
    if (*word < 0 || W > 64) return 0;
    double step = *length / (double)(W - 1);
    for (int k = 0; k < W - 1; k++) orc_dubins_sample(q0, rho, *word, params, (double)k * step, wp + 3 * k);
    memcpy(wp + 3 * (W - 1), q1, sizeof(double) * 3);
    for (int k = 0; k < W; k++) { pts[2 * k] = wp[3 * k]; pts[2 * k + 1] = wp[3 * k + 1]; }
    return orc_check_collision(pts, W, w);
}
/* ---------------------------------------------------------------- mode 3: Dubins-RRT, best parent ------ */
/* PARITY UNPINNED.  The reference only hints at a Dubins steer (commented-out call into the absent PyPI `dubins`
 * module, path_planning/rrt_dubins.py:238-251) driven from the nearest-node branch of the loop (:136-141); the
 * north star asks for "the six-word Dubins steer from each tree node to each sampled AUV state ... with
 * min-reductions for best-parent selection".  This is the build's own definition, restated
 * here independently of csrc/plan.cu and the straightforward way: every candidate's path cost is evaluated by
 * generate_final_course + habitat_shark_cost_func on the explicit waypoint list, not incrementally.
 *
 * One iteration:
 *  1. sample = get_random_mps (:333-343): x, y uniform in the boundary's bounding box, theta in [-pi, pi], size
 *     (4 draws, size unused).
 *  2. candidate parents: the nearest node (get_closest_mps :505-513, strict <, lowest index) and, in index order,
 *     the nodes within near_radius of the sample, at most 32 candidates in all.  If the nearest node's
 *     traj_time_stamp exceeds max_traj_time the iteration is skipped (`continue`, :138-139); other candidates past
 *     it are dropped.
 *  3. per candidate: six-word Dubins path (turning radius dubins_rho) to the sample, cut at arclength
 *     s_end = min(L, dubins_eta); W = dubins_w waypoints at s_k = k s_end / (W-1), k = 0..W-1 (k = 0 is the parent);
 *     the new node is the last waypoint; traj_time_stamp advances by s / v, length by s.
 *  4. RRT.check_collision on the W points; unsafe candidates are out.
 *  5. cost of the path root -> new node through that parent (cost.habitat_shark_cost_func on generate_final_course,
 *     with the planner's bin filter :161-166); the candidate with the smallest total wins, ties to the lowest node
 *     index.  The node is appended with that parent.
 *  6. if its traj_time_stamp >= max_traj_time - 30 it is a candidate plan (:158-171, strict <).
 * Trace rows: parent = chosen parent (-1: no safe candidate), safe, nwp = W, leaf = the new node (through the
 * nearest node when nothing was safe), upos = stream position before the iteration's draws. */
int orc_exploring_dubins(const orc_world_t *w, const double start[5], orc_stream_t *rng,
                         const orc_plan_params_t *p, orc_trace_t *out) {
    const int I = p->iterations, W = p->dubins_w, NC = 32;
    if (W < 2 || W > 32 || !(p->dubins_rho > 0) || !(p->dubins_eta > 0) || !(p->v > 0)) return ORC_KEY_ERROR;
    node_t *nodes = (node_t *)malloc(sizeof(node_t) * (size_t)(I + 2));
    wp_t *wps = (wp_t *)malloc(sizeof(wp_t) * ((size_t)I + 2) * (size_t)W);
    wp_t *course = (wp_t *)malloc(sizeof(wp_t) * ((size_t)(I + 2) * (size_t)(W + 1) + 2));
    double *cpts = (double *)malloc(sizeof(double) * 3 * ((size_t)(I + 2) * (size_t)(W + 1) + 2));
    uint8_t *bin_mask = (uint8_t *)malloc((size_t)(w->T > 0 ? w->T : 1));
    int n_nodes = 0, n_wps = 0, status = ORC_OK;
    nodes[n_nodes++] = (node_t){{start[0], start[1], start[2], 0.0, start[3], start[4]}, -1, 0, 0, 0.0};
    double minx = INFINITY, miny = INFINITY, maxx = -INFINITY, maxy = -INFINITY;
    for (int i = 0; i < w->E; i++) {
        minx = fmin(minx, w->poly[2 * i]); maxx = fmax(maxx, w->poly[2 * i]);
        miny = fmin(miny, w->poly[2 * i + 1]); maxy = fmax(maxy, w->poly[2 * i + 1]);
    }
    double opt_cost[4] = {INFINITY, 0, 0, 0}, opt_len = 0.0;
    int opt_node = -1, opt_iter = -1;
    out->n_cost_evals = 0; out->n_waypoints_total = 0;
    int it = 0;
    int64_t upos0 = 0, guard = 0, guard_max = 64LL * I + 1024;
    while (it < I && guard++ < guard_max) {
        const double sx = uniform(rng, minx, maxx), sy = uniform(rng, miny, maxy);     /* :336-339 */
        const double sth = uniform(rng, -M_PI, M_PI);
        (void)uniform(rng, 0.0, 15.0);
        int nearest = 0;                                                               /* :505-513 */
        {
            double dx = sx - nodes[0].s.x, dy = sy - nodes[0].s.y, md = sqrt(pow(dx, 2.0) + pow(dy, 2.0));
            for (int i = 0; i < n_nodes; i++) {
                dx = sx - nodes[i].s.x; dy = sy - nodes[i].s.y;
                double d = sqrt(pow(dx, 2.0) + pow(dy, 2.0));
                if (d < md) { md = d; nearest = i; }
            }
        }
        if (nodes[nearest].s.t > p->max_traj_time) { upos0 = rng->pos; continue; }     /* :138-139 */
        int cand[32], nc = 0;
        cand[nc++] = nearest;
        for (int i = 0; i < n_nodes && nc < NC; i++) {
            if (i == nearest) continue;
            double dx = sx - nodes[i].s.x, dy = sy - nodes[i].s.y;
            if (sqrt(pow(dx, 2.0) + pow(dy, 2.0)) <= p->near_radius) cand[nc++] = i;
        }
        int best = -1; double best_c[4] = {INFINITY, 0, 0, 0}; wp_t best_wp[32]; double best_send = 0.0;
        wp_t near_leaf = nodes[nearest].s;
        const double q1[3] = {sx, sy, sth};
        for (int c = 0; c < nc; c++) {
            const int pi_ = cand[c];
            const node_t *pn = &nodes[pi_];
            if (pn->s.t > p->max_traj_time) continue;
            const double q0[3] = {pn->s.x, pn->s.y, pn->s.th};
            double prm[3], L;
            const int word = orc_dubins_shortest(q0, q1, p->dubins_rho, prm, &L, NULL);
            if (word < 0) continue;
            const double s_end = L < p->dubins_eta ? L : p->dubins_eta;
            wp_t wp[32]; double pts[64];
            for (int k = 0; k < W; k++) {
                const double sk = (double)k * (s_end / (double)(W - 1));
                double q[3];
                orc_dubins_sample(q0, p->dubins_rho, word, prm, sk, q);
                wp[k] = (wp_t){q[0], q[1], q[2], p->v, pn->s.t + sk / p->v, pn->s.len + sk};
                pts[2 * k] = q[0]; pts[2 * k + 1] = q[1];
            }
            out->n_waypoints_total += W;
            if (pi_ == nearest) near_leaf = wp[W - 1];
            if (!orc_check_collision(pts, W, w)) continue;
            /* tentatively hang the node under this parent and cost the whole path */
            nodes[n_nodes] = (node_t){wp[W - 1], pi_, W - 1, n_wps, 0.0};
            memcpy(wps + n_wps, wp + 1, sizeof(wp_t) * (size_t)(W - 1));
            const int n = final_course(nodes, wps, n_nodes, course);
            const double t0 = start[3], t1 = wp[W - 1].t;
            for (int b = 0; b < w->T; b++) {                                           /* :164-166 */
                double b0 = w->bins[2 * b], b1 = w->bins[2 * b + 1];
                bin_mask[b] = (uint8_t)((t0 >= b0 && t0 <= b1) || (b0 >= t0 && b1 <= t1) || (t1 >= b0 && t1 <= b1));
            }
            for (int k = 0; k < n; k++) { cpts[3 * k] = course[k].x; cpts[3 * k + 1] = course[k].y; cpts[3 * k + 2] = course[k].t; }
            double cc[4];
            orc_cost(cpts, n, t1, w, bin_mask, p->weights, cc);
            if (cc[0] < best_c[0] || (cc[0] == best_c[0] && best >= 0 && pi_ < best)) {
                memcpy(best_c, cc, sizeof(cc)); best = pi_; memcpy(best_wp, wp, sizeof(wp_t) * (size_t)W); best_send = s_end;
            }
        }
        (void)best_send;
        if (out->parent) out->parent[it] = best;
        if (out->safe) out->safe[it] = (uint8_t)(best >= 0);
        if (out->nwp) out->nwp[it] = W;
        if (out->upos) out->upos[it] = upos0;
        if (out->leaf) {
            const wp_t lf = best >= 0 ? best_wp[W - 1] : near_leaf;
            double *o = out->leaf + 5 * (size_t)it;
            o[0] = lf.x; o[1] = lf.y; o[2] = lf.th; o[3] = lf.t; o[4] = lf.len;
        }
        if (best >= 0) {
            const int id = n_nodes++;
            nodes[id] = (node_t){best_wp[W - 1], best, W - 1, n_wps, 0.0};
            memcpy(wps + n_wps, best_wp + 1, sizeof(wp_t) * (size_t)(W - 1));
            n_wps += W - 1;
            if (best_wp[W - 1].t >= p->max_traj_time - 30) {                           /* :158 */
                if (out->cost_evals) {
                    double *r = out->cost_evals + 6 * (size_t)out->n_cost_evals;
                    r[0] = it; r[1] = best_c[0]; r[2] = best_c[1]; r[3] = best_c[2]; r[4] = best_c[3];
                    r[5] = final_course(nodes, wps, id, course);
                }
                out->n_cost_evals++;
                if (best_c[0] < opt_cost[0]) {                                         /* :169 */
                    memcpy(opt_cost, best_c, sizeof(best_c)); opt_len = best_wp[W - 1].len; opt_node = id; opt_iter = it;
                }
            }
        }
        it++;
        upos0 = rng->pos;
        if (rng->exhausted) { status = ORC_STREAM_END; break; }
    }
    out->n_nodes = n_nodes; out->n_uniforms = rng->pos;
    out->best_iter = opt_iter; out->best_node = opt_node; out->n_path = 0;
    if (status == ORC_OK && opt_node < 0) status = ORC_NO_PATH;
    if (opt_node >= 0) {
        out->result[0] = opt_len; memcpy(out->result + 1, opt_cost, sizeof(opt_cost));
        const int n = final_course(nodes, wps, opt_node, course);
        out->n_path = n;
        if (out->path) for (int k = 0; k < n && k < out->path_cap; k++)
            memcpy(out->path + 6 * (size_t)k, &course[n - 1 - k], sizeof(wp_t));
    }
    free(nodes); free(wps); free(course); free(cpts); free(bin_mask);
    return status;
}

typedef struct {
    const orc_world_t *w; const double *q0, *q1; double rho; int W; uint8_t *safe, *word; double *length;
} orc_db_t;
static void orc_db_body(int64_t i, void *vc) {
    orc_db_t *c = (orc_db_t *)vc;
    double wp[3 * 64], prm[3], len; int wd;
    int s = orc_edge_dubins(c->w, c->q0 + 3 * i, c->q1 + 3 * i, c->rho, c->W, wp, &wd, prm, &len);
    if (c->safe) c->safe[i] = (uint8_t)s;
    if (c->word) c->word[i] = (uint8_t)wd;
    if (c->length) c->length[i] = len;
}
void orc_edges_dubins_batch(const orc_world_t *w, const double *q0, const double *q1, int64_t n,
                            double rho, int W, uint8_t *safe, uint8_t *word, double *length,
                            int nthreads) {
    orc_db_t c = {w, q0, q1, rho, W, safe, word, length};
    orc_parallel_for(n, 256, nthreads, orc_db_body, &c);
}
/* arc-steer edges on the counter stream: edge i uses stream seed seeds[i] from position 0 */
typedef struct {
    const orc_world_t *w; const double *parents; const uint64_t *seeds; const orc_steer_params_t *sp;
    int f32u; uint8_t *safe; int32_t *nwp_out; double *leaf_out;
} orc_ab_t;
static void orc_ab_body(int64_t i, void *vc) {
    orc_ab_t *c = (orc_ab_t *)vc;
    orc_stream_t rng = {NULL, 0, orc_stream_key(c->seeds[i]), c->f32u, 0, 0};
    const int maxp = (int)ceil(c->sp->freq) + 2;
    double leaf[5]; int nwp;
    double *wp = (double *)malloc(sizeof(double) * 8 * (size_t)maxp), *pts = wp + 6 * (size_t)maxp;
    orc_steer_arc(c->parents + 5 * i, &rng, c->sp, leaf, wp, &nwp);
    pts[0] = c->parents[5 * i]; pts[1] = c->parents[5 * i + 1];
    for (int k = 0; k < nwp; k++) { pts[2 * k + 2] = wp[6 * k]; pts[2 * k + 3] = wp[6 * k + 1]; }
    int s = orc_check_collision(pts, nwp + 1, c->w);
    if (c->safe) c->safe[i] = (uint8_t)s;
    if (c->nwp_out) c->nwp_out[i] = nwp + 1;
    if (c->leaf_out) memcpy(c->leaf_out + 5 * i, leaf, sizeof(leaf));
    free(wp);
}
void orc_edges_arc_batch(const orc_world_t *w, const double *parents, const uint64_t *seeds,
                         int64_t n, const orc_steer_params_t *sp, int f32u, uint8_t *safe,
                         int32_t *nwp_out, double *leaf_out, int nthreads) {
    orc_ab_t c = {w, parents, seeds, sp, f32u, safe, nwp_out, leaf_out};
    orc_parallel_for(n, 256, nthreads, orc_ab_body, &c);
}

/* ---------------------------------------------------------------- SharkOccupancyGrid.convert -- */
/* path_planning/sharkOccupancyGrid.py:47-71 (convert), :145-172 (constructGrid), :174-204
 * (constructAUVGrid), :206-243 (constructSharkOccupancyGrid), :245-299 (time bins, convert2DArr),
 * :307-320 (createBinList).  Cells are polygons (cell_off/cell_xy, in cell_list order); a point
 * belongs to the FIRST cell with `point.within(cell) or cell.touches(point)` (:232), i.e. closed
 * containment decided exactly.  tracks: shark s = points trk_off[s]..trk_off[s+1], rows (x, y, t).
 * out_grid [T][rows][cols] = resultArr; returns T (or -1 if the buffer is too small). */
void orc_occupancy_dims(const double bounds[4], double cell_size, double bin_interval, const double *trk,
                        const int64_t *trk_off, int S, int *T, int *rows, int *cols) {
    double longest = 0;
    for (int s = 0; s < S; s++)
        if (trk_off[s + 1] > trk_off[s] && trk[3 * (trk_off[s + 1] - 1) + 2] > longest) longest = trk[3 * (trk_off[s + 1] - 1) + 2];
    *T = (int)floor(longest / bin_interval);                                           /* :314 */
    *cols = (int)(ceil(bounds[2] - bounds[0]) / cell_size) + 1;                           /* :160 */
    *rows = (int)(ceil(bounds[3] - bounds[1]) / cell_size) + 1;
}
static int cell_contains_closed(const double *xy, int n, double px, double py) {
    return orc_point_in_polygon(xy, n, px, py) >= 0;
}
int orc_occupancy(const double *cell_xy, const int64_t *cell_off, int C, const double bounds[4], double cell_size,
                  double bin_interval, double detect_range, const double *trk, const int64_t *trk_off, int S,
                  double *out_grid, int64_t out_cap) {
    int T, rows, cols;
    orc_occupancy_dims(bounds, cell_size, bin_interval, trk, trk_off, S, &T, &rows, &cols);
    if ((int64_t)T * rows * cols > out_cap) return -1;
    const int count = (int)ceil(detect_range / cell_size);                             /* :189 */
    int *crow = (int *)malloc(sizeof(int) * (size_t)(C > 0 ? C : 1)), *ccol = (int *)malloc(sizeof(int) * (size_t)(C > 0 ? C : 1));
    for (int c = 0; c < C; c++) {                                                      /* cellToIndex :294-299 */
        double lowx = INFINITY, lowy = INFINITY;
        for (int64_t k = cell_off[c]; k < cell_off[c + 1]; k++) { lowx = fmin(lowx, cell_xy[2 * k]); lowy = fmin(lowy, cell_xy[2 * k + 1]); }
        ccol[c] = (int)((lowx - bounds[0]) / cell_size);
        crow[c] = (int)((lowy - bounds[1]) / cell_size);
    }
    double *occ = (double *)malloc(sizeof(double) * (size_t)rows * cols), *auv = (double *)malloc(sizeof(double) * (size_t)rows * cols);
    for (int b = 0; b < T; b++) {
        double *grid = out_grid + (size_t)b * rows * cols;
        for (int i = 0; i < rows * cols; i++) grid[i] = 0.0;
        for (int s = 0; s < S; s++) {
            /* constructSharkOccupancyGrid on the points of shark s that fall in bin b (first-match bin, :253-258) */
            for (int i = 0; i < rows * cols; i++) occ[i] = 0.0;
            for (int c = 0; c < C; c++) occ[crow[c] * cols + ccol[c]] = 0.01;
            int64_t npts = 0;
            for (int64_t k = trk_off[s]; k < trk_off[s + 1]; k++) {
                double t = trk[3 * k + 2];
                int tb = -1;
                for (int bb = 0; bb < T; bb++) if (t >= bb * bin_interval && t <= (bb + 1) * bin_interval) { tb = bb; break; }
                if (tb != b) continue;
                npts++;
                for (int c = 0; c < C; c++)
                    if (cell_contains_closed(cell_xy + 2 * cell_off[c], (int)(cell_off[c + 1] - cell_off[c]), trk[3 * k], trk[3 * k + 1])) {
                        occ[crow[c] * cols + ccol[c]] += 1;
                        break;
                    }
            }
            double nor = (double)npts + C * 0.01;                                       /* :222 */
            for (int i = 0; i < rows * cols; i++) occ[i] = occ[i] / nor;
            /* constructAUVGrid */
            for (int i = 0; i < rows * cols; i++) auv[i] = 0.0;
            for (int c = 0; c < C; c++) {
                int row = crow[c], col = ccol[c];
                for (int i = 0; i < 4 * count; i++) {
                    int rt = row - 2 * count + i;
                    for (int j = 0; j < 4 * count; j++) {
                        int ct = col - 2 * count + j;
                        if (rt >= 0 && rt < rows && ct >= 0 && ct < cols)
                            if (sqrt(pow((double)(rt - row), 2.0) + pow((double)(ct - col), 2.0)) <= count)
                                auv[row * cols + col] += occ[rt * cols + ct];
                    }
                }
            }
            for (int i = 0; i < rows * cols; i++) grid[i] = grid[i] + auv[i];           /* :166 */
        }
        for (int i = 0; i < rows * cols; i++) grid[i] = grid[i] / S;                    /* :170 */
    }
    free(crow); free(ccol); free(occ); free(auv);
    return T;
}

/* ================================================================ gym_rrt Planner_RRT ====== */
/* gym_rrt/envs/rrt_dubins.py: the goal-directed planner RRTEnv.step drives (one node per step).
 * PINNED against tests/golden/gym_plan.npz (the unmodified class run on the sample sequence, with
 * random.choice(seq) = seq[int(random() * len(seq))]). */
typedef struct {
    double x0, y0, x1, y1;          /* boundary_point[0], boundary_point[1] */
    int K; const double *circles;   /* [K][3] obstacle x, y, size in list order */
    double goal_x, goal_y;
    double exp_rate, dist_to_end, diff_max, freq, cell_side;
    int subsections;
} orc_gym_world_t;

/* math.hypot(x, y) as CPython 3.12 evaluates it (Modules/mathmodule.c vector_norm): power-of-two
 * scaling, double-length accumulation of the squares, one differential correction step. */
double orc_py_hypot(double a, double b) {
    double v[2] = {fabs(a), fabs(b)};
    double max = v[0] > v[1] ? v[0] : v[1];
    if (isinf(v[0]) || isinf(v[1])) return INFINITY;
    if (isnan(v[0]) || isnan(v[1])) return NAN;
    if (max == 0.0) return max;
    int max_e; frexp(max, &max_e);
    if (max_e < -1023) return orc_py_hypot(ldexp(a, 1074), ldexp(b, 1074)) * 0x1.0p-1074; /* subnormal inputs; never on this path */
    double scale = ldexp(1.0, -max_e), csum = 1.0, frac1 = 0.0, frac2 = 0.0;
    for (int i = 0; i < 2; i++) {
        double x = v[i] * scale;
        double hi = x * x, lo = fma(x, x, -hi);
        double s = csum + hi, e = (csum - s) + hi;       /* dl_fast_sum: |csum| >= |hi| */
        csum = s; frac1 += lo; frac2 += e;
    }
    double h = sqrt(csum - 1.0 + (frac1 + frac2));
    double hi = -h * h, lo = fma(-h, h, -hi);
    double s = csum + hi, e = (csum - s) + hi;
    csum = s; frac1 += lo; frac2 += e;
    double x = csum - 1.0 + (frac1 + frac2);
    h += x / (2.0 * h);
    return h / scale;
}

/* Planner_RRT.angle_wrap (:420-428): recursion unrolled, same additions */
static double gym_angle_wrap(double a) {
    while (!(-M_PI <= a && a <= M_PI)) {
        if (a > M_PI) a += (-2.0 * M_PI);
        else if (a < -M_PI) a += (2.0 * M_PI);
        else return a; /* NaN */
    }
    return a;
}

/* discretize_env (:77-93): rows = int(height) // int(cell_side), cols likewise */
void orc_gym_grid_shape(const orc_gym_world_t *w, int *rows, int *cols) {
    long cs = (long)w->cell_side;
    long hh = (long)(w->y1 - w->y0), ww = (long)(w->x1 - w->x0);
    /* Python // floors; operands are non-negative here */
    *rows = cs > 0 ? (int)(hh / cs) : 0;
    *cols = cs > 0 ? (int)(ww / cs) : 0;
}

/* add_node_to_grid (:108-154): flat sub-cell id (row * cols + col) * subsections + sub, or -1 when
 * the node falls past the last row / column (the early returns), or -2 for an IndexError. */
static int gym_subcell(const orc_gym_world_t *w, int rows, int cols, double x, double y, double theta) {
    long row = (long)(y / w->cell_side), col = (long)(x / w->cell_side);             /* :115-116 */
    if (row >= rows) return -1;                                                     /* :118 */
    if (col >= cols) return -1;                                                     /* :122 */
    if (row < 0) { row += rows; if (row < 0) return -2; }                           /* list[-k] */
    if (col < 0) { col += cols; if (col < 0) return -2; }
    double delta = (2.0 * M_PI) / (double)w->subsections;                           /* grid_cell_rrt.py:52 */
    double raw = theta / delta;                                                     /* :127 */
    long sub = (long)floor(raw);                                                    /* :130 */
    if (sub < 0) sub = (long)(w->subsections + sub);                                /* :134-135 */
    if (sub == w->subsections) sub -= 1;                                            /* :137-152 (after the input() prompt) */
    if (sub < 0) { sub += w->subsections; if (sub < 0) return -2; }
    if (sub >= w->subsections) return -2;
    return (int)((row * cols + col) * w->subsections + sub);
}

/* check_within_boundary (:458-471) */
static int gym_within(const orc_gym_world_t *w, double x, double y) {
    return (x >= w->x0) && (x <= w->x1) && (y >= w->y0) && (y <= w->y1);
}

/* check_collision_free (:430-449) on pts[n][>=2] with row stride `ld`: dList never reset */
static int gym_collision_free(const orc_gym_world_t *w, const double *pts, int n, int ld) {
    double running = INFINITY;
    for (int k = 0; k < w->K; k++) {
        double ox = w->circles[3 * k], oy = w->circles[3 * k + 1];
        if (n == 0) return -1;   /* min([]) ValueError */
        for (int i = 0; i < n; i++) {
            double dx = pts[ld * i] - ox, dy = pts[ld * i + 1] - oy;                /* get_distance_angle(obstacle, point) */
            double d = sqrt(pow(dx, 2.0) + pow(dy, 2.0));
            if (d < running) running = d;
        }
        if (running <= w->circles[3 * k + 2]) return 0;
    }
    for (int i = 0; i < n; i++)
        if (!gym_within(w, pts[ld * i], pts[ld * i + 1])) return 0;
    return 1;
}

/* connect_to_goal_curve_alt (:375-418).  Returns 1 and the arc description, 0 for `return None`,
 * ORC_ZERO_DIV+100 for a ZeroDivisionError.  arc = {x_C, y_C, radius, ang_vel, theta_0, length},
 * point i = (x_C + r sin(w i + th0), y_C - r cos(w i + th0), w i + th0), i = 0..n_expand. */
static int gym_goal_arc(const orc_gym_world_t *w, double x, double y, double th, double arc[6], int *n_expand) {
    double theta0 = th;
    double gdx = w->goal_x - x, gdy = w->goal_y - y;
    double theta = atan2(gdy, gdx);                                                 /* get_distance_angle :485 */
    double diff = gym_angle_wrap(theta - theta0);
    if (fabs(diff) > M_PI / 2.0) return 0;                                          /* :382 */
    double r_G = orc_py_hypot(gdx, gdy);                                            /* :386 */
    double phi_G = atan2(gdy, gdx);
    double phi, radius;
    if (phi_G - th != 0.0) phi = 2.0 * gym_angle_wrap(phi_G - th); else return 0;   /* :391-395 */
    double sn = sin(phi_G - th);
    if (sn != 0.0) radius = r_G / (2.0 * sn); else return 0;                        /* :397-400 */
    double length = radius * phi;
    if (phi > M_PI) { phi -= 2.0 * M_PI; length = -radius * phi; }
    else if (phi < -M_PI) { phi += 2.0 * M_PI; length = -radius * phi; }
    double le = length / w->exp_rate;
    if (le == 0.0) return 100 + ORC_ZERO_DIV;
    double ang_vel = phi / le;                                                      /* :411 */
    arc[0] = x - radius * sin(th);                                                  /* :414 */
    arc[1] = y + radius * cos(th);
    arc[2] = radius; arc[3] = ang_vel; arc[4] = theta0; arc[5] = length;
    *n_expand = (int)floor(le);                                                     /* :417 */
    return 1;
}
static void gym_arc_point(const double arc[6], int i, double p[3]) {
    double a = arc[3] * (double)i + arc[4];
    p[0] = arc[0] + arc[2] * sin(a);
    p[1] = arc[1] - arc[2] * cos(a);
    p[2] = a;
}

#define GYM_TRACE_W 11
/* Planner_RRT.planning (:157-196) when actions == NULL, else the RRTEnv.step loop
 * (gym_rrt/envs/rrt_env.py:206-224) over flat sub-cell ids (empty cells are skipped).
 * nodes [max_step+1][4] = x, y, theta, traj_time_stamp; node_cell [max_step+1] flat sub-cell or -1;
 * occupied [max_step+1] flat ids in first-occupancy order; counts NULL or [rows*cols*subsections];
 * trace NULL or [max_step][11] = parent, nwp, accepted, done, n_nodes, n_occupied, n_uniforms (of generate_one_node),
 * candidate x, y, theta, t; path [path_cap][3] goal -> start (generate_final_course :318-328).
 * rec[0..5] = steps, found, n_nodes, n_occupied, n_path, last node; arc_len = final_node.length. */
int orc_gym_plan(const orc_gym_world_t *w, const double start[3], orc_stream_t *rng, int max_step,
                 const int32_t *actions, int32_t rec[6], double *nodes, int32_t *parents, int32_t *node_cell,
                 int32_t *occupied, int32_t *counts, double *trace, double *path, int path_cap, double *arc_len) {
    int rows, cols; orc_gym_grid_shape(w, &rows, &cols);
    int maxp = (int)ceil(w->freq) + 2;
    int cap = max_step + 1, status = ORC_OK;
    double *wps = (double *)malloc(sizeof(double) * 3 * (size_t)cap * maxp);
    int *nwp = (int *)calloc(cap, sizeof(int));
    int *cell_count = (int *)calloc(cap, sizeof(int));   /* per occupied entry */
    double *pts = (double *)malloc(sizeof(double) * 3 * (maxp + 1));
    int n = 1, n_occ = 0, steps = 0, found = 0, n_path = 0, goal_checked = -1;
    if (counts) memset(counts, 0, sizeof(int32_t) * (size_t)rows * cols * w->subsections);
    nodes[0] = start[0]; nodes[1] = start[1]; nodes[2] = start[2]; nodes[3] = 0.0; parents[0] = -1;
    *arc_len = 0.0;
    {   /* __init__ :57 add_node_to_grid(start) */
        int c = gym_subcell(w, rows, cols, start[0], start[1], start[2]);
        node_cell[0] = c;
        if (c == -2) { status = ORC_KEY_ERROR; goto done; }
        if (c >= 0) { occupied[0] = c; cell_count[0] = 1; n_occ = 1; if (counts) counts[c] = 1; }
    }
    for (int it = 0; it < max_step; it++) {
        int cell, oi = -1;
        if (!actions) {
            if (n_occ == 0) { status = ORC_KEY_ERROR; break; }                     /* random.choice([]) IndexError */
            oi = (int)(next_u(rng) * (double)n_occ);                               /* :176 */
            cell = occupied[oi];
        } else {
            cell = actions[it];
            for (int j = 0; j < n_occ; j++) if (occupied[j] == cell) { oi = j; break; }
            if (oi < 0) continue;                                                  /* node_array == [] (:209-215) */
        }
        /* generate_one_node :198-238 */
        int64_t pos0 = rng->pos;
        int k = (int)(next_u(rng) * (double)cell_count[oi]);                       /* :217 random.choice(node_array) */
        int pn = -1;
        for (int j = 0, seen = 0; j < n; j++) if (node_cell[j] == cell) { if (seen == k) { pn = j; break; } seen++; }
        /* steer :241-283 */
        double x = nodes[4 * pn], y = nodes[4 * pn + 1], th = nodes[4 * pn + 2], t = nodes[4 * pn + 3];
        int m = 0;
        pts[0] = x; pts[1] = y; pts[2] = th;                                       /* path[0] = mps */
        double n_expand = floor(uniform(rng, 0.0, w->freq) / 1.0);                 /* :251-252 */
        for (int i = 0; i < (int)n_expand; i++) {
            double dist = uniform(rng, 0.0, w->dist_to_end);
            double diff = uniform(rng, -w->diff_max, w->diff_max);
            if (fabs(dist) > fabs(diff)) {                                         /* :258 */
                double s1 = dist + diff, s2 = dist - diff, den = -s1 + s2;
                if (den == 0.0) { status = ORC_ZERO_DIV; break; }
                double radius = (s1 + s2) / den;
                if (2.0 * radius == 0.0) { status = ORC_ZERO_DIV; break; }
                double phi = (s1 + s2) / (2.0 * radius);
                double ori = th;
                th = gym_angle_wrap(th + phi);                                     /* :265 */
                double dx = radius * (sin(th) - sin(ori));
                double dy = radius * (-cos(th) + cos(ori));
                x += dx; y += dy;
                t += sqrt(pow(dx, 2.0) + pow(dy, 2.0)) / 1.0;                      /* velocity = 1 (:241, :274) */
                m++;
                pts[3 * m] = x; pts[3 * m + 1] = y; pts[3 * m + 2] = th;
            }
        }
        if (status != ORC_OK) break;
        if (rng->exhausted) { status = ORC_STREAM_END; break; }
        int accepted = 0, is_done = 0;
        int cf = gym_collision_free(w, pts, m + 1, 3);                             /* :222 */
        if (cf == 1) {
            nodes[4 * n] = x; nodes[4 * n + 1] = y; nodes[4 * n + 2] = th; nodes[4 * n + 3] = t;
            parents[n] = pn; nwp[n] = m;
            memcpy(wps + 3 * (size_t)n * maxp, pts + 3, sizeof(double) * 3 * m);
            int c = gym_subcell(w, rows, cols, x, y, th);                          /* :226 */
            node_cell[n] = c;
            if (c == -2) { status = ORC_KEY_ERROR; break; }
            if (c >= 0) {
                int oj = -1;
                for (int j = 0; j < n_occ; j++) if (occupied[j] == c) { oj = j; break; }
                if (oj < 0) { oj = n_occ++; occupied[oj] = c; cell_count[oj] = 0; } /* :153-154 */
                cell_count[oj]++;
                if (counts) counts[c]++;
            }
            n++; accepted = 1;
        }
        steps++;
        /* connect_to_goal_curve_alt(self.mps_list[-1]) (:229): a pure function of the last node, so
         * it is re-evaluated only when the last node changed */
        int last = n - 1;
        if (goal_checked != last) {
            goal_checked = last;
            double arc[6]; int ne = 0;
            int g = gym_goal_arc(w, nodes[4 * last], nodes[4 * last + 1], nodes[4 * last + 2], arc, &ne);
            if (g >= 100) { status = g - 100; break; }
            if (g == 1) {
                /* check_collision_free(final_node) (:232): obstacle-major running minimum over the arc points */
                int ok = 1; double running = INFINITY, p[3];
                if (ne + 1 <= 0 && w->K > 0) { status = ORC_KEY_ERROR; break; }
                for (int kk = 0; kk < w->K && ok; kk++) {
                    for (int i = 0; i <= ne; i++) {
                        gym_arc_point(arc, i, p);
                        double dx = p[0] - w->circles[3 * kk], dy = p[1] - w->circles[3 * kk + 1];
                        double d = sqrt(pow(dx, 2.0) + pow(dy, 2.0));
                        if (d < running) running = d;
                    }
                    if (running <= w->circles[3 * kk + 2]) ok = 0;
                }
                for (int i = 0; i <= ne && ok; i++) { gym_arc_point(arc, i, p); if (!gym_within(w, p[0], p[1])) ok = 0; }
                if (ok) {
                    is_done = 1; found = 1; *arc_len = arc[5];
                    /* generate_final_course (:318-328): [final] + reversed(final.path) + reversed(node.path) ... */
                    double p2[3];
                    if (ne >= 0) gym_arc_point(arc, ne, p2); else { p2[0] = nodes[4 * last]; p2[1] = nodes[4 * last + 1]; p2[2] = nodes[4 * last + 2]; }
                    #define GYM_PUSH(a, b, c) do { if (n_path < path_cap) { path[3 * n_path] = (a); path[3 * n_path + 1] = (b); path[3 * n_path + 2] = (c); } n_path++; } while (0)
                    GYM_PUSH(p2[0], p2[1], p2[2]);
                    for (int i = ne; i >= 0; i--) { gym_arc_point(arc, i, p2); GYM_PUSH(p2[0], p2[1], p2[2]); }
                    for (int j = last; parents[j] >= 0; j = parents[j]) {
                        const double *wj = wps + 3 * (size_t)j * maxp;
                        for (int i = nwp[j] - 1; i >= 0; i--) GYM_PUSH(wj[3 * i], wj[3 * i + 1], wj[3 * i + 2]);
                        int pj = parents[j];
                        GYM_PUSH(nodes[4 * pj], nodes[4 * pj + 1], nodes[4 * pj + 2]);
                    }
                    #undef GYM_PUSH
                }
            }
        }
        if (trace) {
            double *tr = trace + GYM_TRACE_W * (size_t)(steps - 1);
            tr[0] = pn; tr[1] = m; tr[2] = accepted; tr[3] = is_done; tr[4] = n; tr[5] = n_occ;
            tr[6] = (double)(rng->pos - pos0); tr[7] = x; tr[8] = y; tr[9] = th; tr[10] = t;
        }
        if (is_done) break;
    }
done:
    rec[0] = steps; rec[1] = found; rec[2] = n; rec[3] = n_occ; rec[4] = n_path; rec[5] = n - 1;
    if (status == ORC_OK && n_path > path_cap) status = 5; /* overflow */
    free(wps); free(nwp); free(cell_count); free(pts);
    return status;
}

/* batch of independent episodes (own start / goal / seed each), pthread-parallel: the CPU baseline */
typedef struct { const orc_gym_world_t *w; const double *starts, *goals; const uint64_t *seeds; int max_step;
                 int32_t *recs; int32_t *status; } orc_gym_batch_t;
static void orc_gym_body(int64_t q, void *vc) {
    orc_gym_batch_t *c = (orc_gym_batch_t *)vc;
    orc_gym_world_t w = *c->w; w.goal_x = c->goals[2 * q]; w.goal_y = c->goals[2 * q + 1];
    int cap = c->max_step + 1;
    double *nodes = (double *)malloc(sizeof(double) * 4 * cap);
    int32_t *ibuf = (int32_t *)malloc(sizeof(int32_t) * 3 * cap);
    double arc_len;
    orc_stream_t s = {NULL, 0, orc_stream_key(c->seeds[q]), 0, 0, 0};
    c->status[q] = orc_gym_plan(&w, c->starts + 3 * q, &s, c->max_step, NULL, c->recs + 6 * q, nodes, ibuf, ibuf + cap,
                                ibuf + 2 * cap, NULL, NULL, NULL, 0, &arc_len);
    if (c->status[q] == 5) c->status[q] = ORC_OK;   /* path not requested */
    free(nodes); free(ibuf);
}
void orc_gym_plan_batch(const orc_gym_world_t *w, const double *starts, const double *goals, const uint64_t *seeds,
                        int Q, int max_step, int nthreads, int32_t *recs, int32_t *status) {
    orc_gym_batch_t c = {w, starts, goals, seeds, max_step, recs, status};
    orc_parallel_for(Q, 4, nthreads, orc_gym_body, &c);
}

/* ================================================================ lattice A* (fixed length, SOG) == */
/* path_planning/astar_fixLenSOG.py: class astar, method astar (:551-657) with its helpers
 * (:178-264, :380-414, :436-547).  Deterministic (no random draws); IEEE add / mul / sqrt only, so the
 * restatement is expected to agree with the reference BIT FOR BIT.
 * PINNED against tests/golden/astar.npz (the unmodified module run by oracle/harness.py).
 * Inputs the reference derives through unpinned third parties are taken as inputs here: the boundary
 * centroid (shapely/GEOS, :191) and the cell bounds after Python's round(v, 2) (:494-495). */
typedef struct {
    int K; const double *circles;     /* [K][3] obstacle_list order */
    int E; const double *boundary;    /* [E][2] boundary_list corners */
    double centroid[2];
    int H; const double *habitats;    /* [H][3] habitat_list */
    int T; const double *bins;        /* [T][2] shark-grid time bins, dict order */
    int C; const double *cells_r;     /* [C][4] cell bounds rounded to 2 decimals, dict order */
    const double *probs;              /* [T][C] */
    const double *topn;               /* [T][C+1] prefix sums over probabilities sorted descending (orc_astar_topn) */
} orc_astar_world_t;

static int cmp_desc(const void *a, const void *b) {
    double x = *(const double *)a, y = *(const double *)b;
    return x < y ? 1 : (x > y ? -1 : 0);
}
/* get_top_n_prob (:520-536) for every n: sort descending, add one by one starting from int 0 */
void orc_astar_topn(const double *probs, int T, int C, double *topn) {
    double *tmp = (double *)malloc(sizeof(double) * (C > 0 ? C : 1));
    for (int t = 0; t < T; t++) {
        memcpy(tmp, probs + (size_t)t * C, sizeof(double) * C);
        qsort(tmp, C, sizeof(double), cmp_desc);
        double total = 0.0;
        topn[(size_t)t * (C + 1)] = 0.0;
        for (int i = 0; i < C; i++) { total += tmp[i]; topn[(size_t)t * (C + 1) + i + 1] = total; }
    }
    free(tmp);
}

/* same_side / point_in_triangle / within_bounds (:178-203); np.cross of 2-vectors = a0*b1 - a1*b0 */
static int astar_same_side(const double p1[2], const double p2[2], const double a[2], const double b[2]) {
    double ux = b[0] - a[0], uy = b[1] - a[1];
    double cp1 = ux * (p1[1] - a[1]) - uy * (p1[0] - a[0]);
    double cp2 = ux * (p2[1] - a[1]) - uy * (p2[0] - a[0]);
    return cp1 * cp2 >= 0.0;
}
static int astar_in_triangle(const double p[2], const double a[2], const double b[2], const double c[2]) {
    return astar_same_side(p, a, b, c) && astar_same_side(p, b, a, c) && astar_same_side(p, c, a, b);
}
int orc_astar_within_bounds(const orc_astar_world_t *w, double x, double y) {
    double p[2] = {x, y};
    for (int i = 0; i < w->E; i++) {
        const double *a = w->boundary + 2 * i, *b = w->boundary + 2 * ((i + 1) % w->E);
        if (astar_in_triangle(p, a, b, w->centroid)) return 1;
    }
    return 0;
}
/* collision_free (:205-221): per obstacle, no running minimum */
int orc_astar_collision_free(const orc_astar_world_t *w, double x, double y) {
    for (int k = 0; k < w->K; k++) {
        double dx = x - w->circles[3 * k], dy = y - w->circles[3 * k + 1];
        if (sqrt(pow(dx, 2.0) + pow(dy, 2.0)) <= w->circles[3 * k + 2]) return 0;
    }
    return 1;
}
/* get_cell_prob's key search (:486-505): first cell, in dict order, whose rounded box passes the abs test */
int orc_astar_find_cell(const orc_astar_world_t *w, double x, double y) {
    for (int c = 0; c < w->C; c++) {
        const double *b = w->cells_r + 4 * c;
        double dx = fabs(b[0] - b[2]), dy = fabs(b[1] - b[3]);
        if (fabs(x - b[0]) <= dx && fabs(x - b[2]) <= dx && fabs(y - b[1]) <= dy && fabs(y - b[3]) <= dy) return c;
    }
    return -1;
}
static double astar_euclid(double ax, double ay, double bx, double by) {             /* module-level :18-29 */
    double dx = fabs(ax - bx), dy = fabs(ay - by);
    return sqrt(dx * dx + dy * dy);
}
/* Walkable (:223-246); -1 when the loop could not terminate */
static int astar_walkable(const orc_astar_world_t *w, double cx, double cy, double px, double py) {
    double sx = cx, sy = cy;
    double stepx = (double)(long)(fabs(px - cx) / 5.0), stepy = (double)(long)(fabs(py - cy) / 5.0);
    long guard = 0;
    while (sx <= px && sy <= py) {
        double ix = sx, iy = sy;
        sx += stepx; sy += stepy;
        if (!orc_astar_collision_free(w, ix, iy)) return 0;
        if (++guard > 10000000L) return -1;
    }
    return 1;
}

#define ASTAR_PATH_W 6
/* rec = {n_expanded, n_nodes, n_path, n_smooth}; path rows start -> goal: x, y, pathLen, time_stamp, cost, f;
 * keep[i] = 1 when smoothPath (:416-452) keeps trajectory point i; expand_order NULL or [node_cap] node ids in
 * the order they left the open list; node_xy NULL or [node_cap][2].  Status: ORC_OK, ORC_NO_PATH (open list ran
 * empty: returns None), ORC_KEY_ERROR (AttributeError :489 no time bin / TypeError :602 no cell / IndexError
 * :532, :419, :652), 5 = node_cap / path_cap too small. */
int orc_astar(const orc_astar_world_t *w, const double start[2], double velocity, double limit, const double weights[4],
              int node_cap, int path_cap, int32_t rec[4], double *cost_out, double *path, uint8_t *keep,
              int32_t *expand_order, double *node_xy) {
    static const double DX[8] = {0, 0, -10, 10, -10, -10, 10, 10}, DY[8] = {-10, 10, 0, 0, -10, 10, -10, 10};   /* :255 */
    const double w2 = weights[1], w3 = weights[2], w4 = weights[3];
    double *nx = (double *)malloc(sizeof(double) * 5 * (size_t)node_cap), *ny = nx + node_cap, *nlen = ny + node_cap,
           *nf = nlen + node_cap, *ncost = nf + node_cap;
    int32_t *nts = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)node_cap), *npar = nts + node_cap;
    uint8_t *alive = (uint8_t *)calloc(node_cap, 1);
    uint8_t *visited = (uint8_t *)calloc(600 * 600, 1);                              /* :121 */
    int n = 0, n_exp = 0, status = ORC_NO_PATH, goal = -1;
    rec[0] = rec[1] = rec[2] = rec[3] = 0; *cost_out = 0.0;
    nx[0] = start[0]; ny[0] = start[1]; nlen[0] = 0.0; nf[0] = 0.0; ncost[0] = 0.0; nts[0] = 0; npar[0] = -1; alive[0] = 1; n = 1;
    for (;;) {
        int cur = -1;
        for (int i = 0; i < n; i++)                                                  /* :577-583 first strict minimum */
            if (alive[i] && (cur < 0 || nf[i] < nf[cur])) cur = i;
        if (cur < 0) break;                                                          /* while len(open_list) > 0 */
        alive[cur] = 0;
        if (expand_order) expand_order[n_exp] = cur;
        n_exp++;
        if (fabs(nlen[cur] - limit) <= 10.0) { goal = cur; status = ORC_OK; break; }  /* :588 */
        for (int d = 0; d < 8; d++) {
            double px = nx[cur] + DX[d], py = ny[cur] + DY[d];                       /* :261 */
            if (!orc_astar_within_bounds(w, px, py)) continue;
            if (!orc_astar_collision_free(w, px, py)) continue;
            double plen = nlen[cur] + astar_euclid(nx[cur], ny[cur], px, py);        /* :636 */
            double dist_left = fabs(limit - plen);
            long ts = (long)(plen / velocity);                                       /* :638 int() */
            int tb = -1;
            for (int t = 0; t < w->T; t++)                                           /* findCurrSOG :454-468 */
                if ((double)ts <= w->bins[2 * t + 1] && (double)ts >= w->bins[2 * t]) { tb = t; break; }
            if (tb < 0) { status = ORC_KEY_ERROR; goto done; }
            int cell = orc_astar_find_cell(w, px, py);
            if (cell < 0) { status = ORC_KEY_ERROR; goto done; }
            double g = ncost[cur] - w4 * w->probs[(size_t)tb * w->C + cell];          /* :641 */
            long nn = (long)dist_left;
            if (nn > w->C) { status = ORC_KEY_ERROR; goto done; }                     /* probabilities[index] IndexError */
            double h = -w2 * dist_left - w3 * (double)w->H - w4 * w->topn[(size_t)tb * (w->C + 1) + nn];   /* :644 */
            double f = g + h;
            long xi = (long)(px + 500.0), yi = (long)(py + 200.0);                   /* get_indices :405-414 */
            if (xi < 0) xi += 600;
            if (yi < 0) yi += 600;
            if (xi < 0 || xi >= 600 || yi < 0 || yi >= 600) { status = ORC_KEY_ERROR; goto done; }
            if (!visited[xi * 600 + yi]) {                                           /* :654-656 */
                if (n >= node_cap) { status = 5; goto done; }
                nx[n] = px; ny[n] = py; nlen[n] = plen; nf[n] = f; ncost[n] = g; nts[n] = (int32_t)ts; npar[n] = cur; alive[n] = 1;
                n++;
                visited[xi * 600 + yi] = 1;
            }
        }
    }
done:
    rec[0] = n_exp; rec[1] = n;
    if (node_xy) for (int i = 0; i < n; i++) { node_xy[2 * i] = nx[i]; node_xy[2 * i + 1] = ny[i]; }
    if (status == ORC_OK) {
        int np_ = 0;
        for (int j = goal; j >= 0; j = npar[j]) np_++;
        rec[2] = np_;
        *cost_out = ncost[goal];                                                     /* cost[0] */
        if (np_ > path_cap) status = 5;
        else if (np_ < 2) status = ORC_KEY_ERROR;                                     /* smoothPath: trajectory[1] IndexError */
        else {
            int i = np_ - 1;
            for (int j = goal; j >= 0; j = npar[j], i--) {
                double *r = path + ASTAR_PATH_W * i;
                r[0] = nx[j]; r[1] = ny[j]; r[2] = nlen[j]; r[3] = (double)nts[j]; r[4] = ncost[j]; r[5] = nf[j];
                keep[i] = 1;
            }
            /* smoothPath (:416-452) */
            int index = 0, check = 0, curp;
            index += 1; curp = index;
            while (index < np_ - 1) {
                int wk = astar_walkable(w, path[ASTAR_PATH_W * check], path[ASTAR_PATH_W * check + 1],
                                        path[ASTAR_PATH_W * curp], path[ASTAR_PATH_W * curp + 1]);
                if (wk < 0) { status = ORC_KEY_ERROR; break; }
                if (wk) {
                    int inside = 0;
                    for (int hh = 0; hh < w->H && !inside; hh++)
                        if (astar_euclid(w->habitats[3 * hh], w->habitats[3 * hh + 1], path[ASTAR_PATH_W * curp],
                                         path[ASTAR_PATH_W * curp + 1]) <= w->habitats[3 * hh + 2]) inside = 1;
                    if (!inside) keep[curp] = 0;
                    index += 1; curp = index;
                } else {
                    check = curp; index += 1; curp = index;
                }
            }
            int ns = 0;
            for (int k = 0; k < np_; k++) ns += keep[k];
            rec[3] = ns;
        }
    }
    free(nx); free(nts); free(alive); free(visited);
    return status;
}

/* batch of independent queries {start x, y, path_len_limit, w1..w4, velocity}[8], pthread-parallel */
typedef struct { const orc_astar_world_t *w; const double *queries; int node_cap, path_cap; int32_t *recs; double *cost;
                 int32_t *status; } orc_astar_batch_t;
static void orc_astar_body(int64_t q, void *vc) {
    orc_astar_batch_t *c = (orc_astar_batch_t *)vc;
    const double *Q = c->queries + 8 * q;
    double *path = (double *)malloc(sizeof(double) * ASTAR_PATH_W * (size_t)c->path_cap);
    uint8_t *keep = (uint8_t *)malloc(c->path_cap);
    c->status[q] = orc_astar(c->w, Q, Q[7], Q[2], Q + 3, c->node_cap, c->path_cap, c->recs + 4 * q, c->cost + q, path, keep, NULL, NULL);
    free(path); free(keep);
}
void orc_astar_batch(const orc_astar_world_t *w, const double *queries, int Q, int node_cap, int path_cap, int nthreads,
                     int32_t *recs, double *cost, int32_t *status) {
    orc_astar_batch_t c = {w, queries, node_cap, path_cap, recs, cost, status};
    orc_parallel_for(Q, 1, nthreads, orc_astar_body, &c);
}
