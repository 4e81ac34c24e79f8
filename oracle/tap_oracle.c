/*
 * tap_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, CPU restatement of the TAP-Net packing-environment step
 * (Juzhan/TAP-Net @ 6eded31).  It exists only as the checker for the CUDA path:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it.  The product (tap-net_b200/) never links, loads
 * or calls anything in this file and has no CPU fallback.
 *
 * The restatement is deliberately LITERAL: it keeps the reference's voxel
 * container (0 empty / -1 "empty under a block" / k+1 block id), its EMS lists,
 * its shared `visited` list and, for MACS, the incrementally edited
 * `level_free_space` interval lists.  The CUDA kernels use a reduced
 * heightmap-only formulation; agreement between the two is what the parity
 * tests establish.
 *
 * Pinning: validated against (a) the golden vectors G1-G5 of SURVEY.md section 4,
 * (b) fixtures generated from the imported Python reference
 * (tests/golden/make_golden.py) and (c) live differential runs against the
 * Python reference when /root/reference is present (tests/test_oracle_vs_reference.py).
 * 2D paths are fully pinned.  3D: everything except the ConvexHull +
 * matplotlib Path.contains_point branch of tools.is_stable is pinned; that
 * branch is "parity unpinned" (matplotlib is absent here).
 *
 * Reference citations are file:line into /root/reference.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#include <pthread.h>

#define TAPO_MAXW 32   /* container width / length limit of the oracle */
#define TAPO_MAXN 256  /* blocks per episode limit */

enum { TAPO_LB_GREEDY = 0, TAPO_MACS = 1, TAPO_LB = 2 };
enum { TAPO_HM_FULL = 0, TAPO_HM_ZERO = 1, TAPO_HM_DIFF = 2 };

typedef struct { int len; int v[TAPO_MAXW + 4]; } ilist;
typedef struct { int len; int v[TAPO_MAXN + 4]; } xlist;   /* LB: per-level x lists grow by appends (tools.py:1745-1749) */

typedef struct tapo_env {
    int dim, W, L, H, n;
    int strategy, hm_type;
    char reward_type[48];
    int hard, useP, useS, mcs_in, mcs_start;
    int *container;   /* [W][H] or [W][L][H]   (tools.py:3629) */
    int *heightmap;   /* [W] or [W][L]         (tools.py:3630) */
    int *positions;   /* [n][dim]              (tools.py:3628) */
    int *blocks;      /* [n][dim] appended     (tools.py:3631,3674) */
    unsigned char *stable; /* [n]              (tools.py:3633) */
    long long valid, empty; /* tools.py:3635-3636 */
    int k;            /* current_blocks_num tools.py:3653 */
    ilist *lfs;       /* [H] MACS 2D level_free_space tools.py:3641-3643 */
    ilist *lfs3;      /* [H][L] MACS 3D level_free_space tools.py:3644-3648, every row [0, W-1] */
    xlist *lbl;       /* LB: [H] (2D) or [H][L] (3D) level_free_space tools.py:3649-3653, initial [0] */
    int error;        /* sticky: 1 = numpy would have raised IndexError, 2 = limits */
} tapo_env;

/* ------------------------------------------------------------------ */
/* reward-type string tests, exactly the substring tests the reference does */
static int ends_with(const char *s, const char *suf) {
    size_t a = strlen(s), b = strlen(suf);
    return a >= b && strcmp(s + a - b, suf) == 0;
}

static void parse_reward(tapo_env *e, const char *rt) {
    strncpy(e->reward_type, rt, sizeof(e->reward_type) - 1);
    e->hard = ends_with(rt, "hard");           /* tools.py:2113 */
    e->useP = strchr(rt, 'P') != NULL;         /* tools.py:2135 */
    e->useS = strchr(rt, 'S') != NULL;         /* tools.py:2138 */
    e->mcs_in = strstr(rt, "mcs") != NULL;     /* tools.py:2718 */
    e->mcs_start = strncmp(rt, "mcs", 3) == 0; /* tools.py:2709 */
}

/* ------------------------------------------------------------------ */
/* tools.py:3611-3661 Container.__init__  /  :3858-3885 clear_container */
void tapo_env_clear(tapo_env *e) {
    int cells = e->dim == 2 ? e->W : e->W * e->L;
    memset(e->container, 0, sizeof(int) * (size_t)cells * e->H);
    memset(e->heightmap, 0, sizeof(int) * cells);
    memset(e->positions, 0, sizeof(int) * e->n * e->dim);
    memset(e->blocks, 0, sizeof(int) * e->n * e->dim);
    memset(e->stable, 0, e->n);
    e->valid = 0; e->empty = 0; e->k = 0; e->error = 0;
    if (e->lfs) {
        for (int z = 0; z < e->H; z++) { /* [0, W-1] per level */
            e->lfs[z].len = 2; e->lfs[z].v[0] = 0; e->lfs[z].v[1] = e->W - 1;
        }
    }
    if (e->lfs3) {
        for (int i = 0; i < e->H * e->L; i++) { e->lfs3[i].len = 2; e->lfs3[i].v[0] = 0; e->lfs3[i].v[1] = e->W - 1; }
    }
    if (e->lbl) {
        int cnt = e->dim == 2 ? e->H : e->H * e->L;
        for (int i = 0; i < cnt; i++) { e->lbl[i].len = 1; e->lbl[i].v[0] = 0; }
    }
}

tapo_env *tapo_env_new(int dim, int W, int L, int H, int n, const char *reward_type,
                       int hm_type, int strategy) {
    if (W > TAPO_MAXW || L > TAPO_MAXW || n > TAPO_MAXN || (dim != 2 && dim != 3)) return NULL;
    tapo_env *e = (tapo_env *)calloc(1, sizeof(tapo_env));
    e->dim = dim; e->W = W; e->L = dim == 3 ? L : 1; e->H = H; e->n = n;
    e->hm_type = hm_type;
    parse_reward(e, reward_type);
    /* tools.py:3617-3620: the reward type overrides the packing strategy */
    if (!strcmp(reward_type, "C+P+S-mul-soft") || !strcmp(reward_type, "C+P+S-mul-hard") ||
        !strcmp(reward_type, "C+P+S-mcs-soft") || !strcmp(reward_type, "C+P+S-mcs-hard"))
        strategy = TAPO_MACS;
    e->strategy = strategy;
    int cells = e->W * e->L;
    e->container = (int *)malloc(sizeof(int) * (size_t)cells * H);
    e->heightmap = (int *)malloc(sizeof(int) * cells);
    e->positions = (int *)malloc(sizeof(int) * n * dim);
    e->blocks = (int *)malloc(sizeof(int) * n * dim);
    e->stable = (unsigned char *)malloc(n);
    e->lfs = NULL;
    if (strategy == TAPO_MACS && dim == 2) e->lfs = (ilist *)malloc(sizeof(ilist) * H);
    e->lfs3 = NULL;
    if (strategy == TAPO_MACS && dim == 3) e->lfs3 = (ilist *)malloc(sizeof(ilist) * (size_t)H * e->L);
    e->lbl = NULL;
    if (strategy == TAPO_LB) e->lbl = (xlist *)malloc(sizeof(xlist) * (size_t)H * (dim == 2 ? 1 : e->L));
    tapo_env_clear(e);
    return e;
}

void tapo_env_free(tapo_env *e) {
    if (!e) return;
    free(e->container); free(e->heightmap); free(e->positions); free(e->blocks);
    free(e->stable); free(e->lfs); free(e->lfs3); free(e->lbl); free(e);
}

/* ------------------------------------------------------------------ */
/* tools.py:839-868 is_stable_2d(support, obj_left, obj_width) */
static int is_stable_2d(const int *support, int stride, int slen, int obj_left, int obj_width) {
    /* object_center = obj_left + obj_width/2 ; compare in doubled integers.
     * slen = len(support) (numpy clips the slice at the container wall) */
    int left_index = obj_left, right_index = obj_left + obj_width;
    for (int i = 0; i < slen; i++) { if (support[i * stride] == 0) left_index++; else break; }
    for (int i = slen - 1; i >= 0; i--) { if (support[i * stride] == 0) right_index--; else break; }
    int c2 = 2 * obj_left + obj_width;
    if (c2 <= 2 * left_index || c2 >= 2 * right_index) return 0;
    return 1;
}

/* tools.py:736-744 and :754-762: two-point rule.  a..d are DOUBLED (exact) */
static int two_point_rule(int a, int b, int c, int d) {
    if (b == 0 || d == 0) {
        if (b != d) return 0;
        return (a < 0) != (c < 0);
    }
    /* a/b == c/d in IEEE double: a,b,c,d are small integers (|.| <= 64); the
     * quotients of the doubled values equal the quotients of the originals.
     * Use the same double division the reference performs. */
    double q1 = (double)a / (double)b, q2 = (double)c / (double)d;
    return q1 == q2 && ((a < 0) != (c < 0)) && ((b < 0) != (d < 0));
}

static long long cross_ll(int ox, int oy, int ax, int ay, int bx, int by) {
    return (long long)(ax - ox) * (by - oy) - (long long)(ay - oy) * (bx - ox);
}

/* tools.py:710-765 is_stable(block, position, container) */
static int is_stable_3d(const tapo_env *e, int bx, int by, int px, int py, int pz) {
    if (pz == 0) return 1;
    int x1 = px, x2 = px + bx - 1, y1 = py, y2 = py + by - 1, z = pz - 1;
    int cx2 = x1 + x2, cy2 = y1 + y2; /* doubled centre */
    int pts[TAPO_MAXW * TAPO_MAXW][2]; int np_ = 0;
    for (int x = x1; x <= x2; x++)
        for (int y = y1; y <= y2; y++)
            if (e->container[(x * e->L + y) * e->H + z] > 0) { pts[np_][0] = x; pts[np_][1] = y; np_++; }
    if (2 * np_ > bx * by) return 1;            /* len(points) > bx*by/2  (:730) */
    if (np_ == 0 || np_ == 1) return 0;          /* :732 */
    if (np_ == 2) {
        return two_point_rule(cx2 - 2 * pts[0][0], cy2 - 2 * pts[0][1],
                              cx2 - 2 * pts[1][0], cy2 - 2 * pts[1][1]);
    }
    /* >= 3 points: ConvexHull (:749).  Qhull raises on degenerate (all collinear)
     * input -> fall back to the 2-point rule on argmin/argmax BY X ONLY (:752-762). */
    int collinear = 1;
    for (int i = 2; i < np_ && collinear; i++)
        if (cross_ll(pts[0][0], pts[0][1], pts[1][0], pts[1][1], pts[i][0], pts[i][1]) != 0) collinear = 0;
    if (collinear) {
        int imin = 0, imax = 0; /* np.argmin / np.argmax: first occurrence */
        for (int i = 1; i < np_; i++) {
            if (pts[i][0] < pts[imin][0]) imin = i;
            if (pts[i][0] > pts[imax][0]) imax = i;
        }
        return two_point_rule(cx2 - 2 * pts[imin][0], cy2 - 2 * pts[imin][1],
                              cx2 - 2 * pts[imax][0], cy2 - 2 * pts[imax][1]);
    }
    /* Convex hull, counter-clockwise, extreme points only (what Qhull's 2-d
     * `vertices` holds; checked exhaustively against scipy in
     * tests/test_oracle_vs_reference.py).  Andrew monotone chain; pts are
     * already sorted x-major, y-minor by construction. */
    int hull[2 * TAPO_MAXW * TAPO_MAXW][2]; int hn = 0;
    for (int i = 0; i < np_; i++) {
        while (hn >= 2 && cross_ll(hull[hn-2][0], hull[hn-2][1], hull[hn-1][0], hull[hn-1][1], pts[i][0], pts[i][1]) <= 0) hn--;
        hull[hn][0] = pts[i][0]; hull[hn][1] = pts[i][1]; hn++;
    }
    int lower = hn + 1;
    for (int i = np_ - 2; i >= 0; i--) {
        while (hn >= lower && cross_ll(hull[hn-2][0], hull[hn-2][1], hull[hn-1][0], hull[hn-1][1], pts[i][0], pts[i][1]) <= 0) hn--;
        hull[hn][0] = pts[i][0]; hull[hn][1] = pts[i][1]; hn++;
    }
    hn--; /* last == first */
    /* matplotlib.path.Path(...).contains_point(centre) (:764-765): crossing test
     * of src/_path.h point_in_path_impl, polygon implicitly closed, radius 0.
     * UNPINNED (matplotlib absent).  Doubled-integer arithmetic, exact. */
    int tx = cx2, ty = cy2, inside = 0;
    int vtx0 = 2 * hull[0][0], vty0 = 2 * hull[0][1];
    int yflag0 = vty0 >= ty;
    for (int kk = 1; kk <= hn; kk++) {
        int vtx1 = 2 * hull[kk % hn][0], vty1 = 2 * hull[kk % hn][1];
        int yflag1 = vty1 >= ty;
        if (yflag0 != yflag1) {
            long long lhs = (long long)(vty1 - ty) * (vtx0 - vtx1);
            long long rhs = (long long)(vtx1 - tx) * (vty0 - vty1);
            if ((lhs >= rhs) == yflag1) inside ^= 1;
        }
        yflag0 = yflag1; vtx0 = vtx1; vty0 = vty1;
    }
    return inside;
}

/* exported for the exhaustive hull test: footprint occupancy bitmask -> stable */
int tapo_is_stable_3d_mask(int bx, int by, const unsigned char *occ /*[bx][by]*/) {
    tapo_env e; memset(&e, 0, sizeof(e));
    e.dim = 3; e.W = bx; e.L = by; e.H = 2;
    int cont[TAPO_MAXW * TAPO_MAXW * 2];
    for (int x = 0; x < bx; x++) for (int y = 0; y < by; y++) {
        cont[(x * by + y) * 2 + 0] = occ[x * by + y] ? 1 : 0;
        cont[(x * by + y) * 2 + 1] = 0;
    }
    e.container = cont;
    return is_stable_3d(&e, bx, by, 0, 0, 1);
}

/* bulk form for the exhaustive tests of the CUDA-side restatement: out[i] = is_stable for support
 * bitmask masks[i] (bit x*by+y, x-major) */
void tapo_is_stable_3d_masks(int bx, int by, const unsigned *masks, int count, unsigned char *out) {
    unsigned char occ[TAPO_MAXW * TAPO_MAXW];
    for (int i = 0; i < count; i++) {
        for (int c = 0; c < bx * by; c++) occ[c] = (masks[i] >> c) & 1u;
        out[i] = (unsigned char)tapo_is_stable_3d_mask(bx, by, occ);
    }
}

/* ------------------------------------------------------------------ */
/* stable (insertion) sort of EMS records by one key -- Python list.sort is stable */
static void stable_sort_by(int (*rec)[4], int cnt, int key) {
    for (int i = 1; i < cnt; i++) {
        int t[4]; memcpy(t, rec[i], sizeof(t));
        int j = i - 1;
        while (j >= 0 && rec[j][key] > t[key]) { memcpy(rec[j + 1], rec[j], sizeof(t)); j--; }
        memcpy(rec[j + 1], t, sizeof(t));
    }
}

#define MAX_EMS (2 * TAPO_MAXW * TAPO_MAXW + 2 * TAPO_MAXN + 8)
#define E2D (TAPO_MAXW + 2)

/* np.argmax over doubles: first maximum; NaN handling as numpy (first NaN wins) */
static int argmax_first(const double *v, int cnt) {
    int best = 0;
    for (int i = 1; i < cnt; i++) {
        if (isnan(v[best])) break;
        if (v[i] > v[best] || isnan(v[i])) best = i;
    }
    return best;
}

/* tools.py:2027-2176 calc_one_position_lb_greedy_2d */
static void lbg_step_2d(tapo_env *e, int bx, int bz) {
    const int W = e->W, H = e->H, k = e->k;
    int *h = e->heightmap, *c = e->container;
    long long valid = e->valid + (long long)bx * bz;      /* :2061 */
    int ems[E2D][4]; int ne = 0;
    /* :2067-2073 hm_diff, ems_x_list = [0] + nonzero(hm_diff) */
    int xs[TAPO_MAXW + 1]; int nx = 0; xs[nx++] = 0;
    for (int x = 1; x < W; x++) if (h[x] - h[x - 1] != 0) xs[nx++] = x;
    for (int i = 0; i < nx; i++) {                         /* :2075-2078 */
        int x = xs[i];
        if (x + bx > W) break;
        int z = h[x]; for (int q = x; q < x + bx; q++) if (h[q] > z) z = h[q];
        ems[ne][0] = x; ems[ne][1] = z; ne++;
    }
    stable_sort_by(ems, ne, 1);                            /* :2080-2081 */
    if (ne == 0) { e->stable[k] = 0; return; }             /* :2084-2087 */
    int pos[E2D][2]; unsigned char settle[E2D], stab[E2D];
    double comp[E2D], pyr[E2D], stb[E2D]; long long empty_ems[E2D];
    static __thread int hm_ems[TAPO_MAXW + 2][TAPO_MAXW];  /* 2D: ne <= W */
    int visited[TAPO_MAXW * TAPO_MAXW + 4][2]; int nv = 0;
    int nsettled = 0;
    const int X = W - bx + 1;                              /* :2143 */
    for (int i = 0; i < ne; i++) {
        settle[i] = 0; stab[i] = 0; comp[i] = pyr[i] = stb[i] = 0.0; empty_ems[i] = e->empty;
        pos[i][0] = pos[i][1] = 0;
        memcpy(hm_ems[i], h, sizeof(int) * W);             /* :2146 */
        int _z = ems[i][1];
        for (int _x = ems[i][0]; _x < X; _x++) {           /* :2148-2150 */
            if (settle[i]) break;
            /* check_position :2103-2121 */
            int seen = 0;
            for (int v = 0; v < nv; v++) if (visited[v][0] == _x && visited[v][1] == _z) { seen = 1; break; }
            if (seen) continue;
            if (_z > 0) {
                if (_z - 1 >= H) { e->error = 1; return; }
                int allz = 1;
                for (int q = _x; q < _x + bx; q++) if (c[q * H + _z - 1] != 0) { allz = 0; break; }
                if (allz) continue;
            }
            visited[nv][0] = _x; visited[nv][1] = _z; nv++;
            if (_z >= H) { e->error = 1; return; }        /* numpy IndexError */
            int freerow = 1;
            for (int q = _x; q < _x + bx; q++) if (c[q * H + _z] != 0) { freerow = 0; break; }
            if (!freerow) continue;
            if (_z > 0) {
                if (!is_stable_2d(&c[_x * H + _z - 1], H, bx, _x, bx)) { if (e->hard) continue; }
                else stab[i] = 1;
            } else stab[i] = 1;
            pos[i][0] = _x; pos[i][1] = _z;
            for (int q = _x; q < _x + bx; q++) hm_ems[i][q] = _z + bz;
            settle[i] = 1;
        }
        if (settle[i]) {                                   /* calc_C_P_S :2124-2140 */
            nsettled++;
            int _x = pos[i][0]; _z = pos[i][1];
            int height = hm_ems[i][0]; for (int q = 1; q < W; q++) if (hm_ems[i][q] > height) height = hm_ems[i][q];
            long long bbox = (long long)height * W;
            comp[i] = (double)valid / (double)bbox;
            long long cnt = 0;
            int zlim = _z < H ? _z : H;
            for (int q = _x; q < _x + bx; q++) for (int zz = 0; zz < zlim; zz++) if (c[q * H + zz] == 0) cnt++;
            empty_ems[i] += cnt;
            if (e->useP) pyr[i] = (double)valid / (double)(empty_ems[i] + valid);
            if (e->useS) {
                int sn = 0; for (int q = 0; q < k; q++) sn += e->stable[q];
                sn += stab[i];
                stb[i] = (double)sn / (double)(k + 1);
            }
        }
    }
    if (nsettled == 0) { e->stable[k] = 0; return; }       /* :2155-2158 */
    double ratio[E2D];
    for (int i = 0; i < ne; i++) ratio[i] = (comp[i] + pyr[i]) + stb[i];   /* :2161 */
    int best = argmax_first(ratio, ne);                    /* :2162 (while-loop :2163 is dead) */
    int _x = pos[best][0], _z = pos[best][1];
    for (int q = _x; q < _x + bx; q++) {                   /* :2168-2169 */
        for (int zz = _z; zz < _z + bz && zz < H; zz++) c[q * H + zz] = k + 1;
        for (int zz = 0; zz < _z && zz < H; zz++) if (c[q * H + zz] == 0) c[q * H + zz] = -1;
    }
    e->positions[k * 2 + 0] = _x; e->positions[k * 2 + 1] = _z;
    e->stable[k] = stab[best];
    memcpy(h, hm_ems[best], sizeof(int) * W);
    e->empty = empty_ems[best];
    e->valid = valid;
}

/* tools.py:2178-2351 calc_one_position_lb_greedy_3d */
static void lbg_step_3d(tapo_env *e, int bx, int by, int bz) {
    const int W = e->W, L = e->L, H = e->H, k = e->k;
    int *h = e->heightmap, *c = e->container;
#define HM(x, y) h[(x) * L + (y)]
#define CT(x, y, z) c[((x) * L + (y)) * H + (z)]
    long long valid = e->valid + (long long)bx * by * bz;  /* :2212 */
    /* :2219-2225 */
    static __thread int dx[TAPO_MAXW * TAPO_MAXW], dy[TAPO_MAXW * TAPO_MAXW];
    for (int x = 0; x < W; x++) for (int y = 0; y < L; y++) {
        dx[x * L + y] = x == 0 ? 0 : HM(x, y) - HM(x - 1, y);
        dy[x * L + y] = y == 0 ? 0 : HM(x, y) - HM(x, y - 1);
    }
    /* :2228-2246 */
    static __thread int xy[MAX_EMS][4]; int nxy = 0;
    xy[nxy][0] = 0; xy[nxy][1] = 0; nxy++;
    for (int x = 0; x < W; x++) for (int y = 0; y < L; y++) {   /* ems_x_list, row-major */
        if (dx[x * L + y] == 0) continue;
        if (y != 0 && dx[x * L + y - 1] != 0) {
            if (HM(x, y) == HM(x, y - 1) && dx[x * L + y] == dx[x * L + y - 1]) continue;
        }
        xy[nxy][0] = x; xy[nxy][1] = y; nxy++;
    }
    for (int x = 0; x < W; x++) for (int y = 0; y < L; y++) {   /* ems_y_list */
        if (dy[x * L + y] == 0) continue;
        if (x != 0 && dy[(x - 1) * L + y] != 0) {
            if (HM(x, y) == HM(x - 1, y) && dx[x * L + y] == dx[(x - 1) * L + y]) continue; /* sic: hm_diff_x (:2243) */
        }
        int dup = 0;
        for (int i = 0; i < nxy; i++) if (xy[i][0] == x && xy[i][1] == y) { dup = 1; break; }
        if (!dup) { xy[nxy][0] = x; xy[nxy][1] = y; nxy++; }
    }
    stable_sort_by(xy, nxy, 1);                            /* :2249-2250 */
    static __thread int ems[MAX_EMS][4]; int ne = 0;
    for (int i = 0; i < nxy; i++) {                        /* :2253-2258 */
        int x = xy[i][0], y = xy[i][1];
        if (x + bx > W || y + by > L) continue;
        int z = HM(x, y);
        for (int p = x; p < x + bx; p++) for (int q = y; q < y + by; q++) if (HM(p, q) > z) z = HM(p, q);
        ems[ne][0] = x; ems[ne][1] = y; ems[ne][2] = z; ne++;
    }
    stable_sort_by(ems, ne, 2);                            /* :2261-2262 */
    if (ne == 0) { e->stable[k] = 0; return; }             /* :2265-2268 */
    static __thread int pos[MAX_EMS][3]; static __thread unsigned char settle[MAX_EMS], stab[MAX_EMS];
    static __thread double comp[MAX_EMS], pyr[MAX_EMS], stb[MAX_EMS]; static __thread long long empty_ems[MAX_EMS];
    static __thread int visited[MAX_EMS * 8][3]; int nv = 0;
    int nsettled = 0;
    const int X = W - bx + 1, Y = L - by + 1;
    for (int i = 0; i < ne; i++) {
        settle[i] = 0; stab[i] = 0; comp[i] = pyr[i] = stb[i] = 0.0; empty_ems[i] = e->empty;
        int X0 = ems[i][0], Y0 = ems[i][1], _z = ems[i][2];
        for (int _x = X0; _x < X && !settle[i]; _x++) for (int _y = Y0; _y < Y; _y++) {   /* :2324 */
            if (settle[i]) break;
            /* check_position :2284-2297 */
            int seen = 0;
            for (int v = 0; v < nv; v++) if (visited[v][0] == _x && visited[v][1] == _y && visited[v][2] == _z) { seen = 1; break; }
            if (seen) continue;
            if (_z > 0) {
                if (_z - 1 >= H) { e->error = 1; return; }
                int allz = 1;
                for (int p = _x; p < _x + bx && allz; p++) for (int q = _y; q < _y + by; q++) if (CT(p, q, _z - 1) != 0) { allz = 0; break; }
                if (allz) continue;
            }
            if (nv >= MAX_EMS * 8) { e->error = 2; return; }
            visited[nv][0] = _x; visited[nv][1] = _y; visited[nv][2] = _z; nv++;
            if (_z >= H) { e->error = 1; return; }
            int freerow = 1;
            for (int p = _x; p < _x + bx && freerow; p++) for (int q = _y; q < _y + by; q++) if (CT(p, q, _z) != 0) { freerow = 0; break; }
            if (!freerow) continue;
            if (!is_stable_3d(e, bx, by, _x, _y, _z)) { if (e->hard) continue; }
            else stab[i] = 1;
            pos[i][0] = _x; pos[i][1] = _y; pos[i][2] = _z;
            settle[i] = 1;
        }
        if (settle[i]) {                                   /* calc_C_P_S :2300-2316 */
            nsettled++;
            int _x = pos[i][0], _y = pos[i][1]; _z = pos[i][2];
            int height = 0;
            for (int p = 0; p < W; p++) for (int q = 0; q < L; q++) {
                int hv = (p >= _x && p < _x + bx && q >= _y && q < _y + by) ? _z + bz : HM(p, q);
                if (hv > height) height = hv;
            }
            long long bbox = (long long)height * W * L;
            comp[i] = (double)valid / (double)bbox;
            long long cnt = 0; int zlim = _z < H ? _z : H;
            for (int p = _x; p < _x + bx; p++) for (int q = _y; q < _y + by; q++) for (int zz = 0; zz < zlim; zz++) if (CT(p, q, zz) == 0) cnt++;
            empty_ems[i] += cnt;
            if (e->useP) pyr[i] = (double)valid / (double)(empty_ems[i] + valid);
            if (e->useS) {
                int sn = 0; for (int q = 0; q < k; q++) sn += e->stable[q];
                sn += stab[i];
                stb[i] = (double)sn / (double)(k + 1);
            }
        }
    }
    if (nsettled == 0) { e->stable[k] = 0; return; }       /* :2330-2333 */
    static __thread double ratio[MAX_EMS];
    for (int i = 0; i < ne; i++) ratio[i] = (comp[i] + pyr[i]) + stb[i];
    int best = argmax_first(ratio, ne);
    int _x = pos[best][0], _y = pos[best][1], _z = pos[best][2];
    for (int p = _x; p < _x + bx; p++) for (int q = _y; q < _y + by; q++) {   /* :2342-2343 */
        for (int zz = _z; zz < _z + bz && zz < H; zz++) CT(p, q, zz) = k + 1;
        for (int zz = 0; zz < _z && zz < H; zz++) if (CT(p, q, zz) == 0) CT(p, q, zz) = -1;
        HM(p, q) = _z + bz;
    }
    e->positions[k * 3 + 0] = _x; e->positions[k * 3 + 1] = _y; e->positions[k * 3 + 2] = _z;
    e->stable[k] = stab[best];
    e->empty = empty_ems[best];
    e->valid = valid;
#undef HM
#undef CT
}

/* ------------------------------------------------------------------ */
/* Python-list helpers for MACS level_free_space */
static int il_index(const ilist *l, int v) { for (int i = 0; i < l->len; i++) if (l->v[i] == v) return i; return -1; }
static void il_remove(ilist *l, int v) { int i = il_index(l, v); if (i < 0) return; for (; i + 1 < l->len; i++) l->v[i] = l->v[i + 1]; l->len--; }
static void il_sort(ilist *l) { for (int i = 1; i < l->len; i++) { int t = l->v[i], j = i - 1; while (j >= 0 && l->v[j] > t) { l->v[j + 1] = l->v[j]; j--; } l->v[j + 1] = t; } }
static int il_eq(const ilist *a, const ilist *b) { if (a->len != b->len) return 0; for (int i = 0; i < a->len; i++) if (a->v[i] != b->v[i]) return 0; return 1; }

/* tools.py:2610-2660 update_level_free_space(pos) -> into `out` (deep copy) */
static void macs_update_lfs(const tapo_env *e, ilist *out, int _x, int _z, int bx, int bz, int *err) {
    const int H = e->H;
    memcpy(out, e->lfs, sizeof(ilist) * H);
    int xx = _x + bx - 1;
    for (int zz = _z; zz < _z + bz; zz++) {
        if (zz >= H) { *err = 1; return; }                 /* list IndexError */
        ilist *fs = &out[zz];
        int idx = il_index(fs, _x);
        if (idx >= 0) {
            if ((idx + 1) % 2 == 1) {
                if (il_index(fs, xx) >= 0) {
                    if (bx == 1) {
                        if (fs->v[idx + 1] == _x) { il_remove(fs, _x); il_remove(fs, _x); }
                        else fs->v[idx] = _x + 1;
                    } else { il_remove(fs, _x); il_remove(fs, xx); }
                } else fs->v[idx] = xx + 1;
            } else fs->v[idx] = _x - 1;
        } else {
            int ix = il_index(fs, xx);
            if (ix >= 0) fs->v[ix] = _x - 1;
            else {
                if (fs->len + 2 > TAPO_MAXW + 4) { *err = 2; return; }
                fs->v[fs->len++] = _x - 1; fs->v[fs->len++] = xx + 1; il_sort(fs);
            }
        }
    }
    for (int zz = 0; zz < _z && zz < H; zz++) {
        ilist *fs = &out[zz];
        ilist snap = *fs;                                  /* `spaces` is built before editing */
        for (int s = 0; s + 1 < snap.len; s += 2) {
            int x1 = snap.v[s], x2 = snap.v[s + 1];
            if (x1 == x2) {
                if (x1 >= _x && x1 <= xx) { il_remove(fs, x1); il_remove(fs, x1); }
            } else if (bx == 1) {
                if (_x == x1) { int i = il_index(fs, x1); if (i >= 0) fs->v[i] = _x + 1; }
                else if (_x == x2) { int i = il_index(fs, x2); if (i >= 0) fs->v[i] = xx - 1; }
            } else if (_x <= x1 && x2 <= xx) { il_remove(fs, x1); il_remove(fs, x2); }
            else if (_x <= x1 && x1 <= xx) { int i = il_index(fs, x1); if (i >= 0) fs->v[i] = xx + 1; }
            else if (_x <= x2 && x2 <= xx) { int i = il_index(fs, x2); if (i >= 0) fs->v[i] = _x - 1; }
        }
    }
}

/* tools.py:2667-2678 calc_maximal_usable_spaces(lfs, H) */
static long long macs_usable(const ilist *lfs, int Hlim) {
    long long score = 0;
    for (int hh = 0; hh < Hlim; hh++) {
        int best = 0;
        for (int s = 0; s + 1 < lfs[hh].len; s += 2) { int len = lfs[hh].v[s + 1] - lfs[hh].v[s]; if (len > best) best = len; }
        score += best;
    }
    return score;
}

/* tools.py:2456-2749 calc_one_position_mcs_2d */
static void macs_step_2d(tapo_env *e) {
    const int W = e->W, H = e->H, k = e->k;
    int *h = e->heightmap, *c = e->container;
    const int bx = e->blocks[k * 2 + 0], bz = e->blocks[k * 2 + 1];
    long long valid = e->valid + (long long)bx * bz;       /* :2513 */
    static __thread int ems[MAX_EMS][4]; int ne = 0;
    /* list A :2518-2529 */
    for (int z = 0; z < H; z++) {
        const ilist *fs = &e->lfs[z];
        if (z + bz > H) break;
        else if (z > 0 && il_eq(&e->lfs[z - 1], fs)) continue;
        for (int s = 0; s + 1 < fs->len; s += 2) {
            int x1 = fs->v[s], x2 = fs->v[s + 1];
            if (x1 + bx > W) break;
            if (z > 0) {
                int idx = il_index(&e->lfs[z - 1], x1);
                if (idx >= 0) { idx += 1; if (idx % 2 == 1 && idx < e->lfs[z - 1].len && x2 == e->lfs[z - 1].v[idx]) continue; }
            }
            if (ne >= MAX_EMS) { e->error = 2; return; }
            ems[ne][0] = x1; ems[ne][1] = z; ems[ne][2] = x2; ems[ne][3] = z; ne++;
        }
    }
    /* list B :2531-2555 */
    for (int b = 0; b < k; b++) {
        int x = e->positions[b * 2], z = e->positions[b * 2 + 1];
        int xx = e->blocks[b * 2], zz = e->blocks[b * 2 + 1];
        int t = z + zz;
        if (t < H) {
            int full = 1;
            for (int q = x; q < x + xx && q < W; q++) if (c[q * H + t] != 0) { full = 0; break; }
            if (full) {
                int dup = 0;
                for (int i = 0; i < ne; i++) if (ems[i][0] == x && ems[i][1] == t && ems[i][2] == x + xx - 1 && ems[i][3] == t) { dup = 1; break; }
                if (!dup) { if (ne >= MAX_EMS) { e->error = 2; return; } ems[ne][0] = x; ems[ne][1] = t; ems[ne][2] = x + xx - 1; ems[ne][3] = t; ne++; }
            } else {
                if (x + xx > W) { e->error = 1; return; }
                if (c[x * H + t] == 0) {                   /* left */
                    if (x > 0 && c[(x - 1) * H + t] == 0) {
                        int x2 = x;
                        for (x2 = x; x2 < x + xx; x2++) {
                            if (x2 == W - 1) break;
                            if (c[(x2 + 1) * H + t] != 0) break;
                        }
                        if (x2 == x + xx) x2 = x + xx - 1;  /* loop ran to completion */
                        if (ne >= MAX_EMS) { e->error = 2; return; }
                        ems[ne][0] = x; ems[ne][1] = t; ems[ne][2] = x2; ems[ne][3] = t; ne++;
                    }
                }
                if (c[(x + xx - 1) * H + t] == 0) {        /* right */
                    if (x + xx < W && c[(x + xx) * H + t] == 0) {
                        int x1 = x + xx - 1;
                        for (x1 = x + xx - 1; x1 >= x; x1--) {
                            if (x1 == 0) break;
                            if (c[(x1 - 1) * H + t] != 0) break;
                        }
                        if (x1 < x) x1 = x;                 /* loop ran to completion */
                        if (ne >= MAX_EMS) { e->error = 2; return; }
                        ems[ne][0] = x1; ems[ne][1] = t; ems[ne][2] = x + xx - 1; ems[ne][3] = t; ne++;
                    }
                }
            }
        }
    }
    const int nc = ne * 2;
    static __thread int pos[2 * MAX_EMS][2]; static __thread unsigned char settle[2 * MAX_EMS], stab[2 * MAX_EMS];
    static __thread double comp[2 * MAX_EMS], pyr[2 * MAX_EMS], stb[2 * MAX_EMS]; static __thread long long empty_ems[2 * MAX_EMS];
    static __thread int hmmax[2 * MAX_EMS];   /* max of heightmap_ems[index] (0 if never touched) */
    static __thread int visited[65536][2]; int nv = 0;
    for (int i = 0; i < nc; i++) { settle[i] = stab[i] = 0; comp[i] = pyr[i] = stb[i] = 0.0; empty_ems[i] = e->empty; pos[i][0] = pos[i][1] = 0; hmmax[i] = 0; }
    int hmax0 = 0; for (int q = 0; q < W; q++) if (h[q] > hmax0) hmax0 = h[q];
    const int X = W - bx + 1;
    int nsettled = 0;
    for (int ei = 0; ei < ne; ei++) {
        int X1 = ems[ei][0], Z = ems[ei][1], X2 = ems[ei][2];
        for (int side = 0; side < 2; side++) {
            int index = ei * 2 + side;
            int lo, hi, step;
            if (side == 0) { if (!(X1 < X)) continue; lo = X1; hi = X; step = 1; }
            else { if (!(X2 - bx + 2 > 0)) continue; lo = X2 - bx + 1; hi = -1; step = -1; }
            hmmax[index] = hmax0;                          /* heightmap.copy() */
            for (int _x = lo; _x != hi; _x += step) {
                if (settle[index]) break;
                /* check_position :2569-2588 */
                int seen = 0;
                for (int v = 0; v < nv; v++) if (visited[v][0] == _x && visited[v][1] == Z) { seen = 1; break; }
                if (seen) continue;
                const int xe = _x + bx < W ? _x + bx : W;     /* numpy clips slices at the wall */
                if (Z > 0) {
                    if (Z - 1 >= H) { e->error = 1; return; }
                    int allz = 1;
                    for (int q = _x; q < xe; q++) if (c[q * H + Z - 1] != 0) { allz = 0; break; }
                    if (allz) continue;
                }
                if (nv >= (int)(sizeof(visited) / sizeof(visited[0]))) { e->error = 2; return; }
                visited[nv][0] = _x; visited[nv][1] = Z; nv++;
                int freeall = 1;
                for (int q = _x; q < xe && freeall; q++) for (int zz = Z; zz < Z + bz && zz < H; zz++) if (c[q * H + zz] != 0) { freeall = 0; break; }
                if (!freeall) continue;
                if (Z > 0) {
                    if (!is_stable_2d(&c[_x * H + Z - 1], H, xe > _x ? xe - _x : 0, _x, bx)) { if (e->hard) continue; }
                    else stab[index] = 1;
                } else stab[index] = 1;
                pos[index][0] = _x; pos[index][1] = Z; settle[index] = 1;
            }
            if (settle[index]) {                           /* calc_C_P_S :2590-2604 */
                nsettled++;
                int _x = pos[index][0], _z = pos[index][1];
                int height = 0;
                for (int q = 0; q < W; q++) { int hv = (q >= _x && q < _x + bx) ? _z + bz : h[q]; if (hv > height) height = hv; }
                hmmax[index] = height;
                const int xe = _x + bx < W ? _x + bx : W;
                if (_z + bx > height) height = _z + bz;    /* sic :2594 */
                long long bbox = (long long)height * W;
                comp[index] = (double)valid / (double)bbox;
                long long cnt = 0; int zlim = _z < H ? _z : H;
                for (int q = _x; q < xe; q++) for (int zz = 0; zz < zlim; zz++) if (c[q * H + zz] == 0) cnt++;
                empty_ems[index] += cnt;
                if (e->useP) pyr[index] = (double)valid / (double)(empty_ems[index] + valid);
                if (e->useS) {
                    int sn = 0; for (int q = 0; q < k; q++) sn += e->stable[q];
                    sn += stab[index];
                    stb[index] = (double)sn / (double)(k + 1);
                }
            }
        }
    }
    if (nsettled == 0) { e->stable[k] = 0; return; }       /* :2703-2706 */
    static __thread double ratio[2 * MAX_EMS];
    for (int i = 0; i < nc; i++) ratio[i] = e->mcs_start ? 0.0 : (comp[i] + pyr[i]) + stb[i];   /* :2709-2712 */
    double best_score = ratio[0]; for (int i = 1; i < nc; i++) if (ratio[i] > best_score) best_score = ratio[i];
    static __thread int cands[2 * MAX_EMS]; int ncand = 0;
    for (int i = 0; i < nc; i++) if (ratio[i] == best_score) cands[ncand++] = i;
    int best_index;
    if (ncand > 1 && e->mcs_in) {                          /* :2718-2731 */
        int max_height = 0; for (int i = 0; i < nc; i++) if (hmmax[i] > max_height) max_height = hmmax[i];
        if (max_height > H) { e->error = 1; return; }
        static __thread long long mus[2 * MAX_EMS];
        ilist *tmp = (ilist *)malloc(sizeof(ilist) * H);
        for (int i = 0; i < ncand; i++) {
            mus[i] = 0;
            if (settle[cands[i]]) {
                int err = 0;
                macs_update_lfs(e, tmp, pos[cands[i]][0], pos[cands[i]][1], bx, bz, &err);
                if (err) { e->error = err; free(tmp); return; }
                mus[i] = macs_usable(tmp, max_height);
            }
        }
        free(tmp);
        int bi = 0; for (int i = 1; i < ncand; i++) if (mus[i] > mus[bi]) bi = i;
        best_index = cands[bi];
        while (!settle[best_index]) {
            mus[bi] = -1;
            bi = 0; for (int i = 1; i < ncand; i++) if (mus[i] > mus[bi]) bi = i;
            best_index = cands[bi];
        }
    } else {                                               /* :2732-2736 */
        int ci = 0; best_index = cands[0];
        while (!settle[best_index]) { ci++; best_index = cands[ci]; }
    }
    int _x = pos[best_index][0], _z = pos[best_index][1];
    e->positions[k * 2] = _x; e->positions[k * 2 + 1] = _z;  /* :2739-2747 */
    e->stable[k] = stab[best_index];
    e->empty = empty_ems[best_index];
    for (int q = _x; q < _x + bx && q < W; q++) {
        for (int zz = _z; zz < _z + bz && zz < H; zz++) c[q * H + zz] = k + 1;
        for (int zz = 0; zz < _z && zz < H; zz++) if (c[q * H + zz] == 0) c[q * H + zz] = -1;
    }
    {
        ilist *tmp = (ilist *)malloc(sizeof(ilist) * H); int err = 0;
        macs_update_lfs(e, tmp, _x, _z, bx, bz, &err);
        if (err) e->error = err; else memcpy(e->lfs, tmp, sizeof(ilist) * H);
        free(tmp);
    }
    for (int q = _x; q < _x + bx && q < W; q++) h[q] = _z + bz;
    e->valid = valid;
}


/* ------------------------------------------------------------------ */
/* tools.py:2751-3165 calc_one_position_mcs_3d -- LITERAL, including
 *   - the read of a loop variable that is stale at that point of the Python source (`x1` at :2869),
 *   - numpy's slice clipping / empty-slice semantics ((empty == 0).all() is True),
 *   - the z (not z+zz) written into the EMS found on top of a partly covered block (:2939-2940),
 *   - the `_z + block_x` height typo (:2976), the shared `visited` list, phantom (0,0,0) positions of unplaced blocks.
 * e->error = 1 where the reference would raise (IndexError / UnboundLocalError). */
#define M3_MAX_EMS 8192
typedef struct { int v[6]; } ems6;
#define C3(x, y, z) c[(((x) * L + (y)) * H) + (z)]

/* (container[xa:xb, y, z] == 0).all() with Python slice semantics on the x axis */
static int m3_row_free(const int *c, int W, int L, int H, int xa, int xb, int y, int z) {
    if (xa < 0) { xa += W; if (xa < 0) xa = 0; }           /* a negative bound counts from the wall (the edited interval lists */
    if (xb < 0) { xb += W; if (xb < 0) xb = 0; }           /* can hold -1: update_level_free_space writes _x - 1, :3016-3020)  */
    if (xb > W) xb = W;
    for (int x = xa; x < xb; x++) if (C3(x, y, z) != 0) return 0;
    return 1;
}
static int m3_has(const ems6 *l, int n, int a, int b, int cc, int d, int e_, int f) {
    for (int i = 0; i < n; i++) if (l[i].v[0] == a && l[i].v[1] == b && l[i].v[2] == cc && l[i].v[3] == d && l[i].v[4] == e_ && l[i].v[5] == f) return 1;
    return 0;
}
#define M3_PUSH(a, b, cc, d, e_, f) do { if (ne >= M3_MAX_EMS) { e->error = 2; return; } ems[ne].v[0] = (a); ems[ne].v[1] = (b); ems[ne].v[2] = (cc); ems[ne].v[3] = (d); ems[ne].v[4] = (e_); ems[ne].v[5] = (f); ne++; } while (0)

/* one row of update_level_free_space, levels _z .. _z+bz-1 (:3000-3024) */
static void m3_lfs_occupy(ilist *fs, int _x, int xx, int bx, int *err) {
    int idx = il_index(fs, _x);
    if (idx >= 0) {
        if ((idx + 1) % 2 == 1) {
            if (il_index(fs, xx) >= 0) {
                if (bx == 1) {
                    if (idx + 1 < fs->len && fs->v[idx + 1] == _x) { il_remove(fs, _x); il_remove(fs, _x); }
                    else fs->v[idx] = _x + 1;
                } else { il_remove(fs, _x); il_remove(fs, xx); }
            } else fs->v[idx] = xx + 1;
        } else fs->v[idx] = _x - 1;
    } else {
        int ix = il_index(fs, xx);
        if (ix >= 0) fs->v[ix] = _x - 1;
        else {
            if (fs->len + 2 > TAPO_MAXW + 4) { *err = 2; return; }
            fs->v[fs->len++] = _x - 1; fs->v[fs->len++] = xx + 1; il_sort(fs);
        }
    }
}
/* one row of the levels below the block, 0 .. _z-1 (:3026-3042) */
static void m3_lfs_under(ilist *fs, int _x, int xx, int bx) {
    ilist snap = *fs;                                      /* `spaces` is built before editing */
    for (int s = 0; s + 1 < snap.len; s += 2) {
        int x1 = snap.v[s], x2 = snap.v[s + 1];
        if (x1 == x2) {
            if (x1 >= _x && x1 <= xx) { il_remove(fs, x1); il_remove(fs, x1); }
        } else if (bx == 1) {
            if (_x == x1) { int i = il_index(fs, x1); if (i >= 0) fs->v[i] = _x + 1; }
            else if (_x == x2) { int i = il_index(fs, x2); if (i >= 0) fs->v[i] = xx - 1; }
        } else if (_x <= x1 && x2 <= xx) { il_remove(fs, x1); il_remove(fs, x2); }
        else if (_x <= x1 && x1 <= xx) { int i = il_index(fs, x1); if (i >= 0) fs->v[i] = xx + 1; }
        else if (_x <= x2 && x2 <= xx) { int i = il_index(fs, x2); if (i >= 0) fs->v[i] = _x - 1; }
    }
}

/* calc_maximal_usable_spaces(ctn, Hlim) :3052-3080: per level the largest histogram rectangle of empty cells among the
 * scanned anchors; `ctn` = container with the candidate block written in (update_container :3046-3050) */
static long long m3_usable(const int *c, int W, int L, int H, int Hlim) {
    static __thread int hist[TAPO_MAXW][TAPO_MAXW];
    long long score = 0;
    for (int hh = 0; hh < Hlim; hh++) {
        int level_max = 0;
        for (int i = W - 1; i >= 0; i--)
            for (int j = 0; j < L; j++) {
                int hot = C3(i, j, hh) == 0;
                if (i == W - 1) hist[i][j] = hot;
                else if (!hot) hist[i][j] = 0;
                else hist[i][j] = hist[i + 1][j] + hot;
            }
        for (int i = 0; i < W; i++)
            for (int j = 0; j < L; j++) {
                if (hist[i][j] == 0) continue;
                if (j > 0 && hist[i][j] == hist[i][j - 1]) continue;
                int j2, j1;
                for (j2 = j; j2 < L; j2++) { if (j2 == L - 1) break; if (hist[i][j2 + 1] < hist[i][j]) break; }
                for (j1 = j; j1 >= 0; j1--) { if (j1 == 0) break; if (hist[i][j1 - 1] < hist[i][j]) break; }
                int area = hist[i][j] * (j2 - j1 + 1);
                if (area > level_max) level_max = area;
            }
        score += level_max;
    }
    return score;
}

static int m3_dbg_max_ne = 0, m3_dbg_max_nv = 0, m3_dbg_max_zlevels = 0;
int tapo_dbg_m3(int which) { return which == 0 ? m3_dbg_max_ne : (which == 1 ? m3_dbg_max_nv : m3_dbg_max_zlevels); }

static void macs_step_3d(tapo_env *e) {
    const int W = e->W, L = e->L, H = e->H, k = e->k;
    int *h = e->heightmap, *c = e->container;
    ilist *lfs = e->lfs3;
    const int bx = e->blocks[k * 3], by = e->blocks[k * 3 + 1], bz = e->blocks[k * 3 + 2];
    long long valid = e->valid + (long long)bx * by * bz;  /* :2806 */
    static __thread ems6 ems[M3_MAX_EMS]; int ne = 0;
    /* Python function-scope loop variables that outlive their loops */
    int x1 = 0, x2 = 0, y1 = 0, y2 = 0, x1_bound = 0;

    /* ---- EMS from level_free_space :2810-2838 ---- */
    for (int z = 0; z < H; z++) {
        const ilist *fsx = &lfs[z * L];
        if (z + bz > H) break;
        else if (z > 0) { int same = 1; for (int y = 0; y < L; y++) if (!il_eq(&lfs[(z - 1) * L + y], &fsx[y])) { same = 0; break; } if (same) continue; }
        for (int y = 0; y < L; y++) {
            const ilist *fs = &fsx[y];
            if (y + by > L) break;
            else if (y > 0 && il_eq(&fsx[y - 1], fs)) continue;
            for (int s = 0; s + 1 < fs->len; s += 2) {
                x1 = fs->v[s]; x2 = fs->v[s + 1]; x1_bound = 1;
                if (x1 + bx > W) break;
                if (y > 0) {
                    int idx = il_index(&fsx[y - 1], x1);
                    if (idx >= 0) { idx += 1; if (idx % 2 == 1 && idx < fsx[y - 1].len && x2 == fsx[y - 1].v[idx]) continue; }
                }
                if (z > 0) {
                    const ilist *lo = &lfs[(z - 1) * L + y];
                    int idx = il_index(lo, x1);
                    if (idx >= 0) { idx += 1; if (idx % 2 == 1 && idx < lo->len && x2 == lo->v[idx]) continue; }
                }
                int xspace = 1;
                for (y2 = y; y2 < L; y2++) {
                    if (y2 == L - 1) break;
                    if (!m3_row_free(c, W, L, H, x1, x2 + 1, y2 + 1, z)) break;
                    if (xspace && !(il_index(&fsx[y2 + 1], x1) >= 0 && il_index(&fsx[y2 + 1], x2) >= 0)) {
                        xspace = 0;                        /* next to settled blocks along the x axis */
                        M3_PUSH(x1, y, z, x2, y2, z);
                    }
                }
                M3_PUSH(x1, y, z, x2, y2, z);
            }
        }
    }

    /* ---- EMS next to the settled blocks :2841-2940 (every previous block, unplaced ones at their phantom (0,0,0)) ---- */
    for (int b = 0; b < k; b++) {
        const int x = e->positions[b * 3], y = e->positions[b * 3 + 1], z = e->positions[b * 3 + 2];
        const int xx = e->blocks[b * 3], yy = e->blocks[b * 3 + 1], zz = e->blocks[b * 3 + 2];
        if (z >= H) { e->error = 1; return; }
        /* upon along the y axis */
        if (y + yy < L) {
            if (m3_row_free(c, W, L, H, x, x + xx, y + yy, z)) {                       /* full */
                if ((x > 0 && C3(x - 1, y + yy, z) == 0) || (x + xx < W && C3(x + xx, y + yy, z) == 0)) {
                    for (y2 = y + yy; y2 < L; y2++) { if (y2 == L - 1) break; if (!m3_row_free(c, W, L, H, x, x + xx, y2 + 1, z)) break; }
                    M3_PUSH(x, y + yy, z, x + xx - 1, y2, z);
                }
            } else {                                                                  /* part */
                if (x + xx > W) { e->error = 1; return; }                             /* container[x+xx-1, ...] below would raise */
                if (C3(x, y + yy, z) == 0) {                                          /* left */
                    if (x > 0 && C3(x - 1, y + yy, z) == 0) {
                        for (x2 = x; x2 < x + xx; x2++) { if (x2 == W - 1) break; if (C3(x2 + 1, y + yy, z) != 0) break; }
                        if (x2 == x + xx) x2 = x + xx - 1;                             /* loop ran to completion */
                        if (!x1_bound) { e->error = 1; return; }                      /* UnboundLocalError */
                        for (y2 = y + yy; y2 < L; y2++) { if (y2 == L - 1) break; if (!m3_row_free(c, W, L, H, x1 /* stale, :2869 */, x2 + 1, y2 + 1, z)) break; }
                        M3_PUSH(x, y + yy, z, x2, y2, z);
                    }
                }
                if (C3(x + xx - 1, y + yy, z) == 0) {                                 /* right */
                    if (x + xx < W && C3(x + xx, y + yy, z) == 0) {
                        for (x1 = x + xx - 1; x1 >= x; x1--) { if (x1 == 0) break; if (C3(x1 - 1, y + yy, z) != 0) break; }
                        if (x1 < x) x1 = x;
                        x1_bound = 1;
                        for (y2 = y + yy; y2 < L; y2++) { if (y2 == L - 1) break; if (!m3_row_free(c, W, L, H, x1, x + xx, y2 + 1, z)) break; }
                        M3_PUSH(x1, y + yy, z, x + xx - 1, y2, z);
                    }
                }
            }
        }
        /* under along the y axis */
        if (y > 0) {
            if (m3_row_free(c, W, L, H, x, x + xx, y - 1, z)) {                        /* full */
                if ((x > 0 && C3(x - 1, y - 1, z) == 0) || (x + xx < W && C3(x + xx, y - 1, z) == 0)) {
                    for (y1 = y - 1; y1 >= 0; y1--) { if (y1 == 0) break; if (!m3_row_free(c, W, L, H, x, x + xx, y1 - 1, z)) break; }
                    M3_PUSH(x, y1, z, x + xx - 1, y - 1, z);
                }
            } else {
                if (x + xx > W) { e->error = 1; return; }
                if (C3(x, y - 1, z) == 0) {                                           /* left */
                    if (x > 0 && C3(x - 1, y - 1, z) == 0) {
                        for (x2 = x; x2 < x + xx; x2++) { if (x2 == W - 1) break; if (C3(x2 + 1, y - 1, z) != 0) break; }
                        if (x2 == x + xx) x2 = x + xx - 1;
                        for (y1 = y - 1; y1 >= 0; y1--) { if (y1 == 0) break; if (!m3_row_free(c, W, L, H, x, x2 + 1, y1 - 1, z)) break; }
                        M3_PUSH(x, y1, z, x2, y - 1, z);
                    }
                }
                if (C3(x + xx - 1, y - 1, z) == 0) {                                  /* right */
                    if (x + xx < W && C3(x + xx, y - 1, z) == 0) {
                        for (x1 = x + xx - 1; x1 >= x; x1--) { if (x1 == 0) break; if (C3(x1 - 1, y - 1, z) != 0) break; }
                        if (x1 < x) x1 = x;
                        x1_bound = 1;
                        for (y1 = y - 1; y1 >= 0; y1--) { if (y1 == 0) break; if (!m3_row_free(c, W, L, H, x1, x + xx, y1 - 1, z)) break; }
                        M3_PUSH(x1, y1, z, x + xx - 1, y - 1, z);
                    }
                }
            }
        }
        /* upon along the z axis (top) */
        if (z + zz < H) {
            const int t = z + zz;
            int full = 1;
            for (int q = x; q < x + xx && q < W && full; q++) for (int r = y; r < y + yy && r < L; r++) if (C3(q, r, t) != 0) { full = 0; break; }
            if (full) {
                if (!m3_has(ems, ne, x, y, t, x + xx - 1, y + yy - 1, t)) M3_PUSH(x, y, t, x + xx - 1, y + yy - 1, t);
            } else {
                if (x + xx > W || y + yy > L) { e->error = 1; return; }               /* hotmap is clipped, histmap[i, j] raises */
                static __thread int hot[TAPO_MAXW][TAPO_MAXW], hist[TAPO_MAXW][TAPO_MAXW];
                for (int i = 0; i < xx; i++) for (int j = 0; j < yy; j++) hot[i][j] = C3(x + i, y + j, t) == 0;
                for (int i = xx - 1; i >= 0; i--)
                    for (int j = 0; j < yy; j++) {
                        if (i == xx - 1) hist[i][j] = hot[i][j];
                        else if (hot[i][j] == 0) hist[i][j] = 0;
                        else hist[i][j] = hist[i + 1][j] + hot[i][j];
                    }
                for (int i = 0; i < xx; i++)
                    for (int j = 0; j < yy; j++) {
                        if (hist[i][j] == 0) continue;
                        if (j > 0 && hist[i][j] == hist[i][j - 1]) continue;
                        if (i > 0) { int eq = 1; for (int r = y + j; r < y + yy; r++) if (C3(x + i, r, t) != C3(x + i - 1, r, t)) { eq = 0; break; } if (eq) continue; }
                        int i2 = i + hist[i][j] - 1, j2, j1;
                        for (j2 = j; j2 < yy; j2++) { if (j2 == yy - 1) break; if (hist[i][j2 + 1] < hist[i][j]) break; }
                        if (i > 0) { int eq = 1; for (int r = y + j; r < y + j2; r++) if (C3(x + i, r, t) != C3(x + i - 1, r, t)) { eq = 0; break; } if (eq) continue; }   /* empty slice when j2 == j */
                        for (j1 = j; j1 >= 0; j1--) { if (j1 == 0) break; if (hist[i][j1 - 1] < hist[i][j]) break; }
                        if (!m3_has(ems, ne, x + i, y + j1, z, x + i2, y + j2, z)) M3_PUSH(x + i, y + j1, z /* sic */, x + i2, y + j2, z);
                    }
            }
        }
    }

    /* ---- candidates: four corners per EMS :2943-3126 ---- */
    const int nc = ne * 4;
    static __thread int pos[4 * M3_MAX_EMS][3]; static __thread unsigned char settle[4 * M3_MAX_EMS], stab[4 * M3_MAX_EMS], tried[4 * M3_MAX_EMS];
    static __thread double comp[4 * M3_MAX_EMS], pyr[4 * M3_MAX_EMS], stb[4 * M3_MAX_EMS]; static __thread long long empty_ems[4 * M3_MAX_EMS];
    static __thread int hmmax[4 * M3_MAX_EMS];
    static __thread int visited[1 << 16][3]; int nv = 0;
    for (int i = 0; i < nc; i++) { settle[i] = stab[i] = tried[i] = 0; comp[i] = pyr[i] = stb[i] = 0.0; empty_ems[i] = e->empty; pos[i][0] = pos[i][1] = pos[i][2] = 0; hmmax[i] = 0; }
    int hmax0 = 0; for (int q = 0; q < W * L; q++) if (h[q] > hmax0) hmax0 = h[q];
    const int X = W - bx + 1, Y = L - by + 1;
    int nsettled = 0;
    for (int ei = 0; ei < ne; ei++) {
        const int X1 = ems[ei].v[0], Y1 = ems[ei].v[1], Z = ems[ei].v[2], X2 = ems[ei].v[3], Y2 = ems[ei].v[4];
        for (int corner = 0; corner < 4; corner++) {
            const int index = ei * 4 + corner;
            const int xr = X2 - bx + 2, yr = Y2 - by + 2;    /* reversed(range(0, xr)) / reversed(range(0, yr)) */
            int ok;
            if (corner == 0) ok = X1 < X && Y1 < Y;
            else if (corner == 1) ok = xr > 0 && Y1 < Y;
            else if (corner == 2) ok = xr > 0 && yr > 0;
            else ok = X1 < X && yr > 0;
            if (!ok) continue;
            tried[index] = 1; hmmax[index] = hmax0;         /* heightmap.copy() */
            /* itertools.product order: corner 0 (x asc, y asc), 1 (y asc, x desc), 2 (x desc, y desc), 3 (y desc, x asc) */
            const int na = (corner == 0) ? (X - X1) : (corner == 1) ? (Y - Y1) : (corner == 2) ? xr : yr;
            const int nb = (corner == 0) ? (Y - Y1) : (corner == 1) ? xr : (corner == 2) ? yr : (X - X1);
            for (int ia = 0; ia < na && !settle[index]; ia++) {
                for (int ib = 0; ib < nb && !settle[index]; ib++) {
                    int _x, _y;
                    if (corner == 0) { _x = X1 + ia; _y = Y1 + ib; }
                    else if (corner == 1) { _y = Y1 + ia; _x = xr - 1 - ib; }
                    else if (corner == 2) { _x = xr - 1 - ia; _y = yr - 1 - ib; }
                    else { _y = yr - 1 - ia; _x = X1 + ib; }
                    /* check_position :2955-2968 */
                    int seen = 0;
                    for (int v = 0; v < nv; v++) if (visited[v][0] == _x && visited[v][1] == _y && visited[v][2] == Z) { seen = 1; break; }
                    if (seen) continue;
                    if (_x < 0 || _y < 0) { e->error = 1; return; }
                    const int xe = _x + bx < W ? _x + bx : W, ye = _y + by < L ? _y + by : L;
                    if (Z > 0) {
                        if (Z - 1 >= H) { e->error = 1; return; }
                        int allz = 1;
                        for (int q = _x; q < xe && allz; q++) for (int r = _y; r < ye; r++) if (C3(q, r, Z - 1) != 0) { allz = 0; break; }
                        if (allz) continue;
                    }
                    if (nv >= (1 << 16)) { e->error = 2; return; }
                    visited[nv][0] = _x; visited[nv][1] = _y; visited[nv][2] = Z; nv++;
                    int freeall = 1;
                    for (int q = _x; q < xe && freeall; q++) for (int r = _y; r < ye && freeall; r++) for (int t = Z; t < Z + bz && t < H; t++) if (C3(q, r, t) != 0) { freeall = 0; break; }
                    if (!freeall) continue;
                    if (xe - _x != bx || ye - _y != by) { e->error = 1; return; }     /* is_stable would index outside the container */
                    if (!is_stable_3d(e, bx, by, _x, _y, Z)) { if (e->hard) continue; }
                    else stab[index] = 1;
                    pos[index][0] = _x; pos[index][1] = _y; pos[index][2] = Z; settle[index] = 1;
                }
            }
            if (settle[index]) {                            /* calc_C_P_S :2971-2987 */
                nsettled++;
                const int _x = pos[index][0], _y = pos[index][1], _z = pos[index][2];
                int height = 0;
                for (int q = 0; q < W; q++) for (int r = 0; r < L; r++) {
                    int hv = (q >= _x && q < _x + bx && r >= _y && r < _y + by) ? _z + bz : h[q * L + r];
                    if (hv > height) height = hv;
                }
                hmmax[index] = height;
                if (_z + bx > height) height = _z + bz;     /* sic :2976 */
                const long long bbox = (long long)height * W * L;
                comp[index] = (double)valid / (double)bbox;
                long long cnt = 0; const int zlim = _z < H ? _z : H;
                for (int q = _x; q < _x + bx; q++) for (int r = _y; r < _y + by; r++) for (int t = 0; t < zlim; t++) if (C3(q, r, t) == 0) cnt++;
                empty_ems[index] += cnt;
                if (e->useP) pyr[index] = (double)valid / (double)(empty_ems[index] + valid);
                if (e->useS) {
                    int sn = 0; for (int q = 0; q < k; q++) sn += e->stable[q];
                    sn += stab[index];
                    stb[index] = (double)sn / (double)(k + 1);
                }
            }
        }
    }
    if (ne > m3_dbg_max_ne) m3_dbg_max_ne = ne;
    if (nv > m3_dbg_max_nv) m3_dbg_max_nv = nv;
    { int zs[256], nz = 0; for (int v = 0; v < nv; v++) { int f = 0; for (int q = 0; q < nz; q++) if (zs[q] == visited[v][2]) f = 1; if (!f && nz < 256) zs[nz++] = visited[v][2]; } if (nz > m3_dbg_max_zlevels) m3_dbg_max_zlevels = nz; }
    if (nsettled == 0) { e->stable[k] = 0; return; }        /* :3129-3132 */

    /* ---- choose :3135-3158 ---- */
    static __thread double ratio[4 * M3_MAX_EMS];
    for (int i = 0; i < nc; i++) ratio[i] = e->mcs_start ? 0.0 : (comp[i] + pyr[i]) + stb[i];
    double best_score = ratio[0]; for (int i = 1; i < nc; i++) if (ratio[i] > best_score) best_score = ratio[i];
    static __thread int cands[4 * M3_MAX_EMS]; int ncand = 0;
    for (int i = 0; i < nc; i++) if (ratio[i] == best_score) cands[ncand++] = i;
    int best_index;
    if (ncand > 1 && e->mcs_in) {
        int max_height = 0; for (int i = 0; i < nc; i++) if (hmmax[i] > max_height) max_height = hmmax[i];   /* np.max(heightmap_ems) */
        if (max_height > H) { e->error = 1; return; }
        static __thread long long mus[4 * M3_MAX_EMS];
        int *tmp = (int *)malloc(sizeof(int) * (size_t)W * L * H);
        for (int i = 0; i < ncand; i++) {
            mus[i] = 0;
            if (settle[cands[i]]) {
                memcpy(tmp, c, sizeof(int) * (size_t)W * L * H);
                const int _x = pos[cands[i]][0], _y = pos[cands[i]][1], _z = pos[cands[i]][2];
                for (int q = _x; q < _x + bx; q++) for (int r = _y; r < _y + by; r++) {         /* update_container :3046-3050 */
                    for (int t = _z; t < _z + bz && t < H; t++) tmp[((q * L + r) * H) + t] = k + 1;
                    for (int t = 0; t < _z && t < H; t++) if (tmp[((q * L + r) * H) + t] == 0) tmp[((q * L + r) * H) + t] = -1;
                }
                mus[i] = m3_usable(tmp, W, L, H, max_height);
            }
        }
        free(tmp);
        int bi = 0; for (int i = 1; i < ncand; i++) if (mus[i] > mus[bi]) bi = i;
        best_index = cands[bi];
        while (!settle[best_index]) {
            mus[bi] = -1;
            bi = 0; for (int i = 1; i < ncand; i++) if (mus[i] > mus[bi]) bi = i;
            best_index = cands[bi];
        }
    } else {
        int ci = 0; best_index = cands[0];
        while (!settle[best_index]) { ci++; best_index = cands[ci]; }
    }

    /* ---- commit :3161-3171 ---- */
    const int _x = pos[best_index][0], _y = pos[best_index][1], _z = pos[best_index][2];
    if (_z + bz > H) { e->error = 1; return; }              /* update_level_free_space indexes level _z+bz-1 */
    e->positions[k * 3] = _x; e->positions[k * 3 + 1] = _y; e->positions[k * 3 + 2] = _z;
    e->stable[k] = stab[best_index];
    e->empty = empty_ems[best_index];
    for (int q = _x; q < _x + bx; q++) for (int r = _y; r < _y + by; r++) {
        for (int t = _z; t < _z + bz; t++) C3(q, r, t) = k + 1;
        for (int t = 0; t < _z; t++) if (C3(q, r, t) == 0) C3(q, r, t) = -1;
    }
    {
        const int xx = _x + bx - 1; int err = 0;
        for (int t = _z; t < _z + bz; t++) for (int r = _y; r < _y + by; r++) m3_lfs_occupy(&lfs[t * L + r], _x, xx, bx, &err);
        for (int t = 0; t < _z; t++) for (int r = _y; r < _y + by; r++) m3_lfs_under(&lfs[t * L + r], _x, xx, bx);
        if (err) e->error = err;
    }
    for (int q = _x; q < _x + bx; q++) for (int r = _y; r < _y + by; r++) h[q * L + r] = _z + bz;
    e->valid = valid;
}
#undef C3
#undef M3_PUSH

/* ------------------------------------------------------------------ */
/* LB ("abandoned" but selectable with packing_strategy='LB'): tools.py:1602-1754 (2D), :1756-1914 (3D).
 * Driven through Container.add_new_block (:3683-3686), which never stores the returned bounding_box
 * (`# self.bounding_box = bounding_box`, :3706): every call starts from bounding_box = zeros, so the
 * compactness of a candidate is valid/((_z+bz)*W) -- THIS block's top, not the packing height. */
static int xl_has(const xlist *l, int v) { for (int i = 0; i < l->len; i++) if (l->v[i] == v) return 1; return 0; }
static void xl_remove(xlist *l, int v) { for (int i = 0; i < l->len; i++) if (l->v[i] == v) { for (; i + 1 < l->len; i++) l->v[i] = l->v[i + 1]; l->len--; return; } }

static void lb_step_2d(tapo_env *e) {
    const int W = e->W, H = e->H, k = e->k;
    int *h = e->heightmap, *c = e->container;
    const int bx = e->blocks[k * 2 + 0], bz = e->blocks[k * 2 + 1];
    /* :1640-1650 first-block initialisation when the container is entirely empty */
    int allzero = 1; for (int i = 0; i < W * H && allzero; i++) if (c[i] != 0) allzero = 0;
    if (allzero) { e->valid = 0; e->empty = 0; for (int z = 0; z < H; z++) { e->lbl[z].len = 1; e->lbl[z].v[0] = 0; } memset(h, 0, sizeof(int) * W); }
    long long valid = e->valid + (long long)bx * bz;          /* :1655 */
    static __thread int ems[MAX_EMS][4]; int ne = 0;
    for (int z = 0; z < H; z++) {                              /* :1659-1666 */
        const xlist *fs = &e->lbl[z];
        if (z + bz > H) break;
        else if (z > 0) { int rowzero = 1; for (int x = 0; x < W; x++) if (c[x * H + z - 1] != 0) { rowzero = 0; break; } if (rowzero) break; }
        for (int i = 0; i < fs->len; i++) {
            int x = fs->v[i];
            if (x + bx > W) break;
            if (z > 0 && xl_has(&e->lbl[z - 1], x)) {
                int same = 1; for (int q = x; q < W; q++) if (c[q * H + z] != c[q * H + z - 1]) { same = 0; break; }
                if (same) continue;
            }
            if (ne >= MAX_EMS) { e->error = 2; return; }
            ems[ne][0] = x; ems[ne][1] = z; ne++;
        }
    }
    for (int b = 0; b < k; b++) {                              /* :1668-1677 */
        int x = e->positions[b * 2], z = e->positions[b * 2 + 1], zz = e->blocks[b * 2 + 1];
        if (z + zz < H) {
            if (c[x * H + z + zz] == 0) {
                int dup = 0; for (int i = 0; i < ne; i++) if (ems[i][0] == x && ems[i][1] == z + zz) { dup = 1; break; }
                if (!dup) { if (ne >= MAX_EMS) { e->error = 2; return; } ems[ne][0] = x; ems[ne][1] = z + zz; ne++; }
            }
        }
    }
    static __thread int pos[MAX_EMS][2]; static __thread unsigned char settle[MAX_EMS], stab[MAX_EMS];
    static __thread double comp[MAX_EMS], pyr[MAX_EMS], stb[MAX_EMS]; static __thread long long empty_ems[MAX_EMS];
    const int X = W - bx + 1; int nsettled = 0;
    for (int i = 0; i < ne; i++) {                             /* :1690-1730 */
        settle[i] = stab[i] = 0; comp[i] = pyr[i] = stb[i] = 0.0; empty_ems[i] = e->empty;
        int _z = ems[i][1];
        for (int _x = ems[i][0]; _x < X; _x++) {
            if (settle[i]) break;
            int freeall = 1;                                   /* numpy clips the z slice at H */
            for (int q = _x; q < _x + bx && freeall; q++) for (int zz = _z; zz < _z + bz && zz < H; zz++) if (c[q * H + zz] != 0) { freeall = 0; break; }
            if (!freeall) continue;
            if (_z > 0) {
                if (!is_stable_2d(&c[_x * H + _z - 1], H, bx, _x, bx)) { if (e->hard) continue; }
                else stab[i] = 1;
            } else stab[i] = 1;
            pos[i][0] = _x; pos[i][1] = _z; settle[i] = 1;
        }
        if (settle[i]) {
            nsettled++;
            int _x = pos[i][0]; _z = pos[i][1];
            long long bbox = (long long)(_z + bz) * W;         /* bounding_box arrives as zeros: see the note above */
            comp[i] = (double)valid / (double)bbox;
            long long cnt = 0; for (int q = _x; q < _x + bx; q++) for (int zz = 0; zz < _z && zz < H; zz++) if (c[q * H + zz] == 0) cnt++;
            empty_ems[i] += cnt;
            if (e->useP) pyr[i] = (double)valid / (double)(empty_ems[i] + valid);
            if (e->useS) { int sn = 0; for (int q = 0; q < k; q++) sn += e->stable[q]; sn += stab[i]; stb[i] = (double)sn / (double)(k + 1); }
        }
    }
    if (nsettled == 0) { e->stable[k] = 0; return; }           /* :1733-1736 */
    static __thread double ratio[MAX_EMS];
    for (int i = 0; i < ne; i++) ratio[i] = (comp[i] + pyr[i]) + stb[i];
    int best = argmax_first(ratio, ne);                        /* the remove-loop at :1741-1743 never runs: unsettled scores are 0.0 */
    int _x = pos[best][0], _z = pos[best][1];
    if (_z + bz > H) { e->error = 1; return; }                 /* level_free_space[_z+zz] IndexError below */
    e->positions[k * 2] = _x; e->positions[k * 2 + 1] = _z;
    e->stable[k] = stab[best];
    e->empty = empty_ems[best];
    for (int q = _x; q < _x + bx; q++) {
        for (int zz = _z; zz < _z + bz; zz++) c[q * H + zz] = k + 1;
        for (int zz = 0; zz < _z; zz++) if (c[q * H + zz] == 0) c[q * H + zz] = -1;
    }
    for (int zz = 0; zz < bz; zz++) {                          /* :1757-1761 */
        xlist *fs = &e->lbl[_z + zz];
        if (xl_has(fs, _x)) xl_remove(fs, _x);
        if (_x + bx < W && c[(_x + bx) * H + _z + zz] == 0) { if (fs->len >= TAPO_MAXN + 4) { e->error = 2; return; } fs->v[fs->len++] = _x + bx; }
    }
    for (int q = _x; q < _x + bx; q++) h[q] = _z + bz;         /* :1764 */
    e->valid = valid;
}

static void lb_step_3d(tapo_env *e) {
    const int W = e->W, L = e->L, H = e->H, k = e->k;
    int *h = e->heightmap, *c = e->container;
#define CT(x, y, z) c[((x) * L + (y)) * H + (z)]
#define LF(z, y) e->lbl[(z) * L + (y)]
    const int bx = e->blocks[k * 3], by = e->blocks[k * 3 + 1], bz = e->blocks[k * 3 + 2];
    int allzero = 1; for (int i = 0; i < W * L * H && allzero; i++) if (c[i] != 0) allzero = 0;
    if (allzero) { e->valid = 0; e->empty = 0; for (int i = 0; i < H * L; i++) { e->lbl[i].len = 1; e->lbl[i].v[0] = 0; } memset(h, 0, sizeof(int) * W * L); }
    long long valid = e->valid + (long long)bx * by * bz;
    static __thread int ems[MAX_EMS * 4][4]; int ne = 0;
    const int EMAX = MAX_EMS * 4;
    for (int z = 0; z < H; z++) {                              /* :1810-1823 */
        if (z + bz > H) break;
        else if (z > 0) { int zero = 1; for (int x = 0; x < W && zero; x++) for (int y = 0; y < L; y++) if (CT(x, y, z - 1) != 0) { zero = 0; break; } if (zero) break; }
        for (int y = 0; y < L; y++) {
            const xlist *fs = &LF(z, y);
            if (y + by > L) break;
            else if (y > 0 && LF(z, y - 1).len == 1 && LF(z, y - 1).v[0] == 0) continue;   /* free_space_x[y-1] == [0] */
            for (int i = 0; i < fs->len; i++) {
                int x = fs->v[i];
                if (x + bx > W) break;
                if (y > 0 && xl_has(&LF(z, y - 1), x)) { int same = 1; for (int q = x; q < W; q++) if (CT(q, y, z) != CT(q, y - 1, z)) { same = 0; break; } if (same) continue; }
                if (z > 0 && xl_has(&LF(z - 1, y), x)) { int same = 1; for (int q = x; q < W && same; q++) for (int r = y; r < L; r++) if (CT(q, r, z) != CT(q, r, z - 1)) { same = 0; break; } if (same) continue; }
                if (ne >= EMAX) { e->error = 2; return; }
                ems[ne][0] = x; ems[ne][1] = y; ems[ne][2] = z; ne++;
            }
        }
    }
    for (int b = 0; b < k; b++) {                              /* :1824-1833 */
        int x = e->positions[b * 3], y = e->positions[b * 3 + 1], z = e->positions[b * 3 + 2];
        int yy = e->blocks[b * 3 + 1], zz = e->blocks[b * 3 + 2];
        if (y + yy < L) {
            if (CT(x, y + yy, z) == 0) { int dup = 0; for (int i = 0; i < ne; i++) if (ems[i][0] == x && ems[i][1] == y + yy && ems[i][2] == z) { dup = 1; break; }
                if (!dup) { if (ne >= EMAX) { e->error = 2; return; } ems[ne][0] = x; ems[ne][1] = y + yy; ems[ne][2] = z; ne++; } }
        }
        if (z + zz < H) {
            if (CT(x, y, z + zz) == 0) { int dup = 0; for (int i = 0; i < ne; i++) if (ems[i][0] == x && ems[i][1] == y && ems[i][2] == z + zz) { dup = 1; break; }
                if (!dup) { if (ne >= EMAX) { e->error = 2; return; } ems[ne][0] = x; ems[ne][1] = y; ems[ne][2] = z + zz; ne++; } }
        }
    }
    static __thread int pos[MAX_EMS * 4][3]; static __thread unsigned char settle[MAX_EMS * 4], stab[MAX_EMS * 4];
    static __thread double comp[MAX_EMS * 4], pyr[MAX_EMS * 4], stb[MAX_EMS * 4]; static __thread long long empty_ems[MAX_EMS * 4];
    const int X = W - bx + 1, Y = L - by + 1; int nsettled = 0;
    for (int i = 0; i < ne; i++) {
        settle[i] = stab[i] = 0; comp[i] = pyr[i] = stb[i] = 0.0; empty_ems[i] = e->empty;
        int _z = ems[i][2];
        for (int _x = ems[i][0]; _x < X && !settle[i]; _x++) for (int _y = ems[i][1]; _y < Y; _y++) {
            if (settle[i]) break;
            int freeall = 1;
            for (int p = _x; p < _x + bx && freeall; p++) for (int q = _y; q < _y + by && freeall; q++) for (int zz = _z; zz < _z + bz && zz < H; zz++) if (CT(p, q, zz) != 0) { freeall = 0; break; }
            if (!freeall) continue;
            if (!is_stable_3d(e, bx, by, _x, _y, _z)) { if (e->hard) continue; }
            else stab[i] = 1;
            pos[i][0] = _x; pos[i][1] = _y; pos[i][2] = _z; settle[i] = 1;
        }
        if (settle[i]) {
            nsettled++;
            int _x = pos[i][0], _y = pos[i][1]; _z = pos[i][2];
            long long bbox = (long long)(_z + bz) * W * L;
            comp[i] = (double)valid / (double)bbox;
            long long cnt = 0; for (int p = _x; p < _x + bx; p++) for (int q = _y; q < _y + by; q++) for (int zz = 0; zz < _z && zz < H; zz++) if (CT(p, q, zz) == 0) cnt++;
            empty_ems[i] += cnt;
            if (e->useP) pyr[i] = (double)valid / (double)(empty_ems[i] + valid);
            if (e->useS) { int sn = 0; for (int q = 0; q < k; q++) sn += e->stable[q]; sn += stab[i]; stb[i] = (double)sn / (double)(k + 1); }
        }
    }
    if (nsettled == 0) { e->stable[k] = 0; return; }
    static __thread double ratio[MAX_EMS * 4];
    for (int i = 0; i < ne; i++) ratio[i] = (comp[i] + pyr[i]) + stb[i];
    int best = argmax_first(ratio, ne);
    int _x = pos[best][0], _y = pos[best][1], _z = pos[best][2];
    if (_z + bz > H) { e->error = 1; return; }
    e->positions[k * 3] = _x; e->positions[k * 3 + 1] = _y; e->positions[k * 3 + 2] = _z;
    e->stable[k] = stab[best];
    e->empty = empty_ems[best];
    for (int p = _x; p < _x + bx; p++) for (int q = _y; q < _y + by; q++) {
        for (int zz = _z; zz < _z + bz; zz++) CT(p, q, zz) = k + 1;
        for (int zz = 0; zz < _z; zz++) if (CT(p, q, zz) == 0) CT(p, q, zz) = -1;
    }
    for (int zz = 0; zz < bz; zz++) for (int yy = 0; yy < by; yy++) {       /* :1905-1909 */
        xlist *fs = &LF(_z + zz, _y + yy);
        if (xl_has(fs, _x)) xl_remove(fs, _x);
        if (_x + bx < W && CT(_x + bx, _y + yy, _z + zz) == 0) { if (fs->len >= TAPO_MAXN + 4) { e->error = 2; return; } fs->v[fs->len++] = _x + bx; }
    }
    for (int p = _x; p < _x + bx; p++) for (int q = _y; q < _y + by; q++) h[p * L + q] = _z + bz;
    e->valid = valid;
#undef CT
#undef LF
}

/* ------------------------------------------------------------------ */
/* heightmap encodings, tools.py:3716-3743.  Returns number of ints written. */
int tapo_env_encode_heightmap(const tapo_env *e, int *out) {
    const int W = e->W, L = e->L; const int *h = e->heightmap;
    if (e->dim == 2) {
        if (e->hm_type == TAPO_HM_FULL) { memcpy(out, h, sizeof(int) * W); return W; }
        if (e->hm_type == TAPO_HM_ZERO) { int m = h[0]; for (int i = 1; i < W; i++) if (h[i] < m) m = h[i]; for (int i = 0; i < W; i++) out[i] = h[i] - m; return W; }
        for (int i = 0; i + 1 < W; i++) out[i] = h[i + 1] - h[i];   /* :3738-3743 */
        return W - 1;
    }
    if (e->hm_type == TAPO_HM_FULL) { memcpy(out, h, sizeof(int) * W * L); return W * L; }
    if (e->hm_type == TAPO_HM_ZERO) { int m = h[0]; for (int i = 1; i < W * L; i++) if (h[i] < m) m = h[i]; for (int i = 0; i < W * L; i++) out[i] = h[i] - m; return W * L; }
    for (int x = 0; x < W; x++) for (int y = 0; y < L; y++) {       /* :3721-3737 */
        out[x * L + y] = x == 0 ? 0 : h[x * L + y] - h[(x - 1) * L + y];
        out[W * L + x * L + y] = y == 0 ? 0 : h[x * L + y] - h[x * L + y - 1];
    }
    return 2 * W * L;
}

/* tools.py:3663-3744 Container.add_new_block(block).  `block` is the float
 * row the caller hands over (model.py:412); `.astype(int)` truncates (:3689). */
int tapo_env_add_new_block(tapo_env *e, const float *block, int *hm_out) {
    if (e->k >= e->n) { e->error = 1; return -1; }         /* IndexError on rotate_state[k] */
    for (int d = 0; d < e->dim; d++) e->blocks[e->k * e->dim + d] = (int)block[d];
    const int *b = &e->blocks[e->k * e->dim];
    if (e->strategy == TAPO_MACS) {
        if (e->dim == 2) macs_step_2d(e); else macs_step_3d(e);
    } else if (e->strategy == TAPO_LB) {
        if (e->dim == 2) lb_step_2d(e); else lb_step_3d(e);
    } else {
        if (e->dim == 2) lbg_step_2d(e, b[0], b[1]); else lbg_step_3d(e, b[0], b[1], b[2]);
    }
    e->k += 1;                                             /* :3713, even when placement failed */
    return tapo_env_encode_heightmap(e, hm_out);
}

/* tools.py:3887-3906 calc_CPS */
void tapo_env_calc_cps(const tapo_env *e, double *C, double *P, double *S) {
    if (e->k == 0) { *C = *P = *S = 0.0; return; }
    int cells = e->W * e->L, height = e->heightmap[0];
    for (int i = 1; i < cells; i++) if (e->heightmap[i] > height) height = e->heightmap[i];
    long long box = (long long)cells * height;
    int sn = 0; for (int i = 0; i < e->n; i++) sn += e->stable[i];
    *C = (double)e->valid / (double)box;
    *P = (double)e->valid / (double)(e->empty + e->valid);
    *S = (double)sn / (double)e->k;
}

/* tools.py:3908-3966 calc_ratio: first matching branch of the if/elif chain */
double tapo_env_calc_ratio(const tapo_env *e) {
    double C, P, S, ratio; const char *rt = e->reward_type;
    tapo_env_calc_cps(e, &C, &P, &S);
    if (!strcmp(rt, "comp")) ratio = C;
    else if (!strcmp(rt, "soft") || !strcmp(rt, "hard")) ratio = C * S;
    else if (!strcmp(rt, "pyrm")) ratio = C + P;
    else if (!strcmp(rt, "pyrm-soft") || !strcmp(rt, "pyrm-hard") || !strcmp(rt, "mcs-soft") || !strcmp(rt, "mcs-hard")) ratio = (C + P) * S;
    else if (!strcmp(rt, "pyrm-soft-sum") || !strcmp(rt, "pyrm-hard-sum")) ratio = C + P + S;
    else if (!strcmp(rt, "pyrm-soft-SUM") || !strcmp(rt, "pyrm-hard-SUM")) ratio = 2 * C + P + S;
    else if (!strcmp(rt, "CPS")) ratio = C * P * S;
    else if (!strncmp(rt, "C+P", 3)) ratio = C + P + S;   /* every 'C+P-*' / 'C+P+S-*' row of the table */
    else return NAN;                                       /* reference prints and raises NameError */
    if (!strcmp(rt, "C+P-lb-soft")) return (C + P) / 2;    /* :3961-3962 */
    return ratio / 3;
}

/* accessors for the ctypes wrapper */
int tapo_env_k(const tapo_env *e) { return e->k; }
int tapo_env_error(const tapo_env *e) { return e->error; }
long long tapo_env_valid(const tapo_env *e) { return e->valid; }
long long tapo_env_empty(const tapo_env *e) { return e->empty; }
const int *tapo_env_heightmap(const tapo_env *e) { return e->heightmap; }
const int *tapo_env_positions(const tapo_env *e) { return e->positions; }
const int *tapo_env_container(const tapo_env *e) { return e->container; }
const unsigned char *tapo_env_stable(const tapo_env *e) { return e->stable; }
int tapo_env_strategy(const tapo_env *e) { return e->strategy; }
/* MACS: copy level z of level_free_space, returns its length */
int tapo_env_lfs3(const tapo_env *e, int z, int y, int *out) { if (!e->lfs3 || z >= e->H || y >= e->L) return -1; const ilist *l = &e->lfs3[z * e->L + y]; memcpy(out, l->v, sizeof(int) * l->len); return l->len; }
int tapo_env_lfs(const tapo_env *e, int z, int *out) { if (!e->lfs || z >= e->H) return -1; memcpy(out, e->lfs[z].v, sizeof(int) * e->lfs[z].len); return e->lfs[z].len; }

/* ------------------------------------------------------------------ */
/* pack.py:333-376 update_dynamic (out-of-place: clone + `update_time` row scatters) */
void tapo_update_dynamic(const float *dynamic, const float *static_, const int64_t *ptr,
                         int B, int rows, int S, int srows, int n, int update_time, float *out) {
    for (int b = 0; b < B; b++) {
        const float *d = dynamic + (size_t)b * rows * S; float *o = out + (size_t)b * rows * S;
        memcpy(o, d, sizeof(float) * rows * S);                                   /* :370 */
        long real = (long)static_[(size_t)b * srows * S + 0 * S + ptr[b]];        /* :347 */
        for (int i = 0; i < update_time; i++) {                                   /* :372-374 */
            long r = real + (long)n * i;
            if (r >= 0 && r < rows) memset(o + r * S, 0, sizeof(float) * S);
        }
    }
}

/* pack.py:297-331 update_mask; also model.py:297-307 when ptr == NULL (initial mask) */
void tapo_update_mask(const float *mask, const float *dynamic, const int64_t *ptr,
                      int B, int rows, int S, int n, int R, float *new_mask, float *chosen_mask) {
    for (int b = 0; b < B; b++) {
        const float *d = dynamic + (size_t)b * rows * S;
        float *cm = chosen_mask + (size_t)b * S, *nm = new_mask + (size_t)b * S;
        for (int j = 0; j < S; j++) cm[j] = mask ? mask[(size_t)b * S + j] : 1.0f;
        if (ptr) {
            long real = ptr[b];
            while (real >= n) real -= n;                                          /* :314-316 */
            for (int i = 0; i < R; i++) if (real + (long)n * i < S) cm[real + (long)n * i] = 0.0f;   /* :320-321 */
        }
        for (int j = 0; j < S; j++) {
            float mv = 0.f, sm = 0.f, lg = 0.f;                                   /* :324-326 .sum(1) */
            for (int i = 0; i < n; i++) mv += d[(size_t)i * S + j];
            if (rows >= 3 * n) {
                for (int i = 0; i < n; i++) sm += d[(size_t)(n + i) * S + j];
                for (int i = 0; i < n; i++) lg += d[(size_t)(2 * n + i) * S + j];
            }
            float dm = sm * lg + mv;                                              /* :327-328 */
            nm[j] = dm != 0.f ? 0.f : cm[j];                                      /* :329 */
        }
    }
}

/* ------------------------------------------------------------------ */
/* Whole-batch episode driver: the CPU baseline of bench.py ("port").
 * Per env and per decode step: update_dynamic (clone) + update_mask +
 * add_new_block, then calc_ratio -- the same per-step work the reference's
 * decode loop performs (model.py:376-453, :509-510).  Environments are
 * independent -> the batch is cut into `nthreads` contiguous slices, one
 * pthread each (the reference itself is single-threaded).
 * ptr_seq is [steps][B] int64.  Outputs (any may be NULL):
 *   heightmap_out [B][cells] int32 (final), pos_out [B][n][dim] int32,
 *   stable_out [B][n] u8, reward_out [B] float (calc_ratio, fp64 -> fp32),
 *   cur_mask_out / mask_out [B][S] float (after the last step),
 *   dynamic_out [B][rows][S] float (after the last step),
 *   dec_dyn_out [B][enc] int32 (last returned heightmap encoding).
 * Returns 0, or the first non-zero env error. */
typedef struct {
    int dim, W, L, H, n, R, hm_type, strategy, B, steps, b0, b1, status;
    int nwin, cap;   /* rolling-style: nwin windows of n blocks into ONE container of capacity cap (rolling.py:702-703) */
    const char *reward_type;
    const float *static_, *dynamic; const int64_t *ptr_seq;
    int *heightmap_out, *pos_out; unsigned char *stable_out; float *reward_out;
    float *cur_mask_out, *mask_out, *dynamic_out; int *dec_dyn_out;
} ep_job;

static void *ep_worker(void *arg) {
    ep_job *j = (ep_job *)arg;
    const int dim = j->dim, W = j->W, L = j->L, n = j->n, R = j->R, B = j->B;
    const int S = n * R, rows = 3 * n, srows = 1 + dim;
    const int cells = dim == 2 ? W : W * L;
    const int enc = dim == 2 ? (j->hm_type == TAPO_HM_DIFF ? W - 1 : W) : (j->hm_type == TAPO_HM_DIFF ? 2 * W * L : W * L);
    tapo_env *e = tapo_env_new(dim, W, L, j->H, j->cap, j->reward_type, j->hm_type, j->strategy);
    if (!e) { j->status = 2; return NULL; }
    float *da = (float *)malloc(sizeof(float) * rows * S), *db = (float *)malloc(sizeof(float) * rows * S);
    float *m = (float *)malloc(sizeof(float) * S), *cm = (float *)malloc(sizeof(float) * S), *nm = (float *)malloc(sizeof(float) * S);
    int *hm = (int *)calloc(2 * cells + 4, sizeof(int));
    for (int b = j->b0; b < j->b1; b++) {
        tapo_env_clear(e);                                           /* model.py:294 */
        for (int w = 0; w < j->nwin; w++) {
            const float *st = j->static_ + ((size_t)w * B + b) * srows * S;
            memcpy(da, j->dynamic + ((size_t)w * B + b) * rows * S, sizeof(float) * rows * S);
            tapo_update_mask(NULL, da, NULL, 1, rows, S, n, R, nm, m);  /* model.py:297-307; m = ones */
            for (int t = 0; t < j->steps; t++) {
                int64_t p = j->ptr_seq[((size_t)w * j->steps + t) * B + b];
                tapo_update_dynamic(da, st, &p, 1, rows, S, srows, n, 3, db);      /* model.py:376 */
                tapo_update_mask(m, db, &p, 1, rows, S, n, R, nm, cm);             /* model.py:384 */
                memcpy(m, cm, sizeof(float) * S);
                float blk[3]; for (int d = 0; d < dim; d++) blk[d] = st[(size_t)(1 + d) * S + p];   /* model.py:404-412 */
                tapo_env_add_new_block(e, blk, hm);                                /* model.py:453 */
                float *tsw = da; da = db; db = tsw;
            }
        }
        if (e->error && !j->status) j->status = e->error;
        if (j->heightmap_out) memcpy(j->heightmap_out + (size_t)b * cells, e->heightmap, sizeof(int) * cells);
        if (j->pos_out) memcpy(j->pos_out + (size_t)b * j->cap * dim, e->positions, sizeof(int) * j->cap * dim);
        if (j->stable_out) memcpy(j->stable_out + (size_t)b * j->cap, e->stable, j->cap);
        if (j->reward_out) j->reward_out[b] = (float)tapo_env_calc_ratio(e);   /* model.py:509-510 */
        if (j->cur_mask_out) memcpy(j->cur_mask_out + (size_t)b * S, nm, sizeof(float) * S);
        if (j->mask_out) memcpy(j->mask_out + (size_t)b * S, m, sizeof(float) * S);
        if (j->dynamic_out) memcpy(j->dynamic_out + (size_t)b * rows * S, da, sizeof(float) * rows * S);
        if (j->dec_dyn_out) memcpy(j->dec_dyn_out + (size_t)b * enc, hm, sizeof(int) * enc);
    }
    tapo_env_free(e); free(da); free(db); free(m); free(cm); free(nm); free(hm);
    return NULL;
}

int tapo_episode_batch(int dim, int W, int L, int H, int n, int R, const char *reward_type, int hm_type, int strategy,
                       int B, int steps, const float *static_, const float *dynamic, const int64_t *ptr_seq,
                       int *heightmap_out, int *pos_out, unsigned char *stable_out, float *reward_out,
                       float *cur_mask_out, float *mask_out, float *dynamic_out, int *dec_dyn_out, int nthreads,
                       int nwin, int cap) {
    if (nwin < 1) nwin = 1;
    if (cap < n) cap = n;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > B) nthreads = B > 0 ? B : 1;
    if (nthreads > 1024) nthreads = 1024;
    ep_job *jobs = (ep_job *)calloc(nthreads, sizeof(ep_job));
    pthread_t *th = (pthread_t *)calloc(nthreads, sizeof(pthread_t));
    int status = 0;
    for (int t = 0; t < nthreads; t++) {
        ep_job *j = &jobs[t];
        j->dim = dim; j->W = W; j->L = L; j->H = H; j->n = n; j->R = R; j->hm_type = hm_type; j->strategy = strategy;
        j->B = B; j->steps = steps; j->reward_type = reward_type; j->nwin = nwin; j->cap = cap;
        j->b0 = (int)((long long)B * t / nthreads); j->b1 = (int)((long long)B * (t + 1) / nthreads);
        j->static_ = static_; j->dynamic = dynamic; j->ptr_seq = ptr_seq;
        j->heightmap_out = heightmap_out; j->pos_out = pos_out; j->stable_out = stable_out; j->reward_out = reward_out;
        j->cur_mask_out = cur_mask_out; j->mask_out = mask_out; j->dynamic_out = dynamic_out; j->dec_dyn_out = dec_dyn_out;
        if (t > 0) pthread_create(&th[t], NULL, ep_worker, j);
    }
    ep_worker(&jobs[0]);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
    for (int t = 0; t < nthreads; t++) if (jobs[t].status && !status) status = jobs[t].status;
    free(jobs); free(th);
    return status;
}
