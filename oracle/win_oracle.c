/*
 * win_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as tap_oracle.c).
 *
 * Plain-C, CPU restatement of the rolling-window logic of Juzhan/TAP-Net @ 6eded31:
 * generate.InitialContainer.sub_deps_graph / convert_to_input / remove_block / is_last_graph
 * (generate.py:1674-1823), driven per decode step by rolling.validate (rolling.py:575-658).
 * The restatement is LITERAL: explicit adjacency matrices, Python-list semantics for
 * `after_nodes_list` and `sub_graph_nodes`, in-degrees by counting edges.  The CUDA kernel
 * (tap-net_b200/csrc/window.cuh) uses 64-bit node masks instead.
 *
 * Third-party behaviour on this path (neither is under /root/reference):
 *   - networkx (README.md lists it unversioned; 3.6.1 installed here): `G.subgraph(nodes).copy()`
 *     (generate.py:1684-1688) enumerates its nodes through networkx.classes.coreviews.FilterAtlas.__iter__,
 *     which iterates the *set* built from `nodes` when 2*len(set) < len(G) and the graph's own (ascending)
 *     node order otherwise.
 *   - CPython (3.12 here) Objects/setobject.c: the iteration order of that set of small ints
 *     (hash(i) == i; table 8 -> 32 -> 128 slots, linear probing over 9 neighbours, then the perturbed
 *     recurrence i = 5i + 1 + perturb).  tapo_pyset_order() restates set_add_entry / set_table_resize /
 *     set_insert_clean; tests pin it against the running interpreter's own `set`.
 *   `G_to_deps` (generate.py:1750-1753) indexes the dependency matrices by that enumeration while `static`
 *   uses the SORTED node list (:1766, :1781-1790): with total=50, window=10 the two orders differ in most
 *   windows.  This is reference behaviour and is reproduced (order mode 0); mode 1 = sorted everywhere.
 *
 * Pinning: live differential runs against generate.InitialContainer (tests/test_oracle_vs_reference.py) and the
 * recorded trajectories tests/golden/traj_rolling_*.npz (tests/golden/make_golden.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define TAPW_MAXT 64     /* total blocks per instance */
#define TAPW_MAXSET 512  /* set table slots */

/* ---- CPython set insertion order for distinct small non-negative ints ---- */
#define PYSET_LINEAR_PROBES 9
#define PYSET_PERTURB_SHIFT 5
#define PYSET_MINSIZE 8

static void pyset_insert_clean(int *table, size_t mask, int key) {      /* setobject.c set_insert_clean */
    size_t perturb = (size_t)key, i = (size_t)key & mask;
    for (;;) {
        if (table[i] < 0) { table[i] = key; return; }
        if (i + PYSET_LINEAR_PROBES <= mask)
            for (size_t j = 1; j <= PYSET_LINEAR_PROBES; j++)
                if (table[i + j] < 0) { table[i + j] = key; return; }
        perturb >>= PYSET_PERTURB_SHIFT;
        i = (i * 5 + 1 + perturb) & mask;
    }
}

/* keys[k] (distinct or not, 0 <= key) inserted in order into an empty set; out = iteration order. Returns count. */
int tapo_pyset_order(const int *keys, int k, int *out) {
    static __thread int ta[TAPW_MAXSET], tb[TAPW_MAXSET];
    int *table = ta, *other = tb;
    size_t mask = PYSET_MINSIZE - 1;
    int fill = 0;
    for (size_t i = 0; i <= mask; i++) table[i] = -1;
    for (int q = 0; q < k; q++) {                                       /* set_add_entry */
        const int key = keys[q];
        size_t perturb = (size_t)key, i = (size_t)key & mask;
        int found = 0, placed = 0;
        while (!found && !placed) {
            size_t probes = (i + PYSET_LINEAR_PROBES <= mask) ? PYSET_LINEAR_PROBES : 0;
            for (size_t j = 0; j <= probes; j++) {
                if (table[i + j] < 0) { table[i + j] = key; fill++; placed = 1; break; }
                if (table[i + j] == key) { found = 1; break; }
            }
            if (found || placed) break;
            perturb >>= PYSET_PERTURB_SHIFT;
            i = (i * 5 + 1 + perturb) & mask;
        }
        if (found) continue;
        if ((size_t)fill * 5 < mask * 3) continue;
        size_t minused = (size_t)fill * 4, newsize = PYSET_MINSIZE;      /* set_table_resize(used*4), used <= 50000 */
        while (newsize <= minused) newsize <<= 1;
        if (newsize > TAPW_MAXSET) return -1;
        for (size_t s = 0; s < newsize; s++) other[s] = -1;
        for (size_t s = 0; s <= mask; s++) if (table[s] >= 0) pyset_insert_clean(other, newsize - 1, table[s]);
        int *t = table; table = other; other = t;
        mask = newsize - 1;
    }
    int cnt = 0;
    for (size_t s = 0; s <= mask; s++) if (table[s] >= 0) out[cnt++] = table[s];
    return cnt;
}

/* ---- InitialContainer ---- */
typedef struct tapo_win {
    int T, n, dim, R, order_mode;
    unsigned char *adj[5];        /* [T][T]: adj[g][u*T+v] != 0 <=> edge u -> v in G_move/G_left/G_right/G_forward/G_backward */
    int *blocks;                  /* [R*T][dim]  self.blocks (generate.py:1615) */
    unsigned char in_gm[TAPW_MAXT];   /* node still in self.gm (generate.py:1666) */
    int after[TAPW_MAXT], after_len;  /* self.after_nodes_list (generate.py:1672) */
    int win[TAPW_MAXT], win_len;      /* self.sub_graph_nodes  (generate.py:1673) */
    int error;                    /* 1: the reference would spin forever (no in-degree-0 node), 2: window not full at convert */
} tapo_win;

tapo_win *tapo_win_new(int T, int n, int dim, const unsigned char *adj5 /*[5][T][T]*/, const int *blocks /*[R*T][dim]*/,
                       int order_mode) {
    if (T < 1 || T > TAPW_MAXT || n < 1 || n > T || (dim != 2 && dim != 3)) return NULL;
    tapo_win *w = (tapo_win *)calloc(1, sizeof(tapo_win));
    w->T = T; w->n = n; w->dim = dim; w->R = dim == 2 ? 2 : 6; w->order_mode = order_mode;
    for (int g = 0; g < 5; g++) {
        w->adj[g] = (unsigned char *)malloc((size_t)T * T);
        memcpy(w->adj[g], adj5 + (size_t)g * T * T, (size_t)T * T);
    }
    w->blocks = (int *)malloc(sizeof(int) * w->R * T * dim);
    memcpy(w->blocks, blocks, sizeof(int) * w->R * T * dim);
    for (int v = 0; v < T; v++) { w->in_gm[v] = 1; w->after[v] = v; }   /* generate.py:1666,1672 */
    w->after_len = T; w->win_len = 0;
    return w;
}

void tapo_win_free(tapo_win *w) {
    if (!w) return;
    for (int g = 0; g < 5; g++) free(w->adj[g]);
    free(w->blocks); free(w);
}

static int list_has(const int *l, int len, int v) { for (int i = 0; i < len; i++) if (l[i] == v) return 1; return 0; }
static void list_remove(int *l, int *len, int v) {
    for (int i = 0; i < *len; i++) if (l[i] == v) { for (; i + 1 < *len; i++) l[i] = l[i + 1]; (*len)--; return; }
}

/* decompose(nodes) generate.py:1682-1724: the five sub-matrices [n][n] in the subgraph's node enumeration */
static void win_decompose(tapo_win *w, int *mats /*[5][n][n]*/) {
    const int T = w->T, n = w->n;
    int P[TAPW_MAXT], np_;
    /* node order of G.subgraph(nodes).copy(): FilterAtlas.__iter__ (see header) */
    int distinct = 0;
    { unsigned char seen[TAPW_MAXT] = {0}; for (int i = 0; i < w->win_len; i++) if (!seen[w->win[i]]) { seen[w->win[i]] = 1; distinct++; } }
    if (w->order_mode == 0 && 2 * distinct < T) {
        np_ = tapo_pyset_order(w->win, w->win_len, P);
    } else {
        np_ = 0;
        for (int v = 0; v < T; v++) if (list_has(w->win, w->win_len, v)) P[np_++] = v;
    }
    for (int g = 0; g < 5; g++) {
        int *M = mats + (size_t)g * n * n;
        for (int i = 0; i < np_ && i < n; i++)
            for (int j = 0; j < np_ && j < n; j++)
                if (w->adj[g][P[i] * T + P[j]]) M[i * n + j] = 1;           /* G_to_deps generate.py:1750-1753 */
        if (g == 0) continue;
        for (int i = 0; i < np_ && i < n; i++)                               /* generate.py:1690-1705 */
            for (int u = 0; u < T; u++)
                if (w->adj[g][u * T + P[i]] && list_has(w->after, w->after_len, u)) M[i * n + i] = 1;
    }
    for (int i = 0; i < w->win_len; i++) w->in_gm[w->win[i]] = 0;          /* generate.py:1713-1723 */
}

/* sub_deps_graph generate.py:1674-1768 */
static void win_sub_deps_graph(tapo_win *w, int *mats /*[5][n][n], zeroed here*/) {
    const int T = w->T, n = w->n;
    memset(mats, 0, sizeof(int) * 5 * n * n);
    unsigned char alive[TAPW_MAXT];
    int alive_cnt = 0;
    for (int v = 0; v < T; v++) { alive[v] = w->in_gm[v]; alive_cnt += alive[v]; }   /* gm_copy = self.gm.copy() */
    int stop = 0;
    while (alive_cnt > 0 && !stop) {
        int nodes[TAPW_MAXT], nn = 0;
        if (alive_cnt == 1) { for (int v = 0; v < T; v++) if (alive[v]) nodes[nn++] = v; }
        else {
            for (int v = 0; v < T; v++) {
                if (!alive[v]) continue;
                int deg = 0;
                for (int u = 0; u < T; u++) if (alive[u] && w->adj[0][u * T + v]) deg++;
                if (deg == 0) nodes[nn++] = v;
            }
            if (nn == 0) { w->error |= 1; break; }                          /* the reference loops forever */
        }
        for (int q = 0; q < nn; q++) {
            const int node = nodes[q];
            if (w->win_len == n) { win_decompose(w, mats); stop = 1; break; }
            w->win[w->win_len++] = node;                                    /* :1742 */
            alive[node] = 0; alive_cnt--;                                   /* :1743 */
            list_remove(w->after, &w->after_len, node);                     /* :1744 */
            if (w->win_len == n) { win_decompose(w, mats); stop = 1; break; }
        }
    }
    for (int i = 1; i < w->win_len; i++) {                                  /* self.sub_graph_nodes.sort() :1766 */
        int t = w->win[i], j = i - 1;
        while (j >= 0 && w->win[j] > t) { w->win[j + 1] = w->win[j]; j--; }
        w->win[j + 1] = t;
    }
}

static const int PERM2[2][2] = {{0, 1}, {1, 0}};
static const int PERM3[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};   /* itertools.permutations */

/* convert_to_input generate.py:1770-1808 ('bot' input).  static_out [1+dim][S], dynamic_out [3n][S] (float).
 * Returns 0, or 2 when the window could not be filled (the reference raises in np.concatenate). */
int tapo_win_convert_to_input(tapo_win *w, float *static_out, float *dynamic_out) {
    const int T = w->T, n = w->n, dim = w->dim, R = w->R, S = n * R;
    int *mats = (int *)malloc(sizeof(int) * 5 * n * n);
    win_sub_deps_graph(w, mats);
    memset(static_out, 0, sizeof(float) * (1 + dim) * S);
    memset(dynamic_out, 0, sizeof(float) * 3 * n * S);
    int rc = 0;
    if (w->win_len != n) { w->error |= 2; rc = 2; }
    for (int r = 0; r < R; r++) {
        const int last = dim == 2 ? PERM2[r][1] : PERM3[r][2];
        const int ga = last == 0 ? 1 : (last == 1 ? 3 : -1), gb = last == 0 ? 2 : (last == 1 ? 4 : -1);   /* :1793-1806; 2D up/down = zeros = empty forward/backward */
        for (int i = 0; i < n; i++) {
            const int col = r * n + i;
            static_out[col] = (float)i;                                     /* static_index :1781-1786 */
            if (i < w->win_len)
                for (int d = 0; d < dim; d++)
                    static_out[(1 + d) * S + col] = (float)w->blocks[(w->win[i] + r * T) * dim + d];   /* blocks[rotate_order] :1779,1788 */
            for (int row = 0; row < n; row++) {
                dynamic_out[(size_t)row * S + col] = (float)mats[(0 * n + row) * n + i];
                if (ga >= 0 && !(dim == 2 && last == 1)) {
                    dynamic_out[(size_t)(n + row) * S + col] = (float)mats[((size_t)ga * n + row) * n + i];
                    dynamic_out[(size_t)(2 * n + row) * S + col] = (float)mats[((size_t)gb * n + row) * n + i];
                }
            }
        }
    }
    free(mats);
    return rc;
}

void tapo_win_remove_block(tapo_win *w, int block_id) { list_remove(w->win, &w->win_len, block_id); }   /* :1810-1822 */
int tapo_win_is_last_graph(const tapo_win *w) { return w->after_len == 0; }                             /* :1824-1825 */
int tapo_win_nodes(const tapo_win *w, int *out) { memcpy(out, w->win, sizeof(int) * w->win_len); return w->win_len; }
int tapo_win_error(const tapo_win *w) { return w->error; }

/* ------------------------------------------------------------------ */
/* Rolling episode driver (CPU baseline of bench.py for the rolling workload): rolling.validate's loop
 * (rolling.py:575-640) without the network.  Per instance: ONE container of capacity T; T-n+1 windows; the first
 * T-n are decoded for one step (one_step=True -> max_steps = 1, rolling.py:349-350), the last one completely.
 * Each decode step still performs update_dynamic + update_mask (rolling.py:404-412) + add_new_block (:436).
 * ptr_seq [T][B] int64: step t of instance b.  Uses the environment functions of tap_oracle.c. */
typedef struct tapo_env tapo_env;
tapo_env *tapo_env_new(int dim, int W, int L, int H, int n, const char *reward_type, int hm_type, int strategy);
void tapo_env_free(tapo_env *e);
void tapo_env_clear(tapo_env *e);
int tapo_env_add_new_block(tapo_env *e, const float *block, int *hm_out);
double tapo_env_calc_ratio(const tapo_env *e);
int tapo_env_error(const tapo_env *e);
const int *tapo_env_heightmap(const tapo_env *e);
void tapo_update_dynamic(const float *dynamic, const float *static_, const int64_t *ptr, int B, int rows, int S, int srows,
                         int n, int update_time, float *out);
void tapo_update_mask(const float *mask, const float *dynamic, const int64_t *ptr, int B, int rows, int S, int n, int R,
                      float *new_mask, float *chosen_mask);

typedef struct {
    int dim, W, L, H, T, n, hm_type, strategy, order_mode, B, b0, b1, status;
    const char *reward_type;
    const unsigned char *adj; const int *blocks; const int64_t *ptr_seq;
    float *reward_out; int *heightmap_out;
} roll_job;

static void *roll_worker(void *arg) {
    roll_job *j = (roll_job *)arg;
    const int dim = j->dim, T = j->T, n = j->n, R = dim == 2 ? 2 : 6, S = n * R, rows = 3 * n, srows = 1 + dim, B = j->B;
    const int cells = dim == 2 ? j->W : j->W * j->L;
    tapo_env *e = tapo_env_new(dim, j->W, j->L, j->H, T, j->reward_type, j->hm_type, j->strategy);
    float *st = (float *)malloc(sizeof(float) * srows * S), *da = (float *)malloc(sizeof(float) * rows * S), *db = (float *)malloc(sizeof(float) * rows * S);
    float *m = (float *)malloc(sizeof(float) * S), *cm = (float *)malloc(sizeof(float) * S), *nm = (float *)malloc(sizeof(float) * S);
    int *hm = (int *)calloc(2 * cells + 4, sizeof(int));
    for (int b = j->b0; b < j->b1; b++) {
        tapo_win *w = tapo_win_new(T, n, dim, j->adj + (size_t)b * 5 * T * T, j->blocks + (size_t)b * R * T * dim, j->order_mode);
        tapo_env_clear(e);
        int t = 0, one_step = 1;
        while (one_step && t < T) {
            if (tapo_win_convert_to_input(w, st, da)) { j->status = 3; break; }     /* rolling.py:593 */
            if (tapo_win_is_last_graph(w)) one_step = 0;                            /* rolling.py:598-599 */
            tapo_update_mask(NULL, da, NULL, 1, rows, S, n, R, nm, m);              /* rolling.py:325-335 */
            const int steps = one_step ? 1 : n;
            int64_t p = 0;
            for (int q = 0; q < steps && t < T; q++, t++) {
                p = j->ptr_seq[(size_t)t * B + b];
                tapo_update_dynamic(da, st, &p, 1, rows, S, srows, n, 3, db);       /* rolling.py:404-406 */
                tapo_update_mask(m, db, &p, 1, rows, S, n, R, nm, cm);              /* rolling.py:409-412 */
                memcpy(m, cm, sizeof(float) * S);
                float blk[3]; for (int d = 0; d < dim; d++) blk[d] = st[(size_t)(1 + d) * S + p];
                tapo_env_add_new_block(e, blk, hm);                                 /* rolling.py:436 */
                float *sw = da; da = db; db = sw;
            }
            int nodes[TAPW_MAXT]; tapo_win_nodes(w, nodes);
            tapo_win_remove_block(w, nodes[p % n]);                                 /* rolling.py:636-640 */
        }
        if ((tapo_env_error(e) || tapo_win_error(w)) && !j->status) j->status = tapo_env_error(e) ? tapo_env_error(e) : 4;
        if (j->reward_out) j->reward_out[b] = (float)tapo_env_calc_ratio(e);
        if (j->heightmap_out) memcpy(j->heightmap_out + (size_t)b * cells, tapo_env_heightmap(e), sizeof(int) * cells);
        tapo_win_free(w);
    }
    tapo_env_free(e); free(st); free(da); free(db); free(m); free(cm); free(nm); free(hm);
    return NULL;
}

int tapo_rolling_batch(int dim, int W, int L, int H, int T, int n, const char *reward_type, int hm_type, int strategy,
                       int order_mode, int B, const unsigned char *adj /*[B][5][T][T]*/, const int *blocks /*[B][R*T][dim]*/,
                       const int64_t *ptr_seq /*[T][B]*/, float *reward_out, int *heightmap_out, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > B) nthreads = B > 0 ? B : 1;
    if (nthreads > 1024) nthreads = 1024;
    roll_job *jobs = (roll_job *)calloc(nthreads, sizeof(roll_job));
    pthread_t *th = (pthread_t *)calloc(nthreads, sizeof(pthread_t));
    for (int t = 0; t < nthreads; t++) {
        roll_job *j = &jobs[t];
        j->dim = dim; j->W = W; j->L = L; j->H = H; j->T = T; j->n = n; j->hm_type = hm_type; j->strategy = strategy;
        j->order_mode = order_mode; j->B = B; j->reward_type = reward_type; j->adj = adj; j->blocks = blocks; j->ptr_seq = ptr_seq;
        j->reward_out = reward_out; j->heightmap_out = heightmap_out;
        j->b0 = (int)((long long)B * t / nthreads); j->b1 = (int)((long long)B * (t + 1) / nthreads);
        if (t > 0) pthread_create(&th[t], NULL, roll_worker, j);
    }
    roll_worker(&jobs[0]);
    int status = 0;
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
    for (int t = 0; t < nthreads; t++) if (jobs[t].status && !status) status = jobs[t].status;
    free(jobs); free(th);
    return status;
}
