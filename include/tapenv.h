/*
 * tapenv.h -- C ABI of the B200-native TAP packing-environment step.
 *
 * Drop-in boundary for ONE hot path of Juzhan/TAP-Net: the per-decode-step
 * environment transition that model.py's pointer-network decode loop performs
 * (model.py:376-458, :499-515).  The reference has no FFI layer; its "operator
 * API" is a set of plain Python callables and one class.  Every entry point below
 * cites the reference interface it replaces (file:line into the reference tree).
 * The Python binding a maintainer adds is a ctypes stub -- see INTEGRATION.md and
 * tap-net_b200/tapenv/_capi.py.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / C++ types.  All data pointers are
 *    DEVICE pointers (sm_100a, HBM) unless the name ends in _host.
 *  - the caller owns every buffer (torch allocates them); the library never
 *    allocates, frees or retains memory, has no global state, and is re-entrant.
 *  - kernels run on the CURRENT CUDA device of the calling thread; pointers must belong to it.
 *  - every function is stream-ordered on `stream` (a cudaStream_t passed as
 *    void*; NULL = legacy default stream) and never synchronises.
 *  - outputs never alias inputs (the reference clones: pack.py:318,323,370).
 *  - return value: TAPENV_OK (0) or a negative TAPENV_E* code; nothing is
 *    launched when an error is returned.
 *  - tensors are dense row-major with the reference's layouts:
 *      static   f32 [B, static_rows, S]   row 0 = block id, rows 1..dim = edge lengths   (pack.py:144-147,186)
 *      dynamic  f32 [B, dyn_rows, S]      3 bands of n rows: move | rot-small | rot-large (pack.py:195); 'simple'/'rot': the
 *                                         move band only; legacy 'rot-old': move band + ONE rotate-state row (pack.py:218-223)
 *      mask     f32 [B, S]                0/1                                               (model.py:297)
 *      ptr      i64 [B]                   chosen candidate column                           (model.py:365-371)
 *    S = n * rotate_types candidates, rotation-major (column j = r*n + i).
 *  - `dynamic` entries must be non-negative (the dataset writes 0/1, pack.py:101-223):
 *    the accessibility test `move_sum + small_sum*large_sum != 0` (pack.py:324-329) is
 *    evaluated as `any(move) | (any(small) & any(large))`.
 */
#ifndef TAPENV_H_
#define TAPENV_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TAPENV_VERSION 101

/* error codes */
#define TAPENV_OK 0
#define TAPENV_EINVAL (-1)    /* NULL pointer / bad size */
#define TAPENV_EENUM (-2)     /* unknown reward_type / packing_strategy / heightmap_type / input_type */
#define TAPENV_ELIMIT (-3)    /* shape outside the compiled limits (see tapenv_limits) */
#define TAPENV_ESHAPE (-4)    /* S != n*R, dyn_rows too small, ... */
#define TAPENV_ECUDA (-5)     /* the launch itself failed (cudaGetLastError) */
#define TAPENV_EUNSUPPORTED (-6) /* not served by this entry point: the fused step / episode / rolling entries with the legacy
                                    'rot-old' layout (the reference's own decode loop raises there: model.py:391-392 hands
                                    Container.add_new_block 1+dim values, tools.py:2060); the tensor operators serve it */

/* packing_strategy (tools.py:3607, :3679-3701) */
#define TAPENV_LB_GREEDY 0
#define TAPENV_MACS 1           /* 2D: heightmap + block history; 3D: voxel grid + interval lists in the state (tools.py:2751-3165) */
#define TAPENV_LB 2           /* the older corner-list strategy, tools.py:1602-1914; keeps a voxel grid in the state */
/* heightmap_type (tools.py:3716-3743) */
#define TAPENV_HM_FULL 0
#define TAPENV_HM_ZERO 1
#define TAPENV_HM_DIFF 2
/* reward_flags: the substring tests the reference performs on reward_type */
#define TAPENV_RF_HARD 1      /* reward_type.endswith('hard')   tools.py:2113 */
#define TAPENV_RF_P 2         /* 'P' in reward_type             tools.py:2135 */
#define TAPENV_RF_S 4         /* 'S' in reward_type             tools.py:2138 */
#define TAPENV_RF_MCS_IN 8    /* 'mcs' in reward_type           tools.py:2718 */
#define TAPENV_RF_MCS_START 16/* reward_type.startswith('mcs')  tools.py:2709 */
/* ratio_mode: which branch of Container.calc_ratio applies (tools.py:3908-3966) */
#define TAPENV_RATIO_C 0            /* 'comp'                      -> C/3            */
#define TAPENV_RATIO_CS 1           /* 'soft','hard'               -> C*S/3          */
#define TAPENV_RATIO_C_P 2          /* 'pyrm'                      -> (C+P)/3        */
#define TAPENV_RATIO_CP_S 3         /* 'pyrm-soft','mcs-hard',...  -> (C+P)*S/3      */
#define TAPENV_RATIO_SUM 4          /* every 'C+P*-*' type, '*-sum'-> (C+P+S)/3      */
#define TAPENV_RATIO_2C_SUM 5       /* '*-SUM'                     -> (2C+P+S)/3     */
#define TAPENV_RATIO_CPS 6          /* 'CPS'                       -> C*P*S/3        */
#define TAPENV_RATIO_CP_HALF 7      /* 'C+P-lb-soft'               -> (C+P)/2  (:3961) */

typedef struct tapenv_config {
    int32_t batch;          /* B: environments in this call                                      */
    int32_t blocks_num;     /* n: blocks per episode            (Container blocks_num, tools.py:3611) */
    int32_t dim;            /* 2 or 3                           (len(container_size))              */
    int32_t rotate_types;   /* R: dim! when allow_rot else 1    (pack.py:306-309)                  */
    int32_t width;          /* container_size[0]                                                   */
    int32_t length;         /* container_size[1] in 3D, 1 in 2D                                    */
    int32_t height;         /* container_size[-1]                                                  */
    int32_t strategy;       /* TAPENV_LB_GREEDY | TAPENV_MACS | TAPENV_LB (after the reward-type override, tools.py:3617-3620) */
    int32_t heightmap_type; /* TAPENV_HM_*                                                         */
    int32_t reward_flags;   /* TAPENV_RF_* bits                                                    */
    int32_t ratio_mode;     /* TAPENV_RATIO_*                                                      */
    int32_t static_rows;    /* rows of `static`: 1+dim ('mul-with': 2+dim)                         */
    int32_t dyn_rows;       /* rows of `dynamic`: 3n for 'bot'-like inputs, n for 'simple'/'rot', n+1 for 'rot-old' */
    int32_t update_time;    /* bands zeroed by update_dynamic: 3 or 1 (pack.py:349)                */
    int32_t capacity;       /* blocks one container can take: rows of positions/blocks/stable per environment.
                               tapenv_config_init sets it to blocks_num; rolling inference keeps ONE container for
                               total_blocks_num blocks while the network window stays at blocks_num
                               (rolling.py:702-703) -- set it, then re-check with tapenv_config_check.        */
} tapenv_config;

/* Byte offsets of the arrays inside the opaque per-batch state buffer. */
typedef struct tapenv_state_layout {
    size_t scalars;    /* i32 [B,4]  valid_size, empty_size, stable count, current_blocks_num (tools.py:3635-3653) */
    size_t heightmap;  /* i32 [B,W] or [B,W,L]                                        (tools.py:3630) */
    size_t positions;  /* i32 [B,capacity,dim]                                        (tools.py:3628) */
    size_t blocks;     /* i32 [B,capacity,dim]  blocks in arrival order               (tools.py:3631,3674) */
    size_t stable;     /* u8  [B,capacity]                                            (tools.py:3633) */
    size_t flags;      /* i32 [B] sticky anomaly bits (the reference would raise IndexError in each case):
                          1 = a stack grew above container height, 2 = more than `capacity` blocks were added,
                          4 = a pointer outside [0,S) was passed to the fused step */
    size_t voxels;     /* i16 [B,cells,H]  LB and MACS 3D: 0 empty, -1 empty under a block, k+1 block id   (tools.py:3629) */
    size_t lists;      /* i8  [B,nlists,max(capacity+2, width+4)]  LB: level_free_space x lists (tools.py:3649-3653); MACS 3D: the
                          per-(level,row) interval lists (tools.py:3644-3648); byte 0 = length */
    size_t pending;    /* f32 [B,4]  reserved (r01 handed the gathered block of LB / MACS 3D to a second kernel through it) */
    size_t total;      /* == tapenv_state_bytes() */
} tapenv_state_layout;

/* compiled limits */
typedef struct tapenv_limits {
    int32_t max_width_2d;      /* W  <= this for dim 2 */
    int32_t max_cells_3d;      /* W*L <= this for dim 3 */
    int32_t max_candidates;    /* S  <= this */
    int32_t max_blocks;        /* n  <= this */
} tapenv_limits;

int tapenv_version(void);
const char *tapenv_strerror(int code);
void tapenv_get_limits(tapenv_limits *out);

/*
 * Fill `cfg` from the reference's own (string-typed) options.  Replaces the
 * option handling spread over tools.Container.__init__ (tools.py:3611-3661),
 * pack.update_mask / update_dynamic (pack.py:285-309, :338-365) and
 * Container.calc_ratio (tools.py:3908-3966).  Unknown strings -> TAPENV_EENUM
 * (the reference prints "... OHHH" and then fails with NameError).
 *   container_size: int[dim];  allow_rot: 0/1
 */
int tapenv_config_init(tapenv_config *cfg, int32_t batch, int32_t blocks_num, int32_t dim,
                       int32_t allow_rot, const int32_t *container_size,
                       const char *reward_type, const char *packing_strategy,
                       const char *heightmap_type, const char *input_type);

/* Validate a config against shapes and compiled limits. */
int tapenv_config_check(const tapenv_config *cfg);

/* Size / layout of the state buffer for cfg->batch environments. */
size_t tapenv_state_bytes(const tapenv_config *cfg);
int tapenv_state_get_layout(const tapenv_config *cfg, tapenv_state_layout *out);

/* Elements per environment of the encoded heightmap that add_new_block returns
 * (tools.py:3716-3743): 2D full/zero W, 2D diff W-1, 3D full/zero W*L, 3D diff 2*W*L. */
int32_t tapenv_encoded_heightmap_len(const tapenv_config *cfg);

/*
 * K0 reset.  Replaces `[tools.Container(...) for _ in range(batch_size)]`
 * (model.py:294 -> tools.py:3611-3661; also clear_container :3858-3885) and the
 * initial accessibility mask (model.py:297-307).
 *   dynamic may be NULL (state reset only); then the mask outputs are ignored.
 *   cur_mask_out f32 [B,S] : current_mask at t=0 ; mask_out f32 [B,S] : ones.
 */
int tapenv_reset(const tapenv_config *cfg, void *state, const float *dynamic,
                 float *cur_mask_out, float *mask_out, void *stream);

/* The initial accessibility mask alone (model.py:297-307 == rolling.py:325-335), without touching any
 * container state: what a rolling-window driver needs each time the window is refilled while ONE container
 * keeps filling up (rolling.py:702-703, :607).  cur_mask_out f32 [B,S]; mask_out f32 [B,S] = ones (may be NULL). */
int tapenv_initial_mask(const tapenv_config *cfg, const float *dynamic, float *cur_mask_out, float *mask_out,
                        void *stream);

/* K0 for a packed host format.  The dataset tensors are small integers (static, pack.py:144-147) and 0/1 flags
 * (dynamic, pack.py:101-223) stored as fp32; a loader may keep and upload them as u8 / bit rows (20x fewer PCIe
 * bytes per batch than trainer.py:189-192's `.cuda()` of the fp32 tensors) and expand them on the device:
 *   static_u8    u8  [B, static_rows, S]
 *   dynamic_bits u32 [B, tapenv_packed_words()]: bit (row*S + col) -- word (q >> 5), bit (q & 31) -- <=> dynamic[row,col] == 1
 * Writes the fp32 tensors (static_out [B,static_rows,S], dynamic_out [B,dyn_rows,S]), clears the containers when
 * state != NULL (tools.py:3611-3661) and emits the initial masks (model.py:297-307) -- tapenv_reset on packed input. */
int32_t tapenv_packed_words(const tapenv_config *cfg);
int tapenv_reset_packed(const tapenv_config *cfg, void *state, const uint8_t *static_u8, const uint32_t *dynamic_bits,
                        float *static_out, float *dynamic_out, float *cur_mask_out, float *mask_out, void *stream);

/* pack.update_dynamic(dynamic, static, chosen_idx, input_type, allow_rot) (pack.py:333-376).
 * Out-of-place; block id is read from static[:,0,ptr] (pack.py:347). */
int tapenv_update_dynamic(const tapenv_config *cfg, const float *dynamic, const float *static_,
                          const int64_t *ptr, float *dynamic_out, void *stream);

/* pack.update_mask(mask, dynamic, static, chosen_idx, input_type, allow_rot) (pack.py:276-331).
 * `dynamic` is the ALREADY UPDATED tensor; block id is ptr mod n (pack.py:314-316).
 * Returns (new_mask, chosen_mask) = (new_mask_out, chosen_mask_out). */
int tapenv_update_mask(const tapenv_config *cfg, const float *mask, const float *dynamic,
                       const int64_t *ptr, float *new_mask_out, float *chosen_mask_out, void *stream);

/* Batched tools.Container.add_new_block(block, is_rotate) (tools.py:3663-3744) for all B
 * environments: placement scan (tools.py:2027-2351 / :2456-2749), stability test
 * (tools.py:710-765, :839-868), state commit, heightmap encoding.
 *   blocks f32 [B,dim] -- the rows model.py:412 hands to the Python loop at :452-453
 *   dec_dynamic_out f32 [B, tapenv_encoded_heightmap_len()] -- what model.py:456-463 builds */
int tapenv_add_blocks(const tapenv_config *cfg, void *state, const float *blocks,
                      float *dec_dynamic_out, void *stream);

/* The fused hot path: update_dynamic + update_mask + gather of the chosen block
 * (model.py:404-406) + add_new_block, ONE launch per decode step.
 *   dec_static_out f32 [B, static_rows-1] = static[:,1:,ptr]  (may be NULL)
 *   all other arguments as above. */
int tapenv_step(const tapenv_config *cfg, void *state, const int64_t *ptr, const float *static_,
                const float *dynamic_in, const float *mask_in,
                float *dynamic_out, float *cur_mask_out, float *mask_out,
                float *dec_static_out, float *dec_dynamic_out, void *stream);

/* tapenv_step + Container.calc_ratio() of the state the step leaves behind: what model.py:499-515 computes right after the
 * LAST decode step, without a separate launch.  reward_out f32 [B] as tapenv_reward. */
int tapenv_step_reward(const tapenv_config *cfg, void *state, const int64_t *ptr, const float *static_,
                       const float *dynamic_in, const float *mask_in, float *dynamic_out, float *cur_mask_out,
                       float *mask_out, float *dec_static_out, float *dec_dynamic_out, float *reward_out, void *stream);

/* The deterministic (sum r, sum r^2, B) of an existing reward vector, and -- with comm != NULL -- its cross-GPU exchange
 * (see tapenv_reward_allreduce): the second half of tapenv_reward / tapenv_reward_allreduce for rewards that
 * tapenv_step_reward already produced.  partial_sums_out may be NULL when comm is given. */
struct tapenv_peer_comm;
int tapenv_reward_sums(const tapenv_config *cfg, const float *reward, double *partial_sums_out, double *total_sums_out,
                       const struct tapenv_peer_comm *comm, void *stream);

/* Container.calc_ratio() for every environment (tools.py:3887-3966; consumed at
 * model.py:499-515): reward_out f32 [B] = (float) ratio (fp64 -> fp32, NOT negated).
 * partial_sums_out (may be NULL) f64 [3] = (sum r, sum r^2, B) reduced in a fixed
 * order -- the per-rank operand of the end-of-episode all-reduce (trainer.py:216-225). */
int tapenv_reward(const tapenv_config *cfg, const void *state, float *reward_out,
                  double *partial_sums_out, void *stream);

/* K6 fused with its collective: Container.calc_ratio for every environment, the deterministic per-rank
 * (sum r, sum r^2, count) AND the cross-GPU reduction of those triples in the SAME launch sequence, over
 * NVLink peer memory instead of a separate NCCL call (the statistics feed the critic baseline,
 * trainer.py:216-225; the reference is single-process and has no collective to cite).
 *   comm: every rank allocates tapenv_comm_bytes() of zero-initialised, peer-mapped device memory (e.g.
 *         torch symmetric memory) and passes ALL ranks' base pointers, own rank included, in rank order.
 *   Each call posts this rank's triple into every peer's buffer (st + fence + sequence flag), waits for
 *   the triples of all ranks for the same call number and sums them IN RANK ORDER: total_sums_out f64 [3] is
 *   bit-identical on every rank.  All ranks must issue the same sequence of calls.  Stream-ordered, capturable
 *   in a CUDA graph (the call counter lives in device memory). */
#define TAPENV_COMM_MAX_RANKS 8
typedef struct tapenv_peer_comm {
    int32_t world, rank;
    void *peer[TAPENV_COMM_MAX_RANKS];
} tapenv_peer_comm;
size_t tapenv_comm_bytes(void);
/* Byte offset, inside every rank's exchange buffer, of two u64 words: a STICKY status (0 = healthy; bit 0 = a call gave up
 * polling for a peer -- its total_sums_out are NaN) followed by the number of the first call that failed.  A call polls
 * for at most TAPENV_EXCHANGE_TIMEOUT_MS (environment, default 10000 ms, measured on %globaltimer): the bound must exceed the
 * worst rank skew of the job (checkpointing / logging on one rank, a loader stall, first-iteration warm-up); it exists so
 * that a peer which never calls cannot hang the GPU.  After a timeout the ranks' call counters may disagree: treat the
 * status as fatal for the exchange (re-create the buffers, or fall back to an NCCL all-gather of the triples). */
size_t tapenv_comm_status_offset(void);
int tapenv_reward_allreduce(const tapenv_config *cfg, const void *state, float *reward_out, double *partial_sums_out,
                            double *total_sums_out, const tapenv_peer_comm *comm, void *stream);

/* Whole-episode entry (tools.calc_positions_lb_greedy tools.py:2393-2449,
 * calc_positions_mcs :3213-3315, pack.reward pack.py:378-473): run `steps` decode steps
 * for a known pointer sequence in ONE launch, without materialising the
 * intermediate dynamic tensors.
 *   ptr_seq i64 [steps,B].  Resets the state first.  Outputs (each may be NULL):
 *   reward_out f32 [B] (calc_ratio), cur_mask_out/mask_out f32 [B,S] after the last
 *   step, dec_dynamic_out f32 [B,enc] last encoded heightmap. */
int tapenv_episode(const tapenv_config *cfg, void *state, const float *static_, const float *dynamic,
                   const int64_t *ptr_seq, int32_t steps, float *reward_out,
                   float *cur_mask_out, float *mask_out, float *dec_dynamic_out, void *stream);


/* ---- two-container inputs: input_type 'mul' / 'mul-with' (model.py:286-292, :396-447, :503-507) ----------------------
 * `static` has 2+dim rows: block id, edge lengths, target container id (0 = A, 1 = B; pack.py:212-216).  Every environment
 * owns TWO containers (two state buffers of the same config).  The tensor side equals 'bot' (pack.py:300-302, :354-357), so
 * tapenv_update_dynamic / tapenv_update_mask / tapenv_reset apply unchanged (reset each state).
 *   tapenv_step_mul: fused decode step; the chosen block goes into the container its candidate names, the decoder input
 *     is both heightmaps: dec_dynamic_out f32 [B, 2, enc] = cat(A, B) (model.py:421-447).  dec_static_out f32
 *     [B, dec_static_rows]: dec_static_rows = dim for 'mul' (static[:,1:-1,:]) or dim+1 for 'mul-with' (static[:,1:,:]),
 *     model.py:388-394.  A target id other than 0/1 sets flag 4 in both states (the reference fails at :437).
 *   tapenv_add_blocks_mul: the unfused placement (blocks f32 [B,dim], target_ids f32 [B]).
 *   tapenv_reward_mul: reward_out f32 [B] = (calc_ratio(A) + calc_ratio(B)) / 2, accumulated in fp32 as model.py:503-507 does. */
int tapenv_step_mul(const tapenv_config *cfg, void *state_a, void *state_b, const int64_t *ptr, const float *static_,
                    const float *dynamic_in, const float *mask_in, float *dynamic_out, float *cur_mask_out,
                    float *mask_out, float *dec_static_out, int32_t dec_static_rows, float *dec_dynamic_out, void *stream);
int tapenv_add_blocks_mul(const tapenv_config *cfg, void *state_a, void *state_b, const float *blocks,
                          const float *target_ids, float *dec_dynamic_out, void *stream);
int tapenv_reward_mul(const tapenv_config *cfg, const void *state_a, const void *state_b, float *reward_out, void *stream);

/* ---- rolling window --------------------------------------------------------------------------------------------
 * generate.InitialContainer (generate.py:1589-1825) for B instances: the window of `window` (= child_graph_size) nodes
 * the network sees while ONE container takes all `total_blocks` blocks (rolling.py:575-640, :702-703).
 *   pred    u64 [B,5,T]: bit u of pred[b,g,v] <=> edge u -> v in G_move / G_left / G_right / G_forward / G_backward
 *           (generate.py:1621-1664: deps_g[u,v] == True).  T = total_blocks <= 64.
 *   blocks  i32 [B,R*T,dim]: self.blocks, rotation-major (rolling.py:483-485, generate.py:1615).
 *   wstate  tapenv_window_state_bytes() bytes, 64 per instance: the InitialContainer fields that change
 *           (gm's node set, after_nodes_list, sub_graph_nodes).  u32 words per instance: [0:2) removed-from-gm mask,
 *           [2:4) after_nodes_list mask, [4:12) sub_graph_nodes as bytes (0xff = unused), [12] len, [13] sticky flags:
 *           1 = no in-degree-0 node (the reference would loop forever), 2 = window not full at convert_to_input (the
 *           reference raises in np.concatenate), 4 = pointer outside the window.
 * node_order: how the induced sub-graphs enumerate their nodes (this decides which row/column of `dynamic` a dependency
 *   lands in).  REFERENCE reproduces what the reference does under networkx >= 2 / CPython 3: networkx's subgraph view
 *   iterates the Python *set* of the window nodes when 2*window < total_blocks (coreviews.FilterAtlas.__iter__), so the
 *   order is CPython's set layout (Objects/setobject.c) -- while `static` uses the sorted node list.  SORTED uses the
 *   ascending order everywhere (the evident intent). */
#define TAPENV_WINDOW_ORDER_REFERENCE 0
#define TAPENV_WINDOW_ORDER_SORTED 1
typedef struct tapenv_window_config {
    int32_t batch;          /* B */
    int32_t total_blocks;   /* T: InitialContainer blocks_num        (generate.py:1590) */
    int32_t window;         /* n: child_graph_size                   (generate.py:1590) */
    int32_t dim;            /* 2 or 3 */
    int32_t rotate_types;   /* factorial(dim)                        (generate.py:1610) */
    int32_t node_order;     /* TAPENV_WINDOW_ORDER_* */
    int32_t blocks_are_rotations; /* != 0: blocks[b, r*T+i, d] == blocks[b, i, perm_r[d]] with perm_r the r-th of
                               itertools.permutations(range(dim)) -- how generate.generate_blocks writes every dataset
                               (rolling.py:60-63).  The kernel then reads T*dim instead of R*T*dim values per instance.
                               Set it only after checking the array; 0 is always correct. */
} tapenv_window_config;
size_t tapenv_window_state_bytes(const tapenv_window_config *wcfg);
/* InitialContainer.__init__ tail (generate.py:1666-1673): nothing removed, every node in after_nodes_list, empty window. */
int tapenv_window_reset(const tapenv_window_config *wcfg, void *wstate, void *stream);
/* [remove_block(sub_graph_nodes[prev_ptr mod n]) when prev_ptr != NULL (rolling.py:636-640, generate.py:1810-1822)] +
 * convert_to_input() (generate.py:1770-1808, 'bot' input) + is_last_graph() (:1824) for every instance.
 *   static_out f32 [B,1+dim,S], dynamic_out f32 [B,3n,S];
 *   cur_mask_out / mask_out f32 [B,S] (may be NULL): the initial masks rolling.DRL.forward derives (rolling.py:325-335);
 *   nodes_out i32 [B,n] (may be NULL): sub_graph_nodes (sorted; -1 = unused);
 *   remaining_out i32 [B] (may be NULL): len(after_nodes_list); is_last_graph() <=> 0. */
int tapenv_window_next(const tapenv_window_config *wcfg, void *wstate, const uint64_t *pred, const int32_t *blocks,
                       const int64_t *prev_ptr, float *static_out, float *dynamic_out, float *cur_mask_out,
                       float *mask_out, int32_t *nodes_out, int32_t *remaining_out, void *stream);
/* One rolling decode step in ONE launch: the block the pointer selects in the CURRENT window (rolling.py:417-431) goes
 * into the container (add_new_block, :436), then tapenv_window_next(prev_ptr = ptr).  update_dynamic / update_mask are
 * not evaluated: with one_step=True their results never leave rolling.DRL.forward (rolling.py:404-412, :449-453).
 *   dec_static_out f32 [B,dim] (may be NULL), dec_dynamic_out f32 [B,enc] (may be NULL); other outputs as above. */
int tapenv_rolling_step(const tapenv_config *cfg, void *state, const tapenv_window_config *wcfg, void *wstate,
                        const uint64_t *pred, const int32_t *blocks, const int64_t *ptr, float *dec_static_out,
                        float *dec_dynamic_out, float *static_out, float *dynamic_out, float *cur_mask_out,
                        float *mask_out, int32_t *nodes_out, int32_t *remaining_out, void *stream);


#ifdef __cplusplus
}
#endif
#endif /* TAPENV_H_ */
