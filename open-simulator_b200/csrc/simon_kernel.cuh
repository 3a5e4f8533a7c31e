// simon_kernel.cuh — persistent thread-block-cluster placement kernel (sm_100a).
//
// One scenario = one thread-block cluster (up to 16 CTAs, DSMEM).  Every node of the scenario is owned by
// one thread slot for the whole kernel; its static columns, its dynamic NodeInfo aggregates and the
// per-class cached values live in that CTA's shared memory, so a placement decision touches no global
// memory on its critical path: filter -> score -> cluster-wide reductions over DSMEM -> argmax -> commit.
//
// Semantics follow oracle/simon_oracle.c line by line (which cites the reference file:line of every
// plugin); the data layout is include/simon_gpu.h.  Integer work is exact; the five float64 plugin
// formulas use IEEE double with contraction disabled (-fmad=false) and truncation by cast.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/simon_gpu.h"

namespace cg = cooperative_groups;

#define SK_MAX_ENT 32       // list entries (ports + constraints + terms) cached per node for one class
#define SK_MAX_HARD 8
#define SK_MAX_SOFT 8
#define SK_MAX_CS 16
#define SK_NV 16            // values per cluster all-reduce (8 scalars + 8 domain-bitmask words)
#define SK_MAXW 8            // domain-bitmask words available per decision
#define SK_PAYLOAD 10       // winner payload words: T domains + flags

enum { EK_PORT = 0, EK_HARD, EK_SOFT, EK_AFF, EK_ANTI, EK_EXIST, EK_SCORE };

struct SkScenario {
    const uint32_t *order;     // [n_active] scenario order -> node index (nullptr: identity)
    const int32_t *rank_of;    // [N] node index -> scenario rank or -1 (nullptr: identity)
    uint32_t n_active;
    uint32_t pad;
    // dynamic state, node-indexed
    int64_t *req_mcpu, *req_mem, *req_eph, *nz_mcpu, *nz_mem, *req_scalar, *gpu_used;
    int32_t *num_pods;
    int32_t *cnt;              // counters, all domains
    int32_t *cnt_total;        // [n_counters]
    // scratch tables [SK_MAX_SOFT or SK_MAX_HARD][max_dom]
    int32_t *tp, *fcount, *size;   // size: [SK_MAX_SOFT]
    uint8_t *hard_reg;
    // outputs
    int32_t *out_node;         // [P]
    int64_t *out_score;        // [P] or nullptr
    uint32_t *fail_counts;     // [max_fail][SIMON_N_FAIL_CODES]
    uint32_t *fail_pod;        // [max_fail]
    uint32_t *n_fail;          // [1]
    uint32_t *n_sched;         // [1]
    unsigned long long *clk;   // [2] start/stop globaltimer of this cluster (optional)
};

struct SkParams {
    // snapshot
    uint32_t N, K, WL, WT, T, NC, n_log, max_dom;
    const uint32_t *topo_ndom;
    const int64_t *alloc_mcpu, *alloc_mem, *alloc_eph, *alloc_scalar;
    const int32_t *alloc_pods;
    const uint32_t *node_flags;
    const uint64_t *label_bits, *taint_hard, *taint_soft;
    const int32_t *topo_dom, *node_class, *gpu_count;
    const int64_t *gpu_dev_mem, *gpu_total_mem;
    const double *log_table;
    // pod set
    uint32_t n_classes, n_pods, n_counters, max_blob_words;
    const uint64_t *class_off;
    const int64_t *class_blob;
    const int32_t *pod_class, *pod_fixed, *pod_guard;
    const uint64_t *cnt_off;
    const int64_t *simon_raw;
    const int32_t *extra_score;
    // launch
    uint32_t first, count, max_fail, npt, emax, record_scores;
    const SkScenario *scen;    // [gridDim.x / cluster size]
};

// ---------------------------------------------------------------------------------------------------------
// shared memory carve-up (per CTA).  L = npt * blockDim.x node slots.
struct SkSmem {
    int64_t *alloc_mcpu, *alloc_mem, *alloc_eph, *req_mcpu, *req_mem, *req_eph, *nz_mcpu, *nz_mem, *raw_simon;
    int32_t *alloc_pods, *num_pods, *node_g, *raw_na, *raw_tt, *extra, *raw_pts, *raw_ipa;
    int32_t *dom;        // [T][L]
    int32_t *val;        // [emax][L]
    uint8_t *st_code, *nflags, *regbits;
    int64_t *blob;       // [max_blob_words]
    long long *inbox;    // [2][SK_MAX_CS][SK_NV + SK_PAYLOAD]
    long long *wpart;    // [32][SK_NV + SK_PAYLOAD]
    int32_t *ent;        // [7][SK_MAX_ENT]: kind, k, t, a, b, inc, bitmask word offset (-1: table method)
};

#define NF_SEL_OK 1
#define NF_IGNORED 2
#define NF_HARDKEYS 4
#define NF_COUNTED 8
#define NF_FEASIBLE 16
#define NF_VALID 32

__host__ __device__ inline size_t sk_align(size_t x) { return (x + 15) & ~(size_t)15; }

__host__ __device__ inline size_t sk_smem_bytes(uint32_t L, uint32_t T, uint32_t emax, uint32_t blob_words) {
    size_t b = 0;
    b += 9 * sk_align(8ull * L);
    b += 8 * sk_align(4ull * L);
    b += sk_align(4ull * T * L);
    b += sk_align(4ull * (emax ? emax : 1) * L);
    b += 3 * sk_align(L);
    b += sk_align(8ull * blob_words);
    b += sk_align(8ull * 2 * SK_MAX_CS * (SK_NV + SK_PAYLOAD));
    b += sk_align(8ull * 32 * (SK_NV + SK_PAYLOAD));
    b += sk_align(4ull * 7 * SK_MAX_ENT);
    return b + 64;
}

__device__ inline void sk_carve(SkSmem &S, unsigned char *base, uint32_t L, uint32_t T, uint32_t emax, uint32_t blob_words) {
    unsigned char *p = base;
    auto take = [&](size_t bytes) { unsigned char *q = p; p += sk_align(bytes); return q; };
    S.alloc_mcpu = (int64_t *)take(8ull * L); S.alloc_mem = (int64_t *)take(8ull * L); S.alloc_eph = (int64_t *)take(8ull * L);
    S.req_mcpu = (int64_t *)take(8ull * L); S.req_mem = (int64_t *)take(8ull * L); S.req_eph = (int64_t *)take(8ull * L);
    S.nz_mcpu = (int64_t *)take(8ull * L); S.nz_mem = (int64_t *)take(8ull * L); S.raw_simon = (int64_t *)take(8ull * L);
    S.alloc_pods = (int32_t *)take(4ull * L); S.num_pods = (int32_t *)take(4ull * L); S.node_g = (int32_t *)take(4ull * L);
    S.raw_na = (int32_t *)take(4ull * L); S.raw_tt = (int32_t *)take(4ull * L); S.extra = (int32_t *)take(4ull * L);
    S.raw_pts = (int32_t *)take(4ull * L); S.raw_ipa = (int32_t *)take(4ull * L);
    S.dom = (int32_t *)take(4ull * T * L);
    S.val = (int32_t *)take(4ull * (emax ? emax : 1) * L);
    S.st_code = (uint8_t *)take(L); S.nflags = (uint8_t *)take(L); S.regbits = (uint8_t *)take(L);
    S.blob = (int64_t *)take(8ull * blob_words);
    S.inbox = (long long *)take(8ull * 2 * SK_MAX_CS * (SK_NV + SK_PAYLOAD));
    S.wpart = (long long *)take(8ull * 32 * (SK_NV + SK_PAYLOAD));
    S.ent = (int32_t *)take(4ull * 7 * SK_MAX_ENT);
}

// ---------------------------------------------------------------------------------------------------------
// cluster-wide all-reduce of NVAL int64 values (+ optional payload that follows value 0 when value 0 is a max)
// ops: 0 sum, 1 min, 2 max, 3 bitwise or.
struct SkRed {
    SkSmem *S;
    cg::cluster_group *cluster;
    uint32_t crank, CS, phase;
};

__device__ inline long long sk_identity(int op) { return (op == 0 || op == 3) ? 0 : (op == 1 ? INT64_MAX : INT64_MIN); }

__device__ inline long long sk_combine(long long a, long long b, int op) {
    return op == 0 ? a + b : (op == 1 ? (a < b ? a : b) : (op == 2 ? (a > b ? a : b) : (a | b)));
}

template <int NVAL, bool PAYLOAD>
__device__ inline void sk_allreduce(SkRed &R, long long (&v)[NVAL], const int (&op)[NVAL], long long (&pl)[SK_PAYLOAD]) {
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    constexpr int W = SK_NV + SK_PAYLOAD;
    // 1) warp level
#pragma unroll
    for (int i = 0; i < NVAL; i++) {
        long long x = v[i];
        if (PAYLOAD && i == 0) {
            // arg-max: find the winning lane, broadcast its payload
            long long m = x;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { long long y = __shfl_xor_sync(0xffffffffu, m, o); m = y > m ? y : m; }
            unsigned who = __ffs(__ballot_sync(0xffffffffu, x == m)) - 1;
#pragma unroll
            for (int j = 0; j < SK_PAYLOAD; j++) pl[j] = __shfl_sync(0xffffffffu, pl[j], who);
            v[i] = m;
        } else {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x = sk_combine(x, __shfl_xor_sync(0xffffffffu, x, o), op[i]);
            v[i] = x;
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NVAL; i++) R.S->wpart[warp * W + i] = v[i];
        if (PAYLOAD) {
#pragma unroll
            for (int j = 0; j < SK_PAYLOAD; j++) R.S->wpart[warp * W + SK_NV + j] = pl[j];
        }
    }
    __syncthreads();
    // 2) CTA level by warp 0, then publish to every CTA of the cluster through DSMEM
    const uint32_t buf = R.phase & 1;
    if (warp == 0) {
        long long cv[NVAL];
        long long cpl[SK_PAYLOAD];
#pragma unroll
        for (int i = 0; i < NVAL; i++) {
            long long x = lane < nwarp ? R.S->wpart[lane * W + i] : sk_identity(op[i]);
            if (PAYLOAD && i == 0) {
                long long m = x;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { long long y = __shfl_xor_sync(0xffffffffu, m, o); m = y > m ? y : m; }
                unsigned who = __ffs(__ballot_sync(0xffffffffu, x == m)) - 1;
#pragma unroll
                for (int j = 0; j < SK_PAYLOAD; j++) {
                    long long p = lane < nwarp ? R.S->wpart[lane * W + SK_NV + j] : 0;
                    cpl[j] = __shfl_sync(0xffffffffu, p, who);
                }
                cv[i] = m;
            } else {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) x = sk_combine(x, __shfl_xor_sync(0xffffffffu, x, o), op[i]);
                cv[i] = x;
            }
        }
        if (lane < R.CS) {
            long long *dst = R.cluster->map_shared_rank(R.S->inbox, lane) + (buf * SK_MAX_CS + R.crank) * W;
#pragma unroll
            for (int i = 0; i < NVAL; i++) dst[i] = cv[i];
            if (PAYLOAD) {
#pragma unroll
                for (int j = 0; j < SK_PAYLOAD; j++) dst[SK_NV + j] = cpl[j];
            }
        }
    }
    R.cluster->sync();
    // 3) every warp folds the CS partials (lane = source CTA)
#pragma unroll
    for (int i = 0; i < NVAL; i++) {
        long long x = lane < R.CS ? R.S->inbox[(buf * SK_MAX_CS + lane) * W + i] : sk_identity(op[i]);
        if (PAYLOAD && i == 0) {
            long long m = x;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { long long y = __shfl_xor_sync(0xffffffffu, m, o); m = y > m ? y : m; }
            unsigned who = __ffs(__ballot_sync(0xffffffffu, x == m)) - 1;
#pragma unroll
            for (int j = 0; j < SK_PAYLOAD; j++) {
                long long p = lane < R.CS ? R.S->inbox[(buf * SK_MAX_CS + lane) * W + SK_NV + j] : 0;
                pl[j] = __shfl_sync(0xffffffffu, p, who);
            }
            v[i] = m;
        } else {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x = sk_combine(x, __shfl_xor_sync(0xffffffffu, x, o), op[i]);
            v[i] = x;
        }
    }
    R.phase++;
}
