// simon_kernel.cu — the placement kernel (see simon_kernel.cuh for the design summary).
#include "simon_kernel.cuh"

// ---- small helpers --------------------------------------------------------------------------------------
__device__ __forceinline__ int32_t ldcg32(const int32_t *p) { return __ldcg(p); }

__device__ __forceinline__ int64_t f2i(double x) {
    // Go int64(float64): in range -> truncation; out of range/NaN -> MinInt64 (amd64 CVTTSD2SQ)
    if (!(x >= -9223372036854775808.0 && x < 9223372036854775808.0)) return INT64_MIN;
    return (int64_t)x;
}

__device__ __forceinline__ unsigned long long sk_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

struct ReqCtx {
    const uint64_t *label_bits;
    uint32_t N;
};

__device__ inline bool req_eval(const int64_t *&p, const ReqCtx &c, uint32_t g) {
    int64_t head = *p++;
    int op = (int)(head & 0xff);
    if (op == SIMON_REQ_NODE_IS || op == SIMON_REQ_NODE_ISNOT) {
        int64_t idx = *p++;
        return op == SIMON_REQ_NODE_IS ? ((int64_t)g == idx) : ((int64_t)g != idx);
    }
    int nw = (int)((head >> 8) & 0xff);
    bool any = false;
    for (int i = 0; i < nw; i++) {
        int64_t wi = *p++;
        uint64_t mask = (uint64_t)*p++;
        if (__ldg(c.label_bits + (uint64_t)wi * c.N + g) & mask) any = true;
    }
    return op == SIMON_REQ_ANY ? any : !any;
}

__device__ inline bool term_eval(const int64_t *&p, const ReqCtx &c, uint32_t g) {
    int64_t nreq = *p++;
    bool ok = true;
    for (int64_t i = 0; i < nreq; i++)
        if (!req_eval(p, c, g)) ok = false;
    return ok;
}

// PodMatchesNodeSelectorAndAffinityTerms (helper/node_affinity.go:27-71) over the compiled program
__device__ inline bool selection_ok(const int64_t *cw, const ReqCtx &c, uint32_t g) {
    const int64_t *p = cw + cw[SCW_OFF_SEL];
    bool ok = true;
    int64_t n_ns = *p++;
    for (int64_t q = 0; q < n_ns; q++)
        if (!req_eval(p, c, g)) ok = false;
    int64_t has_required = *p++, n_terms = *p++;
    bool any = false;
    for (int64_t q = 0; q < n_terms; q++)
        if (term_eval(p, c, g)) any = true;
    if (has_required && !any) ok = false;
    return ok;
}

// eligibility of node g for the spreading counts of class `sig` (scoring.go:140-148): passes the class's node
// selection and carries every soft topology key of that class
__device__ inline bool elig_eval(const SkParams &P, int64_t sig, const ReqCtx &c, uint32_t g) {
    const int64_t *cw = P.class_blob + P.class_off[sig];
    if (!selection_ok(cw, c, g)) return false;
    const int64_t *soft = cw + cw[SCW_OFF_PTS_SOFT];
    for (int64_t j = 0; j < cw[SCW_N_PTS_SOFT]; j++)
        if (P.topo_dom[(uint64_t)soft[5 * j + 1] * P.N + g] < 0) return false;
    return true;
}

// GpuNodeInfo.AllocateGpuId (pkg/type/open-gpu-share/cache/gpunodeinfo.go:232-290); returns slot count, 0 = none
__device__ inline int gpu_allocate(const SkParams &P, const SkScenario &SC, int64_t req_mem, int64_t req_num, uint32_t g, int *out) {
    int ndev = P.gpu_count ? P.gpu_count[g] : 0;
    if (req_mem <= 0 || req_num <= 0 || ndev <= 0) return 0;
    int64_t avail[SIMON_MAX_GPU_DEV];
    for (int d = 0; d < SIMON_MAX_GPU_DEV; d++)
        avail[d] = d < ndev ? P.gpu_dev_mem[g] - SC.gpu_used[(uint64_t)d * P.N + g] : -1;
    if (req_num == 1) {
        int cand = -1;
        int64_t cm = 0;
        for (int d = 0; d < ndev; d++)
            if (avail[d] >= req_mem && (cand < 0 || avail[d] < cm)) { cand = d; cm = avail[d]; }
        if (cand < 0) return 0;
        out[0] = cand;
        return 1;
    }
    int d = 0, got = 0;
    while (d < ndev && got < req_num) {
        if (avail[d] >= req_mem) { if (got < 64) out[got] = d; avail[d] -= req_mem; got++; }
        else d++;
    }
    return got == req_num ? got : 0;
}

#define ENT(row, e) S.ent[(row) * SK_MAX_ENT + (e)]
enum { ER_KIND = 0, ER_K, ER_T, ER_A, ER_B, ER_INC, ER_WOFF };

// ---------------------------------------------------------------------------------------------------------
extern "C" __global__ void __launch_bounds__(1024, 1) simon_place_kernel(const SkParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t CS = cluster.num_blocks();
    const uint32_t crank = cluster.block_rank();
    const uint32_t scen_id = blockIdx.x / CS;
    const SkScenario SC = P.scen[scen_id];
    const uint32_t TPB = blockDim.x, tid = threadIdx.x;
    const uint32_t CT = CS * TPB;
    const uint32_t gtid = crank * TPB + tid;
    const uint32_t NPT = P.npt, L = NPT * TPB;
    const uint32_t N = P.N, NA = SC.n_active, T = P.T, K = P.K, WT = P.WT;

    SkSmem S;
    sk_carve(S, smem_raw, L, T, P.emax, P.max_blob_words);
    SkRed R{&S, &cluster, crank, CS, 0};
    const ReqCtx RC{P.label_bits, N};
    const bool leader = (gtid == 0);

    if (leader && SC.clk) SC.clk[0] = sk_globaltimer();

    // ---- load node state into shared memory -----------------------------------------------------------
    for (uint32_t s = 0; s < NPT; s++) {
        uint32_t idx = s * TPB + tid;
        uint32_t r = s * CT + gtid;
        uint8_t nf = 0;
        int32_t g = -1;
        if (r < NA) {
            g = SC.order ? (int32_t)SC.order[r] : (int32_t)r;
            nf = NF_VALID;
            S.alloc_mcpu[idx] = P.alloc_mcpu[g]; S.alloc_mem[idx] = P.alloc_mem[g]; S.alloc_eph[idx] = P.alloc_eph[g];
            S.alloc_pods[idx] = P.alloc_pods[g];
            S.req_mcpu[idx] = SC.req_mcpu[g]; S.req_mem[idx] = SC.req_mem[g]; S.req_eph[idx] = SC.req_eph[g];
            S.nz_mcpu[idx] = SC.nz_mcpu[g]; S.nz_mem[idx] = SC.nz_mem[g]; S.num_pods[idx] = SC.num_pods[g];
            for (uint32_t t = 0; t < T; t++) S.dom[t * L + idx] = P.topo_dom[(uint64_t)t * N + g];
        }
        S.node_g[idx] = g;
        S.nflags[idx] = nf;
        S.st_code[idx] = 0;
        S.regbits[idx] = 0;
    }
    __syncthreads();

    int32_t cur_class = -1;
    uint32_t n_fail = 0, n_sched = 0;
    // per-class uniform state (identical in every thread)
    uint32_t E = 0, n_ports = 0, n_hard = 0, n_soft = 0, n_aff = 0, n_isc = 0;
    uint32_t e_hard = 0, e_soft = 0, e_aff = 0, e_anti = 0, e_exist = 0, e_isc = 0;
    int64_t aff_total = 0;
    uint32_t cflags = 0;
    bool any_table = false;

    const uint32_t end = P.first + P.count;
    uint32_t i = P.first;
    int32_t nx_cls = P.pod_class[i], nx_fixed = P.pod_fixed[i], nx_guard = P.pod_guard[i];
    while (i < end) {
        const int32_t cls = nx_cls, fixed = nx_fixed;
        const int64_t guard = nx_guard;
        if (i + 1 < end) { nx_cls = P.pod_class[i + 1]; nx_fixed = P.pod_fixed[i + 1]; nx_guard = P.pod_guard[i + 1]; }
        bool exists = true;
        if (guard == -2) exists = false;
        else if (guard >= 0) exists = SC.rank_of ? (SC.rank_of[guard] >= 0) : true;

        if (!exists || fixed != -1) {
            // ---- batch of pods that bypass scheduling (spec.nodeName preset) or do not exist in this scenario ----
            uint32_t j = i;
            while (j < end) {
                int32_t f2 = P.pod_fixed[j];
                int64_t g2 = P.pod_guard[j];
                bool ex2 = g2 == -2 ? false : (g2 >= 0 ? (SC.rank_of ? SC.rank_of[g2] >= 0 : true) : true);
                if (ex2 && f2 == -1) break;
                j++;
            }
            for (uint32_t q = i; q < j; q++) {
                int32_t c2 = P.pod_class[q], f2 = P.pod_fixed[q];
                const int64_t *cw2 = P.class_blob + P.class_off[c2];
                int64_t g2 = P.pod_guard[q];
                bool ex2 = g2 == -2 ? false : (g2 >= 0 ? (SC.rank_of ? SC.rank_of[g2] >= 0 : true) : true);
                int32_t res;
                if (!ex2) res = -3;
                else if (f2 < 0) res = -2;
                else {
                    int32_t r = SC.rank_of ? SC.rank_of[f2] : f2;
                    if (r < 0) res = -2;
                    else {
                        res = f2;
                        if ((uint32_t)r % CT == gtid) {       // I own that node: account the pod
                            uint32_t idx = ((uint32_t)r / CT) * TPB + tid;
                            S.req_mcpu[idx] += cw2[SCW_REQ_MCPU]; S.req_mem[idx] += cw2[SCW_REQ_MEM]; S.req_eph[idx] += cw2[SCW_REQ_EPH];
                            S.nz_mcpu[idx] += cw2[SCW_NZ_MCPU]; S.nz_mem[idx] += cw2[SCW_NZ_MEM]; S.num_pods[idx] += 1;
                            const int64_t *sc = cw2 + cw2[SCW_OFF_SCALARS];
                            for (uint32_t k = 0; k < K; k++) SC.req_scalar[(uint64_t)k * N + f2] += sc[k];
                            const int64_t *inc = cw2 + cw2[SCW_OFF_INC];
                            for (int64_t u = 0; u < cw2[SCW_N_INC]; u++) {
                                int64_t k = inc[3 * u], t = inc[3 * u + 1], sig = inc[3 * u + 2];
                                int32_t d = S.dom[t * L + idx];
                                if (d < 0) continue;
                                if (sig >= 0 && !elig_eval(P, sig, RC, (uint32_t)f2)) continue;
                                atomicAdd(&SC.cnt[P.cnt_off[k] + d], 1);
                                atomicAdd(&SC.cnt_total[k], 1);
                            }
                        }
                    }
                }
                if (leader) { SC.out_node[q] = res; if (SC.out_score) SC.out_score[q] = 0; }
            }
            cur_class = -1;
            i = j;
            if (i < end) { nx_cls = P.pod_class[i]; nx_fixed = P.pod_fixed[i]; nx_guard = P.pod_guard[i]; }
            continue;
        }

        // =================================================================================================
        // class change: stage the class record, evaluate everything that is static per (class, node), and
        // (re)load the cached counter values of this class's constraint/term entries
        // =================================================================================================
        if (cls != cur_class) {
            // the previous commit's counter updates (global atomics by the owner thread) must be visible
            __threadfence();
            cluster.sync();
            const int64_t *gw = P.class_blob + P.class_off[cls];
            const uint32_t words = (uint32_t)(P.class_off[cls + 1] - P.class_off[cls]);
            for (uint32_t w = tid; w < words; w += TPB) S.blob[w] = gw[w];
            __syncthreads();
            const int64_t *cw = S.blob;
            cflags = (uint32_t)cw[SCW_FLAGS];
            n_ports = (uint32_t)cw[SCW_N_PORTS]; n_hard = (uint32_t)cw[SCW_N_PTS_HARD]; n_soft = (uint32_t)cw[SCW_N_PTS_SOFT];
            n_aff = (uint32_t)cw[SCW_N_IPA_AFF]; n_isc = (uint32_t)cw[SCW_N_IPA_SCORE];
            const uint32_t n_anti = (uint32_t)cw[SCW_N_IPA_ANTI], n_exist = (uint32_t)cw[SCW_N_IPA_EXIST];
            e_hard = n_ports; e_soft = e_hard + n_hard; e_aff = e_soft + n_soft; e_anti = e_aff + n_aff;
            e_exist = e_anti + n_anti; e_isc = e_exist + n_exist; E = e_isc + n_isc;
            // bitmask word budget for the "number of distinct domains among feasible nodes" of soft constraints
            any_table = false;
            {
                uint32_t woff = 0;
                const int64_t *q = cw + cw[SCW_OFF_PTS_SOFT];
                for (uint32_t js = 0; js < n_soft; js++) {
                    if (q[5 * js + 3]) continue;
                    uint32_t nw = (P.topo_ndom[q[5 * js + 1]] + 63) / 64;
                    if (woff + nw <= SK_MAXW) woff += nw; else any_table = true;
                }
            }
            if (tid < E) {
                uint32_t e = tid;
                int32_t kind, k, t, a = 0, b = 0, woff = -1;
                if (e < e_hard) { kind = EK_PORT; k = (int32_t)cw[cw[SCW_OFF_PORTS] + e]; t = 0; }
                else if (e < e_soft) { const int64_t *q = cw + cw[SCW_OFF_PTS_HARD] + 4 * (e - e_hard); kind = EK_HARD; k = (int32_t)q[0]; t = (int32_t)q[1]; a = (int32_t)q[2]; b = (int32_t)q[3]; }
                else if (e < e_aff) {
                    const int64_t *q0 = cw + cw[SCW_OFF_PTS_SOFT];
                    const int64_t *q = q0 + 5 * (e - e_soft);
                    kind = EK_SOFT; t = (int32_t)q[1]; a = (int32_t)q[2]; b = (int32_t)q[3];
                    k = b ? (int32_t)q[0] : (int32_t)q[4];
                    if (!b) {
                        uint32_t wo = 0;
                        for (uint32_t js = 0; js < e - e_soft; js++)
                            if (!q0[5 * js + 3]) { uint32_t nw = (P.topo_ndom[q0[5 * js + 1]] + 63) / 64; if (wo + nw <= SK_MAXW) wo += nw; }
                        uint32_t nw = (P.topo_ndom[t] + 63) / 64;
                        woff = (wo + nw <= SK_MAXW) ? (int32_t)wo : -1;
                    }
                }
                else if (e < e_anti) { const int64_t *q = cw + cw[SCW_OFF_IPA_AFF] + 2 * (e - e_aff); kind = EK_AFF; k = (int32_t)q[0]; t = (int32_t)q[1]; }
                else if (e < e_exist) { const int64_t *q = cw + cw[SCW_OFF_IPA_ANTI] + 2 * (e - e_anti); kind = EK_ANTI; k = (int32_t)q[0]; t = (int32_t)q[1]; }
                else if (e < e_isc) { const int64_t *q = cw + cw[SCW_OFF_IPA_EXIST] + 2 * (e - e_exist); kind = EK_EXIST; k = (int32_t)q[0]; t = (int32_t)q[1]; }
                else { const int64_t *q = cw + cw[SCW_OFF_IPA_SCORE] + 3 * (e - e_isc); kind = EK_SCORE; k = (int32_t)q[0]; t = (int32_t)q[1]; a = (int32_t)q[2]; }
                int32_t inc = 0;
                int32_t tk = (kind == EK_SOFT && b) ? 0 : t;     // hostname constraints count on the node-level counter
                const int64_t *il = cw + cw[SCW_OFF_INC];
                for (int64_t u = 0; u < cw[SCW_N_INC]; u++)
                    if (il[3 * u] == k && il[3 * u + 1] == tk) inc = 1;
                ENT(ER_KIND, e) = kind; ENT(ER_K, e) = k; ENT(ER_T, e) = t; ENT(ER_A, e) = a; ENT(ER_B, e) = b;
                ENT(ER_INC, e) = inc; ENT(ER_WOFF, e) = woff;
            }
            __syncthreads();
            // node-static evaluation + cached counter values
            const int64_t *tol = cw + cw[SCW_OFF_TOL];
            const int64_t *simon_row = P.simon_raw + (uint64_t)cw[SCW_STATIC_ROW] * P.NC;
            const int32_t *extra = cw[SCW_EXTRA_ROW] >= 0 ? P.extra_score + (uint64_t)cw[SCW_EXTRA_ROW] * N : nullptr;
            for (uint32_t s = 0; s < NPT; s++) {
                uint32_t idx = s * TPB + tid;
                uint8_t nf = S.nflags[idx] & NF_VALID;
                if (!nf) continue;
                uint32_t g = (uint32_t)S.node_g[idx];
                bool ok = selection_ok(cw, RC, g);
                if (ok) nf |= NF_SEL_OK;
                uint8_t code = 0;
                if ((P.node_flags[g] & SIMON_NODE_UNSCHEDULABLE) && !(cflags & SIMON_CLS_TOL_UNSCHED)) code = 1;
                if (!code && cw[SCW_NODE_NAME] != -1 && cw[SCW_NODE_NAME] != (int64_t)g) code = 2;
                if (!code)
                    for (uint32_t w = 0; w < WT; w++)
                        if (P.taint_hard[(uint64_t)w * N + g] & ~(uint64_t)tol[w]) code = 3;
                if (!code && !ok) code = 4;
                S.st_code[idx] = code;
                const int64_t *p = cw + cw[SCW_OFF_PREF];
                int64_t na = 0;
                for (int64_t q = 0; q < cw[SCW_N_PREF]; q++) { int64_t w = *p++; if (term_eval(p, RC, g)) na += w; }
                S.raw_na[idx] = (int32_t)na;
                int tt = 0;
                for (uint32_t w = 0; w < WT; w++) tt += __popcll(P.taint_soft[(uint64_t)w * N + g] & ~(uint64_t)tol[WT + w]);
                S.raw_tt[idx] = tt;
                S.raw_simon[idx] = simon_row[P.node_class[g]];
                S.extra[idx] = extra ? extra[g] : 1000000;
                bool hk = true, ign = false;
                for (uint32_t jh = 0; jh < n_hard; jh++) if (S.dom[ENT(ER_T, e_hard + jh) * L + idx] < 0) hk = false;
                for (uint32_t js = 0; js < n_soft; js++) if (S.dom[ENT(ER_T, e_soft + js) * L + idx] < 0) ign = true;
                if (hk) nf |= NF_HARDKEYS;
                if (ign) nf |= NF_IGNORED;
                S.nflags[idx] = nf;
                for (uint32_t e = 0; e < E; e++) {
                    int32_t kind = ENT(ER_KIND, e), k = ENT(ER_K, e), t = ENT(ER_T, e);
                    int32_t d = (kind == EK_SOFT && ENT(ER_B, e)) ? (int32_t)g : S.dom[t * L + idx];
                    S.val[e * L + idx] = d >= 0 ? ldcg32(&SC.cnt[P.cnt_off[k] + d]) : 0;
                }
                S.regbits[idx] = 0;
            }
            if (n_hard > 0 || any_table) {
                // rare: DoNotSchedule constraints need the set of registered domains (filtering.go:221-243);
                // soft constraints over very large topologies fall back to per-domain tables
                for (uint32_t jh = 0; jh < n_hard; jh++) {
                    uint32_t nd = P.topo_ndom[ENT(ER_T, e_hard + jh)];
                    for (uint32_t d = gtid; d < nd; d += CT) SC.hard_reg[(uint64_t)jh * P.max_dom + d] = 0;
                }
                for (uint32_t js = 0; js < n_soft; js++) {
                    if (ENT(ER_B, e_soft + js) || ENT(ER_WOFF, e_soft + js) >= 0) continue;
                    uint32_t nd = P.topo_ndom[ENT(ER_T, e_soft + js)];
                    for (uint32_t d = gtid; d < nd; d += CT) SC.fcount[(uint64_t)js * P.max_dom + d] = 0;
                    if (gtid == 0) SC.size[js] = 0;
                }
                __threadfence();
                cluster.sync();
                for (uint32_t s = 0; s < NPT; s++) {
                    uint32_t idx = s * TPB + tid;
                    uint8_t nf = S.nflags[idx];
                    if ((nf & (NF_VALID | NF_SEL_OK | NF_HARDKEYS)) != (NF_VALID | NF_SEL_OK | NF_HARDKEYS)) continue;
                    for (uint32_t jh = 0; jh < n_hard; jh++)
                        SC.hard_reg[(uint64_t)jh * P.max_dom + S.dom[ENT(ER_T, e_hard + jh) * L + idx]] = 1;
                }
                __threadfence();
                cluster.sync();
                for (uint32_t s = 0; s < NPT; s++) {
                    uint32_t idx = s * TPB + tid;
                    if (!(S.nflags[idx] & NF_VALID)) continue;
                    uint8_t rb = 0;
                    for (uint32_t jh = 0; jh < n_hard; jh++) {
                        int32_t d = S.dom[ENT(ER_T, e_hard + jh) * L + idx];
                        if (d >= 0 && __ldcg(&SC.hard_reg[(uint64_t)jh * P.max_dom + d])) rb |= (uint8_t)(1u << jh);
                    }
                    S.regbits[idx] = rb;
                }
            }
            aff_total = 0;
            for (uint32_t e = e_aff; e < e_anti; e++) aff_total += ldcg32(&SC.cnt_total[ENT(ER_K, e)]);
            cur_class = cls;
        }
        const int64_t *cw = S.blob;

        // =================================================================================================
        // one placement decision
        // =================================================================================================
        long long pl[SK_PAYLOAD];
        // ---- R1: global minimum of every hard spread constraint over its registered domains ----
        int32_t hard_min[SK_MAX_HARD];
        if (n_hard > 0) {
            long long hv[SK_MAX_HARD];
            const int hop[SK_MAX_HARD] = {1, 1, 1, 1, 1, 1, 1, 1};
#pragma unroll
            for (int q = 0; q < SK_MAX_HARD; q++) hv[q] = INT32_MAX;
            for (uint32_t s = 0; s < NPT; s++) {
                uint32_t idx = s * TPB + tid;
                uint8_t nf = S.nflags[idx];
                if ((nf & (NF_VALID | NF_SEL_OK | NF_HARDKEYS)) != (NF_VALID | NF_SEL_OK | NF_HARDKEYS)) continue;
#pragma unroll
                for (int q = 0; q < SK_MAX_HARD; q++)
                    if ((uint32_t)q < n_hard) { long long v = S.val[(e_hard + q) * L + idx]; hv[q] = v < hv[q] ? v : hv[q]; }
            }
            sk_allreduce<SK_MAX_HARD, false>(R, hv, hop, pl);
#pragma unroll
            for (int q = 0; q < SK_MAX_HARD; q++) hard_min[q] = (int32_t)hv[q];
        }

        // ---- P1: filters on cached state; raw scores that do not depend on other nodes ----
        const int64_t *sc_req = cw + cw[SCW_OFF_SCALARS];
        auto filter_node = [&](uint32_t idx) -> uint32_t {
            uint32_t reasons = 0;
            if (S.st_code[idx]) return 1u << SFC_STATIC;
            for (uint32_t e = 0; e < n_ports; e++)
                if (S.val[e * L + idx] > 0) return 1u << SFC_PORTS;
            if (S.num_pods[idx] + 1 > S.alloc_pods[idx]) reasons |= 1u << SFC_TOO_MANY_PODS;
            if (cflags & SIMON_CLS_HAS_REQUEST) {
                if (S.alloc_mcpu[idx] < cw[SCW_REQ_MCPU] + S.req_mcpu[idx]) reasons |= 1u << SFC_CPU;
                if (S.alloc_mem[idx] < cw[SCW_REQ_MEM] + S.req_mem[idx]) reasons |= 1u << SFC_MEM;
                if (S.alloc_eph[idx] < cw[SCW_REQ_EPH] + S.req_eph[idx]) reasons |= 1u << SFC_EPH;
                if (K) {
                    uint32_t g = (uint32_t)S.node_g[idx];
                    for (uint32_t k = 0; k < K; k++)
                        if (sc_req[k] != 0 && P.alloc_scalar[(uint64_t)k * N + g] < sc_req[k] + SC.req_scalar[(uint64_t)k * N + g])
                            reasons |= 1u << (SFC_SCALAR0 + k);
                }
            }
            if (reasons) return reasons;
            for (uint32_t jh = 0; jh < n_hard; jh++) {
                uint32_t e = e_hard + jh;
                int32_t d = S.dom[ENT(ER_T, e) * L + idx];
                if (d < 0) return 1u << SFC_PTS_MISSING;
                int64_t match = (S.regbits[idx] >> jh) & 1 ? S.val[e * L + idx] : 0;
                int64_t skew = match + ENT(ER_B, e) - (int64_t)hard_min[jh];
                if (skew > ENT(ER_A, e)) return 1u << SFC_PTS_SKEW;
            }
            if (n_aff) {
                bool pods_exist = true, missing = false;
                for (uint32_t e = e_aff; e < e_anti; e++) {
                    int32_t d = S.dom[ENT(ER_T, e) * L + idx];
                    if (d < 0) { missing = true; break; }
                    if (S.val[e * L + idx] <= 0) pods_exist = false;
                }
                bool ok = true;
                if (missing) ok = false;
                else if (!pods_exist) ok = (aff_total == 0 && (cflags & SIMON_CLS_IPA_SELF_MATCH));
                if (!ok) return 1u << SFC_IPA_AFF;
            }
            for (uint32_t e = e_anti; e < e_exist; e++)
                if (S.dom[ENT(ER_T, e) * L + idx] >= 0 && S.val[e * L + idx] > 0) return 1u << SFC_IPA_ANTI;
            for (uint32_t e = e_exist; e < e_isc; e++)
                if (S.dom[ENT(ER_T, e) * L + idx] >= 0 && S.val[e * L + idx] > 0) return 1u << SFC_IPA_EXIST;
            if (cw[SCW_GPU_MEM] > 0) {
                uint32_t g = (uint32_t)S.node_g[idx];
                int slots[64];
                int64_t total = P.gpu_total_mem ? P.gpu_total_mem[g] : 0;
                if (total < cw[SCW_GPU_MEM] || gpu_allocate(P, SC, cw[SCW_GPU_MEM], cw[SCW_GPU_COUNT], g, slots) == 0) return 1u << SFC_GPU;
            }
            return 0;
        };

        long long rv[SK_NV];
        // F, n_ignored (sum); na, tt, simon, ipa (max); simon, ipa (min); 8 domain-bitmask words (or)
        const int rop[SK_NV] = {0, 0, 2, 2, 2, 2, 1, 1, 3, 3, 3, 3, 3, 3, 3, 3};
        rv[0] = 0; rv[1] = 0; rv[2] = 0; rv[3] = 0; rv[4] = -INT64_MAX; rv[5] = 0; rv[6] = INT64_MAX; rv[7] = 0;
#pragma unroll
        for (int q = 8; q < SK_NV; q++) rv[q] = 0;
        for (uint32_t s = 0; s < NPT; s++) {
            uint32_t idx = s * TPB + tid;
            uint8_t nf = S.nflags[idx];
            if (!(nf & NF_VALID)) continue;
            bool feas = filter_node(idx) == 0;
            bool counted = feas && !(nf & NF_IGNORED);
            if (counted)
                for (uint32_t js = 0; js < n_soft; js++) {
                    uint32_t e = e_soft + js;
                    int32_t wo = ENT(ER_WOFF, e);
                    if (wo < 0) continue;
                    int32_t d = S.dom[ENT(ER_T, e) * L + idx];
                    uint32_t w = (uint32_t)wo + ((uint32_t)d >> 6);
#pragma unroll
                    for (int q = 0; q < SK_MAXW; q++)
                        if ((uint32_t)q == w) rv[8 + q] |= (long long)(1ull << (d & 63));
                }
            if (any_table && counted != ((nf & NF_COUNTED) != 0)) {
                for (uint32_t js = 0; js < n_soft; js++) {
                    uint32_t e = e_soft + js;
                    if (ENT(ER_B, e) || ENT(ER_WOFF, e) >= 0) continue;
                    int32_t d = S.dom[ENT(ER_T, e) * L + idx];
                    if (counted) { if (atomicAdd(&SC.fcount[(uint64_t)js * P.max_dom + d], 1) == 0) atomicAdd(&SC.size[js], 1); }
                    else { if (atomicAdd(&SC.fcount[(uint64_t)js * P.max_dom + d], -1) == 1) atomicAdd(&SC.size[js], -1); }
                }
            }
            nf = (nf & ~(NF_COUNTED | NF_FEASIBLE)) | (counted ? NF_COUNTED : 0) | (feas ? NF_FEASIBLE : 0);
            S.nflags[idx] = nf;
            if (feas) {
                int64_t ip = 0;
                for (uint32_t e = e_isc; e < E; e++)
                    if (S.dom[ENT(ER_T, e) * L + idx] >= 0) ip += (int64_t)ENT(ER_A, e) * S.val[e * L + idx];
                S.raw_ipa[idx] = (int32_t)ip;
                rv[0] += 1;
                if (nf & NF_IGNORED) rv[1] += 1;
                long long na = S.raw_na[idx], tt = S.raw_tt[idx], sm = S.raw_simon[idx];
                rv[2] = na > rv[2] ? na : rv[2]; rv[3] = tt > rv[3] ? tt : rv[3];
                rv[4] = sm > rv[4] ? sm : rv[4]; rv[6] = sm < rv[6] ? sm : rv[6];
                rv[5] = ip > rv[5] ? ip : rv[5]; rv[7] = ip < rv[7] ? ip : rv[7];
            }
        }
        if (any_table) __threadfence();
        sk_allreduce<SK_NV, false>(R, rv, rop, pl);
        const int64_t F = rv[0], n_ign = rv[1], na_max = rv[2], tt_max = rv[3], simon_max = rv[4], ipa_max = rv[5],
                      simon_min = rv[6], ipa_min = rv[7];

        if (F == 0) {
            // ---- unschedulable: histogram of per-node failure reasons (FitError, generic_scheduler.go:72-90) ----
            if (n_fail < P.max_fail && SC.fail_counts) {
                for (uint32_t s = 0; s < NPT; s++) {
                    uint32_t idx = s * TPB + tid;
                    if (!(S.nflags[idx] & NF_VALID)) continue;
                    uint32_t rs = filter_node(idx);
                    while (rs) { int b = __ffs(rs) - 1; rs &= rs - 1; atomicAdd(&SC.fail_counts[(uint64_t)n_fail * SIMON_N_FAIL_CODES + b], 1u); }
                }
                if (leader) SC.fail_pod[n_fail] = i;
            }
            if (leader) { SC.out_node[i] = -1; if (SC.out_score) SC.out_score[i] = 0; }
            n_fail++;
            i++;
            continue;
        }

        // ---- P2: PodTopologySpread raw scores (scoring.go:175-208) ----
        long long pv[2];
        const int pop[2] = {1, 2};
        pv[0] = INT64_MAX; pv[1] = 0;
        if (n_soft > 0 && F > 1) {
            double soft_w[SK_MAX_SOFT];
#pragma unroll
            for (int js = 0; js < SK_MAX_SOFT; js++) {
                soft_w[js] = 0.0;
                if ((uint32_t)js < n_soft) {
                    uint32_t e = e_soft + js;
                    int64_t size;
                    if (ENT(ER_B, e)) size = F - n_ign;
                    else if (ENT(ER_WOFF, e) >= 0) {
                        uint32_t nw = (P.topo_ndom[ENT(ER_T, e)] + 63) / 64;
                        size = 0;
#pragma unroll
                        for (int q = 0; q < SK_MAXW; q++)
                            if (q >= ENT(ER_WOFF, e) && (uint32_t)q < (uint32_t)ENT(ER_WOFF, e) + nw) size += __popcll((unsigned long long)rv[8 + q]);
                    } else size = (int64_t)ldcg32(&SC.size[js]);
                    soft_w[js] = P.log_table[size + 2];
                }
            }
            for (uint32_t s = 0; s < NPT; s++) {
                uint32_t idx = s * TPB + tid;
                uint8_t nf = S.nflags[idx];
                if (!(nf & NF_FEASIBLE)) continue;
                int64_t raw = 0;
                if (!(nf & NF_IGNORED)) {
                    double score = 0.0;
#pragma unroll
                    for (int js = 0; js < SK_MAX_SOFT; js++)
                        if ((uint32_t)js < n_soft) {
                            uint32_t e = e_soft + js;
                            double sfc = (double)S.val[e * L + idx] * soft_w[js] + (double)(ENT(ER_A, e) - 1);
                            score = score + sfc;
                        }
                    raw = f2i(score);
                    pv[0] = raw < pv[0] ? raw : pv[0];
                    pv[1] = raw > pv[1] ? raw : pv[1];
                }
                S.raw_pts[idx] = (int32_t)raw;
            }
            sk_allreduce<2, false>(R, pv, pop, pl);
        }
        const int64_t pts_min = pv[0], pts_max = pv[1];

        // ---- P3: totals and arg-max (first node in scenario order wins ties) ----
        long long best[1];
        const int bop[1] = {2};
        best[0] = -1;
#pragma unroll
        for (int q = 0; q < SK_PAYLOAD; q++) pl[q] = 0;
        const int64_t simon_range = simon_max - simon_min, ipa_diff = ipa_max - ipa_min;
        for (uint32_t s = 0; s < NPT; s++) {
            uint32_t idx = s * TPB + tid;
            uint8_t nf = S.nflags[idx];
            if (!(nf & NF_FEASIBLE)) continue;
            int64_t total = 0;
            if (F > 1) {
                int64_t capc = S.alloc_mcpu[idx], rqc = S.nz_mcpu[idx] + cw[SCW_SCORE_MCPU];
                int64_t capm = S.alloc_mem[idx], rqm = S.nz_mem[idx] + cw[SCW_SCORE_MEM];
                int64_t s1 = (capc == 0 || rqc > capc) ? 0 : ((capc - rqc) * 100) / capc;
                int64_t s2 = (capm == 0 || rqm > capm) ? 0 : ((capm - rqm) * 100) / capm;
                int64_t la = (s1 + s2) / 2;
                double cf = capc == 0 ? 1.0 : (double)rqc / (double)capc;
                double mf = capm == 0 ? 1.0 : (double)rqm / (double)capm;
                int64_t ba = 0;
                if (!(cf >= 1.0 || mf >= 1.0)) ba = f2i((1.0 - fabs(cf - mf)) * 100.0);
                int64_t na = na_max == 0 ? S.raw_na[idx] : (100 * (int64_t)S.raw_na[idx]) / na_max;
                int64_t tt = tt_max == 0 ? 100 : 100 - (100 * (int64_t)S.raw_tt[idx]) / tt_max;
                int64_t sm = simon_range == 0 ? 0 : ((S.raw_simon[idx] - simon_min) * 100) / simon_range;
                int64_t ip = 0;
                if (ipa_diff > 0) ip = f2i(100.0 * ((double)((int64_t)S.raw_ipa[idx] - ipa_min) / (double)ipa_diff));
                int64_t pts;
                if (n_soft == 0) pts = 100;
                else if (nf & NF_IGNORED) pts = 0;
                else if (pts_max == 0) pts = 100;
                else pts = (100 * (pts_max + pts_min - (int64_t)S.raw_pts[idx])) / pts_max;
                total = ba + la + ip + na + 2 * pts + tt + 2 * sm + (int64_t)S.extra[idx];
            }
            uint32_t r = s * CT + gtid;
            long long key = (long long)(((unsigned long long)total << 24) | (unsigned long long)(0xFFFFFFu - r));
            if (key > best[0]) {
                best[0] = key;
#pragma unroll
                for (int t = 0; t < SIMON_MAX_TOPOS; t++) pl[t] = (uint32_t)t < T ? S.dom[t * L + idx] : -1;
                pl[8] = nf;
            }
        }
        sk_allreduce<1, true>(R, best, bop, pl);
        const uint32_t win_r = 0xFFFFFFu - (uint32_t)((unsigned long long)best[0] & 0xFFFFFFu);
        const int64_t win_total = (int64_t)((unsigned long long)best[0] >> 24);
        const bool win_ignored = ((uint32_t)pl[8] & NF_IGNORED) != 0;

        // ---- commit (AssumePod / NodeInfo.AddPod) ----
        if (win_r % CT == gtid) {
            uint32_t idx = (win_r / CT) * TPB + tid;
            uint32_t g = (uint32_t)S.node_g[idx];
            S.req_mcpu[idx] += cw[SCW_REQ_MCPU]; S.req_mem[idx] += cw[SCW_REQ_MEM]; S.req_eph[idx] += cw[SCW_REQ_EPH];
            S.nz_mcpu[idx] += cw[SCW_NZ_MCPU]; S.nz_mem[idx] += cw[SCW_NZ_MEM]; S.num_pods[idx] += 1;
            for (uint32_t k = 0; k < K; k++) SC.req_scalar[(uint64_t)k * N + g] += sc_req[k];
            const int64_t *inc = cw + cw[SCW_OFF_INC];
            for (int64_t u = 0; u < cw[SCW_N_INC]; u++) {
                int64_t k = inc[3 * u], t = inc[3 * u + 1], sig = inc[3 * u + 2];
                int32_t d = S.dom[t * L + idx];
                if (d < 0) continue;
                if (sig >= 0) {
                    bool el = (sig == (int64_t)cls) ? !win_ignored : elig_eval(P, sig, RC, g);   // the winner passed NodeAffinity
                    if (!el) continue;
                }
                atomicAdd(&SC.cnt[P.cnt_off[k] + d], 1);
                atomicAdd(&SC.cnt_total[k], 1);
            }
            if (cw[SCW_GPU_MEM] > 0) {
                int slots[64];
                int ns = gpu_allocate(P, SC, cw[SCW_GPU_MEM], cw[SCW_GPU_COUNT], g, slots);
                for (int q = 0; q < ns && q < 64; q++) SC.gpu_used[(uint64_t)slots[q] * N + g] += cw[SCW_GPU_MEM];
            }
            SC.out_node[i] = (int32_t)g;
            if (SC.out_score) SC.out_score[i] = F > 1 ? win_total : 0;
        }
        // every thread folds the winner into its cached counter values
        for (uint32_t e = 0; e < E; e++) {
            if (!ENT(ER_INC, e)) continue;
            int32_t kind = ENT(ER_KIND, e), t = ENT(ER_T, e);
            bool host = kind == EK_SOFT && ENT(ER_B, e);
            int32_t wd = host ? (int32_t)pl[0] : (int32_t)pl[t];
            if (kind == EK_SOFT && !host && win_ignored) continue;
            if (wd < 0) continue;
            const int32_t *dcol = S.dom + (host ? 0 : t) * L;
            for (uint32_t s = 0; s < NPT; s++) {
                uint32_t idx = s * TPB + tid;
                if ((S.nflags[idx] & NF_VALID) && dcol[idx] == wd) S.val[e * L + idx] += 1;
            }
            if (kind == EK_AFF) aff_total += 1;
        }
        n_sched++;
        i++;
    }

    // ---- write the dynamic state back -------------------------------------------------------------------
    for (uint32_t s = 0; s < NPT; s++) {
        uint32_t idx = s * TPB + tid;
        if (!(S.nflags[idx] & NF_VALID)) continue;
        uint32_t g = (uint32_t)S.node_g[idx];
        SC.req_mcpu[g] = S.req_mcpu[idx]; SC.req_mem[g] = S.req_mem[idx]; SC.req_eph[g] = S.req_eph[idx];
        SC.nz_mcpu[g] = S.nz_mcpu[idx]; SC.nz_mem[g] = S.nz_mem[idx]; SC.num_pods[g] = S.num_pods[idx];
    }
    if (leader) {
        if (SC.n_fail) *SC.n_fail = n_fail;
        if (SC.n_sched) *SC.n_sched = n_sched;
        if (SC.clk) SC.clk[1] = sk_globaltimer();
    }
    cluster.sync();   // no CTA may exit while peers can still address its shared memory
}
