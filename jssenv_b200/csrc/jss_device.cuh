// jss_device.cuh -- sm_100a device code of the batched job-shop environment.
//
// One warp simulates one environment.  Lane l owns the KJ = ceil(J/32) (1, 2 or 4)
// consecutive jobs KJ*l .. KJ*l+KJ-1 (so its slice of every per-job array is ONE
// 4/8/16-byte vector load/store and its slice of real_obs is 7 consecutive
// vectors), and lane m (< M <= 32) owns machine m.  Legal / no-op-blocked job sets
// live as warp-uniform ballot words.  Everything the reference keeps redundantly
// (event queue, illegal_actions[M][J], machine_legal, counters, needed_machine,
// total_perform) is re-derived in registers.
//
// Reference semantics implemented here (file:line relative to the reference):
//   env_reset_regs      JSSEnv/envs/jss_env.py:145-181  reset
//   env_advance         JSSEnv/envs/jss_env.py:495-637  increase_time_step
//   env_prioritize      JSSEnv/envs/jss_env.py:183-254  _prioritization_non_final
//   env_check_no_op     JSSEnv/envs/jss_env.py:256-401  _check_no_op
//   env_step            JSSEnv/envs/jss_env.py:403-481  step (+483-493, 639-653)
//   env_emit            JSSEnv/envs/jss_env.py:102-134  observation / mask
//   env_select_action   JSSEnv/dispatching.py:92-408    rules; README.md:58-60 sampler
//
// The same source is compiled for the host by tests/emu (32 fibers per warp,
// collectives emulated) so the logic can be differential-tested against the
// oracle without a GPU; that build is test-only and is never loaded by the
// product library.
#pragma once
#include <stdint.h>

#include "../../include/jss_b200.h"
#include "jss_rng.h"
#include "jss_types.h"

#define JSS_FULL 0xffffffffu
#define JSS_INF 0x7fffffff

#ifndef JSS_DEV
#define JSS_DEV __device__ __forceinline__
#endif

struct InstView {  // instance tables staged in shared memory
    const uint16_t *ops;
    const int32_t *len;
    const uint16_t *rem;
    int J, M, max_time_op, max_time_jobs, sum_op;
};

template <int KJ>
struct EnvRegs {
    int todo[KJ], tufco[KJ], idle_last[KJ], total_idle[KJ], col4[KJ];
    uint32_t op[KJ];       // packed current op of each owned job, JSS_OP_NONE if none
    uint32_t L[KJ], B[KJ];  // warp-uniform: legal / no-op-blocked ballots
    int tuam;               // lane m: time_until_available_machine[m]
    int t;                  // current_time_step
    uint32_t flags;
    int ep_steps, ep_return;
};

// ---- small helpers --------------------------------------------------------------
template <int KJ, typename T>
JSS_DEV T jss_sel(const T (&a)[KJ], int i) {  // a[i] without dynamic register indexing
    T r = a[0];
#pragma unroll
    for (int k = 1; k < KJ; k++) r = (i == k) ? a[k] : r;
    return r;
}
JSS_DEV uint32_t jss_bit(uint32_t mask, uint32_t pos) { return pos < 32u ? (mask >> pos) & 1u : 0u; }
JSS_DEV uint32_t jss_op_m(uint32_t op) { return op >> JSS_OP_SHIFT; }
JSS_DEV int jss_op_d(uint32_t op) { return (int)(op & JSS_OP_DMASK); }
JSS_DEV uint32_t jss_op_at(const InstView &iv, int j, int ts) { return iv.ops[j * iv.M + ts]; }

template <int KJ>
JSS_DEV uint32_t jss_any(const uint32_t (&a)[KJ]) {
    uint32_t r = a[0];
#pragma unroll
    for (int k = 1; k < KJ; k++) r |= a[k];
    return r;
}
template <int KJ>
JSS_DEV int jss_count(const uint32_t (&a)[KJ]) {
    int r = 0;
#pragma unroll
    for (int k = 0; k < KJ; k++) r += __popc(a[k]);
    return r;
}

// vector access to a lane's KJ-word slice
template <int KJ>
JSS_DEV void jss_ld(const int32_t *p, int (&o)[KJ]);
template <>
JSS_DEV void jss_ld<1>(const int32_t *p, int (&o)[1]) { o[0] = *p; }
template <>
JSS_DEV void jss_ld<2>(const int32_t *p, int (&o)[2]) {
    int2 v = *reinterpret_cast<const int2 *>(p);
    o[0] = v.x; o[1] = v.y;
}
template <>
JSS_DEV void jss_ld<4>(const int32_t *p, int (&o)[4]) {
    int4 v = *reinterpret_cast<const int4 *>(p);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <int KJ>
JSS_DEV void jss_st(int32_t *p, const int (&o)[KJ]);
template <>
JSS_DEV void jss_st<1>(int32_t *p, const int (&o)[1]) { *p = o[0]; }
template <>
JSS_DEV void jss_st<2>(int32_t *p, const int (&o)[2]) { *reinterpret_cast<int2 *>(p) = make_int2(o[0], o[1]); }
template <>
JSS_DEV void jss_st<4>(int32_t *p, const int (&o)[4]) {
    *reinterpret_cast<int4 *>(p) = make_int4(o[0], o[1], o[2], o[3]);
}

// ---- state block <-> registers -----------------------------------------------------
template <int KJ>
JSS_DEV void env_derive_ops(const InstView &iv, EnvRegs<KJ> &s, int lane) {
#pragma unroll
    for (int i = 0; i < KJ; i++) {
        int j = KJ * lane + i;
        s.op[i] = (j < iv.J && s.todo[i] < iv.M) ? jss_op_at(iv, j, s.todo[i]) : JSS_OP_NONE;
    }
}

template <int KJ>
JSS_DEV void env_load(const JssParams &p, const InstView &iv, int env, int lane, EnvRegs<KJ> &s) {
    const int32_t *blk = p.state + (size_t)env * p.block_words;
    const int Jc = p.Jcap;
    if (KJ * lane < Jc) {
        jss_ld<KJ>(blk + KJ * lane, s.todo);
        jss_ld<KJ>(blk + Jc + KJ * lane, s.tufco);
        jss_ld<KJ>(blk + 2 * Jc + KJ * lane, s.idle_last);
        jss_ld<KJ>(blk + 3 * Jc + KJ * lane, s.total_idle);
        jss_ld<KJ>(blk + 4 * Jc + KJ * lane, s.col4);
    }
#pragma unroll
    for (int i = 0; i < KJ; i++)
        if (KJ * lane + i >= iv.J) {  // padding slot: behaves like a finished job
            s.todo[i] = iv.M; s.tufco[i] = 0; s.idle_last[i] = 0; s.total_idle[i] = 0; s.col4[i] = 0;
        }
    s.tuam = (lane < p.Mcap) ? blk[5 * Jc + lane] : 0;
    if (lane >= iv.M) s.tuam = 0;
    const int4 *tail = reinterpret_cast<const int4 *>(blk + 5 * Jc + p.Mcap);
    int4 l4 = tail[0], b4 = tail[1], h4 = tail[2];
    const uint32_t lw[4] = {(uint32_t)l4.x, (uint32_t)l4.y, (uint32_t)l4.z, (uint32_t)l4.w};
    const uint32_t bw[4] = {(uint32_t)b4.x, (uint32_t)b4.y, (uint32_t)b4.z, (uint32_t)b4.w};
#pragma unroll
    for (int i = 0; i < KJ; i++) { s.L[i] = lw[i]; s.B[i] = bw[i]; }
    s.t = h4.x; s.flags = (uint32_t)h4.y; s.ep_steps = h4.z; s.ep_return = h4.w;
    env_derive_ops<KJ>(iv, s, lane);
}

// policy kernels read only what the rule looks at: the ballots + header always, todo for
// every rule (current op / remaining work / remaining ops), idle_last for FIFO
template <int KJ>
JSS_DEV void env_load_for_policy(const JssParams &p, const InstView &iv, int env, int lane, EnvRegs<KJ> &s,
                                 int rule) {
    const int32_t *blk = p.state + (size_t)env * p.block_words;
    const int Jc = p.Jcap;
#pragma unroll
    for (int i = 0; i < KJ; i++) { s.todo[i] = iv.M; s.tufco[i] = 0; s.idle_last[i] = 0; s.total_idle[i] = 0; s.col4[i] = 0; }
    if (rule != JSS_RULE_RANDOM && KJ * lane < Jc) {
        jss_ld<KJ>(blk + KJ * lane, s.todo);
        if (rule == JSS_RULE_FIFO) jss_ld<KJ>(blk + 2 * Jc + KJ * lane, s.idle_last);
    }
#pragma unroll
    for (int i = 0; i < KJ; i++)
        if (KJ * lane + i >= iv.J) { s.todo[i] = iv.M; s.idle_last[i] = 0; }
    s.tuam = 0;
    const int4 *tail = reinterpret_cast<const int4 *>(blk + 5 * Jc + p.Mcap);
    int4 l4 = tail[0], h4 = tail[2];
    const uint32_t lw[4] = {(uint32_t)l4.x, (uint32_t)l4.y, (uint32_t)l4.z, (uint32_t)l4.w};
#pragma unroll
    for (int i = 0; i < KJ; i++) { s.L[i] = lw[i]; s.B[i] = 0u; }
    s.t = h4.x; s.flags = (uint32_t)h4.y; s.ep_steps = h4.z; s.ep_return = h4.w;
    if (rule != JSS_RULE_RANDOM) env_derive_ops<KJ>(iv, s, lane);
    else {
#pragma unroll
        for (int i = 0; i < KJ; i++) s.op[i] = JSS_OP_NONE;
    }
}

template <int KJ>
JSS_DEV void env_store(const JssParams &p, int env, int lane, const EnvRegs<KJ> &s) {
    int32_t *blk = p.state + (size_t)env * p.block_words;
    const int Jc = p.Jcap;
    if (KJ * lane < Jc) {
        jss_st<KJ>(blk + KJ * lane, s.todo);
        jss_st<KJ>(blk + Jc + KJ * lane, s.tufco);
        jss_st<KJ>(blk + 2 * Jc + KJ * lane, s.idle_last);
        jss_st<KJ>(blk + 3 * Jc + KJ * lane, s.total_idle);
        jss_st<KJ>(blk + 4 * Jc + KJ * lane, s.col4);
    }
    if (lane < p.Mcap) blk[5 * Jc + lane] = s.tuam;
    if (lane < 3) {
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < KJ; i++) w[i] = s.L[i];
        } else if (lane == 1) {
#pragma unroll
            for (int i = 0; i < KJ; i++) w[i] = s.B[i];
        } else {
            w[0] = (uint32_t)s.t; w[1] = s.flags; w[2] = (uint32_t)s.ep_steps; w[3] = (uint32_t)s.ep_return;
        }
        reinterpret_cast<int4 *>(blk + 5 * Jc + p.Mcap)[lane] = make_int4((int)w[0], (int)w[1], (int)w[2], (int)w[3]);
    }
}

// jss_env.py:145-181
template <int KJ>
JSS_DEV void env_reset_regs(const InstView &iv, EnvRegs<KJ> &s, int lane) {
#pragma unroll
    for (int i = 0; i < KJ; i++) {
        const bool valid = KJ * lane + i < iv.J;
        s.todo[i] = valid ? 0 : iv.M;
        s.tufco[i] = 0; s.idle_last[i] = 0; s.total_idle[i] = 0; s.col4[i] = 0;
        s.L[i] = __ballot_sync(JSS_FULL, valid);  // every job legal, no-op illegal (:160-161)
        s.B[i] = 0u;
    }
    s.tuam = 0; s.t = 0; s.flags = 0u; s.ep_steps = 0; s.ep_return = 0;
    env_derive_ops<KJ>(iv, s, lane);
}

// ---- increase_time_step (jss_env.py:495-637); returns hole_planning ---------------
template <int KJ>
JSS_DEV int env_advance(const InstView &iv, EnvRegs<KJ> &s, int lane) {
    // next event = smallest positive machine countdown (the sorted event list of
    // the reference always equals {t + tuam[m] : tuam[m] > 0}, appendix A.1)
    const int diff = (int)__reduce_min_sync(JSS_FULL, (unsigned)(s.tuam > 0 ? s.tuam : JSS_INF));
    const int tuam_old = s.tuam;
    const int hole = (int)__reduce_add_sync(JSS_FULL, (unsigned)((lane < iv.M && tuam_old < diff) ? diff - tuam_old : 0));
    s.t += diff;
#pragma unroll
    for (int i = 0; i < KJ; i++) {
        const int j = KJ * lane + i;
        const int was = s.tufco[i];
        bool finished = false;
        if (was > 0) {                                    // running (:529)
            const int left = was - diff;
            s.tufco[i] = left > 0 ? left : 0;
            if (left <= 0) {                              // op done (:550)
                s.total_idle[i] += diff - was;
                s.idle_last[i] = diff - was;
                s.todo[i] += 1;
                finished = true;
                s.op[i] = (s.todo[i] < iv.M) ? jss_op_at(iv, j, s.todo[i]) : JSS_OP_NONE;
            }
        } else if (s.todo[i] < iv.M) {                    // waiting (:594)
            s.total_idle[i] += diff;
            s.idle_last[i] += diff;
        }
        // real_obs[:,4] uses the PRE-decrement countdown of the next machine (:569-578)
        const int tq = __shfl_sync(JSS_FULL, tuam_old, (int)(jss_op_m(s.op[i]) & 31u));
        if (finished) {
            if (s.op[i] != JSS_OP_NONE) { const int w = tq - diff; s.col4[i] = w > 0 ? w : 0; }
            else s.col4[i] = iv.max_time_op;              // encodes 1.0 (:586)
        }
    }
    { const int left = tuam_old - diff; s.tuam = left > 0 ? left : 0; }
    const uint32_t free_m = __ballot_sync(JSS_FULL, lane < iv.M && s.tuam == 0);
#pragma unroll
    for (int i = 0; i < KJ; i++) {                        // legalisation (:616-634)
        const bool mine = s.op[i] != JSS_OP_NONE && jss_bit(free_m, jss_op_m(s.op[i])) &&
                          !jss_bit(s.B[i], (uint32_t)lane);
        s.L[i] |= __ballot_sync(JSS_FULL, mine);
    }
    return hole;
}

// machines that have at least one legal job (== machine_legal of the reference)
template <int KJ>
JSS_DEV uint32_t env_machine_legal(const EnvRegs<KJ> &s, int lane) {
    uint32_t mine = 0u;
#pragma unroll
    for (int i = 0; i < KJ; i++)
        if (jss_bit(s.L[i], (uint32_t)lane)) mine |= 1u << (jss_op_m(s.op[i]) & 31u);
    return __reduce_or_sync(JSS_FULL, mine);
}

// ---- _prioritization_non_final (jss_env.py:183-254) --------------------------------
template <int KJ>
JSS_DEV void env_prioritize(const InstView &iv, EnvRegs<KJ> &s, int lane) {
    bool fin[KJ];
    uint32_t fin_m = 0u;
#pragma unroll
    for (int i = 0; i < KJ; i++) {
        fin[i] = jss_bit(s.L[i], (uint32_t)lane) && s.todo[i] == iv.M - 1;
        if (fin[i]) fin_m |= 1u << (jss_op_m(s.op[i]) & 31u);
    }
    fin_m = __reduce_or_sync(JSS_FULL, fin_m);            // machines wanted by a legal FINAL op
    if (fin_m == 0u) return;                              // nothing can be de-legalised
    const uint32_t free_m = __ballot_sync(JSS_FULL, lane < iv.M && s.tuam == 0);
    int cand_d[KJ];                                       // duration if legal non-final op whose NEXT machine is free
#pragma unroll
    for (int i = 0; i < KJ; i++) {
        cand_d[i] = JSS_INF;
        if (jss_bit(s.L[i], (uint32_t)lane) && !fin[i]) {
            const uint32_t nxt = jss_op_at(iv, KJ * lane + i, s.todo[i] + 1);
            if (jss_bit(free_m, jss_op_m(nxt))) cand_d[i] = jss_op_d(s.op[i]);   // :234-239
        }
    }
    bool kill[KJ];
#pragma unroll
    for (int i = 0; i < KJ; i++) kill[i] = false;
    while (fin_m) {                                       // per machine with a legal final op
        const uint32_t m = (uint32_t)(__ffs((int)fin_m) - 1);
        fin_m &= fin_m - 1u;
        int mn = JSS_INF;
#pragma unroll
        for (int i = 0; i < KJ; i++)
            if (cand_d[i] != JSS_INF && jss_op_m(s.op[i]) == m) mn = min(mn, cand_d[i]);
        mn = (int)__reduce_min_sync(JSS_FULL, (unsigned)mn);   // min_non_final (:238)
        if (mn != JSS_INF) {
#pragma unroll
            for (int i = 0; i < KJ; i++)
                if (fin[i] && jss_op_m(s.op[i]) == m && jss_op_d(s.op[i]) > mn) kill[i] = true;  // :252
        }
    }
#pragma unroll
    for (int i = 0; i < KJ; i++) s.L[i] &= ~__ballot_sync(JSS_FULL, kill[i]);
}

// ---- _check_no_op (jss_env.py:256-401); returns legal_actions[J] -----------------
template <int KJ>
JSS_DEV bool env_check_no_op(const InstView &iv, const EnvRegs<KJ> &s, int lane, uint32_t ML) {
    const int nlegal = jss_count<KJ>(s.L);
    const unsigned mn = __reduce_min_sync(JSS_FULL, (unsigned)(s.tuam > 0 ? s.tuam : JSS_INF));
    if (mn == (unsigned)JSS_INF || __popc(ML) > 3 || nlegal > 4) return false;   // gate :284-288
    const int next_event = s.t + (int)mn;                                        // :293
    int maxh = s.t;                                                              // :296
    int lm0 = -1, lm1 = -1, lm2 = -1;                   // the <= 3 legal machines and their horizons
    const int hinit = s.t + iv.max_time_op;             // :300-302
    int h0 = hinit, h1 = hinit, h2 = hinit;
    // pass 1 (:305-321): legal jobs in ascending job index = ascending lane, then slot
    uint32_t lanes = jss_any<KJ>(s.L);
    while (lanes) {
        const int l = __ffs((int)lanes) - 1;
        lanes &= lanes - 1u;
#pragma unroll
        for (int i = 0; i < KJ; i++) {
            if (jss_bit(s.L[i], (uint32_t)l)) {         // warp-uniform branch
                const uint32_t o = __shfl_sync(JSS_FULL, s.op[i], l);
                const int m = (int)jss_op_m(o);
                const int end = s.t + jss_op_d(o);
                if (end < next_event) return false;     // :314-315
                int cur;
                if (lm0 == m || lm0 < 0) { lm0 = m; h0 = min(h0, end); cur = h0; }
                else if (lm1 == m || lm1 < 0) { lm1 = m; h1 = min(h1, end); cur = h1; }
                else { lm2 = m; h2 = min(h2, end); cur = h2; }
                maxh = max(maxh, cur);                  // :321
            }
        }
    }
    // pass 2 (:324-401): jobs that are not legal now but may need a legal machine soon
    uint32_t want = 0u;
#pragma unroll
    for (int i = 0; i < KJ; i++) {
        // countdown of the job's current machine (case 2, :374-377)
        const int tq = __shfl_sync(JSS_FULL, s.tuam, (int)(jss_op_m(s.op[i]) & 31u));
        const int j = KJ * lane + i;
        if (j < iv.J && !jss_bit(s.L[i], (uint32_t)lane)) {
            int ts = 0, tm = 0;
            bool go = false;
            if (s.tufco[i] > 0 && s.todo[i] + 1 < iv.M) {            // case 1 (:327-337)
                ts = s.todo[i] + 1; tm = s.t + s.tufco[i]; go = true;
            } else if (!jss_bit(s.B[i], (uint32_t)lane) && s.todo[i] < iv.M) {  // case 2 (:366-377)
                ts = s.todo[i]; tm = s.t + tq; go = true;
            }
            if (go) {
                while (ts < iv.M - 1 && maxh > tm) {                 // :340-342 / :380-382
                    const uint32_t o = jss_op_at(iv, j, ts);
                    const int m = (int)jss_op_m(o);
                    if (jss_bit(ML, (uint32_t)m)) {
                        const int hz = (m == lm0) ? h0 : (m == lm1) ? h1 : h2;
                        if (hz > tm) want |= 1u << m;               // machine_next.add (:351 / :391)
                    }
                    tm += jss_op_d(o);
                    ts += 1;
                }
            }
        }
    }
    want = __reduce_or_sync(JSS_FULL, want);
    return ML != 0u && want == ML;                      // len(machine_next) == nb_machine_legal
}

// ---- observation / mask / reward (jss_env.py:102-134, 483-493) ---------------------
struct StepOut {
    int raw_reward;
    bool wrote;
};

template <int KJ>
JSS_DEV void env_emit_obs(const JssParams &p, const InstView &iv, const EnvRegs<KJ> &s, int env, int lane,
                          float *scratch) {
    const float mto = (float)iv.max_time_op, mtj = (float)iv.max_time_jobs, sop = (float)iv.sum_op;
    const float fM = (float)iv.M;
    float v[KJ * 7];
#pragma unroll
    for (int i = 0; i < KJ; i++) {
        const int j = KJ * lane + i;
        const bool valid = j < iv.J;
        // total_perform_op_time_jobs == t - total_idle while the job is unfinished,
        // jobs_length[j] afterwards (every advance adds `difference` to exactly one of
        // the two counters until the job completes)
        const int perf = valid ? (s.todo[i] < iv.M ? s.t - s.total_idle[i] : iv.len[j]) : 0;
        v[7 * i + 0] = jss_bit(s.L[i], (uint32_t)lane) ? 1.0f : 0.0f;
        v[7 * i + 1] = __fdiv_rn((float)s.tufco[i], mto);
        v[7 * i + 2] = valid ? __fdiv_rn((float)s.todo[i], fM) : 0.0f;
        v[7 * i + 3] = __fdiv_rn((float)perf, mtj);
        v[7 * i + 4] = __fdiv_rn((float)s.col4[i], mto);
        v[7 * i + 5] = __fdiv_rn((float)s.idle_last[i], sop);
        v[7 * i + 6] = __fdiv_rn((float)s.total_idle[i], sop);
    }
    // stage the lane's 7*KJ floats (one contiguous, conflict-free run per lane) ...
    float *mine = scratch + 7 * KJ * lane;
    if (KJ * lane >= iv.J) {
        // lanes past the last job stage nothing (scratch holds 7 * roundup(J, 4) floats)
    } else if (KJ == 4) {
#pragma unroll
        for (int q = 0; q < 7; q++)
            reinterpret_cast<float4 *>(mine)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    } else if (KJ == 2) {
#pragma unroll
        for (int q = 0; q < 7; q++) reinterpret_cast<float2 *>(mine)[q] = make_float2(v[2 * q], v[2 * q + 1]);
    } else {
#pragma unroll
        for (int q = 0; q < 7; q++) mine[q] = v[q];
    }
    __syncwarp();
    // ... and stream the J*7 floats of the env out with fully coalesced stores
    float *dst = p.obs + (size_t)env * p.jobs_max * 7;
    const int n = iv.J * 7;
    int done_elems = 0;
    if ((p.jobs_max & 3) == 0) {
        const int n4 = n >> 2;
        for (int k = lane; k < n4; k += 32)
            reinterpret_cast<float4 *>(dst)[k] = reinterpret_cast<const float4 *>(scratch)[k];
        done_elems = n4 << 2;
    }
    for (int k = done_elems + lane; k < n; k += 32) dst[k] = scratch[k];
    __syncwarp();
}

template <int KJ>
JSS_DEV void env_emit_mask(const JssParams &p, const InstView &iv, const EnvRegs<KJ> &s, int env, int lane,
                           bool noop) {
    uint8_t *row = p.mask + (size_t)env * p.mask_stride;
    if (KJ * lane <= iv.J) {
        uint32_t w = 0u;
#pragma unroll
        for (int i = 0; i < KJ; i++) {
            const int j = KJ * lane + i;
            const uint32_t b = (j < iv.J) ? jss_bit(s.L[i], (uint32_t)lane) : (j == iv.J ? (noop ? 1u : 0u) : 0u);
            w |= b << (8 * i);
        }
        if (KJ == 4) *reinterpret_cast<uint32_t *>(row + 4 * lane) = w;
        else if (KJ == 2) *reinterpret_cast<uint16_t *>(row + 2 * lane) = (uint16_t)w;
        else row[lane] = (uint8_t)w;
    }
    if (iv.J == 32 * KJ && lane == 0) row[iv.J] = noop ? 1 : 0;   // no lane owns byte J
}

template <int KJ>
JSS_DEV void env_emit_scalars(const JssParams &p, const InstView &iv, const EnvRegs<KJ> &s, int env, int lane,
                              int raw_reward) {
    if (lane == 0) {
        p.reward[env] = __fdiv_rn((float)raw_reward, (float)iv.max_time_op);   // :483-493
        p.reward_raw[env] = raw_reward;
        p.done[env] = (s.flags & JSS_FLAG_DONE) ? 1 : 0;
        p.time[env] = s.t;
        p.flags[env] = s.flags;
    }
}

// episode bookkeeping when _is_done() turns true (jss_env.py:649-652)
template <int KJ>
JSS_DEV void env_episode_end(const JssParams &p, const EnvRegs<KJ> &s, int env, int lane) {
    if (lane == 0) {
        const int prev = p.episode_count[env];
        p.episode_count[env] = prev + 1;
        p.last_makespan[env] = s.t;
        p.last_return[env] = s.ep_return;
        int64_t *acc = p.acc + (size_t)env * 4;
        acc[0] += s.ep_steps;
        acc[1] += s.t;
        acc[2] += s.ep_return;
        const uint64_t mm = (uint64_t)acc[3];
        uint32_t mn = prev ? (uint32_t)(mm >> 32) : 0xffffffffu, mx = prev ? (uint32_t)mm : 0u;
        if ((uint32_t)s.t < mn) mn = (uint32_t)s.t;
        if ((uint32_t)s.t > mx) mx = (uint32_t)s.t;
        acc[3] = (int64_t)(((uint64_t)mn << 32) | mx);
    }
}

// ---- step (jss_env.py:403-481) ---------------------------------------------------------
// Returns true if the env changed (outputs must be re-emitted).
template <int KJ>
JSS_DEV bool env_step(const JssParams &p, const InstView &iv, EnvRegs<KJ> &s, int env, int lane, int action,
                      int &raw_reward) {
    raw_reward = 0;
    if (action == JSS_ACTION_SKIP) return false;
    if (s.flags & JSS_FLAG_DONE) {
        if (p.create_flags & JSS_CREATE_AUTO_RESET) {
            env_reset_regs<KJ>(iv, s, lane);
            if (p.solution) {
                int32_t *sol = p.solution + (size_t)env * p.jobs_max * p.machines_max;
                for (int k = lane; k < p.jobs_max * p.machines_max; k += 32) sol[k] = -1;
            }
            return true;
        }
        return false;  // frozen until reset
    }
    const bool pending = __ballot_sync(JSS_FULL, s.tuam > 0) != 0u;
    int holes = 0;
    if (action == JSS_ACTION_ADVANCE) {                  // raw increase_time_step()
        if (!pending) { s.flags |= JSS_FLAG_ERROR; return false; }
        holes = env_advance<KJ>(iv, s, lane);
        raw_reward = -holes;
        // the heuristics and _is_done do NOT run here; legal_actions[J] keeps its value
        return true;
    }
    if (action == iv.J) {                                // no-op (:419-440)
        if (!pending) { s.flags |= JSS_FLAG_ERROR; return false; }   // IndexError at :517
#pragma unroll
        for (int i = 0; i < KJ; i++) { s.B[i] |= s.L[i]; s.L[i] = 0u; }   // :422-428
        bool more = true;
        do {                                             // :429-430
            holes += env_advance<KJ>(iv, s, lane);
            more = __ballot_sync(JSS_FULL, s.tuam > 0) != 0u;
        } while (jss_any<KJ>(s.L) == 0u && more);
        if (jss_any<KJ>(s.L) == 0u) s.flags |= JSS_FLAG_ERROR;     // the reference raises here
        raw_reward = -holes;
    } else {                                             // job allocation (:441-481)
        if (action < 0 || action > iv.J) { s.flags |= JSS_FLAG_ERROR; return false; }
        const int la = action / KJ, ia = action % KJ;
        const uint32_t opa = __shfl_sync(JSS_FULL, jss_sel<KJ>(s.op, ia), la);
        const int todo_a = __shfl_sync(JSS_FULL, jss_sel<KJ>(s.todo, ia), la);
        const bool legal_a = jss_bit(jss_sel<KJ>(s.L, ia), (uint32_t)la);
        if (opa == JSS_OP_NONE || !legal_a) { s.flags |= JSS_FLAG_ERROR; return false; }
        const uint32_t m_a = jss_op_m(opa);
        const int d_a = jss_op_d(opa);
        if ((uint32_t)lane == m_a) s.tuam = d_a;                     // :446
        if (lane == la) {
#pragma unroll
            for (int i = 0; i < KJ; i++) if (i == ia) s.tufco[i] = d_a;   // :447
            if (p.solution)                                           // :454
                p.solution[((size_t)env * p.jobs_max + action) * p.machines_max + todo_a] = s.t;
        }
#pragma unroll
        for (int i = 0; i < KJ; i++) {
            // every job waiting for machine m_a: no longer legal (:455-461), no longer
            // no-op-blocked (:464-467; illegal_actions[m][j] implies needed_machine[j]==m)
            const uint32_t need = __ballot_sync(JSS_FULL, s.op[i] != JSS_OP_NONE && jss_op_m(s.op[i]) == m_a);
            s.L[i] &= ~need; s.B[i] &= ~need;
        }
        bool more = true;                                // tuam[m_a] = d_a > 0
        while (jss_any<KJ>(s.L) == 0u && more) {         // :469-470
            holes += env_advance<KJ>(iv, s, lane);
            more = __ballot_sync(JSS_FULL, s.tuam > 0) != 0u;
        }
        raw_reward = d_a - holes;
    }
    env_prioritize<KJ>(iv, s, lane);                     // :432 / :471
    const uint32_t ML = env_machine_legal<KJ>(s, lane);
    const bool noop = env_check_no_op<KJ>(iv, s, lane, ML);   // :433 / :472
    s.flags &= ~(JSS_FLAG_NOOP_LEGAL | JSS_FLAG_DONE);
    if (noop) s.flags |= JSS_FLAG_NOOP_LEGAL;
    s.ep_steps += 1;
    s.ep_return += raw_reward;
    if (jss_any<KJ>(s.L) == 0u) {                        // _is_done (:649)
        s.flags |= JSS_FLAG_DONE;
        env_episode_end<KJ>(p, s, env, lane);
    }
    return true;
}

// ---- policies (JSSEnv/dispatching.py; README.md:58-60) ---------------------------------
template <int KJ>
JSS_DEV int env_select_action(const InstView &iv, const EnvRegs<KJ> &s, int lane, int rule, int coin_mode,
                              uint32_t h) {
    const int njobs = jss_count<KJ>(s.L);
    const bool noop = (s.flags & JSS_FLAG_NOOP_LEGAL) != 0u;
    if (s.flags & JSS_FLAG_DONE) return 0;               // ignored by step (auto-reset or frozen)
    if (njobs == 0) return noop ? iv.J : JSS_ACTION_SKIP;   // "only the no-op is legal" (e.g. :96-97)
    if (rule == JSS_RULE_RANDOM) {
        // uniform over the set bits of action_mask, indexed in ascending action order
        const uint32_t r = jss_pick(h, (uint32_t)(njobs + (noop ? 1 : 0)));
        if ((int)r == njobs) return iv.J;
        const uint32_t lt = (1u << lane) - 1u;
        uint32_t before = 0u, mine = 0u;
#pragma unroll
        for (int i = 0; i < KJ; i++) {
            before += (uint32_t)__popc(s.L[i] & lt);
            mine += jss_bit(s.L[i], (uint32_t)lane);
        }
        const bool own = r >= before && r < before + mine;
        int act = 0;
        if (own) {
            uint32_t k = r - before;
#pragma unroll
            for (int i = KJ - 1; i >= 0; i--) {          // k-th set slot of this lane
                uint32_t below = 0u;
#pragma unroll
                for (int q = 0; q < KJ; q++) if (q < i) below += jss_bit(s.L[q], (uint32_t)lane);
                if (jss_bit(s.L[i], (uint32_t)lane) && below == k) act = KJ * lane + i;
            }
        }
        const uint32_t who = __ballot_sync(JSS_FULL, own);
        return __shfl_sync(JSS_FULL, act, __ffs((int)who) - 1);
    }
    int best;
    if (rule == JSS_RULE_CR) {                           // dispatching.py:365-408, float64 like Python
        double key = 1.0 / 0.0;
        int kj = 1 << 20;
#pragma unroll
        for (int i = 0; i < KJ; i++) {
            const int j = KJ * lane + i;
            if (jss_bit(s.L[i], (uint32_t)lane)) {
                const double due = (double)iv.len[j] * 1.5;                          // :357-360
                const int remaining = iv.rem[j * (iv.M + 1) + s.todo[i]];            // :387-388
                const double ratio = remaining > 0 ? (due - (double)s.t) / (double)remaining : 1.0 / 0.0;
                if (ratio < key) { key = ratio; kj = j; }                            // strict <, first index wins
            }
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const double ok = __shfl_xor_sync(JSS_FULL, key, off);
            const int oj = __shfl_xor_sync(JSS_FULL, kj, off);
            if (ok < key || (ok == key && oj < kj)) { key = ok; kj = oj; }
        }
        best = kj;
    } else {
        // integer keys; composite (key << 8 | tie) so one REDUX picks value and first index
        const bool minimise = (rule == JSS_RULE_SPT || rule == JSS_RULE_LWR || rule == JSS_RULE_LOR);
        uint32_t comp = minimise ? 0xffffffffu : 0u;
#pragma unroll
        for (int i = 0; i < KJ; i++) {
            const int j = KJ * lane + i;
            if (jss_bit(s.L[i], (uint32_t)lane)) {
                uint32_t key;
                if (rule == JSS_RULE_SPT) key = (uint32_t)jss_op_d(s.op[i]);                    // :105-108
                else if (rule == JSS_RULE_FIFO) key = (uint32_t)s.idle_last[i];                 // :146-148
                else if (rule == JSS_RULE_MWR || rule == JSS_RULE_LWR)
                    key = iv.rem[j * (iv.M + 1) + s.todo[i]];                                   // :188-191 / :231-234
                else key = (uint32_t)(iv.M - s.todo[i]);                                        // :273 / :314
                const uint32_t c = minimise ? ((key << 8) | (uint32_t)j) : ((key << 8) | (uint32_t)(255 - j));
                comp = minimise ? min(comp, c) : max(comp, c);
            }
        }
        comp = minimise ? __reduce_min_sync(JSS_FULL, comp) : __reduce_max_sync(JSS_FULL, comp);
        best = minimise ? (int)(comp & 255u) : 255 - (int)(comp & 255u);
    }
    if (noop && coin_mode == JSS_COIN_DEVICE && h < JSS_COIN_THRESHOLD) return iv.J;   // e.g. :113-114
    return best;
}

// ---- canonical export / import (snapshot & restore; host attribute views) -------------
template <int KJ>
JSS_DEV void env_export(const JssParams &p, const InstView &iv, const EnvRegs<KJ> &s, int env, int lane) {
    const size_t jb = (size_t)env * p.jobs_max;
#pragma unroll
    for (int i = 0; i < KJ; i++) {
        const int j = KJ * lane + i;
        if (j < iv.J) {
            p.x_todo[jb + j] = s.todo[i]; p.x_tufco[jb + j] = s.tufco[i];
            p.x_idle_last[jb + j] = s.idle_last[i]; p.x_total_idle[jb + j] = s.total_idle[i];
            p.x_col4[jb + j] = s.col4[i];
            p.x_legal[jb + j] = (uint8_t)jss_bit(s.L[i], (uint32_t)lane);
            p.x_blocked[jb + j] = (uint8_t)jss_bit(s.B[i], (uint32_t)lane);
        }
    }
    if (lane < iv.M) p.x_tuam[(size_t)env * p.machines_max + lane] = s.tuam;
    if (lane == 0) { p.time[env] = s.t; p.flags[env] = s.flags; }
}

template <int KJ>
JSS_DEV void env_import(const JssParams &p, const InstView &iv, EnvRegs<KJ> &s, int env, int lane) {
    const size_t jb = (size_t)env * p.jobs_max;
#pragma unroll
    for (int i = 0; i < KJ; i++) {
        const int j = KJ * lane + i;
        const bool valid = j < iv.J;
        s.todo[i] = valid ? p.x_todo[jb + j] : iv.M;
        s.tufco[i] = valid ? p.x_tufco[jb + j] : 0;
        s.idle_last[i] = valid ? p.x_idle_last[jb + j] : 0;
        s.total_idle[i] = valid ? p.x_total_idle[jb + j] : 0;
        s.col4[i] = valid ? p.x_col4[jb + j] : 0;
        s.L[i] = __ballot_sync(JSS_FULL, valid && p.x_legal[jb + j] != 0);
        s.B[i] = __ballot_sync(JSS_FULL, valid && p.x_blocked[jb + j] != 0);
    }
    s.tuam = lane < iv.M ? p.x_tuam[(size_t)env * p.machines_max + lane] : 0;
    s.t = p.time[env];
    s.flags = p.flags[env];
    env_derive_ops<KJ>(iv, s, lane);
}

// ---- CTA-level driver -----------------------------------------------------------------------
struct JssSmemLayout {      // element counts; every region starts 16-byte aligned
    int32_t ops_elems, len_elems, rem_elems, scratch_words;
};

JSS_DEV void jss_stage_instance(const JssParams &p, const JssInstDesc &d, uint16_t *sm_ops, int32_t *sm_len,
                                uint16_t *sm_rem, bool want_rem) {
    const int tid = threadIdx.x, nt = blockDim.x;
    {   // pools are padded so whole uint4 copies stay in-bounds
        const uint4 *src = reinterpret_cast<const uint4 *>(p.ops_pool + d.ops_off);
        uint4 *dst = reinterpret_cast<uint4 *>(sm_ops);
        const int n = (d.J * d.M + 7) >> 3;
        for (int k = tid; k < n; k += nt) dst[k] = src[k];
    }
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.len_pool + d.len_off);
        uint4 *dst = reinterpret_cast<uint4 *>(sm_len);
        const int n = (d.J + 3) >> 2;
        for (int k = tid; k < n; k += nt) dst[k] = src[k];
    }
    if (want_rem) {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.rem_pool + d.rem_off);
        uint4 *dst = reinterpret_cast<uint4 *>(sm_rem);
        const int n = (d.J * (d.M + 1) + 7) >> 3;
        for (int k = tid; k < n; k += nt) dst[k] = src[k];
    }
}

template <int KJ, int MODE>
JSS_DEV void jss_process_env(const JssParams &p, const JssLaunch &a, const InstView &iv, int env, int lane,
                             float *scratch) {
    EnvRegs<KJ> s;
    // MODE is a compile-time kernel variant (the hot step kernel carries no policy /
    // rollout / export code); JSS_MODE_RESET instantiates the rarely used rest
    const int mode = (MODE == JSS_MODE_RESET) ? a.mode : MODE;
    if (mode == JSS_MODE_RESET) {
        if (a.env_mask && a.env_mask[env] == 0) return;
        env_reset_regs<KJ>(iv, s, lane);
        if (p.solution) {
            int32_t *sol = p.solution + (size_t)env * p.jobs_max * p.machines_max;
            for (int k = lane; k < p.jobs_max * p.machines_max; k += 32) sol[k] = -1;
        }
        env_store<KJ>(p, env, lane, s);
        env_emit_obs<KJ>(p, iv, s, env, lane, scratch);
        env_emit_mask<KJ>(p, iv, s, env, lane, false);
        env_emit_scalars<KJ>(p, iv, s, env, lane, 0);
        return;
    }
    if (mode == JSS_MODE_IMPORT) {
        if (a.env_mask && a.env_mask[env] == 0) return;
        env_import<KJ>(p, iv, s, env, lane);
        s.ep_steps = 0; s.ep_return = 0;
        env_store<KJ>(p, env, lane, s);
        env_emit_obs<KJ>(p, iv, s, env, lane, scratch);
        env_emit_mask<KJ>(p, iv, s, env, lane, (s.flags & JSS_FLAG_NOOP_LEGAL) != 0u);
        if (lane == 0) p.done[env] = (s.flags & JSS_FLAG_DONE) ? 1 : 0;
        return;
    }
    const uint64_t genv = p.env_id_base + (uint64_t)env;
    if (mode == JSS_MODE_POLICY) env_load_for_policy<KJ>(p, iv, env, lane, s, a.rule);
    else env_load<KJ>(p, iv, env, lane, s);
    if (mode == JSS_MODE_EXPORT) { env_export<KJ>(p, iv, s, env, lane); return; }
    if (mode == JSS_MODE_POLICY) {
        const uint32_t h = jss_hash3(a.seed, genv, a.step_index);
        const int act = env_select_action<KJ>(iv, s, lane, a.rule, a.coin_mode, h);
        if (lane == 0) a.actions_out[env] = act;
        return;
    }
    if (mode == JSS_MODE_STEP) {
        const int action = a.actions[env];
        int raw = 0;
        const uint32_t flags_in = s.flags;
        const bool changed = env_step<KJ>(p, iv, s, env, lane, action, raw);
        if (changed) {
            env_store<KJ>(p, env, lane, s);
            env_emit_obs<KJ>(p, iv, s, env, lane, scratch);
            env_emit_mask<KJ>(p, iv, s, env, lane, (s.flags & JSS_FLAG_NOOP_LEGAL) != 0u);
            env_emit_scalars<KJ>(p, iv, s, env, lane, raw);
        } else if (s.flags != flags_in) {                // only the sticky error bit changed
            if (lane == 0) {
                p.state[(size_t)env * p.block_words + 5 * p.Jcap + p.Mcap + 8 + JSS_HDR_FLAGS] = (int32_t)s.flags;
                p.flags[env] = s.flags; p.reward[env] = 0.0f; p.reward_raw[env] = 0;
            }
        }
        return;
    }
    // JSS_MODE_ROLLOUT: n_steps x (policy -> step) with the state held in registers
    bool dirty = false;
    int raw = 0;
    for (int k = 0; k < a.n_steps; k++) {
        const uint32_t h = jss_hash3(a.seed, genv, a.step_index + (uint64_t)k);
        const int act = env_select_action<KJ>(iv, s, lane, a.rule, a.coin_mode, h);
        int r = 0;
        const bool changed = env_step<KJ>(p, iv, s, env, lane, act, r);
        if (changed) {
            raw = r; dirty = true;
            if (a.write_obs) {
                env_emit_obs<KJ>(p, iv, s, env, lane, scratch);
                env_emit_mask<KJ>(p, iv, s, env, lane, (s.flags & JSS_FLAG_NOOP_LEGAL) != 0u);
                env_emit_scalars<KJ>(p, iv, s, env, lane, r);
            }
        }
    }
    if (dirty) {
        env_store<KJ>(p, env, lane, s);
        if (!a.write_obs) {
            env_emit_obs<KJ>(p, iv, s, env, lane, scratch);
            env_emit_mask<KJ>(p, iv, s, env, lane, (s.flags & JSS_FLAG_NOOP_LEGAL) != 0u);
            env_emit_scalars<KJ>(p, iv, s, env, lane, raw);
        }
    }
}

template <int KJ, int MODE>
__global__ void __launch_bounds__(JSS_WARPS_PER_CTA * 32)
jss_env_kernel(const JssParams p, const JssLaunch a, const JssSmemLayout sl) {
    JSS_SMEM_DECL(jss_smem);
    uint16_t *sm_ops = reinterpret_cast<uint16_t *>(jss_smem);
    int32_t *sm_len = reinterpret_cast<int32_t *>(sm_ops + sl.ops_elems);
    uint16_t *sm_rem = reinterpret_cast<uint16_t *>(sm_len + sl.len_elems);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *scratch = reinterpret_cast<float *>(sm_rem + sl.rem_elems) + (size_t)warp * sl.scratch_words;
    const bool want_rem = (MODE == JSS_MODE_POLICY || MODE == JSS_MODE_ROLLOUT) &&
                          (a.rule == JSS_RULE_MWR || a.rule == JSS_RULE_LWR || a.rule == JSS_RULE_CR);
    int staged = -1;
    InstView iv;
    iv.ops = sm_ops; iv.len = sm_len; iv.rem = sm_rem;
    iv.J = iv.M = iv.max_time_op = iv.max_time_jobs = iv.sum_op = 0;
    for (int tile = a.tile_begin + (int)blockIdx.x; tile < a.tile_end; tile += (int)gridDim.x) {
        const JssTile td = p.tiles[tile];
        const int inst = td.inst_count >> 8, count = td.inst_count & 255;
        if (inst != staged) {                            // CTA-uniform
            __syncthreads();
            const JssInstDesc d = p.inst[inst];
            jss_stage_instance(p, d, sm_ops, sm_len, sm_rem, want_rem);
            iv.J = d.J; iv.M = d.M; iv.max_time_op = d.max_time_op; iv.max_time_jobs = d.max_time_jobs;
            iv.sum_op = d.sum_op;
            staged = inst;
            __syncthreads();
        }
        if (warp < count) jss_process_env<KJ, MODE>(p, a, iv, p.order[td.first + warp], lane, scratch);
    }
}

// ---- per-shard statistics (one block; N is at most a few 100k) ---------------------------
__global__ void jss_stats_kernel(const JssParams p, unsigned long long *out) {
    // out[0..7] pre-initialised by the host: sums 0, min = ~0ull
    unsigned long long ep = 0, steps = 0, smk = 0, ndone = 0, nerr = 0;
    long long sret = 0;
    unsigned long long mn = ~0ull, mx = 0;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < p.n_envs; e += gridDim.x * blockDim.x) {
        const int64_t *acc = p.acc + (size_t)e * 4;
        const int32_t *hdr = p.state + (size_t)e * p.block_words + 5 * p.Jcap + p.Mcap + 8;
        ep += (unsigned long long)p.episode_count[e];
        steps += (unsigned long long)acc[0] + (unsigned long long)(((uint32_t)hdr[JSS_HDR_FLAGS] & JSS_FLAG_DONE) ? 0 : hdr[JSS_HDR_EP_STEPS]);
        smk += (unsigned long long)acc[1];
        sret += acc[2];
        if (p.episode_count[e] > 0) {
            const uint64_t mm = (uint64_t)acc[3];
            mn = min(mn, (unsigned long long)(mm >> 32));
            mx = max(mx, (unsigned long long)(mm & 0xffffffffull));
        }
        ndone += ((uint32_t)hdr[JSS_HDR_FLAGS] & JSS_FLAG_DONE) ? 1 : 0;
        nerr += ((uint32_t)hdr[JSS_HDR_FLAGS] & JSS_FLAG_ERROR) ? 1 : 0;
    }
    atomicAdd(&out[0], ep);
    atomicAdd(&out[1], steps);
    atomicAdd(&out[2], smk);
    atomicMin(&out[3], mn);
    atomicMax(&out[4], mx);
    atomicAdd(&out[5], (unsigned long long)sret);
    atomicAdd(&out[6], ndone);
    atomicAdd(&out[7], nerr);
}
