/*
 * jss_oracle.c -- CPU restatement of the reference job-shop environment.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this file's shared
 * object.  The product path (jssenv_b200/, the C-ABI in include/jss_b200.h) never
 * links, imports or falls back to it.
 *
 * What it restates (all citations relative to /root/reference):
 *   JSSEnv/envs/jss_env.py:72-95    instance parse / derived scalars  -> jsso_create
 *   JSSEnv/envs/jss_env.py:121-134  _get_current_state_representation -> jsso_observe
 *   JSSEnv/envs/jss_env.py:145-181  reset                             -> jsso_reset
 *   JSSEnv/envs/jss_env.py:183-254  _prioritization_non_final         -> prioritization_non_final
 *   JSSEnv/envs/jss_env.py:256-401  _check_no_op                      -> check_no_op
 *   JSSEnv/envs/jss_env.py:403-481  step                              -> jsso_step
 *   JSSEnv/envs/jss_env.py:483-493  _reward_scaler                    -> (inline, reward / max_time_op)
 *   JSSEnv/envs/jss_env.py:495-637  increase_time_step                -> jsso_increase_time_step
 *   JSSEnv/envs/jss_env.py:639-653  _is_done                          -> is_done
 *   JSSEnv/dispatching.py:92-408    SPT/FIFO/MWR/LWR/MOR/LOR/CR       -> jsso_rule_action
 *
 * The restatement is deliberately LITERAL: it keeps the reference's redundant
 * state (sorted event list, illegal_actions[M][J] matrix, running counters, the
 * incrementally-updated float64 `state` matrix) and its loop order, so that it is
 * an independent check of the reduced-state CUDA kernels.  Integers are int64
 * (NumPy `int`), observations/rewards are float64, exactly as in the reference.
 *
 * Parity pinning: tests/test_oracle_golden.py checks this file against (i) the
 * reference's 12 optimal-makespan replays (tests/test_solutions.py) and (ii)
 * step-by-step traces recorded from the unmodified Python reference by
 * oracle/gen_golden.py (fixtures in tests/golden/).
 *
 * Python exceptions of the reference are mapped to return codes:
 *   JSSO_ERR_EMPTY_QUEUE  list.pop(0) on an empty event list (jss_env.py:517)
 *   JSSO_ERR_JOB_FINISHED instance_matrix[action][M] IndexError (jss_env.py:444)
 */
#define _POSIX_C_SOURCE 200809L
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define JSSO_OK 0
#define JSSO_ERR_EMPTY_QUEUE (-1)
#define JSSO_ERR_JOB_FINISHED (-2)
#define JSSO_ERR_BAD_ARG (-3)

typedef struct jsso {
    int jobs, machines;
    double cr_due_date_factor; /* CriticalRatio.__init__(due_date_factor=1.5), dispatching.py:337-349 */
    int64_t *inst_machine;   /* instance_matrix[j][op][0]  [J*M] */
    int64_t *inst_time;      /* instance_matrix[j][op][1]  [J*M] */
    int64_t *jobs_length;    /* [J] */
    int64_t max_time_op, max_time_jobs, sum_op;
    int64_t nb_legal_actions, nb_machine_legal;
    int64_t *solution;       /* [J*M], -1 = unscheduled */
    int64_t last_time_step;  /* -1 stands for float('inf') */
    int64_t current_time_step;
    int64_t *next_time_step; /* sorted unique event list */
    int64_t *next_jobs;
    int n_next;
    uint8_t *legal_actions;  /* [J+1] */
    int64_t *time_until_available_machine;      /* [M] */
    int64_t *time_until_finish_current_op_jobs; /* [J] */
    int64_t *todo_time_step_job;                /* [J] */
    int64_t *total_perform_op_time_jobs;        /* [J] */
    int64_t *needed_machine_jobs;               /* [J] */
    int64_t *total_idle_time_jobs;              /* [J] */
    int64_t *idle_time_jobs_last_op;            /* [J] */
    uint8_t *illegal_actions;                   /* [M*J] */
    uint8_t *action_illegal_no_op;              /* [J] */
    uint8_t *machine_legal;                     /* [M] */
    double *state;                              /* [J*7] */
    int64_t *scratch_horizon;                   /* [M] */
    uint8_t *scratch_set;                       /* [M] */
} jsso;

#define IM(o, j, op) ((o)->inst_machine[(size_t)(j) * (o)->machines + (op)])
#define IT(o, j, op) ((o)->inst_time[(size_t)(j) * (o)->machines + (op)])
#define ST(o, j, c) ((o)->state[(size_t)(j) * 7 + (c)])

static void *zalloc(size_t n) { return calloc(n ? n : 1, 1); }

/* jss_env.py:72-95 -- the text parse itself is host Python; this takes the
 * parsed (machine, time) pairs and derives jobs_length / max_time_op /
 * max_time_jobs / sum_op the way the parse loop does (lines 86-89). */
jsso *jsso_create(int jobs, int machines, const int32_t *machine, const int32_t *time) {
    if (jobs <= 0 || machines <= 1) return NULL; /* asserts at jss_env.py:93-94 */
    jsso *o = (jsso *)zalloc(sizeof(jsso));
    size_t J = (size_t)jobs, M = (size_t)machines;
    o->jobs = jobs;
    o->machines = machines;
    o->cr_due_date_factor = 1.5;
    o->inst_machine = (int64_t *)zalloc(J * M * 8);
    o->inst_time = (int64_t *)zalloc(J * M * 8);
    o->jobs_length = (int64_t *)zalloc(J * 8);
    for (size_t j = 0; j < J; j++)
        for (size_t i = 0; i < M; i++) {
            int64_t m = machine[j * M + i], t = time[j * M + i];
            o->inst_machine[j * M + i] = m;
            o->inst_time[j * M + i] = t;
            if (t > o->max_time_op) o->max_time_op = t; /* :86 */
            o->jobs_length[j] += t;                     /* :87 */
            o->sum_op += t;                             /* :88 */
        }
    for (size_t j = 0; j < J; j++)
        if (o->jobs_length[j] > o->max_time_jobs) o->max_time_jobs = o->jobs_length[j]; /* :89 */
    o->solution = (int64_t *)zalloc(J * M * 8);
    o->next_time_step = (int64_t *)zalloc((J * M + 2) * 8);
    o->next_jobs = (int64_t *)zalloc((J * M + 2) * 8);
    o->legal_actions = (uint8_t *)zalloc(J + 1);
    o->time_until_available_machine = (int64_t *)zalloc(M * 8);
    o->time_until_finish_current_op_jobs = (int64_t *)zalloc(J * 8);
    o->todo_time_step_job = (int64_t *)zalloc(J * 8);
    o->total_perform_op_time_jobs = (int64_t *)zalloc(J * 8);
    o->needed_machine_jobs = (int64_t *)zalloc(J * 8);
    o->total_idle_time_jobs = (int64_t *)zalloc(J * 8);
    o->idle_time_jobs_last_op = (int64_t *)zalloc(J * 8);
    o->illegal_actions = (uint8_t *)zalloc(M * J);
    o->action_illegal_no_op = (uint8_t *)zalloc(J);
    o->machine_legal = (uint8_t *)zalloc(M);
    o->state = (double *)zalloc(J * 7 * 8);
    o->scratch_horizon = (int64_t *)zalloc(M * 8);
    o->scratch_set = (uint8_t *)zalloc(M);
    o->last_time_step = -1;    /* float('inf') at :53 */
    o->current_time_step = -1; /* float('inf') at :54 */
    return o;
}

void jsso_destroy(jsso *o) {
    if (!o) return;
    free(o->inst_machine); free(o->inst_time); free(o->jobs_length); free(o->solution);
    free(o->next_time_step); free(o->next_jobs); free(o->legal_actions);
    free(o->time_until_available_machine); free(o->time_until_finish_current_op_jobs);
    free(o->todo_time_step_job); free(o->total_perform_op_time_jobs);
    free(o->needed_machine_jobs); free(o->total_idle_time_jobs);
    free(o->idle_time_jobs_last_op); free(o->illegal_actions);
    free(o->action_illegal_no_op); free(o->machine_legal); free(o->state);
    free(o->scratch_horizon); free(o->scratch_set);
    free(o);
}

/* jss_env.py:121-134 */
static void get_current_state_representation(jsso *o) {
    for (int j = 0; j < o->jobs; j++) ST(o, j, 0) = o->legal_actions[j] ? 1.0 : 0.0; /* :130 */
}

/* jss_env.py:145-181 */
void jsso_reset(jsso *o) {
    int J = o->jobs, M = o->machines;
    o->current_time_step = 0;                  /* :154 */
    o->n_next = 0;                             /* :155-156 */
    o->nb_legal_actions = J;                   /* :157 */
    o->nb_machine_legal = 0;                   /* :158 */
    memset(o->legal_actions, 1, (size_t)J + 1); /* :160 */
    o->legal_actions[J] = 0;                   /* :161 */
    for (size_t i = 0; i < (size_t)J * M; i++) o->solution[i] = -1; /* :163 */
    memset(o->time_until_available_machine, 0, (size_t)M * 8);
    memset(o->time_until_finish_current_op_jobs, 0, (size_t)J * 8);
    memset(o->todo_time_step_job, 0, (size_t)J * 8);
    memset(o->total_perform_op_time_jobs, 0, (size_t)J * 8);
    memset(o->needed_machine_jobs, 0, (size_t)J * 8);
    memset(o->total_idle_time_jobs, 0, (size_t)J * 8);
    memset(o->idle_time_jobs_last_op, 0, (size_t)J * 8);
    memset(o->illegal_actions, 0, (size_t)M * J);
    memset(o->action_illegal_no_op, 0, (size_t)J);
    memset(o->machine_legal, 0, (size_t)M);
    for (int job = 0; job < J; job++) {        /* :174-179 */
        int64_t needed_machine = IM(o, job, 0);
        o->needed_machine_jobs[job] = needed_machine;
        if (!o->machine_legal[needed_machine]) {
            o->machine_legal[needed_machine] = 1;
            o->nb_machine_legal += 1;
        }
    }
    memset(o->state, 0, (size_t)J * 7 * 8);    /* :180 */
    get_current_state_representation(o);       /* :181 */
}

/* jss_env.py:183-254 */
static void prioritization_non_final(jsso *o) {
    int J = o->jobs, M = o->machines;
    if (o->nb_machine_legal < 1) return; /* :202 */
    for (int machine = 0; machine < M; machine++) {
        if (!o->machine_legal[machine]) continue; /* :204 */
        int n_non_final = 0;
        int64_t min_non_final = INT64_MAX; /* float('inf') :208 */
        /* first sweep (:211-239): classify; final jobs are revisited below in
         * ascending order exactly like the `final_job` list */
        for (int job = 0; job < J; job++) {
            if (o->needed_machine_jobs[job] == machine && o->legal_actions[job]) {
                if (o->todo_time_step_job[job] == M - 1) {
                    /* final_job.append(job) :218 */
                } else {
                    int64_t cur = o->todo_time_step_job[job];
                    int64_t time_needed_legal = IT(o, job, cur);
                    int64_t machine_needed_nextstep = IM(o, job, cur + 1);
                    if (o->time_until_available_machine[machine_needed_nextstep] == 0) { /* :234-236 */
                        if (time_needed_legal < min_non_final) min_non_final = time_needed_legal;
                        n_non_final++;
                    }
                }
            }
        }
        if (n_non_final > 0) { /* :243 */
            for (int job = 0; job < J; job++) {
                /* membership in final_job was decided BEFORE any legality change;
                 * de-legalising a final job cannot change another job's class */
                if (o->needed_machine_jobs[job] == machine && o->legal_actions[job] &&
                    o->todo_time_step_job[job] == M - 1) {
                    int64_t time_needed_legal = IT(o, job, o->todo_time_step_job[job]);
                    if (time_needed_legal > min_non_final) { /* :252 */
                        o->legal_actions[job] = 0;
                        o->nb_legal_actions -= 1;
                    }
                }
            }
        }
    }
}

/* jss_env.py:256-401 */
static void check_no_op(jsso *o) {
    int J = o->jobs, M = o->machines;
    o->legal_actions[J] = 0; /* :278 */
    if (!(o->n_next > 0 && o->nb_machine_legal <= 3 && o->nb_legal_actions <= 4)) return; /* :284-288 */
    uint8_t *machine_next = o->scratch_set; /* set() :290 */
    int64_t machine_next_len = 0;
    memset(machine_next, 0, (size_t)M);
    int64_t next_time_step = o->next_time_step[0]; /* :293 */
    int64_t max_horizon = o->current_time_step;    /* :296 */
    int64_t *max_horizon_machine = o->scratch_horizon;
    for (int m = 0; m < M; m++) max_horizon_machine[m] = o->current_time_step + o->max_time_op; /* :300-302 */
    for (int job = 0; job < J; job++) { /* :305-321 */
        if (o->legal_actions[job]) {
            int64_t time_step = o->todo_time_step_job[job];
            int64_t machine_needed = IM(o, job, time_step);
            int64_t time_needed = IT(o, job, time_step);
            int64_t end_job = o->current_time_step + time_needed;
            if (end_job < next_time_step) return; /* :314-315 */
            if (end_job < max_horizon_machine[machine_needed]) max_horizon_machine[machine_needed] = end_job;
            if (max_horizon_machine[machine_needed] > max_horizon) max_horizon = max_horizon_machine[machine_needed];
        }
    }
    for (int job = 0; job < J; job++) { /* :324-401 */
        if (o->legal_actions[job]) continue;
        int64_t time_step, time_needed;
        if (o->time_until_finish_current_op_jobs[job] > 0 &&
            o->todo_time_step_job[job] + 1 < M) { /* :327-330 */
            time_step = o->todo_time_step_job[job] + 1;
            time_needed = o->current_time_step + o->time_until_finish_current_op_jobs[job];
        } else if (!o->action_illegal_no_op[job] && o->todo_time_step_job[job] < M) { /* :366-369 */
            time_step = o->todo_time_step_job[job];
            int64_t machine_needed = IM(o, job, time_step);
            time_needed = o->current_time_step + o->time_until_available_machine[machine_needed];
        } else {
            continue;
        }
        while (time_step < M - 1 && max_horizon > time_needed) { /* :340-342 / :380-382 */
            int64_t machine_needed = IM(o, job, time_step);
            if (max_horizon_machine[machine_needed] > time_needed && o->machine_legal[machine_needed]) {
                if (!machine_next[machine_needed]) { machine_next[machine_needed] = 1; machine_next_len++; }
                if (machine_next_len == o->nb_machine_legal) { /* :357 / :395 */
                    o->legal_actions[J] = 1;
                    return;
                }
            }
            time_needed += IT(o, job, time_step);
            time_step += 1;
        }
    }
}

/* jss_env.py:495-637.  *hole receives hole_planning. */
int jsso_increase_time_step(jsso *o, int64_t *hole) {
    int J = o->jobs, M = o->machines;
    int64_t hole_planning = 0;
    if (o->n_next == 0) return JSSO_ERR_EMPTY_QUEUE; /* list.pop(0) IndexError :517 */
    int64_t next_time_step_to_pick = o->next_time_step[0];
    memmove(o->next_time_step, o->next_time_step + 1, (size_t)(o->n_next - 1) * 8);
    memmove(o->next_jobs, o->next_jobs + 1, (size_t)(o->n_next - 1) * 8);
    o->n_next--;
    int64_t difference = next_time_step_to_pick - o->current_time_step; /* :521 */
    o->current_time_step = next_time_step_to_pick;
    for (int job = 0; job < J; job++) { /* :525-601 */
        int64_t was_left_time = o->time_until_finish_current_op_jobs[job];
        if (was_left_time > 0) {
            int64_t performed_op_job = difference < was_left_time ? difference : was_left_time;
            int64_t left = o->time_until_finish_current_op_jobs[job] - difference;
            o->time_until_finish_current_op_jobs[job] = left > 0 ? left : 0;
            ST(o, job, 1) = (double)o->time_until_finish_current_op_jobs[job] / (double)o->max_time_op;
            o->total_perform_op_time_jobs[job] += performed_op_job;
            ST(o, job, 3) = (double)o->total_perform_op_time_jobs[job] / (double)o->max_time_jobs;
            if (o->time_until_finish_current_op_jobs[job] == 0) {
                o->total_idle_time_jobs[job] += difference - was_left_time;
                ST(o, job, 6) = (double)o->total_idle_time_jobs[job] / (double)o->sum_op;
                o->idle_time_jobs_last_op[job] = difference - was_left_time;
                ST(o, job, 5) = (double)o->idle_time_jobs_last_op[job] / (double)o->sum_op;
                o->todo_time_step_job[job] += 1;
                ST(o, job, 2) = (double)o->todo_time_step_job[job] / (double)M;
                if (o->todo_time_step_job[job] < M) {
                    o->needed_machine_jobs[job] = IM(o, job, o->todo_time_step_job[job]);
                    int64_t w = o->time_until_available_machine[o->needed_machine_jobs[job]] - difference;
                    ST(o, job, 4) = (double)(w > 0 ? w : 0) / (double)o->max_time_op; /* :569-578 */
                } else {
                    o->needed_machine_jobs[job] = -1;
                    ST(o, job, 4) = 1.0; /* :586 */
                    if (o->legal_actions[job]) {
                        o->legal_actions[job] = 0;
                        o->nb_legal_actions -= 1;
                    }
                }
            }
        } else if (o->todo_time_step_job[job] < M) { /* :594 */
            o->total_idle_time_jobs[job] += difference;
            o->idle_time_jobs_last_op[job] += difference;
            ST(o, job, 5) = (double)o->idle_time_jobs_last_op[job] / (double)o->sum_op;
            ST(o, job, 6) = (double)o->total_idle_time_jobs[job] / (double)o->sum_op;
        }
    }
    for (int machine = 0; machine < M; machine++) { /* :604-634 */
        if (o->time_until_available_machine[machine] < difference) {
            int64_t empty = difference - o->time_until_available_machine[machine];
            hole_planning += empty;
        }
        int64_t left = o->time_until_available_machine[machine] - difference;
        o->time_until_available_machine[machine] = left > 0 ? left : 0;
        if (o->time_until_available_machine[machine] == 0) {
            for (int job = 0; job < J; job++) {
                if (o->needed_machine_jobs[job] == machine && !o->legal_actions[job] &&
                    !o->illegal_actions[(size_t)machine * J + job]) {
                    o->legal_actions[job] = 1;
                    o->nb_legal_actions += 1;
                    if (!o->machine_legal[machine]) {
                        o->machine_legal[machine] = 1;
                        o->nb_machine_legal += 1;
                    }
                }
            }
        }
    }
    *hole = hole_planning;
    return JSSO_OK;
}

/* jss_env.py:639-653 */
static int is_done(jsso *o) {
    if (o->nb_legal_actions == 0) {
        o->last_time_step = o->current_time_step;
        return 1;
    }
    return 0;
}

/* jss_env.py:403-481.  Outputs: scaled reward (float64), raw reward (the float
 * `reward` before _reward_scaler; always integer-valued), done. */
int jsso_step(jsso *o, int action, double *scaled_reward, int64_t *raw_reward, int *done) {
    int J = o->jobs, M = o->machines;
    double reward = 0.0; /* :418 */
    int64_t hole;
    int rc;
    if (action < 0 || action > J) return JSSO_ERR_BAD_ARG;
    if (action == J) { /* :419-440 */
        o->nb_machine_legal = 0;
        o->nb_legal_actions = 0;
        for (int job = 0; job < J; job++) {
            if (o->legal_actions[job]) {
                o->legal_actions[job] = 0;
                int64_t needed_machine = o->needed_machine_jobs[job];
                o->machine_legal[needed_machine] = 0;
                o->illegal_actions[(size_t)needed_machine * J + job] = 1;
                o->action_illegal_no_op[job] = 1;
            }
        }
        while (o->nb_machine_legal == 0) { /* :429-430 (no empty-queue guard) */
            rc = jsso_increase_time_step(o, &hole);
            if (rc != JSSO_OK) return rc;
            reward -= (double)hole;
        }
        *scaled_reward = reward / (double)o->max_time_op; /* :431 */
        prioritization_non_final(o);
        check_no_op(o);
    } else { /* :441-481 */
        int64_t current_time_step_job = o->todo_time_step_job[action];
        if (current_time_step_job >= M) return JSSO_ERR_JOB_FINISHED; /* IndexError :444 */
        int64_t machine_needed = o->needed_machine_jobs[action];
        int64_t time_needed = IT(o, action, current_time_step_job);
        reward += (double)time_needed;
        o->time_until_available_machine[machine_needed] = time_needed;
        o->time_until_finish_current_op_jobs[action] = time_needed;
        ST(o, action, 1) = (double)time_needed / (double)o->max_time_op; /* :448 */
        int64_t to_add_time_step = o->current_time_step + time_needed;
        int present = 0, index = 0; /* :450-453: membership test + bisect_left */
        while (index < o->n_next && o->next_time_step[index] < to_add_time_step) index++;
        if (index < o->n_next && o->next_time_step[index] == to_add_time_step) present = 1;
        if (!present) {
            memmove(o->next_time_step + index + 1, o->next_time_step + index, (size_t)(o->n_next - index) * 8);
            memmove(o->next_jobs + index + 1, o->next_jobs + index, (size_t)(o->n_next - index) * 8);
            o->next_time_step[index] = to_add_time_step;
            o->next_jobs[index] = action;
            o->n_next++;
        }
        o->solution[(size_t)action * M + current_time_step_job] = o->current_time_step; /* :454 */
        for (int job = 0; job < J; job++) { /* :455-461 */
            if (o->needed_machine_jobs[job] == machine_needed && o->legal_actions[job]) {
                o->legal_actions[job] = 0;
                o->nb_legal_actions -= 1;
            }
        }
        o->nb_machine_legal -= 1; /* :462 */
        o->machine_legal[machine_needed] = 0;
        for (int job = 0; job < J; job++) { /* :464-467 */
            if (o->illegal_actions[(size_t)machine_needed * J + job]) {
                o->action_illegal_no_op[job] = 0;
                o->illegal_actions[(size_t)machine_needed * J + job] = 0;
            }
        }
        while (o->nb_machine_legal == 0 && o->n_next > 0) { /* :469-470 */
            rc = jsso_increase_time_step(o, &hole);
            if (rc != JSSO_OK) return rc;
            reward -= (double)hole;
        }
        prioritization_non_final(o); /* :471 */
        check_no_op(o);              /* :472 */
        *scaled_reward = reward / (double)o->max_time_op; /* :474 */
    }
    *raw_reward = (int64_t)reward;
    get_current_state_representation(o);
    *done = is_done(o);
    return JSSO_OK;
}

/* ------------------------------------------------------------------------
 * Dispatching rules, JSSEnv/dispatching.py.  `rule`: 0 SPT (92-116), 1 FIFO
 * (133-156), 2 MWR (173-199), 3 LWR (216-242), 4 MOR (259-283), 5 LOR (300-324),
 * 6 CR (365-408).  The reference draws np.random.random() ONLY when the no-op is
 * legal (short-circuit `and`, e.g. :113); the caller supplies that uniform in
 * `u` and learns through *consumed whether the reference would have drawn it.
 * ------------------------------------------------------------------------ */
int jsso_rule_action(jsso *o, int rule, double u, int *consumed) {
    int J = o->jobs, M = o->machines;
    *consumed = 0;
    int sum = 0;
    for (int i = 0; i <= J; i++) sum += o->legal_actions[i];
    if (sum == 1 && o->legal_actions[J]) return J; /* e.g. :96-97 */
    int best_job = -1;
    double best = 0.0;
    int have = 0;
    for (int job = 0; job < J; job++) {
        if (!o->legal_actions[job]) continue;
        int64_t todo = o->todo_time_step_job[job];
        double key; /* all integer keys are exactly representable */
        int minimise;
        int64_t remaining = 0;
        switch (rule) {
        case 0: key = (double)IT(o, job, todo); minimise = 1; break;          /* :105-108 */
        case 1: key = (double)o->idle_time_jobs_last_op[job]; minimise = 0; break; /* :146-148 */
        case 2: case 3:
            for (int64_t op = todo; op < M; op++) remaining += IT(o, job, op); /* :188-189 / :231-232 */
            key = (double)remaining; minimise = (rule == 3); break;
        case 4: case 5:
            key = (double)(M - todo); minimise = (rule == 5); break;           /* :273 / :314 */
        case 6: {
            int64_t total_time = 0;
            for (int op = 0; op < M; op++) total_time += IT(o, job, op);      /* :357 */
            double due_date = (double)total_time * o->cr_due_date_factor;      /* :360 */
            for (int64_t op = todo; op < M; op++) remaining += IT(o, job, op); /* :387-388 */
            double time_remaining = due_date - (double)o->current_time_step;   /* :391 */
            if (remaining > 0) key = time_remaining / (double)remaining;       /* :396 */
            else key = 1.0 / 0.0;                                              /* :398 */
            minimise = 1; break;
        }
        default: return -1;
        }
        /* strict comparisons from +inf / -1 starts: first index wins ties */
        if (minimise) {
            if (!have) { if (key < 1.0 / 0.0) { best = key; best_job = job; have = 1; } }
            else if (key < best) { best = key; best_job = job; }
        } else {
            if (!have) { if (key > -1.0) { best = key; best_job = job; have = 1; } }
            else if (key > best) { best = key; best_job = job; }
        }
    }
    if (o->legal_actions[J]) { /* e.g. :113-114 */
        *consumed = 1;
        if (u < 0.1) return J;
    }
    return best_job;
}

/* ---- accessors for the ctypes wrapper --------------------------------- */
void jsso_set_cr_due_date_factor(jsso *o, double f) { o->cr_due_date_factor = f; }
int jsso_jobs(const jsso *o) { return o->jobs; }
int jsso_machines(const jsso *o) { return o->machines; }
int64_t jsso_scalar(const jsso *o, int which) {
    switch (which) {
    case 0: return o->max_time_op;
    case 1: return o->max_time_jobs;
    case 2: return o->sum_op;
    case 3: return o->nb_legal_actions;
    case 4: return o->nb_machine_legal;
    case 5: return o->current_time_step;
    case 6: return o->last_time_step;
    case 7: return o->n_next;
    default: return -1;
    }
}
const void *jsso_array(const jsso *o, int which) {
    switch (which) {
    case 0: return o->legal_actions;
    case 1: return o->state;
    case 2: return o->time_until_available_machine;
    case 3: return o->time_until_finish_current_op_jobs;
    case 4: return o->todo_time_step_job;
    case 5: return o->total_perform_op_time_jobs;
    case 6: return o->needed_machine_jobs;
    case 7: return o->total_idle_time_jobs;
    case 8: return o->idle_time_jobs_last_op;
    case 9: return o->illegal_actions;
    case 10: return o->action_illegal_no_op;
    case 11: return o->machine_legal;
    case 12: return o->solution;
    case 13: return o->next_time_step;
    case 14: return o->jobs_length;
    case 15: return o->inst_machine;
    case 16: return o->inst_time;
    default: return NULL;
    }
}

/* ------------------------------------------------------------------------
 * Bulk driver used by bench.py's CPU legs: runs `n_steps` env steps with the
 * masked-uniform random policy (reference idiom README.md:58-60: uniform over
 * the set bits of action_mask), auto-reset on done.  The RNG is the same
 * counter hash the CUDA policy kernel uses (jssenv_b200/csrc/jss_rng.h restated
 * here so the oracle stays self-contained).  Returns the number of completed
 * episodes; *sum_makespan accumulates their makespans.
 * ------------------------------------------------------------------------ */
static inline uint32_t jsso_fold32(uint64_t v) { return (uint32_t)v ^ ((uint32_t)(v >> 32) * 0x7FEB352Du); }
static inline uint32_t jsso_hash3(uint64_t seed, uint64_t env, uint64_t ctr) {
    uint32_t h = jsso_fold32(seed) ^ (jsso_fold32(env) * 0x9E3779B1u + 0x85EBCA77u) ^
                 (jsso_fold32(ctr) * 0xC2B2AE3Du + 0x27D4EB2Fu);
    h ^= h >> 16; h *= 0x85EBCA6Bu;
    h ^= h >> 13; h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h;
}

int jsso_masked_random_action(const jsso *o, uint64_t seed, uint64_t env, uint64_t ctr) {
    int J = o->jobs, cnt = 0;
    for (int i = 0; i <= J; i++) cnt += o->legal_actions[i];
    if (cnt == 0) return -1;
    uint32_t r = (uint32_t)(((uint64_t)jsso_hash3(seed, env, ctr) * (uint64_t)cnt) >> 32);
    for (int i = 0; i <= J; i++)
        if (o->legal_actions[i]) { if (r == 0) return i; r--; }
    return -1;
}

int64_t jsso_run_random(jsso *o, uint64_t seed, uint64_t env, int64_t n_steps,
                        int64_t *episodes, int64_t *sum_makespan) {
    int64_t steps = 0, ctr = 0;
    double r; int64_t raw; int done = 0;
    jsso_reset(o);
    while (steps < n_steps) {
        int a = jsso_masked_random_action(o, seed, env, (uint64_t)ctr);
        ctr++;
        if (a < 0 || jsso_step(o, a, &r, &raw, &done) != JSSO_OK) return -1;
        steps++;
        if (done) {
            *episodes += 1;
            *sum_makespan += o->current_time_step;
            jsso_reset(o);
        }
    }
    return steps;
}

/* Same driver, bounded by wall time instead of a step count (bench.py's CPU legs run one
 * of these per host thread for a fixed window).  Returns the steps executed. */
int64_t jsso_run_random_timed(jsso *o, uint64_t seed, uint64_t env, double seconds,
                              int64_t *episodes, int64_t *sum_makespan) {
    struct timespec t0, t1;
    int64_t steps = 0, ctr = 0;
    double r; int64_t raw; int done = 0;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    jsso_reset(o);
    for (;;) {
        for (int k = 0; k < 512; k++) {
            int a = jsso_masked_random_action(o, seed, env, (uint64_t)ctr);
            ctr++;
            if (a < 0 || jsso_step(o, a, &r, &raw, &done) != JSSO_OK) return -1;
            steps++;
            if (done) {
                *episodes += 1;
                *sum_makespan += o->current_time_step;
                jsso_reset(o);
            }
        }
        clock_gettime(CLOCK_MONOTONIC, &t1);
        if ((double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec) >= seconds) break;
    }
    return steps;
}
