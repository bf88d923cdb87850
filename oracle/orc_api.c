/*
 * orc_api.c -- TEST INFRASTRUCTURE.  Flat entry points for the ctypes binding (oracle/pyoracle.py):
 * choose a draw source by kind, optionally record every draw into orc_tables, collect emitted rows.
 */
#include "bayesrr_oracle.h"
#include <string.h>
#include <time.h>

typedef struct { double *rows; int64_t max_rows, n_rows; } collect_t;
static void collect_sink(void *ctx, const double *row, int64_t len)
{
    collect_t *c = (collect_t *)ctx;
    if (c->rows && c->n_rows < c->max_rows) memcpy(c->rows + c->n_rows * len, row, (size_t)len * 8);
    c->n_rows++;
}

enum { SRC_PHILOX = 0, SRC_SEQ = 1, SRC_REPLAY = 2 };

typedef struct { orc_philox px; orc_seq sq; orc_recorder rec; orc_draws d; } src_t;
static const orc_draws *make_src(src_t *s, int kind, uint64_t seed, orc_tables *tbl, int record)
{
    if (kind == SRC_PHILOX) { orc_philox_init(&s->px, seed); s->d = orc_philox_source(&s->px); }
    else if (kind == SRC_SEQ) { orc_seq_init(&s->sq, seed); s->d = orc_seq_source(&s->sq); }
    else { s->d = orc_replay_source(tbl); return &s->d; }
    if (record && tbl) { s->rec.inner = s->d; s->rec.t = tbl; s->d = orc_record_source(&s->rec); }
    return &s->d;
}

#define RUNNER(name, argt, fn)                                                                   \
    int name(const argt *a, int kind, uint64_t seed, orc_tables *tbl, int record, double *rows,  \
             int64_t max_rows, int64_t *n_rows, double *seconds)                                 \
    {                                                                                            \
        src_t s; collect_t c = { rows, max_rows, 0 };                                            \
        const orc_draws *d = make_src(&s, kind, seed, tbl, record);                              \
        struct timespec t0, t1; clock_gettime(CLOCK_MONOTONIC, &t0);                             \
        int rc = fn(a, d, collect_sink, &c);                                                     \
        clock_gettime(CLOCK_MONOTONIC, &t1);                                                     \
        if (seconds) *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec); \
        if (n_rows) *n_rows = c.n_rows;                                                          \
        return rc;                                                                               \
    }
RUNNER(orc_api_v2, orc_v2_args, orc_v2_run)
RUNNER(orc_api_groups, orc_groups_args, orc_groups_run)
RUNNER(orc_api_grstart, orc_grstart_args, orc_grstart_run)
RUNNER(orc_api_horseshoe, orc_hs_args, orc_horseshoe_run)

/* raw draw access for unit tests of the generator itself */
double orc_api_px_uniform(uint64_t seed, int stream, int64_t it, int64_t idx)
{ orc_philox p; orc_philox_init(&p, seed); orc_draws d = orc_philox_source(&p); return d.uniform(d.ctx, stream, it, idx); }
double orc_api_px_normal(uint64_t seed, int stream, int64_t it, int64_t idx)
{ orc_philox p; orc_philox_init(&p, seed); orc_draws d = orc_philox_source(&p); return d.normal(d.ctx, stream, it, idx); }
double orc_api_px_gamma(uint64_t seed, int stream, int64_t it, int64_t idx, double shape)
{ orc_philox p; orc_philox_init(&p, seed); orc_draws d = orc_philox_source(&p); return d.gamma(d.ctx, stream, it, idx, shape); }
void orc_api_px_shuffle(uint64_t seed, int stream, int64_t it, int32_t *order, int64_t n)
{ orc_philox p; orc_philox_init(&p, seed); orc_draws d = orc_philox_source(&p); d.shuffle(d.ctx, stream, it, order, n); }
