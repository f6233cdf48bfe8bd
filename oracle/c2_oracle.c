/*
 * c2_oracle.c -- CPU restatement of CRISPResso2's per-read hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in crispresso2_b200/ may import, link or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against
 *   (1) the reference's own known-answer tests
 *       (/root/reference/tests/unit_tests/test_CRISPResso2Align.py:35-332,
 *        test_CRISPRessoCOREResources.py:11-238), re-stated in tests/,
 *   (2) golden vectors produced by the compiled reference itself
 *       (tests/golden/ *.json, generator tests/golden/gen_golden.py), and
 *   (3) when oracle/_ref is present, live fuzzing against the compiled
 *       reference Cython modules.
 *
 * What is restated (reference file:line):
 *   c2o_global_align   <- CRISPResso2/CRISPResso2Align.pyx:101-434
 *   c2o_find_indels    <- CRISPResso2/CRISPRessoCOREResources.pyx:68-187
 *
 * The code below is written from the algorithm description in SURVEY.md
 * section 3.2/3.3; it is not a transcription of the Cython source.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { ST_M = 1, ST_I = 2, ST_J = 3, ST_UNSET = 0 };

typedef struct {
    int32_t *m, *x, *y;      /* scores: match, gap-in-ref ("I"), gap-in-read ("J") */
    uint8_t *pm, *px, *py;   /* back pointers (ST_*) ; ST_UNSET = never written by the reference */
    size_t stride;
} planes_t;

#define AT(p, i, j) ((p)[(size_t)(i) * pl.stride + (size_t)(j)])

/* tie rule of Align.pyx:216-229 -- I beats J beats M on equal scores */
static inline void pick3(int32_t mv, int32_t iv, int32_t jv, int32_t *best, uint8_t *from)
{
    if (mv > jv) {
        if (mv > iv) { *best = mv; *from = ST_M; } else { *best = iv; *from = ST_I; }
    } else {
        if (jv > iv) { *best = jv; *from = ST_J; } else { *best = iv; *from = ST_I; }
    }
}

/*
 * Returns 0 on success.
 *  -1 : gap_incentive length mismatch is the caller's business (checked in Python, Align.pyx:124-126)
 *  -2 : empty sequence or symbol outside the matrix (reference behaviour undefined)
 *  -3 : traceback walked onto a pointer the reference never initialises (undefined in the reference)
 *  -4 : out of memory
 * out_j / out_i must hold I+J bytes.  Strings are returned left-to-right.
 */
int c2o_global_align(const uint8_t *seqj, int64_t J, const uint8_t *seqi, int64_t I,
                     const int64_t *matrix, int64_t msize,
                     const int64_t *gi, int32_t gap_open, int32_t gap_extend,
                     uint8_t *out_j, uint8_t *out_i, int32_t *aln_len, int32_t *n_match)
{
    if (I < 1 || J < 1) return -2;
    for (int64_t k = 0; k < I; k++) if (seqi[k] >= msize) return -2;
    for (int64_t k = 0; k < J; k++) if (seqj[k] >= msize) return -2;

    planes_t pl;
    size_t cells = (size_t)(I + 1) * (size_t)(J + 1);
    pl.stride = (size_t)(J + 1);
    pl.m = malloc(cells * sizeof(int32_t)); pl.x = malloc(cells * sizeof(int32_t)); pl.y = malloc(cells * sizeof(int32_t));
    pl.pm = calloc(cells, 1); pl.px = calloc(cells, 1); pl.py = calloc(cells, 1);
    if (!pl.m || !pl.x || !pl.y || !pl.pm || !pl.px || !pl.py) {
        free(pl.m); free(pl.x); free(pl.y); free(pl.pm); free(pl.px); free(pl.py);
        return -4;
    }

    /* sentinel of Align.pyx:150: product formed in 64 bits, stored in a C int */
    const int32_t NEG = (int32_t)((int64_t)gap_open * J * I);

    /* borders, Align.pyx:153-176 */
    for (int64_t j = 1; j <= J; j++) {
        AT(pl.m, 0, j) = NEG;  AT(pl.pm, 0, j) = ST_I;
        AT(pl.x, 0, j) = (int32_t)((int64_t)gap_extend * j + gi[0]);  AT(pl.px, 0, j) = ST_I;
        AT(pl.y, 0, j) = NEG;                                        /* py[0][j] stays unset */
    }
    for (int64_t i = 1; i <= I; i++) {
        AT(pl.m, i, 0) = NEG;  AT(pl.pm, i, 0) = ST_J;
        AT(pl.y, i, 0) = (int32_t)((int64_t)gap_extend * i + gi[0]);  AT(pl.py, i, 0) = ST_J;
        AT(pl.x, i, 0) = NEG;                                        /* px[i][0] stays unset */
    }
    AT(pl.m, 0, 0) = 0; AT(pl.x, 0, 0) = NEG; AT(pl.y, 0, 0) = NEG;

    /* fill; one loop with the "free opening on the last row / last column" rule
       (Align.pyx:187-228 interior, :234-273 last column, :277-317 last row) */
    for (int64_t i = 1; i <= I; i++) {
        const int64_t *mrow = matrix + (int64_t)seqi[i - 1] * msize;
        for (int64_t j = 1; j <= J; j++) {
            const int64_t open = (i == I || j == J) ? gap_extend : gap_open;
            int32_t a, b;
            /* gap in reference (read base consumed): incentive of row i on open AND extend */
            a = (int32_t)(open + (int64_t)AT(pl.m, i, j - 1) + gi[i]);
            b = (int32_t)((int64_t)gap_extend + AT(pl.x, i, j - 1) + gi[i]);
            if (a > b) { AT(pl.x, i, j) = a; AT(pl.px, i, j) = ST_M; } else { AT(pl.x, i, j) = b; AT(pl.px, i, j) = ST_I; }
            /* gap in read (reference base consumed): incentive of row i-1, on open only */
            a = (int32_t)(open + (int64_t)AT(pl.m, i - 1, j) + gi[i - 1]);
            b = (int32_t)((int64_t)gap_extend + AT(pl.y, i - 1, j));
            if (a > b) { AT(pl.y, i, j) = a; AT(pl.py, i, j) = ST_M; } else { AT(pl.y, i, j) = b; AT(pl.py, i, j) = ST_J; }
            /* diagonal */
            const int64_t s = mrow[seqj[j - 1]];
            int32_t mv = (int32_t)(AT(pl.m, i - 1, j - 1) + s);
            int32_t iv = (int32_t)(AT(pl.x, i - 1, j - 1) + s);
            int32_t jv = (int32_t)(AT(pl.y, i - 1, j - 1) + s);
            pick3(mv, iv, jv, &AT(pl.m, i, j), &AT(pl.pm, i, j));
        }
    }

    /* traceback, Align.pyx:338-421 (emitted right-to-left, reversed at the end) */
    int64_t i = I, j = J, n = 0;
    int32_t matches = 0, rc = 0;
    uint8_t state, dummy_from; int32_t dummy_best;
    pick3(AT(pl.m, I, J), AT(pl.x, I, J), AT(pl.y, I, J), &dummy_best, &dummy_from);
    state = dummy_from;
    while (i > 0 || j > 0) {
        const uint8_t ci = seqi[i > 0 ? i - 1 : 0];
        const uint8_t cj = seqj[j > 0 ? j - 1 : 0];
        if (state == ST_M) {
            if (i == 0 || j == 0) { rc = -3; break; }   /* reference would re-emit seq[0]; undefined zone */
            state = AT(pl.pm, i, j);
            out_j[n] = cj; out_i[n] = ci;
            if (ci == cj) matches++;
            i--; j--;
        } else if (state == ST_J) {
            if (i == 0) { rc = -3; break; }
            state = AT(pl.py, i, j);
            out_j[n] = '-'; out_i[n] = ci;
            i--;
        } else if (state == ST_I) {
            if (j == 0) { rc = -3; break; }
            state = AT(pl.px, i, j);
            out_j[n] = cj; out_i[n] = '-';
            j--;
        } else { rc = -3; break; }
        n++;
    }
    if (rc == 0) {
        for (int64_t a = 0, b = n - 1; a < b; a++, b--) {
            uint8_t t = out_j[a]; out_j[a] = out_j[b]; out_j[b] = t;
            t = out_i[a]; out_i[a] = out_i[b]; out_i[b] = t;
        }
        *aln_len = (int32_t)n; *n_match = matches;
    }
    free(pl.m); free(pl.x); free(pl.y); free(pl.pm); free(pl.px); free(pl.py);
    return rc;
}

/* ------------------------------------------------------------------------- */

typedef struct {
    /* every list has capacity 2*n (n = alignment columns) except coordinates (pairs -> 2*n ints) */
    int32_t *ref_positions;            int32_t n_ref_positions;
    int32_t *all_sub_pos;              int32_t n_all_sub;       uint8_t *all_sub_val;
    int32_t *sub_pos;                  int32_t n_sub;           uint8_t *sub_val;
    int32_t *all_del_pos;              int32_t n_all_del_pos;
    int32_t *all_del_coord;            int32_t n_all_del_coord; /* pairs */
    int32_t *del_pos;                  int32_t n_del_pos;
    int32_t *del_coord;                int32_t n_del_coord;     /* pairs */
    int32_t *del_sizes;                int32_t n_del_sizes;
    int32_t *all_ins_pos;              int32_t n_all_ins_pos;
    int32_t *all_ins_left;             int32_t n_all_ins_left;
    int32_t *ins_pos;                  int32_t n_ins_pos;
    int32_t *ins_coord;                int32_t n_ins_coord;     /* pairs */
    int32_t *ins_sizes;                int32_t n_ins_sizes;
    int64_t substitution_n, deletion_n, insertion_n;
} c2o_edits_t;

static inline int in_set(const uint8_t *mask, int64_t mask_len, int64_t v)
{
    return v >= 0 && v < mask_len && mask[v];
}

static int range_hits(const uint8_t *mask, int64_t mask_len, int64_t lo, int64_t hi)
{
    if (lo < 0) lo = 0;
    if (hi > mask_len) hi = mask_len;
    for (int64_t p = lo; p < hi; p++) if (mask[p]) return 1;
    return 0;
}

/*
 * One left-to-right pass over the aligned pair (COREResources.pyx:108-163).
 * include_mask[p] != 0  <=>  p is in the quantification window (mask_len entries).
 * Lists whose capacity could be exceeded by a pathological hand-made pair
 * (deleted ranges) are guarded with cap.
 */
int c2o_find_indels(const uint8_t *read_al, const uint8_t *ref_al, int64_t n,
                    const uint8_t *include_mask, int64_t mask_len,
                    c2o_edits_t *o, int64_t cap)
{
    int64_t idx = 0, start_del = -1, start_ins = -1, cur_ins = 0;
    o->n_ref_positions = o->n_all_sub = o->n_sub = 0;
    o->n_all_del_pos = o->n_all_del_coord = o->n_del_pos = o->n_del_coord = o->n_del_sizes = 0;
    o->n_all_ins_pos = o->n_all_ins_left = o->n_ins_pos = o->n_ins_coord = o->n_ins_sizes = 0;
    int64_t del_total = 0, ins_total = 0;

    for (int64_t c = 0; c < n; c++) {
        const uint8_t r = ref_al[c], q = read_al[c];
        if (r != '-') {
            o->ref_positions[o->n_ref_positions++] = (int32_t)idx;
            if (r != q && q != '-' && q != 'N') {
                o->all_sub_pos[o->n_all_sub] = (int32_t)idx; o->all_sub_val[o->n_all_sub++] = q;
                if (in_set(include_mask, mask_len, idx)) { o->sub_pos[o->n_sub] = (int32_t)idx; o->sub_val[o->n_sub++] = q; }
            }
            if (start_ins != -1) {             /* an insertion run ends here */
                o->all_ins_left[o->n_all_ins_left++] = (int32_t)start_ins;
                o->all_ins_pos[o->n_all_ins_pos++] = (int32_t)start_ins;
                o->all_ins_pos[o->n_all_ins_pos++] = (int32_t)idx;
                if (in_set(include_mask, mask_len, start_ins) && in_set(include_mask, mask_len, idx)) {
                    o->ins_coord[2 * o->n_ins_coord] = (int32_t)start_ins; o->ins_coord[2 * o->n_ins_coord + 1] = (int32_t)idx; o->n_ins_coord++;
                    o->ins_pos[o->n_ins_pos++] = (int32_t)start_ins;
                    o->ins_pos[o->n_ins_pos++] = (int32_t)idx;
                    o->ins_sizes[o->n_ins_sizes++] = (int32_t)cur_ins; ins_total += cur_ins;
                }
                start_ins = -1;
            }
            cur_ins = 0;
            idx++;
        } else {
            o->ref_positions[o->n_ref_positions++] = (int32_t)(idx == 0 ? -1 : -idx);
            if (idx > 0 && start_ins == -1) start_ins = idx - 1;
            cur_ins++;
        }

        if (q == '-' && start_del == -1) {
            start_del = (c >= 1) ? o->ref_positions[c] : 0;
        } else if (q != '-' && start_del != -1) {
            const int64_t end_del = o->ref_positions[c];
            const int hit = range_hits(include_mask, mask_len, start_del, end_del);
            for (int64_t p = start_del; p < end_del; p++) {
                if (o->n_all_del_pos >= cap) return -5;
                o->all_del_pos[o->n_all_del_pos++] = (int32_t)p;
                if (hit) o->del_pos[o->n_del_pos++] = (int32_t)p;
            }
            o->all_del_coord[2 * o->n_all_del_coord] = (int32_t)start_del; o->all_del_coord[2 * o->n_all_del_coord + 1] = (int32_t)end_del; o->n_all_del_coord++;
            if (hit) {
                o->del_coord[2 * o->n_del_coord] = (int32_t)start_del; o->del_coord[2 * o->n_del_coord + 1] = (int32_t)end_del; o->n_del_coord++;
                o->del_sizes[o->n_del_sizes++] = (int32_t)(end_del - start_del); del_total += end_del - start_del;
            }
            start_del = -1;
        }
    }
    if (start_del != -1) {                     /* deletion running off the right end */
        const int64_t end_del = (int64_t)o->ref_positions[n - 1] + 1;
        const int hit = range_hits(include_mask, mask_len, start_del, end_del);
        for (int64_t p = start_del; p < end_del; p++) {
            if (o->n_all_del_pos >= cap) return -5;
            o->all_del_pos[o->n_all_del_pos++] = (int32_t)p;
            if (hit) o->del_pos[o->n_del_pos++] = (int32_t)p;
        }
        o->all_del_coord[2 * o->n_all_del_coord] = (int32_t)start_del; o->all_del_coord[2 * o->n_all_del_coord + 1] = (int32_t)end_del; o->n_all_del_coord++;
        if (hit) {
            o->del_coord[2 * o->n_del_coord] = (int32_t)start_del; o->del_coord[2 * o->n_del_coord + 1] = (int32_t)end_del; o->n_del_coord++;
            o->del_sizes[o->n_del_sizes++] = (int32_t)(end_del - start_del); del_total += end_del - start_del;
        }
    }
    o->substitution_n = o->n_sub;
    o->deletion_n = del_total;
    o->insertion_n = ins_total;
    return 0;
}

/* ------------------------------------------------------------------------- */
/*
 * Throughput helper for bench.py's cpu_baseline "port" leg: align + classify a
 * batch of reads against one amplicon, single thread, no Python in the loop.
 * Returns the number of reads processed; sums a checksum so the work cannot be
 * optimised away.
 */
int64_t c2o_batch_align_classify(const uint8_t *reads, const int64_t *offsets, int64_t n_reads,
                                 const uint8_t *ref, int64_t I,
                                 const int64_t *matrix, int64_t msize, const int64_t *gi,
                                 int32_t gap_open, int32_t gap_extend,
                                 const uint8_t *include_mask, int64_t *checksum)
{
    int64_t maxJ = 0;
    for (int64_t r = 0; r < n_reads; r++) { int64_t L = offsets[r + 1] - offsets[r]; if (L > maxJ) maxJ = L; }
    const int64_t cap = I + maxJ + 8;
    uint8_t *aj = malloc(cap), *ai = malloc(cap);
    c2o_edits_t e; memset(&e, 0, sizeof e);
    int32_t *pool = malloc(sizeof(int32_t) * cap * 2 * 13);
    uint8_t *vals = malloc(cap * 4);
    int32_t *p = pool;
    e.ref_positions = p; p += 2 * cap; e.all_sub_pos = p; p += 2 * cap; e.sub_pos = p; p += 2 * cap;
    e.all_del_pos = p; p += 2 * cap; e.all_del_coord = p; p += 2 * cap; e.del_pos = p; p += 2 * cap;
    e.del_coord = p; p += 2 * cap; e.del_sizes = p; p += 2 * cap; e.all_ins_pos = p; p += 2 * cap;
    e.all_ins_left = p; p += 2 * cap; e.ins_pos = p; p += 2 * cap; e.ins_coord = p; p += 2 * cap; e.ins_sizes = p;
    e.all_sub_val = vals; e.sub_val = vals + 2 * cap;
    int64_t done = 0, sum = 0;
    for (int64_t r = 0; r < n_reads; r++) {
        int32_t n = 0, m = 0;
        int rc = c2o_global_align(reads + offsets[r], offsets[r + 1] - offsets[r], ref, I, matrix, msize, gi,
                                  gap_open, gap_extend, aj, ai, &n, &m);
        if (rc == 0) {
            c2o_find_indels(aj, ai, n, include_mask, I, &e, 2 * cap);
            sum += m + 3 * n + e.n_all_sub + 7 * e.n_all_del_pos + 11 * e.n_all_ins_pos;
        }
        done++;
    }
    *checksum = sum;
    free(aj); free(ai); free(pool); free(vals);
    return done;
}
