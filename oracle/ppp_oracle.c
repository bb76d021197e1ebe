/* ppp_oracle.c -- CPU restatement of PingPongPro's hot path in plain C.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under pingpongpro_b200/ may link, import or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg do,
 * and only as the checker.
 *
 * PARITY STATUS: the reference repository holds no tests, golden vectors or fixtures
 * (SURVEY.md §4), so this restatement is pinned against OUTPUTS OF THE REFERENCE ITSELF:
 * the unmodified /root/reference/pingpongpro.cpp compiled by oracle/Makefile into
 * oracle/_ref/ppp_harness, whose dumps are committed as tests/golden/ fixtures (made by
 * tests/golden/make_golden.py) and compared bit-for-bit in tests/test_oracle_golden.py.
 *
 * Every function cites the reference lines it follows.  Arithmetic types follow the
 * C++ usual arithmetic conversions of the reference under g++ -O2 -std=gnu++17, x86-64
 * SSE2 (FLT_EVAL_METHOD 0, no FMA contraction: build with -ffp-contract=off, no -march).
 * Sorted arrays replace the reference's std::map / std::list containers; iteration order
 * (strand, contig, position ascending) is the same.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint32_t key; uint32_t meta; } ppo_read; /* meta: bit0 minus strand, bit1 A10, bits 2.. NH */

typedef struct { uint32_t contig, strand, start, end; } ppo_interval;

typedef struct {
    uint32_t contig, strand, start, end;
    float p, q, reads_plus, reads_minus;
    float hist[64];
    double z64, p64; /* zValue (:1266) and pValue (:1267-1269) BEFORE the narrowing store of :1271 (NaN / 1 on the shortcut of :1250) */
} ppo_tp;

typedef struct { uint32_t strand, contig, key; float reads; uint32_t a10; } stack_t;

typedef struct {
    uint32_t ov, contig, pos, bin_pre, bin, local, bias;
    float rp, rm, fdr;
} sig_t;

typedef struct { uint32_t contig, key, meta; uint64_t idx; } rec_t;

typedef struct ppo {
    int n_contigs, min_ov, max_ov, pp_ov, mode, W;
    rec_t* recs; size_t n_recs, cap_recs;
    stack_t* stacks; size_t n_stacks;
    size_t strand_begin[3];
    double total;
    float* freq; uint32_t freq_len; /* dense: freq[h], 0 when absent */
    float max_height_score;
    float* grouped_pre;  /* [W][1000][2][2] */
    float* grouped_post; /* [W][C][2][2] */
    float* fdr_table;    /* [W][C][2][2] */
    uint32_t bin_map[1000];
    uint32_t C;
    sig_t* sigs; size_t n_sigs, cap_sigs;
} ppo;

#define BINS 1000u

ppo* ppo_new(int n_contigs, int min_ov, int max_ov, int pp_ov, int mode) {
    ppo* o = (ppo*)calloc(1, sizeof(ppo));
    o->n_contigs = n_contigs; o->min_ov = min_ov; o->max_ov = max_ov; o->pp_ov = pp_ov; o->mode = mode;
    o->W = max_ov - min_ov + 1;
    return o;
}

void ppo_free(ppo* o) {
    if (!o) return;
    free(o->recs); free(o->stacks); free(o->freq); free(o->grouped_pre); free(o->grouped_post); free(o->fdr_table); free(o->sigs);
    free(o);
}

/* Records must be added in file order (all calls together define the order). */
void ppo_add(ppo* o, int contig, const ppo_read* r, size_t n) {
    if (o->n_recs + n > o->cap_recs) {
        o->cap_recs = (o->n_recs + n) * 2 + 1024;
        o->recs = (rec_t*)realloc(o->recs, o->cap_recs * sizeof(rec_t));
    }
    for (size_t i = 0; i < n; ++i) {
        rec_t* d = &o->recs[o->n_recs];
        d->contig = (uint32_t)contig; d->key = r[i].key; d->meta = r[i].meta; d->idx = o->n_recs;
        o->n_recs++;
    }
}

/* readWeight, pingpongpro.cpp:430-456 */
static float read_weight(int mode, uint32_t nh) {
    if (mode == 2) return 1;                        /* unique  :432-436 */
    if (mode == 0) return (float)(1.0 / nh);        /* weighted:449 (double divide, narrowed) */
    return (nh == 1) ? 1 : 0;                       /* discard :454 */
}

static int cmp_rec(const void* a, const void* b) {
    const rec_t* x = (const rec_t*)a; const rec_t* y = (const rec_t*)b;
    uint32_t sx = x->meta & 1u, sy = y->meta & 1u;
    if (sx != sy) return sx < sy ? -1 : 1;
    if (x->contig != y->contig) return x->contig < y->contig ? -1 : 1;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0);
}

/* Stack accumulation, pingpongpro.cpp:458-488.  The std::map<strand, map<contig, map<pos>>> becomes an array
 * sorted by (strand, contig, key); within a stack the float adds run in file order as in :487. */
int ppo_build(ppo* o) {
    /* total in file order (:488), over records with w > 0 (:458) */
    o->total = 0;
    for (size_t i = 0; i < o->n_recs; ++i) {
        float w = read_weight(o->mode, o->recs[i].meta >> 2);
        if (w <= 0) continue;
        o->total += w;
    }
    qsort(o->recs, o->n_recs, sizeof(rec_t), cmp_rec);
    free(o->stacks);
    o->stacks = (stack_t*)malloc((o->n_recs + 1) * sizeof(stack_t));
    o->n_stacks = 0;
    for (size_t i = 0; i < o->n_recs; ++i) {
        const rec_t* r = &o->recs[i];
        float w = read_weight(o->mode, r->meta >> 2);
        if (w <= 0) continue;                                  /* :458 */
        uint32_t strand = r->meta & 1u;
        stack_t* s = o->n_stacks ? &o->stacks[o->n_stacks - 1] : NULL;
        if (!s || s->strand != strand || s->contig != r->contig || s->key != r->key) {
            s = &o->stacks[o->n_stacks++];
            s->strand = strand; s->contig = r->contig; s->key = r->key; s->reads = 0; s->a10 = 0;  /* :95-98 */
        }
        if (r->meta & 2u) s->a10 = 1;                          /* :471 / :483 */
        s->reads += w;                                         /* :487 */
    }
    o->strand_begin[0] = 0;
    size_t k = 0;
    while (k < o->n_stacks && o->stacks[k].strand == 0) ++k;
    o->strand_begin[1] = k; o->strand_begin[2] = o->n_stacks;
    return 0;
}

static inline uint32_t round_height(float reads) { return (uint32_t)(0.5 + reads); } /* map key conversion, :508 / :582 */

static void push_sig(ppo* o, sig_t s) {
    if (o->n_sigs == o->cap_sigs) {
        o->cap_sigs = o->cap_sigs * 2 + 4096;
        o->sigs = (sig_t*)realloc(o->sigs, o->cap_sigs * sizeof(sig_t));
    }
    o->sigs[o->n_sigs++] = s;
}

static int cmp_sig(const void* a, const void* b) {
    const sig_t* x = (const sig_t*)a; const sig_t* y = (const sig_t*)b;
    if (x->ov != y->ov) return x->ov < y->ov ? -1 : 1;
    if (x->contig != y->contig) return x->contig < y->contig ? -1 : 1;
    if (x->pos != y->pos) return x->pos < y->pos ? -1 : 1;
    return 0;
}

/* mapHeightsToScores (:502-509) + countStacksByGroup (:523-624). */
int ppo_scan(ppo* o) {
    const int W = o->W;
    /* --- height -> frequency, :505-508 (float counter: saturates at 2^24) */
    uint32_t max_h = 0;
    for (size_t i = 0; i < o->n_stacks; ++i) { uint32_t h = round_height(o->stacks[i].reads); if (h > max_h) max_h = h; }
    free(o->freq);
    o->freq_len = max_h + 1;
    o->freq = (float*)calloc(o->freq_len, sizeof(float));
    for (size_t i = 0; i < o->n_stacks; ++i) o->freq[round_height(o->stacks[i].reads)] += 1;

    free(o->grouped_pre);
    o->grouped_pre = (float*)calloc((size_t)W * BINS * 4, sizeof(float));
    o->n_sigs = 0;

    /* --- :547-551 */
    float max_hs = 0;
    for (uint32_t h = 0; h < o->freq_len; ++h) if (o->freq[h] > max_hs) max_hs = o->freq[h];
    max_hs = log10f(max_hs * max_hs);
    o->max_height_score = max_hs;

    /* --- :554-612; the minus strand of the same contig is searched by binary search */
    const stack_t* plus = o->stacks + o->strand_begin[0];
    size_t n_plus = o->strand_begin[1] - o->strand_begin[0];
    const stack_t* minus = o->stacks + o->strand_begin[1];
    size_t n_minus = o->strand_begin[2] - o->strand_begin[1];
    const stack_t** found = (const stack_t**)malloc(sizeof(stack_t*) * (size_t)W);
    size_t mlo = 0;  /* first minus stack of the current contig */
    for (size_t i = 0; i < n_plus; ++i) {
        const stack_t* p = &plus[i];
        while (mlo < n_minus && minus[mlo].contig < p->contig) ++mlo;
        if (mlo >= n_minus || minus[mlo].contig != p->contig) continue;   /* :556-557 */
        float mean = 0, mx = 0;
        for (int ov = o->min_ov; ov <= o->max_ov; ++ov) {                  /* :564-577 */
            uint64_t want = (uint64_t)p->key + (uint64_t)ov;
            const stack_t* hit = NULL;
            size_t lo = mlo, hi = n_minus;
            while (lo < hi) {
                size_t mid = lo + (hi - lo) / 2;
                if (minus[mid].contig != p->contig) { hi = mid; continue; }
                if (minus[mid].key < want) lo = mid + 1; else hi = mid;
            }
            if (lo < n_minus && minus[lo].contig == p->contig && minus[lo].key == want) hit = &minus[lo];
            found[ov - o->min_ov] = hit;
            if (hit) {
                mean += hit->reads;
                if (hit->reads > mx) mx = hit->reads;
            }
        }
        mean /= (float)(size_t)W;                                           /* :578 */
        if (mx > 0) {                                                       /* :580 */
            float hs_plus = o->freq[round_height(p->reads)];                 /* :582 */
            for (int ov = o->min_ov; ov <= o->max_ov; ++ov) {
                const stack_t* m = found[ov - o->min_ov];
                if (!m) continue;
                float hs = hs_plus * o->freq[round_height(m->reads)];       /* :589 */
                uint32_t bin = (uint32_t)(int)(0.5 + log10f(hs) / max_hs * (float)(size_t)(BINS - 1)); /* :591-594 */
                float local = (m->reads - (mean - m->reads / (float)(size_t)W)) / mx;          /* :597 */
                uint32_t lbin = (local < 0.2) ? 1u : 0u;                     /* :599 */
                uint32_t bbin = (p->a10 || m->a10) ? 0u : 1u;                /* :602 */
                if (bin < BINS)  /* the reference would write out of bounds otherwise */
                    o->grouped_pre[(((size_t)(ov - o->min_ov) * BINS + bin) * 2 + bbin) * 2 + lbin] += 1; /* :605 */
                sig_t s = { (uint32_t)(ov - o->min_ov), p->contig, p->key, bin, bin, lbin, bbin, p->reads, m->reads, 0.0f };
                push_sig(o, s);                                              /* :608 */
            }
        }
    }
    free(found);
    qsort(o->sigs, o->n_sigs, sizeof(sig_t), cmp_sig);  /* order of pingPongSignaturesByOverlap[ov][contig] lists */
    return 0;
}

/* collapseBins (:631-686) + calculateFDRs (:695-754). */
int ppo_score(ppo* o) {
    const int W = o->W;
    const float* g = o->grouped_pre;
    float* c = (float*)malloc((size_t)W * BINS * 4 * sizeof(float));
    memcpy(c, g, (size_t)W * BINS * 4 * sizeof(float));                      /* :634 */
    uint32_t cb = 0, bin = 0;
    while (bin < BINS) {                                                     /* :641 */
        for (int ov = 0; ov < W; ++ov) for (int k = 0; k < 4; ++k) c[((size_t)ov * BINS + cb) * 4 + k] = 0;  /* :644-647 */
        unsigned empty;
        do {
            empty = 0;
            for (int ov = 0; ov < W; ++ov)
                for (int k = 0; k < 4; ++k) {
                    float* cell = &c[((size_t)ov * BINS + cb) * 4 + k];
                    *cell += g[((size_t)ov * BINS + bin) * 4 + k];           /* :658 */
                    if (*cell <= 0) empty++;                                 /* :661 */
                }
            o->bin_map[bin] = cb;                                            /* :666 */
            bin++;
        } while (empty > 0 && bin < BINS);                                   /* :670 */
        cb++;
    }
    o->C = cb;
    free(o->grouped_post); free(o->fdr_table);
    o->grouped_post = (float*)malloc((size_t)W * cb * 4 * sizeof(float));
    o->fdr_table = (float*)malloc((size_t)W * cb * 4 * sizeof(float));
    for (int ov = 0; ov < W; ++ov) for (uint32_t b = 0; b < cb; ++b) for (int k = 0; k < 4; ++k)
        o->grouped_post[((size_t)ov * cb + b) * 4 + k] = c[((size_t)ov * BINS + b) * 4 + k];               /* :676-677 */
    free(c);
    for (size_t i = 0; i < o->n_sigs; ++i) if (o->sigs[i].bin_pre < BINS) o->sigs[i].bin = o->bin_map[o->sigs[i].bin_pre];  /* :679-682 */

    /* --- FDR per cell, :700-747 */
    const int pp = o->pp_ov - o->min_ov;
    const double pi = 3.14159265358979323846;                                /* :50 */
    for (uint32_t b = 0; b < cb; ++b) for (int k = 0; k < 4; ++k) {
        #define CELL(ov) (o->grouped_post[((size_t)(ov) * cb + b) * 4 + k])
        float mean = 0;
        for (int ov = 0; ov < W; ++ov) if (ov != pp) mean += CELL(ov);       /* :705-708 */
        mean = mean / (float)(size_t)(W - 1);                                /* :709 */
        float sd = 0;
        for (int ov = 0; ov < W; ++ov) if (ov != pp) { double d = (double)(CELL(ov) - mean); sd += d * d; }  /* :712-715, pow(float,2) is double */
        sd = sqrt(1.0 / (double)(size_t)(W - 1 - 1) * sd);                   /* :716 */
        if (sd <= 1E-10) sd = 1E-10;                                         /* :717-718 */
        for (int ov = 0; ov < W; ++ov) {
            double fdr = 0;
            float cnt = CELL(ov);
            for (double x = -5.0; x <= +5.0; x += 0.01) {                    /* :724 */
                if (x >= (cnt - mean) / sd)                                  /* :726 */
                    fdr += 0.01 * 1 / sqrt(2 * pi) * exp(-0.5 * x * x) * 1;  /* :728 */
                else
                    fdr += 0.01 * 1 / sqrt(2 * pi) * exp(-0.5 * x * x) * (x * sd + mean) / cnt;  /* :732 */
            }
            if (fdr < 0) fdr = 0; else if (fdr > 1) fdr = 1;                 /* :736-743 */
            o->fdr_table[((size_t)ov * cb + b) * 4 + k] = fdr;               /* :745 */
        }
        #undef CELL
    }
    for (size_t i = 0; i < o->n_sigs; ++i) {                                 /* :750-753 */
        sig_t* s = &o->sigs[i];
        s->fdr = o->fdr_table[(((size_t)s->ov * cb + s->bin) * 2 + s->bias) * 2 + s->local];
    }
    return 0;
}

typedef struct { ppo_tp t; size_t order; } tp_sort_t;

static int cmp_tp_pos(const void* a, const void* b) {
    const tp_sort_t* x = (const tp_sort_t*)a; const tp_sort_t* y = (const tp_sort_t*)b;
    if (x->t.contig != y->t.contig) return x->t.contig < y->t.contig ? -1 : 1;
    if (x->t.start != y->t.start) return x->t.start < y->t.start ? -1 : 1;     /* :206-212 */
    if (x->t.end != y->t.end) return x->t.end < y->t.end ? -1 : 1;
    return x->order < y->order ? -1 : (x->order > y->order ? 1 : 0);            /* list::sort is stable */
}
static int cmp_tp_p(const void* a, const void* b) {
    const tp_sort_t* x = *(const tp_sort_t* const*)a; const tp_sort_t* y = *(const tp_sort_t* const*)b;
    if (x->t.p < y->t.p) return -1;                                            /* :1184 */
    if (y->t.p < x->t.p) return 1;
    return x->order < y->order ? -1 : (x->order > y->order ? 1 : 0);
}

/* findSuppressedTransposons (:1195-1313).  `out` receives the transposons in the reference's output order
 * (contig, start, end).  The std::list iterator walk of :1215-1237 dereferences end(); with libstdc++ the
 * position read there is the list's size.  That behaviour is emulated literally
 * (sentinel position = number of signatures in the list) so that overlapping intervals give what the
 * reference binary gives. */
int ppo_transposons(ppo* o, const ppo_interval* iv, size_t n, ppo_tp* out) {
    const int W = o->W;
    const int pp = o->pp_ov - o->min_ov;
    const double pi = 3.14159265358979323846;
    tp_sort_t* t = (tp_sort_t*)calloc(n + 1, sizeof(tp_sort_t));
    for (size_t i = 0; i < n; ++i) {
        t[i].t.contig = iv[i].contig; t[i].t.strand = iv[i].strand; t[i].t.start = iv[i].start; t[i].t.end = iv[i].end;
        t[i].t.p = 1; t[i].t.q = 1; t[i].order = i;                           /* :187 */
        t[i].t.z64 = NAN; t[i].t.p64 = 1.0;
    }
    qsort(t, n, sizeof(tp_sort_t), cmp_tp_pos);                               /* :1174 + map order */
    for (size_t i = 0; i < n; ++i) t[i].order = i;
    /* per (overlap, contig) signature ranges in o->sigs (sorted by ov, contig, pos) */
    size_t a = 0;
    while (a < n) {
        size_t bnd = a;
        while (bnd < n && t[bnd].t.contig == t[a].t.contig) ++bnd;
        uint32_t contig = t[a].t.contig;
        for (int ov = 0; ov < W; ++ov) {
            size_t lo = 0, hi = o->n_sigs;  /* find [s0, s1) of (ov, contig) */
            sig_t key = { (uint32_t)ov, contig, 0, 0, 0, 0, 0, 0, 0, 0 };
            while (lo < hi) { size_t mid = (lo + hi) / 2; if (cmp_sig(&o->sigs[mid], &key) < 0) lo = mid + 1; else hi = mid; }
            size_t s0 = lo;
            size_t s1 = s0;
            while (s1 < o->n_sigs && o->sigs[s1].ov == (uint32_t)ov && o->sigs[s1].contig == contig) ++s1;
            const sig_t* L = o->sigs + s0;
            size_t len = s1 - s0;
            #define POS(i) ((i) < len ? L[(i)].pos : (uint32_t)len)   /* end() sentinel reads the list size */
            size_t it = 0;                                                     /* :1204 / :1206 */
            for (size_t k = a; k < bnd; ++k) {
                ppo_tp* tp = &t[k].t;
                while (it != 0 && POS(it) > tp->start) --it;                   /* :1215-1216 */
                while (it != len && POS(it) < tp->start) ++it;                 /* :1217-1218 */
                float sum = 0;
                if (POS(it) >= tp->start) {                                    /* :1222 */
                    while (POS(it) <= tp->end && it != len) {                  /* :1224 */
                        sum += (L[it].rp + L[it].rm) * (1 - L[it].fdr);        /* :1227 */
                        if (ov == pp) { tp->reads_plus += L[it].rp; tp->reads_minus += L[it].rm; }  /* :1230-1234 */
                        ++it;
                    }
                }
                tp->hist[ov] = sum;                                            /* :1240 */
            }
            #undef POS
        }
        a = bnd;
    }
    for (size_t k = 0; k < n; ++k) {
        ppo_tp* tp = &t[k].t;
        float mean = 0;
        for (int ov = 0; ov < W; ++ov) if (ov != pp) mean += tp->hist[ov];     /* :1244-1247 */
        mean = mean / (float)(size_t)(W - 1);                                  /* :1248 */
        if (mean == 0 && tp->hist[pp] == 0) { tp->p = 1; continue; }           /* :1250-1253 */
        float sd = 0;
        for (int ov = 0; ov < W; ++ov) if (ov != pp) { double d = (double)(tp->hist[ov] - mean); sd += d * d; }  /* :1257-1260 */
        sd = sqrt(1.0 / (double)(size_t)(W - 1 - 1) * sd);                     /* :1261 */
        if (sd <= 1E-10) sd = 1E-10;                                           /* :1262-1263 */
        double z = (tp->hist[pp] - mean) / sd;                                 /* :1266 */
        double p = 0;
        if (z < 1e13) {  /* beyond that x += 0.01 no longer advances and the reference never returns */
            for (double x = z; x <= z + 5.0; x += 0.01)                        /* :1268 */
                p += 0.01 * 1 / sqrt(2 * pi) * exp(-0.5 * x * x);              /* :1269 */
        }
        tp->z64 = z; tp->p64 = p;
        tp->p = p;                                                             /* :1271 */
    }
    /* q-values, :1276-1312 */
    tp_sort_t** byp = (tp_sort_t**)malloc((n + 1) * sizeof(tp_sort_t*));
    for (size_t k = 0; k < n; ++k) byp[k] = &t[k];
    qsort(byp, n, sizeof(tp_sort_t*), cmp_tp_p);                               /* stable by construction (:1281) */
    unsigned i = 1;
    float prev_p = 1, prev_q = 1;
    for (size_t r = n; r-- > 0;) {                                             /* reverse iteration :1286-1287 */
        ppo_tp* tp = &byp[r]->t;
        if (tp->hist[pp] / (tp->reads_plus + tp->reads_minus) < 0.2) tp->q = 1; /* :1289-1291 */
        else if (tp->p == prev_p) tp->q = prev_q;                              /* :1293-1298 */
        else {
            tp->q = tp->p * (float)n / (float)i;                               /* :1302 (float * size_t / unsigned) */
            if (tp->q > 1) tp->q = 1;
            prev_p = tp->p; prev_q = tp->q;
        }
        i++;
    }
    free(byp);
    for (size_t k = 0; k < n; ++k) out[k] = t[k].t;
    free(t);
    return 0;
}

/* ------------------------------------------------------------------ accessors */
size_t ppo_n_stacks(const ppo* o) { return o->n_stacks; }
void ppo_get_stacks(const ppo* o, uint32_t* strand, uint32_t* contig, uint32_t* key, float* reads, uint32_t* a10) {
    for (size_t i = 0; i < o->n_stacks; ++i) {
        strand[i] = o->stacks[i].strand; contig[i] = o->stacks[i].contig; key[i] = o->stacks[i].key;
        reads[i] = o->stacks[i].reads; a10[i] = o->stacks[i].a10;
    }
}
double ppo_total(const ppo* o) { return o->total; }
uint32_t ppo_freq_len(const ppo* o) { return o->freq_len; }
void ppo_get_freq(const ppo* o, float* out) { memcpy(out, o->freq, o->freq_len * sizeof(float)); }
float ppo_max_height_score(const ppo* o) { return o->max_height_score; }
void ppo_get_grouped_pre(const ppo* o, float* out) { memcpy(out, o->grouped_pre, (size_t)o->W * BINS * 4 * sizeof(float)); }
uint32_t ppo_n_bins(const ppo* o) { return o->C; }
void ppo_get_grouped_post(const ppo* o, float* out) { memcpy(out, o->grouped_post, (size_t)o->W * o->C * 4 * sizeof(float)); }
void ppo_get_fdr_table(const ppo* o, float* out) { memcpy(out, o->fdr_table, (size_t)o->W * o->C * 4 * sizeof(float)); }
void ppo_get_bin_map(const ppo* o, uint32_t* out) { memcpy(out, o->bin_map, sizeof o->bin_map); }
size_t ppo_n_sigs(const ppo* o) { return o->n_sigs; }
/* columns: ov, contig, pos, bin_pre, bin, local, bias as uint32[n_sigs][7]; rp, rm, fdr as float[n_sigs][3] */
void ppo_get_sigs(const ppo* o, uint32_t* ints, float* flts) {
    for (size_t i = 0; i < o->n_sigs; ++i) {
        const sig_t* s = &o->sigs[i];
        uint32_t* a = ints + 7 * i;
        a[0] = s->ov; a[1] = s->contig; a[2] = s->pos; a[3] = s->bin_pre; a[4] = s->bin; a[5] = s->local; a[6] = s->bias;
        flts[3 * i] = s->rp; flts[3 * i + 1] = s->rm; flts[3 * i + 2] = s->fdr;
    }
}
