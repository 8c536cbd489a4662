/* oracle/famseq_oracle.c -- TEST INFRASTRUCTURE ONLY (see famseq_oracle.h).
 *
 * A from-scratch C restatement of the reference engine's arithmetic, written so that every
 * multiply / add / divide happens in the same order as in /root/reference/src/family.cpp
 * (compiled there for plain x86-64, i.e. without FMA contraction; this file is built with
 * -ffp-contract=off).  With the libc rand() stream it reproduces the reference bit for bit,
 * which is what tests/test_oracle_vs_ref.py asserts.  It is deliberately simple and slow.
 */
#include "famseq_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* Mendelian transmission tables                                                              */
/* ------------------------------------------------------------------------------------------ */

#define T3(t, g, a, b) ((t)[(g) * 9 + (a) * 3 + (b)])

/* Autosomal table, family.cpp:447-550 (calPCP2S with two alleles).  Genotype codes:
 * 0 = RR (alleles 0,0), 1 = RA (0,1), 2 = AA (1,1).  The table is built as the sum over the four
 * (maternal haplotype, paternal haplotype) choices of an outer product of gamete distributions,
 * accumulated in the reference's order -- which is why it is not bit-symmetric. */
static void table_autosome(double mu, double *t) {
    static const int allele[3][2] = {{0, 0}, {0, 1}, {1, 1}};
    static const int code[2][2] = {{0, 1}, {1, 2}};
    memset(t, 0, 27 * sizeof(double));
    for (int a = 0; a < 3; a++) {
        for (int b = 0; b < 3; b++) {
            if (mu == 0) { /* family.cpp:472-490: exact quarters */
                for (int ha = 0; ha < 2; ha++)
                    for (int hb = 0; hb < 2; hb++) {
                        int g = code[allele[a][ha]][allele[b][hb]];
                        T3(t, g, a, b) = T3(t, g, a, b) + 0.25;
                    }
                continue;
            }
            /* family.cpp:497-544.  gamete[k] = Pr(parent transmits allele k | chosen haplotype) / 2 */
            const double flip = mu / (2 * (2 - 1));
            const double keep = (1 - mu) / 2;
            double pm[2] = {flip, flip}, pf[2] = {flip, flip};
            /* the reference mutates the two gamete vectors in place between the four passes */
            for (int pass = 0; pass < 4; pass++) {
                switch (pass) {
                case 0:
                    pm[allele[a][0]] = keep;
                    pf[allele[b][0]] = keep;
                    break;
                case 1:
                    pf[allele[b][0]] = flip;
                    pf[allele[b][1]] = keep;
                    break;
                case 2:
                    pm[allele[a][0]] = flip;
                    pf[allele[b][1]] = flip;
                    pm[allele[a][1]] = keep;
                    pf[allele[b][0]] = keep;
                    break;
                default:
                    pf[allele[b][0]] = flip;
                    pf[allele[b][1]] = keep;
                    break;
                }
                for (int k = 0; k < 2; k++)
                    for (int l = 0; l < 2; l++) {
                        int g = code[k][l];
                        T3(t, g, a, b) = T3(t, g, a, b) + pm[k] * pf[l];
                    }
            }
        }
    }
}

/* X chromosome, daughter: family.cpp:383-416.  Father column b is his single X (0 or 2);
 * the het-father column stays zero. */
static void table_x_daughter(double mu, double *t) {
    const double q = 1.0 - mu;
    memset(t, 0, 27 * sizeof(double));
    T3(t, 0, 0, 0) = q * q;
    T3(t, 1, 0, 0) = 2 * mu * q;
    T3(t, 2, 0, 0) = mu * mu;

    T3(t, 0, 0, 2) = q * mu;
    T3(t, 1, 0, 2) = q * q + mu * mu;
    T3(t, 2, 0, 2) = q * mu;

    T3(t, 0, 1, 0) = q * q / 2 + mu * q / 2;
    T3(t, 1, 1, 0) = mu * q + q * q / 2 + mu * mu / 2;
    T3(t, 2, 1, 0) = mu * mu / 2 + mu * q / 2;

    T3(t, 0, 1, 2) = mu * mu / 2 + mu * q / 2;
    T3(t, 1, 1, 2) = mu * q + q * q / 2 + mu * mu / 2;
    T3(t, 2, 1, 2) = q * q / 2 + mu * q / 2;

    T3(t, 0, 2, 0) = q * mu;
    T3(t, 1, 2, 0) = q * q + mu * mu;
    T3(t, 2, 2, 0) = q * mu;

    T3(t, 0, 2, 2) = mu * mu;
    T3(t, 1, 2, 2) = 2 * mu * q;
    T3(t, 2, 2, 2) = q * q;
}

/* X chromosome, son: family.cpp:418-445.  Depends on the mother only; a het son is impossible. */
static void table_x_son(double mu, double *t) {
    memset(t, 0, 27 * sizeof(double));
    for (int b = 0; b < 3; b += 2) {
        T3(t, 0, 0, b) = 1 - mu;
        T3(t, 2, 0, b) = mu;
        T3(t, 0, 1, b) = 0.5;
        T3(t, 2, 1, b) = 0.5;
        T3(t, 0, 2, b) = mu;
        T3(t, 2, 2, b) = 1 - mu;
    }
}

void fso_tables(double mrate, double *pcp2, double *pcp2xf, double *pcp2xm) {
    table_autosome(mrate, pcp2);
    table_x_daughter(mrate, pcp2xf);
    table_x_son(mrate, pcp2xm);
}

/* ------------------------------------------------------------------------------------------ */
/* Family state                                                                               */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    int N, S;
    int *mo, *fa;        /* parent rows, -1 for founders          (family.cpp:291-350) */
    int *gender;         /* 1 male, 2 female                                            */
    int *nchild, *child; /* child[i*N + k], ped order                                  */
    int *nsp, *sp;       /* spouse[i*N + k], order of first joint child                */
    const int *cols;     /* sequenced columns -> ped row                               */
    double tA[27], tXf[27], tXm[27];
    double prior[4][3];
    double lc;
    double *lk;     /* N x 3, current variant */
    double *single; /* N x 3 */
    double *post;   /* N x 3 */
    /* ES scratch */
    double *ant;  /* N x 3, -1 = not yet                                  */
    double *pos;  /* [i][j][g]  -> pos[(i*N + j)*3 + g], -1 = not yet     */
    char *ant_busy, *pos_busy;
    int loop;
} fam_t;

static int build_topology(fam_t *f, const int *id, const int *mid, const int *fid) {
    const int N = f->N;
    for (int i = 0; i < N; i++) {
        f->nchild[i] = 0;
        f->nsp[i] = 0;
    }
    for (int i = 0; i < N; i++) {
        int m = -1, p = -1;
        for (int j = 0; j < N; j++) { /* no break: the LAST row with a matching id wins */
            if (mid[i] == id[j]) m = j;
            if (fid[i] == id[j]) p = j;
        }
        if ((m < 0) != (p < 0)) return FSO_E_HALF_PARENTS;
        f->mo[i] = m;
        f->fa[i] = p;
        if (m < 0) continue;
        f->child[m * N + f->nchild[m]++] = i;
        f->child[p * N + f->nchild[p]++] = i;
        int seen = 0;
        for (int k = 0; k < f->nsp[m]; k++)
            if (f->sp[m * N + k] == p) {
                seen = 1;
                break;
            }
        if (!seen) {
            f->sp[m * N + f->nsp[m]++] = p;
            f->sp[p * N + f->nsp[p]++] = m;
        }
    }
    for (int i = 0; i < N; i++) /* checkPed, family.cpp:204-219 */
        if (f->mo[i] >= 0 && (f->gender[f->mo[i]] != 2 || f->gender[f->fa[i]] != 1)) return FSO_E_GENDER;
    return FSO_OK;
}

static fam_t *fam_new(int N, int S) {
    fam_t *f = (fam_t *)calloc(1, sizeof(fam_t));
    f->N = N;
    f->S = S;
    f->mo = (int *)calloc(N, sizeof(int));
    f->fa = (int *)calloc(N, sizeof(int));
    f->gender = (int *)calloc(N, sizeof(int));
    f->nchild = (int *)calloc(N, sizeof(int));
    f->child = (int *)calloc((size_t)N * N + 1, sizeof(int));
    f->nsp = (int *)calloc(N, sizeof(int));
    f->sp = (int *)calloc((size_t)N * N + 1, sizeof(int));
    f->lk = (double *)calloc((size_t)N * 3, sizeof(double));
    f->single = (double *)calloc((size_t)N * 3, sizeof(double));
    f->post = (double *)calloc((size_t)N * 3, sizeof(double));
    f->ant = (double *)calloc((size_t)N * 3, sizeof(double));
    f->pos = (double *)calloc((size_t)N * N * 3 + 1, sizeof(double));
    f->ant_busy = (char *)calloc((size_t)N * 3, 1);
    f->pos_busy = (char *)calloc((size_t)N * N * 3 + 1, 1);
    return f;
}

static void fam_free(fam_t *f) {
    free(f->mo);
    free(f->fa);
    free(f->gender);
    free(f->nchild);
    free(f->child);
    free(f->nsp);
    free(f->sp);
    free(f->lk);
    free(f->single);
    free(f->post);
    free(f->ant);
    free(f->pos);
    free(f->ant_busy);
    free(f->pos_busy);
    free(f);
}

/* prior vector of individual i for this variant (family.cpp:1415-1481, :1052-1062) */
static const double *prior_of(const fam_t *f, int i, int known, int chrx) {
    if (chrx && f->gender[i] == 1) return known ? f->prior[3] : f->prior[2];
    return known ? f->prior[1] : f->prior[0];
}

/* transmission table for child c (family.cpp:1065-1072 and the ...X twins) */
static const double *table_of(const fam_t *f, int c, int chrx) {
    if (!chrx) return f->tA;
    return f->gender[c] == 1 ? f->tXm : f->tXf;
}

static double row_sum(const double *r) { /* dMatrix::sum_row, dMatrix.h:154-162 */
    double s = 0;
    for (int g = 0; g < 3; g++) s = s + r[g];
    return s;
}

/* family.cpp:1405-1499.  Writes dst (N x 3); returns 0 when some row sum is <= 0. */
static int single_posterior(const fam_t *f, int known, int chrx, double *dst) {
    for (int i = 0; i < f->N; i++) {
        const double *pr = prior_of(f, i, known, chrx);
        for (int g = 0; g < 3; g++) dst[i * 3 + g] = f->lk[i * 3 + g] * pr[g];
    }
    for (int i = 0; i < f->N; i++) {
        double s = row_sum(dst + i * 3);
        if (s <= 0) return 0;
        for (int g = 0; g < 3; g++) dst[i * 3 + g] = dst[i * 3 + g] / s;
    }
    return 1;
}

/* LRC gate, family.cpp:767-789 (same text at :1140-1162 and :1949-1971).
 * Returns 1 when the pedigree is to be ignored (every sequenced sample is "certain"). */
static int lrc_says_single(const fam_t *f) {
    for (int s = 0; s < f->S; s++) {
        const double *l = f->lk + f->cols[s] * 3;
        double big = 0, sum = 0;
        for (int g = 0; g < 3; g++) {
            if (big < l[g]) big = l[g];
            sum = sum + l[g];
        }
        big = big / sum;
        if (big < f->lc) return 0;
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* BN: exhaustive enumeration, family.cpp:882-954 (autosome) and :990-1120 (chrX)             */
/* ------------------------------------------------------------------------------------------ */
static int run_bn(fam_t *f, int known, int chrx) {
    const int N = f->N;
    int *g = (int *)calloc(N, sizeof(int));
    double *fac = (double *)calloc(N, sizeof(double));
    memset(f->post, 0, sizeof(double) * N * 3);
    for (;;) {
        for (int i = 0; i < N; i++) {
            if (f->mo[i] < 0)
                fac[i] = prior_of(f, i, known, chrx)[g[i]] * f->lk[i * 3 + g[i]];
            else
                fac[i] = T3(table_of(f, i, chrx), g[i], g[f->mo[i]], g[f->fa[i]]) * f->lk[i * 3 + g[i]];
        }
        double joint = 10000000;
        for (int i = 0; i < N; i++) joint = joint * fac[i];
        for (int i = 0; i < N; i++) f->post[i * 3 + g[i]] = f->post[i * 3 + g[i]] + joint;
        int d = 0; /* odometer, ped row 0 is the fastest digit (family.cpp:923-940) */
        while (d < N) {
            if (++g[d] == 3) {
                g[d] = 0;
                d++;
            } else
                break;
        }
        if (d == N) break;
    }
    free(g);
    free(fac);
    for (int i = 0; i < N; i++) {
        double s = row_sum(f->post + i * 3);
        if (s <= 0) return 0;
        for (int k = 0; k < 3; k++) f->post[i * 3 + k] = f->post[i * 3 + k] / s;
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* ES peeling: family.cpp:1255-1316 driver, :1501-1649 / :1651-1781 anterior,                 */
/* :1783-1845 / :1847-1930 posterior.  Memoised mutual recursion.                             */
/* ------------------------------------------------------------------------------------------ */
static double es_pos(fam_t *f, int i, int gi, int j, int chrx);

/* Pr(everything "above" i, g_i = gi).  Founders are pre-filled with their prior. */
static double es_ant(fam_t *f, int i, int gi, int chrx) {
    const int N = f->N;
    if (f->ant[i * 3 + gi] >= 0) return f->ant[i * 3 + gi];
    if (f->loop) return 0;
    if (f->ant_busy[i * 3 + gi]) {
        f->loop = 1;
        return 0;
    }
    f->ant_busy[i * 3 + gi] = 1;
    const int m = f->mo[i], p = f->fa[i];
    const double *ti = table_of(f, i, chrx);
    double over_m = 0;
    for (int a = 0; a < 3; a++) { /* mother's genotype */
        double over_f = 0;
        for (int b = 0; b < 3; b++) { /* father's genotype */
            double sibs = 1;
            for (int k = 0; k < f->nchild[m]; k++) { /* full sibs of i, in the mother's child order */
                int c = f->child[m * N + k];
                if (c == i || f->fa[c] != p) continue;
                const double *tc = table_of(f, c, chrx);
                double sc = 0;
                for (int l = 0; l < 3; l++) {
                    double below = 1;
                    for (int s = 0; s < f->nsp[c]; s++) below = below * es_pos(f, c, l, f->sp[c * N + s], chrx);
                    sc = sc + below * f->lk[c * 3 + l] * T3(tc, l, a, b);
                }
                sibs = sibs * sc;
            }
            double other_f = 1;
            for (int s = 0; s < f->nsp[p]; s++)
                if (f->sp[p * N + s] != m) other_f = other_f * es_pos(f, p, b, f->sp[p * N + s], chrx);
            over_f = over_f + es_ant(f, p, b, chrx) * f->lk[p * 3 + b] * other_f * T3(ti, gi, a, b) * sibs;
        }
        double other_m = 1;
        for (int s = 0; s < f->nsp[m]; s++)
            if (f->sp[m * N + s] != p) other_m = other_m * es_pos(f, m, a, f->sp[m * N + s], chrx);
        over_m = over_m + es_ant(f, m, a, chrx) * f->lk[m * 3 + a] * other_m * over_f;
    }
    f->ant_busy[i * 3 + gi] = 0;
    f->ant[i * 3 + gi] = over_m;
    return over_m;
}

/* Pr(everything "below" i through the marriage with j | g_i = gi). */
static double es_pos(fam_t *f, int i, int gi, int j, int chrx) {
    const int N = f->N;
    const size_t slot = ((size_t)i * N + j) * 3 + gi;
    if (f->pos[slot] >= 0) return f->pos[slot];
    if (f->loop) return 0;
    if (f->pos_busy[slot]) {
        f->loop = 1;
        return 0;
    }
    f->pos_busy[slot] = 1;
    double over_j = 0;
    for (int b = 0; b < 3; b++) { /* spouse's genotype */
        double other = 1;
        for (int s = 0; s < f->nsp[j]; s++)
            if (f->sp[j * N + s] != i) other = other * es_pos(f, j, b, f->sp[j * N + s], chrx);
        double kids = 1;
        for (int k = 0; k < f->nchild[i]; k++) { /* joint children, in i's child order */
            int c = f->child[i * N + k];
            if (f->mo[c] != j && f->fa[c] != j) continue;
            const double *tc = table_of(f, c, chrx);
            double sc = 0;
            for (int l = 0; l < 3; l++) {
                double below = 1;
                for (int s = 0; s < f->nsp[c]; s++) below = below * es_pos(f, c, l, f->sp[c * N + s], chrx);
                /* autosome: (g_i, g_j) whatever the sexes (family.cpp:1836);
                 * chrX: (mother, father) (family.cpp:1900-1921) */
                double tr = (chrx && f->gender[i] == 1) ? T3(tc, l, b, gi) : T3(tc, l, gi, b);
                sc = sc + tr * f->lk[c * 3 + l] * below;
            }
            kids = kids * sc;
        }
        over_j = over_j + es_ant(f, j, b, chrx) * f->lk[j * 3 + b] * other * kids;
    }
    f->pos_busy[slot] = 0;
    f->pos[slot] = over_j;
    return over_j;
}

/* returns 1 ok, 0 "false", and sets f->loop on a cyclic pedigree */
static int run_es(fam_t *f, int known, int chrx) {
    const int N = f->N;
    for (int k = 0; k < N * 3; k++) f->ant[k] = -1;
    for (size_t k = 0; k < (size_t)N * N * 3; k++) f->pos[k] = -1;
    memset(f->ant_busy, 0, (size_t)N * 3);
    memset(f->pos_busy, 0, (size_t)N * N * 3);
    f->loop = 0;
    for (int i = 0; i < N; i++)
        if (f->mo[i] < 0) {
            const double *pr = prior_of(f, i, known, chrx);
            for (int g = 0; g < 3; g++) f->ant[i * 3 + g] = pr[g];
        }
    memset(f->post, 0, sizeof(double) * N * 3);
    for (int i = 0; i < N; i++) {
        for (int g = 0; g < 3; g++) {
            double v = 1;
            for (int s = 0; s < f->nsp[i]; s++) v = v * es_pos(f, i, g, f->sp[i * N + s], chrx);
            v = v * f->lk[i * 3 + g] * es_ant(f, i, g, chrx);
            f->post[i * 3 + g] = v;
            if (f->loop) return 0;
        }
        double s = row_sum(f->post + i * 3);
        if (s == 0) return 0; /* note ==, family.cpp:1306 */
        for (int g = 0; g < 3; g++) f->post[i * 3 + g] = f->post[i * 3 + g] / s;
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* Gibbs sampler: family.cpp:1932-2096 driver, :2098-2299 one sweep                           */
/* ------------------------------------------------------------------------------------------ */
static inline uint32_t mulhilo(uint32_t a, uint32_t b, uint32_t *hi) {
    uint64_t p = (uint64_t)a * b;
    *hi = (uint32_t)(p >> 32);
    return (uint32_t)p;
}

/* Philox4x32-10 (Salmon et al., SC'11), the same constants as in the CUDA kernel. */
/* file.cpp:588-590 (= :825-827): lkTmp = atof(field); lkTmp = pow(10.0, -fabs(lkTmp) / 10.0); */
double fso_pl_decode(double x) { return pow(10.0, -fabs(x) / 10.0); }

void fso_pl_table(double *out, int n) {
    for (int k = 0; k < n; k++) out[k] = fso_pl_decode((double)k);
}

/* file.cpp:702-749 (= :938-997, :1814-1873): a posterior is printed as fabs(-10*log10(p)) through ostream's default
 * formatting (printf's %g, six significant digits), or as 99999 when -10*log10(p) is +infinity. */
int fso_phred_text(double p, char *buf) {
    const double v = -10 * log10(p);
    if (v == INFINITY) return sprintf(buf, "99999");
    return sprintf(buf, "%g", fabs(v));
}

void fso_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint32_t hi0, hi1;
        uint32_t lo0 = mulhilo(0xD2511F53u, c0, &hi0);
        uint32_t lo1 = mulhilo(0xCD9E8D57u, c2, &hi1);
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0;
        c1 = n1;
        c2 = n2;
        c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

/* Random numbers.  LIBC: rand(), exactly the reference's calls in the reference's order.  PHILOX: the word of
 * member i in sweep s of global site gv is word i%4 of Philox4x32-10(counter = (s, i/4, gv_lo, gv_hi), key = seed);
 * s = 0 initialises the genotypes, s >= 1 are the sweeps.  The CUDA kernel uses the same indexing, so a value
 * depends on (seed, site, sweep, member) only and not on the order in which members are visited. */
typedef struct {
    int kind;
    uint32_t key[2], gv[2];
    uint32_t buf[4];
    int64_t buf_sweep, buf_block; /* which Philox block `buf` holds */
} rng_t;

static void rng_start(rng_t *r, int kind, int64_t seed, int64_t gv) {
    r->kind = kind;
    r->key[0] = (uint32_t)((uint64_t)seed & 0xffffffffu);
    r->key[1] = (uint32_t)((uint64_t)seed >> 32);
    r->gv[0] = (uint32_t)((uint64_t)gv & 0xffffffffu);
    r->gv[1] = (uint32_t)((uint64_t)gv >> 32);
    r->buf_sweep = r->buf_block = -1;
}

static uint32_t rng_u32(rng_t *r, int sweep, int member) {
    if (r->buf_sweep != sweep || r->buf_block != member / 4) {
        const uint32_t ctr[4] = {(uint32_t)sweep, (uint32_t)(member / 4), r->gv[0], r->gv[1]};
        fso_philox4x32(ctr, r->key, r->buf);
        r->buf_sweep = sweep;
        r->buf_block = member / 4;
    }
    return r->buf[member % 4];
}

static int rng_init_genotype(rng_t *r, int member) {
    if (r->kind == FSO_RNG_LIBC) return rand() % 3; /* family.cpp:2063-2067 */
    return (int)(rng_u32(r, 0, member) % 3u);
}

static double rng_uniform(rng_t *r, int sweep, int member) {
    if (r->kind == FSO_RNG_LIBC) return (double)rand() / (double)RAND_MAX; /* family.cpp:2161 */
    return ((double)(rng_u32(r, sweep, member) >> 1) + 0.5) * (1.0 / 2147483648.0); /* 31 bits, like rand() / RAND_MAX */
}

/* Test probe: smallest and largest weight sum `s` met by the chains of the last fso_run(MCMC) call (all its variants).
 * The generated CUDA Gibbs kernel hands a chain back to the table-driven kernel when a sum leaves the positive normal
 * range [2^-963, 2^963); the tests use this probe to check that it does so exactly when it must. */
static double g_sum_min = 0, g_sum_max = 0;
static long long g_steps = 0, g_changes = 0; /* Gibbs steps / steps that changed the member's genotype, same call */
void fso_mcmc_sum_range(double out[2]) {
    out[0] = g_sum_min;
    out[1] = g_sum_max;
}
void fso_mcmc_change_count(long long out[2]) {
    out[0] = g_steps;
    out[1] = g_changes;
}

/* one sweep over all individuals in ped order */
static void gibbs_sweep(const fam_t *f, int *cur, double *acc, int known, int chrx, rng_t *rng, int sweep) {
    const int N = f->N;
    for (int i = 0; i < N; i++) {
        double w[3] = {1000000, 1000000, 1000000};
        const int male = f->gender[i] == 1;
        for (int g = 0; g < 3; g++) {
            if (f->mo[i] < 0)
                w[g] = w[g] * prior_of(f, i, known, chrx)[g] * f->lk[i * 3 + g];
            else
                w[g] = w[g] * T3(table_of(f, i, chrx), g, cur[f->mo[i]], cur[f->fa[i]]) * f->lk[i * 3 + g];
            /* chrX: only males get the children factor (reference quirk, family.cpp:2230-2257) */
            if (chrx && !male) continue;
            for (int k = 0; k < f->nchild[i]; k++) {
                int c = f->child[i * N + k];
                const double *tc = table_of(f, c, chrx);
                if (male)
                    w[g] = w[g] * T3(tc, cur[c], cur[f->mo[c]], g);
                else
                    w[g] = w[g] * T3(tc, cur[c], g, cur[f->fa[c]]);
            }
        }
        double s = 0;
        for (int g = 0; g < 3; g++) s = s + w[g];
        if (!(s >= g_sum_min)) g_sum_min = s; /* NaN sticks */
        if (!(s <= g_sum_max)) g_sum_max = s;
        if (s <= 0)
            w[0] = w[1] = w[2] = 0;
        else
            for (int g = 0; g < 3; g++) w[g] = w[g] / s;
        double rd = rng_uniform(rng, sweep, i);
        const int before = cur[i];
        if (rd < w[0])
            cur[i] = 0;
        else if (rd > (1.0 - w[2]))
            cur[i] = 2;
        else
            cur[i] = 1;
        g_steps++;
        g_changes += cur[i] != before;
        for (int g = 0; g < 3; g++) acc[i * 3 + g] = acc[i * 3 + g] + w[g];
    }
}

static int run_mcmc(fam_t *f, int known, int chrx, int burn, int rep, rng_t *rng) {
    const int N = f->N;
    int *cur = (int *)calloc(N, sizeof(int));
    double *acc = (double *)calloc((size_t)N * 3, sizeof(double));
    for (int i = 0; i < N; i++) cur[i] = rng_init_genotype(rng, i);
    for (int t = 0; t < burn; t++) gibbs_sweep(f, cur, acc, known, chrx, rng, 1 + t);
    memset(acc, 0, sizeof(double) * N * 3);
    for (int t = 0; t < rep; t++) gibbs_sweep(f, cur, acc, known, chrx, rng, 1 + burn + t);
    int ok = 1;
    for (int i = 0; i < N && ok; i++) {
        for (int g = 0; g < 3; g++) f->post[i * 3 + g] = acc[i * 3 + g] / rep;
        if (row_sum(f->post + i * 3) <= 0) ok = 0;
    }
    free(cur);
    free(acc);
    return ok;
}

/* ------------------------------------------------------------------------------------------ */
/* Batch entry points                                                                         */
/* ------------------------------------------------------------------------------------------ */
int fso_topology(int N, const int *ped_id, const int *ped_mid, const int *ped_fid, const int *gender,
                 int *mother, int *father) {
    fam_t *f = fam_new(N, 0);
    memcpy(f->gender, gender, sizeof(int) * N);
    int rc = build_topology(f, ped_id, ped_mid, ped_fid);
    if (rc == FSO_OK || rc == FSO_E_GENDER) {
        memcpy(mother, f->mo, sizeof(int) * N);
        memcpy(father, f->fa, sizeof(int) * N);
    }
    fam_free(f);
    return rc;
}

int fso_run(int method, int N, const int *ped_id, const int *ped_mid, const int *ped_fid, const int *gender,
            int S, const int *cols, double mrate, double lc, const double *priors, int64_t V,
            const uint8_t *flags, const double *lk, int burn, int rep, int rng_kind, int64_t seed,
            int64_t v_offset, double *post, double *single, int32_t *gt, uint8_t *status, double *post_full,
            double *single_full) {
    if (N <= 0 || S < 0 || method < 1 || method > 3) return FSO_E_ARG;
    fam_t *f = fam_new(N, S);
    memcpy(f->gender, gender, sizeof(int) * N);
    int rc = build_topology(f, ped_id, ped_mid, ped_fid);
    if (rc != FSO_OK) {
        fam_free(f);
        return rc;
    }
    f->cols = cols;
    f->lc = lc;
    memcpy(f->prior, priors, sizeof(double) * 12);
    fso_tables(mrate, f->tA, f->tXf, f->tXm);
    if (method == FSO_MCMC && rng_kind == FSO_RNG_LIBC && seed >= 0) srand((unsigned)seed);
    g_sum_min = INFINITY;
    g_sum_max = -INFINITY;
    g_steps = g_changes = 0;

    for (int64_t v = 0; v < V; v++) {
        const int known = flags[v] & 1, chrx = (flags[v] >> 1) & 1;
        for (int k = 0; k < N * 3; k++) f->lk[k] = 1; /* unsequenced members: (1,1,1), file.cpp:565 */
        for (int s = 0; s < S; s++)
            for (int g = 0; g < 3; g++) f->lk[cols[s] * 3 + g] = lk[(v * S + s) * 3 + g];

        int ok = single_posterior(f, known, chrx, f->single);
        if (ok) {
            if (lrc_says_single(f)) { /* family.cpp:793-878: FPP := individual posterior */
                ok = single_posterior(f, known, chrx, f->post);
            } else if (method == FSO_BN) {
                ok = run_bn(f, known, chrx);
            } else if (method == FSO_ES) {
                ok = run_es(f, known, chrx);
                if (f->loop) {
                    fam_free(f);
                    return FSO_E_LOOP;
                }
            } else {
                rng_t rng;
                rng_start(&rng, rng_kind, seed, v_offset + v);
                ok = run_mcmc(f, known, chrx, burn, rep, &rng);
            }
        }
        status[v] = ok ? 0 : 1;
        for (int s = 0; s < S; s++) {
            const double *pp = f->post + cols[s] * 3;
            for (int g = 0; g < 3; g++) {
                post[(v * S + s) * 3 + g] = ok ? pp[g] : 0;
                single[(v * S + s) * 3 + g] = ok ? f->single[cols[s] * 3 + g] : 0;
            }
            /* get_postRlt, family.cpp:636-665: strict '<' from -1, first maximum wins, NaN -> -1 */
            double big = -1;
            int arg = -1;
            for (int g = 0; g < 3; g++)
                if (big < pp[g]) {
                    big = pp[g];
                    arg = g;
                }
            gt[v * S + s] = ok ? arg : -1;
        }
        if (post_full)
            for (int k = 0; k < N * 3; k++) post_full[v * N * 3 + k] = ok ? f->post[k] : 0;
        if (single_full)
            for (int k = 0; k < N * 3; k++) single_full[v * N * 3 + k] = ok ? f->single[k] : 0;
    }
    fam_free(f);
    return FSO_OK;
}
