/* oracle/famseq_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the reference's per-variant pedigree posterior engine
 * (`class family`, /root/reference/src/family.cpp).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.  The product
 * (famseq_b200/) never links, imports or executes anything from oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks this restatement bit-for-bit
 * against the unmodified reference engine (oracle/_ref/ref_harness) for BN, ES and MCMC
 * (same libc rand() stream) on TestData-derived and synthetic inputs; the committed golden
 * vectors under tests/golden/ were produced by the reference itself.
 */
#ifndef FAMSEQ_ORACLE_H_
#define FAMSEQ_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { FSO_BN = 1, FSO_ES = 2, FSO_MCMC = 3 };
enum { FSO_RNG_LIBC = 0, FSO_RNG_PHILOX = 1 };

/* error codes (negative) */
enum {
    FSO_OK = 0,
    FSO_E_HALF_PARENTS = -1, /* family.cpp:318-322 "not a fulfill family" */
    FSO_E_GENDER = -2,       /* family.cpp:204-219 checkPed */
    FSO_E_LOOP = -3,         /* ES on a looped pedigree: the reference recurses forever */
    FSO_E_ARG = -4
};

/* Mendelian transmission tables, layout t[g*9 + a*3 + b] = Pr(child = g | mother = a, father = b).
 * family.cpp:447-550 (autosome), :383-416 (X, female child), :418-445 (X, male child). */
void fso_tables(double mrate, double *pcp2, double *pcp2xf, double *pcp2xm);

/* Run one method over a batch.
 *  ped_id/ped_mid/ped_fid/gender : the N ped rows (ids as in the ped file; 0 = no parent)
 *  cols[S]   : ped row of every sequenced input column, in input-column order (mapV2P >= 0 entries)
 *  priors    : [4][3] = genoProbN, genoProbK, genoProbXN, genoProbXK
 *  flags[v]  : bit0 Known, bit1 chrX
 *  lk        : [V][S][3] raw likelihoods; unsequenced members are (1,1,1)
 *  rng_kind  : MCMC only. LIBC = rand() exactly as the reference (srand(seed) first if seed >= 0);
 *              PHILOX = counter-based: word i%4 of Philox(counter (sweep, i/4, site), key seed), as the CUDA kernel
 *  post/single [V][S][3], gt [V][S], status[V] (1 = the reference would have returned false);
 *  post_full/single_full [V][N][3] may be NULL.
 * Returns 0 or a negative FSO_E_* code. */
int fso_run(int method, int N, const int *ped_id, const int *ped_mid, const int *ped_fid, const int *gender,
            int S, const int *cols, double mrate, double lc, const double *priors, int64_t V,
            const uint8_t *flags, const double *lk, int burn, int rep, int rng_kind, int64_t seed,
            int64_t v_offset, double *post, double *single, int32_t *gt, uint8_t *status, double *post_full,
            double *single_full);

/* Topology as the reference builds it (family.cpp:291-350): mother/father row (-1 = founder).
 * Returns 0 or a negative code; used by tests of the host-side pedigree builder. */
int fso_topology(int N, const int *ped_id, const int *ped_mid, const int *ped_fid, const int *gender,
                 int *mother, int *father);

/* The VCF driver's likelihood decode (file.cpp:588-590, :825-827): pow(10, -fabs(x)/10) of a PL / GL field. */
double fso_pl_decode(double x);
/* ... for every integer PL 0 .. n-1 at once (the table the engine's fs_run_pl decodes through). */
void fso_pl_table(double *out, int n);

/* The drivers' Phred text of one posterior (file.cpp:702-749): "%g" of fabs(-10*log10(p)), "99999" for +inf.
 * buf needs 32 bytes; returns the length. */
int fso_phred_text(double p, char *buf);

/* Test probe: {min, max} of the Gibbs weight sums met during the last fso_run(FSO_MCMC) call (over all its variants). */
void fso_mcmc_sum_range(double out[2]);
/* Test probe: {Gibbs steps, steps that changed the member's genotype} of the last fso_run(FSO_MCMC) call. */
void fso_mcmc_change_count(long long out[2]);

/* Philox4x32-10 block, exposed so the host/CUDA implementations can be checked against it. */
void fso_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif
