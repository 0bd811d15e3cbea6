/*
 * kmerlr_oracle.h -- CPU ORACLE for the kmerLr hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the reference's algorithm (pbenner/kmerLr, Go) for the
 * hot path named by BASELINE.json:north_star.  It is used ONLY by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs, as the
 * CHECKER and the CPU baseline.  The product (kmerlr_b200/) never includes, links or calls it.
 *
 * Parity status: PINNED for k-mer indices / class names / counts (kmerLr_test.go:40-43,55-66),
 * CoeffIndex (round trip), leapfrog lambda (README.md:39 = 2.496875), the TestKmers6 tie group
 * and Go<=1.18 sort order (kmerLr_test.go:205-206), the loss/prediction known answers
 * (kmerLr_test.go:224,228,248), the TestKmers5 loss under the standardizer (kmerLr_test.go:186, numpy
 * restatement of the transform in oracle.py).  The reference itself cannot be built here (no Go toolchain;
 * gonetics / autodiff are not vendored), so there is no oracle/_ref.  UNPINNED (no reference
 * test exists): non-ACGT input handling, --complement / --reverse alone, MaxAmbiguous >= 0,
 * estimate_proximal / estimate_coordinate (dead code in the reference, no test; estimate_coordinate is
 * restated in numpy, oracle.py:coordinate, and pinned by optimality properties only), genomic scoring.
 *
 * Every function cites the reference file:line it follows.
 */
#ifndef KMERLR_ORACLE_H
#define KMERLR_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* NewKmerCounter(M, N, complement, reverse, revcomp, maxAmbiguous, alphabet)  kmerLr_learn.go:94 */
typedef struct {
  int32_t M, N;
  int32_t complement, reverse, revcomp;
  int32_t binarize;      /* scan_sequence: IdentifyKmers instead of CountKmers  kmerLr_data.go:257-263 */
  int32_t alphabet;      /* 0 = NucleotideAlphabet, 1 = GappedNucleotideAlphabet */
  int32_t max_ambiguous; /* -1 = nil (unlimited) */
} ko_config;

typedef struct ko_matrix ko_matrix;

/* ---- stage 1: FASTA sequences -> sparse rows (kmerLr_data.go:197-357) ---------------------- */

/* compile_training_data / compile_test_data restated.  seq = concatenated ASCII, off[n+1].
 * frozen_k/frozen_code (n_frozen > 0): counter frozen to that class list, column order = list
 * order (SetKmers, kmerLr_data.go:317-319,333).  features (n_features > 0): explicit feature
 * list, pairs index into the class list (convert_counts, kmerLr_data.go:210-229).
 * faithful != 0 walks all union classes per sample with one map lookup each, exactly like
 * convert_counts (kmerLr_data.go:204-209); faithful == 0 gives the same rows faster. */
ko_matrix *ko_extract(const ko_config *cfg, const uint8_t *seq, const int64_t *off, int64_t n,
                      const int32_t *frozen_k, const uint64_t *frozen_code, int64_t n_frozen,
                      const int32_t *features, int64_t n_features, int threads, int faithful);
void    ko_matrix_free(ko_matrix *m);
/* n rows, m columns (without the bias column), nnz (without bias entries), n_classes */
void    ko_matrix_info(const ko_matrix *m, int64_t *n, int64_t *ncol, int64_t *nnz, int64_t *n_classes);
void    ko_matrix_classes(const ko_matrix *m, int32_t *k_out, uint64_t *code_out);
/* CSR without the bias: Go index j+1 <-> col j; values are the float64 the reference stores */
void    ko_matrix_rows(const ko_matrix *m, int64_t *rowptr, int32_t *col, double *val);
/* build a matrix from CSR (used for the scoresLr-style dense goldens and for tests) */
ko_matrix *ko_matrix_from_csr(int64_t n, int64_t ncol, const int64_t *rowptr, const int32_t *col, const double *val);

/* class name as printed by the reference ("gntanc|gntanc"); returns strlen */
int     ko_class_name(const ko_config *cfg, int32_t k, uint64_t code, char *buf, int buflen);
/* canonical class code of the k-mer given as letter codes (a=0,c=1,g=2,t=3,n=4) */
uint64_t ko_class_code(const ko_config *cfg, const uint8_t *letters, int32_t k);

/* ---- CoeffIndex (kmerLr_coefficients_index.go:26-54) ---------------------------------------- */
int64_t ko_coeff_dim(int64_t n);
int64_t ko_coeff_ind2sub(int64_t n, int64_t k1, int64_t k2);
void    ko_coeff_sub2ind(int64_t n, int64_t i, int64_t *k1, int64_t *k2);

/* ---- stage 2: logistic regression (kmerLr_logistic_regression.go:47-272) -------------------- */
/* theta has ncol+1 entries (or CoeffIndex(ncol).Dim() when cooccurrence != 0) */
void    ko_linear_pdf(const ko_matrix *m, const double *theta, int cooccurrence, double *out);
void    ko_log_pdf   (const ko_matrix *m, const double *theta, int cooccurrence, double *out);
void    ko_gradient  (const ko_matrix *m, const uint8_t *labels, const double *theta, int64_t ntheta,
                      const double cw[2], double lambda, int cooccurrence, double *g);
double  ko_loss      (const ko_matrix *m, const uint8_t *labels, const double *theta, int64_t ntheta,
                      const double cw[2], double lambda, int cooccurrence);
void    ko_class_weights(const uint8_t *labels, int64_t n, double cw[2]); /* kmerLr_data.go:178-193 */

/* ---- selection (kmerLr_feature_selection.go:78-134,179-219,309-343; kmerLr_sort.go:120-131) -- */
#define KO_TIE_GO118 0   /* emulate Go <= 1.18 sort.Sort(sort.Reverse(AbsFloatInt)) */
#define KO_TIE_INDEX 1   /* |g| descending, then coefficient index ascending */
/* x is sorted in place (like the reference), idx receives the permutation */
void    ko_nlargest_abs(double *x, int64_t *idx, int64_t len, int tie);
/* featureSelector.Select: active_idx are full-space coefficient indices (>= 1) of the current
 * model with their theta; mask (ntheta bytes) receives b; returns ok */
int     ko_select(const ko_matrix *m, const uint8_t *labels, const double cw[2], int cooccurrence,
                  int64_t N, double theta0, const int64_t *active_idx, const double *active_theta,
                  int64_t n_active, int tie, double epsilon_lambda, double prev_lambda,
                  uint8_t *mask, int64_t ntheta, double *lambda_out, int64_t *c_out, double *g_out);
/* featureSelection.Data: sel[0] = 0 (bias), ascending coefficient indices */
ko_matrix *ko_reduce(const ko_matrix *m, const int64_t *sel, int64_t nsel);

/* ---- proximal gradient (kmerLr_estimator_proximal.go:30-120, aliasing fixed; hook :46-99) --- */
typedef struct {
  double  loss_old, loss_new;  /* hook closure state, persists across epochs */
} ko_hook_state;
double  ko_step_size(const ko_matrix *m, double l2, double step_factor);
/* returns iterations done; theta updated in place; trace (may be NULL) gets [iter] = loss */
int64_t ko_proxgrad(const ko_matrix *reduced, const uint8_t *labels, double *theta, const double cw[2],
                    double lambda, double l2, double step_factor, double epsilon, double epsilon_loss,
                    int64_t max_iter, ko_hook_state *hook, double *delta_out);

/* estimate_loop (kmerLr_estimator.go:209-255) with ISTA as the inner solver.  State in/out:
 * theta0 + active set.  path_lambda/path_nsel record one entry per epoch (cap path_cap). */
typedef struct {
  int64_t  n_active;       /* number of reduced features (including theta == 0 ones) */
  int64_t *active_idx;     /* full-space coefficient index of each (cap = state_cap) */
  double  *theta;          /* theta[0] = bias, theta[1+i] for active_idx[i]            */
  int64_t  state_cap;
  ko_hook_state hook;
  double   l1reg_over_n;   /* obj.L1Reg / n carried between targets */
} ko_estimator;
int64_t ko_estimate_loop(const ko_matrix *m, const uint8_t *labels, const double cw[2], int cooccurrence,
                         int64_t N, int tie, double epsilon_lambda, double l2, double step_factor,
                         double epsilon, double epsilon_loss, int64_t max_iter, int64_t max_epochs,
                         ko_estimator *est, double *path_lambda, int64_t *path_iters, int64_t path_cap);

/* ---- stage 3: genomic sliding-window scoring (kmerLr_predict_genomic.go:134-171) ------------ */
#define KO_SUMMARY_NONE 0
#define KO_SUMMARY_MEAN 1
#define KO_SUMMARY_PRODUCT 2
#define KO_SUMMARY_MIN 3
#define KO_SUMMARY_MAX 4
typedef struct {
  ko_config       cfg;
  int64_t         n_classes;
  const int32_t  *class_k;
  const uint64_t *class_code;
  int64_t         n_features;
  const int32_t  *features;    /* n_features x 2 */
  int64_t         n_members;
  const double   *theta;       /* n_members x (n_features+1) */
  int32_t         summary;
} ko_model;
/* number of output slots of a region of length len: n = len-W > 0 ? n/step+1 : 0  (:152-156) */
int64_t ko_window_slots(int64_t len, int64_t W, int64_t step);
/* out is laid out region after region, ko_window_slots() each; untouched slots are 0.0 */
void    ko_score_windows(const ko_model *models, int n_models, const uint8_t *seq, const int64_t *region_off,
                         int64_t n_regions, int64_t W, int64_t step, double *out, int threads);

#ifdef __cplusplus
}
#endif
#endif
