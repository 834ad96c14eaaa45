/*
 * ba_oracle.h -- CPU restatement ("port") of the reference's LM bundle-adjustment hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (libcuba_b200.so) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement against the
 * reference README's published chi^2 table for ba_kitti_00 (README.md:141-150, kernel NONE) and,
 * on the GPU box, against the reference's own code compiled unmodified (oracle/_ref/libcuba_ref.so).
 *
 * Every function cites the reference file:line it restates (paths relative to /root/reference).
 */
#ifndef BA_ORACLE_H
#define BA_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* Flat problem, exactly what CudaBlockSolver::initialize (src/cuda_bundle_adjustment.cpp:115-261)
 * produces: vertices indexed by iP / iL (free first, fixed appended), edges carry (iP, iL). */
typedef struct bao_problem {
	int Pall, numP;          /* all poses / free poses  (iP >= numP  <=> fixed)        */
	int Lall, numL;          /* all landmarks / free landmarks                          */
	const double* q;         /* [4*Pall] quaternion x,y,z,w                             */
	const double* t;         /* [3*Pall]                                                */
	const double* cam;       /* [5*Pall] fx,fy,cx,cy,bf                                 */
	const double* Xw;        /* [3*Lall]                                                */
	int E2;                  /* monocular edges (edge ids 0..E2-1)                      */
	const int* idx2;         /* [2*E2] (iP,iL)                                          */
	const double* meas2;     /* [2*E2]                                                  */
	const double* omega2;    /* [E2] scalar information                                 */
	int E3;                  /* stereo edges (edge ids E2..E2+E3-1)                     */
	const int* idx3;         /* [2*E3]                                                  */
	const double* meas3;     /* [3*E3]                                                  */
	const double* omega3;    /* [E3]                                                    */
} bao_problem;

typedef struct bao bao;

/* robust kernel types: 0 NONE, 1 HUBER, 2 TUKEY (src/cuda_block_solver.cu:679-727) */
bao* bao_create(const bao_problem* prob, const int rk_type[2], const double rk_delta[2]);
void bao_destroy(bao* h);

/* whole optimize() (src/cuda_bundle_adjustment.cpp:793-857). Returns number of batch statistics written.
 * chi2_out[niter], lambda_out[niter] (lambda after each outer iteration), trials_out[niter] (trial count). */
int bao_optimize(bao* h, int niter, double* chi2_out, double* lambda_out, int* trials_out);

/* stage-wise entry points for stage parity tests */
double bao_compute_errors(bao* h);                 /* cpp:368-382, cu:732-786 */
void bao_build_system(bao* h);                     /* cpp:384-410, cu:788-839 */
double bao_max_diagonal(bao* h);                   /* cpp:412-418, cu:877-904 */
int bao_solve(bao* h, double lambda);              /* cpp:432-481 (direct block Cholesky instead of cuSOLVER) */
void bao_update(bao* h);                           /* cpp:483-492, cu:1045-1068 */
double bao_compute_scale(bao* h, double lambda);   /* cu:1070-1091 (without the +1e-3) */
void bao_chi_sqs(bao* h, double* out);             /* cu:841-875; out[E2+E3] in edge-id order */

/* sizes */
int bao_nhpl(const bao* h);     /* free-free edges = Hpl blocks */
int bao_nblk(const bao* h);     /* upper-triangular Hsc blocks  */
int bao_nmul(const bao* h);     /* block products               */

/* index structures (bit-exact contract) */
void bao_get_hpl_structure(const bao* h, int* colPtr /*numL+1*/, int* rowInd /*nhpl*/, int* edge2Hpl /*E2+E3, -1 if none*/);
void bao_get_hsc_structure(const bao* h, int* rowPtr /*numP+1*/, int* colInd /*nblk*/);

/* numeric arrays, column-major blocks like the reference (MatView, cu:79-85) */
void bao_get_state(const bao* h, double* q, double* t, double* Xw);
void bao_set_state(bao* h, const double* q, const double* t, const double* Xw);
void bao_get_system(const bao* h, double* Hpp /*36*numP*/, double* bp /*6*numP*/, double* Hll /*9*numL*/,
	double* bl /*3*numL*/, double* Hpl /*18*nhpl, CSC order*/);
void bao_get_schur(const bao* h, double* Hsc /*36*nblk*/, double* bsc /*6*numP*/, double* invHll /*9*numL*/);
void bao_get_delta(const bao* h, double* xp /*6*numP*/, double* xl /*3*numL*/);

/* number of OpenMP threads actually used by the edge loops (1 when built without -fopenmp) */
int bao_threads(void);

#ifdef __cplusplus
}
#endif
#endif
