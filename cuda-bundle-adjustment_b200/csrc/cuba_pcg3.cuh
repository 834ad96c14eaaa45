// cuba_pcg3.cuh -- third-generation block-Jacobi PCG: k_pcg2 without grid barriers inside the iteration.
//
// Same algorithm and data placement as k_pcg2 (split-preconditioned Chronopoulos-Gear CG, A^ resident in
// shared memory, one CTA per SM).  What changes is the exchange between CTAs: every value another CTA needs
// (the six entries of w = A^ r per owned row, and the two partial inner products per CTA) is published as
// "low-latency" words -- each fp64 is split in two 32-bit halves, each half travels in an 8-byte word next
// to a 32-bit sequence tag (the pass number) -- so a consumer simply polls the payload in L2 until both tags
// match.  8-byte stores are single-copy atomic, hence no __threadfence / MEMBAR / L1 invalidation and no
// barrier: an iteration costs one L2 store->load hop instead of a GPU-wide barrier (~2 us on B200).
// r and s of the needed columns are kept in shared memory by every consumer and advanced locally.
// Buffers are double-buffered by pass parity; a CTA can never run more than one pass ahead of any other,
// because it needs everybody's partial sums of pass k-1 to start pass k.
#pragma once

#include "cuba_pcg2.cuh"

namespace cuba_b200 {

#ifdef CUBA_PCG_TIMING
#define PCG_T(var) const long long var = clock64()
#define PCG_ACC(slot, a, b) do { if (threadIdx.x == 0) tacc[slot] += (b) - (a); } while (0)
#else
#define PCG_T(var)
#define PCG_ACC(slot, a, b)
#endif

constexpr unsigned int PCG3_SPIN_LIMIT = 1u << 24;   // ~1 s of polling: a lost peer aborts the solve instead of hanging

__device__ __forceinline__ void ll_store(unsigned long long* slot, double v, unsigned int tag)
{
	const unsigned long long b = (unsigned long long)__double_as_longlong(v);
	const unsigned long long lo = (b & 0xffffffffull) | ((unsigned long long)tag << 32);
	const unsigned long long hi = (b >> 32) | ((unsigned long long)tag << 32);
	asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" :: "l"(slot), "l"(lo), "l"(hi) : "memory");
}
__device__ __forceinline__ bool ll_try_load(const unsigned long long* slot, unsigned int tag, double& v)
{
	unsigned long long lo, hi;
	asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(slot) : "memory");
	if ((unsigned int)(lo >> 32) != tag || (unsigned int)(hi >> 32) != tag) return false;
	v = __longlong_as_double((long long)((lo & 0xffffffffull) | (hi << 32)));
	return true;
}
// polls until the tag matches; returns false when the solve was aborted (or this thread gave up and aborts it)
__device__ __forceinline__ bool ll_wait(const unsigned long long* slot, unsigned int tag, double& v, int* abortFlag)
{
	for (unsigned int spin = 0;; spin++) {
		if (ll_try_load(slot, tag, v)) return true;
		if ((spin & 1023u) == 1023u) {
			if (*(volatile int*)abortFlag) return false;
			if (spin >= PCG3_SPIN_LIMIT) { atomicExch(abortFlag, 1); return false; }
		}
	}
}

template <typename T>
struct Pcg3Args {
	Pcg2Args<T> base;
	unsigned long long* wFlag;   // [2][6*numP][2]  published w entries
	unsigned long long* pFlag;   // [2][PCG3_REPL][2*G][2]  published partial inner products (replicated board)
	int* abortFlag;              // zeroed before the launch together with wFlag/pFlag
	long long* timing;           // [G][8] per-phase clock64 sums (only with -DCUBA_PCG_TIMING)
};

constexpr int PCG3_BLOCK = 256;                 // threads per CTA: 255 registers each, enough for two 6x6 fp64 blocks
constexpr int PCG3_BPT = 2;                     // register-resident A^ blocks per thread
constexpr int PCG3_REGBLK = PCG3_BLOCK * PCG3_BPT;
constexpr int PCG3_CHUNK = PCG3_REGBLK;         // block products staged per round
constexpr int PCG3_REPL = 8;                    // replicas of the partial-product board (readers per L2 line / 8)
constexpr int PCG3_WPT = 4;                     // polled w items per thread (need list up to 170 columns without extra rounds)

template <typename T>
__global__ void __launch_bounds__(PCG3_BLOCK, 1) k_pcg3(const Pcg3Args<T> aa)
{
	const Pcg2Args<T>& a = aa.base;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	T* s_blk = reinterpret_cast<T*>(smem_raw);                          // [36][capBlocks] element-major cache of the blocks past the registers
	T* s_r = s_blk + (size_t)a.capBlocks * 36;                          // [needMax][6] residual of the needed columns
	T* s_s = s_r + (size_t)a.needMax * 6;                               // [needMax][6] s = w + beta s
	T* s_p = s_s + (size_t)a.needMax * 6;                               // [maxRows][6] search direction of the own rows
	T* s_y = s_p + (size_t)a.maxRows * 6;                               // [maxRows][6] iterate (hat space) of the own rows
	T* s_c = s_y + (size_t)a.maxRows * 6;                               // [PCG3_CHUNK][6] block-product contributions
	int* s_loc = reinterpret_cast<int*>(s_c + (size_t)PCG3_CHUNK * 6);  // [capBlocks]  need index of a cached block's column (<0: diagonal)
	int* s_rowPtr = s_loc + a.capBlocks;                                // [maxRows+1]
	int* s_need = s_rowPtr + a.maxRows + 1;                             // [needMax] global column of each need entry
	int* s_own = s_need + a.needMax;                                    // [needMax] local own row of a need entry, or -1
	int* s_diag = s_own + a.needMax;                                    // [maxRows] need index of each own row
	__shared__ double s_red[PCG3_BLOCK / 32][2];
	__shared__ double s_w2[16][2];
	__shared__ double s_bc[2];
	__shared__ unsigned int s_gen;
	__shared__ int s_abort;

	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const int G = gridDim.x, cta = blockIdx.x;
	const int row0 = a.ctaRow[cta], row1 = a.ctaRow[cta + 1], nrows = row1 - row0;
	const int need0 = a.needPtr[cta], nneed = a.needPtr[cta + 1] - need0;
	const int blk0 = a.fRowPtr[row0], nblkCta = a.fRowPtr[row1] - blk0;
	// shared-memory cache for the blocks past the register-resident first PCG3_REGBLK ones
	const int ncached = nblkCta > PCG3_REGBLK ? (nblkCta - PCG3_REGBLK < a.capBlocks ? nblkCta - PCG3_REGBLK : a.capBlocks) : 0;
	const size_t n6 = 6 * (size_t)a.numP;
	if (tid == 0) { s_gen = ld_acquire_u32(&a.bar->gen); s_abort = 0; }
	for (int i = tid; i <= nrows; i += PCG3_BLOCK) s_rowPtr[i] = a.fRowPtr[row0 + i] - blk0;
	for (int i = tid; i < nneed; i += PCG3_BLOCK) {
		const int j = a.needCol[need0 + i];
		s_need[i] = j;
		s_own[i] = (j >= row0 && j < row1) ? j - row0 : -1;
	}
	for (int i = tid; i < nrows * 6; i += PCG3_BLOCK) { s_p[i] = T(0); s_y[i] = T(0); }
	__syncthreads();
	for (int i = tid; i < nneed; i += PCG3_BLOCK) if (s_own[i] >= 0) s_diag[s_own[i]] = i;
	unsigned int gen = s_gen;

	// ---- S1: factor the diagonal blocks of the own rows, b^ = L^-1 b ------------------------------------
	int bad = 0;
	for (int i = row0 + tid; i < row1; i += PCG3_BLOCK) {
		int d = -1;
		for (int n = a.fRowPtr[i]; n < a.fRowPtr[i + 1]; n++) if (a.fColInd[n] == i) { d = n; break; }
		T Li[36];
		bool ok = d >= 0 && chol6_inverse_factor(a.fVal + 36 * (size_t)d, Li);
		if (!ok) { bad = 1; for (int e = 0; e < 36; e++) Li[e] = (e % 7) == 0 ? T(1) : T(0); }
		for (int e = 0; e < 36; e++) a.Linv[36 * (size_t)i + e] = Li[e];
		for (int r = 0; r < 6; r++) {
			T s = T(0);
			for (int c = 0; c <= r; c++) s += Li[c * 6 + r] * a.b[6 * (size_t)i + c];
			a.R0[6 * (size_t)i + r] = s;
		}
	}
	{
		const int anyBad = __syncthreads_or(bad);
		if (tid == 0) a.partial[(size_t)cta * 2] = (double)anyBad;
	}
	grid_barrier(a.bar, G, gen);
	double nbad = 0;
	if (tid < 32) {
		for (int i = tid; i < G; i += 32) nbad += __ldcg(a.partial + (size_t)i * 2);
		nbad = warp_sum(nbad);
		if (tid == 0) s_bc[0] = nbad;
	}
	__syncthreads();
	nbad = s_bc[0];
	__syncthreads();

	// ---- S2: A^_ij = L_i^-1 S_ij L_j^-T for the own rows; r0 of the needed columns.
	//      Blocks n < PCG3_REGBLK stay in REGISTERS for the whole solve (thread n % PCG3_BLOCK, slot n / PCG3_BLOCK),
	//      later ones go to shared memory (element-major) and, past its capacity, to the global copy.
	T breg[PCG3_BPT][36];
	int myLoc[PCG3_BPT];
#pragma unroll
	for (int u = 0; u < PCG3_BPT; u++) {
		myLoc[u] = -1;
#pragma unroll
		for (int e = 0; e < 36; e++) breg[u][e] = T(0);
	}
	{
		auto transform = [&](int n, T* out) {
			const int g = blk0 + n;
			int lo = 0, hi = nrows - 1;
			while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (s_rowPtr[mid] <= n) lo = mid; else hi = mid - 1; }
			const int i = row0 + lo, j = a.fColInd[g];
			const T* B = a.fVal + 36 * (size_t)g;
			const T* Li = a.Linv + 36 * (size_t)i;
			const T* Lj = a.Linv + 36 * (size_t)j;
			T tmp[36];
			for (int c = 0; c < 6; c++)
				for (int r = 0; r < 6; r++) {
					T s = T(0);
					for (int k = 0; k <= r; k++) s += Li[k * 6 + r] * B[c * 6 + k];
					tmp[c * 6 + r] = s;
				}
			for (int c = 0; c < 6; c++)
				for (int r = 0; r < 6; r++) {
					T s = T(0);
					for (int k = 0; k <= c; k++) s += tmp[k * 6 + r] * __ldcg(Lj + k * 6 + c);
					out[c * 6 + r] = s;
				}
		};
#pragma unroll
		for (int u = 0; u < PCG3_BPT; u++) {
			const int n = u * PCG3_BLOCK + tid;
			if (n < nblkCta) {
				T out[36];
				transform(n, out);
#pragma unroll
				for (int e = 0; e < 36; e++) breg[u][e] = out[e];
				myLoc[u] = a.fLocal[blk0 + n];
			}
		}
		for (int n = PCG3_REGBLK + tid; n < nblkCta; n += PCG3_BLOCK) {
			T out[36];
			transform(n, out);
			const int m = n - PCG3_REGBLK;
			if (m < ncached) {
				for (int e = 0; e < 36; e++) s_blk[(size_t)e * a.capBlocks + m] = out[e];
				s_loc[m] = a.fLocal[blk0 + n];
			} else {
				for (int e = 0; e < 36; e++) a.fHat[36 * (size_t)(blk0 + n) + e] = out[e];
			}
		}
	}
	for (int wi = tid; wi < nneed * 6; wi += PCG3_BLOCK) {
		const int c = wi / 6, comp = wi - 6 * c;
		s_r[wi] = __ldcg(a.R0 + 6 * (size_t)s_need[c] + comp);
		s_s[wi] = T(0);
	}
	__syncthreads();

	int status = 1, it = 0;
	double gamma = 0, gamma0 = 0, alpha = 0, beta = 0;
#ifdef CUBA_PCG_TIMING
	long long tacc[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
#endif
	const int nw2 = (2 * G + 31) >> 5;       // warps' worth of polled partial products (CTA 0)
	// row sums: tpp threads per (row, component) pair, a power of two
	int tpp = 1;
	while (tpp < 8 && nrows * 6 * tpp * 2 <= PCG3_BLOCK) tpp *= 2;
	if (nbad > 0) status = 2;
	else {
		// pass k = -1: w0 = A^ r0 and the first inner products; pass k >= 0: CG iteration k.
		// Values published at the end of pass k-1 carry the tag k+1 and live in parity (k+1)&1.
		for (int k = -1;; k++) {
			if (k >= 0) {
				const unsigned int tag = (unsigned int)(k + 1);
				const int par = (k + 1) & 1;
				PCG_T(t0);
				// ---- one polling round per thread: up to PCG3_WPT w items and up to two partial / total words, loads in flight together.
				//      Every CTA sums everybody's partial products itself, in the same fixed order (one L2 hop).  With all G CTAs
				//      polling the same 2G slots the hot L2 lines cost ~2.5 us per round on B200 (tools/microbench), so the board
				//      is published in PCG3_REPL replicas and CTA c reads replica c % PCG3_REPL.
				// slots are kept as 32-bit offsets (in 16-byte words) to keep the two register-resident blocks out of local memory
				double wv[PCG3_WPT], pv[2];
				unsigned int woff[PCG3_WPT], poff[2];
				unsigned int pend = 0;                             // bit i: w item i pending; bits 8,9: partial words pending
#pragma unroll
				for (int u = 0; u < PCG3_WPT; u++) {
					const int wi = u * PCG3_BLOCK + tid;
					wv[u] = 0; woff[u] = 0;
					if (wi < nneed * 6) {
						const int c = wi / 6, comp = wi - 6 * c;
						woff[u] = (unsigned int)(par * (int)n6 + 6 * s_need[c] + comp);
						pend |= 1u << u;
					}
				}
#pragma unroll
				for (int u = 0; u < 2; u++) {
					const int pi = u * PCG3_BLOCK + tid;
					pv[u] = 0; poff[u] = 0;
					if (pi < 2 * G) {
						poff[u] = (unsigned int)((par * PCG3_REPL + (cta % PCG3_REPL)) * 2 * G + pi);
						pend |= 0x100u << u;
					}
				}
				bool ok = true;
				// the w entries are published before the partial products (which need the producer's block reduction), so
				// they are polled first; this also halves the registers live in either polling loop
				for (unsigned int spin = 0; pend & 0xffu; spin++) {
#pragma unroll
					for (int u = 0; u < PCG3_WPT; u++) if ((pend >> u) & 1u) { if (ll_try_load(aa.wFlag + 2 * (size_t)woff[u], tag, wv[u])) pend &= ~(1u << u); }
					if ((spin & 1023u) == 1023u) {
						if (*(volatile int*)aa.abortFlag) { ok = false; break; }
						if (spin >= PCG3_SPIN_LIMIT) { atomicExch(aa.abortFlag, 1); ok = false; break; }
					}
				}
				for (unsigned int spin = 0; ok && (pend & 0x300u); spin++) {
#pragma unroll
					for (int u = 0; u < 2; u++) if ((pend >> (8 + u)) & 1u) { if (ll_try_load(aa.pFlag + 2 * (size_t)poff[u], tag, pv[u])) pend &= ~(0x100u << u); }
					if ((spin & 1023u) == 1023u) {
						if (*(volatile int*)aa.abortFlag) { ok = false; break; }
						if (spin >= PCG3_SPIN_LIMIT) { atomicExch(aa.abortFlag, 1); ok = false; break; }
					}
				}
				PCG_T(t1);
				if (!ok) s_abort = 1;
				PCG_T(t2);
				PCG_ACC(0, t0, t1); PCG_ACC(1, t1, t2);
				double gnew = 0, delta = 0;
				// ---- fixed-order sum of everybody's partial products (slot = u*PCG3_BLOCK + tid; even slots gamma', odd delta):
				//      parity-preserving butterfly inside each warp straight from the polled registers, then 16 warp partials ----
#pragma unroll
				for (int u = 0; u < 2; u++) {
					double v = pv[u];
					v += __shfl_xor_sync(0xffffffffu, v, 2);
					v += __shfl_xor_sync(0xffffffffu, v, 4);
					v += __shfl_xor_sync(0xffffffffu, v, 8);
					v += __shfl_xor_sync(0xffffffffu, v, 16);
					if (lane < 2) s_w2[u * (PCG3_BLOCK / 32) + wid][lane] = v;
				}
				__syncthreads();
				for (int w = 0; w < nw2; w++) { gnew += s_w2[w][0]; delta += s_w2[w][1]; }
				if (s_abort) { status = 3; break; }
				if (!(gnew == gnew) || !(delta == delta)) { status = 2; break; }
				if (k == 0) {
					gamma0 = gamma = gnew;
					if (gamma0 <= 0) { status = 0; break; }
					if (!(delta > 0)) { status = 2; break; }
					alpha = gamma / delta; beta = 0;
				} else {
					it = k;
					if (gnew <= a.tol2 * gamma0) { gamma = gnew; status = 0; break; }
					// beta = g'/g and alpha' = g' / (delta - beta g'/alpha) with the two divisions independent of each other
					beta = gnew / gamma;
					const double ga = gamma * alpha;
					const double den = delta * ga - gnew * gnew;      // = ga * (delta - beta g'/alpha)
					if (!(den > 0) || !(ga > 0)) { gamma = gnew; status = 2; break; }
					alpha = gnew * ga / den;
					gamma = gnew;
				}
				if (k >= a.maxIters) { status = 1; break; }
				PCG_T(t3);
				PCG_ACC(2, t2, t3);
				// ---- advance s, r (all needed columns) and p, y (own rows) in shared memory ----
				for (int wi = tid, u = 0; wi < nneed * 6; wi += PCG3_BLOCK, u++) {
					double wvv = 0;
					if (u < PCG3_WPT) {
#pragma unroll
						for (int q = 0; q < PCG3_WPT; q++) if (q == u) wvv = wv[q];
					} else {
						const int c = wi / 6, comp = wi - 6 * c;
						if (!ll_wait(aa.wFlag + 2 * ((size_t)par * n6 + 6 * (size_t)s_need[c] + comp), tag, wvv, aa.abortFlag)) { s_abort = 1; wvv = 0; }
					}
					const T rold = s_r[wi];
					const T snew = (T)wvv + (T)beta * s_s[wi];
					s_s[wi] = snew;
					s_r[wi] = rold - (T)alpha * snew;
					const int own = s_own[wi / 6];
					if (own >= 0) {
						const int o = own * 6 + (wi % 6);
						const T p = rold + (T)beta * s_p[o];
						s_p[o] = p;
						s_y[o] += (T)alpha * p;
					}
				}
				__syncthreads();
				if (s_abort) { status = 3; break; }
				PCG_T(t4);
				PCG_ACC(3, t3, t4);
			}
			PCG_T(t5);
			// ---- w_{k+1} = A^ r_{k+1} for the own rows: block products from registers, then per-row sums ----
			const unsigned int otag = (unsigned int)(k + 2);
			const int opar = (k + 2) & 1;
			T wacc = T(0);                                       // (row, component) partial of this thread's slice
			for (int cs = 0; cs < nblkCta; cs += PCG3_CHUNK) {
				if (cs > 0) __syncthreads();                     // the previous chunk's row sums are done
#pragma unroll
				for (int u = 0; u < PCG3_BPT; u++) {
					const int n = cs + u * PCG3_BLOCK + tid;
					T y[6] = { T(0), T(0), T(0), T(0), T(0), T(0) };
					if (cs == 0) {
						if (myLoc[u] >= 0) {
							const T* rj = s_r + 6 * (size_t)myLoc[u];
#pragma unroll
							for (int c = 0; c < 6; c++) {
								const T rc = rj[c];
#pragma unroll
								for (int r = 0; r < 6; r++) y[r] += breg[u][c * 6 + r] * rc;
							}
						}
					} else if (n < nblkCta) {
						const int m = n - PCG3_REGBLK;
						const bool cached = m < ncached;
						const int loc = cached ? s_loc[m] : a.fLocal[blk0 + n];
						if (loc >= 0) {
							const T* rj = s_r + 6 * (size_t)loc;
							if (cached) {
								const T* B = s_blk + m;
								const size_t st = (size_t)a.capBlocks;
#pragma unroll
								for (int c = 0; c < 6; c++) {
									const T rc = rj[c];
#pragma unroll
									for (int r = 0; r < 6; r++) y[r] += B[(c * 6 + r) * st] * rc;
								}
							} else {
								const T* B = a.fHat + 36 * (size_t)(blk0 + n);
#pragma unroll
								for (int c = 0; c < 6; c++) {
									const T rc = rj[c];
#pragma unroll
									for (int r = 0; r < 6; r++) y[r] += B[c * 6 + r] * rc;
								}
							}
						}
					}
					// component-major staging: the row sums read consecutive blocks of one component
#pragma unroll
					for (int r = 0; r < 6; r++) s_c[r * PCG3_CHUNK + u * PCG3_BLOCK + tid] = y[r];
				}
				__syncthreads();
				if (tid < nrows * 6 * tpp) {                     // nrows*6 <= PCG3_BLOCK (checked on the host)
					const int pair = tid / tpp, sub = tid - pair * tpp;
					const int li = pair / 6, comp = pair - 6 * li;
					int n0 = s_rowPtr[li], n1 = s_rowPtr[li + 1];
					n0 = (n0 > cs ? n0 : cs) - cs;
					n1 = (n1 < cs + PCG3_CHUNK ? n1 : cs + PCG3_CHUNK) - cs;
					T s0 = T(0), s1 = T(0);
					const T* col = s_c + comp * PCG3_CHUNK;
					int q = n0 + sub;
					for (; q + tpp < n1; q += 2 * tpp) { s0 += col[q]; s1 += col[q + tpp]; }
					if (q < n1) s0 += col[q];
					wacc += s0 + s1;
				}
			}
			// combine the tpp slices of each (row, component) pair: lanes of one pair are adjacent
			for (int o = 1; o < tpp; o <<= 1) wacc += __shfl_xor_sync(0xffffffffu, wacc, o);
			PCG_T(t6);
			double pg = 0, pd = 0;
			if (tid < nrows * 6 * tpp && (tid % tpp) == 0) {
				const int pair = tid / tpp;
				const int li = pair / 6, comp = pair - 6 * li;
				const T ri = s_r[6 * (size_t)s_diag[li] + comp];
				const T wv1 = wacc + ri;                            // A^_ii = I
				ll_store(aa.wFlag + 2 * ((size_t)opar * n6 + 6 * (size_t)(row0 + li) + comp), (double)wv1, otag);
				pg = (double)ri * (double)ri;
				pd = (double)wv1 * (double)ri;
			}
			pg = warp_sum(pg); pd = warp_sum(pd);
			if (lane == 0) { s_red[wid][0] = pg; s_red[wid][1] = pd; }
			__syncthreads();
			PCG_T(t7);
			if (tid < 2 * PCG3_REPL) {                           // thread (replica, word): every replica gets both words
				const int rep = tid >> 1, word = tid & 1;
				double v = 0;
				for (int w = 0; w < PCG3_BLOCK / 32; w++) v += s_red[w][word];
				ll_store(aa.pFlag + 2 * (((size_t)opar * PCG3_REPL + rep) * 2 * G + 2 * (size_t)cta + word), v, otag);
			}
			PCG_T(t8);
			PCG_ACC(4, t5, t6); PCG_ACC(5, t6, t7); PCG_ACC(6, t7, t8);
			// s_red / s_c are rewritten only after the next pass's __syncthreads
		}
	}
	// ---- x = L^-T y for the own rows ----
	__syncthreads();
	for (int wi = tid; wi < nrows * 6; wi += PCG3_BLOCK) {
		const int li = wi / 6, r = wi % 6;
		const T* Li = a.Linv + 36 * (size_t)(row0 + li);
		T s = T(0);
		for (int c = r; c < 6; c++) s += Li[r * 6 + c] * s_y[6 * li + c];   // (L^-T)(r,c) = Li(c,r)
		a.x[6 * (size_t)(row0 + li) + r] = s;
	}
#ifdef CUBA_PCG_TIMING
	if (tid == 0 && aa.timing) { for (int i = 0; i < 7; i++) aa.timing[(size_t)cta * 8 + i] = tacc[i]; aa.timing[(size_t)cta * 8 + 7] = it; }
#endif
	if (cta == 0 && tid == 0) { a.status->iters = it; a.status->status = status; a.status->rz0 = gamma0; a.status->rz = gamma; }
}

}  // namespace cuba_b200
