// cuba_schur2.cuh -- tile-local Schur complement (second generation).
//
// k_schur gathers both 144-byte Hpl blocks of every block product from L2 (0.8 GB per launch on a
// ba_kitti_00-sized graph: each block is fetched ~8 times).  Here the products are regrouped by
// (landmark tile, destination block): a CTA loads the contiguous Hpl range of its tile into shared memory ONCE,
// inverts the tile's Hll blocks, and one thread per (tile, destination) segment accumulates that segment's
// products from shared memory into a partial 6x6 block (+ the bsc contribution on diagonal destinations).
// A second kernel sums the partials of every destination in a fixed order (tiles ascending) and applies the
// same epilogue as k_schur.  No atomics, deterministic.  Traffic: Hpl once + partial blocks twice.
#pragma once

#include "cuba_kernels.cuh"

namespace cuba_b200 {
namespace schur2 {

constexpr int TL = 128;        // threads per tile CTA == max Hpl blocks of a tile chunk staged at once
constexpr int PW = 42;         // doubles per partial: 36 (block) + 6 (bsc part, zero off the diagonal)

// key = (tile << 32) | destination block, for every real product of the destination-sorted list
__global__ void k_keys(const int* __restrict__ prodPtr, int nblk, const int* __restrict__ prodI, int N, const TileInfo* __restrict__ info, int ntiles,
	unsigned long long* key, int* val)
{
	const int n = blockIdx.x * blockDim.x + threadIdx.x;
	if (n >= N) return;
	val[n] = n;
	const int i = prodI[n];
	if (i < 0) { key[n] = ~0ull; return; }
	int lo = 0, hi = ntiles - 1;                 // first tile with h1 > i
	while (lo < hi) { const int mid = (lo + hi) >> 1; if (info[mid].h1 > i) hi = mid; else lo = mid + 1; }
	const int tile = lo;
	lo = 0; hi = nblk - 1;                       // destination k: last k with prodPtr[k] <= n
	while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (prodPtr[mid] <= n) lo = mid; else hi = mid - 1; }
	key[n] = ((unsigned long long)(unsigned)tile << 32) | (unsigned)lo;
}

__global__ void k_heads(const unsigned long long* __restrict__ key, int N, int* head)
{
	const int n = blockIdx.x * blockDim.x + threadIdx.x;
	if (n >= N) return;
	head[n] = (key[n] != ~0ull && (n == 0 || key[n] != key[n - 1])) ? 1 : 0;
}

struct Counts { int nseg; int nvalid; };

__global__ void k_counts(const unsigned long long* __restrict__ key, const int* __restrict__ head, const int* __restrict__ segId, int N, Counts* out)
{
	if (blockIdx.x != 0 || threadIdx.x != 0) return;
	out->nseg = N > 0 ? segId[N - 1] + head[N - 1] : 0;
	int lo = 0, hi = N;                          // first sentinel
	while (lo < hi) { const int mid = (lo + hi) >> 1; if (key[mid] == ~0ull) hi = mid; else lo = mid + 1; }
	out->nvalid = lo;
}

__global__ void k_segments(const unsigned long long* __restrict__ key, const int* __restrict__ valSorted, const int* __restrict__ head,
	const int* __restrict__ segId, const int* __restrict__ prodI, const int* __restrict__ prodJ, int N, int nseg, int nvalid,
	int* segStart, int* segTile, int* segDest, int* p2i, int* p2j, unsigned long long* key3, int* val3)
{
	const int n = blockIdx.x * blockDim.x + threadIdx.x;
	if (n > N) return;
	if (n == N) { segStart[nseg] = nvalid; return; }
	if (key[n] == ~0ull) return;
	if (head[n]) {
		const int s = segId[n];
		const int tile = (int)(key[n] >> 32), dest = (int)(key[n] & 0xffffffffu);
		segStart[s] = n; segTile[s] = tile; segDest[s] = dest;
		key3[s] = ((unsigned long long)(unsigned)dest << 32) | (unsigned)tile;
		val3[s] = s;
	}
	const int src = valSorted[n];
	p2i[n] = prodI[src]; p2j[n] = prodJ[src];
}

// ptr[i] = first segment (sorted by the given 32-bit field, ascending) with field >= i
__global__ void k_ptr_from_field(const int* __restrict__ field, int n, int m, int* ptr)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i > m) return;
	int lo = 0, hi = n;
	while (lo < hi) { const int mid = (lo + hi) >> 1; if (field[mid] < i) lo = mid + 1; else hi = mid; }
	ptr[i] = lo;
}

__global__ void k_rank(const unsigned long long* __restrict__ key3Sorted, const int* __restrict__ val3Sorted, int nseg, int* segRank, int* rankDest)
{
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= nseg) return;
	segRank[val3Sorted[r]] = r;
	rankDest[r] = (int)(key3Sorted[r] >> 32);
}

template <typename T>
struct TileArgs {
	const T* Hpl; const T* Hll; const T* bl;
	const TileInfo* info;
	const int* hplLm;
	const int* tileSegPtr; const int* segStart; const int* segDest; const int* segRank; const int* p2i; const int* p2j;
	const int* blkRow; const int* blkCol;
	int numL;
	T lambda;
	T* invHll; T* partial;
};

template <typename T>
__global__ void __launch_bounds__(TL, 3) k_schur_tiles(const TileArgs<T> a)
{
	__shared__ __align__(16) T s_A[TL * 18];      // the tile's Hpl blocks
	__shared__ T s_inv[TL * 6];                   // inverse of (Hll + lambda I), 6 unique entries per landmark
	__shared__ T s_bl[TL * 3];
	__shared__ int s_lm[TL];                      // local landmark of each block
	const int tid = threadIdx.x, t = blockIdx.x;
	const TileInfo ti = a.info[t];
	const int nb = ti.h1 - ti.h0;
	int nl = ti.l1 - ti.l0;
	if (ti.l0 + nl > a.numL) nl = a.numL - ti.l0;       // the pseudo-landmark of the fixed ones has no Hll
	if (nb <= 0) {
		// still publish the inverses of the tile's landmarks (used by the back-substitution)
		for (int j = tid; j < nl; j += TL) {
			const T* H = a.Hll + 9 * (size_t)(ti.l0 + j);
			T B[6];
			sym3_inverse<T>(H[0] + a.lambda, H[3], H[6], H[4] + a.lambda, H[7], H[8] + a.lambda, B);
			T* o = a.invHll + 9 * (size_t)(ti.l0 + j);
			o[0] = B[0]; o[1] = B[1]; o[2] = B[2]; o[3] = B[1]; o[4] = B[3]; o[5] = B[4]; o[6] = B[2]; o[7] = B[4]; o[8] = B[5];
		}
		return;
	}
	// a tile stages at most TL blocks (tiles are cut so that this holds except for giant landmarks, see below)
	const int nbs = nb < TL ? nb : TL;
	for (int i = tid; i < nbs * 9; i += TL) {
		T x, y;
		ld2(a.Hpl + 18 * (size_t)ti.h0 + 2 * i, x, y);
		s_A[2 * i] = x; s_A[2 * i + 1] = y;
	}
	for (int i = tid; i < nbs; i += TL) s_lm[i] = a.hplLm[ti.h0 + i] - ti.l0;
	for (int j = tid; j < nl; j += TL) {
		const T* H = a.Hll + 9 * (size_t)(ti.l0 + j);
		T B[6];
		sym3_inverse<T>(H[0] + a.lambda, H[3], H[6], H[4] + a.lambda, H[7], H[8] + a.lambda, B);
		T* o = a.invHll + 9 * (size_t)(ti.l0 + j);
		o[0] = B[0]; o[1] = B[1]; o[2] = B[2]; o[3] = B[1]; o[4] = B[3]; o[5] = B[4]; o[6] = B[2]; o[7] = B[4]; o[8] = B[5];
		if (j < TL) {
#pragma unroll
			for (int k = 0; k < 6; k++) s_inv[6 * j + k] = B[k];
			s_bl[3 * j] = a.bl[3 * (size_t)(ti.l0 + j)]; s_bl[3 * j + 1] = a.bl[3 * (size_t)(ti.l0 + j) + 1]; s_bl[3 * j + 2] = a.bl[3 * (size_t)(ti.l0 + j) + 2];
		}
	}
	__syncthreads();
	const int s0 = a.tileSegPtr[t], s1 = a.tileSegPtr[t + 1];
	for (int s = s0 + tid; s < s1; s += TL) {
		const int k = a.segDest[s];
		const bool diag = a.blkRow[k] == a.blkCol[k];
		T C[36], v[6];
#pragma unroll
		for (int i = 0; i < 36; i++) C[i] = T(0);
#pragma unroll
		for (int i = 0; i < 6; i++) v[i] = T(0);
		const int n1 = a.segStart[s + 1];
		for (int n = a.segStart[s]; n < n1; n++) {
			const int bi = a.p2i[n] - ti.h0, bj = a.p2j[n] - ti.h0;
			T Aj[18], inv[6], b3[3];
			const T* Ai;
			T Aig[18];
			int lml;
			if (bi < TL && bj < TL) {            // staged (always, unless one landmark has more than TL observations)
				Ai = s_A + 18 * bi;
#pragma unroll
				for (int x = 0; x < 18; x++) Aj[x] = s_A[18 * bj + x];
				lml = s_lm[bi];
			} else {
				const T* gi = a.Hpl + 18 * (size_t)(ti.h0 + bi);
				const T* gj = a.Hpl + 18 * (size_t)(ti.h0 + bj);
#pragma unroll
				for (int x = 0; x < 18; x++) { Aig[x] = gi[x]; Aj[x] = gj[x]; }
				Ai = Aig;
				lml = a.hplLm[ti.h0 + bi] - ti.l0;
			}
			if (lml < TL) {
#pragma unroll
				for (int x = 0; x < 6; x++) inv[x] = s_inv[6 * lml + x];
				b3[0] = s_bl[3 * lml]; b3[1] = s_bl[3 * lml + 1]; b3[2] = s_bl[3 * lml + 2];
			} else {
				const T* iv = a.invHll + 9 * (size_t)(ti.l0 + lml);
				inv[0] = iv[0]; inv[1] = iv[3]; inv[2] = iv[6]; inv[3] = iv[4]; inv[4] = iv[7]; inv[5] = iv[8];
				const T* bb = a.bl + 3 * (size_t)(ti.l0 + lml);
				b3[0] = bb[0]; b3[1] = bb[1]; b3[2] = bb[2];
			}
#pragma unroll
			for (int r = 0; r < 6; r++) {
				const T w0 = Ai[r] * inv[0] + Ai[6 + r] * inv[1] + Ai[12 + r] * inv[2];
				const T w1 = Ai[r] * inv[1] + Ai[6 + r] * inv[3] + Ai[12 + r] * inv[4];
				const T w2 = Ai[r] * inv[2] + Ai[6 + r] * inv[4] + Ai[12 + r] * inv[5];
#pragma unroll
				for (int c = 0; c < 6; c++) C[c * 6 + r] += w0 * Aj[c] + w1 * Aj[6 + c] + w2 * Aj[12 + c];
				if (diag) v[r] += w0 * b3[0] + w1 * b3[1] + w2 * b3[2];
			}
		}
		T* out = a.partial + (size_t)PW * a.segRank[s];
#pragma unroll
		for (int i = 0; i < 36; i += 2) st2(out + i, C[i], C[i + 1]);
#pragma unroll
		for (int i = 0; i < 6; i += 2) st2(out + 36 + i, v[i], v[i + 1]);
	}
}

template <typename T>
struct ReduceArgs {
	const T* partial; const int* destSegPtr;
	const T* Hpp; const T* bp;
	const int* blkRow; const int* blkCol; const int* u2f; const int* u2fT;
	int nblk; T lambda; int addDiag;
	T* fVal; T* bsc;
};

// one warp per destination block: fixed-order sum of its partials (tiles ascending), then k_schur's epilogue
template <typename T>
__global__ void __launch_bounds__(128) k_schur_reduce(const ReduceArgs<T> a)
{
	const int lane = threadIdx.x & 31;
	const int k = blockIdx.x * 4 + (threadIdx.x >> 5);
	if (k >= a.nblk) return;
	const int ra = a.blkRow[k], cb = a.blkCol[k];
	const bool diag = ra == cb;
	const int r0 = a.destSegPtr[k], r1 = a.destSegPtr[k + 1];
	for (int e = lane; e < PW; e += 32) {
		T s = T(0);
		for (int r = r0; r < r1; r++) s += a.partial[(size_t)PW * r + e];
		if (e < 36) {
			const int c = e / 6, rr = e - 6 * c;
			T val = -s;
			if (diag && a.addDiag) val += a.Hpp[36 * (size_t)ra + e] + (rr == c ? a.lambda : T(0));
			a.fVal[36 * (size_t)a.u2f[k] + e] = val;
			if (!diag) a.fVal[36 * (size_t)a.u2fT[k] + rr * 6 + c] = val;
		} else if (diag) {
			const int rr = e - 36;
			a.bsc[6 * (size_t)ra + rr] = (a.addDiag ? a.bp[6 * (size_t)ra + rr] : T(0)) - s;
		}
	}
}

}  // namespace schur2
}  // namespace cuba_b200
