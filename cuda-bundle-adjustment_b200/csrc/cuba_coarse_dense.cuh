// cuba_coarse_dense.cuh -- explicit inverse of the coarse matrix Ac = Z^T S Z of the two-level PCG (cuba_pcg5.cuh) by a dense,
// tile-parallel Cholesky on the whole chip.
//
// The cluster kernels of cuba_pcg4.cuh keep the packed 6x6-block triangle in the shared memory of 8 / 16 CTAs and walk it one
// block column at a time: 148 aggregates = 148 x (diagonal factorisation by one warp + two cluster barriers) = 3.0 ms, plus
// 0.6 ms for the triangular inverse (profiles/r02_pcg_compare_v2.log).  The matrix is tiny (888 x 888 fp64 = 6.3 MB, L2
// resident) and the arithmetic is 0.7 GFLOP in total: the time is all dependent steps.  This kernel cuts the number of
// dependent steps from 148 + 148 to 56 + 56 by working on plain 16 x 16 scalar tiles of the dense matrix, and spreads every
// step over all SMs (one persistent cooperative kernel, the grid barrier of cuba_pcg2.cuh between steps):
//   phase 0  packed blocks -> dense lower triangle M (padded to a multiple of 16 with a unit diagonal);
//   phase 1  right-looking Cholesky: per step k every CTA factors the 16 x 16 diagonal tile itself (one warp, 1-2 us, no
//            barrier), then one WARP per trailing tile (i, j) forms the two panel tiles it needs on the fly
//            (P_i = M_ik L_kk^-T: redundant across tiles, but it removes the panel barrier) and updates M_ij -= P_i P_j^T;
//            the warp of tile (i, i) also stores P_i as the factor's tile L_ik;  ONE grid barrier per step;
//   phase 2  W = L^-1, one CTA per tile column (columns are independent, a column is sequential in i);
//   phase 3  Ac^-1 = W^T W, one warp per tile, written in fp32 to both triangles (the PCG applies it in fp32, see cuba_pcg4.cuh).
// Fixed summation order everywhere: bit-reproducible.  On a non-positive pivot the inverse is zeroed (block-Jacobi alone).
#pragma once

#include "cuba_pcg4.cuh"

namespace cuba_b200 {
namespace cdense {

constexpr int NB = 16;                 // tile edge
constexpr int WARPS = 8;
constexpr int TS = NB + 1;             // padded row stride of a staged tile

struct Args {
	const double* AcP;                 // packed lower block triangle, block (ib >= jb) at (ib (ib+1)/2 + jb) * 36, column-major 6x6
	int A;                             // aggregates: n = 6 A
	double* M;                         // [np][np] column-major work matrix (lower triangle), zeroed by the host before the launch
	double* Lm;                        // [np][np] factor L (lower)
	double* Dinv;                      // [np/16][256] inverses of the diagonal tiles of L, row-major 16x16
	double* W;                         // [np][np] L^-1 (lower)
	float* AcInv;                      // [n][n] out
	int* info;                         // 0 ok, 1 not positive definite
	GridBar* bar;
};

__device__ __forceinline__ double ldcg(const double* p) { return __ldcg(p); }

// lower Cholesky factor of the 16x16 tile in sD (row-major, stride TS) in place, its inverse (lower) into sLi; one warp.
// Returns false on a non-positive pivot.
__device__ __forceinline__ bool chol16(double* sD, double* sLi, int lane)
{
	bool ok = true;
	double rdiag = 0.0;                 // lane j ends up with 1 / L(j,j)
	for (int j = 0; j < NB; j++) {
		const double d = sD[j * TS + j];
		if (!(d > 0)) { ok = false; break; }
		const double sq = sqrt(d), rs = 1.0 / sq;            // one division per column, the scaling multiplies
		__syncwarp();
		if (lane == j) { sD[j * TS + j] = sq; rdiag = rs; }
		else if (lane > j && lane < NB) sD[lane * TS + j] = sD[lane * TS + j] * rs;
		__syncwarp();
		if (lane > j && lane < NB) {
			const double lrj = sD[lane * TS + j];
			for (int c = j + 1; c <= lane; c++) sD[lane * TS + c] -= lrj * sD[c * TS + j];
		}
		__syncwarp();
	}
	if (!ok) return false;
	// column q of the inverse by forward substitution (lane q); the reciprocals of the diagonal come by shuffle
	double col[NB];
#pragma unroll
	for (int i = 0; i < NB; i++) col[i] = 0.0;
	const int q = lane;
#pragma unroll
	for (int i = 0; i < NB; i++) {
		const double ri = __shfl_sync(0xffffffffu, rdiag, i);
		if (q < NB && i >= q) {
			if (i == q) col[i] = ri;
			else {
				double s = 0;
#pragma unroll
				for (int k = 0; k < NB; k++) if (k >= q && k < i) s += sD[i * TS + k] * col[k];
				col[i] = -s * ri;
			}
		}
	}
	if (q < NB) {
#pragma unroll
		for (int i = 0; i < NB; i++) sLi[i * TS + q] = i >= q ? col[i] : 0.0;
	}
	__syncwarp();
	return true;
}

__global__ void __launch_bounds__(WARPS * 32, 1) k_coarse_dense(const Args a)
{
	__shared__ double sD[NB * TS], sLi[NB * TS];
	__shared__ double sP[WARPS][2][NB * TS];         // per warp: two staged tiles (phase 2: [w][1] = partial sums of warp w)
	__shared__ int s_fail;
	__shared__ unsigned int s_gen;
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const int G = gridDim.x, cta = blockIdx.x;
	const int n = 6 * a.A, nt = (n + NB - 1) / NB, np = nt * NB;
	const int nwarps = G * WARPS, gw = cta * WARPS + wid;
	if (tid == 0) { s_gen = ld_acquire_u32(&a.bar->gen); s_fail = 0; }
	__syncthreads();
	unsigned int gen = s_gen;
	// lane -> entries of a tile: row r = lane / 2, columns c0 .. c0 + 7
	const int r = lane >> 1, c0 = (lane & 1) * 8;

	// ---- phase 0: packed 6x6 blocks -> dense lower triangle; unit diagonal in the padding ----
	{
		const int nblkP = a.A * (a.A + 1) / 2;
		for (long long e = (long long)cta * blockDim.x + tid; e < (long long)nblkP * 36; e += (long long)G * blockDim.x) {
			const int b = (int)(e / 36), rc = (int)(e - 36LL * b), c = rc / 6, rr = rc - 6 * c;
			int ib = (int)((sqrt(8.0 * b + 1.0) - 1.0) * 0.5);
			while ((ib + 1) * (ib + 2) / 2 <= b) ib++;
			while (ib * (ib + 1) / 2 > b) ib--;
			const int jb = b - ib * (ib + 1) / 2;
			const int row = 6 * ib + rr, col = 6 * jb + c;
			if (row >= col) __stcg(a.M + (size_t)col * np + row, a.AcP[e]);
		}
		for (int i = n + cta * blockDim.x + tid; i < np; i += G * blockDim.x) __stcg(a.M + (size_t)i * np + i, 1.0);
	}
	grid_barrier(a.bar, G, gen);

	// ---- phase 1: Cholesky ----
	for (int k = 0; k < nt; k++) {
		if (wid == 0) {
			for (int e = lane; e < NB * NB; e += 32) { const int rr = e % NB, cc = e / NB; sD[rr * TS + cc] = rr >= cc ? ldcg(a.M + (size_t)(k * NB + cc) * np + k * NB + rr) : 0.0; }
			__syncwarp();
			const bool ok = chol16(sD, sLi, lane);
			if (!ok && lane == 0) s_fail = 1;
			if (ok && cta == 0) {
				for (int e = lane; e < NB * NB; e += 32) {
					const int rr = e % NB, cc = e / NB;
					__stcg(a.Lm + (size_t)(k * NB + cc) * np + k * NB + rr, rr >= cc ? sD[rr * TS + cc] : 0.0);
					__stcg(a.Dinv + (size_t)k * NB * NB + rr * NB + cc, sLi[rr * TS + cc]);
				}
			}
		}
		__syncthreads();
		if (s_fail) break;                                 // every CTA sees the same pivots
		const int m = nt - k - 1;                          // trailing tile rows
		const int ntile = m * (m + 1) / 2;
		double* Pi = sP[wid][0];
		double* Pj = sP[wid][1];
		for (int t = gw; t < ntile; t += nwarps) {
			int li = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
			while ((li + 1) * (li + 2) / 2 <= t) li++;
			while (li * (li + 1) / 2 > t) li--;
			const int lj = t - li * (li + 1) / 2;
			const int i = k + 1 + li, j = k + 1 + lj;
			// P_i = M_ik L_kk^-T : P(r, c) = sum_{m <= c} M_ik(r, m) Linv(c, m); this lane: row r, columns c0..c0+7
			{
				double arow[NB];
#pragma unroll
				for (int mm = 0; mm < NB; mm++) arow[mm] = ldcg(a.M + (size_t)(k * NB + mm) * np + i * NB + r);
#pragma unroll
				for (int c = 0; c < 8; c++) {
					const int cc = c0 + c;
					double s = 0;
#pragma unroll
					for (int mm = 0; mm < NB; mm++) if (mm <= cc) s += arow[mm] * sLi[cc * TS + mm];
					Pi[r * TS + cc] = s;
				}
				if (j != i) {
#pragma unroll
					for (int mm = 0; mm < NB; mm++) arow[mm] = ldcg(a.M + (size_t)(k * NB + mm) * np + j * NB + r);
#pragma unroll
					for (int c = 0; c < 8; c++) {
						const int cc = c0 + c;
						double s = 0;
#pragma unroll
						for (int mm = 0; mm < NB; mm++) if (mm <= cc) s += arow[mm] * sLi[cc * TS + mm];
						Pj[r * TS + cc] = s;
					}
				}
			}
			__syncwarp();
			const double* Q = j != i ? Pj : Pi;
			// M_ij -= P_i P_j^T ; the diagonal tile's warp also publishes L_ik = P_i
#pragma unroll
			for (int c = 0; c < 8; c++) {
				const int cc = c0 + c;
				double s = 0;
#pragma unroll
				for (int mm = 0; mm < NB; mm++) s += Pi[r * TS + mm] * Q[cc * TS + mm];
				double* dst = a.M + (size_t)(j * NB + cc) * np + i * NB + r;
				__stcg(dst, ldcg(dst) - s);
				if (j == i) __stcg(a.Lm + (size_t)(k * NB + cc) * np + i * NB + r, Pi[r * TS + cc]);
			}
			__syncwarp();
		}
		grid_barrier(a.bar, G, gen);
	}
	if (s_fail) {
		for (long long e = (long long)cta * blockDim.x + tid; e < (long long)n * n; e += (long long)G * blockDim.x) a.AcInv[e] = 0.f;
		if (cta == 0 && tid == 0) *a.info = 1;
		return;
	}

	// ---- phase 2: W = L^-1, tile column jt per CTA: W(jt,jt) = Linv_jj, W(i,jt) = -Linv_ii sum_{k=jt}^{i-1} L(i,k) W(k,jt) ----
	for (int jt = cta; jt < nt; jt += G) {
		for (int e = tid; e < NB * NB; e += blockDim.x) { const int rr = e / NB, cc = e % NB; __stcg(a.W + (size_t)(jt * NB + cc) * np + jt * NB + rr, ldcg(a.Dinv + (size_t)jt * NB * NB + e)); }
		__syncthreads();
		for (int i = jt + 1; i < nt; i++) {
			// partial sums of this warp over its k's; this lane: row r, columns c0..c0+7
			double acc[8];
#pragma unroll
			for (int c = 0; c < 8; c++) acc[c] = 0.0;
			for (int k = jt + wid; k < i; k += WARPS) {
				double lrow[NB];
#pragma unroll
				for (int mm = 0; mm < NB; mm++) lrow[mm] = ldcg(a.Lm + (size_t)(k * NB + mm) * np + i * NB + r);     // L(i,k)(r, mm)
#pragma unroll
				for (int c = 0; c < 8; c++) {
					const double* wc = a.W + (size_t)(jt * NB + c0 + c) * np + k * NB;                           // W(k,jt)(:, c)
					double s = 0;
#pragma unroll
					for (int mm = 0; mm < NB; mm++) s += lrow[mm] * ldcg(wc + mm);
					acc[c] += s;
				}
			}
#pragma unroll
			for (int c = 0; c < 8; c++) sP[wid][1][r * NB + c0 + c] = acc[c];
			__syncthreads();
			if (wid == 0) {
				// S = sum of the partials (warp order), W(i,jt) = -Linv_ii S
				double* S = sP[0][0];
#pragma unroll
				for (int c = 0; c < 8; c++) {
					double s = 0;
#pragma unroll
					for (int w = 0; w < WARPS; w++) s += sP[w][1][r * NB + c0 + c];
					S[r * TS + c0 + c] = s;
				}
				__syncwarp();
				double lrow[NB];
#pragma unroll
				for (int mm = 0; mm < NB; mm++) lrow[mm] = ldcg(a.Dinv + (size_t)i * NB * NB + r * NB + mm);            // Linv_ii(r, mm), mm <= r
#pragma unroll
				for (int c = 0; c < 8; c++) {
					double s = 0;
#pragma unroll
					for (int mm = 0; mm < NB; mm++) if (mm <= r) s += lrow[mm] * S[mm * TS + c0 + c];
					__stcg(a.W + (size_t)(jt * NB + c0 + c) * np + i * NB + r, -s);
				}
			}
			__syncthreads();
		}
	}
	grid_barrier(a.bar, G, gen);

	// ---- phase 3: Ac^-1 = W^T W: tile (i >= j) = sum_{k >= i} W(k,i)^T W(k,j); one warp per tile ----
	{
		const int ntile = nt * (nt + 1) / 2;
		double* Wi = sP[wid][0];
		double* Wj = sP[wid][1];
		for (int t = gw; t < ntile; t += nwarps) {
			int i = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
			while ((i + 1) * (i + 2) / 2 <= t) i++;
			while (i * (i + 1) / 2 > t) i--;
			const int j = t - i * (i + 1) / 2;
			double acc[8];
#pragma unroll
			for (int c = 0; c < 8; c++) acc[c] = 0.0;
			for (int k = i; k < nt; k++) {
				// stage W(k,i) and W(k,j) (row-major in shared memory); lane loads row r, columns c0..c0+7 of each
#pragma unroll
				for (int c = 0; c < 8; c++) {
					Wi[r * TS + c0 + c] = ldcg(a.W + (size_t)(i * NB + c0 + c) * np + k * NB + r);
					Wj[r * TS + c0 + c] = ldcg(a.W + (size_t)(j * NB + c0 + c) * np + k * NB + r);
				}
				__syncwarp();
				// out(r, cc) += sum_m W(k,i)(m, r) W(k,j)(m, cc)
#pragma unroll
				for (int c = 0; c < 8; c++) {
					double s = 0;
#pragma unroll
					for (int mm = 0; mm < NB; mm++) s += Wi[mm * TS + r] * Wj[mm * TS + c0 + c];
					acc[c] += s;
				}
				__syncwarp();
			}
#pragma unroll
			for (int c = 0; c < 8; c++) {
				const int row = i * NB + r, col = j * NB + c0 + c;
				if (row < n && col < n) {
					const float v = (float)acc[c];
					a.AcInv[(size_t)row * n + col] = v;
					a.AcInv[(size_t)col * n + row] = v;
				}
			}
		}
	}
	if (cta == 0 && tid == 0) *a.info = 0;
}

}  // namespace cdense
}  // namespace cuba_b200
