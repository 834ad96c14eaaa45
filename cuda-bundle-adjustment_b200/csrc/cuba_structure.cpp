// cuba_structure.cpp -- see cuba_structure.h
#include "cuba_structure.h"

#include <algorithm>

namespace cuba_b200 {

bool build_structure(int Pall, int numP, int Lall, int numL, int E2, const int32_t* idx2, int E3, const int32_t* idx3,
	int rank, int world, int tileEdges, Structure& S, const char** err)
{
	S = Structure();
	if (Pall < 0 || Lall < 0 || numP < 0 || numL < 0 || numP > Pall || numL > Lall || E2 < 0 || E3 < 0 || world < 1 || rank < 0 ||
		rank >= world || tileEdges < 1) {
		*err = "build_structure: invalid sizes";
		return false;
	}
	const int E = E2 + E3;
	S.Pall = Pall; S.numP = numP; S.Lall = Lall; S.numL = numL; S.E2 = E2; S.E3 = E3; S.E = E;
	auto IP = [&](int u) { return u < E2 ? idx2[2 * (size_t)u] : idx3[2 * (size_t)(u - E2)]; };
	auto IL = [&](int u) { return u < E2 ? idx2[2 * (size_t)u + 1] : idx3[2 * (size_t)(u - E2) + 1]; };
	for (int u = 0; u < E; u++) {
		const int ip = IP(u), il = IL(u);
		if (ip < 0 || ip >= Pall || il < 0 || il >= Lall) { *err = "build_structure: edge index out of range"; return false; }
		if (ip >= numP && il >= numL) { *err = "build_structure: edge with both ends fixed"; return false; }
	}

	// 1. canonical order: stable counting sort by iP, then by iL  ->  sorted by (iL, iP, edge id)
	std::vector<int> cnt(Pall + 1, 0), byP(E);
	for (int u = 0; u < E; u++) cnt[IP(u) + 1]++;
	for (int p = 0; p < Pall; p++) cnt[p + 1] += cnt[p];
	for (int u = 0; u < E; u++) byP[cnt[IP(u)]++] = u;
	std::vector<int> lmPtrG(Lall + 1, 0);
	for (int u = 0; u < E; u++) lmPtrG[IL(u) + 1]++;
	for (int l = 0; l < Lall; l++) lmPtrG[l + 1] += lmPtrG[l];
	std::vector<int> orderG(E);
	{
		std::vector<int> fill(lmPtrG.begin(), lmPtrG.end() - 1);
		for (int k = 0; k < E; k++) { const int u = byP[k]; orderG[fill[IL(u)]++] = u; }
	}

	// 2. Hpl CSC (global): block index = rank of the edge among free-free edges in canonical order
	S.hplColPtr.assign(numL + 1, 0);
	S.edge2Hpl.assign(E, -1);
	std::vector<int> hplLmG;
	S.hplRowInd.reserve(E); hplLmG.reserve(E);
	for (int k = 0; k < E; k++) {
		const int u = orderG[k];
		const int ip = IP(u), il = IL(u);
		if (ip < numP && il < numL) {
			S.edge2Hpl[u] = (int)S.hplRowInd.size();
			S.hplRowInd.push_back(ip);
			hplLmG.push_back(il);
			S.hplColPtr[il + 1]++;
		}
	}
	for (int l = 0; l < numL; l++) S.hplColPtr[l + 1] += S.hplColPtr[l];
	S.nhpl = (int)S.hplRowInd.size();

	// 3. landmark shard of this rank, balanced by edge count, snapped to landmark boundaries
	auto boundary = [&](int r) {
		if (r <= 0) return 0;
		if (r >= world) return Lall;
		const long long target = (long long)E * r / world;
		return (int)(std::lower_bound(lmPtrG.begin(), lmPtrG.end(), (int)target) - lmPtrG.begin());
	};
	S.lmBeg = std::min(boundary(rank), Lall);
	S.lmEnd = std::min(std::max(boundary(rank + 1), S.lmBeg), Lall);
	const int kBeg = lmPtrG[S.lmBeg], kEnd = lmPtrG[S.lmEnd];
	S.eLocal = kEnd - kBeg;
	auto hplAt = [&](int l) { return l < numL ? S.hplColPtr[l] : S.nhpl; };
	S.hplBase = hplAt(S.lmBeg);
	S.nhplLocal = hplAt(S.lmEnd) - S.hplBase;
	S.order.resize(S.eLocal); S.e_ip.resize(S.eLocal); S.e_il.resize(S.eLocal); S.e_hpl.resize(S.eLocal);
	{
		int rankFF = 0;   // rank among the shard's free-free edges; edges without a block store -1-rank
		for (int k = kBeg; k < kEnd; k++) {
			const int u = orderG[k], e = k - kBeg;
			S.order[e] = u;
			S.e_ip[e] = IP(u) | (u >= E2 ? (int)0x80000000u : 0);
			S.e_il[e] = IL(u);
			if (S.edge2Hpl[u] >= 0) { S.e_hpl[e] = S.edge2Hpl[u] - S.hplBase; rankFF = S.e_hpl[e] + 1; }
			else S.e_hpl[e] = -1 - rankFF;
		}
	}
	S.lmPtr.resize(Lall + 1);
	for (int l = 0; l <= Lall; l++) S.lmPtr[l] = std::min(std::max(lmPtrG[l], kBeg), kEnd) - kBeg;
	S.hplLm.assign(hplLmG.begin() + S.hplBase, hplLmG.begin() + S.hplBase + S.nhplLocal);

	// 4. landmark tiles: whole landmarks, <= tileEdges edges and <= tileEdges landmarks per tile;
	//    a landmark with more edges than that gets a tile of its own
	{
		int l = S.lmBeg;
		while (l < S.lmEnd) {
			S.tileLm.push_back(l);
			int edges = 0, n = 0;
			while (l < S.lmEnd && n < tileEdges) {
				const int d = S.lmPtr[l + 1] - S.lmPtr[l];
				if (n > 0 && edges + d > tileEdges) break;
				edges += d; n++; l++;
				if (edges >= tileEdges) break;
			}
		}
		S.tileLm.push_back(S.lmEnd);
	}

	// 5. pose-major copy of the shard's edges with a free pose
	S.posePtr.assign(numP + 1, 0);
	for (int e = 0; e < S.eLocal; e++) { const int ip = S.e_ip[e] & 0x7fffffff; if (ip < numP) S.posePtr[ip + 1]++; }
	for (int p = 0; p < numP; p++) S.posePtr[p + 1] += S.posePtr[p];
	S.p_src.resize(S.posePtr[numP]); S.p_il.resize(S.posePtr[numP]);
	{
		std::vector<int> fill(S.posePtr.begin(), S.posePtr.end() - 1);
		for (int e = 0; e < S.eLocal; e++) {
			const int ip = S.e_ip[e] & 0x7fffffff;
			if (ip < numP) {
				const int pos = fill[ip]++;
				S.p_src[pos] = e;
				S.p_il[pos] = S.e_il[e] | (S.e_ip[e] & (int)0x80000000u);
			}
		}
	}

	// 6. Hsc upper pattern (global) + destination-sorted product list (shard's landmarks)
	std::vector<int> rowPtrH(numP + 1, 0), rowList(S.nhpl);
	for (int h = 0; h < S.nhpl; h++) rowPtrH[S.hplRowInd[h] + 1]++;
	for (int p = 0; p < numP; p++) rowPtrH[p + 1] += rowPtrH[p];
	{
		std::vector<int> fill(rowPtrH.begin(), rowPtrH.end() - 1);
		for (int h = 0; h < S.nhpl; h++) rowList[fill[S.hplRowInd[h]]++] = h;
	}
	std::vector<int> mark(numP, 0), cntLoc(numP, 0), touched;
	std::vector<int> prodCount;
	S.hscRowPtr.assign(numP + 1, 0);
	S.nmul = 0; S.nmulLocal = 0;
	for (int a = 0; a < numP; a++) {
		touched.clear();
		touched.push_back(a); mark[a] = 1;     // the diagonal block always exists
		for (int x = rowPtrH[a]; x < rowPtrH[a + 1]; x++) {
			const int i = rowList[x];
			const int l = hplLmG[i];
			const bool local = l >= S.lmBeg && l < S.lmEnd;
			const int cend = S.hplColPtr[l + 1];
			for (int j = i; j < cend; j++) {
				const int b = S.hplRowInd[j];
				if (!mark[b]) { mark[b] = 1; touched.push_back(b); }
				if (local) cntLoc[b]++;
			}
			S.nmul += cend - i;
		}
		std::sort(touched.begin(), touched.end());
		for (int b : touched) {
			S.blkRow.push_back(a); S.blkCol.push_back(b);
			prodCount.push_back(cntLoc[b]);
			S.nmulLocal += cntLoc[b];
			mark[b] = 0; cntLoc[b] = 0;
		}
		S.hscRowPtr[a + 1] = (int)S.blkRow.size();
	}
	S.nblk = (int)S.blkRow.size();
	S.hscColInd = S.blkCol;
	if (S.nmulLocal > 0x7fffffffLL) { *err = "build_structure: more than 2^31 block products in one shard"; return false; }
	S.prodPtr.assign(S.nblk + 1, 0);
	for (int k = 0; k < S.nblk; k++) S.prodPtr[k + 1] = S.prodPtr[k] + prodCount[k];
	S.prodI.resize((size_t)S.nmulLocal); S.prodJ.resize((size_t)S.nmulLocal);
	{
		std::vector<int> off(numP, 0);
		for (int a = 0; a < numP; a++) {
			for (int k = S.hscRowPtr[a]; k < S.hscRowPtr[a + 1]; k++) off[S.blkCol[k]] = S.prodPtr[k];
			for (int x = rowPtrH[a]; x < rowPtrH[a + 1]; x++) {
				const int i = rowList[x];
				const int l = hplLmG[i];
				if (l < S.lmBeg || l >= S.lmEnd) continue;
				const int cend = S.hplColPtr[l + 1];
				for (int j = i; j < cend; j++) {
					const int pos = off[S.hplRowInd[j]]++;
					S.prodI[pos] = i - S.hplBase;
					S.prodJ[pos] = j - S.hplBase;
				}
			}
		}
	}

	// 7. symmetric-full BSR: row r = [ (a,r) transposed entries, a<r ascending ] ++ [ (r,b) upper entries ]
	{
		std::vector<int> cntT(numP, 0);
		for (int k = 0; k < S.nblk; k++) if (S.blkRow[k] != S.blkCol[k]) cntT[S.blkCol[k]]++;
		S.fRowPtr.assign(numP + 1, 0);
		for (int r = 0; r < numP; r++) S.fRowPtr[r + 1] = S.fRowPtr[r] + cntT[r] + (S.hscRowPtr[r + 1] - S.hscRowPtr[r]);
		S.nfull = S.fRowPtr[numP];
		S.fColInd.resize(S.nfull); S.u2f.resize(S.nblk); S.u2fT.resize(S.nblk);
		std::vector<int> fillT(numP, 0);
		for (int k = 0; k < S.nblk; k++) {
			const int a = S.blkRow[k], b = S.blkCol[k];
			const int pos = S.fRowPtr[a] + cntT[a] + (k - S.hscRowPtr[a]);
			S.fColInd[pos] = b; S.u2f[k] = pos;
			if (a != b) {
				const int posT = S.fRowPtr[b] + fillT[b]++;
				S.fColInd[posT] = a; S.u2fT[k] = posT;
			} else S.u2fT[k] = pos;
		}
	}
	return true;
}

}  // namespace cuba_b200
