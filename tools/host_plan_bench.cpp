// host_plan_bench.cpp -- times the host side of the PCG setup (cuba_structure.cpp: row partition, coarse partition, coarse block
// lists, k_pcg5 plan), the part of set_problem that runs on the CPU while the GPU waits.  Needs no GPU.
//   python tools/dump_flat_problem.py ba_kitti_00 /tmp/k00_flat.bin
//   g++ -O2 -std=c++17 -I cuda-bundle-adjustment_b200/csrc tools/host_plan_bench.cpp cuda-bundle-adjustment_b200/csrc/cuba_structure.cpp -o /tmp/host_plan_bench
//   /tmp/host_plan_bench /tmp/k00_flat.bin
// ba_kitti_00 (1 321 free poses, 81 293 blocks), this container: partition 1.2 -> 0.30 ms, coarse partition 0.24 -> 0.055 ms,
// coarse lists 0.26 -> 0.22 ms; the engine used to build the partition twice (k_pcg3 and the k_pcg5 plan): 3.0 -> 0.6 ms in all.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cuba_structure.h"

using namespace cuba_b200;

template <typename T>
static std::vector<T> rd(FILE* f, size_t n) { std::vector<T> v(n); if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } return v; }

int main(int argc, char** argv)
{
	if (argc < 2) { fprintf(stderr, "usage: host_plan_bench flat_problem.bin\n"); return 2; }
	FILE* f = fopen(argv[1], "rb");
	if (!f) return 2;
	const auto n = rd<int64_t>(f, 6);            // Pall numP Lall numL E2 E3
	const auto idx2 = rd<int32_t>(f, 2 * n[4]); const auto idx3 = rd<int32_t>(f, 2 * n[5]);
	fclose(f);
	Structure S; const char* err = nullptr;
	if (!build_structure((int)n[0], (int)n[1], (int)n[2], (int)n[3], (int)n[4], idx2.data(), (int)n[5], idx3.data(), 0, 1, 128, S, &err)) { fprintf(stderr, "%s\n", err); return 2; }
	printf("numP %d, %d blocks in the symmetric-full pattern\n", S.numP, S.nfull);
	auto now = [] { return std::chrono::steady_clock::now(); };
	auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
	const int G = S.numP < 148 ? S.numP : 148;
	for (int rep = 0; rep < 5; rep++) {
		const auto t0 = now();
		PcgPartition P; build_pcg_partition(S.numP, S.nfull, S.fRowPtr, S.fColInd, G, P);
		const auto t1 = now();
		CoarsePartition C; build_coarse_partition(S.numP, P, 148, C);
		const auto t2 = now();
		build_coarse_lists(S.numP, S.nfull, S.fRowPtr, S.fColInd, C);
		const auto t3 = now();
		Pcg5Plan plan; build_pcg5_plan(S.numP, S.nfull, S.fRowPtr, S.fColInd, 1, 148, 148, 85, plan, &P);
		const auto t4 = now();
		const char* bad = plan.ok ? check_pcg5_plan(S.numP, S.nfull, S.fRowPtr, S.fColInd, plan) : nullptr;
		printf("partition %.3f ms  coarse partition %.3f  coarse lists %.3f | k_pcg5 plan from that partition %.3f ms (ok %d%s%s)\n",
			ms(t0, t1), ms(t1, t2), ms(t2, t3), ms(t3, t4), (int)plan.ok, bad ? ", CHECK FAILED: " : "", bad ? bad : "");
	}
	return 0;
}
