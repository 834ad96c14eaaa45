// host_initialize_bench.cpp -- times cuba::CudaBundleAdjustment::initialize() alone (graph objects -> flat arrays).
// initialize() needs no GPU, so the host side of the drop-in class can be tuned on any machine:
//   g++ -std=c++17 -O2 -DCUBA_FORCE_EIGEN_COMPAT -I include -I samples tools/host_initialize_bench.cpp \
//       -L cuda-bundle-adjustment_b200 -lcuba_b200 -Wl,-rpath,$PWD/cuda-bundle-adjustment_b200 -o /tmp/host_initialize_bench
//   /tmp/host_initialize_bench oracle/_ref/fixtures/ba_kitti_00.cubagraph 20
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cubagraph_reader.h"

int main(int argc, char** argv)
{
	if (argc < 2) { fprintf(stderr, "usage: %s graph.cubagraph [repeats]\n", argv[0]); return 2; }
	const int reps = argc > 2 ? atoi(argv[2]) : 20;
	Storage st;
	auto opt = readGraph(argv[1], st);
	std::vector<double> ms;
	for (int i = 0; i < reps; i++) {
		const auto t0 = std::chrono::steady_clock::now();
		opt->initialize();
		ms.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
	}
	std::sort(ms.begin(), ms.end());
	printf("{\"poses\": %zu, \"landmarks\": %zu, \"edges\": %zu, \"initialize_ms_median\": %.3f, \"initialize_ms_min\": %.3f, \"repeats\": %d}\n",
		opt->nposes(), opt->nlandmarks(), opt->nedges(), ms[ms.size() / 2], ms[0], reps);
	return 0;
}
