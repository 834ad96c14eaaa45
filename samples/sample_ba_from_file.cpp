// sample_ba_from_file.cpp -- the reference's benchmark protocol on the drop-in API.
//
// Mirrors samples/sample_ba_from_file.cpp of the reference (graph -> warm-up initialize()+optimize(1)
// -> timed initialize()+optimize(10) -> time profile + chi2 per iteration, :34-89,159-161) but reads the
// flat .cubagraph format (see cuda-bundle-adjustment_b200/graphio.py) instead of OpenCV's JSON reader,
// and can print one JSON object for the tests (--json).  Built by tests/test_cpp_api.py and by hand:
//   g++ -std=c++17 -O2 -Iinclude samples/sample_ba_from_file.cpp -Lcuda-bundle-adjustment_b200 -lcuba_b200 \
//       -Wl,-rpath,$PWD/cuda-bundle-adjustment_b200 -o sample_ba_from_file
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include <cuda_bundle_adjustment.h>

#include "cubagraph_reader.h"

int main(int argc, char** argv)
{
	if (argc < 2) { printf("Usage: sample_ba_from_file input.cubagraph [--json] [--huber] [--iters N] [--no-warmup] [--repeat K] [--dump state.bin]\n"); return 0; }
	bool json = false, huber = false, warmup = true;
	int iters = 10, repeat = 1;
	const char* dump = nullptr;
	for (int i = 2; i < argc; i++) {
		if (!strcmp(argv[i], "--json")) json = true;
		else if (!strcmp(argv[i], "--huber")) huber = true;
		else if (!strcmp(argv[i], "--no-warmup")) warmup = false;
		else if (!strcmp(argv[i], "--iters") && i + 1 < argc) iters = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--repeat") && i + 1 < argc) repeat = atoi(argv[++i]);     // timed windows (bench.py's e2e_cpp leg)
		else if (!strcmp(argv[i], "--dump") && i + 1 < argc) dump = argv[++i];               // final q,t,Xw in file order (comparison report)
	}
	Storage st;
	auto optimizer = readGraph(argv[1], st);
	if (huber) {  // the g2o comparison sample's kernels (reference samples/sample_comparison_with_g2o.cpp:195-200)
		optimizer->setRobustKernels(cuba::RobustKernelType::HUBER, std::sqrt(5.991), cuba::EdgeType::MONOCULAR);
		optimizer->setRobustKernels(cuba::RobustKernelType::HUBER, std::sqrt(7.815), cuba::EdgeType::STEREO);
	}
	if (warmup) { optimizer->initialize(); optimizer->optimize(1); }   // writes its result back, like the reference

	// the timed window of the reference: initialize() + optimize(n) (samples/sample_ba_from_file.cpp:52-57).  With --repeat K the
	// window is measured K times on the SAME input: the estimate the window starts from is restored (untimed) in between.
	std::vector<double> q0, t0v, X0, secs;
	if (repeat > 1) {
		for (auto& v : st.poses) { for (int k = 0; k < 4; k++) q0.push_back(v->q.coeffs().data()[k]); for (int k = 0; k < 3; k++) t0v.push_back(v->t.data()[k]); }
		for (auto& v : st.landmarks) for (int k = 0; k < 3; k++) X0.push_back(v->Xw.data()[k]);
	}
	double sec = 0;
	for (int rep = 0; rep < repeat; rep++) {
		if (rep > 0) {
			for (size_t i = 0; i < st.poses.size(); i++) { for (int k = 0; k < 4; k++) st.poses[i]->q.coeffs().data()[k] = q0[4 * i + k]; for (int k = 0; k < 3; k++) st.poses[i]->t.data()[k] = t0v[3 * i + k]; }
			for (size_t i = 0; i < st.landmarks.size(); i++) for (int k = 0; k < 3; k++) st.landmarks[i]->Xw.data()[k] = X0[3 * i + k];
		}
		const auto t0 = std::chrono::steady_clock::now();
		optimizer->initialize();
		optimizer->optimize(iters);
		sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		secs.push_back(sec);
	}
	if (dump) {
		FILE* f = fopen(dump, "wb");
		if (!f) { fprintf(stderr, "cannot write %s\n", dump); return 2; }
		for (auto& v : st.poses) fwrite(v->q.coeffs().data(), sizeof(double), 4, f);
		for (auto& v : st.poses) fwrite(v->t.data(), sizeof(double), 3, f);
		for (auto& v : st.landmarks) fwrite(v->Xw.data(), sizeof(double), 3, f);
		fclose(f);
	}

	if (json) {
		printf("{\"nposes\": %zu, \"nlandmarks\": %zu, \"nedges\": %zu, \"seconds\": %.6f, \"seconds_all\": [",
			optimizer->nposes(), optimizer->nlandmarks(), optimizer->nedges(), sec);
		for (size_t i = 0; i < secs.size(); i++) printf("%s%.6f", i ? ", " : "", secs[i]);
		printf("], \"chi2\": [");
		const auto& s = optimizer->batchStatistics();
		for (size_t i = 0; i < s.size(); i++) printf("%s%.17g", i ? ", " : "", s[i].chi2);
		double chiSum = 0;
		for (auto& e : st.mono) chiSum += optimizer->chiSquared(e.get());
		for (auto& e : st.stereo) chiSum += optimizer->chiSquared(e.get());
		printf("], \"sum_edge_chi2\": %.17g, \"t_last\": [%.17g, %.17g, %.17g], \"profile\": {", chiSum,
			st.poses.back()->t.data()[0], st.poses.back()->t.data()[1], st.poses.back()->t.data()[2]);
		bool first = true;
		for (const auto& kv : optimizer->timeProfile()) { printf("%s\"%s\": %.6f", first ? "" : ", ", kv.first.c_str(), kv.second); first = false; }
		printf("}}\n");
		return 0;
	}
	printf("=== Graph size : \nnum poses      : %zu\nnum landmarks  : %zu\nnum edges      : %zu\n\n", optimizer->nposes(), optimizer->nlandmarks(), optimizer->nedges());
	printf("=== Processing time : \nBA total : %.4f[sec]\n\n", sec);
	for (const auto& kv : optimizer->timeProfile()) printf("%-30s : %8.2f[msec]\n", kv.first.c_str(), 1e3 * kv.second);
	printf("\n=== Objective function value : \n");
	for (const auto& s : optimizer->batchStatistics()) printf("iter: %2d, chi2: %.1f\n", s.iteration + 1, s.chi2);
	return 0;
}
